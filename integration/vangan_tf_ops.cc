// TensorFlow custom-op shim over the C ABI of include/vangan_b200.h — the binding a maintainer of the reference
// (psweens/VAN-GAN, TensorFlow 2.10) adds to reach the B200 kernels from inside tf.function graphs.
//
// NOT COMPILED IN THIS IMAGE (no TensorFlow headers here); build where TF is installed:
//   g++ -std=c++17 -shared -fPIC integration/vangan_tf_ops.cc -o vangan_tf_ops.so -I include \
//       $(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags() + tf.sysconfig.get_link_flags()))') \
//       -L van-gan_b200 -lvangan_b200 -DGOOGLE_CUDA=1
// Every kernel only ENQUEUES on the op's compute stream (no host synchronisation), allocates outputs and scratch through the
// OpKernelContext, and maps a non-zero ABI status to errors::Internal.  Tensor layouts are Keras': activations NDHWC,
// Conv3D kernels (kd,kh,kw,Cin,Cout) fp32, bias (Cout), gamma/beta (C).  Python side: integration/vangan_tf.py.
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "vangan_b200.h"

using namespace tensorflow;

namespace {

inline void* StreamOf(OpKernelContext* ctx) { return (void*)ctx->eigen_gpu_device().stream(); }
inline int DtypeOf(const Tensor& t) { return t.dtype() == DT_BFLOAT16 ? VG_BF16 : VG_F32; }
inline const void* Ptr(const Tensor& t) { return t.tensor_data().data(); }
inline void* MutPtr(Tensor* t) { return const_cast<char*>(t->tensor_data().data()); }

#define VG_TF_CHECK(ctx, call)                                                                   \
    do {                                                                                         \
        int rc__ = (call);                                                                       \
        OP_REQUIRES(ctx, rc__ == VG_OK, errors::Internal(#call " failed with status ", rc__));   \
    } while (0)

// ------------------------------------------------------------------------------------------ Conv3D (resunet_model.py:64, discriminator.py:63, building_blocks.py:182)
struct ConvAttrs {
    int k, stride, cout, act, dx_lo, dx_hi;
    explicit ConvAttrs(OpKernelConstruction* c) {
        OP_REQUIRES_OK(c, c->GetAttr("k", &k));
        OP_REQUIRES_OK(c, c->GetAttr("stride", &stride));
        OP_REQUIRES_OK(c, c->GetAttr("cout", &cout));
        OP_REQUIRES_OK(c, c->GetAttr("act", &act));
        OP_REQUIRES_OK(c, c->GetAttr("dx_lo", &dx_lo));
        OP_REQUIRES_OK(c, c->GetAttr("dx_hi", &dx_hi));
    }
    vg_conv3d_desc Desc(const TensorShape& x, int x_dtype) const {
        const int cin = (int)x.dim_size(4);
        return vg_conv3d_desc{(int)x.dim_size(0), (int)x.dim_size(1), (int)x.dim_size(2), (int)x.dim_size(3), cin, cout, k, stride,
                              x_dtype, cout == 1 ? VG_F32 : VG_BF16, act, dx_lo, dx_hi};
    }
};

// w_keras (fp32) -> packed bf16 operand copies; run once after every optimizer step (the Keras layer caches the result)
class VgConv3dPackOp : public OpKernel {
 public:
    explicit VgConv3dPackOp(OpKernelConstruction* c) : OpKernel(c), a_(c) { OP_REQUIRES_OK(c, c->GetAttr("cin", &cin_)); }
    void Compute(OpKernelContext* ctx) override {
        const Tensor& w = ctx->input(0);
        vg_conv3d_desc d{1, 8, 8, 8, cin_, a_.cout, a_.k, a_.stride, cin_ == 1 ? VG_F32 : VG_BF16, a_.cout == 1 ? VG_F32 : VG_BF16, a_.act, 0, 0};
        Tensor *wf = nullptr, *wd = nullptr;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, {(int64_t)std::max<size_t>(vg_conv3d_packed_bytes(&d, 0), 2) / 2}, &wf));
        OP_REQUIRES_OK(ctx, ctx->allocate_output(1, {(int64_t)std::max<size_t>(vg_conv3d_packed_bytes(&d, 1), 2) / 2}, &wd));
        VG_TF_CHECK(ctx, vg_conv3d_pack_weights(&d, w.flat<float>().data(), MutPtr(wf), MutPtr(wd), StreamOf(ctx)));
    }
 private:
    ConvAttrs a_;
    int cin_;
};

class VgConv3dFwdOp : public OpKernel {
 public:
    explicit VgConv3dFwdOp(OpKernelConstruction* c) : OpKernel(c), a_(c) {}
    void Compute(OpKernelContext* ctx) override {
        const Tensor& x = ctx->input(0);   // NDHWC, explicitly padded (VgInstanceNorm / VgPadNoise write the padding)
        vg_conv3d_desc d = a_.Desc(x.shape(), DtypeOf(x));
        Tensor* y = nullptr;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, {d.N, (d.ID - d.K) / d.stride + 1, (d.IH - d.K) / d.stride + 1,
                                                    (d.IW - d.K) / d.stride + 1, d.Cout}, &y));
        VG_TF_CHECK(ctx, vg_conv3d_fwd(&d, Ptr(x), Ptr(ctx->input(1)), ctx->input(2).flat<float>().data(), MutPtr(y), StreamOf(ctx)));
    }
 private:
    ConvAttrs a_;
};

class VgConv3dDgradOp : public OpKernel {   // inputs: dy, w_dgrad (packed), x_shape (int32[5])
 public:
    explicit VgConv3dDgradOp(OpKernelConstruction* c) : OpKernel(c), a_(c) {}
    void Compute(OpKernelContext* ctx) override {
        auto xs = ctx->input(2).flat<int32>();
        TensorShape xshape({xs(0), xs(1), xs(2), xs(3), xs(4)});
        vg_conv3d_desc d = a_.Desc(xshape, xs(4) == 1 ? VG_F32 : VG_BF16);
        Tensor* dx = nullptr;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, xshape, &dx));
        VG_TF_CHECK(ctx, vg_conv3d_dgrad(&d, Ptr(ctx->input(0)), Ptr(ctx->input(1)), MutPtr(dx), StreamOf(ctx)));
    }
 private:
    ConvAttrs a_;
};

class VgConv3dWgradOp : public OpKernel {   // inputs: x, dy -> dw (Keras layout, fp32), dbias
 public:
    explicit VgConv3dWgradOp(OpKernelConstruction* c) : OpKernel(c), a_(c) {}
    void Compute(OpKernelContext* ctx) override {
        const Tensor& x = ctx->input(0);
        vg_conv3d_desc d = a_.Desc(x.shape(), DtypeOf(x));
        Tensor *dw = nullptr, *db = nullptr;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, {d.K, d.K, d.K, d.Cin, d.Cout}, &dw));
        OP_REQUIRES_OK(ctx, ctx->allocate_output(1, {d.Cout}, &db));
        auto stream = ctx->eigen_gpu_device().stream();
        cudaMemsetAsync(MutPtr(dw), 0, dw->TotalBytes(), stream);   // the ABI accumulates (+=)
        cudaMemsetAsync(MutPtr(db), 0, db->TotalBytes(), stream);
        VG_TF_CHECK(ctx, vg_conv3d_wgrad(&d, Ptr(x), Ptr(ctx->input(1)), dw->flat<float>().data(), db->flat<float>().data(), (void*)stream));
    }
 private:
    ConvAttrs a_;
};

// ------------------------------------------------------------------------------------------ InstanceNorm + act + Add + dropout + noise + pad
// (resunet_model.py:23-39,96-100,133-143; discriminator.py:70-72,105-106; building_blocks.py:15-39,166-195)
struct NormAttrs {
    int act, pad_lo, pad_hi, pad_mode;
    float slope, noise_std;
    explicit NormAttrs(OpKernelConstruction* c) {
        OP_REQUIRES_OK(c, c->GetAttr("act", &act));
        OP_REQUIRES_OK(c, c->GetAttr("slope", &slope));
        OP_REQUIRES_OK(c, c->GetAttr("pad_lo", &pad_lo));
        OP_REQUIRES_OK(c, c->GetAttr("pad_hi", &pad_hi));
        OP_REQUIRES_OK(c, c->GetAttr("pad_mode", &pad_mode));
        OP_REQUIRES_OK(c, c->GetAttr("noise_std", &noise_std));
    }
    vg_instnorm_desc Desc(const Tensor& x, unsigned long long seed) const {
        return vg_instnorm_desc{(int)x.dim_size(0), (int)x.dim_size(1), (int)x.dim_size(2), (int)x.dim_size(3), (int)x.dim_size(4),
                                DtypeOf(x), act, slope, pad_lo, pad_hi, pad_mode, noise_std, seed, nullptr};
    }
};

// inputs: x, residual (or scalar placeholder), gamma, beta, drop (N*C or empty), seed (int64 scalar, host) -> y, mean, rstd
class VgInstanceNormOp : public OpKernel {
 public:
    explicit VgInstanceNormOp(OpKernelConstruction* c) : OpKernel(c), a_(c) {}
    void Compute(OpKernelContext* ctx) override {
        const Tensor& x = ctx->input(0);
        const Tensor& res = ctx->input(1);
        const Tensor& drop = ctx->input(4);
        vg_instnorm_desc d = a_.Desc(x, (unsigned long long)ctx->input(5).scalar<int64_t>()());
        const int pp = d.pad_lo + d.pad_hi;
        Tensor *y = nullptr, *mean = nullptr, *rstd = nullptr, ws;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, {d.N, d.D + pp, d.H + pp, d.W + pp, d.C}, &y));
        OP_REQUIRES_OK(ctx, ctx->allocate_output(1, {d.N * d.C}, &mean));
        OP_REQUIRES_OK(ctx, ctx->allocate_output(2, {d.N * d.C}, &rstd));
        const size_t wsb = vg_instnorm_workspace_bytes(d.N, d.D, d.H, d.W, d.C);
        OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_UINT8, {(int64_t)wsb}, &ws));
        void* st = StreamOf(ctx);
        VG_TF_CHECK(ctx, vg_instnorm_stats(Ptr(x), d.dtype, d.N, d.D, d.H, d.W, d.C, mean->flat<float>().data(), rstd->flat<float>().data(),
                                           MutPtr(&ws), wsb, st));
        VG_TF_CHECK(ctx, vg_instnorm_apply(&d, Ptr(x), res.NumElements() > 1 ? Ptr(res) : nullptr, MutPtr(y), mean->flat<float>().data(),
                                           rstd->flat<float>().data(), ctx->input(2).flat<float>().data(), ctx->input(3).flat<float>().data(),
                                           drop.NumElements() ? drop.flat<float>().data() : nullptr, nullptr, st));
    }
 private:
    NormAttrs a_;
};

// inputs: dy, x, mean, rstd, gamma, beta, drop -> dx, dres, dgamma, dbeta
class VgInstanceNormGradOp : public OpKernel {
 public:
    explicit VgInstanceNormGradOp(OpKernelConstruction* c) : OpKernel(c), a_(c) {}
    void Compute(OpKernelContext* ctx) override {
        const Tensor& x = ctx->input(1);
        const Tensor& drop = ctx->input(6);
        vg_instnorm_desc d = a_.Desc(x, 0);
        Tensor *dx = nullptr, *dres = nullptr, *dg = nullptr, *db = nullptr, ws;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, x.shape(), &dx));
        OP_REQUIRES_OK(ctx, ctx->allocate_output(1, x.shape(), &dres));
        OP_REQUIRES_OK(ctx, ctx->allocate_output(2, {d.C}, &dg));
        OP_REQUIRES_OK(ctx, ctx->allocate_output(3, {d.C}, &db));
        const size_t wsb = vg_instnorm_workspace_bytes(d.N, d.D, d.H, d.W, d.C);
        OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_UINT8, {(int64_t)wsb}, &ws));
        auto stream = ctx->eigen_gpu_device().stream();
        cudaMemsetAsync(MutPtr(dg), 0, dg->TotalBytes(), stream);
        cudaMemsetAsync(MutPtr(db), 0, db->TotalBytes(), stream);
        VG_TF_CHECK(ctx, vg_instnorm_bwd(&d, Ptr(ctx->input(0)), Ptr(x), ctx->input(2).flat<float>().data(), ctx->input(3).flat<float>().data(),
                                         ctx->input(4).flat<float>().data(), ctx->input(5).flat<float>().data(),
                                         drop.NumElements() ? drop.flat<float>().data() : nullptr, MutPtr(dx), 0, MutPtr(dres),
                                         dg->flat<float>().data(), db->flat<float>().data(), MutPtr(&ws), wsb, (void*)stream));
    }
 private:
    NormAttrs a_;
};

// ------------------------------------------------------------------------------------------ clDice soft skeleton (clDice_func.py:60-80)
// forward keeps the erosion pyramid E and the skeleton history S for the backward op
class VgSoftSkelOp : public OpKernel {
 public:
    explicit VgSoftSkelOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("iters", &iters_)); }
    void Compute(OpKernelContext* ctx) override {
        const Tensor& x = ctx->input(0);   // [N, D, H, W, 1] fp32
        const int N = (int)x.dim_size(0), D = (int)x.dim_size(1), H = (int)x.dim_size(2), W = (int)x.dim_size(3);
        const int64_t nv = x.NumElements();
        Tensor *skel = nullptr, *E = nullptr, *S = nullptr;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(1, {iters_ + 2, nv}, &E));
        OP_REQUIRES_OK(ctx, ctx->allocate_output(2, {iters_ + 1, nv}, &S));
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, x.shape(), &skel));
        auto stream = ctx->eigen_gpu_device().stream();
        VG_TF_CHECK(ctx, vg_soft_skel_fwd(x.flat<float>().data(), E->flat<float>().data(), S->flat<float>().data(), N, D, H, W, iters_,
                                          (void*)stream));
        cudaMemcpyAsync(MutPtr(skel), S->flat<float>().data() + (size_t)iters_ * nv, nv * sizeof(float), cudaMemcpyDeviceToDevice, stream);
    }
 private:
    int iters_;
};

class VgSoftSkelGradOp : public OpKernel {   // inputs: gskel, E, S -> dx
 public:
    explicit VgSoftSkelGradOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("iters", &iters_)); }
    void Compute(OpKernelContext* ctx) override {
        const Tensor& g = ctx->input(0);
        const int N = (int)g.dim_size(0), D = (int)g.dim_size(1), H = (int)g.dim_size(2), W = (int)g.dim_size(3);
        Tensor* dx = nullptr;
        Tensor ws;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, g.shape(), &dx));
        const size_t wsb = vg_soft_skel_bwd_workspace_bytes(N, D, H, W);
        OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_UINT8, {(int64_t)wsb}, &ws));
        VG_TF_CHECK(ctx, vg_soft_skel_bwd(ctx->input(1).flat<float>().data(), ctx->input(2).flat<float>().data(), g.flat<float>().data(),
                                          dx->flat<float>().data(), MutPtr(&ws), wsb, N, D, H, W, iters_, StreamOf(ctx)));
    }
 private:
    int iters_;
};

}  // namespace

#define VG_CONV_ATTRS ".Attr(\"k: int\").Attr(\"stride: int\").Attr(\"cout: int\").Attr(\"act: int = 0\").Attr(\"dx_lo: int = 0\").Attr(\"dx_hi: int = 0\")"
REGISTER_OP("VgConv3dPack").Input("w: float").Output("w_fwd: bfloat16").Output("w_dgrad: bfloat16")
    .Attr("cin: int").Attr("k: int").Attr("stride: int").Attr("cout: int").Attr("act: int = 0").Attr("dx_lo: int = 0").Attr("dx_hi: int = 0");
REGISTER_OP("VgConv3dFwd").Input("x: T").Input("w_fwd: bfloat16").Input("bias: float").Output("y: bfloat16")
    .Attr("T: {bfloat16, float}").Attr("k: int").Attr("stride: int").Attr("cout: int").Attr("act: int = 0").Attr("dx_lo: int = 0").Attr("dx_hi: int = 0");
REGISTER_OP("VgConv3dDgrad").Input("dy: bfloat16").Input("w_dgrad: bfloat16").Input("x_shape: int32").Output("dx: bfloat16")
    .Attr("k: int").Attr("stride: int").Attr("cout: int").Attr("act: int = 0").Attr("dx_lo: int = 0").Attr("dx_hi: int = 0");
REGISTER_OP("VgConv3dWgrad").Input("x: T").Input("dy: bfloat16").Output("dw: float").Output("dbias: float")
    .Attr("T: {bfloat16, float}").Attr("k: int").Attr("stride: int").Attr("cout: int").Attr("act: int = 0").Attr("dx_lo: int = 0").Attr("dx_hi: int = 0");
REGISTER_OP("VgInstanceNorm").Input("x: T").Input("residual: T").Input("gamma: float").Input("beta: float").Input("drop: float").Input("seed: int64")
    .Output("y: T").Output("mean: float").Output("rstd: float")
    .Attr("T: {bfloat16, float}").Attr("act: int = 0").Attr("slope: float = 0.2").Attr("pad_lo: int = 0").Attr("pad_hi: int = 0")
    .Attr("pad_mode: int = 0").Attr("noise_std: float = 0.0");
REGISTER_OP("VgInstanceNormGrad").Input("dy: T").Input("x: T").Input("mean: float").Input("rstd: float").Input("gamma: float").Input("beta: float")
    .Input("drop: float").Output("dx: T").Output("dres: T").Output("dgamma: float").Output("dbeta: float")
    .Attr("T: {bfloat16, float}").Attr("act: int = 0").Attr("slope: float = 0.2").Attr("pad_lo: int = 0").Attr("pad_hi: int = 0")
    .Attr("pad_mode: int = 0").Attr("noise_std: float = 0.0");
REGISTER_OP("VgSoftSkel").Input("x: float").Output("skel: float").Output("e: float").Output("s: float").Attr("iters: int");
REGISTER_OP("VgSoftSkelGrad").Input("gskel: float").Input("e: float").Input("s: float").Output("dx: float").Attr("iters: int");

REGISTER_KERNEL_BUILDER(Name("VgConv3dPack").Device(DEVICE_GPU), VgConv3dPackOp);
REGISTER_KERNEL_BUILDER(Name("VgConv3dFwd").Device(DEVICE_GPU), VgConv3dFwdOp);
REGISTER_KERNEL_BUILDER(Name("VgConv3dDgrad").Device(DEVICE_GPU).HostMemory("x_shape"), VgConv3dDgradOp);
REGISTER_KERNEL_BUILDER(Name("VgConv3dWgrad").Device(DEVICE_GPU), VgConv3dWgradOp);
REGISTER_KERNEL_BUILDER(Name("VgInstanceNorm").Device(DEVICE_GPU).HostMemory("seed"), VgInstanceNormOp);
REGISTER_KERNEL_BUILDER(Name("VgInstanceNormGrad").Device(DEVICE_GPU), VgInstanceNormGradOp);
REGISTER_KERNEL_BUILDER(Name("VgSoftSkel").Device(DEVICE_GPU), VgSoftSkelOp);
REGISTER_KERNEL_BUILDER(Name("VgSoftSkelGrad").Device(DEVICE_GPU), VgSoftSkelGradOp);
