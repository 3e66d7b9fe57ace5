// Layout / pooling kernels of the V-Net generator variant (vnet_model.py:80-146,199-264): every tensor that reaches a
// convolution is produced ALREADY PADDED by the pass that has to touch it anyway, so ReflectionPadding3D
// (building_blocks.py:15-39), UpSampling3D, concatenate, TF-'same' zero padding and MaxPooling3D never cost a pass of
// their own beyond one read + one write.
//
//   vg_gather_pad      out = pad_p( concat( upsample_up(a), b ) )                 (vnet_model.py:247-252 and the pad at :116,:132)
//   vg_gather_pad_bwd  da  = blocksum_up( fold_p(dout[..., :C0]) ),  db = fold_p(dout[..., C0:])
//   vg_maxpool2_pad    y   = pad_p( MaxPooling3D(2)(x) )                          (vnet_model.py:223)
//   vg_maxpool2_pad_bwd dx = route( fold_p(dy) ) to the first maximum of each 2x2x2 window (scan order d,h,w)
//
// fold_p = adjoint of the padding: REFLECT adds every halo position whose mirror is the voxel (up to 8 terms on the
// shell, 1 in the interior); ZERO keeps the interior only.  bf16 NDHWC, 8 channels (one 128-bit access) per thread.
#include "common.cuh"

namespace {

constexpr int NT = 256;

__device__ __forceinline__ void add8(float* a, const bf16* p) {
    float t[8];
    load8<bf16>(p, t);
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] += t[k];
}

// g[8] = fold of dout (padded dims PD,PH,PW; C channels per voxel, channel offset c0) at unpadded voxel (d,h,w)
__device__ __forceinline__ void fold_gather(const bf16* __restrict__ dn, int D, int H, int W, int pad, int mode, int C, int c0, int d,
                                            int h, int w, float* g) {
    const int PH = H + 2 * pad, PW = W + 2 * pad;
#pragma unroll
    for (int k = 0; k < 8; k++) g[k] = 0.f;
    add8(g, dn + ((size_t)((d + pad) * PH + h + pad) * PW + w + pad) * C + c0);
    if (pad == 1 && mode == VG_PAD_REFLECT && (d == 1 || d == D - 2 || h == 1 || h == H - 2 || w == 1 || w == W - 2)) {
        int dd[3], hh[3], ww[3], nd = 1, nh = 1, nw = 1;
        dd[0] = d + 1; hh[0] = h + 1; ww[0] = w + 1;
        if (d == 1) dd[nd++] = 0;
        if (d == D - 2) dd[nd++] = D + 1;
        if (h == 1) hh[nh++] = 0;
        if (h == H - 2) hh[nh++] = H + 1;
        if (w == 1) ww[nw++] = 0;
        if (w == W - 2) ww[nw++] = W + 1;
        for (int i0 = 0; i0 < nd; i0++)
            for (int i1 = 0; i1 < nh; i1++)
                for (int i2 = 0; i2 < nw; i2++) {
                    if (i0 + i1 + i2 == 0) continue;
                    add8(g, dn + ((size_t)(dd[i0] * PH + hh[i1]) * PW + ww[i2]) * C + c0);
                }
    }
}

// out[N, D+2p, H+2p, W+2p, C0+C1]; a: [N, D/up, H/up, W/up, C0]; b: [N, D, H, W, C1]
__global__ void __launch_bounds__(NT) gather_pad_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out,
                                                        int N, int D, int H, int W, int C0, int C1, int up, int pad, int mode) {
    const int C = C0 + C1, cg = C / 8, cg0 = C0 / 8;
    const int PD = D + 2 * pad, PH = H + 2 * pad, PW = W + 2 * pad;
    const int AD = D / up, AH = H / up, AW = W / up;
    const size_t total = (size_t)N * PD * PH * PW * cg;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        const int c8 = (int)(i % cg);
        size_t v = i / cg;
        const int pw = (int)(v % PW), ph = (int)((v / PW) % PH), pd = (int)((v / ((size_t)PW * PH)) % PD);
        const int n = (int)(v / ((size_t)PW * PH * PD));
        int d = pd - pad, h = ph - pad, w = pw - pad;
        const bool oob = (unsigned)d >= (unsigned)D || (unsigned)h >= (unsigned)H || (unsigned)w >= (unsigned)W;
        bf16x8 val = make_uint4(0u, 0u, 0u, 0u);
        if (!(oob && mode == VG_PAD_ZERO)) {
            if (oob) { d = reflect1(d, D); h = reflect1(h, H); w = reflect1(w, W); }
            if (c8 < cg0)
                val = *reinterpret_cast<const bf16x8*>(a + ((((size_t)n * AD + d / up) * AH + h / up) * AW + w / up) * C0 + c8 * 8);
            else
                val = *reinterpret_cast<const bf16x8*>(b + ((((size_t)n * D + d) * H + h) * W + w) * C1 + (c8 - cg0) * 8);
        }
        *reinterpret_cast<bf16x8*>(out + i * 8) = val;
    }
}

// da[N, D/up, H/up, W/up, C0] = sum over the up^3 block of fold(dout[..., :C0])
__global__ void __launch_bounds__(NT) gather_pad_bwd_a_kernel(const bf16* __restrict__ dout, bf16* __restrict__ da, int N, int D, int H,
                                                              int W, int C0, int C1, int up, int pad, int mode) {
    const int C = C0 + C1, cg0 = C0 / 8;
    const int PD = D + 2 * pad, PH = H + 2 * pad, PW = W + 2 * pad;
    const int AD = D / up, AH = H / up, AW = W / up;
    const size_t total = (size_t)N * AD * AH * AW * cg0;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        const int c8 = (int)(i % cg0);
        size_t v = i / cg0;
        const int w = (int)(v % AW), h = (int)((v / AW) % AH), d = (int)((v / ((size_t)AW * AH)) % AD);
        const int n = (int)(v / ((size_t)AW * AH * AD));
        const bf16* dn = dout + (size_t)n * PD * PH * PW * C;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = 0.f;
        for (int a0 = 0; a0 < up; a0++)
            for (int a1 = 0; a1 < up; a1++)
                for (int a2 = 0; a2 < up; a2++) {
                    float g[8];
                    fold_gather(dn, D, H, W, pad, mode, C, c8 * 8, d * up + a0, h * up + a1, w * up + a2, g);
#pragma unroll
                    for (int k = 0; k < 8; k++) acc[k] += g[k];
                }
        store8<bf16>(da + i * 8, acc);
    }
}

// db[N, D, H, W, C1] = fold(dout[..., C0:])
__global__ void __launch_bounds__(NT) gather_pad_bwd_b_kernel(const bf16* __restrict__ dout, bf16* __restrict__ db, int N, int D, int H,
                                                              int W, int C0, int C1, int pad, int mode) {
    const int C = C0 + C1, cg1 = C1 / 8;
    const int PD = D + 2 * pad, PH = H + 2 * pad, PW = W + 2 * pad;
    const size_t total = (size_t)N * D * H * W * cg1;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        const int c8 = (int)(i % cg1);
        size_t v = i / cg1;
        const int w = (int)(v % W), h = (int)((v / W) % H), d = (int)((v / ((size_t)W * H)) % D);
        const int n = (int)(v / ((size_t)W * H * D));
        float g[8];
        fold_gather(dout + (size_t)n * PD * PH * PW * C, D, H, W, pad, mode, C, C0 + c8 * 8, d, h, w, g);
        store8<bf16>(db + i * 8, g);
    }
}

// y[N, D/2+2p, H/2+2p, W/2+2p, C] = pad(maxpool2(x[N,D,H,W,C]))
__global__ void __launch_bounds__(NT) maxpool2_pad_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int N, int D, int H, int W,
                                                          int C, int pad, int mode) {
    const int cg = C / 8, OD = D / 2, OH = H / 2, OW = W / 2;
    const int PD = OD + 2 * pad, PH = OH + 2 * pad, PW = OW + 2 * pad;
    const size_t total = (size_t)N * PD * PH * PW * cg;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        const int c8 = (int)(i % cg);
        size_t v = i / cg;
        const int pw = (int)(v % PW), ph = (int)((v / PW) % PH), pd = (int)((v / ((size_t)PW * PH)) % PD);
        const int n = (int)(v / ((size_t)PW * PH * PD));
        int d = pd - pad, h = ph - pad, w = pw - pad;
        const bool oob = (unsigned)d >= (unsigned)OD || (unsigned)h >= (unsigned)OH || (unsigned)w >= (unsigned)OW;
        float m[8];
#pragma unroll
        for (int k = 0; k < 8; k++) m[k] = 0.f;
        if (!(oob && mode == VG_PAD_ZERO)) {
            if (oob) { d = reflect1(d, OD); h = reflect1(h, OH); w = reflect1(w, OW); }
#pragma unroll
            for (int k = 0; k < 8; k++) m[k] = -INFINITY;
            for (int a0 = 0; a0 < 2; a0++)
                for (int a1 = 0; a1 < 2; a1++)
                    for (int a2 = 0; a2 < 2; a2++) {
                        float t[8];
                        load8<bf16>(x + ((((size_t)n * D + 2 * d + a0) * H + 2 * h + a1) * W + 2 * w + a2) * C + c8 * 8, t);
#pragma unroll
                        for (int k = 0; k < 8; k++) m[k] = fmaxf(m[k], t[k]);
                    }
        }
        store8<bf16>(y + i * 8, m);
    }
}

// dx[N,D,H,W,C]: the folded gradient of each pooled voxel goes to the FIRST maximum of its window, zero elsewhere
__global__ void __launch_bounds__(NT) maxpool2_pad_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                              bf16* __restrict__ dx, int N, int D, int H, int W, int C, int pad,
                                                              int mode) {
    const int cg = C / 8, OD = D / 2, OH = H / 2, OW = W / 2;
    const int PD = OD + 2 * pad, PH = OH + 2 * pad, PW = OW + 2 * pad;
    const size_t total = (size_t)N * OD * OH * OW * cg;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        const int c8 = (int)(i % cg);
        size_t v = i / cg;
        const int w = (int)(v % OW), h = (int)((v / OW) % OH), d = (int)((v / ((size_t)OW * OH)) % OD);
        const int n = (int)(v / ((size_t)OW * OH * OD));
        float g[8];
        fold_gather(dy + (size_t)n * PD * PH * PW * C, OD, OH, OW, pad, mode, C, c8 * 8, d, h, w, g);
        float t[8][8], m[8];
        int arg[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { m[k] = -INFINITY; arg[k] = 0; }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            load8<bf16>(x + ((((size_t)n * D + 2 * d + (q >> 2)) * H + 2 * h + ((q >> 1) & 1)) * W + 2 * w + (q & 1)) * C + c8 * 8, t[q]);
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (t[q][k] > m[k]) { m[k] = t[q][k]; arg[k] = q; }
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; k++) o[k] = arg[k] == q ? g[k] : 0.f;
            store8<bf16>(dx + ((((size_t)n * D + 2 * d + (q >> 2)) * H + 2 * h + ((q >> 1) & 1)) * W + 2 * w + (q & 1)) * C + c8 * 8, o);
        }
    }
}

inline bool pad_ok(int D, int H, int W, int pad, int mode) {
    if (pad != 0 && pad != 1) return false;
    if (mode != VG_PAD_ZERO && mode != VG_PAD_REFLECT) return false;
    if (pad == 1 && mode == VG_PAD_REFLECT && (D < 2 || H < 2 || W < 2)) return false;
    return true;
}


// Conv3DTranspose(filters, (2,2,2), strides 2, 'same') (vnet_model.py:245): the eight output positions of an input voxel do not
// overlap, so the layer is ONE pointwise GEMM with 8*Cout columns ordered (a, b, c, co) -- run by vg_conv3d_fwd / dgrad / wgrad with
// K = 1 on the tensor cores -- followed by a depth-to-space scatter that also adds the bias.  t: [N, D, H, W, 8*Cout] bf16;
// y: [N, 2D, 2H, 2W, Cout] bf16.  One thread moves 8 channels (128 bits).
__global__ void __launch_bounds__(NT) ct_scatter_kernel(const bf16* __restrict__ t, const float* __restrict__ bias, bf16* __restrict__ y, int N,
                                                        int D, int H, int W, int Co) {
    const int cg = Co / 8;
    const size_t total = (size_t)N * 8 * D * H * W * cg;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        const int c8 = (int)(i % cg);
        size_t v = i / cg;
        const int ow = (int)(v % (2 * W)); v /= 2 * W;
        const int oh = (int)(v % (2 * H)); v /= 2 * H;
        const int od = (int)(v % (2 * D));
        const int n = (int)(v / (2 * D));
        const int pos = ((od & 1) * 2 + (oh & 1)) * 2 + (ow & 1);
        float f[8];
        load8<bf16>(t + ((((size_t)n * D + (od >> 1)) * H + (oh >> 1)) * W + (ow >> 1)) * (8 * Co) + pos * Co + c8 * 8, f);
        if (bias) {
#pragma unroll
            for (int k = 0; k < 8; k++) f[k] += bias[c8 * 8 + k];
        }
        store8<bf16>(y + i * 8, f);
    }
}

// backward of the scatter: dt = space-to-depth(dy); dbias[co] += sum of dy over every output voxel (block partials, one atomic per
// channel per block)
__global__ void __launch_bounds__(NT) ct_gather_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dt, float* __restrict__ dbias, int N,
                                                       int D, int H, int W, int Co) {
    extern __shared__ float sb[];   // Co floats
    const int cg = Co / 8;
    for (int i = threadIdx.x; i < Co; i += NT) sb[i] = 0.f;
    __syncthreads();
    const size_t total = (size_t)N * 8 * D * H * W * cg;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int my_c8 = -1;
    // NT * gridDim.x is a multiple of cg (cg divides NT: Co in {16..256}), so a thread always sees the same channel group
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        const int c8 = (int)(i % cg);
        my_c8 = c8;
        size_t v = i / cg;
        const int ow = (int)(v % (2 * W)); v /= 2 * W;
        const int oh = (int)(v % (2 * H)); v /= 2 * H;
        const int od = (int)(v % (2 * D));
        const int n = (int)(v / (2 * D));
        const int pos = ((od & 1) * 2 + (oh & 1)) * 2 + (ow & 1);
        float f[8];
        load8<bf16>(dy + i * 8, f);
        store8<bf16>(dt + ((((size_t)n * D + (od >> 1)) * H + (oh >> 1)) * W + (ow >> 1)) * (8 * Co) + pos * Co + c8 * 8, f);
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] += f[k];
    }
    if (dbias && my_c8 >= 0) {
#pragma unroll
        for (int k = 0; k < 8; k++) atomicAdd(sb + my_c8 * 8 + k, acc[k]);
    }
    __syncthreads();
    if (dbias)
        for (int i = threadIdx.x; i < Co; i += NT) atomicAdd(dbias + i, sb[i]);
}


// Conv3DTranspose kernel layouts: Keras (2,2,2,Cout,Cin) <-> pointwise-GEMM (Cin, 8*Cout) with columns (a,b,c,co).
// dir 0: gemm = permute(keras) (after load / every optimizer step); dir 1: keras_grad += permute(gemm_grad) (after the wgrad).
__global__ void ct_weights_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cin, int Cout, int dir) {
    const int total = 8 * Cin * Cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ci = i % Cin, co = (i / Cin) % Cout, pos = i / (Cin * Cout);        // i indexes the Keras layout
        const int j = ci * (8 * Cout) + pos * Cout + co;                               // the GEMM layout
        if (dir == 0) dst[j] = src[i];
        else atomicAdd(dst + i, src[j]);   // sub-sweeps of one network may fold their kernel gradients concurrently
    }
}


// UpSampling3D(2) followed by the zero padding of a TF-'same' convolution with an even kernel (k4 s1: 1 before, 2 after --
// building_blocks.upsample, building_blocks.py:240-280): out[N, 2D+lo+hi, 2H+lo+hi, 2W+lo+hi, C] in one pass; and its adjoint
// (da = sum of the 2x2x2 block of the interior of dout).
__global__ void __launch_bounds__(NT) upsample_pad_kernel(const bf16* __restrict__ a, bf16* __restrict__ out, int N, int D, int H, int W, int C,
                                                          int lo, int hi) {
    const int cg = C / 8, PD = 2 * D + lo + hi, PH = 2 * H + lo + hi, PW = 2 * W + lo + hi;
    const size_t total = (size_t)N * PD * PH * PW * cg;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        const int c8 = (int)(i % cg);
        size_t v = i / cg;
        const int pw = (int)(v % PW); v /= PW;
        const int ph = (int)(v % PH); v /= PH;
        const int pd = (int)(v % PD);
        const int n = (int)(v / PD);
        const int d = pd - lo, h = ph - lo, w = pw - lo;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if ((unsigned)d < (unsigned)(2 * D) && (unsigned)h < (unsigned)(2 * H) && (unsigned)w < (unsigned)(2 * W))
            val = *reinterpret_cast<const uint4*>(a + ((((size_t)n * D + (d >> 1)) * H + (h >> 1)) * W + (w >> 1)) * C + c8 * 8);
        *reinterpret_cast<uint4*>(out + i * 8) = val;
    }
}

__global__ void __launch_bounds__(NT) upsample_pad_bwd_kernel(const bf16* __restrict__ dout, bf16* __restrict__ da, int N, int D, int H, int W,
                                                              int C, int lo, int hi) {
    const int cg = C / 8, PH = 2 * H + lo + hi, PW = 2 * W + lo + hi, PD = 2 * D + lo + hi;
    const size_t total = (size_t)N * D * H * W * cg;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        const int c8 = (int)(i % cg);
        size_t v = i / cg;
        const int w = (int)(v % W); v /= W;
        const int h = (int)(v % H); v /= H;
        const int d = (int)(v % D);
        const int n = (int)(v / D);
        float g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int a = 0; a < 2; a++)
            for (int b = 0; b < 2; b++)
                for (int c = 0; c < 2; c++)
                    add8(g, dout + ((((size_t)n * PD + 2 * d + a + lo) * PH + 2 * h + b + lo) * PW + 2 * w + c + lo) * C + c8 * 8);
        store8<bf16>(da + i * 8, g);
    }
}

}  // namespace

extern "C" {

int vg_gather_pad(const void* a, const void* b, void* out, int N, int D, int H, int W, int C0, int C1, int up, int pad, int pad_mode,
                  void* stream) {
    VG_REQUIRE(out && N > 0 && C0 >= 0 && C1 >= 0 && C0 % 8 == 0 && C1 % 8 == 0 && C0 + C1 > 0 && (up == 1 || up == 2));
    VG_REQUIRE((C0 == 0 || a) && (C1 == 0 || b) && D % up == 0 && H % up == 0 && W % up == 0 && pad_ok(D, H, W, pad, pad_mode));
    const size_t total = (size_t)N * (D + 2 * pad) * (H + 2 * pad) * (W + 2 * pad) * ((C0 + C1) / 8);
    gather_pad_kernel<<<vg_grid_for(total, NT, 16), NT, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)b, (bf16*)out, N, D, H, W,
                                                                                  C0, C1, up, pad, pad_mode); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_gather_pad_bwd(const void* dout, void* da, void* db, int N, int D, int H, int W, int C0, int C1, int up, int pad, int pad_mode,
                      void* stream) {
    VG_REQUIRE(dout && N > 0 && C0 % 8 == 0 && C1 % 8 == 0 && C0 + C1 > 0 && (up == 1 || up == 2) && pad_ok(D, H, W, pad, pad_mode));
    cudaStream_t st = (cudaStream_t)stream;
    if (C0 > 0 && da) {
        const size_t t0 = (size_t)N * (D / up) * (H / up) * (W / up) * (C0 / 8);
        gather_pad_bwd_a_kernel<<<vg_grid_for(t0, NT, 16), NT, 0, st>>>((const bf16*)dout, (bf16*)da, N, D, H, W, C0, C1, up, pad, pad_mode); VG_LAUNCHED(1);
    }
    if (C1 > 0 && db) {
        const size_t t1 = (size_t)N * D * H * W * (C1 / 8);
        gather_pad_bwd_b_kernel<<<vg_grid_for(t1, NT, 16), NT, 0, st>>>((const bf16*)dout, (bf16*)db, N, D, H, W, C0, C1, pad, pad_mode); VG_LAUNCHED(1);
    }
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_maxpool2_pad(const void* x, void* y, int N, int D, int H, int W, int C, int pad, int pad_mode, void* stream) {
    VG_REQUIRE(x && y && N > 0 && C % 8 == 0 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0 && pad_ok(D / 2, H / 2, W / 2, pad, pad_mode));
    const size_t total = (size_t)N * (D / 2 + 2 * pad) * (H / 2 + 2 * pad) * (W / 2 + 2 * pad) * (C / 8);
    maxpool2_pad_kernel<<<vg_grid_for(total, NT, 16), NT, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, N, D, H, W, C, pad, pad_mode); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_maxpool2_pad_bwd(const void* x, const void* dy, void* dx, int N, int D, int H, int W, int C, int pad, int pad_mode, void* stream) {
    VG_REQUIRE(x && dy && dx && N > 0 && C % 8 == 0 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0 && pad_ok(D / 2, H / 2, W / 2, pad, pad_mode));
    const size_t total = (size_t)N * (D / 2) * (H / 2) * (W / 2) * (C / 8);
    maxpool2_pad_bwd_kernel<<<vg_grid_for(total, NT, 16), NT, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)dy, (bf16*)dx, N, D, H, W, C,
                                                                                        pad, pad_mode); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_conv3d_transpose_k2s2_scatter(const void* t, const float* bias, void* y, int N, int D, int H, int W, int Cout, void* stream) {
    VG_REQUIRE(t && y && N > 0 && D > 0 && H > 0 && W > 0 && Cout % 8 == 0 && NT % (Cout / 8) == 0);
    const long long total = (long long)N * 8 * D * H * W * (Cout / 8);
    ct_scatter_kernel<<<vg_grid_for(total, NT, 16), NT, 0, (cudaStream_t)stream>>>((const bf16*)t, bias, (bf16*)y, N, D, H, W, Cout); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_conv3d_transpose_k2s2_gather(const void* dy, void* dt, float* dbias, int N, int D, int H, int W, int Cout, void* stream) {
    VG_REQUIRE(dy && dt && N > 0 && D > 0 && H > 0 && W > 0 && Cout % 8 == 0 && NT % (Cout / 8) == 0);
    const long long total = (long long)N * 8 * D * H * W * (Cout / 8);
    ct_gather_kernel<<<vg_grid_for(total, NT, 8), NT, (size_t)Cout * sizeof(float), (cudaStream_t)stream>>>((const bf16*)dy, (bf16*)dt, dbias, N,
                                                                                                     D, H, W, Cout); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_conv3d_transpose_k2s2_weights(const float* src, float* dst, int Cin, int Cout, int dir, void* stream) {
    VG_REQUIRE(src && dst && Cin > 0 && Cout > 0 && (dir == 0 || dir == 1));
    ct_weights_kernel<<<vg_cdiv(8 * Cin * Cout, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, Cin, Cout, dir); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_upsample_pad(const void* a, void* out, int N, int D, int H, int W, int C, int lo, int hi, void* stream) {
    VG_REQUIRE(a && out && N > 0 && D > 0 && H > 0 && W > 0 && C % 8 == 0 && lo >= 0 && hi >= 0);
    const long long total = (long long)N * (2 * D + lo + hi) * (2 * H + lo + hi) * (2 * W + lo + hi) * (C / 8);
    upsample_pad_kernel<<<vg_grid_for(total, NT, 16), NT, 0, (cudaStream_t)stream>>>((const bf16*)a, (bf16*)out, N, D, H, W, C, lo, hi); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_upsample_pad_bwd(const void* dout, void* da, int N, int D, int H, int W, int C, int lo, int hi, void* stream) {
    VG_REQUIRE(dout && da && N > 0 && D > 0 && H > 0 && W > 0 && C % 8 == 0 && lo >= 0 && hi >= 0);
    const long long total = (long long)N * D * H * W * (C / 8);
    upsample_pad_bwd_kernel<<<vg_grid_for(total, NT, 16), NT, 0, (cudaStream_t)stream>>>((const bf16*)dout, (bf16*)da, N, D, H, W, C, lo, hi); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

}  // extern "C"
