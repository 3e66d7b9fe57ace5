/* vangan_b200 — C ABI of the B200-native VAN-GAN volumetric hot path.
 *
 * Every entry point replaces a group of TensorFlow/Keras/tensorflow_addons op kernels that the
 * reference (psweens/VAN-GAN) reaches from VanGan.train_step (vangan.py:380-440) or from
 * GanMonitor.stitch_subvolumes (custom_callback.py:47-223).  The reference has no FFI of its own
 * (it is pure Python over stock TF ops); these are the symbols a TF custom-op shim
 * (OpKernel::Compute forwarding the op context's stream) or the ctypes binding in
 * van-gan_b200/_lib.py binds.  See INTEGRATION.md.
 *
 * Conventions: plain pointers to DEVICE memory, sizes as ints / size_t, `stream` is a
 * cudaStream_t passed as void*.  The caller owns every buffer (inputs, outputs, workspaces); the
 * library allocates nothing, keeps no global state and never synchronises the host.  Return value:
 * 0 = VG_OK, negative = error (no exceptions, no exit).  Activations are NDHWC (Keras
 * channels_last); Conv3D kernels are (kd,kh,kw,Cin,Cout) fp32 exactly as Keras stores them.
 */
#ifndef VANGAN_B200_H
#define VANGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { VG_OK = 0, VG_ERR_INVALID = -1, VG_ERR_UNSUPPORTED = -2, VG_ERR_WORKSPACE = -3, VG_ERR_CUDA = -4 };
enum { VG_F32 = 0, VG_BF16 = 1 };
/* OR-ed into the dtype of the vg_instnorm_* calls: the tensor being normalised is relu(x) — the producer was a
 * Conv3D(activation='relu') (vnet_model.py:118-126,133-141) whose ReLU is applied on load, and whose gradient mask is
 * applied to dx */
enum {
    VG_IN_RELU_INPUT = 0x100,
    VG_IN_BATCH_STATS = 0x200, /* backward reductions over the whole batch: BatchNormalization */
    VG_IN_DY_SCRATCH = 0x400   /* vg_instnorm_bwd may overwrite dy (it folds the reflected halo into the interior in place once instead
                                * of re-gathering it in both passes); without the flag dy is read-only */
};
enum { VG_ACT_NONE = 0, VG_ACT_RELU = 1, VG_ACT_LEAKY = 2, VG_ACT_TANH = 3 };
enum { VG_PAD_ZERO = 0, VG_PAD_REFLECT = 1 };

int vg_abi_version(void);
/* kernels launched by this library so far (monotonic counter, for bench accounting) */
unsigned long long vg_launch_count(void);
/* of those, launches of the tcgen05/TMEM/TMA convolution kernel */
unsigned long long vg_tc_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Conv3D (valid convolution over an explicitly padded input; the padding itself — ReflectionPadding3D
 * or TF 'same' zeros — is written by vg_instnorm_apply / vg_pad_noise, so it never costs a pass).
 * Replaces keras.layers.Conv3D forward / Conv3DBackpropInputV2 / Conv3DBackpropFilterV2 at
 * resunet_model.py:64-65,89-90,96,127,133-134,245; discriminator.py:63-69,108-114;
 * building_blocks.py:182-189.
 * x: [N, ID, IH, IW, Cin]   y: [N, OD, OH, OW, Cout],  O = (I - K)/stride + 1.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int N;
    int ID, IH, IW; /* input spatial dims, padding included */
    int Cin, Cout;
    int K;          /* cubic kernel: 1, 3 or 4 */
    int stride;     /* 1 or 2 */
    int x_dtype;    /* VG_BF16, or VG_F32 when Cin == 1 */
    int y_dtype;    /* VG_BF16, or VG_F32 when Cout == 1 */
    int act;        /* VG_ACT_NONE or VG_ACT_TANH (forward epilogue) */
    int dx_lo, dx_hi; /* dgrad only: dx is needed for spatial indices [dx_lo, I - dx_hi) of every dimension; the rest
                         (zero 'same' padding, whose gradient nobody reads) may be left unwritten.  0,0 = everything */
} vg_conv3d_desc;

/* bytes of the bf16 operand copies of one layer's weights: forward pack and dgrad pack */
size_t vg_conv3d_packed_bytes(const vg_conv3d_desc* d, int for_dgrad);
/* w_keras fp32 (K,K,K,Cin,Cout) -> packed bf16 operand layouts (call after every optimizer step) */
int vg_conv3d_pack_weights(const vg_conv3d_desc* d, const float* w_keras, void* w_fwd, void* w_dgrad, void* stream);
/* The same refresh for a whole network in ONE launch (after every optimizer step: 4 launches per train step instead of ~400).
 * vg_conv3d_pack_jobs expands one layer into packing jobs (HOST call, no CUDA work; jobs_out: max_jobs * vg_pack_job_bytes() bytes,
 * returns the number of jobs written); the caller concatenates the jobs of all layers, builds the exclusive prefix sum of
 * vg_pack_job_total() over them, uploads both once, and calls vg_pack_run(jobs_dev, prefix_dev, njobs, total) per step. */
int vg_conv3d_pack_jobs(const vg_conv3d_desc* d, const float* w_keras, void* w_fwd, void* w_dgrad, void* jobs_out, int max_jobs);
size_t vg_pack_job_bytes(void);
long long vg_pack_job_total(const void* jobs, int index);
int vg_pack_run(const void* jobs_dev, const long long* prefix_dev, int njobs, long long total, void* stream);
/* y = act(conv(x, w) + bias).  `w_fwd` is the packed buffer, or the fp32 Keras kernel when Cin == 1 */
int vg_conv3d_fwd(const vg_conv3d_desc* d, const void* x, const void* w_fwd, const float* bias, void* y, void* stream);
/* dx[N,ID,IH,IW,Cin] = conv_transpose(dy, w)  (gradient w.r.t. the PADDED input; fully overwritten).
 * `w_dgrad` is the packed buffer, or the fp32 Keras kernel when Cout == 1.  dx dtype = x_dtype. */
int vg_conv3d_dgrad(const vg_conv3d_desc* d, const void* dy, const void* w_dgrad, void* dx, void* stream);
/* dw (fp32, Keras layout) += x^T * dy;  dbias (optional, fp32[Cout]) += sum dy */
int vg_conv3d_wgrad(const vg_conv3d_desc* d, const void* x, const void* dy, float* dw, float* dbias, void* stream);

/* ---------------------------------------------------------------------------------------------
 * InstanceNormalization + activation + residual + dropout + noise + padding (tfa InstanceNormalization,
 * Activation/LeakyReLU, Add, SpatialDropout3D, GaussianNoise, ReflectionPadding3D):
 * resunet_model.py:23-39,96-100,133-143; discriminator.py:50-52,70-72,105-106;
 * building_blocks.py:15-39,166-195.
 *   y[pad(p)] = drop[n,c] * act(gamma*(x-mean)*rstd + beta) + residual  (+ noise)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int N, D, H, W, C; /* unpadded input dims */
    int dtype;         /* storage type of x / residual / y / dy / dx */
    int act;           /* VG_ACT_NONE / RELU / LEAKY */
    float slope;       /* LeakyReLU slope */
    int pad_lo, pad_hi, pad_mode; /* padding written around the output (0/0 = none) */
    float noise_std;   /* >0 and noise==NULL: Philox noise generated in-kernel from `seed` */
    unsigned long long seed;
    const unsigned long long* seed_dev; /* optional DEVICE pointer: per-step offset added to `seed` (lets a captured CUDA
                                           graph draw fresh noise on every replay); NULL = none */
} vg_instnorm_desc;

size_t vg_instnorm_workspace_bytes(int N, int D, int H, int W, int C);
int vg_instnorm_stats(const void* x, int dtype, int N, int D, int H, int W, int C, float* mean, float* rstd, void* ws,
                      size_t ws_bytes, void* stream);
int vg_instnorm_apply(const vg_instnorm_desc* d, const void* x, const void* residual, void* y, const float* mean,
                      const float* rstd, const float* gamma, const float* beta, const float* drop, const float* noise,
                      void* stream);
int vg_instnorm_bwd(const vg_instnorm_desc* d, const void* dy, const void* x, const float* mean, const float* rstd,
                    const float* gamma, const float* beta, const float* drop, void* dx, int accumulate_dx, void* dres,
                    float* dgamma, float* dbeta, void* ws, size_t ws_bytes, void* stream);
/* the same + the bias gradients of the convolutions that PRODUCED x and the residual: dx of the norm is dy of that convolution, so
 * dbias_x[c] += sum over (n, voxels) of the stored dx and dbias_res[c] += sum of the stored dres are taken in the same pass instead of a
 * pass of their own over dy inside vg_conv3d_wgrad (call that with dbias = NULL then).  Either may be NULL; dbias_x excludes accumulate_dx. */
int vg_instnorm_bwd_sinks(const vg_instnorm_desc* d, const void* dy, const void* x, const float* mean, const float* rstd,
                          const float* gamma, const float* beta, const float* drop, void* dx, int accumulate_dx, void* dres,
                          float* dgamma, float* dbeta, float* dbias_x, float* dbias_res, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Layout helpers around the convolutions: UpSampling3D(2)+concatenate (resunet_model.py:176,181),
 * input ReflectionPadding3D+GaussianNoise of the discriminator (discriminator.py:50-52), gradient
 * accumulation, tanh backward of the generator head (resunet_model.py:245).
 * ------------------------------------------------------------------------------------------- */
/* out[N,2D,2H,2W,C0+C1] = concat(upsample2(lo[N,D,H,W,C0]), skip[N,2D,2H,2W,C1]) (bf16) */
int vg_upsample_concat(const void* lo, const void* skip, void* out, int N, int D, int H, int W, int C0, int C1, void* stream);
/* dlo = 2x2x2 block-sum of dcat[..., :C0]; dskip (+)= dcat[..., C0:] */
int vg_upsample_concat_bwd(const void* dcat, void* dlo, void* dskip, int accumulate_skip, int N, int D, int H, int W, int C0,
                           int C1, void* stream);
/* y[N,D+2,H+2,W+2] = reflect_pad(x) + noise  (fp32, one channel).  noise may be NULL (then noise_std/seed) */
int vg_pad_noise(const float* x, float* y, int N, int D, int H, int W, const float* noise, float noise_std,
                 unsigned long long seed, const unsigned long long* seed_dev, void* stream);
/* SpatialDropout3D mask (discriminator.py:106, building_blocks.py:195): out[n] = (u >= rate) / (1 - rate), Philox keyed on
 * seed + *seed_dev (seed_dev optional, device) */
int vg_dropout_mask(float* out, int n, float rate, unsigned long long seed, const unsigned long long* seed_dev, void* stream);
/* dx[N,D,H,W] (+)= fold of dy[N,D+2,H+2,W+2] through the reflect padding */
int vg_pad_fold(const float* dy, float* dx, int N, int D, int H, int W, int accumulate, void* stream);
/* a += b  (dtype VG_BF16 or VG_F32) */
int vg_accumulate(void* a, const void* b, size_t n, int dtype, void* stream);
/* out = dy * (1 - y*y) */
int vg_tanh_bwd(const float* dy, const float* y, float* out, size_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * V-Net generator variant (vnet_model.py:80-146,199-264): MaxPooling3D(2), UpSampling3D(2), concatenate,
 * ReflectionPadding3D and TF 'same' zero padding, each fused with the padding of the convolution that consumes it.  bf16.
 * ------------------------------------------------------------------------------------------- */
/* out[N,D+2p,H+2p,W+2p,C0+C1] = pad_p(concat(upsample_up(a[N,D/up,H/up,W/up,C0]), b[N,D,H,W,C1])); up in {1,2}; C0 or C1 may be 0 */
int vg_gather_pad(const void* a, const void* b, void* out, int N, int D, int H, int W, int C0, int C1, int up, int pad, int pad_mode,
                  void* stream);
/* da = up^3 block-sum of fold_p(dout[..., :C0]); db = fold_p(dout[..., C0:]) (either may be NULL) */
int vg_gather_pad_bwd(const void* dout, void* da, void* db, int N, int D, int H, int W, int C0, int C1, int up, int pad, int pad_mode,
                      void* stream);
/* y[N,D/2+2p,H/2+2p,W/2+2p,C] = pad_p(maxpool2(x[N,D,H,W,C])) */
int vg_maxpool2_pad(const void* x, void* y, int N, int D, int H, int W, int C, int pad, int pad_mode, void* stream);
/* dx = fold_p(dy) routed to the first maximum of every 2x2x2 window (scan order d,h,w), zero elsewhere */
int vg_maxpool2_pad_bwd(const void* x, const void* dy, void* dx, int N, int D, int H, int W, int C, int pad, int pad_mode, void* stream);

/* ---------------------------------------------------------------------------------------------
 * clDice soft skeleton (clDice_func.py:8-80).  x: [N,D,H,W] fp32.
 * E: [(iters+2)][N*D*H*W] erosion pyramid, S: [(iters+1)][N*D*H*W]; the skeleton is S[iters].
 * ------------------------------------------------------------------------------------------- */
int vg_soft_skel_fwd(const float* x, float* E, float* S, int N, int D, int H, int W, int iters, void* stream);
/* workspace: six fp32 volumes (G, a, D ping-pong pairs) + 2 x 64 per-level maxima (fixed-point scales of the routing kernel) */
size_t vg_soft_skel_bwd_workspace_bytes(int N, int D, int H, int W);
int vg_soft_skel_bwd(const float* E, const float* S, const float* gskel, float* dx, void* workspace, size_t workspace_bytes,
                     int N, int D, int H, int W, int iters, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Loss arithmetic (loss_functions.py:7-22,56-68,86-117,163-226,255-322; clDice_func.py:83-149;
 * utils.py:27-48).  Reductions accumulate (+=) into caller-zeroed double accumulators.
 * ------------------------------------------------------------------------------------------- */
int vg_minmax(const float* x, int N, size_t V, float* mm, void* enc_ws, void* stream);
int vg_minmax_normalize(const float* x, const float* mm, float* out, int N, size_t V, void* stream);
int vg_minmax_normalize_bwd(const float* x, const float* nrm, const float* mm, const float* g, float* dx, int N, size_t V,
                            void* acc_ws, int accumulate, void* stream);
int vg_sqdiff_sum(const float* a, const float* b, float target, size_t n, double* acc, void* stream);
int vg_lincomb(float* out, size_t n, int accumulate, float c0, const float* x1, float c1, const float* x2, float c2,
               const float* x3, float c3, void* stream);
/* as vg_lincomb with the coefficients {c0,c1,c2,c3} read from device memory */
int vg_lincomb_dev(float* out, size_t n, int accumulate, const float* coef4, const float* x1, const float* x2, const float* x3,
                   void* stream);
/* coef8 (device) <- the six coefficients of the clDice/Dice backward from the seven sums of vg_cldice_sums, times k:
 * d skel_pred = coef[0] + coef[1]*y_true;  d y_pred = coef[2] + coef[3]*y_true + coef[4]*skel_true + coef[5]*d0 */
int vg_cldice_coeffs(const double* acc7, float alpha, float k, float* coef8, void* stream);
int vg_bce_sum(const float* y_true, const float* y_pred, size_t n, double* acc, void* stream);
int vg_bce_bwd(const float* y_true, const float* y_pred, float coef, float* g, size_t n, int accumulate, void* stream);
int vg_cldice_sums(const float* y_true, const float* y_pred, const float* skel_true, const float* skel_pred, size_t n,
                   double* acc7, void* stream);
int vg_ssim_fwd(const float* t, const float* p, int N, int D, int H, int W, double* acc, float* mA, float* mB, float* mC,
                void* stream);
int vg_ssim_bwd(const float* t, const float* p, const float* mA, const float* mB, const float* mC, int N, int D, int H, int W,
                float coef, float* gp, int accumulate, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Keras OptimizerV2 Adam with per-variable clipnorm (vangan.py:220-235, applied by minimize at
 * :426-438).  One flat fp32 buffer per network; seg_offsets[nseg+1] (device) delimits the variables,
 * total = seg_offsets[nseg]; norm_ws: nseg doubles of scratch.
 * ------------------------------------------------------------------------------------------- */
int vg_clip_adam_step(float* w, const float* g, float* m, float* v, const long long* seg_offsets, int nseg, long long total,
                      float lr_t, float beta1, float beta2, float eps, float clipnorm, double* norm_ws, void* stream);
/* same, with the bias-corrected step size lr_t read from device memory (it changes every step; a captured graph cannot
 * carry it as a launch argument) */
int vg_clip_adam_step_dev(float* w, const float* g, float* m, float* v, const long long* seg_offsets, int nseg, long long total,
                          const float* lr_t_dev, float beta1, float beta2, float eps, float clipnorm, double* norm_ws, void* stream);

/* ---------------------------------------------------------------------------------------------
 * V-Net gen_SI variant (vangan.py:135-149): BatchNormalization (vnet_model.py:127-128,142-143) and Conv3DTranspose k2 s2
 * (vnet_model.py:245).
 *
 * BatchNormalization(axis=-1, momentum=0.99, epsilon=1e-3), per-replica batch statistics.  The arithmetic is the InstanceNorm one
 * with statistics over N*D*H*W: vg_batchnorm_stats writes them replicated per (n, c) (training: batch statistics + Keras moving
 * averages; inference: from the moving values), vg_instnorm_apply applies them, vg_batchnorm_bwd is vg_instnorm_bwd with the
 * reductions taken over the whole batch (VG_IN_BATCH_STATS).  ws: vg_batchnorm_workspace_bytes.
 *
 * Conv3DTranspose((2,2,2), strides 2, 'same'): the eight outputs of an input voxel do not overlap, so the layer is one pointwise GEMM
 * with 8*Cout columns ordered (a,b,c,co) -- vg_conv3d_fwd / dgrad / wgrad with K = 1 on a kernel stored as (1,1,1,Cin,8*Cout) --
 * plus a depth-to-space scatter that adds the bias (forward) / a space-to-depth gather that also reduces the bias gradient (backward).
 * The Keras kernel (2,2,2,Cout,Cin) maps to the GEMM layout by w1[ci][((a*2+b)*2+c)*Cout + co] = wk[a][b][c][co][ci].
 * ------------------------------------------------------------------------------------------- */
size_t vg_batchnorm_workspace_bytes(int N, int D, int H, int W, int C);
int vg_batchnorm_stats(const void* x, int dtype, int N, int D, int H, int W, int C, float* mean_nc, float* rstd_nc, float* moving_mean,
                       float* moving_var, float momentum, int training, void* ws, size_t ws_bytes, void* stream);
int vg_batchnorm_bwd(const vg_instnorm_desc* d, const void* dy, const void* x, const float* mean_nc, const float* rstd_nc,
                     const float* gamma, const float* beta, const float* drop, void* dx, int accumulate_dx, void* dres, float* dgamma,
                     float* dbeta, void* ws, size_t ws_bytes, void* stream);
/* y[N,2D,2H,2W,Cout] = depth_to_space(t[N,D,H,W,8*Cout]) + bias */
int vg_conv3d_transpose_k2s2_scatter(const void* t, const float* bias, void* y, int N, int D, int H, int W, int Cout, void* stream);
/* dt = space_to_depth(dy); dbias (optional, fp32[Cout]) += sum dy */
int vg_conv3d_transpose_k2s2_gather(const void* dy, void* dt, float* dbias, int N, int D, int H, int W, int Cout, void* stream);
/* 'resnet' generator (generator.py:7-73): UpSampling3D(2) + the zero padding of the Conv3D(k4, s1, 'same') that follows it
 * (building_blocks.py:240-280; TF pads 1 before / 2 after): out[N, 2D+lo+hi, 2H+lo+hi, 2W+lo+hi, C] bf16, and its adjoint.  The 7x7x7
 * convolutions of that generator (generator.py:38,67: one side has a single channel) go through vg_conv3d_fwd / dgrad / wgrad with K = 7. */
int vg_upsample_pad(const void* a, void* out, int N, int D, int H, int W, int C, int lo, int hi, void* stream);
int vg_upsample_pad_bwd(const void* dout, void* da, int N, int D, int H, int W, int C, int lo, int hi, void* stream);
/* dir 0: gemm[Cin][8*Cout] = permute(keras[2][2][2][Cout][Cin]); dir 1: keras_grad += permute(gemm_grad) */
int vg_conv3d_transpose_k2s2_weights(const float* src, float* dst, int Cin, int Cout, int dir, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Data-parallel exchange (tf.distribute.MirroredStrategy: main.py:22, vangan.py:86; the gradient all-reduce hidden inside
 * `optimizer.minimize`, vangan.py:426-438; strategy.reduce(SUM) of the result dict, vangan.py:459-473).
 * One process per GPU; NCCL over NVLink / NVSwitch, resolved at run time (dlopen of libnccl.so.2; VG_NCCL_LIB overrides), so
 * single-GPU users need no NCCL.  vg_comm is the library's only opaque handle.  Rank 0 obtains a 128-byte id with
 * vg_comm_unique_id and hands it to the other ranks by any host-side channel; every rank then calls vg_comm_init (blocking
 * rendezvous).  Collectives are enqueued on the caller's stream (a communication stream ordered by events; capturable into a CUDA
 * graph), are in place, and SUM across ranks.  world == 1: every call is a no-op returning 0.
 * ------------------------------------------------------------------------------------------- */
typedef struct vg_comm vg_comm;
int vg_comm_available(void);        /* 1 when the NCCL symbols could be resolved */
int vg_comm_nccl_version(void);     /* NCCL_VERSION_CODE of the library in use, 0 when unavailable */
int vg_comm_unique_id(void* id128);
int vg_comm_init(vg_comm** out, const void* id128, int world, int rank, int device);
int vg_comm_world(const vg_comm* c);
int vg_comm_rank(const vg_comm* c);
unsigned long long vg_comm_collectives(const vg_comm* c);   /* messages enqueued so far */
/* fp32 gradient buffer of one network, cut into messages of bucket_elems elements (<= 0: one message), one grouped launch */
int vg_comm_allreduce_bucket(vg_comm* c, float* buf, long long count, long long bucket_elems, void* stream);
/* n fp64 scalars in device memory (the ten-entry result dict) */
int vg_comm_reduce_scalars(vg_comm* c, double* vals, int n, void* stream);
int vg_comm_destroy(vg_comm* c);

/* ---------------------------------------------------------------------------------------------
 * On-device input pipeline (dataset.py:205-251: tf.image.random_crop + random_spatial_augmentation; the random draws stay on the host).
 * vol: [H,W,D] fp32 (a whole .npy volume, uploaded once); out: [kH,kW,kD] = rot90_k(flip_up_down(flip_left_right(crop at (x0,y0,z0)))),
 * with tf.image's 4-D reading of a volume: left_right reverses D, up_down reverses W, rot90 turns the (W, D) plane counter-clockwise.
 * The "retry until max >= 0.8" loop of process_seg_domain (dataset.py:226-246) tests the crop with vg_minmax.
 * ------------------------------------------------------------------------------------------- */
int vg_crop_augment(const float* vol, int H, int W, int D, float* out, int kH, int kW, int kD, int x0, int y0, int z0, int flip_lr,
                    int flip_ud, int rot_k, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Sliding-window stitching (custom_callback.py:123,165-166,177-183,192,202).
 * pred/cnt: [H,W,D] fp32 volumes; win: [B,kH,kW,kD] generator outputs; starts: [B][3] window origins.
 * ------------------------------------------------------------------------------------------- */
/* win[B,kH,kW,kD] = windows of vol[H,W,D] at starts[B][3] */
int vg_stitch_gather(const float* vol, int H, int W, int D, float* win, const int* starts, int B, int kH, int kW, int kD, void* stream);
/* the same from the UN-padded volume [H0,W0,D0] when the reference pads first (np.pad 'symmetric' by xs/ys/zs, custom_callback.py:82-104):
 * starts are coordinates in the padded volume, every coordinate p reads sym(p - pad); vol holds the rows [row0, ...) of the volume */
int vg_stitch_gather_sym(const float* vol, int row0, int H0, int W0, int D0, int xs, int ys, int zs, float* win, const int* starts, int B,
                         int kH, int kW, int kD, void* stream);
int vg_stitch_accumulate(float* pred, float* cnt, int H, int W, int D, const float* win, const int* starts, int B, int kH,
                         int kW, int kD, int pH, int pW, int pD, void* stream);
/* out[oh,ow,od] = 255 * minmax_norm( (pred/cnt)[crop] ): two calls — divide+minmax, then scale */
int vg_stitch_finalize(const float* pred, const float* cnt, int H, int W, int D, int x0, int y0, int z0, int oH, int oW, int oD,
                       float* out, float* mm, void* enc_ws, void* stream);
int vg_stitch_scale(float* out, size_t n, const float* mm, void* stream);
/* Order-exact form of the same loop (custom_callback.py:142-192): every output voxel of rows [row0, row0+rows) adds the windows that
 * cover it in the reference's enumeration order (rows, columns, depth; clamped last window repeated), sequentially in fp32, divides by
 * the count and folds its value into the running min / max (enc_ws: 2 x u32, initialised when init_enc != 0).  wins: [slots][kH,kW,kD]
 * generator outputs; slot_of[(i*nW + j)*nD + k]: enumeration index -> slot; starts_dev: the nH + nW + nD per-axis window starts back to
 * back (<= 128 per axis); all DEVICE pointers.  Result is bit-identical to the numpy loop for any batching / sharding of the generator calls. */
int vg_stitch_gather_sum(const float* wins, const int* slot_of, const int* starts_dev, int nH, int nW, int nD, int kH, int kW, int kD, int pH,
                         int pW, int pD, int x0, int y0, int z0, int row0, int rows, int oW, int oD, float* out, void* enc_ws, int init_enc,
                         void* stream);
/* mm[0], mm[1] = decoded running min / max */
int vg_stitch_minmax_decode(const void* enc_ws, float* mm, void* stream);
/* out_u8 = (uint8)(255 * (in - mm[0]) / (mm[1] - mm[0]))   (custom_callback.py:202-205, complete=False) */
int vg_stitch_scale_u8(const float* in, unsigned char* out, size_t n, const float* mm, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VANGAN_B200_H */
