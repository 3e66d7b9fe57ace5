// InstanceNormalization (+ReLU / LeakyReLU) (+residual add) (+SpatialDropout3D) (+GaussianNoise)
// with the consumer's ReflectionPadding3D / zero 'same' padding written in the same pass.
//
// Replaces tfa.layers.InstanceNormalization (eps 1e-3, biased variance) and the Keras
// Activation / Add / ReflectionPadding3D / GaussianNoise / SpatialDropout3D layers around it
// (resunet_model.py:23-39,42-66,96-100,133-143; building_blocks.py:15-39,166-195;
// discriminator.py:50-52,70-72,105-106).  NDHWC storage, bf16 (product path) or fp32
// (parity-test path) activations, fp32 statistics.  All kernels are HBM-bound: 128-bit accesses,
// 8 channels per thread, per-(n,c) reductions by shared-memory tree + a deterministic
// second-stage reduction (no floating-point atomics).
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int NT = 256;
constexpr float IN_EPS = 1e-3f;

struct Geo {
    int N, D, H, W, C;
    int pad_lo, pad_hi, pad_mode;  // output (fwd) / incoming-gradient (bwd) padding
    int relu_in;                   // the normalised tensor is relu(x) (Conv3D(activation='relu') -> norm, vnet_model.py:118-130)
};

__device__ __forceinline__ float act_fwd(float z, int act, float slope) {
    if (act == VG_ACT_RELU) return fmaxf(z, 0.f);
    if (act == VG_ACT_LEAKY) return z > 0.f ? z : slope * z;
    return z;
}
__device__ __forceinline__ float act_grad(float z, int act, float slope) {
    if (act == VG_ACT_RELU) return z > 0.f ? 1.f : 0.f;
    if (act == VG_ACT_LEAKY) return z > 0.f ? 1.f : slope;
    return 1.f;
}

// Thread mapping shared by all four streaming kernels: a block has nthr = (256/cg)*cg threads
// (cg = C/8 channel groups), thread = (voxel lane, channel group), so a thread's 8 channels -- and their
// scale / shift / mean / rstd -- are fixed for its whole loop and live in registers.  Voxels are walked
// flat (grid-stride over the sample), U at a time, so every thread keeps U independent 128-bit loads in
// flight; (d,h,w) are carried incrementally (no divisions in the loop).  Raw 8-element packets are kept
// packed until they are consumed to hold the register count down.
constexpr int U = 4;

template <typename T> struct Raw;
template <> struct Raw<bf16> { bf16x8 v; };
template <> struct Raw<float> { float4 a, b; };
template <typename T> __device__ __forceinline__ void load_raw(const T* p, Raw<T>& r);
template <> __device__ __forceinline__ void load_raw<bf16>(const bf16* p, Raw<bf16>& r) { r.v = *reinterpret_cast<const uint4*>(p); }
template <> __device__ __forceinline__ void load_raw<float>(const float* p, Raw<float>& r) {
    r.a = reinterpret_cast<const float4*>(p)[0];
    r.b = reinterpret_cast<const float4*>(p)[1];
}
__device__ __forceinline__ void unpack_raw(const Raw<bf16>& r, float* f) { unpack8(r.v, f); }
__device__ __forceinline__ void unpack_raw(const Raw<float>& r, float* f) {
    f[0] = r.a.x; f[1] = r.a.y; f[2] = r.a.z; f[3] = r.a.w; f[4] = r.b.x; f[5] = r.b.y; f[6] = r.b.z; f[7] = r.b.w;
}

// walks voxels v = v0, v0+S, v0+2S, ... of a [PD][PH][PW] volume keeping (pd,ph,pw) without divisions
struct VoxIter {
    int pd, ph, pw, sd, sh, sw, PH, PW;
    __device__ __forceinline__ void init(int v, int S, int PH_, int PW_) {
        PH = PH_; PW = PW_;
        pw = v % PW; int r = v / PW; ph = r % PH; pd = r / PH;
        sw = S % PW; r = S / PW; sh = r % PH; sd = r / PH;
    }
    __device__ __forceinline__ void next() {
        pw += sw;
        int c = pw >= PW;
        pw -= c ? PW : 0;
        ph += sh + c;
        c = ph >= PH;
        ph -= c ? PH : 0;
        pd += sd + c;
    }
};

// Sample order of the FIRST pass of a two-pass operation (statistics before apply, backward reduction before backward apply): last
// sample first.  The producing convolution wrote the tensor in ascending sample order, so its tail is what the 126 MB L2 still holds;
// the pass then ends on sample 0, which is where the second pass (ascending) begins.  The result does not depend on the order
// (every (sample, block) partial has its own slot).  Measured at 8x128^3: statistics 4.8 -> 4.6 ms per step, backward unchanged
// (a sample's x + dy is 137 MB, more than the L2 holds) -- kept because it is free, not because it matters.
__device__ __forceinline__ int first_pass_sample() { return (int)(gridDim.y - 1 - blockIdx.y); }

// ------------------------------------------------------------------ statistics
// partial[(n*nblk + blk)*C*2 + c*2 + {0,1}] = sum / sumsq of (x - shift_c) over the block's voxels,
// shift_c = x[n,0,c] (keeps E[x^2]-E[x]^2 well conditioned)
template <typename T>
__global__ void __launch_bounds__(NT) in_stats_partial_kernel(const T* __restrict__ x, int V, int C,
                                                              float* __restrict__ partial, int relu_in) {
    extern __shared__ float sm[];  // [vlanes][C][2]
    const int n = first_pass_sample(), cg = C / 8;
    const int c8 = threadIdx.x % cg, vl = threadIdx.x / cg, nvl = blockDim.x / cg;
    const T* xn = x + (size_t)n * V * C + c8 * 8;
    float shift[8], s1[8], s2[8];
    load8<T>(xn, shift);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        s1[k] = s2[k] = 0.f;
        if (relu_in) shift[k] = fmaxf(shift[k], 0.f);
    }
    const int S = gridDim.x * nvl;
    for (int v = blockIdx.x * nvl + vl; v < V; v += U * S) {
        Raw<T> raw[U];
#pragma unroll
        for (int u = 0; u < U; u++)
            if (v + u * S < V) load_raw<T>(xn + (size_t)(v + u * S) * C, raw[u]);
#pragma unroll
        for (int u = 0; u < U; u++)
            if (v + u * S < V) {
                float f[8];
                unpack_raw(raw[u], f);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    float d = (relu_in ? fmaxf(f[k], 0.f) : f[k]) - shift[k];
                    s1[k] += d;
                    s2[k] += d * d;
                }
            }
    }
    float* row = sm + ((size_t)vl * C + c8 * 8) * 2;
#pragma unroll
    for (int k = 0; k < 8; k++) { row[2 * k] = s1[k]; row[2 * k + 1] = s2[k]; }
    __syncthreads();
    float* out = partial + ((size_t)n * gridDim.x + blockIdx.x) * C * 2;
    for (int i = threadIdx.x; i < C * 2; i += blockDim.x) {
        float a = 0.f;
        for (int l = 0; l < nvl; l++) a += sm[(size_t)l * C * 2 + i];
        out[i] = a;
    }
}

// one warp per (n,c): lanes stride over the block partials, shuffle-reduce in double
template <typename T>
__global__ void __launch_bounds__(256) in_stats_final_kernel(const T* __restrict__ x, const float* __restrict__ partial,
                                                             size_t V, int C, int nblk, int N, float* __restrict__ mean,
                                                             float* __restrict__ rstd, int relu_in) {
    int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= N * C) return;
    int n = i / C, c = i % C;
    double s1 = 0, s2 = 0;
    for (int b = lane; b < nblk; b += 32) {
        const float* p = partial + (((size_t)n * nblk + b) * C + c) * 2;
        s1 += p[0];
        s2 += p[1];
    }
    s1 = warp_sum_d(s1);
    s2 = warp_sum_d(s2);
    if (lane == 0) {
        double shift = (double)(float)x[(size_t)n * V * C + c];
        if (relu_in && shift < 0) shift = 0;
        double m = s1 / (double)V;
        double var = s2 / (double)V - m * m;
        if (var < 0) var = 0;
        mean[i] = (float)(shift + m);
        rstd[i] = (float)(1.0 / sqrt(var + (double)IN_EPS));
    }
}

// ------------------------------------------------------------------ forward apply
struct ApplyArgs {
    const float *mean, *rstd, *gamma, *beta, *drop, *noise;
    float slope, noise_std;
    int act;
    unsigned long long seed;
    const unsigned long long* seed_dev;   // optional per-step seed offset in device memory
};

template <typename T>
__global__ void __launch_bounds__(NT, 2) in_apply_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y,
                                                      Geo g, ApplyArgs a) {
    const int n = blockIdx.y, cg = g.C / 8;
    const int c8 = threadIdx.x % cg, vl = threadIdx.x / cg, nvl = blockDim.x / cg;
    const int PD = g.D + g.pad_lo + g.pad_hi, PH = g.H + g.pad_lo + g.pad_hi, PW = g.W + g.pad_lo + g.pad_hi;
    const int M = PD * PH * PW;
    float scale[8], shift[8], drop[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int c = c8 * 8 + k, sc = n * g.C + c;
        scale[k] = a.gamma[c] * a.rstd[sc];
        shift[k] = a.beta[c] - a.mean[sc] * scale[k];
        drop[k] = a.drop ? a.drop[sc] : 1.f;
    }
    const size_t in_sample = (size_t)g.D * g.H * g.W * g.C;
    const T* xn = x + (size_t)n * in_sample + c8 * 8;
    const T* rn = res ? res + (size_t)n * in_sample + c8 * 8 : nullptr;
    T* yn = y + (size_t)n * M * g.C + c8 * 8;
    const int S = gridDim.x * nvl;
    VoxIter it;
    it.init(blockIdx.x * nvl + vl, S, PH, PW);
    for (int v = blockIdx.x * nvl + vl; v < M; v += U * S) {
        Raw<T> rx[U], rr[U];
        int src[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            src[u] = -1;
            if (v + u * S < M) {
                int d = it.pd - g.pad_lo, h = it.ph - g.pad_lo, w = it.pw - g.pad_lo;
                const bool oob = (unsigned)d >= (unsigned)g.D || (unsigned)h >= (unsigned)g.H || (unsigned)w >= (unsigned)g.W;
                if (!(oob && g.pad_mode == VG_PAD_ZERO)) {
                    if (oob) { d = reflect1(d, g.D); h = reflect1(h, g.H); w = reflect1(w, g.W); }
                    src[u] = (d * g.H + h) * g.W + w;
                    load_raw<T>(xn + (size_t)src[u] * g.C, rx[u]);
                    if (rn) load_raw<T>(rn + (size_t)src[u] * g.C, rr[u]);
                }
            }
            it.next();
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int vv = v + u * S;
            if (vv >= M) continue;
            float o[8];
            if (src[u] < 0) {
#pragma unroll
                for (int k = 0; k < 8; k++) o[k] = 0.f;
            } else {
                float f[8];
                unpack_raw(rx[u], f);
#pragma unroll
                for (int k = 0; k < 8; k++)
                    o[k] = act_fwd(fmaf(g.relu_in ? fmaxf(f[k], 0.f) : f[k], scale[k], shift[k]), a.act, a.slope) * drop[k];
                if (rn) {
                    unpack_raw(rr[u], f);
#pragma unroll
                    for (int k = 0; k < 8; k++) o[k] += f[k];
                }
                if (a.noise) {
                    // explicit noise tensor: padded layout for REFLECT (noise is added after the pad layer),
                    // unpadded layout for ZERO ('same' convs pad after the noise layer)
                    const float* np = g.pad_mode == VG_PAD_REFLECT ? a.noise + ((size_t)n * M + vv) * g.C + c8 * 8
                                                                   : a.noise + (size_t)n * in_sample + (size_t)src[u] * g.C + c8 * 8;
#pragma unroll
                    for (int k = 0; k < 8; k++) o[k] += np[k];
                } else if (a.noise_std > 0.f) {
                    const unsigned long long i = ((unsigned long long)n * M + vv) * cg + c8;
                    const unsigned long long sd_ = a.seed + (a.seed_dev ? *a.seed_dev : 0ull);
                    uint2 key = make_uint2((uint32_t)sd_, (uint32_t)(sd_ >> 32));
                    uint4 r0 = philox4x32(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 0u, 0x56414e47u), key);
                    uint4 r1 = philox4x32(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 1u, 0x56414e47u), key);
                    float2 n0 = box_muller(r0.x, r0.y), n1 = box_muller(r0.z, r0.w), n2 = box_muller(r1.x, r1.y),
                           n3 = box_muller(r1.z, r1.w);
                    o[0] += a.noise_std * n0.x; o[1] += a.noise_std * n0.y; o[2] += a.noise_std * n1.x; o[3] += a.noise_std * n1.y;
                    o[4] += a.noise_std * n2.x; o[5] += a.noise_std * n2.y; o[6] += a.noise_std * n3.x; o[7] += a.noise_std * n3.y;
                }
            }
            store8<T>(yn + (size_t)vv * g.C, o);
        }
    }
}

// ------------------------------------------------------------------ backward
struct BwdArgs {
    const float *mean, *rstd, *gamma, *beta, *drop;
    float slope;
    int act;
};

// partial[(n*nblk+blk)*C*2 + c*2 + {0,1}] = sum g, sum g*xhat  with g = dy*drop*act'(z).  The sums are linear in
// dy, so the pass walks the PADDED gradient tensor and reads x at the mirrored voxel: no folding needed here.
template <typename T>
__global__ void __launch_bounds__(NT, 2) in_bwd_partial_kernel(const T* __restrict__ dy, const T* __restrict__ x, Geo g,
                                                            BwdArgs a, float* __restrict__ partial) {
    extern __shared__ float sm[];
    const int n = first_pass_sample(), C = g.C, cg = C / 8;
    const int c8 = threadIdx.x % cg, vl = threadIdx.x / cg, nvl = blockDim.x / cg;
    const int PD = g.D + g.pad_lo + g.pad_hi, PH = g.H + g.pad_lo + g.pad_hi, PW = g.W + g.pad_lo + g.pad_hi;
    const int M = PD * PH * PW;
    // per-channel constants kept to three (register budget = loads in flight): z = x*sc + sh decides act', and the sums are
    // taken of g' = dy*act' and g'*(x - mu); the channel factors drop and drop*rstd are applied once at the end
    float mu[8], sc[8], sh[8], s1[8], s2[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int c = c8 * 8 + k;
        mu[k] = a.mean[n * C + c];
        sc[k] = a.gamma[c] * a.rstd[n * C + c];
        sh[k] = a.beta[c] - mu[k] * sc[k];
        s1[k] = s2[k] = 0.f;
    }
    const T* xn = x + (size_t)n * g.D * g.H * g.W * C + c8 * 8;
    const T* dyn = dy + (size_t)n * M * C + c8 * 8;
    const int S = gridDim.x * nvl;
    VoxIter it;
    it.init(blockIdx.x * nvl + vl, S, PH, PW);
    for (int v = blockIdx.x * nvl + vl; v < M; v += U * S) {
        Raw<T> rx[U], rg[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            ok[u] = false;
            if (v + u * S < M) {
                int d = it.pd - g.pad_lo, h = it.ph - g.pad_lo, w = it.pw - g.pad_lo;
                const bool oob = (unsigned)d >= (unsigned)g.D || (unsigned)h >= (unsigned)g.H || (unsigned)w >= (unsigned)g.W;
                if (!(oob && g.pad_mode == VG_PAD_ZERO)) {
                    if (oob) { d = reflect1(d, g.D); h = reflect1(h, g.H); w = reflect1(w, g.W); }
                    ok[u] = true;
                    load_raw<T>(xn + (size_t)((d * g.H + h) * g.W + w) * C, rx[u]);
                    load_raw<T>(dyn + (size_t)(v + u * S) * C, rg[u]);
                }
            }
            it.next();
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            if (ok[u]) {
                float f[8], gy[8];
                unpack_raw(rx[u], f);
                unpack_raw(rg[u], gy);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const float xr = g.relu_in ? fmaxf(f[k], 0.f) : f[k];
                    float gg = gy[k] * act_grad(fmaf(xr, sc[k], sh[k]), a.act, a.slope);
                    s1[k] += gg;
                    s2[k] = fmaf(gg, xr - mu[k], s2[k]);
                }
            }
    }
    float* row = sm + ((size_t)vl * C + c8 * 8) * 2;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int c = c8 * 8 + k;
        const float dr = a.drop ? a.drop[n * C + c] : 1.f;
        row[2 * k] = s1[k] * dr;
        row[2 * k + 1] = s2[k] * dr * a.rstd[n * C + c];
    }
    __syncthreads();
    float* out = partial + ((size_t)n * gridDim.x + blockIdx.x) * C * 2;
    for (int i = threadIdx.x; i < C * 2; i += blockDim.x) {
        float acc = 0.f;
        for (int l = 0; l < nvl; l++) acc += sm[(size_t)l * C * 2 + i];
        out[i] = acc;
    }
}

// sums[(n*C+c)*2+{0,1}] = S1, S2; dgamma[c] += S2, dbeta[c] += S1  (one warp per (n,c); N <= a few atomics per channel).
// batch != 0 (BatchNormalization: statistics over N*D*H*W): one warp per channel adds the partials of ALL samples and stores
// total / N for every n, so that the apply pass (which divides by the per-sample voxel count V) sees total / (N*V).
__global__ void __launch_bounds__(256) in_bwd_final_kernel(const float* __restrict__ partial, int nblk, int N, int C,
                                                           float* __restrict__ sums, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, int batch) {
    int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= (batch ? C : N * C)) return;
    int c = i % C;
    int n = i / C;
    double s1 = 0, s2 = 0;
    const int n0 = batch ? 0 : n, n1 = batch ? N : n + 1;
    for (int nn = n0; nn < n1; nn++)
        for (int b = lane; b < nblk; b += 32) {
            const float* p = partial + (((size_t)nn * nblk + b) * C + c) * 2;
            s1 += p[0];
            s2 += p[1];
        }
    s1 = warp_sum_d(s1);
    s2 = warp_sum_d(s2);
    if (lane == 0) {
        if (batch) {
            for (int nn = 0; nn < N; nn++) {
                sums[((size_t)nn * C + c) * 2] = (float)(s1 / N);
                sums[((size_t)nn * C + c) * 2 + 1] = (float)(s2 / N);
            }
        } else {
            sums[(size_t)i * 2] = (float)s1;
            sums[(size_t)i * 2 + 1] = (float)s2;
        }
        if (dgamma) atomicAdd(dgamma + c, (float)s2);
        if (dbeta) atomicAdd(dbeta + c, (float)s1);
    }
}

// BatchNormalization statistics: stat[c] (computed as ONE instance over the N stacked samples) -> replicated per (n, c), and the
// Keras moving averages: moving = momentum * moving + (1 - momentum) * batch value.  training == 0: mean / rstd from the moving values.
__global__ void bn_expand_kernel(const float* __restrict__ mean_c, const float* __restrict__ rstd_c, int N, int C, float* __restrict__ mean_nc,
                                 float* __restrict__ rstd_nc, float* __restrict__ moving_mean, float* __restrict__ moving_var, float momentum,
                                 int training, float bessel) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float m, r;
    if (training) {
        m = mean_c[c];
        r = rstd_c[c];
        if (moving_mean) {
            // Keras' fused path (5-D NDHWC input) feeds the moving variance with the Bessel-corrected batch variance
            const float var = (1.f / (r * r) - IN_EPS) * bessel;
            moving_mean[c] = momentum * moving_mean[c] + (1.f - momentum) * m;
            moving_var[c] = momentum * moving_var[c] + (1.f - momentum) * var;
        }
    } else {
        m = moving_mean[c];
        r = rsqrtf(moving_var[c] + IN_EPS);
    }
    for (int n = 0; n < N; n++) {
        mean_nc[n * C + c] = m;
        rstd_nc[n * C + c] = r;
    }
}

// ------------------------------------------------------------------ bias-gradient sinks
// dx of a norm IS dy of the convolution that produced x (and dres is dy of the convolution that produced the residual), so the
// per-channel sums those convolutions need for their bias gradients (a pass of its own over dy otherwise: channel_sum_kernel) are
// taken here from the values as they are STORED (bf16-rounded).  Per-thread partial sums -> shared-memory tree over the block's
// voxel lanes -> one atomicAdd per channel per block.
struct Sinks {
    float* dx;     // fp32[C] += sum over (n, voxels) of the stored dx, or NULL
    float* dres;   // same for dres
};
template <typename T> __device__ __forceinline__ float stored(float v);
template <> __device__ __forceinline__ float stored<bf16>(float v) { return __bfloat162float(__float2bfloat16(v)); }
template <> __device__ __forceinline__ float stored<float>(float v) { return v; }

__device__ __forceinline__ void sink_flush(const float* acc, float* sink, int C, int c8, int vl, int nvl, float* sm) {
    // sm: [nvl][C] floats (dynamic shared memory of the apply kernels when a sink is requested)
#pragma unroll
    for (int k = 0; k < 8; k++) sm[(size_t)vl * C + c8 * 8 + k] = acc[k];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float t = 0.f;
        for (int l = 0; l < nvl; l++) t += sm[(size_t)l * C + c];
        atomicAdd(sink + c, t);
    }
    __syncthreads();
}

// dx = gamma*rstd*(g - S1/V - xhat*S2/V) with g = fold(dy)*drop*act'(z); dres = fold(dy) (optional).
// fold: REFLECT -> every padded position whose mirror is this voxel (1 for interior voxels, up to 8 on the
// shell); ZERO -> the interior only.
template <typename T>
__global__ void __launch_bounds__(NT, 2) in_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, Geo g, BwdArgs a,
                                                          const float* __restrict__ sums, T* __restrict__ dx,
                                                          T* __restrict__ dres, int accumulate_dx, Sinks sk) {
    extern __shared__ float sink_sm[];
    float sx[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int n = blockIdx.y, C = g.C, cg = C / 8;
    const int c8 = threadIdx.x % cg, vl = threadIdx.x / cg, nvl = blockDim.x / cg;
    const int PH = g.H + g.pad_lo + g.pad_hi, PW = g.W + g.pad_lo + g.pad_hi, PD = g.D + g.pad_lo + g.pad_hi;
    const int V = g.D * g.H * g.W;
    const float invV = 1.f / (float)V;
    const bool refl = g.pad_mode == VG_PAD_REFLECT && (g.pad_lo | g.pad_hi);
    // dx = P*act'(z)*dy - Q - R*x with z = x*sc + sh:  sc = gamma*rstd, sh = beta - mean*sc, P = sc*drop,
    // R = sc*rstd*S2/V, Q = sc*S1/V - R*mean  (five per-channel constants instead of seven: registers buy loads in flight)
    float sc[8], sh[8], cP[8], cQ[8], cR[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int c = c8 * 8 + k, sci = n * C + c;
        const float mu = a.mean[sci], rs = a.rstd[sci];
        sc[k] = a.gamma[c] * rs;
        sh[k] = a.beta[c] - mu * sc[k];
        cP[k] = sc[k] * (a.drop ? a.drop[sci] : 1.f);
        cR[k] = sc[k] * rs * sums[2 * sci + 1] * invV;
        cQ[k] = sc[k] * sums[2 * sci] * invV - cR[k] * mu;
    }
    const T* xn = x + (size_t)n * V * C + c8 * 8;
    const T* dyn = dy + (size_t)n * PD * PH * PW * C + c8 * 8;
    T* dxn = dx + (size_t)n * V * C + c8 * 8;
    T* drn = dres ? dres + (size_t)n * V * C + c8 * 8 : nullptr;
    const int S = gridDim.x * nvl;
    VoxIter it;
    it.init(blockIdx.x * nvl + vl, S, g.H, g.W);
    for (int v = blockIdx.x * nvl + vl; v < V; v += U * S) {
        Raw<T> rx[U], rg[U];
        int cd[U], ch[U], cw[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            cd[u] = it.pd; ch[u] = it.ph; cw[u] = it.pw;
            if (v + u * S < V) {
                load_raw<T>(xn + (size_t)(v + u * S) * C, rx[u]);
                load_raw<T>(dyn + (size_t)(((cd[u] + g.pad_lo) * PH + ch[u] + g.pad_lo) * PW + cw[u] + g.pad_lo) * C, rg[u]);
            }
            it.next();
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int vv = v + u * S;
            if (vv >= V) continue;
            float f[8], gy[8];
            unpack_raw(rx[u], f);
            unpack_raw(rg[u], gy);
            const int d = cd[u], h = ch[u], w = cw[u];
            if (refl && (d == 1 || d == g.D - 2 || h == 1 || h == g.H - 2 || w == 1 || w == g.W - 2)) {
                // shell voxel: add the mirrored halo positions (rare)
                int dd[3], hh[3], ww[3], nd = 1, nh = 1, nw = 1;
                dd[0] = d + 1; hh[0] = h + 1; ww[0] = w + 1;
                if (d == 1) dd[nd++] = 0;
                if (d == g.D - 2) dd[nd++] = g.D + 1;
                if (h == 1) hh[nh++] = 0;
                if (h == g.H - 2) hh[nh++] = g.H + 1;
                if (w == 1) ww[nw++] = 0;
                if (w == g.W - 2) ww[nw++] = g.W + 1;
                for (int i0 = 0; i0 < nd; i0++)
                    for (int i1 = 0; i1 < nh; i1++)
                        for (int i2 = 0; i2 < nw; i2++) {
                            if (i0 + i1 + i2 == 0) continue;
                            float t[8];
                            load8<T>(dyn + (size_t)((dd[i0] * PH + hh[i1]) * PW + ww[i2]) * C, t);
#pragma unroll
                            for (int k = 0; k < 8; k++) gy[k] += t[k];
                        }
            }
            if (drn) {
                store8<T>(drn + (size_t)vv * C, gy);
                if (sk.dres) {
#pragma unroll
                    for (int k = 0; k < 8; k++) sr[k] += stored<T>(gy[k]);
                }
            }
            float o[8];
            if (accumulate_dx) load8<T>(dxn + (size_t)vv * C, o);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float xr = g.relu_in ? fmaxf(f[k], 0.f) : f[k];
                float gg = gy[k] * act_grad(fmaf(xr, sc[k], sh[k]), a.act, a.slope);
                float val = fmaf(cP[k], gg, -fmaf(cR[k], xr, cQ[k]));
                if (g.relu_in && !(f[k] > 0.f)) val = 0.f;   // gradient through the producer's ReLU
                o[k] = accumulate_dx ? o[k] + val : val;
            }
            store8<T>(dxn + (size_t)vv * C, o);
            if (sk.dx) {
#pragma unroll
                for (int k = 0; k < 8; k++) sx[k] += stored<T>(o[k]);
            }
        }
    }
    if (sk.dx) sink_flush(sx, sk.dx, C, c8, vl, nvl, sink_sm);
    if (sk.dres) sink_flush(sr, sk.dres, C, c8, vl, nvl, sink_sm);
}

// ================================================================== specialised instances (generator hot path)
// Same arithmetic as the generic kernels above with the options fixed at compile time, so the unrolled inner loops are
// straight-line code: SP = 1: InstanceNorm -> ReLU -> ReflectionPadding3D (conv_block, resunet_model.py:42-66);
// SP = 2: InstanceNorm + residual Add, no padding (resunet_model.py:96-100,133-143).  bf16, no dropout / noise.
// SP: 0 = generic (all options read at run time); 1 = InstanceNorm -> ReLU -> ReflectionPadding3D (conv_block, resunet_model.py:42-66);
// 2 = InstanceNorm + residual Add, no padding (resunet_model.py:96-100,133-143).  The specialised instances have a
// straight-line inner loop (the generic one carries ~200 branches per unrolled body).
template <typename T, int SP>
__global__ void __launch_bounds__(NT, 2) in_apply_sp_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y,
                                                      Geo g, ApplyArgs a) {
    const int act = SP == 0 ? a.act : (SP == 1 ? VG_ACT_RELU : VG_ACT_NONE);
    const int pad_lo = SP == 0 ? g.pad_lo : (SP == 1 ? 1 : 0), pad_hi = SP == 0 ? g.pad_hi : (SP == 1 ? 1 : 0);
    const int pad_mode = SP == 0 ? pad_mode : VG_PAD_REFLECT;
    const bool relu_in = SP == 0 && g.relu_in, has_drop = SP == 0 && a.drop != nullptr;
    const int n = blockIdx.y, cg = g.C / 8;
    const int c8 = threadIdx.x % cg, vl = threadIdx.x / cg, nvl = blockDim.x / cg;
    const int PD = g.D + pad_lo + pad_hi, PH = g.H + pad_lo + pad_hi, PW = g.W + pad_lo + pad_hi;
    const int M = PD * PH * PW;
    float scale[8], shift[8], drop[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int c = c8 * 8 + k, sc = n * g.C + c;
        scale[k] = a.gamma[c] * a.rstd[sc];
        shift[k] = a.beta[c] - a.mean[sc] * scale[k];
        drop[k] = has_drop ? a.drop[sc] : 1.f;
    }
    const size_t in_sample = (size_t)g.D * g.H * g.W * g.C;
    const T* xn = x + (size_t)n * in_sample + c8 * 8;
    const T* rn = (SP == 0 ? res != nullptr : SP == 2) ? res + (size_t)n * in_sample + c8 * 8 : nullptr;
    T* yn = y + (size_t)n * M * g.C + c8 * 8;
    const int S = gridDim.x * nvl;
    VoxIter it;
    it.init(blockIdx.x * nvl + vl, S, PH, PW);
    for (int v = blockIdx.x * nvl + vl; v < M; v += U * S) {
        Raw<T> rx[U], rr[U];
        int src[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            src[u] = -1;
            if (v + u * S < M) {
                int d = it.pd - pad_lo, h = it.ph - pad_lo, w = it.pw - pad_lo;
                const bool oob = (unsigned)d >= (unsigned)g.D || (unsigned)h >= (unsigned)g.H || (unsigned)w >= (unsigned)g.W;
                if (!(oob && pad_mode == VG_PAD_ZERO)) {
                    if (oob) { d = reflect1(d, g.D); h = reflect1(h, g.H); w = reflect1(w, g.W); }
                    src[u] = (d * g.H + h) * g.W + w;
                    load_raw<T>(xn + (size_t)src[u] * g.C, rx[u]);
                    if (SP == 2 || (SP == 0 && rn)) load_raw<T>(rn + (size_t)src[u] * g.C, rr[u]);
                }
            }
            it.next();
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int vv = v + u * S;
            if (vv >= M) continue;
            float o[8];
            if (src[u] < 0) {
#pragma unroll
                for (int k = 0; k < 8; k++) o[k] = 0.f;
            } else {
                float f[8];
                unpack_raw(rx[u], f);
#pragma unroll
                for (int k = 0; k < 8; k++)
                    o[k] = act_fwd(fmaf(relu_in ? fmaxf(f[k], 0.f) : f[k], scale[k], shift[k]), act, a.slope) * (SP == 0 ? drop[k] : 1.f);
                if (SP == 2 || (SP == 0 && rn)) {
                    unpack_raw(rr[u], f);
#pragma unroll
                    for (int k = 0; k < 8; k++) o[k] += f[k];
                }
                if (SP == 0 && a.noise) {
                    // explicit noise tensor: padded layout for REFLECT (noise is added after the pad layer),
                    // unpadded layout for ZERO ('same' convs pad after the noise layer)
                    const float* np = pad_mode == VG_PAD_REFLECT ? a.noise + ((size_t)n * M + vv) * g.C + c8 * 8
                                                                   : a.noise + (size_t)n * in_sample + (size_t)src[u] * g.C + c8 * 8;
#pragma unroll
                    for (int k = 0; k < 8; k++) o[k] += np[k];
                } else if (SP == 0 && a.noise_std > 0.f) {
                    const unsigned long long i = ((unsigned long long)n * M + vv) * cg + c8;
                    const unsigned long long sd_ = a.seed + (a.seed_dev ? *a.seed_dev : 0ull);
                    uint2 key = make_uint2((uint32_t)sd_, (uint32_t)(sd_ >> 32));
                    uint4 r0 = philox4x32(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 0u, 0x56414e47u), key);
                    uint4 r1 = philox4x32(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 1u, 0x56414e47u), key);
                    float2 n0 = box_muller(r0.x, r0.y), n1 = box_muller(r0.z, r0.w), n2 = box_muller(r1.x, r1.y),
                           n3 = box_muller(r1.z, r1.w);
                    o[0] += a.noise_std * n0.x; o[1] += a.noise_std * n0.y; o[2] += a.noise_std * n1.x; o[3] += a.noise_std * n1.y;
                    o[4] += a.noise_std * n2.x; o[5] += a.noise_std * n2.y; o[6] += a.noise_std * n3.x; o[7] += a.noise_std * n3.y;
                }
            }
            store8<T>(yn + (size_t)vv * g.C, o);
        }
    }
}

template <typename T, int SP>
__global__ void __launch_bounds__(NT, 2) in_bwd_partial_sp_kernel(const T* __restrict__ dy, const T* __restrict__ x, Geo g,
                                                            BwdArgs a, float* __restrict__ partial) {
    const int act = SP == 0 ? a.act : (SP == 1 ? VG_ACT_RELU : VG_ACT_NONE);
    const int pad_lo = SP == 0 ? g.pad_lo : (SP == 1 ? 1 : 0), pad_hi = SP == 0 ? g.pad_hi : (SP == 1 ? 1 : 0);
    const int pad_mode = SP == 0 ? pad_mode : VG_PAD_REFLECT;
    const bool relu_in = SP == 0 && relu_in;
    extern __shared__ float sm[];
    const int n = first_pass_sample(), C = g.C, cg = C / 8;
    const int c8 = threadIdx.x % cg, vl = threadIdx.x / cg, nvl = blockDim.x / cg;
    const int PD = g.D + pad_lo + pad_hi, PH = g.H + pad_lo + pad_hi, PW = g.W + pad_lo + pad_hi;
    const int M = PD * PH * PW;
    // per-channel constants kept to three (register budget = loads in flight): z = x*sc + sh decides act', and the sums are
    // taken of g' = dy*act' and g'*(x - mu); the channel factors drop and drop*rstd are applied once at the end
    float mu[8], sc[8], sh[8], s1[8], s2[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int c = c8 * 8 + k;
        mu[k] = a.mean[n * C + c];
        sc[k] = a.gamma[c] * a.rstd[n * C + c];
        sh[k] = a.beta[c] - mu[k] * sc[k];
        s1[k] = s2[k] = 0.f;
    }
    const T* xn = x + (size_t)n * g.D * g.H * g.W * C + c8 * 8;
    const T* dyn = dy + (size_t)n * M * C + c8 * 8;
    const int S = gridDim.x * nvl;
    VoxIter it;
    it.init(blockIdx.x * nvl + vl, S, PH, PW);
    for (int v = blockIdx.x * nvl + vl; v < M; v += U * S) {
        Raw<T> rx[U], rg[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            ok[u] = false;
            if (v + u * S < M) {
                int d = it.pd - pad_lo, h = it.ph - pad_lo, w = it.pw - pad_lo;
                const bool oob = (unsigned)d >= (unsigned)g.D || (unsigned)h >= (unsigned)g.H || (unsigned)w >= (unsigned)g.W;
                if (!(oob && pad_mode == VG_PAD_ZERO)) {
                    if (oob) { d = reflect1(d, g.D); h = reflect1(h, g.H); w = reflect1(w, g.W); }
                    ok[u] = true;
                    load_raw<T>(xn + (size_t)((d * g.H + h) * g.W + w) * C, rx[u]);
                    load_raw<T>(dyn + (size_t)(v + u * S) * C, rg[u]);
                }
            }
            it.next();
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            if (ok[u]) {
                float f[8], gy[8];
                unpack_raw(rx[u], f);
                unpack_raw(rg[u], gy);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const float xr = relu_in ? fmaxf(f[k], 0.f) : f[k];
                    float gg = gy[k] * act_grad(fmaf(xr, sc[k], sh[k]), act, a.slope);
                    s1[k] += gg;
                    s2[k] = fmaf(gg, xr - mu[k], s2[k]);
                }
            }
    }
    float* row = sm + ((size_t)vl * C + c8 * 8) * 2;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int c = c8 * 8 + k;
        const float dr = (SP == 0 && a.drop) ? a.drop[n * C + c] : 1.f;
        row[2 * k] = s1[k] * dr;
        row[2 * k + 1] = s2[k] * dr * a.rstd[n * C + c];
    }
    __syncthreads();
    float* out = partial + ((size_t)n * gridDim.x + blockIdx.x) * C * 2;
    for (int i = threadIdx.x; i < C * 2; i += blockDim.x) {
        float acc = 0.f;
        for (int l = 0; l < nvl; l++) acc += sm[(size_t)l * C * 2 + i];
        out[i] = acc;
    }
}

// gy += every halo position of the padded gradient whose reflection is voxel (d,h,w)
template <typename T>
__device__ __noinline__ void shell_fold(const T* __restrict__ dyn, const Geo& g, int PH, int PW, int C, int d, int h, int w, float* gy) {
    int dd[3], hh[3], ww[3], nd = 1, nh = 1, nw = 1;
    dd[0] = d + 1; hh[0] = h + 1; ww[0] = w + 1;
    if (d == 1) dd[nd++] = 0;
    if (d == g.D - 2) dd[nd++] = g.D + 1;
    if (h == 1) hh[nh++] = 0;
    if (h == g.H - 2) hh[nh++] = g.H + 1;
    if (w == 1) ww[nw++] = 0;
    if (w == g.W - 2) ww[nw++] = g.W + 1;
    for (int i0 = 0; i0 < nd; i0++)
        for (int i1 = 0; i1 < nh; i1++)
            for (int i2 = 0; i2 < nw; i2++) {
                if (i0 + i1 + i2 == 0) continue;
                float t[8];
                load8<T>(dyn + (size_t)((dd[i0] * PH + hh[i1]) * PW + ww[i2]) * C, t);
#pragma unroll
                for (int k = 0; k < 8; k++) gy[k] += t[k];
            }
}

// dx = gamma*rstd*(g - S1/V - xhat*S2/V) with g = fold(dy)*drop*act'(z); dres = fold(dy) (optional).
// fold: REFLECT -> every padded position whose mirror is this voxel (1 for interior voxels, up to 8 on the
// shell); ZERO -> the interior only.
template <typename T, int SP, bool FOLDED>
__global__ void __launch_bounds__(NT, 2) in_bwd_apply_sp_kernel(const T* __restrict__ dy, const T* __restrict__ x, Geo g, BwdArgs a,
                                                          const float* __restrict__ sums, T* __restrict__ dx,
                                                          T* __restrict__ dres, int accumulate_dx, Sinks sk) {
    extern __shared__ float sink_sm[];
    float sx[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int act = SP == 0 ? a.act : (SP == 1 ? VG_ACT_RELU : VG_ACT_NONE);
    const int pad_lo = SP == 0 ? g.pad_lo : (SP == 1 ? 1 : 0), pad_hi = SP == 0 ? g.pad_hi : (SP == 1 ? 1 : 0);
    const int pad_mode = SP == 0 ? pad_mode : VG_PAD_REFLECT;
    const bool relu_in = SP == 0 && relu_in;
    const int n = blockIdx.y, C = g.C, cg = C / 8;
    const int c8 = threadIdx.x % cg, vl = threadIdx.x / cg, nvl = blockDim.x / cg;
    const int PH = g.H + pad_lo + pad_hi, PW = g.W + pad_lo + pad_hi, PD = g.D + pad_lo + pad_hi;
    const int V = g.D * g.H * g.W;
    const float invV = 1.f / (float)V;
    const bool refl = pad_mode == VG_PAD_REFLECT && (pad_lo | pad_hi);
    // dx = P*act'(z)*dy - Q - R*x with z = x*sc + sh:  sc = gamma*rstd, sh = beta - mean*sc, P = sc*drop,
    // R = sc*rstd*S2/V, Q = sc*S1/V - R*mean  (five per-channel constants instead of seven: registers buy loads in flight)
    float sc[8], sh[8], cP[8], cQ[8], cR[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int c = c8 * 8 + k, sci = n * C + c;
        const float mu = a.mean[sci], rs = a.rstd[sci];
        sc[k] = a.gamma[c] * rs;
        sh[k] = a.beta[c] - mu * sc[k];
        cP[k] = sc[k] * ((SP == 0 && a.drop) ? a.drop[sci] : 1.f);
        cR[k] = sc[k] * rs * sums[2 * sci + 1] * invV;
        cQ[k] = sc[k] * sums[2 * sci] * invV - cR[k] * mu;
    }
    const T* xn = x + (size_t)n * V * C + c8 * 8;
    const T* dyn = dy + (size_t)n * PD * PH * PW * C + c8 * 8;
    T* dxn = dx + (size_t)n * V * C + c8 * 8;
    T* drn = (SP == 0 ? dres != nullptr : SP == 2) ? dres + (size_t)n * V * C + c8 * 8 : nullptr;
    const int S = gridDim.x * nvl;
    VoxIter it;
    it.init(blockIdx.x * nvl + vl, S, g.H, g.W);
    for (int v = blockIdx.x * nvl + vl; v < V; v += U * S) {
        Raw<T> rx[U], rg[U];
        int cd[U], ch[U], cw[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            cd[u] = it.pd; ch[u] = it.ph; cw[u] = it.pw;
            if (v + u * S < V) {
                load_raw<T>(xn + (size_t)(v + u * S) * C, rx[u]);
                load_raw<T>(dyn + (size_t)(((cd[u] + pad_lo) * PH + ch[u] + pad_lo) * PW + cw[u] + pad_lo) * C, rg[u]);
            }
            it.next();
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int vv = v + u * S;
            if (vv >= V) continue;
            float f[8], gy[8];
            unpack_raw(rx[u], f);
            unpack_raw(rg[u], gy);
            const int d = cd[u], h = ch[u], w = cw[u];
            if (!FOLDED && refl && (d == 1 || d == g.D - 2 || h == 1 || h == g.H - 2 || w == 1 || w == g.W - 2))
                shell_fold<T>(dyn, g, PH, PW, C, d, h, w, gy);   // shell voxel: add the mirrored halo positions (rare, not inlined)
            if (drn) {
                store8<T>(drn + (size_t)vv * C, gy);
                if (sk.dres) {
#pragma unroll
                    for (int k = 0; k < 8; k++) sr[k] += stored<T>(gy[k]);
                }
            }
            float o[8];
            if (accumulate_dx) load8<T>(dxn + (size_t)vv * C, o);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float xr = relu_in ? fmaxf(f[k], 0.f) : f[k];
                float gg = gy[k] * act_grad(fmaf(xr, sc[k], sh[k]), act, a.slope);
                float val = fmaf(cP[k], gg, -fmaf(cR[k], xr, cQ[k]));
                if (relu_in && !(f[k] > 0.f)) val = 0.f;   // gradient through the producer's ReLU
                o[k] = accumulate_dx ? o[k] + val : val;
            }
            store8<T>(dxn + (size_t)vv * C, o);
            if (sk.dx) {
#pragma unroll
                for (int k = 0; k < 8; k++) sx[k] += stored<T>(o[k]);
            }
        }
    }
    if (sk.dx) sink_flush(sx, sk.dx, C, c8, vl, nvl, sink_sm);
    if (sk.dres) sink_flush(sr, sk.dres, C, c8, vl, nvl, sink_sm);
}

// ------------------------------------------------------------------ forward fast path: interior pass + halo pass
// InstanceNorm -> ReLU -> ReflectionPadding3D (conv_block, resunet_model.py:42-66), bf16.  The interior of the padded output is
// written by a pass that walks x linearly (the access pattern of the un-padded variant); the reflected halo is then copied from
// the interior cells it mirrors (6 faces, ~5 % of the tensor, mostly L2 hits).
template <typename T>
__global__ void __launch_bounds__(NT, 2) in_apply_interior_kernel(const T* __restrict__ x, T* __restrict__ y, Geo g, ApplyArgs a) {
    const int n = blockIdx.y, cg = g.C / 8;
    const int c8 = threadIdx.x % cg, vl = threadIdx.x / cg, nvl = blockDim.x / cg;
    const int PD = g.D + 2, PH = g.H + 2, PW = g.W + 2;
    const int V = g.D * g.H * g.W;
    float scale[8], shift[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int c = c8 * 8 + k, sc = n * g.C + c;
        scale[k] = a.gamma[c] * a.rstd[sc];
        shift[k] = a.beta[c] - a.mean[sc] * scale[k];
    }
    const T* xn = x + (size_t)n * V * g.C + c8 * 8;
    T* yn = y + (size_t)n * PD * PH * PW * g.C + c8 * 8;
    const int S = gridDim.x * nvl;
    VoxIter it;
    it.init(blockIdx.x * nvl + vl, S, g.H, g.W);
    for (int v = blockIdx.x * nvl + vl; v < V; v += U * S) {
        Raw<T> rx[U];
        int dst[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            dst[u] = ((it.pd + 1) * PH + it.ph + 1) * PW + it.pw + 1;
            if (v + u * S < V) load_raw<T>(xn + (size_t)(v + u * S) * g.C, rx[u]);
            it.next();
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (v + u * S >= V) continue;
            float f[8], o[8];
            unpack_raw(rx[u], f);
#pragma unroll
            for (int k = 0; k < 8; k++) o[k] = fmaxf(fmaf(f[k], scale[k], shift[k]), 0.f);
            store8<T>(yn + (size_t)dst[u] * g.C, o);
        }
    }
}

// y[halo cell] = y[mirrored interior cell] for a tensor padded by 1 (REFLECT); one thread per (halo cell, 8-channel group)
template <typename T>
__global__ void __launch_bounds__(256) pad_halo_inplace_kernel(T* __restrict__ y, Geo g) {
    const int C = g.C, cg = C / 8;
    const int PD = g.D + 2, PH = g.H + 2, PW = g.W + 2;
    const long long f0 = (long long)PH * PW, f1 = (long long)(PD - 2) * PW, f2 = (long long)(PD - 2) * (PH - 2);
    const long long total = 2 * (f0 + f1 + f2) * cg;
    T* yn = y + (size_t)blockIdx.y * PD * PH * PW * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cg);
        long long v = i / cg;
        int pd, ph, pw;
        if (v < 2 * f0) {                          // pd faces (whole planes)
            pd = v < f0 ? 0 : PD - 1;
            v %= f0;
            ph = (int)(v / PW); pw = (int)(v % PW);
        } else if (v < 2 * (f0 + f1)) {            // ph faces, pd interior
            v -= 2 * f0;
            ph = v < f1 ? 0 : PH - 1;
            v %= f1;
            pd = 1 + (int)(v / PW); pw = (int)(v % PW);
        } else {                                   // pw faces, pd and ph interior
            v -= 2 * (f0 + f1);
            pw = v < f2 ? 0 : PW - 1;
            v %= f2;
            pd = 1 + (int)(v / (PH - 2)); ph = 1 + (int)(v % (PH - 2));
        }
        const int d = reflect1(pd - 1, g.D), h = reflect1(ph - 1, g.H), w = reflect1(pw - 1, g.W);
        const uint4 val = *reinterpret_cast<const uint4*>(yn + ((size_t)((d + 1) * PH + h + 1) * PW + w + 1) * C + c8 * 8);
        *reinterpret_cast<uint4*>(yn + ((size_t)(pd * PH + ph) * PW + pw) * C + c8 * 8) = val;
    }
}

// ------------------------------------------------------------------ reflect-pad fold as a pass of its own (fast path)
// dy (padded by 1, REFLECT) is folded IN PLACE: every interior voxel of the padded tensor that has a coordinate in {1, S-2}
// receives the halo positions that mirror onto it.  Halo cells are only read and shell cells only written, so there is no
// hazard.  One thread per (face voxel, 8-channel group); a voxel on several faces is handled by the first face that owns it
// (d faces, then h, then w).  After this pass both backward passes read the INTERIOR of dy only, with x walked linearly --
// the layout of the un-padded variant, which runs at 93 % of the HBM copy bandwidth (the mirrored / divergent shell accesses
// held the padded variant at 57 %: profiles/r02_instnorm_fold_call16.txt).
template <typename T>
__global__ void __launch_bounds__(256) in_fold_inplace_kernel(T* __restrict__ dy, Geo g) {
    const int C = g.C, cg = C / 8, D = g.D, H = g.H, W = g.W;
    const int PH = H + 2, PW = W + 2;
    const long long per_face[3] = {(long long)H * W, (long long)D * W, (long long)D * H};
    const long long total = 2 * (per_face[0] + per_face[1] + per_face[2]) * cg;
    T* dyn = dy + (size_t)blockIdx.y * (D + 2) * PH * PW * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cg);
        long long v = i / cg;
        int d, h, w;
        if (v < 2 * per_face[0]) {                       // d faces
            d = v < per_face[0] ? 1 : D - 2;
            v %= per_face[0];
            h = (int)(v / W); w = (int)(v % W);
        } else if (v < 2 * (per_face[0] + per_face[1])) {   // h faces (voxels not on a d face)
            v -= 2 * per_face[0];
            h = v < per_face[1] ? 1 : H - 2;
            v %= per_face[1];
            d = (int)(v / W); w = (int)(v % W);
            if (d == 1 || d == D - 2) continue;
        } else {                                         // w faces (voxels on neither a d nor an h face)
            v -= 2 * (per_face[0] + per_face[1]);
            w = v < per_face[2] ? 1 : W - 2;
            v %= per_face[2];
            d = (int)(v / H); h = (int)(v % H);
            if (d == 1 || d == D - 2 || h == 1 || h == H - 2) continue;
        }
        float gy[8];
        T* cell = dyn + ((size_t)((d + 1) * PH + h + 1) * PW + w + 1) * C + c8 * 8;
        load8<T>(cell, gy);
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
                    float t[8];
                    load8<T>(dyn + ((size_t)(dd[i0] * PH + hh[i1]) * PW + ww[i2]) * C + c8 * 8, t);
#pragma unroll
                    for (int k = 0; k < 8; k++) gy[k] += t[k];
                }
        store8<T>(cell, gy);
    }
}

// sums of g' = dy*act'(z) and g'*(x - mu) over the INTERIOR of an already folded padded gradient; x is walked linearly
// (InstanceNorm -> ReLU -> ReflectionPadding3D, bf16, no dropout).  Same partial layout as in_bwd_partial_kernel.
template <typename T>
__global__ void __launch_bounds__(NT, 2) in_bwd_partial_folded_kernel(const T* __restrict__ dy, const T* __restrict__ x, Geo g,
                                                                   BwdArgs a, float* __restrict__ partial) {
    extern __shared__ float sm[];
    const int n = first_pass_sample(), C = g.C, cg = C / 8;
    const int c8 = threadIdx.x % cg, vl = threadIdx.x / cg, nvl = blockDim.x / cg;
    const int PD = g.D + 2, PH = g.H + 2, PW = g.W + 2;
    const int V = g.D * g.H * g.W;
    float mu[8], sc[8], sh[8], s1[8], s2[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int c = c8 * 8 + k;
        mu[k] = a.mean[n * C + c];
        sc[k] = a.gamma[c] * a.rstd[n * C + c];
        sh[k] = a.beta[c] - mu[k] * sc[k];
        s1[k] = s2[k] = 0.f;
    }
    const T* xn = x + (size_t)n * V * C + c8 * 8;
    const T* dyn = dy + (size_t)n * PD * PH * PW * C + c8 * 8;
    const int S = gridDim.x * nvl;
    VoxIter it;
    it.init(blockIdx.x * nvl + vl, S, g.H, g.W);
    for (int v = blockIdx.x * nvl + vl; v < V; v += U * S) {
        Raw<T> rx[U], rg[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (v + u * S < V) {
                load_raw<T>(xn + (size_t)(v + u * S) * C, rx[u]);
                load_raw<T>(dyn + (size_t)(((it.pd + 1) * PH + it.ph + 1) * PW + it.pw + 1) * C, rg[u]);
            }
            it.next();
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            if (v + u * S < V) {
                float f[8], gy[8];
                unpack_raw(rx[u], f);
                unpack_raw(rg[u], gy);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    float gg = fmaf(f[k], sc[k], sh[k]) > 0.f ? gy[k] : 0.f;
                    s1[k] += gg;
                    s2[k] = fmaf(gg, f[k] - mu[k], s2[k]);
                }
            }
    }
    float* row = sm + ((size_t)vl * C + c8 * 8) * 2;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        row[2 * k] = s1[k];
        row[2 * k + 1] = s2[k] * a.rstd[n * C + c8 * 8 + k];
    }
    __syncthreads();
    float* out = partial + ((size_t)n * gridDim.x + blockIdx.x) * C * 2;
    for (int i = threadIdx.x; i < C * 2; i += blockDim.x) {
        float acc = 0.f;
        for (int l = 0; l < nvl; l++) acc += sm[(size_t)l * C * 2 + i];
        out[i] = acc;
    }
}

inline int block_threads(int C) { int cg = C / 8; return (NT / cg) * cg; }
// blocks per sample for a pass over `vox` voxels: ~8 waves of 148 SMs over the whole batch, at least U voxels per lane
inline int pick_grid(long long vox, int N, int C) {
    long long nvl = block_threads(C) / (C / 8);
    // ~8 k voxels per block: between one resident round (2 blocks per SM) for small batches -- a block's prologue (per-channel
    // constants) and epilogue (block reduction) cost ~3 us, as much as streaming 4 k voxels -- and 8 waves for the large ones
    long long total = vox * N / 8192;
    if (total < 148 * 2) total = 148 * 2;
    if (total > 148 * 8) total = 148 * 8;
    long long want = (total + N - 1) / N;
    long long maxb = (vox + nvl * U - 1) / (nvl * U);
    if (want > maxb) want = maxb;
    return want < 1 ? 1 : (int)want;
}

template <typename T>
int stats_impl(const T* x, int N, int D, int H, int W, int C, float* mean, float* rstd, void* ws, size_t ws_bytes, cudaStream_t st,
               int relu_in) {
    const long long V = (long long)D * H * W;
    if (V * C >= (1LL << 31)) return VG_ERR_UNSUPPORTED;
    const int nblk = pick_grid(V, N, C), nthr = block_threads(C);
    size_t need = (size_t)N * nblk * C * 2 * sizeof(float);
    if (ws_bytes < need) return VG_ERR_WORKSPACE;
    size_t smem = (size_t)(nthr / (C / 8)) * C * 2 * sizeof(float);
    in_stats_partial_kernel<T><<<dim3(nblk, N), nthr, smem, st>>>(x, (int)V, C, (float*)ws, relu_in); VG_LAUNCHED(1);
    in_stats_final_kernel<T><<<vg_cdiv(N * C, 8), 256, 0, st>>>(x, (const float*)ws, (size_t)V, C, nblk, N, mean, rstd, relu_in); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

}  // namespace

extern "C" {

size_t vg_instnorm_workspace_bytes(int N, int D, int H, int W, int C) {
    // partials of the widest pass (the backward reduction walks the padded volume: up to 3 + 3 voxels of padding per axis, the zeros
    // of a 7^3 'same' convolution) + per-(n,c) sums
    size_t nblk = (size_t)pick_grid((long long)(D + 6) * (H + 6) * (W + 6), N, C);
    return (size_t)N * nblk * C * 2 * sizeof(float) + (size_t)N * C * 2 * sizeof(float);
}

int vg_instnorm_stats(const void* x, int dtype, int N, int D, int H, int W, int C, float* mean, float* rstd, void* ws,
                      size_t ws_bytes, void* stream) {
    VG_REQUIRE(x && mean && rstd && ws && N > 0 && C % 8 == 0 && C >= 8 && C <= 8 * NT);
    cudaStream_t st = (cudaStream_t)stream;
    const int relu_in = (dtype & VG_IN_RELU_INPUT) ? 1 : 0;
    dtype &= ~VG_IN_RELU_INPUT;
    if (dtype == VG_BF16) return stats_impl<bf16>((const bf16*)x, N, D, H, W, C, mean, rstd, ws, ws_bytes, st, relu_in);
    if (dtype == VG_F32) return stats_impl<float>((const float*)x, N, D, H, W, C, mean, rstd, ws, ws_bytes, st, relu_in);
    return VG_ERR_INVALID;
}

int vg_instnorm_apply(const vg_instnorm_desc* d, const void* x, const void* residual, void* y, const float* mean,
                      const float* rstd, const float* gamma, const float* beta, const float* drop, const float* noise,
                      void* stream) {
    VG_REQUIRE(d && x && y && mean && rstd && gamma && beta);
    VG_REQUIRE(d->C % 8 == 0 && d->C <= 8 * NT && d->pad_lo >= 0 && d->pad_hi >= 0 && d->pad_lo <= 3 && d->pad_hi <= 3);
    if (d->pad_mode == VG_PAD_REFLECT && (d->pad_lo || d->pad_hi))
        VG_REQUIRE(d->pad_lo == 1 && d->pad_hi == 1 && d->D >= 2 && d->H >= 2 && d->W >= 2);
    const int dtype = d->dtype & ~(VG_IN_RELU_INPUT | VG_IN_BATCH_STATS | VG_IN_DY_SCRATCH);
    Geo g{d->N, d->D, d->H, d->W, d->C, d->pad_lo, d->pad_hi, d->pad_mode, (d->dtype & VG_IN_RELU_INPUT) ? 1 : 0};
    ApplyArgs a{mean, rstd, gamma, beta, drop, noise, d->slope, d->noise_std, d->act, d->seed, d->seed_dev};
    const int pp = d->pad_lo + d->pad_hi;
    const long long M = (long long)(d->D + pp) * (d->H + pp) * (d->W + pp);
    VG_REQUIRE(M * d->C < (1LL << 31));
    dim3 grid(pick_grid(M, d->N, d->C), d->N);
    const int nthr = block_threads(d->C);
    cudaStream_t st = (cudaStream_t)stream;
    int sp = 0;
    if (dtype == VG_BF16 && !noise && !(d->noise_std > 0.f) && !drop && !g.relu_in) {
        if (d->act == VG_ACT_RELU && d->pad_lo == 1 && d->pad_hi == 1 && d->pad_mode == VG_PAD_REFLECT && !residual) sp = 1;
        else if (d->act == VG_ACT_NONE && d->pad_lo == 0 && d->pad_hi == 0 && residual) sp = 2;
    }
    static int fold_on = -1;   // VG_IN_FOLD=0: one pass over the padded output with mirrored reads (A/B testing)
    if (fold_on < 0) {
        const char* e = getenv("VG_IN_FOLD");
        fold_on = (e && e[0] == '0') ? 0 : 1;
    }
    if (sp == 1 && fold_on && d->D >= 2 && d->H >= 2 && d->W >= 2) {
        const long long V = (long long)d->D * d->H * d->W;
        const long long halo = 2LL * ((long long)(d->H + 2) * (d->W + 2) + (long long)d->D * (d->W + 2) + (long long)d->D * d->H) * (d->C / 8);
        in_apply_interior_kernel<bf16><<<dim3(pick_grid(V, d->N, d->C), d->N), nthr, 0, st>>>((const bf16*)x, (bf16*)y, g, a);
        pad_halo_inplace_kernel<bf16><<<dim3(vg_cdiv(halo, 256), d->N), 256, 0, st>>>((bf16*)y, g);
        VG_LAUNCHED(2);
    } else if (sp == 1) {
        in_apply_sp_kernel<bf16, 1><<<grid, nthr, 0, st>>>((const bf16*)x, (const bf16*)residual, (bf16*)y, g, a); VG_LAUNCHED(1);
    } else if (sp == 2) {
        in_apply_sp_kernel<bf16, 2><<<grid, nthr, 0, st>>>((const bf16*)x, (const bf16*)residual, (bf16*)y, g, a); VG_LAUNCHED(1);
    } else if (dtype == VG_BF16) {
        in_apply_kernel<bf16><<<grid, nthr, 0, st>>>((const bf16*)x, (const bf16*)residual, (bf16*)y, g, a); VG_LAUNCHED(1);
    } else if (dtype == VG_F32) {
        in_apply_kernel<float><<<grid, nthr, 0, st>>>((const float*)x, (const float*)residual, (float*)y, g, a); VG_LAUNCHED(1);
    } else {
        return VG_ERR_INVALID;
    }
    VG_CHECK_LAUNCH();
    return VG_OK;
}

// dy: gradient w.r.t. the apply output in ITS (padded) layout; dx: gradient w.r.t. x; dres (optional):
// gradient w.r.t. the residual input; dgamma/dbeta (optional) are accumulated (+=).
int vg_instnorm_bwd(const vg_instnorm_desc* d, const void* dy, const void* x, const float* mean, const float* rstd,
                    const float* gamma, const float* beta, const float* drop, void* dx, int accumulate_dx, void* dres,
                    float* dgamma, float* dbeta, void* ws, size_t ws_bytes, void* stream) {
    return vg_instnorm_bwd_sinks(d, dy, x, mean, rstd, gamma, beta, drop, dx, accumulate_dx, dres, dgamma, dbeta, nullptr, nullptr, ws,
                                 ws_bytes, stream);
}

// vg_instnorm_bwd + the bias gradients of the convolutions that produced x / the residual: dbias_x[c] += sum of the stored dx,
// dbias_res[c] += sum of the stored dres (either may be NULL).  Not combinable with accumulate_dx (the sum must be of the
// producer's whole gradient).
int vg_instnorm_bwd_sinks(const vg_instnorm_desc* d, const void* dy, const void* x, const float* mean, const float* rstd,
                          const float* gamma, const float* beta, const float* drop, void* dx, int accumulate_dx, void* dres,
                          float* dgamma, float* dbeta, float* dbias_x, float* dbias_res, void* ws, size_t ws_bytes, void* stream) {
    VG_REQUIRE(d && dy && x && mean && rstd && gamma && beta && dx && ws);
    VG_REQUIRE(!(dbias_x && accumulate_dx) && !(dbias_res && !dres));
    const Sinks sk{dbias_x, dbias_res};
    VG_REQUIRE(d->C % 8 == 0 && d->C <= 8 * NT);
    const int dtype = d->dtype & ~(VG_IN_RELU_INPUT | VG_IN_BATCH_STATS | VG_IN_DY_SCRATCH);
    const int batch = (d->dtype & VG_IN_BATCH_STATS) ? 1 : 0;
    Geo g{d->N, d->D, d->H, d->W, d->C, d->pad_lo, d->pad_hi, d->pad_mode, (d->dtype & VG_IN_RELU_INPUT) ? 1 : 0};
    BwdArgs a{mean, rstd, gamma, beta, drop, d->slope, d->act};
    const int pp = d->pad_lo + d->pad_hi;
    const long long M = (long long)(d->D + pp) * (d->H + pp) * (d->W + pp);
    VG_REQUIRE(M * d->C < (1LL << 31));
    const int nblk = pick_grid(M, d->N, d->C);
    size_t need_p = (size_t)d->N * nblk * d->C * 2 * sizeof(float), need_s = (size_t)d->N * d->C * 2 * sizeof(float);
    if (ws_bytes < need_p + need_s) return VG_ERR_WORKSPACE;
    float* partial = (float*)ws;
    float* sums = (float*)((char*)ws + need_p);
    const int nthr = block_threads(d->C);
    size_t smem = (size_t)(nthr / (d->C / 8)) * d->C * 2 * sizeof(float);
    const size_t sink_smem = (dbias_x || dbias_res) ? (size_t)(nthr / (d->C / 8)) * d->C * sizeof(float) : 0;
    dim3 grid2(pick_grid((long long)d->D * d->H * d->W, d->N, d->C), d->N);
    cudaStream_t st = (cudaStream_t)stream;
    int sp = 0;
    if (dtype == VG_BF16 && !drop && !g.relu_in) {
        if (d->act == VG_ACT_RELU && d->pad_lo == 1 && d->pad_hi == 1 && d->pad_mode == VG_PAD_REFLECT && !dres) sp = 1;
        else if (d->act == VG_ACT_NONE && d->pad_lo == 0 && d->pad_hi == 0 && dres) sp = 2;
    }
    static int fold_on = -1;   // VG_IN_FOLD=0: keep the fold inside the two passes (A/B testing)
    if (fold_on < 0) {
        const char* e = getenv("VG_IN_FOLD");
        fold_on = (e && e[0] == '0') ? 0 : 1;
    }
    const bool can_fold = fold_on && (d->dtype & VG_IN_DY_SCRATCH) && dtype == VG_BF16 && d->pad_mode == VG_PAD_REFLECT && d->pad_lo == 1 &&
                          d->pad_hi == 1 && d->D >= 4 && d->H >= 4 && d->W >= 4;
    if (sp != 1 && can_fold) {
        // generic options (LeakyReLU / dropout / residual gradient: the discriminator's norms): fold once, then the generic passes
        // treat the halo as zero padding (no mirrored reads, no shell branch)
        const long long faces = 2LL * ((long long)d->H * d->W + (long long)d->D * d->W + (long long)d->D * d->H) * (d->C / 8);
        in_fold_inplace_kernel<bf16><<<dim3(vg_cdiv(faces, 256), d->N), 256, 0, st>>>((bf16*)const_cast<void*>(dy), g); VG_LAUNCHED(1);
        g.pad_mode = VG_PAD_ZERO;
    }
    if (sp == 1 && can_fold) {
        // the caller lets dy be overwritten: fold its reflected halo into the interior once, then two linear passes
        const long long faces = 2LL * ((long long)d->H * d->W + (long long)d->D * d->W + (long long)d->D * d->H) * (d->C / 8);
        const int nblk2 = pick_grid((long long)d->D * d->H * d->W, d->N, d->C);
        in_fold_inplace_kernel<bf16><<<dim3(vg_cdiv(faces, 256), d->N), 256, 0, st>>>((bf16*)const_cast<void*>(dy), g);
        in_bwd_partial_folded_kernel<bf16><<<dim3(nblk2, d->N), nthr, smem, st>>>((const bf16*)dy, (const bf16*)x, g, a, partial);
        in_bwd_final_kernel<<<vg_cdiv(d->N * d->C, 8), 256, 0, st>>>(partial, nblk2, d->N, d->C, sums, dgamma, dbeta, batch);
        in_bwd_apply_sp_kernel<bf16, 1, true><<<grid2, nthr, sink_smem, st>>>((const bf16*)dy, (const bf16*)x, g, a, sums, (bf16*)dx, (bf16*)dres, accumulate_dx, sk);
        VG_LAUNCHED(4);
    } else if (sp == 1) {
        in_bwd_partial_sp_kernel<bf16, 1><<<dim3(nblk, d->N), nthr, smem, st>>>((const bf16*)dy, (const bf16*)x, g, a, partial);
        in_bwd_final_kernel<<<vg_cdiv(d->N * d->C, 8), 256, 0, st>>>(partial, nblk, d->N, d->C, sums, dgamma, dbeta, batch);
        in_bwd_apply_sp_kernel<bf16, 1, false><<<grid2, nthr, sink_smem, st>>>((const bf16*)dy, (const bf16*)x, g, a, sums, (bf16*)dx, (bf16*)dres, accumulate_dx, sk);
        VG_LAUNCHED(3);
    } else if (sp == 2) {
        in_bwd_partial_sp_kernel<bf16, 2><<<dim3(nblk, d->N), nthr, smem, st>>>((const bf16*)dy, (const bf16*)x, g, a, partial);
        in_bwd_final_kernel<<<vg_cdiv(d->N * d->C, 8), 256, 0, st>>>(partial, nblk, d->N, d->C, sums, dgamma, dbeta, batch);
        in_bwd_apply_sp_kernel<bf16, 2, false><<<grid2, nthr, sink_smem, st>>>((const bf16*)dy, (const bf16*)x, g, a, sums, (bf16*)dx, (bf16*)dres, accumulate_dx, sk);
        VG_LAUNCHED(3);
    } else if (dtype == VG_BF16) {
        in_bwd_partial_kernel<bf16><<<dim3(nblk, d->N), nthr, smem, st>>>((const bf16*)dy, (const bf16*)x, g, a, partial); VG_LAUNCHED(1);
        in_bwd_final_kernel<<<vg_cdiv(d->N * d->C, 8), 256, 0, st>>>(partial, nblk, d->N, d->C, sums, dgamma, dbeta, batch); VG_LAUNCHED(1);
        in_bwd_apply_kernel<bf16><<<grid2, nthr, sink_smem, st>>>((const bf16*)dy, (const bf16*)x, g, a, sums, (bf16*)dx, (bf16*)dres,
                                                                 accumulate_dx, sk); VG_LAUNCHED(1);
    } else if (dtype == VG_F32) {
        in_bwd_partial_kernel<float><<<dim3(nblk, d->N), nthr, smem, st>>>((const float*)dy, (const float*)x, g, a, partial); VG_LAUNCHED(1);
        in_bwd_final_kernel<<<vg_cdiv(d->N * d->C, 8), 256, 0, st>>>(partial, nblk, d->N, d->C, sums, dgamma, dbeta, batch); VG_LAUNCHED(1);
        in_bwd_apply_kernel<float><<<grid2, nthr, sink_smem, st>>>((const float*)dy, (const float*)x, g, a, sums, (float*)dx, (float*)dres,
                                                                  accumulate_dx, sk); VG_LAUNCHED(1);
    } else {
        return VG_ERR_INVALID;
    }
    VG_CHECK_LAUNCH();
    return VG_OK;
}

// BatchNormalization (vnet_model.py:127-128,142-143: Keras defaults axis=-1, momentum 0.99, epsilon 1e-3, per-replica batch
// statistics).  The normalisation itself is the InstanceNorm arithmetic with statistics taken over the whole local batch, so the
// apply / backward passes are vg_instnorm_apply / vg_instnorm_bwd with (n, c)-replicated statistics and, for the backward,
// VG_IN_BATCH_STATS set in desc.dtype (reductions over N*D*H*W).  mean_nc / rstd_nc: [N*C]; moving_*: [C] (may be NULL when training);
// ws: vg_instnorm_workspace_bytes(1, N*D, H, W, C) + 2*C floats.
int vg_batchnorm_stats(const void* x, int dtype, int N, int D, int H, int W, int C, float* mean_nc, float* rstd_nc, float* moving_mean,
                       float* moving_var, float momentum, int training, void* ws, size_t ws_bytes, void* stream) {
    VG_REQUIRE(x && mean_nc && rstd_nc && N > 0 && C % 8 == 0 && (training || (moving_mean && moving_var)));
    float* mc = nullptr;
    float* rc = nullptr;
    const double cnt = (double)N * D * H * W;
    if (training) {
        VG_REQUIRE(ws && ws_bytes >= (size_t)2 * C * sizeof(float));
        const size_t tail = (size_t)2 * C * sizeof(float);
        mc = (float*)((char*)ws + ws_bytes - tail);
        rc = mc + C;
        int rc_ = vg_instnorm_stats(x, dtype, 1, N * D, H, W, C, mc, rc, ws, ws_bytes - tail, stream);
        if (rc_ != VG_OK) return rc_;
    }
    bn_expand_kernel<<<vg_cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(mc, rc, N, C, mean_nc, rstd_nc, moving_mean, moving_var, momentum,
                                                                      training, cnt > 1.0 ? (float)(cnt / (cnt - 1.0)) : 1.f); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

size_t vg_batchnorm_workspace_bytes(int N, int D, int H, int W, int C) {
    const size_t a = vg_instnorm_workspace_bytes(1, N * D, H, W, C), b = vg_instnorm_workspace_bytes(N, D, H, W, C);
    return (a > b ? a : b) + (size_t)2 * C * sizeof(float);
}

int vg_batchnorm_bwd(const vg_instnorm_desc* d, const void* dy, const void* x, const float* mean_nc, const float* rstd_nc,
                     const float* gamma, const float* beta, const float* drop, void* dx, int accumulate_dx, void* dres, float* dgamma,
                     float* dbeta, void* ws, size_t ws_bytes, void* stream) {
    VG_REQUIRE(d);
    vg_instnorm_desc e = *d;
    e.dtype |= VG_IN_BATCH_STATS;
    return vg_instnorm_bwd(&e, dy, x, mean_nc, rstd_nc, gamma, beta, drop, dx, accumulate_dx, dres, dgamma, dbeta, ws, ws_bytes, stream);
}

}  // extern "C"
