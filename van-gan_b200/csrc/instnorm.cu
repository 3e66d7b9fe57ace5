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
#include "common.cuh"

namespace {

constexpr int NT = 256;
constexpr float IN_EPS = 1e-3f;

struct Geo {
    int N, D, H, W, C;
    int pad_lo, pad_hi, pad_mode;  // output (fwd) / incoming-gradient (bwd) padding
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

// ------------------------------------------------------------------ statistics
// partial[(n*nblk + blk)*C*2 + c*2 + {0,1}] = sum / sumsq of (x - shift_c) over the block's voxels,
// shift_c = x[n,0,c] (keeps E[x^2]-E[x]^2 well conditioned)
template <typename T>
__global__ void __launch_bounds__(NT) in_stats_partial_kernel(const T* __restrict__ x, size_t V, int C, int nblk,
                                                              float* __restrict__ partial) {
    extern __shared__ float sm[];  // [nvl][C][2]
    int n = blockIdx.y, blk = blockIdx.x;
    int cg = C / 8, nvl = NT / cg;
    int g = threadIdx.x % cg, vl = threadIdx.x / cg;
    const T* xn = x + (size_t)n * V * C;
    float shift[8], s1[8], s2[8];
    load8<T>(xn + g * 8, shift);
#pragma unroll
    for (int k = 0; k < 8; k++) s1[k] = s2[k] = 0.f;
    size_t per = (V + nblk - 1) / nblk;
    size_t v0 = (size_t)blk * per, v1 = v0 + per < V ? v0 + per : V;
    if (vl < nvl) {
        for (size_t v = v0 + vl; v < v1; v += nvl) {
            float f[8];
            load8<T>(xn + v * C + g * 8, f);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                float d = f[k] - shift[k];
                s1[k] += d;
                s2[k] += d * d;
            }
        }
        float* row = sm + ((size_t)vl * C + g * 8) * 2;
#pragma unroll
        for (int k = 0; k < 8; k++) { row[2 * k] = s1[k]; row[2 * k + 1] = s2[k]; }
    }
    __syncthreads();
    float* out = partial + ((size_t)n * nblk + blk) * C * 2;
    for (int i = threadIdx.x; i < C * 2; i += NT) {
        float a = 0.f;
        for (int l = 0; l < nvl; l++) a += sm[(size_t)l * C * 2 + i];
        out[i] = a;
    }
}

// one warp per (n,c): lanes stride over the block partials, shuffle-reduce in double
template <typename T>
__global__ void __launch_bounds__(256) in_stats_final_kernel(const T* __restrict__ x, const float* __restrict__ partial,
                                                             size_t V, int C, int nblk, int N, float* __restrict__ mean,
                                                             float* __restrict__ rstd) {
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
};

template <typename T>
__global__ void __launch_bounds__(NT) in_apply_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y,
                                                      Geo g, ApplyArgs a) {
    int cg = g.C / 8;
    int PD = g.D + g.pad_lo + g.pad_hi, PH = g.H + g.pad_lo + g.pad_hi, PW = g.W + g.pad_lo + g.pad_hi;
    size_t total = (size_t)g.N * PD * PH * PW * cg;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        int c8 = (int)(i % cg);
        size_t pv = i / cg;
        int pw = (int)(pv % PW), ph = (int)((pv / PW) % PH), pd = (int)((pv / ((size_t)PW * PH)) % PD);
        int n = (int)(pv / ((size_t)PW * PH * PD));
        int d = pd - g.pad_lo, h = ph - g.pad_lo, w = pw - g.pad_lo;
        bool oob = (unsigned)d >= (unsigned)g.D || (unsigned)h >= (unsigned)g.H || (unsigned)w >= (unsigned)g.W;
        float o[8];
        if (oob && g.pad_mode == VG_PAD_ZERO) {
#pragma unroll
            for (int k = 0; k < 8; k++) o[k] = 0.f;
            store8<T>(y + i * 8, o);
            continue;
        }
        if (oob) { d = reflect1(d, g.D); h = reflect1(h, g.H); w = reflect1(w, g.W); }
        size_t src = ((((size_t)n * g.D + d) * g.H + h) * g.W + w) * g.C + c8 * 8;
        float f[8];
        load8<T>(x + src, f);
        int sc = n * g.C + c8 * 8;
        float r[8];
        if (res) load8<T>(res + src, r);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            float scale = a.gamma[c8 * 8 + k] * a.rstd[sc + k];
            float z = (f[k] - a.mean[sc + k]) * scale + a.beta[c8 * 8 + k];
            float v = act_fwd(z, a.act, a.slope);
            if (a.drop) v *= a.drop[sc + k];
            if (res) v += r[k];
            o[k] = v;
        }
        if (a.noise) {
            // explicit noise tensor: padded layout for REFLECT (noise is added after the pad layer),
            // unpadded layout for ZERO ('same' convs pad after the noise layer)
            size_t ni = g.pad_mode == VG_PAD_REFLECT ? i * 8 : src;
#pragma unroll
            for (int k = 0; k < 8; k++) o[k] += a.noise[ni + k];
        } else if (a.noise_std > 0.f) {
            uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
            uint4 r0 = philox4x32(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 0u, 0x56414e47u), key);
            uint4 r1 = philox4x32(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 1u, 0x56414e47u), key);
            float2 n0 = box_muller(r0.x, r0.y), n1 = box_muller(r0.z, r0.w), n2 = box_muller(r1.x, r1.y),
                   n3 = box_muller(r1.z, r1.w);
            o[0] += a.noise_std * n0.x; o[1] += a.noise_std * n0.y; o[2] += a.noise_std * n1.x; o[3] += a.noise_std * n1.y;
            o[4] += a.noise_std * n2.x; o[5] += a.noise_std * n2.y; o[6] += a.noise_std * n3.x; o[7] += a.noise_std * n3.y;
        }
        store8<T>(y + i * 8, o);
    }
}

// ------------------------------------------------------------------ backward
// gradient w.r.t. the (unpadded) apply output = incoming gradient in padded layout folded back:
// REFLECT: every padded position whose mirror is this voxel; ZERO: the interior only.
template <typename T>
__device__ __forceinline__ void load_folded(const T* __restrict__ dy, const Geo& g, int n, int d, int h, int w, int c8,
                                            float* out) {
    int PH = g.H + g.pad_lo + g.pad_hi, PW = g.W + g.pad_lo + g.pad_hi, PD = g.D + g.pad_lo + g.pad_hi;
    if (g.pad_lo == 0 && g.pad_hi == 0) {
        load8<T>(dy + ((((size_t)n * g.D + d) * g.H + h) * g.W + w) * g.C + c8 * 8, out);
        return;
    }
    if (g.pad_mode == VG_PAD_ZERO) {
        load8<T>(dy + ((((size_t)n * PD + d + g.pad_lo) * PH + h + g.pad_lo) * PW + w + g.pad_lo) * g.C + c8 * 8, out);
        return;
    }
    // REFLECT, pad 1 each side
    int dd[2], hh[2], ww[2], nd = 1, nh = 1, nw = 1;
    dd[0] = d + 1; hh[0] = h + 1; ww[0] = w + 1;
    // (for S==3 a voxel can be the mirror of both borders; handled by the two independent tests)
    int dd2[3], hh2[3], ww2[3];
    dd2[0] = d + 1; nd = 1; if (d == 1) dd2[nd++] = 0; if (d == g.D - 2) dd2[nd++] = g.D + 1;
    hh2[0] = h + 1; nh = 1; if (h == 1) hh2[nh++] = 0; if (h == g.H - 2) hh2[nh++] = g.H + 1;
    ww2[0] = w + 1; nw = 1; if (w == 1) ww2[nw++] = 0; if (w == g.W - 2) ww2[nw++] = g.W + 1;
    (void)dd; (void)hh; (void)ww;
#pragma unroll
    for (int k = 0; k < 8; k++) out[k] = 0.f;
    for (int a = 0; a < nd; a++)
        for (int b = 0; b < nh; b++)
            for (int c = 0; c < nw; c++) {
                float f[8];
                load8<T>(dy + ((((size_t)n * PD + dd2[a]) * PH + hh2[b]) * PW + ww2[c]) * g.C + c8 * 8, f);
#pragma unroll
                for (int k = 0; k < 8; k++) out[k] += f[k];
            }
}

struct BwdArgs {
    const float *mean, *rstd, *gamma, *beta, *drop;
    float slope;
    int act;
};

// partial[(n*nblk+blk)*C*2 + c*2 + {0,1}] = sum g, sum g*xhat  with g = fold(dy)*drop*act'(z)
template <typename T>
__global__ void __launch_bounds__(NT) in_bwd_partial_kernel(const T* __restrict__ dy, const T* __restrict__ x, Geo g,
                                                            BwdArgs a, int nblk, float* __restrict__ partial) {
    extern __shared__ float sm[];
    int n = blockIdx.y, blk = blockIdx.x;
    int C = g.C, cg = C / 8, nvl = NT / cg;
    int c8 = threadIdx.x % cg, vl = threadIdx.x / cg;
    size_t V = (size_t)g.D * g.H * g.W;
    float s1[8], s2[8];
#pragma unroll
    for (int k = 0; k < 8; k++) s1[k] = s2[k] = 0.f;
    size_t per = (V + nblk - 1) / nblk;
    size_t v0 = (size_t)blk * per, v1 = v0 + per < V ? v0 + per : V;
    if (vl < nvl) {
        float mu[8], rs[8], ga[8], be[8], dr[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            mu[k] = a.mean[n * C + c8 * 8 + k]; rs[k] = a.rstd[n * C + c8 * 8 + k];
            ga[k] = a.gamma[c8 * 8 + k]; be[k] = a.beta[c8 * 8 + k];
            dr[k] = a.drop ? a.drop[n * C + c8 * 8 + k] : 1.f;
        }
        for (size_t v = v0 + vl; v < v1; v += nvl) {
            int w = (int)(v % g.W), h = (int)((v / g.W) % g.H), d = (int)(v / ((size_t)g.W * g.H));
            float f[8], gy[8];
            load8<T>(x + ((size_t)n * V + v) * C + c8 * 8, f);
            load_folded<T>(dy, g, n, d, h, w, c8, gy);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                float xh = (f[k] - mu[k]) * rs[k];
                float z = xh * ga[k] + be[k];
                float gg = gy[k] * dr[k] * act_grad(z, a.act, a.slope);
                s1[k] += gg;
                s2[k] += gg * xh;
            }
        }
        float* row = sm + ((size_t)vl * C + c8 * 8) * 2;
#pragma unroll
        for (int k = 0; k < 8; k++) { row[2 * k] = s1[k]; row[2 * k + 1] = s2[k]; }
    }
    __syncthreads();
    float* out = partial + ((size_t)n * nblk + blk) * C * 2;
    for (int i = threadIdx.x; i < C * 2; i += NT) {
        float acc = 0.f;
        for (int l = 0; l < nvl; l++) acc += sm[(size_t)l * C * 2 + i];
        out[i] = acc;
    }
}

// sums[(n*C+c)*2+{0,1}] = S1, S2; dgamma[c] += S2, dbeta[c] += S1  (one warp per (n,c); N <= a few atomics per channel)
__global__ void __launch_bounds__(256) in_bwd_final_kernel(const float* __restrict__ partial, int nblk, int N, int C,
                                                           float* __restrict__ sums, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta) {
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
        sums[(size_t)i * 2] = (float)s1;
        sums[(size_t)i * 2 + 1] = (float)s2;
        if (dgamma) atomicAdd(dgamma + c, (float)s2);
        if (dbeta) atomicAdd(dbeta + c, (float)s1);
    }
}

// dx = gamma*rstd*(g - S1/V - xhat*S2/V); dres = fold(dy) (optional)
template <typename T>
__global__ void __launch_bounds__(NT) in_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, Geo g, BwdArgs a,
                                                          const float* __restrict__ sums, T* __restrict__ dx,
                                                          T* __restrict__ dres, int accumulate_dx) {
    int C = g.C, cg = C / 8;
    size_t V = (size_t)g.D * g.H * g.W;
    size_t total = (size_t)g.N * V * cg;
    float invV = 1.f / (float)V;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        int c8 = (int)(i % cg);
        size_t nv = i / cg;
        size_t v = nv % V;
        int n = (int)(nv / V);
        int w = (int)(v % g.W), h = (int)((v / g.W) % g.H), d = (int)(v / ((size_t)g.W * g.H));
        float f[8], gy[8], o[8];
        load8<T>(x + i * 8, f);
        load_folded<T>(dy, g, n, d, h, w, c8, gy);
        if (dres) store8<T>(dres + i * 8, gy);
        if (accumulate_dx) load8<T>(dx + i * 8, o);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            int sc = n * C + c8 * 8 + k;
            float rs = a.rstd[sc], ga = a.gamma[c8 * 8 + k];
            float xh = (f[k] - a.mean[sc]) * rs;
            float z = xh * ga + a.beta[c8 * 8 + k];
            float gg = gy[k] * (a.drop ? a.drop[sc] : 1.f) * act_grad(z, a.act, a.slope);
            float val = ga * rs * (gg - sums[2 * sc] * invV - xh * sums[2 * sc + 1] * invV);
            o[k] = accumulate_dx ? o[k] + val : val;
        }
        store8<T>(dx + i * 8, o);
    }
}

inline int pick_nblk(size_t V, int N) {
    long long want = (148LL * 6 + N - 1) / N;
    long long maxb = (long long)((V + 255) / 256);
    if (want > maxb) want = maxb;
    if (want < 1) want = 1;
    return (int)want;
}

template <typename T>
int stats_impl(const T* x, int N, size_t V, int C, float* mean, float* rstd, void* ws, size_t ws_bytes, cudaStream_t st) {
    int nblk = pick_nblk(V, N);
    size_t need = (size_t)N * nblk * C * 2 * sizeof(float);
    if (ws_bytes < need) return VG_ERR_WORKSPACE;
    int cg = C / 8, nvl = NT / cg;
    size_t smem = (size_t)nvl * C * 2 * sizeof(float);
    in_stats_partial_kernel<T><<<dim3(nblk, N), NT, smem, st>>>(x, V, C, nblk, (float*)ws); VG_LAUNCHED(1);
    in_stats_final_kernel<T><<<vg_cdiv(N * C, 8), 256, 0, st>>>(x, (const float*)ws, V, C, nblk, N, mean, rstd); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

}  // namespace

extern "C" {

size_t vg_instnorm_workspace_bytes(int N, int D, int H, int W, int C) {
    size_t V = (size_t)D * H * W;
    return (size_t)N * pick_nblk(V, N) * C * 2 * sizeof(float) + (size_t)N * C * 2 * sizeof(float);
}

int vg_instnorm_stats(const void* x, int dtype, int N, int D, int H, int W, int C, float* mean, float* rstd, void* ws,
                      size_t ws_bytes, void* stream) {
    VG_REQUIRE(x && mean && rstd && ws && N > 0 && C % 8 == 0 && C >= 8 && C <= 8 * NT);
    size_t V = (size_t)D * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == VG_BF16) return stats_impl<bf16>((const bf16*)x, N, V, C, mean, rstd, ws, ws_bytes, st);
    if (dtype == VG_F32) return stats_impl<float>((const float*)x, N, V, C, mean, rstd, ws, ws_bytes, st);
    return VG_ERR_INVALID;
}

int vg_instnorm_apply(const vg_instnorm_desc* d, const void* x, const void* residual, void* y, const float* mean,
                      const float* rstd, const float* gamma, const float* beta, const float* drop, const float* noise,
                      void* stream) {
    VG_REQUIRE(d && x && y && mean && rstd && gamma && beta);
    VG_REQUIRE(d->C % 8 == 0 && d->pad_lo >= 0 && d->pad_hi >= 0);
    if (d->pad_mode == VG_PAD_REFLECT && (d->pad_lo || d->pad_hi))
        VG_REQUIRE(d->pad_lo == 1 && d->pad_hi == 1 && d->D >= 2 && d->H >= 2 && d->W >= 2);
    Geo g{d->N, d->D, d->H, d->W, d->C, d->pad_lo, d->pad_hi, d->pad_mode};
    ApplyArgs a{mean, rstd, gamma, beta, drop, noise, d->slope, d->noise_std, d->act, d->seed};
    size_t P = (size_t)(d->D + d->pad_lo + d->pad_hi) * (d->H + d->pad_lo + d->pad_hi) * (d->W + d->pad_lo + d->pad_hi);
    size_t total = (size_t)d->N * P * (d->C / 8);
    int grid = vg_grid_for(total, NT, 16);
    cudaStream_t st = (cudaStream_t)stream;
    if (d->dtype == VG_BF16) {
        in_apply_kernel<bf16><<<grid, NT, 0, st>>>((const bf16*)x, (const bf16*)residual, (bf16*)y, g, a); VG_LAUNCHED(1);
    }
    else if (d->dtype == VG_F32) {
        in_apply_kernel<float><<<grid, NT, 0, st>>>((const float*)x, (const float*)residual, (float*)y, g, a); VG_LAUNCHED(1);
    }
    else
        return VG_ERR_INVALID;
    VG_CHECK_LAUNCH();
    return VG_OK;
}

// dy: gradient w.r.t. the apply output in ITS (padded) layout; dx: gradient w.r.t. x; dres (optional):
// gradient w.r.t. the residual input; dgamma/dbeta (optional) are accumulated (+=).
int vg_instnorm_bwd(const vg_instnorm_desc* d, const void* dy, const void* x, const float* mean, const float* rstd,
                    const float* gamma, const float* beta, const float* drop, void* dx, int accumulate_dx, void* dres,
                    float* dgamma, float* dbeta, void* ws, size_t ws_bytes, void* stream) {
    VG_REQUIRE(d && dy && x && mean && rstd && gamma && beta && dx && ws);
    VG_REQUIRE(d->C % 8 == 0 && d->C <= 8 * NT);
    Geo g{d->N, d->D, d->H, d->W, d->C, d->pad_lo, d->pad_hi, d->pad_mode};
    BwdArgs a{mean, rstd, gamma, beta, drop, d->slope, d->act};
    size_t V = (size_t)d->D * d->H * d->W;
    int nblk = pick_nblk(V, d->N);
    size_t need_p = (size_t)d->N * nblk * d->C * 2 * sizeof(float), need_s = (size_t)d->N * d->C * 2 * sizeof(float);
    if (ws_bytes < need_p + need_s) return VG_ERR_WORKSPACE;
    float* partial = (float*)ws;
    float* sums = (float*)((char*)ws + need_p);
    int cg = d->C / 8, nvl = NT / cg;
    size_t smem = (size_t)nvl * d->C * 2 * sizeof(float);
    size_t total = (size_t)d->N * V * cg;
    int grid = vg_grid_for(total, NT, 16);
    cudaStream_t st = (cudaStream_t)stream;
    if (d->dtype == VG_BF16) {
        in_bwd_partial_kernel<bf16><<<dim3(nblk, d->N), NT, smem, st>>>((const bf16*)dy, (const bf16*)x, g, a, nblk, partial); VG_LAUNCHED(1);
        in_bwd_final_kernel<<<vg_cdiv(d->N * d->C, 8), 256, 0, st>>>(partial, nblk, d->N, d->C, sums, dgamma, dbeta); VG_LAUNCHED(1);
        in_bwd_apply_kernel<bf16><<<grid, NT, 0, st>>>((const bf16*)dy, (const bf16*)x, g, a, sums, (bf16*)dx, (bf16*)dres,
                                                      accumulate_dx); VG_LAUNCHED(1);
    } else if (d->dtype == VG_F32) {
        in_bwd_partial_kernel<float><<<dim3(nblk, d->N), NT, smem, st>>>((const float*)dy, (const float*)x, g, a, nblk, partial); VG_LAUNCHED(1);
        in_bwd_final_kernel<<<vg_cdiv(d->N * d->C, 8), 256, 0, st>>>(partial, nblk, d->N, d->C, sums, dgamma, dbeta); VG_LAUNCHED(1);
        in_bwd_apply_kernel<float><<<grid, NT, 0, st>>>((const float*)dy, (const float*)x, g, a, sums, (float*)dx, (float*)dres,
                                                       accumulate_dx); VG_LAUNCHED(1);
    } else {
        return VG_ERR_INVALID;
    }
    VG_CHECK_LAUNCH();
    return VG_OK;
}

}  // extern "C"
