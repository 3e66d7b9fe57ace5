// Cycle / SSIM / clDice / LSGAN loss arithmetic on fp32 single-channel volumes.
//
// Replaces the TF op graph behind loss_functions.py:7-22,56-68,86-117,163-226,255-322,
// clDice_func.py:83-149 and utils.py:27-48 (reference).  All kernels are HBM-bound streaming
// passes with 128-bit accesses where the layout allows; reductions go warp-shuffle -> block ->
// one double atomicAdd per block.  Scalars are combined on the host side of the ABI
// (van-gan_b200/loss_functions.py) exactly in the reference's order.
#include "common.cuh"

namespace {

constexpr int NT = 256;

__device__ __forceinline__ uint32_t enc_f(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void minmax_init_kernel(uint32_t* enc, int N) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) {
        enc[2 * i] = 0xffffffffu;  // min slot
        enc[2 * i + 1] = 0u;       // max slot
    }
}

__global__ void __launch_bounds__(NT) minmax_kernel(const float* __restrict__ x, size_t V, uint32_t* enc) {
    int n = blockIdx.y;
    const float* p = x + (size_t)n * V;
    float mn = INFINITY, mx = -INFINITY;
    size_t V4 = (V % 4 == 0) ? V / 4 : 0;  // vector path only when every sample stays 16-byte aligned
    const float4* p4 = reinterpret_cast<const float4*>(p);
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < V4; i += (size_t)gridDim.x * NT) {
        float4 v = __ldg(p4 + i);
        mn = fminf(fminf(mn, fminf(v.x, v.y)), fminf(v.z, v.w));
        mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
    if (blockIdx.x == 0)
        for (size_t i = V4 * 4 + threadIdx.x; i < V; i += NT) { mn = fminf(mn, p[i]); mx = fmaxf(mx, p[i]); }
    mn = warp_min(mn);
    mx = warp_max(mx);
    __shared__ float smn[NT / 32], smx[NT / 32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { smn[w] = mn; smx[w] = mx; }
    __syncthreads();
    if (w == 0) {
        mn = lane < NT / 32 ? smn[lane] : INFINITY;
        mx = lane < NT / 32 ? smx[lane] : -INFINITY;
        mn = warp_min(mn);
        mx = warp_max(mx);
        if (lane == 0) {
            atomicMin(enc + 2 * n, enc_f(mn));
            atomicMax(enc + 2 * n + 1, enc_f(mx));
        }
    }
}

__global__ void minmax_decode_kernel(const uint32_t* enc, float* mm, int N) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 * N) mm[i] = dec_f(enc[i]);
}

// out = (x - min) / (max - min), per sample (utils.py:48; IEEE sub and div, no epsilon)
__global__ void __launch_bounds__(NT) normalize_kernel(const float* __restrict__ x, const float* __restrict__ mm,
                                                       float* __restrict__ out, size_t V) {
    int n = blockIdx.y;
    float mn = mm[2 * n], r = __fsub_rn(mm[2 * n + 1], mn);
    const float* p = x + (size_t)n * V;
    float* o = out + (size_t)n * V;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < V; i += (size_t)gridDim.x * NT)
        o[i] = __fdiv_rn(__fsub_rn(p[i], mn), r);
}

// per-sample sums needed by the min-max-norm backward: sum g, sum g*n, #min ties, #max ties
__global__ void __launch_bounds__(NT) mmnorm_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ nrm,
                                                               const float* __restrict__ mm, const float* __restrict__ g,
                                                               size_t V, double* acc /* [N][4] */) {
    int n = blockIdx.y;
    float mn = mm[2 * n], mx = mm[2 * n + 1];
    size_t off = (size_t)n * V;
    double sg = 0, sgn = 0, cmin = 0, cmax = 0;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < V; i += (size_t)gridDim.x * NT) {
        float gi = g[off + i], xi = x[off + i];
        sg += gi;
        sgn += (double)gi * nrm[off + i];
        cmin += xi == mn;
        cmax += xi == mx;
    }
    __shared__ double sh[32];
    sg = block_sum_d(sg, sh);
    sgn = block_sum_d(sgn, sh);
    cmin = block_sum_d(cmin, sh);
    cmax = block_sum_d(cmax, sh);
    if (threadIdx.x == 0) {
        atomicAdd(acc + 4 * n + 0, sg);
        atomicAdd(acc + 4 * n + 1, sgn);
        atomicAdd(acc + 4 * n + 2, cmin);
        atomicAdd(acc + 4 * n + 3, cmax);
    }
}

// dx = g/r - [x==min] * sum(g*(1-n)) / (r*cnt_min) - [x==max] * sum(g*n) / (r*cnt_max)
__global__ void __launch_bounds__(NT) mmnorm_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ mm,
                                                              const float* __restrict__ g, const double* __restrict__ acc,
                                                              float* __restrict__ dx, size_t V, int accumulate) {
    int n = blockIdx.y;
    float mn = mm[2 * n], mx = mm[2 * n + 1];
    double r = (double)mx - (double)mn;
    double sg = acc[4 * n], sgn = acc[4 * n + 1], cmin = acc[4 * n + 2], cmax = acc[4 * n + 3];
    float inv_r = (float)(1.0 / r);
    float fix_min = (float)(-(sg - sgn) / (r * cmin));
    float fix_max = (float)(-sgn / (r * cmax));
    size_t off = (size_t)n * V;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < V; i += (size_t)gridDim.x * NT) {
        float xi = x[off + i];
        float d = g[off + i] * inv_r;
        if (xi == mn) d += fix_min;
        if (xi == mx) d += fix_max;
        dx[off + i] = accumulate ? dx[off + i] + d : d;
    }
}

// sum (a - b)^2  (b may be null -> constant target t)
__global__ void __launch_bounds__(NT) sqdiff_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, float t,
                                                        size_t n, double* acc) {
    double s = 0;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < n; i += (size_t)gridDim.x * NT) {
        float d = a[i] - (b ? b[i] : t);
        s += (double)d * d;
    }
    __shared__ double sh[32];
    s = block_sum_d(s, sh);
    if (threadIdx.x == 0) atomicAdd(acc, s);
}

// out (+)= c0 + c1*x1 + c2*x2 + c3*x3 (null inputs skipped)
__global__ void __launch_bounds__(NT) lincomb_kernel(float* __restrict__ out, size_t n, int accumulate, float c0,
                                                     const float* __restrict__ x1, float c1, const float* __restrict__ x2,
                                                     float c2, const float* __restrict__ x3, float c3) {
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < n; i += (size_t)gridDim.x * NT) {
        float v = c0;
        if (x1) v += c1 * x1[i];
        if (x2) v += c2 * x2[i];
        if (x3) v += c3 * x3[i];
        out[i] = accumulate ? out[i] + v : v;
    }
}

constexpr float BCE_EPS = 1e-7f;

// Keras binary_crossentropy(from_logits=False): clip p to [eps,1-eps]; -(y log(p+eps) + (1-y) log(1-p+eps))
__global__ void __launch_bounds__(NT) bce_sum_kernel(const float* __restrict__ y, const float* __restrict__ p, size_t n,
                                                     double* acc) {
    double s = 0;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < n; i += (size_t)gridDim.x * NT) {
        float yi = y[i];
        float pc = fminf(fmaxf(p[i], BCE_EPS), 1.0f - BCE_EPS);
        float l = yi * logf(pc + BCE_EPS) + (1.0f - yi) * logf(1.0f - pc + BCE_EPS);
        s -= (double)l;
    }
    __shared__ double sh[32];
    s = block_sum_d(s, sh);
    if (threadIdx.x == 0) atomicAdd(acc, s);
}

__global__ void __launch_bounds__(NT) bce_bwd_kernel(const float* __restrict__ y, const float* __restrict__ p, float coef,
                                                     float* __restrict__ g, size_t n, int accumulate) {
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < n; i += (size_t)gridDim.x * NT) {
        float yi = y[i], pi = p[i];
        float pc = fminf(fmaxf(pi, BCE_EPS), 1.0f - BCE_EPS);
        float d = 0.f;
        if (pi >= BCE_EPS && pi <= 1.0f - BCE_EPS) d = -coef * (yi / (pc + BCE_EPS) - (1.0f - yi) / (1.0f - pc + BCE_EPS));
        g[i] = accumulate ? g[i] + d : d;
    }
}

// acc[0..6] += { sum skp*yt, sum skp, sum skt*yp, sum skt, sum yt*yp, sum yt, sum yp }
__global__ void __launch_bounds__(NT) cldice_sums_kernel(const float* __restrict__ yt, const float* __restrict__ yp,
                                                         const float* __restrict__ skt, const float* __restrict__ skp,
                                                         size_t n, double* acc) {
    double s[7] = {0, 0, 0, 0, 0, 0, 0};
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < n; i += (size_t)gridDim.x * NT) {
        float a = yt[i], b = yp[i], c = skt[i], d = skp[i];
        s[0] += (double)d * a; s[1] += d; s[2] += (double)c * b; s[3] += c; s[4] += (double)a * b; s[5] += a; s[6] += b;
    }
    __shared__ double sh[32];
#pragma unroll
    for (int k = 0; k < 7; k++) {
        double v = block_sum_d(s[k], sh);
        if (threadIdx.x == 0) atomicAdd(acc + k, v);
    }
}

// ---------------------------------------------------------------- SSIM (3x3x3 Gaussian, zero 'SAME' padding)
constexpr int SX = 32, SY = 8, SZ = 8;
constexpr int SHX = SX + 2, SHY = SY + 2, SHZ = SZ + 2;

struct Taps {
    float g[3];
};

__device__ __forceinline__ void ssim_point(float mt, float mp, float ett, float epp, float etp, float& S, float& dmu,
                                           float& depp, float& detp) {
    const float c1 = 1e-4f, c2 = 9e-4f;  // (0.01*1)^2, (0.03*1)^2
    float stt = ett - mt * mt, spp = epp - mp * mp, stp = etp - mt * mp;
    float num1 = 2.f * mt * mp + c1, num2 = 2.f * stp + c2;
    float den1 = mt * mt + mp * mp + c1, den2 = stt + spp + c2;
    float inv = 1.f / (den1 * den2);
    S = num1 * num2 * inv;
    // partials w.r.t. mu_p (holding E[pp], E[tp] fixed), E[pp], E[tp]
    dmu = (2.f * mt * num2 - 2.f * mt * num1) * inv - S * (2.f * mp / den1 - 2.f * mp / den2);
    depp = -S / den2;
    detp = 2.f * num1 * inv;
}

// forward: acc += sum(1 - ssim); optionally writes the three partial-derivative maps for the backward
__global__ void __launch_bounds__(NT) ssim_fwd_kernel(const float* __restrict__ t, const float* __restrict__ p, int D, int H,
                                                      int W, Taps tp, double* acc, float* __restrict__ mA,
                                                      float* __restrict__ mB, float* __restrict__ mC, int tiles_x,
                                                      int tiles_y, int tiles_z) {
    __shared__ float st[SHZ * SHY * SHX], sp[SHZ * SHY * SHX];
    int b = blockIdx.x;
    int tx = b % tiles_x, ty = (b / tiles_x) % tiles_y, tz = (b / (tiles_x * tiles_y)) % tiles_z;
    int n = b / (tiles_x * tiles_y * tiles_z);
    size_t voff = (size_t)n * D * H * W;
    int z0 = tz * SZ, y0 = ty * SY, x0 = tx * SX;
    for (int i = threadIdx.x; i < SHZ * SHY * SHX; i += NT) {
        int lx = i % SHX, ly = (i / SHX) % SHY, lz = i / (SHX * SHY);
        int z = z0 + lz - 1, y = y0 + ly - 1, x = x0 + lx - 1;
        bool in = (unsigned)z < (unsigned)D && (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
        size_t g = voff + ((size_t)z * H + y) * W + x;
        st[i] = in ? __ldg(t + g) : 0.f;
        sp[i] = in ? __ldg(p + g) : 0.f;
    }
    __syncthreads();
    double local = 0;
    for (int i = threadIdx.x; i < SZ * SY * SX; i += NT) {
        int lx = i % SX, ly = (i / SX) % SY, lz = i / (SX * SY);
        int z = z0 + lz, y = y0 + ly, x = x0 + lx;
        if (z >= D || y >= H || x >= W) continue;
        float mt = 0, mp = 0, ett = 0, epp = 0, etp = 0;
#pragma unroll
        for (int dz = 0; dz < 3; dz++)
#pragma unroll
            for (int dy = 0; dy < 3; dy++)
#pragma unroll
                for (int dx = 0; dx < 3; dx++) {
                    float w = tp.g[dz] * tp.g[dy] * tp.g[dx];
                    int j = ((lz + dz) * SHY + (ly + dy)) * SHX + (lx + dx);
                    float a = st[j], c = sp[j];
                    mt += w * a; mp += w * c; ett += w * a * a; epp += w * c * c; etp += w * a * c;
                }
        float S, dmu, depp, detp;
        ssim_point(mt, mp, ett, epp, etp, S, dmu, depp, detp);
        local += 1.0 - (double)S;
        if (mA) {
            size_t g = voff + ((size_t)z * H + y) * W + x;
            mA[g] = dmu; mB[g] = depp; mC[g] = detp;
        }
    }
    __shared__ double sh[32];
    local = block_sum_d(local, sh);
    if (threadIdx.x == 0) atomicAdd(acc, local);
}

// backward: g_p[j] (+)= -coef * ( blur(A)[j] + 2 p[j] blur(B)[j] + t[j] blur(C)[j] )
__global__ void __launch_bounds__(NT) ssim_bwd_kernel(const float* __restrict__ t, const float* __restrict__ p,
                                                      const float* __restrict__ mA, const float* __restrict__ mB,
                                                      const float* __restrict__ mC, int D, int H, int W, Taps tp, float coef,
                                                      float* __restrict__ gp, int accumulate, int tiles_x, int tiles_y,
                                                      int tiles_z) {
    __shared__ float sa[SHZ * SHY * SHX], sb[SHZ * SHY * SHX], sc[SHZ * SHY * SHX];
    int b = blockIdx.x;
    int tx = b % tiles_x, ty = (b / tiles_x) % tiles_y, tz = (b / (tiles_x * tiles_y)) % tiles_z;
    int n = b / (tiles_x * tiles_y * tiles_z);
    size_t voff = (size_t)n * D * H * W;
    int z0 = tz * SZ, y0 = ty * SY, x0 = tx * SX;
    for (int i = threadIdx.x; i < SHZ * SHY * SHX; i += NT) {
        int lx = i % SHX, ly = (i / SHX) % SHY, lz = i / (SHX * SHY);
        int z = z0 + lz - 1, y = y0 + ly - 1, x = x0 + lx - 1;
        bool in = (unsigned)z < (unsigned)D && (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
        size_t g = voff + ((size_t)z * H + y) * W + x;
        sa[i] = in ? __ldg(mA + g) : 0.f;
        sb[i] = in ? __ldg(mB + g) : 0.f;
        sc[i] = in ? __ldg(mC + g) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SZ * SY * SX; i += NT) {
        int lx = i % SX, ly = (i / SX) % SY, lz = i / (SX * SY);
        int z = z0 + lz, y = y0 + ly, x = x0 + lx;
        if (z >= D || y >= H || x >= W) continue;
        float ba = 0, bb = 0, bc = 0;
#pragma unroll
        for (int dz = 0; dz < 3; dz++)
#pragma unroll
            for (int dy = 0; dy < 3; dy++)
#pragma unroll
                for (int dx = 0; dx < 3; dx++) {
                    float w = tp.g[dz] * tp.g[dy] * tp.g[dx];
                    int j = ((lz + dz) * SHY + (ly + dy)) * SHX + (lx + dx);
                    ba += w * sa[j]; bb += w * sb[j]; bc += w * sc[j];
                }
        size_t g = voff + ((size_t)z * H + y) * W + x;
        float d = -coef * (ba + 2.f * p[g] * bb + t[g] * bc);
        gp[g] = accumulate ? gp[g] + d : d;
    }
}

inline Taps make_taps() {
    // loss_functions.py:89-92: grid = [-1,0,1], sigma = 1.5, normalised
    Taps t;
    double g[3], s = 0;
    for (int i = 0; i < 3; i++) { double x = (i - 1) / 1.5; g[i] = exp(-0.5 * x * x) / (1.5 * sqrt(2.0 * M_PI)); s += g[i]; }
    for (int i = 0; i < 3; i++) t.g[i] = (float)(g[i] / s);
    return t;
}

inline dim3 grid2(size_t V, int N) { return dim3(vg_grid_for((long long)V / 4 + 1, NT, 4), N); }

}  // namespace

extern "C" {

// mm[N][2] = per-sample {min,max}; enc_ws: 2*N uint32 scratch
int vg_minmax(const float* x, int N, size_t V, float* mm, void* enc_ws, void* stream) {
    VG_REQUIRE(x && mm && enc_ws && N > 0 && V > 0);
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* enc = (uint32_t*)enc_ws;
    minmax_init_kernel<<<vg_cdiv(N, 128), 128, 0, st>>>(enc, N); VG_LAUNCHED(1);
    minmax_kernel<<<grid2(V, N), NT, 0, st>>>(x, V, enc); VG_LAUNCHED(1);
    minmax_decode_kernel<<<vg_cdiv(2 * N, 128), 128, 0, st>>>(enc, mm, N); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_minmax_normalize(const float* x, const float* mm, float* out, int N, size_t V, void* stream) {
    VG_REQUIRE(x && mm && out && N > 0);
    normalize_kernel<<<grid2(V, N), NT, 0, (cudaStream_t)stream>>>(x, mm, out, V); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

// dx (+)= d(normalised)/dx applied to g; acc_ws: N*4 doubles of scratch (zeroed here)
int vg_minmax_normalize_bwd(const float* x, const float* nrm, const float* mm, const float* g, float* dx, int N, size_t V,
                            void* acc_ws, int accumulate, void* stream) {
    VG_REQUIRE(x && nrm && mm && g && dx && acc_ws);
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(acc_ws, 0, (size_t)N * 4 * sizeof(double), st) != cudaSuccess) return VG_ERR_CUDA;
    mmnorm_bwd_reduce_kernel<<<grid2(V, N), NT, 0, st>>>(x, nrm, mm, g, V, (double*)acc_ws); VG_LAUNCHED(1);
    mmnorm_bwd_apply_kernel<<<grid2(V, N), NT, 0, st>>>(x, mm, g, (const double*)acc_ws, dx, V, accumulate); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_sqdiff_sum(const float* a, const float* b, float target, size_t n, double* acc, void* stream) {
    VG_REQUIRE(a && acc);
    sqdiff_sum_kernel<<<vg_grid_for(n, NT * 4, 4), NT, 0, (cudaStream_t)stream>>>(a, b, target, n, acc); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

// out = c[0] + c[1]*x1 + c[2]*x2 + c[3]*x3 with the four coefficients read from DEVICE memory (so a backward pass whose
// coefficients depend on reduced sums needs no host round trip)
__global__ void __launch_bounds__(NT) lincomb_dev_kernel(float* __restrict__ out, size_t n, int accumulate, const float* __restrict__ c,
                                                         const float* __restrict__ x1, const float* __restrict__ x2,
                                                         const float* __restrict__ x3) {
    const float c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3];
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < n; i += (size_t)gridDim.x * NT) {
        float v = c0;
        if (x1) v += c1 * x1[i];
        if (x2) v += c2 * x2[i];
        if (x3) v += c3 * x3[i];
        out[i] = accumulate ? out[i] + v : v;
    }
}

// coefficients of the clDice / Dice backward from the seven sums (clDice_func.py:83-119), k = upstream scale:
//   d skel_pred : coef[0] + coef[1]*y_true        d y_pred : coef[2] + coef[3]*y_true + coef[4]*skel_true + coef[5]*d0
__global__ void cldice_coeffs_kernel(const double* __restrict__ a, float alpha, float k, float* __restrict__ coef) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double smooth = 1.0;
    const double s0 = a[0], s1 = a[1], s2 = a[2], s3 = a[3], s4 = a[4], s5 = a[5], s6 = a[6];
    const double P = (s0 + smooth) / (s1 + smooth), R = (s2 + smooth) / (s3 + smooth);
    const double dcl_dP = -2.0 * R * R / ((P + R) * (P + R)), dcl_dR = -2.0 * P * P / ((P + R) * (P + R));
    const double den = s5 + s6 + smooth;
    const double kc = (double)k * alpha, kd = (double)k * (1.0 - alpha);
    coef[0] = (float)(-kc * dcl_dP * (s0 + smooth) / ((s1 + smooth) * (s1 + smooth)));
    coef[1] = (float)(kc * dcl_dP / (s1 + smooth));
    coef[2] = (float)(kd * (2.0 * s4 + smooth) / (den * den));
    coef[3] = (float)(-2.0 * kd / den);
    coef[4] = (float)(kc * dcl_dR / (s3 + smooth));
    coef[5] = 1.0f;
    coef[6] = 0.f;
    coef[7] = 0.f;
}

int vg_lincomb(float* out, size_t n, int accumulate, float c0, const float* x1, float c1, const float* x2, float c2,
               const float* x3, float c3, void* stream) {
    VG_REQUIRE(out);
    lincomb_kernel<<<vg_grid_for(n, NT * 4, 4), NT, 0, (cudaStream_t)stream>>>(out, n, accumulate, c0, x1, c1, x2, c2, x3, c3); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_bce_sum(const float* y_true, const float* y_pred, size_t n, double* acc, void* stream) {
    VG_REQUIRE(y_true && y_pred && acc);
    bce_sum_kernel<<<vg_grid_for(n, NT * 4, 4), NT, 0, (cudaStream_t)stream>>>(y_true, y_pred, n, acc); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_bce_bwd(const float* y_true, const float* y_pred, float coef, float* g, size_t n, int accumulate, void* stream) {
    VG_REQUIRE(y_true && y_pred && g);
    bce_bwd_kernel<<<vg_grid_for(n, NT * 4, 4), NT, 0, (cudaStream_t)stream>>>(y_true, y_pred, coef, g, n, accumulate); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_cldice_sums(const float* y_true, const float* y_pred, const float* skel_true, const float* skel_pred, size_t n,
                   double* acc7, void* stream) {
    VG_REQUIRE(y_true && y_pred && skel_true && skel_pred && acc7);
    cldice_sums_kernel<<<vg_grid_for(n, NT * 4, 4), NT, 0, (cudaStream_t)stream>>>(y_true, y_pred, skel_true, skel_pred, n, acc7); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

// acc += sum(1-ssim(t,p)); mA/mB/mC (optional, all or none) receive the backward partials
int vg_ssim_fwd(const float* t, const float* p, int N, int D, int H, int W, double* acc, float* mA, float* mB, float* mC,
                void* stream) {
    VG_REQUIRE(t && p && acc && N > 0);
    VG_REQUIRE((mA && mB && mC) || (!mA && !mB && !mC));
    int tx = vg_cdiv(W, SX), ty = vg_cdiv(H, SY), tz = vg_cdiv(D, SZ);
    ssim_fwd_kernel<<<tx * ty * tz * N, NT, 0, (cudaStream_t)stream>>>(t, p, D, H, W, make_taps(), acc, mA, mB, mC, tx, ty, tz); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

// g_p (+)= coef * d sum(1-ssim) / d p
int vg_ssim_bwd(const float* t, const float* p, const float* mA, const float* mB, const float* mC, int N, int D, int H, int W,
                float coef, float* gp, int accumulate, void* stream) {
    VG_REQUIRE(t && p && mA && mB && mC && gp);
    int tx = vg_cdiv(W, SX), ty = vg_cdiv(H, SY), tz = vg_cdiv(D, SZ);
    ssim_bwd_kernel<<<tx * ty * tz * N, NT, 0, (cudaStream_t)stream>>>(t, p, mA, mB, mC, D, H, W, make_taps(), coef, gp,
                                                                      accumulate, tx, ty, tz); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_lincomb_dev(float* out, size_t n, int accumulate, const float* coef4, const float* x1, const float* x2, const float* x3,
                   void* stream) {
    VG_REQUIRE(out && coef4);
    lincomb_dev_kernel<<<vg_grid_for(n, NT * 4, 4), NT, 0, (cudaStream_t)stream>>>(out, n, accumulate, coef4, x1, x2, x3); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_cldice_coeffs(const double* acc7, float alpha, float k, float* coef8, void* stream) {
    VG_REQUIRE(acc7 && coef8);
    cldice_coeffs_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(acc7, alpha, k, coef8); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

}  // extern "C"
