// 7x7x7 convolutions of the 'resnet' generator (generator.py:38 Conv3D(32, (7,7,7)) on the single-channel input, generator.py:67
// Conv3D(1, (7,7,7), padding='same') + tanh on the 32-channel decoder output): one side of these layers has ONE channel, so they are
// direct (CUDA-core) kernels like the other Cin == 1 / Cout == 1 shapes -- 343 taps of a single-channel tensor do not make a GEMM the
// tensor cores can be fed with.  Forward of the Cin == 1 layer is cin1_fwd_kernel<7, 1> (conv_mma.cu) and the input gradient of the
// Cout == 1 layer is cout1_dgrad_kernel (generic K); this file adds the two contractions that were missing for K = 7:
//
//   many_to_one   y1[p] = bias + sum_t sum_c M[p + sgn*t][c] * w[t][c]      forward of Cout == 1 (sgn = +1, M = padded x)
//                                                                           input gradient of Cin == 1 (sgn = -1, M = dy, zero outside)
//   wgrad_one     dw[t][c] += sum_m M[m][c] * S[m + sgn*t + off]            weight gradient of both: M = dy, S = x (Cin == 1: sgn +1)
//                                                                           or M = x, S = dy (Cout == 1: sgn -1, zero outside)
// M: bf16 NDHWC with C % 8 == 0; S / y1: fp32 single channel.  Stride 1 only.
#include "common.cuh"

namespace {

constexpr int NT = 256;

// one thread per output voxel; the weights of all taps sit in shared memory as fp32 [T][C] and are read as broadcasts
template <int C>
__global__ void __launch_bounds__(NT) many_to_one_kernel(const bf16* __restrict__ M, const bf16* __restrict__ w, int w_tstride,
                                                         const float* __restrict__ bias, float* __restrict__ y, int N, int MD, int MH, int MW,
                                                         int YD, int YH, int YW, int K, int sgn, int act) {
    extern __shared__ float sw[];   // [K^3][C]
    const int T = K * K * K;
    for (int i = threadIdx.x; i < T * C; i += NT) sw[i] = __bfloat162float(w[(size_t)(i / C) * w_tstride + (i % C)]);
    __syncthreads();
    const size_t total = (size_t)N * YD * YH * YW;
    const float b0 = bias ? bias[0] : 0.f;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        const int pw = (int)(i % YW);
        size_t r = i / YW;
        const int ph = (int)(r % YH); r /= YH;
        const int pd = (int)(r % YD);
        const int n = (int)(r / YD);
        const bf16* Mn = M + (size_t)n * MD * MH * MW * C;
        float acc = b0;
        for (int kd = 0; kd < K; kd++) {
            const int d = pd + sgn * kd;
            if ((unsigned)d >= (unsigned)MD) continue;
            for (int kh = 0; kh < K; kh++) {
                const int h = ph + sgn * kh;
                if ((unsigned)h >= (unsigned)MH) continue;
                const bf16* row = Mn + ((size_t)d * MH + h) * MW * C;
                const float* wr = sw + ((kd * K + kh) * K) * C;
                for (int kw = 0; kw < K; kw++) {
                    const int x = pw + sgn * kw;
                    if ((unsigned)x >= (unsigned)MW) continue;
#pragma unroll
                    for (int c8 = 0; c8 < C / 8; c8++) {
                        float f[8];
                        load8<bf16>(row + (size_t)x * C + c8 * 8, f);
                        const float4 w0 = *reinterpret_cast<const float4*>(wr + kw * C + c8 * 8);
                        const float4 w1 = *reinterpret_cast<const float4*>(wr + kw * C + c8 * 8 + 4);
                        acc = fmaf(f[0], w0.x, acc); acc = fmaf(f[1], w0.y, acc); acc = fmaf(f[2], w0.z, acc); acc = fmaf(f[3], w0.w, acc);
                        acc = fmaf(f[4], w1.x, acc); acc = fmaf(f[5], w1.y, acc); acc = fmaf(f[6], w1.z, acc); acc = fmaf(f[7], w1.w, acc);
                    }
                }
            }
        }
        y[i] = act == VG_ACT_TANH ? tanhf(acc) : acc;
    }
}

// blockIdx.y = (kd, kh); a thread owns an 8-channel group and the K taps along w: K x 8 accumulators.  The block walks a chunk of M's
// voxels, every 16-byte cell of M is read once per (kd, kh) pair and meets the K single-channel values S[m_w + sgn*kw + off].
template <int K>
__global__ void __launch_bounds__(NT) wgrad_one_kernel(const bf16* __restrict__ M, const float* __restrict__ S, float* __restrict__ dw, int N,
                                                       int MD, int MH, int MW, int C, int SD, int SH, int SW, int sgn, int off,
                                                       size_t per_block) {
    extern __shared__ float sred[];   // [nvl][C] per kw pass
    const int kd = blockIdx.y / K, kh = blockIdx.y % K;
    const int cg = C / 8, nvl = NT / cg;
    const int c8 = threadIdx.x % cg, vl = threadIdx.x / cg;
    const size_t V = (size_t)N * MD * MH * MW;
    const size_t v0 = (size_t)blockIdx.x * per_block, v1 = v0 + per_block < V ? v0 + per_block : V;
    float acc[K][8];
#pragma unroll
    for (int q = 0; q < K; q++)
#pragma unroll
        for (int k = 0; k < 8; k++) acc[q][k] = 0.f;
    if (vl < nvl)
        for (size_t v = v0 + vl; v < v1; v += nvl) {
            const int mw = (int)(v % MW);
            size_t r = v / MW;
            const int mh = (int)(r % MH); r /= MH;
            const int md = (int)(r % MD);
            const int n = (int)(r / MD);
            const int sd = md + sgn * kd + off, sh = mh + sgn * kh + off;
            if ((unsigned)sd >= (unsigned)SD || (unsigned)sh >= (unsigned)SH) continue;
            const float* srow = S + (((size_t)n * SD + sd) * SH + sh) * SW;
            float f[8];
            load8<bf16>(M + v * C + c8 * 8, f);
#pragma unroll
            for (int q = 0; q < K; q++) {
                const int sx = mw + sgn * q + off;
                const float s = (unsigned)sx < (unsigned)SW ? __ldg(srow + sx) : 0.f;
#pragma unroll
                for (int k = 0; k < 8; k++) acc[q][k] = fmaf(s, f[k], acc[q][k]);
            }
        }
    for (int q = 0; q < K; q++) {
        __syncthreads();
        if (vl < nvl) {
#pragma unroll
            for (int k = 0; k < 8; k++) sred[(size_t)vl * C + c8 * 8 + k] = acc[q][k];
        }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += NT) {
            float t = 0.f;
            for (int l = 0; l < nvl; l++) t += sred[(size_t)l * C + c];
            atomicAdd(dw + (size_t)((kd * K + kh) * K + q) * C + c, t);
        }
    }
}

}  // namespace

// y1 = conv(M, w): forward of a Cout == 1 layer over an explicitly padded input (sgn = +1), or the input gradient of a Cin == 1 layer
// (sgn = -1: M = dy, positions outside dy contribute zero).  w: bf16, element (t, c) at w[t * w_tstride + c].
int vg_k7_many_to_one(const bf16* M, const bf16* w, int w_tstride, const float* bias, float* y, int N, int MD, int MH, int MW, int C, int YD,
                      int YH, int YW, int K, int sgn, int act, cudaStream_t st) {
    if (K != 7 || (C != 16 && C != 32 && C != 64)) return VG_ERR_UNSUPPORTED;
    const size_t smem = (size_t)K * K * K * C * sizeof(float);
    const size_t total = (size_t)N * YD * YH * YW;
    const int grid = vg_grid_for(total, NT, 8);
#define VG_M21(CC)                                                                                                                    \
    do {                                                                                                                              \
        static VgPerDevice attr;                                                                                                     \
        if (!attr.done()) {                                                                                                                  \
            if (cudaFuncSetAttribute(many_to_one_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess) \
                return VG_ERR_CUDA;                                                                                                   \
            attr.mark();                                                                                                              \
        }                                                                                                                             \
        many_to_one_kernel<CC><<<grid, NT, smem, st>>>(M, w, w_tstride, bias, y, N, MD, MH, MW, YD, YH, YW, K, sgn, act);            \
    } while (0)
    if (C == 16) VG_M21(16);
    else if (C == 32) VG_M21(32);
    else VG_M21(64);
#undef VG_M21
    VG_LAUNCHED(1);
    return VG_OK;
}

// dw[t][c] += sum over M's voxels m of M[m][c] * S[m + sgn*t + off] (zero outside S).  Cin == 1 layer: M = dy, S = x, sgn = +1, off = 0;
// Cout == 1 layer: M = x (padded input), S = dy, sgn = -1, off = 0.
int vg_k7_wgrad_one(const bf16* M, const float* S, float* dw, int N, int MD, int MH, int MW, int C, int SD, int SH, int SW, int K, int sgn,
                    int off, cudaStream_t st) {
    if (K != 7 || C % 8 || C > 8 * NT) return VG_ERR_UNSUPPORTED;
    const size_t V = (size_t)N * MD * MH * MW;
    int nbx = (148 * 8 + K * K - 1) / (K * K);
    size_t per_block = (V + nbx - 1) / nbx;
    if (per_block < 256) per_block = 256;
    const size_t smem = (size_t)(NT / (C / 8)) * C * sizeof(float);
    wgrad_one_kernel<7><<<dim3(vg_cdiv((long long)V, (long long)per_block), K * K), NT, smem, st>>>(M, S, dw, N, MD, MH, MW, C, SD, SH, SW, sgn,
                                                                                               off, per_block); VG_LAUNCHED(1);
    return VG_OK;
}
