// HBM-bound convolution shapes: 1x1x1 kernels, single-channel outputs (generator head, PatchGAN logits)
// and the single-channel-input dgrad of the PatchGAN's first layer.
//
// These replace cuDNN behind the Keras layers whose arithmetic intensity is far below the tensor-core ridge
// ("MEMORY-bound: all 1x1 shortcuts, head"): the decoder shortcuts `Conv3D(f,(1,1,1))` of
// resunet_model.py:127, the head `Conv3D(1,(1,1,1),activation='tanh')` of resunet_model.py:245, the logits
// conv of discriminator.py:107-114 and the input gradient of discriminator.py:63-69 (k4 s2, 1 -> 64).
// Their roofline is one pass over the wide tensor, so they are written as streaming kernels:
//   k1_fwd_kernel    y[v][co] = x[v][:] W          mma.sync fragments loaded straight from global (no smem, no barriers)
//   k1_wgrad_kernel  dW = X^T dY                   warp-private cp.async rings + ldmatrix.trans, fp32 atomics at the end
//   cout1_*          dot products / outer products over one 8-channel packet per thread
//   cin1_dgrad_s2    all 8 stride-parity classes in ONE mma: M = 16 voxels, K = taps x Cout, N = 8 classes
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cpa16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm4(uint32_t a, uint32_t* r) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm4_t(uint32_t a, uint32_t* r) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t ldg32(const bf16* p) { return __ldg(reinterpret_cast<const unsigned int*>(p)); }

// ------------------------------------------------------------------------------------------ 1x1x1 forward
// y[v][co] = sum_ci x[v][ci] * w[ci][co] + b[co]; wp = bf16 [Np][Cin] (row = co, ci contiguous: the mma.sync operand pack).
// A warp owns 16-voxel tiles; the A fragments are 4-byte loads of full 32-byte sectors, two tiles in flight per warp.
template <int KS, int NT>
__global__ void __launch_bounds__(256) k1_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ wp, const float* __restrict__ bias,
                                                     bf16* __restrict__ y, long long nvox) {
    constexpr int CIN = KS * 16, COUT = NT * 8, U = 2;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    uint32_t b[KS][NT][2];
#pragma unroll
    for (int ks = 0; ks < KS; ks++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            const bf16* wr = wp + (size_t)(nt * 8 + g) * CIN + ks * 16 + 2 * t;
            b[ks][nt][0] = ldg32(wr);
            b[ks][nt][1] = ldg32(wr + 8);
        }
    float bv[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
        bv[nt][0] = bias ? bias[nt * 8 + 2 * t] : 0.f;
        bv[nt][1] = bias ? bias[nt * 8 + 2 * t + 1] : 0.f;
    }
    const long long ntiles = (nvox + 15) >> 4;
    const long long nw = (long long)gridDim.x * 8;
    for (long long tile = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); tile < ntiles; tile += U * nw) {
        uint32_t a[U][KS][4];
        long long v0[U];
        bool ok0[U], ok1[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long long tl = tile + u * nw;
            v0[u] = tl * 16 + g;
            ok0[u] = tl < ntiles && v0[u] < nvox;
            ok1[u] = tl < ntiles && v0[u] + 8 < nvox;
            const bf16* r0 = x + (size_t)(ok0[u] ? v0[u] : 0) * CIN + 2 * t;
            const bf16* r1 = x + (size_t)(ok1[u] ? v0[u] + 8 : 0) * CIN + 2 * t;
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
                a[u][ks][0] = ok0[u] ? ldg32(r0 + ks * 16) : 0u;
                a[u][ks][1] = ok1[u] ? ldg32(r1 + ks * 16) : 0u;
                a[u][ks][2] = ok0[u] ? ldg32(r0 + ks * 16 + 8) : 0u;
                a[u][ks][3] = ok1[u] ? ldg32(r1 + ks * 16 + 8) : 0u;
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            float c[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                c[nt][0] = bv[nt][0]; c[nt][1] = bv[nt][1]; c[nt][2] = bv[nt][0]; c[nt][3] = bv[nt][1];
            }
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
#pragma unroll
                for (int nt = 0; nt < NT; nt++) mma16816(c[nt], a[u][ks], b[ks][nt][0], b[ks][nt][1]);
            bf16* y0 = y + (size_t)v0[u] * COUT + 2 * t;
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                if (ok0[u]) *reinterpret_cast<uint32_t*>(y0 + nt * 8) = pack2_bf16(c[nt][0], c[nt][1]);
                if (ok1[u]) *reinterpret_cast<uint32_t*>(y0 + 8 * COUT + nt * 8) = pack2_bf16(c[nt][2], c[nt][3]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ 1x1x1 weight gradient
// dw[ci][co] += sum_v x[src(v)][ci] * dy[v][co]   (src(v) = v for stride 1, the even-index voxel for stride 2).
// The reduction index (voxel) is the slow index of both operands, so both go through ldmatrix.trans.  Every WARP runs its own
// cp.async ring over 16-voxel slabs (no block barriers in the main loop); row pitches are odd multiples of 16 bytes so the
// eight row addresses of one ldmatrix phase fall into eight distinct bank groups.
template <int KS, int NT>
__global__ void __launch_bounds__(256) k1_wgrad_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dw,
                                                       int nvox, int stride, int OD, int OH, int OW, int ID, int IH, int IW) {
    constexpr int CIN = KS * 16, COUT = NT * 8;
    constexpr int CHX = CIN / 8, CHY = COUT / 8, CH = CHX + CHY;
    constexpr int PX = CHX | 1, PY = CHY | 1;
    constexpr int SLAB = 16 * (PX + PY) * 16;   // bytes per stage per warp
    constexpr int STAGES = 3;
    extern __shared__ __align__(16) uint8_t k1w_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t wbase = sm_u32(k1w_smem) + (uint32_t)warp * STAGES * SLAB;
    float acc[KS][NT][4];
#pragma unroll
    for (int ks = 0; ks < KS; ks++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[ks][nt][e] = 0.f;
    const int nslabs = (nvox + 15) >> 4;
    const int nw = gridDim.x * 8;
    const int first = blockIdx.x * 8 + warp;

    auto issue = [&](int slab, int stage) {
        if (slab < nslabs) {
            const uint32_t sb = wbase + (uint32_t)stage * SLAB;
#pragma unroll
            for (int i0 = 0; i0 < 16 * CH; i0 += 32) {
                const int i = i0 + lane;
                if ((16 * CH) % 32 == 0 || i < 16 * CH) {
                    const int vl = i / CH, c = i - vl * CH;
                    const int v = slab * 16 + vl;
                    const bool ok = v < nvox;
                    if (c < CHX) {
                        long long xv = v;
                        if (stride == 2) {
                            const int ow = v % OW, r = v / OW;
                            const int oh = r % OH, r2 = r / OH;
                            const int od = r2 % OD, n = r2 / OD;
                            xv = (((long long)n * ID + 2 * od) * IH + 2 * oh) * IW + 2 * ow;
                        }
                        cpa16(sb + (uint32_t)(vl * PX + c) * 16, ok ? x + xv * CIN + c * 8 : x, ok ? 16 : 0);
                    } else {
                        cpa16(sb + (uint32_t)(16 * PX + vl * PY + (c - CHX)) * 16, ok ? dy + (long long)v * COUT + (c - CHX) * 8 : dy, ok ? 16 : 0);
                    }
                }
            }
        }
        cpa_commit();
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) issue(first + s * nw, s);
    int stage = 0;
    const int j = lane >> 3, r = lane & 7;
    for (int slab = first; slab < nslabs; slab += nw) {
        int pst = stage + STAGES - 1;
        if (pst >= STAGES) pst -= STAGES;
        issue(slab + (STAGES - 1) * nw, pst);
        cpa_wait<STAGES - 1>();
        __syncwarp();
        const uint32_t sb = wbase + (uint32_t)stage * SLAB;
        uint32_t bf[NT][2];
#pragma unroll
        for (int np = 0; np < NT / 2; np++) {
            uint32_t q[4];
            ldsm4_t(sb + (uint32_t)(16 * PX + ((j & 1) * 8 + r) * PY + 2 * np + (j >> 1)) * 16, q);
            bf[2 * np][0] = q[0]; bf[2 * np][1] = q[1]; bf[2 * np + 1][0] = q[2]; bf[2 * np + 1][1] = q[3];
        }
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
            uint32_t a[4];
            ldsm4_t(sb + (uint32_t)(((j >> 1) * 8 + r) * PX + 2 * ks + (j & 1)) * 16, a);
#pragma unroll
            for (int nt = 0; nt < NT; nt++) mma16816(acc[ks][nt], a, bf[nt][0], bf[nt][1]);
        }
        __syncwarp();
        if (++stage == STAGES) stage = 0;
    }
    cpa_wait<0>();
    __syncthreads();
    // block reduction through shared memory (the rings are dead now), then one atomic per output per block
    float* red = reinterpret_cast<float*>(k1w_smem);
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int ks = 0; ks < KS; ks++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            float* o = red + (size_t)warp * CIN * COUT + (ks * 16 + g) * COUT + nt * 8 + 2 * t;
            o[0] = acc[ks][nt][0]; o[1] = acc[ks][nt][1];
            o[8 * COUT] = acc[ks][nt][2]; o[8 * COUT + 1] = acc[ks][nt][3];
        }
    __syncthreads();
    for (int i = threadIdx.x; i < CIN * COUT; i += 256) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) s += red[(size_t)w * CIN * COUT + i];
        atomicAdd(dw + i, s);
    }
}

template <int KS, int NT>
constexpr size_t k1w_smem_bytes() {
    constexpr size_t ring = (size_t)8 * 3 * 16 * (((KS * 2) | 1) + (NT | 1)) * 16;
    constexpr size_t red = (size_t)8 * KS * 16 * NT * 8 * 4;
    return ring > red ? ring : red;
}

template <int KS, int NT>
int launch_k1_wgrad(const bf16* x, const bf16* dy, float* dw, int nvox, int stride, int OD, int OH, int OW, int ID, int IH, int IW,
                    cudaStream_t st) {
    constexpr size_t smem = k1w_smem_bytes<KS, NT>();
    static VgPerDevice attr;
    if (!attr.done()) {
        if (cudaFuncSetAttribute(k1_wgrad_kernel<KS, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return VG_ERR_CUDA;
        attr.mark();
    }
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm > 3) per_sm = 3;
    if (per_sm < 1) per_sm = 1;
    int grid = 148 * per_sm;
    const int nslabs = (nvox + 15) / 16;
    if (grid * 8 > nslabs) grid = (nslabs + 7) / 8;
    k1_wgrad_kernel<KS, NT><<<grid, 256, smem, st>>>(x, dy, dw, nvox, stride, OD, OH, OW, ID, IH, IW);
    VG_LAUNCHED(1);
    return VG_OK;
}

// ------------------------------------------------------------------------------------------ Cout == 1, K == 1 (generator head)
// y[v] = act(sum_c x[v][c] * w[c] + b): CIN/8 lanes per voxel, one 128-bit load each (fully coalesced), shuffle reduction.
template <int CIN>
__global__ void __launch_bounds__(256) cout1_k1_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ wp, const float* __restrict__ bias,
                                                           float* __restrict__ y, size_t nvox, int act) {
    constexpr int CG = CIN / 8;
    const int c8 = threadIdx.x % CG;
    float w[8];
    load8<bf16>(wp + c8 * 8, w);
    const float b0 = bias ? bias[0] : 0.f;
    const size_t total = nvox * CG, step = (size_t)gridDim.x * 256;
    const size_t iters = (total + step - 1) / step;
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    for (size_t it = 0; it < iters; it++, i += step) {   // uniform trip count: the shuffles below need full warps
        float s = 0.f;
        if (i < total) {
            float f[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), f);
#pragma unroll
            for (int k = 0; k < 8; k++) s = fmaf(f[k], w[k], s);
        }
#pragma unroll
        for (int o = 1; o < CG; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (c8 == 0 && i < total) {
            s += b0;
            y[i / CG] = act == VG_ACT_TANH ? tanhf(s) : s;
        }
    }
}

// dx[v][c] = dy[v] * w[c]   (w fp32 [Cin])
__global__ void __launch_bounds__(256) cout1_k1_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, bf16* __restrict__ dx,
                                                             size_t nvox, int Cin) {
    const int cg = Cin / 8;   // host guarantees 256 % cg == 0, so the channel group of a thread is loop-invariant
    const int c8 = threadIdx.x % cg;
    float wv[8];
    load8<float>(w + c8 * 8, wv);
    const size_t total = nvox * cg;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const float g = __ldg(dy + i / cg);
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; k++) o[k] = g * wv[k];
        reinterpret_cast<uint4*>(dx)[i] = pack8(o);
    }
}

// dw[c] += sum_v x[v][c] * dy[v]
__global__ void __launch_bounds__(256) cout1_k1_wgrad_kernel(const bf16* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                                                             size_t nvox, int Cin) {
    __shared__ float sred[256 * 8];
    const int cg = Cin / 8;   // host guarantees 256 % cg == 0: a thread's channel group (threadIdx.x % cg) is loop-invariant
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = 0.f;
    const size_t total = nvox * cg, step = (size_t)gridDim.x * 256;
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    for (; i + 3 * step < total; i += 4 * step) {
        uint4 p[4];
        float g[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            p[u] = __ldg(reinterpret_cast<const uint4*>(x) + i + u * step);
            g[u] = __ldg(dy + (i + u * step) / cg);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            float f[8];
            unpack8(p[u], f);
#pragma unroll
            for (int k = 0; k < 8; k++) a[k] = fmaf(g[u], f[k], a[k]);
        }
    }
    for (; i < total; i += step) {
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), f);
        const float g = __ldg(dy + i / cg);
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = fmaf(g, f[k], a[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) sred[threadIdx.x * 8 + k] = a[k];
    __syncthreads();
    for (int c = threadIdx.x; c < Cin; c += 256) {
        float s = 0.f;
        for (int l = c / 8; l < 256; l += cg) s += sred[l * 8 + (c & 7)];
        atomicAdd(dw + c, s);
    }
}

// ------------------------------------------------------------------------------------------ Cout == 1, K > 1 weight gradient (PatchGAN logits conv)
// dw[t][ci] += sum_o x[o+t][ci] * dy[o].  A block owns a run of x rows (n, pd, ph) and ONE kd (blockIdx.y); a thread owns an
// 8-channel packet and a w-lane and keeps the K*K (kh, kw) partial sums of its packet in registers, so x is read K times in
// total (once per kd) instead of K^3 times.
template <int K>
__global__ void __launch_bounds__(256) cout1_wgrad_rows_kernel(const bf16* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                                                               int N, int ID, int IH, int IW, int OD, int OH, int OW, int Cin, int rows_per_block) {
    const int cg = Cin / 8, nwl = 256 / cg;
    const int c8 = threadIdx.x % cg, wl = threadIdx.x / cg;
    const int kd = blockIdx.y;
    float acc[K * K][8];
#pragma unroll
    for (int q = 0; q < K * K; q++)
#pragma unroll
        for (int k = 0; k < 8; k++) acc[q][k] = 0.f;
    const int rows = N * ID * IH;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    if (wl < nwl)
        for (int row = r0; row < r1; row++) {
            const int ph = row % IH, r = row / IH;
            const int pd = r % ID, n = r / ID;
            const int od = pd - kd;
            if ((unsigned)od >= (unsigned)OD) continue;
            const bf16* xr = x + (size_t)row * IW * Cin + c8 * 8;
            const float* dyn = dy + ((size_t)n * OD + od) * OH * OW;
            for (int pw = wl; pw < IW; pw += nwl) {
                float f[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(xr + (size_t)pw * Cin)), f);
#pragma unroll
                for (int kh = 0; kh < K; kh++) {
                    const int oh = ph - kh;
                    if ((unsigned)oh >= (unsigned)OH) continue;
#pragma unroll
                    for (int kw = 0; kw < K; kw++) {
                        const int ow = pw - kw;
                        if ((unsigned)ow >= (unsigned)OW) continue;
                        const float g = __ldg(dyn + (size_t)oh * OW + ow);
#pragma unroll
                        for (int k = 0; k < 8; k++) acc[kh * K + kw][k] = fmaf(g, f[k], acc[kh * K + kw][k]);
                    }
                }
            }
        }
    // the w-lanes of a block are summed in shared memory first: one atomic per (block, tap, channel) instead of one per thread
    // (1 x 18^3 x 512: 9 M atomics on 13.8 k addresses were the whole run time of the launch)
    __shared__ float sred[256 * 8];
#pragma unroll
    for (int q = 0; q < K * K; q++) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; k++) sred[(wl * cg + c8) * 8 + k] = wl < nwl ? acc[q][k] : 0.f;
        __syncthreads();
        for (int i = threadIdx.x; i < Cin; i += 256) {
            float t = 0.f;
            for (int l = 0; l < nwl; l++) t += sred[l * Cin + i];
            atomicAdd(dw + (size_t)(kd * K * K + q) * Cin + i, t);
        }
    }
}

// ------------------------------------------------------------------------------------------ Cin == 1, K == 4, stride 2 dgrad (PatchGAN d0)
// dx[2p'+a] = sum_{t' in {0,1}^3} sum_co dy[p' - t'][co] * w[a + 2t'][co]   for the 8 parity classes a.
// The dy rows a voxel p' needs do not depend on the class, so the classes are the N dimension of one mma:
//   M = 16 consecutive p'w, K = 8 taps x Cout, N = 8 classes  ->  D[p'w][class], and classes (2t, 2t+1) of a thread are the
// two w-parities, i.e. two ADJACENT output voxels (one 8-byte store).  A block stages the (4+1)x(4+1)x(16+1) dy brick once.
template <int NC>   // NC = Cout / 16
__global__ void __launch_bounds__(256) cin1_dgrad_s2_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ wd, float* __restrict__ dx,
                                                            int N, int ID, int IH, int IW, int OD, int OH, int OW, int nbd, int nbh, int nbw) {
    constexpr int COUT = NC * 16, BD = 4, BH = 4, BW = 16;
    constexpr int ED = BD + 1, EH = BH + 1, EW = BW + 1;
    constexpr int PITCH = COUT * 2 + 16;
    constexpr int NCH = COUT / 8;
    extern __shared__ __align__(16) uint8_t c1d_smem[];
    const uint32_t sbase = sm_u32(c1d_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    // B fragments: column n = class g; k = (tap t', channel).  Pack layout (pack_dgrad_kernel): [class][t'][NpI = 64][Cout], row ci = 0.
    uint32_t b[8][NC][2];
#pragma unroll
    for (int tp = 0; tp < 8; tp++)
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const bf16* wr = wd + ((size_t)g * 8 + tp) * 64 * COUT + c * 16 + 2 * t;
            b[tp][c][0] = ldg32(wr);
            b[tp][c][1] = ldg32(wr + 8);
        }
    const int nbricks = N * nbd * nbh * nbw;
    const int j = lane >> 3, r = lane & 7;
    for (int brick = blockIdx.x; brick < nbricks; brick += gridDim.x) {
        int q = brick;
        const int bw = q % nbw; q /= nbw;
        const int bh = q % nbh; q /= nbh;
        const int bd = q % nbd;
        const int n = q / nbd;
        const int pd0 = bd * BD, ph0 = bh * BH, pw0 = bw * BW;
        __syncthreads();   // the previous brick has been consumed
        const bf16* dyn = dy + (size_t)n * OD * OH * OW * COUT;
        for (int i = threadIdx.x; i < ED * EH * EW * NCH; i += 256) {
            const int row = i / NCH, ch = i - row * NCH;
            const int lw = row % EW, r2 = row / EW;
            const int lh = r2 % EH, ld = r2 / EH;
            const int od = pd0 + ld - 1, oh = ph0 + lh - 1, ow = pw0 + lw - 1;
            const bool ok = (unsigned)od < (unsigned)OD && (unsigned)oh < (unsigned)OH && (unsigned)ow < (unsigned)OW;
            cpa16(sbase + (uint32_t)row * PITCH + ch * 16, ok ? dyn + (((size_t)od * OH + oh) * OW + ow) * COUT + ch * 8 : dy, ok ? 16 : 0);
        }
        cpa_commit();
        cpa_wait<0>();
        __syncthreads();
#pragma unroll 1
        for (int mt = warp; mt < BD * BH; mt += 8) {
            const int ld = mt >> 2, lh = mt & 3;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int tp = 0; tp < 8; tp++) {
                const int td = tp >> 2, th = (tp >> 1) & 1, tw = tp & 1;
                const int rowi = ((ld - td + 1) * EH + (lh - th + 1)) * EW + ((j & 1) * 8 + r - tw + 1);
                const uint32_t abase = sbase + (uint32_t)rowi * PITCH + (j >> 1) * 16;
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    uint32_t a[4];
                    ldsm4(abase + c * 32, a);
                    mma16816(acc, a, b[tp][c][0], b[tp][c][1]);
                }
            }
            // thread (g, t): rows p'w = g, g+8; columns = classes 2t, 2t+1 = (ad, ah) = (t>>1, t&1), aw = 0, 1
            const int pd = 2 * (pd0 + ld) + (t >> 1), ph = 2 * (ph0 + lh) + (t & 1);
            if (pd < ID && ph < IH) {
                float* orow = dx + (((size_t)n * ID + pd) * IH + ph) * IW;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int pw = 2 * (pw0 + g + 8 * h);
                    if (pw < IW) orow[pw] = acc[2 * h];
                    if (pw + 1 < IW) orow[pw + 1] = acc[2 * h + 1];
                }
            }
        }
    }
}


// ------------------------------------------------------------------------------------------ Cin == 1 weight gradient on tensor cores
// dw[tap][co] += sum_o x[S*o + tap] * dy[o][co]   (stem 1->16 k3 s1, PatchGAN d0 1->64 k4 s2; resunet_model.py:89, discriminator.py:63).
// GEMM view: M = taps (27 -> 32, 64), N = Cout, K = 16 consecutive output voxels of one row.  x is fp32 and dw must be fp32-exact
// (the reference differentiates fp32 inputs), so x is split on the fly into THREE bf16 planes h + m + l = x (exact to 2^-25) and
// every k-step issues three MMAs per tile; dy is already bf16.  The A fragment of tap (kd,kh,kw) is a pair of adjacent elements
// of the x row (kd, S*oh+kh), de-interleaved by w-parity for stride 2: an aligned 32-bit shared load, or two loads and a funnel
// shift for odd offsets.  The B fragment is dy^T through ldmatrix.trans (XOR-swizzled 16-byte chunks).  A spare M row of ones
// yields the bias gradient for free when taps < 16*MT.
__device__ __forceinline__ void split3(float v, uint32_t& h, uint32_t& m, uint32_t& l) {
    const __nv_bfloat16 hb = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(hb);
    const __nv_bfloat16 mb = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(mb);
    const __nv_bfloat16 lb = __float2bfloat16_rn(r2);
    h = __bfloat16_as_ushort(hb); m = __bfloat16_as_ushort(mb); l = __bfloat16_as_ushort(lb);
}

template <int K, int S, int COUT>
__global__ void __launch_bounds__(256) cin1_wgrad_tc_kernel(const float* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dw,
                                                            float* __restrict__ dbias, int N, int ID, int IH, int IW, int OD, int OH, int OW,
                                                            int WP, int OWp) {
    constexpr int T = K * K * K, MT = (T + 15) / 16, NG = COUT / 16, RW = 8 / NG, NCHK = COUT / 8;
    constexpr int BH = S == 1 ? 8 : 4, HB = (BH - 1) * S + K;
    constexpr bool ONES = T < MT * 16;
    constexpr int SWZ_SH = NCHK == 8 ? 0 : (NCHK == 4 ? 1 : 2);
    extern __shared__ __align__(16) uint8_t c1w_smem[];
    const int plane_elems = S * K * HB * WP;                // one split of one copy
    // [2 copies][3 splits][S][K][HB][WP]: copy 0 holds P[i], copy 1 holds P[i+1], so that the element pair of ANY tap offset
    // starts on an even index = one aligned 32-bit shared load (no funnel shift, no select)
    uint16_t* xs = reinterpret_cast<uint16_t*>(c1w_smem);
    const uint32_t xs_u32 = sm_u32(c1w_smem);
    const uint32_t dy_u32 = xs_u32 + (uint32_t)(6 * plane_elems * 2 + 15) / 16 * 16;   // [BH][OWp][NCHK] 16-byte chunks, swizzled
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int ng = warp % NG, rw = warp / NG;
    // per-lane tap constants: element offset of the tap inside a split plane (row oh_l = 0, voxel 0), -1 = zero row, -2 = ones row
    int toff[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
            const int tap = mt * 16 + g + 8 * hf;
            if (tap < T) {
                const int kw = tap % K, kh = (tap / K) % K, kd = tap / (K * K);
                const int shf = kw / S;
                toff[mt][hf] = (shf & 1) * 3 * plane_elems + (((kw % S) * K + kd) * HB + kh) * WP + (shf & ~1);
            } else toff[mt][hf] = (ONES && tap == T) ? -2 : -1;
        }
    float acc[MT][2][4];
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[mt][nt][e] = 0.f;
    const int nhb = (OH + BH - 1) / BH;
    const int nbricks = N * OD * nhb;
    const int ksteps = BH * (OWp / 16);
    const int j = lane >> 3, r = lane & 7;
    for (int brick = blockIdx.x; brick < nbricks; brick += gridDim.x) {
        const int hb = brick % nhb, q = brick / nhb;
        const int od = q % OD, n = q / OD;
        const int oh0 = hb * BH;
        __syncthreads();   // previous brick consumed
        // dy tile: rows oh0..oh0+BH-1, voxels 0..OWp-1 (zero beyond OH / OW)
        for (int i = threadIdx.x; i < BH * OWp * NCHK; i += 256) {
            const int c = i % NCHK, v = (i / NCHK) % OWp, row = i / (NCHK * OWp);
            const bool ok = oh0 + row < OH && v < OW;
            const bf16* src = dy + ((((size_t)n * OD + od) * OH + oh0 + row) * OW + v) * COUT + c * 8;
            cpa16(dy_u32 + (uint32_t)((row * OWp + v) * NCHK + (c ^ ((v >> SWZ_SH) & (NCHK - 1)))) * 16, ok ? src : dy, ok ? 16 : 0);
        }
        cpa_commit();
        // x halo -> three bf16 planes per w-parity; a work item = (d, h, pair index): 2*S consecutive floats -> S pairs
        const int npair = WP / 2;
        for (int i = threadIdx.x; i < K * HB * npair; i += 256) {
            const int pi = i % npair, h = (i / npair) % HB, d = i / (npair * HB);
            const int xd = od * S + d, xh = oh0 * S + h;
            const bool rok = xh < IH;   // xd < ID always (od < OD)
            const float* xr = x + (((size_t)n * ID + xd) * IH + (rok ? xh : 0)) * IW;
            float v[3 * S];
#pragma unroll
            for (int e = 0; e < 3 * S; e++) {
                const int w = 2 * S * pi + e;
                v[e] = (rok && w < IW) ? __ldg(xr + w) : 0.f;
            }
#pragma unroll
            for (int par = 0; par < S; par++) {
                uint32_t h0, m0, l0, h1, m1, l1, h2, m2, l2;
                split3(v[par], h0, m0, l0);
                split3(v[par + S], h1, m1, l1);
                split3(v[par + 2 * S], h2, m2, l2);
                uint32_t* dst = reinterpret_cast<uint32_t*>(xs + ((par * K + d) * HB + h) * WP + 2 * pi);
                dst[0] = h0 | (h1 << 16);                       // copy 0: (P[2pi], P[2pi+1])
                dst[plane_elems / 2] = m0 | (m1 << 16);
                dst[plane_elems] = l0 | (l1 << 16);
                uint32_t* dso = dst + 3 * plane_elems / 2;      // copy 1: (P[2pi+1], P[2pi+2])
                dso[0] = h1 | (h2 << 16);
                dso[plane_elems / 2] = m1 | (m2 << 16);
                dso[plane_elems] = l1 | (l2 << 16);
            }
        }
        cpa_wait<0>();
        __syncthreads();
        for (int ks = rw; ks < ksteps; ks += RW) {
            const int row = ks / (OWp / 16), w0 = (ks % (OWp / 16)) * 16;
            uint32_t bq[4];
            {
                const int v = w0 + (j & 1) * 8 + r, c = 2 * ng + (j >> 1);
                ldsm4_t(dy_u32 + (uint32_t)((row * OWp + v) * NCHK + (c ^ ((v >> SWZ_SH) & (NCHK - 1)))) * 16, bq);
            }
            const int rbase = row * S * WP + w0 + 2 * t;
#pragma unroll
            for (int sp = 0; sp < 3; sp++) {
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    uint32_t a[4];
#pragma unroll
                    for (int hf = 0; hf < 2; hf++) {
                        const int to = toff[mt][hf];
                        if (to >= 0) {
                            const uint32_t* wp0 = reinterpret_cast<const uint32_t*>(xs + sp * plane_elems + to + rbase);   // even index
                            a[hf] = wp0[0];
                            a[hf + 2] = wp0[4];
                        } else {
                            a[hf] = a[hf + 2] = (to == -2 && sp == 0) ? 0x3F803F80u : 0u;
                        }
                    }
                    mma16816(acc[mt][0], a, bq[0], bq[1]);
                    mma16816(acc[mt][1], a, bq[2], bq[3]);
                }
            }
        }
    }
    // block reduction over the row-warps, then one atomic per output per block
    __syncthreads();
    float* red = reinterpret_cast<float*>(c1w_smem);   // [8 warps][MT*16][16]
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            float* o = red + ((size_t)warp * MT * 16 + mt * 16 + g) * 16 + nt * 8 + 2 * t;
            o[0] = acc[mt][nt][0]; o[1] = acc[mt][nt][1];
            o[8 * 16] = acc[mt][nt][2]; o[8 * 16 + 1] = acc[mt][nt][3];
        }
    __syncthreads();
    for (int i = threadIdx.x; i < NG * MT * 16 * 16; i += 256) {
        const int col = i % 16, tap = (i / 16) % (MT * 16), gq = i / (16 * MT * 16);
        float sacc = 0.f;
#pragma unroll
        for (int w = 0; w < RW; w++) sacc += red[((size_t)(w * NG + gq) * MT * 16 + tap) * 16 + col];
        if (tap < T) atomicAdd(dw + (size_t)tap * COUT + gq * 16 + col, sacc);
        else if (ONES && tap == T && dbias) atomicAdd(dbias + gq * 16 + col, sacc);
    }
}

template <int K, int S, int COUT>
int launch_cin1_wgrad_tc(const float* x, const bf16* dy, float* dw, float* dbias, int N, int ID, int IH, int IW, int OD, int OH, int OW,
                         cudaStream_t st) {
    constexpr int T = K * K * K, MT = (T + 15) / 16, BH = S == 1 ? 8 : 4, HB = (BH - 1) * S + K, NCHK = COUT / 8;
    const int OWp = (OW + 15) / 16 * 16;
    const int WP = (OWp + (K - 1) / S + 2 + 1) & ~1;
    if ((IW + S - 1) / S + 1 > WP) return VG_ERR_UNSUPPORTED;
    const size_t planes = ((size_t)6 * S * K * HB * WP * 2 + 15) / 16 * 16;
    size_t smem = planes + (size_t)BH * OWp * NCHK * 16;
    const size_t red = (size_t)8 * MT * 16 * 16 * 4;
    if (smem < red) smem = red;
    if (smem > 200 * 1024) return VG_ERR_UNSUPPORTED;
    static VgPerDevice attr;
    if (!attr.done()) {
        if (cudaFuncSetAttribute(cin1_wgrad_tc_kernel<K, S, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
            return VG_ERR_CUDA;
        attr.mark();
    }
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm > 3) per_sm = 3;
    if (per_sm < 1) per_sm = 1;
    const int nbricks = N * OD * ((OH + BH - 1) / BH);
    int grid = 148 * per_sm;
    if (grid > nbricks) grid = nbricks;
    cin1_wgrad_tc_kernel<K, S, COUT><<<grid, 256, smem, st>>>(x, dy, dw, dbias, N, ID, IH, IW, OD, OH, OW, WP, OWp);
    VG_LAUNCHED(1);
    return VG_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------- dispatch (called from conv_mma.cu)
// Every function returns VG_ERR_UNSUPPORTED when the shape is not one of its cases; the caller then takes the general path.

int vg_small_k1_fwd(const bf16* x, const bf16* wp, const float* bias, bf16* y, long long nvox, int Cin, int Cout, cudaStream_t st) {
    const int grid = 148 * 3;
    if (Cin == 48 && Cout == 16) k1_fwd_kernel<3, 2><<<grid, 256, 0, st>>>(x, wp, bias, y, nvox);
    else if (Cin == 96 && Cout == 32) k1_fwd_kernel<6, 4><<<148 * 2, 256, 0, st>>>(x, wp, bias, y, nvox);
    else if (Cin == 16 && Cout == 16) k1_fwd_kernel<1, 2><<<grid, 256, 0, st>>>(x, wp, bias, y, nvox);
    else if (Cin == 32 && Cout == 32) k1_fwd_kernel<2, 4><<<grid, 256, 0, st>>>(x, wp, bias, y, nvox);
    else if (Cin == 16 && Cout == 48) k1_fwd_kernel<1, 6><<<grid, 256, 0, st>>>(x, wp, bias, y, nvox);        // input gradients of the
    else if (Cin == 32 && Cout == 96) k1_fwd_kernel<2, 12><<<148 * 2, 256, 0, st>>>(x, wp, bias, y, nvox);    // decoder shortcuts
    else return VG_ERR_UNSUPPORTED;
    VG_LAUNCHED(1);
    return VG_OK;
}

int vg_small_k1_wgrad(const bf16* x, const bf16* dy, float* dw, int N, int ID, int IH, int IW, int Cin, int OD, int OH, int OW, int Cout,
                      int stride, cudaStream_t st) {
    const long long nv = (long long)N * OD * OH * OW;
    if (nv > 0x3fffffff) return VG_ERR_UNSUPPORTED;
    const int nvox = (int)nv;
#define VG_K1W(KS, NT) return launch_k1_wgrad<KS, NT>(x, dy, dw, nvox, stride, OD, OH, OW, ID, IH, IW, st)
    if (Cin == 48 && Cout == 16) VG_K1W(3, 2);
    if (Cin == 96 && Cout == 32) VG_K1W(6, 4);
    if (Cin == 16 && Cout == 32) VG_K1W(1, 4);
    if (Cin == 32 && Cout == 64) VG_K1W(2, 8);
    if (Cin == 16 && Cout == 16) VG_K1W(1, 2);
    if (Cin == 32 && Cout == 32) VG_K1W(2, 4);
#undef VG_K1W
    return VG_ERR_UNSUPPORTED;
}

int vg_small_cout1_k1_fwd(const bf16* x, const bf16* wp, const float* bias, float* y, size_t nvox, int Cin, int act, cudaStream_t st) {
    if (Cin != 16 && Cin != 32) return VG_ERR_UNSUPPORTED;
    const int grid = vg_grid_for((long long)(nvox * (Cin / 8)), 256, 16);
    if (Cin == 16) cout1_k1_fwd_kernel<16><<<grid, 256, 0, st>>>(x, wp, bias, y, nvox, act);
    else cout1_k1_fwd_kernel<32><<<grid, 256, 0, st>>>(x, wp, bias, y, nvox, act);
    VG_LAUNCHED(1);
    return VG_OK;
}

int vg_small_cout1_k1_dgrad(const float* dy, const float* w, bf16* dx, size_t nvox, int Cin, cudaStream_t st) {
    if (Cin % 8 || 256 % (Cin / 8)) return VG_ERR_UNSUPPORTED;
    cout1_k1_dgrad_kernel<<<vg_grid_for((long long)(nvox * (Cin / 8)), 256, 16), 256, 0, st>>>(dy, w, dx, nvox, Cin);
    VG_LAUNCHED(1);
    return VG_OK;
}

int vg_small_cout1_k1_wgrad(const bf16* x, const float* dy, float* dw, size_t nvox, int Cin, cudaStream_t st) {
    if (Cin % 8 || 256 % (Cin / 8)) return VG_ERR_UNSUPPORTED;
    cout1_k1_wgrad_kernel<<<vg_grid_for((long long)(nvox * (Cin / 8)), 256 * 4, 4), 256, 0, st>>>(x, dy, dw, nvox, Cin);
    VG_LAUNCHED(1);
    return VG_OK;
}

int vg_small_cout1_wgrad(const bf16* x, const float* dy, float* dw, int N, int ID, int IH, int IW, int OD, int OH, int OW, int Cin, int K,
                         cudaStream_t st) {
    if (Cin % 8 || Cin > 2048 || 256 % (Cin / 8) || (K != 3 && K != 4)) return VG_ERR_UNSUPPORTED;
    const int rows = N * ID * IH;
    int rpb = (rows * K + 148 * 2 - 1) / (148 * 2);   // ~2 blocks per SM over the (rows, kd) grid (every block ends with K*K*Cin atomics)
    if (rpb < 1) rpb = 1;
    const dim3 grid((rows + rpb - 1) / rpb, K);
    if (K == 3) cout1_wgrad_rows_kernel<3><<<grid, 256, 0, st>>>(x, dy, dw, N, ID, IH, IW, OD, OH, OW, Cin, rpb);
    else cout1_wgrad_rows_kernel<4><<<grid, 256, 0, st>>>(x, dy, dw, N, ID, IH, IW, OD, OH, OW, Cin, rpb);
    VG_LAUNCHED(1);
    return VG_OK;
}

int vg_small_cin1_dgrad_s2(const bf16* dy, const bf16* wd, float* dx, int N, int ID, int IH, int IW, int OD, int OH, int OW, int Cout, int K,
                           cudaStream_t st) {
    if (K != 4 || Cout != 64) return VG_ERR_UNSUPPORTED;
    constexpr int NC = 4;
    constexpr size_t smem = (size_t)5 * 5 * 17 * (NC * 32 + 16);
    static VgPerDevice attr;
    if (!attr.done()) {
        if (cudaFuncSetAttribute(cin1_dgrad_s2_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return VG_ERR_CUDA;
        attr.mark();
    }
    const int GD = (ID + 1) / 2, GH = (IH + 1) / 2, GW = (IW + 1) / 2;
    const int nbd = (GD + 3) / 4, nbh = (GH + 3) / 4, nbw = (GW + 15) / 16;
    const long long nbricks = (long long)N * nbd * nbh * nbw;
    if (nbricks > 0x3fffffff) return VG_ERR_UNSUPPORTED;
    const int grid = nbricks < 148 * 3 ? (int)nbricks : 148 * 3;
    cin1_dgrad_s2_kernel<NC><<<grid, 256, smem, st>>>(dy, wd, dx, N, ID, IH, IW, OD, OH, OW, nbd, nbh, nbw);
    VG_LAUNCHED(1);
    return VG_OK;
}

// Returns VG_OK with *bias_done = 1 when the bias gradient was produced by the same launch.
int vg_small_cin1_wgrad(const float* x, const bf16* dy, float* dw, float* dbias, int N, int ID, int IH, int IW, int OD, int OH, int OW,
                        int Cout, int K, int stride, int* bias_done, cudaStream_t st) {
    *bias_done = 0;
    if (K == 3 && stride == 1 && Cout == 16) {
        int rc = launch_cin1_wgrad_tc<3, 1, 16>(x, dy, dw, dbias, N, ID, IH, IW, OD, OH, OW, st);
        if (rc == VG_OK) *bias_done = 1;
        return rc;
    }
    if (K == 4 && stride == 2 && Cout == 64) return launch_cin1_wgrad_tc<4, 2, 64>(x, dy, dw, nullptr, N, ID, IH, IW, OD, OH, OW, st);
    return VG_ERR_UNSUPPORTED;
}
