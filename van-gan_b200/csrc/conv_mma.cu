// Conv3D forward / dgrad / wgrad as implicit GEMM on bf16 tensor cores (warp-level mma.sync path).
//
// Replaces cuDNN Conv3D / Conv3DBackpropInputV2 / Conv3DBackpropFilterV2 behind the Keras layers at
// resunet_model.py:64-65,89-90,96,127,133-134,245, discriminator.py:63-69,91-114 and
// building_blocks.py:182-189 (reference).  Every convolution on the hot path is a VALID
// convolution over an explicitly padded NDHWC tensor (the padding is written by the producer
// kernel), so one gather-form kernel covers:
//   forward          M = output voxels, N = Cout, K = taps*Cin, source step = stride per voxel
//   dgrad            M = (padded) input voxels of one stride-parity class, N = Cin, K = taps'*Cout,
//                    source step -1 per tap; stride-2 layers are decomposed into 8 parity classes so
//                    no multiply-by-zero work is issued
// A CTA owns a 4x4x8 brick of M (128 rows).  The source halo brick of one K-chunk of channels is
// staged ONCE in shared memory (cp.async, XOR-swizzled 16-byte chunks, zero-filled outside the
// tensor) and re-used by all taps through per-lane ldmatrix row addresses, so the 27x / 64x tap
// re-reads never leave the SM.  Stages are double-buffered across K-chunks.
// wgrad is dW = A^T * dY with the same halo brick as A (ldmatrix.trans), each warp owning a subset
// of the taps, reduction over voxels split across CTAs and finished with fp32 atomics.
//
// The tcgen05/TMEM path for the wide-channel layers lives in conv_tc.cu; this file is the
// general-shape path and the numerical cross-check for it.
#include <stdlib.h>

#include "pack_elem.cuh"
#include "common.cuh"

// conv_k7.cu
int vg_k7_many_to_one(const bf16* M, const bf16* w, int w_tstride, const float* bias, float* y, int N, int MD, int MH, int MW, int C, int YD,
                      int YH, int YW, int K, int sgn, int act, cudaStream_t st);
int vg_k7_wgrad_one(const bf16* M, const float* S, float* dw, int N, int MD, int MH, int MW, int C, int SD, int SH, int SW, int K, int sgn,
                    int off, cudaStream_t st);

namespace {

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void ldsm_x4(uint32_t a, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2(uint32_t a, uint32_t& r0, uint32_t& r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t a, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// XOR swizzle of the 16-byte chunk index inside a row of NCH chunks so that 8 consecutive rows hit
// 8 distinct bank groups (NCH = 2, 4 or 8)
template <int NCH>
__device__ __forceinline__ int swz(int row, int chunk) {
    constexpr int SH = NCH == 8 ? 0 : (NCH == 4 ? 1 : 2);
    return chunk ^ ((row >> SH) & (NCH - 1));
}

constexpr int BD = 4, BH = 4, BW = 8;  // M brick (128 rows)
constexpr int NPAD = 64;               // packed weight rows are padded to a multiple of this

// ------------------------------------------------------------------------------------------ gather conv
struct GConv {
    const bf16* x;
    const bf16* w;  // [T][Np][Cx]
    void* y;
    const float* bias;
    int N, XD, XH, XW, Cx;
    int YD, YH, YW, Cy;
    int GD, GH, GW;
    int TD, TH, TW;
    int so, st;
    int oso, ood, ooh, oow;
    int out_f32, act, Np;
};

template <int CK, int NTW>
__global__ void __launch_bounds__(128) gconv_kernel(GConv p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int NCTA = 8 * NTW, NCH = CK / 8;
    const int ED = (BD - 1) * p.so + p.TD, EH = (BH - 1) * p.so + p.TH, EW = (BW - 1) * p.so + p.TW;
    const int EV = ED * EH * EW;
    const int T = p.TD * p.TH * p.TW;
    const uint32_t xbytes = (uint32_t)EV * CK * 2, wbytes = (uint32_t)T * NCTA * CK * 2;
    const uint32_t stage_bytes = xbytes + wbytes;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const int gtw = (p.GW + BW - 1) / BW, gth = (p.GH + BH - 1) / BH, gtd = (p.GD + BD - 1) / BD;
    int b = blockIdx.x;
    const int tw_ = b % gtw; b /= gtw;
    const int th_ = b % gth; b /= gth;
    const int td_ = b % gtd;
    const int n = b / gtd;
    const int g0d = td_ * BD, g0h = th_ * BH, g0w = tw_ * BW;
    // origin of the source brick
    const int sd0 = g0d * p.so - (p.st < 0 ? p.TD - 1 : 0);
    const int sh0 = g0h * p.so - (p.st < 0 ? p.TH - 1 : 0);
    const int sw0 = g0w * p.so - (p.st < 0 ? p.TW - 1 : 0);
    const int n0 = blockIdx.y * NCTA;
    const uint32_t sbase = smem_u32(smem);
    const bf16* xn = p.x + (size_t)n * p.XD * p.XH * p.XW * p.Cx;

    auto load_stage = [&](int s, int c0) {
        const uint32_t xs = sbase + s * stage_bytes, ws = xs + xbytes;
        for (int i = tid; i < EV * NCH; i += 128) {
            int vox = i / NCH, ch = i % NCH;
            int lw = vox % EW, r = vox / EW;
            int lh = r % EH, ld = r / EH;
            int sd = sd0 + ld, sh = sh0 + lh, sw = sw0 + lw;
            bool ok = (unsigned)sd < (unsigned)p.XD && (unsigned)sh < (unsigned)p.XH && (unsigned)sw < (unsigned)p.XW;
            const bf16* src = ok ? xn + (((size_t)sd * p.XH + sh) * p.XW + sw) * p.Cx + c0 + ch * 8 : p.x;
            cp_async16(xs + (uint32_t)(vox * NCH + swz<NCH>(vox, ch)) * 16, src, ok ? 16 : 0);
        }
        for (int i = tid; i < T * NCTA * NCH; i += 128) {
            int row = i / NCH, ch = i % NCH;
            int t = row / NCTA, nn = row % NCTA;
            const bf16* src = p.w + ((size_t)t * p.Np + n0 + nn) * p.Cx + c0 + ch * 8;
            cp_async16(ws + (uint32_t)(row * NCH + swz<NCH>(row, ch)) * 16, src, 16);
        }
    };

    float acc[2][NTW][4];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int c = 0; c < NTW; c++)
#pragma unroll
            for (int k = 0; k < 4; k++) acc[a][c][k] = 0.f;

    // per-lane A row (ldmatrix address provider): matrix mi = lane/8, row r = lane%8
    const int mi = lane >> 3, r8 = lane & 7;
    int vbase[2];
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
        int lh = mt * 2 + (mi & 1);
        vbase[mt] = ((warp * p.so) * EH + lh * p.so) * EW + r8 * p.so;
    }
    const int a_chunk = mi >> 1;
    // B provider: x4 -> (ntile pair, kchunk): matrix mi: ntile = mi>>1, kchunk = mi&1
    const int b_row = (mi >> 1) * 8 + r8, b_chunk = mi & 1;

    const int nchunks = p.Cx / CK;
    load_stage(0, 0);
    cp_async_commit();
    for (int c = 0; c < nchunks; c++) {
        if (c + 1 < nchunks) {
            load_stage((c + 1) & 1, (c + 1) * CK);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const uint32_t xs = sbase + (c & 1) * stage_bytes, ws = xs + xbytes;
        int t = 0;
        for (int td = 0; td < p.TD; td++)
            for (int th = 0; th < p.TH; th++)
                for (int tw = 0; tw < p.TW; tw++, t++) {
                    const int od = p.st > 0 ? td : p.TD - 1 - td, oh = p.st > 0 ? th : p.TH - 1 - th,
                              ow = p.st > 0 ? tw : p.TW - 1 - tw;
                    const int toff = (od * EH + oh) * EW + ow;
#pragma unroll
                    for (int kk = 0; kk < CK / 16; kk++) {
                        uint32_t a[2][4];
#pragma unroll
                        for (int mt = 0; mt < 2; mt++) {
                            int vox = vbase[mt] + toff;
                            ldsm_x4(xs + (uint32_t)(vox * NCH + swz<NCH>(vox, kk * 2 + a_chunk)) * 16, a[mt][0], a[mt][1],
                                    a[mt][2], a[mt][3]);
                        }
                        if constexpr (NTW == 1) {
                            int row = t * NCTA + r8;
                            uint32_t b0, b1;
                            ldsm_x2(ws + (uint32_t)(row * NCH + swz<NCH>(row, kk * 2 + (mi & 1))) * 16, b0, b1);
                            mma_bf16(acc[0][0], a[0], b0, b1);
                            mma_bf16(acc[1][0], a[1], b0, b1);
                        } else {
#pragma unroll
                            for (int np = 0; np < NTW / 2; np++) {
                                int row = t * NCTA + np * 16 + b_row;
                                uint32_t b0, b1, b2, b3;
                                ldsm_x4(ws + (uint32_t)(row * NCH + swz<NCH>(row, kk * 2 + b_chunk)) * 16, b0, b1, b2, b3);
                                mma_bf16(acc[0][2 * np], a[0], b0, b1);
                                mma_bf16(acc[1][2 * np], a[1], b0, b1);
                                mma_bf16(acc[0][2 * np + 1], a[0], b2, b3);
                                mma_bf16(acc[1][2 * np + 1], a[1], b2, b3);
                            }
                        }
                    }
                }
        __syncthreads();
    }

    // epilogue: c0,c1 -> row lane/4, cols (lane%4)*2+{0,1}; c2,c3 -> row+8
    const int gd = g0d + warp;
    if (gd >= p.GD) return;
    const int yd = gd * p.oso + p.ood;
    const int gw = g0w + (lane >> 2);
    const int yw = gw * p.oso + p.oow;
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
            const int gh = g0h + mt * 2 + hf;
            if (gh >= p.GH || gw >= p.GW) continue;
            const int yh = gh * p.oso + p.ooh;
            if (yd >= p.YD || yh >= p.YH || yw >= p.YW) continue;
            const size_t row = ((((size_t)n * p.YD + yd) * p.YH + yh) * p.YW + yw) * p.Cy;
#pragma unroll
            for (int nt = 0; nt < NTW; nt++) {
                const int col = n0 + nt * 8 + (lane & 3) * 2;
                float v0 = acc[mt][nt][hf * 2], v1 = acc[mt][nt][hf * 2 + 1];
                if (p.bias) {
                    if (col < p.Cy) v0 += p.bias[col];
                    if (col + 1 < p.Cy) v1 += p.bias[col + 1];
                }
                if (p.act == VG_ACT_TANH) { v0 = tanhf(v0); v1 = tanhf(v1); }
                if (p.out_f32) {
                    float* o = (float*)p.y + row + col;
                    if (col < p.Cy) o[0] = v0;
                    if (col + 1 < p.Cy) o[1] = v1;
                } else {
                    bf16* o = (bf16*)p.y + row + col;
                    if (col + 1 < p.Cy) *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(v0, v1);
                    else if (col < p.Cy) o[0] = __float2bfloat16(v0);
                }
            }
        }
}

struct GCfg {
    int ck, ntw, stages;
    size_t smem;
};

inline bool pick_gconv_cfg(const GConv& p, int ncols, GCfg& out) {
    const int ED = (BD - 1) * p.so + p.TD, EH = (BH - 1) * p.so + p.TH, EW = (BW - 1) * p.so + p.TW;
    const size_t EV = (size_t)ED * EH * EW, T = (size_t)p.TD * p.TH * p.TW;
    const int ntws[4] = {8, 4, 2, 1}, cks[3] = {64, 32, 16};
    int n8 = (ncols + 7) / 8;
    for (int a = 0; a < 4; a++) {
        int ntw = ntws[a];
        if (ntw > n8 && ntw != 1 && (ntw / 2) >= n8) continue;  // do not over-pad N
        for (int b = 0; b < 3; b++) {
            int ck = cks[b];
            if (p.Cx % ck) continue;
            size_t bytes = EV * ck * 2 + T * 8 * ntw * ck * 2;
            int stages = (p.Cx / ck) > 1 ? 2 : 1;
            if (bytes * stages <= 200 * 1024) {
                out = {ck, ntw, stages, bytes * stages};
                return true;
            }
        }
    }
    return false;
}

template <int CK, int NTW>
int launch_gconv_t(const GConv& p, const GCfg& c, cudaStream_t st) {
    static VgPerDevice attr_done;
    if (!attr_done.done()) {
        if (cudaFuncSetAttribute(gconv_kernel<CK, NTW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess)
            return VG_ERR_CUDA;
        attr_done.mark();
    }
    const int gtw = (p.GW + BW - 1) / BW, gth = (p.GH + BH - 1) / BH, gtd = (p.GD + BD - 1) / BD;
    dim3 grid((unsigned)((size_t)gtw * gth * gtd * p.N), (unsigned)((p.Cy + 8 * NTW - 1) / (8 * NTW)));
    gconv_kernel<CK, NTW><<<grid, 128, c.smem, st>>>(p); VG_LAUNCHED(1);
    return VG_OK;
}

int launch_gconv(const GConv& p, cudaStream_t st) {
    GCfg c;
    if (p.Cx % 16 != 0 || !pick_gconv_cfg(p, p.Cy, c)) return VG_ERR_UNSUPPORTED;
#define VG_CASE(CK, NTW) \
    if (c.ck == CK && c.ntw == NTW) return launch_gconv_t<CK, NTW>(p, c, st);
    VG_CASE(16, 1) VG_CASE(16, 2) VG_CASE(16, 4) VG_CASE(16, 8)
    VG_CASE(32, 1) VG_CASE(32, 2) VG_CASE(32, 4) VG_CASE(32, 8)
    VG_CASE(64, 1) VG_CASE(64, 2) VG_CASE(64, 4) VG_CASE(64, 8)
#undef VG_CASE
    return VG_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------ wgrad
struct WGrad {
    const bf16* x;
    const bf16* dy;
    float* dw;
    int N, XD, XH, XW, Cx;
    int OD, OH, OW, Cy;
    int K, so;
    int nbricks, gtd, gth, gtw;
};

// 256 threads; warp w owns taps w, w+8, ... (TPW per warp).  For K==1 (one tap) the warps split the
// eight 16-voxel k-steps of the brick instead.
template <int TPW, int NCW>
__global__ void __launch_bounds__(256) wgrad_kernel(WGrad p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int CKI = 16, NCHX = 2, NC = 8 * NCW, NCHY = NC / 8;
    const int E_D = (BD - 1) * p.so + p.K, E_H = (BH - 1) * p.so + p.K, E_W = (BW - 1) * p.so + p.K;
    const int EV = E_D * E_H * E_W;
    const int T = p.K * p.K * p.K;
    const uint32_t xbytes = (uint32_t)EV * CKI * 2, ybytes = 128 * NC * 2;
    const uint32_t stage_bytes = xbytes + ybytes;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ci0 = blockIdx.y * CKI, co0 = blockIdx.z * NC;
    const uint32_t sbase = smem_u32(smem);

    auto load_stage = [&](int s, int brick) {
        int b = brick;
        const int tw_ = b % p.gtw; b /= p.gtw;
        const int th_ = b % p.gth; b /= p.gth;
        const int td_ = b % p.gtd;
        const int n = b / p.gtd;
        const int g0d = td_ * BD, g0h = th_ * BH, g0w = tw_ * BW;
        const uint32_t xs = sbase + s * stage_bytes, ys = xs + xbytes;
        const bf16* xn = p.x + (size_t)n * p.XD * p.XH * p.XW * p.Cx;
        for (int i = tid; i < EV * NCHX; i += 256) {
            int vox = i / NCHX, ch = i % NCHX;
            int lw = vox % E_W, r = vox / E_W;
            int lh = r % E_H, ld = r / E_H;
            int sd = g0d * p.so + ld, sh = g0h * p.so + lh, sw = g0w * p.so + lw;
            bool ok = sd < p.XD && sh < p.XH && sw < p.XW;
            const bf16* src = ok ? xn + (((size_t)sd * p.XH + sh) * p.XW + sw) * p.Cx + ci0 + ch * 8 : p.x;
            cp_async16(xs + (uint32_t)(vox * NCHX + swz<NCHX>(vox, ch)) * 16, src, ok ? 16 : 0);
        }
        const bf16* yn = p.dy + (size_t)n * p.OD * p.OH * p.OW * p.Cy;
        for (int i = tid; i < 128 * NCHY; i += 256) {
            int m = i / NCHY, ch = i % NCHY;
            int lw = m % BW, lh = (m / BW) % BH, ld = m / (BW * BH);
            int od = g0d + ld, oh = g0h + lh, ow = g0w + lw;
            bool ok = od < p.OD && oh < p.OH && ow < p.OW;   // rows outside the output are zero -> no contribution
            const bf16* src = ok ? yn + (((size_t)od * p.OH + oh) * p.OW + ow) * p.Cy + co0 + ch * 8 : p.dy;
            cp_async16(ys + (uint32_t)(m * NCHY + swz<NCHY>(m, ch)) * 16, src, ok ? 16 : 0);
        }
    };

    float acc[TPW][NCW][4];
#pragma unroll
    for (int a = 0; a < TPW; a++)
#pragma unroll
        for (int c = 0; c < NCW; c++)
#pragma unroll
            for (int k = 0; k < 4; k++) acc[a][c][k] = 0.f;

    const int mi = lane >> 3, r8 = lane & 7;
    const bool ksplit = (T == 1);

    int brick = blockIdx.x;
    if (brick < p.nbricks) {
        load_stage(0, brick);
        cp_async_commit();
    }
    int it = 0;
    for (; brick < p.nbricks; brick += gridDim.x, it++) {
        int nxt = brick + gridDim.x;
        if (nxt < p.nbricks) {
            load_stage((it + 1) & 1, nxt);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const uint32_t xs = sbase + (it & 1) * stage_bytes, ys = xs + xbytes;
#pragma unroll 1
        for (int ks = 0; ks < 8; ks++) {
            if (ksplit && (ks & 7) != warp) continue;
            // B fragments (dY, k = 16 voxels of k-step ks): rows m = ks*16 + (mi&1)*8 + r8, chunk = j + (mi>>1)
            uint32_t bfr[NCW][2];
#pragma unroll
            for (int j = 0; j < NCW; j += 2) {
                int m = ks * 16 + (mi & 1) * 8 + r8;
                int ch = j + (mi >> 1);
                if constexpr (NCW == 1) ch = 0;
                uint32_t b0, b1, b2, b3;
                ldsm_x4_t(ys + (uint32_t)(m * NCHY + swz<NCHY>(m, ch % NCHY)) * 16, b0, b1, b2, b3);
                bfr[j][0] = b0; bfr[j][1] = b1;
                if (j + 1 < NCW) { bfr[j + 1][0] = b2; bfr[j + 1][1] = b3; }
            }
            const int ld = ks >> 1, lh = (ks & 1) * 2 + (mi >> 1), lw = r8;
            const int vb = ((ld * p.so) * E_H + lh * p.so) * E_W + lw * p.so;
#pragma unroll
            for (int a = 0; a < TPW; a++) {
                int t = ksplit ? 0 : warp + a * 8;
                if (t >= T) break;
                int tw = t % p.K, th = (t / p.K) % p.K, td = t / (p.K * p.K);
                int vox = vb + (td * E_H + th) * E_W + tw;
                uint32_t af[4];
                ldsm_x4_t(xs + (uint32_t)(vox * NCHX + swz<NCHX>(vox, mi & 1)) * 16, af[0], af[1], af[2], af[3]);
#pragma unroll
                for (int j = 0; j < NCW; j++) mma_bf16(acc[a][j], af, bfr[j][0], bfr[j][1]);
            }
        }
        __syncthreads();
    }
    // dw[t][ci][co] += acc
#pragma unroll
    for (int a = 0; a < TPW; a++) {
        int t = ksplit ? 0 : warp + a * 8;
        if (t >= T) break;
#pragma unroll
        for (int j = 0; j < NCW; j++) {
            int col = co0 + j * 8 + (lane & 3) * 2;
            int row = ci0 + (lane >> 2);
            float* o = p.dw + ((size_t)t * p.Cx + row) * p.Cy + col;
            atomicAdd(o, acc[a][j][0]);
            atomicAdd(o + 1, acc[a][j][1]);
            atomicAdd(o + (size_t)8 * p.Cy, acc[a][j][2]);
            atomicAdd(o + (size_t)8 * p.Cy + 1, acc[a][j][3]);
        }
    }
}

template <int TPW, int NCW>
int launch_wgrad_t(const WGrad& p, cudaStream_t st) {
    static VgPerDevice attr_done;
    if (!attr_done.done()) {
        if (cudaFuncSetAttribute(wgrad_kernel<TPW, NCW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess)
            return VG_ERR_CUDA;
        attr_done.mark();
    }
    const int E_D = (BD - 1) * p.so + p.K, E_H = (BH - 1) * p.so + p.K, E_W = (BW - 1) * p.so + p.K;
    size_t stage = (size_t)E_D * E_H * E_W * 32 + 128 * 8 * NCW * 2;
    int tiles = (p.Cx / 16) * (p.Cy / (8 * NCW));
    int nsplit = (148 * 4 + tiles - 1) / tiles;
    if (nsplit > p.nbricks) nsplit = p.nbricks;
    if (nsplit < 1) nsplit = 1;
    dim3 grid(nsplit, p.Cx / 16, p.Cy / (8 * NCW));
    wgrad_kernel<TPW, NCW><<<grid, 256, 2 * stage, st>>>(p); VG_LAUNCHED(1);
    return VG_OK;
}

// ------------------------------------------------------------------------------------------ weight packing
// fwd pack: Wf[t][Np][Cin] (row = output channel, K contiguous); dgrad pack: per parity class,
// Wd[class][t'][NpI][Cout] (row = input channel, K = output channel contiguous)
// one job per launch (per-layer API) ...
__global__ void pack_job_kernel(const vg_pack_job job) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)job.total; i += (size_t)gridDim.x * blockDim.x)
        job.out[i] = __float2bfloat16(pack_elem(job, i));
}

// ... or every job of a network in ONE launch.  The packed elements of all jobs form one global index range; a block takes a
// contiguous span of it (so a 16x16 shortcut and the 8.4 M-element PatchGAN layer load the grid equally), finds the job of its first
// element by binary search in the prefix array and walks on from there.
constexpr int PACK_SPAN = 4096;   // elements per block iteration
__global__ void __launch_bounds__(256) pack_multi_kernel(const vg_pack_job* __restrict__ jobs, const long long* __restrict__ prefix, int njobs,
                                                         long long total) {
    for (long long s0 = (long long)blockIdx.x * PACK_SPAN; s0 < total; s0 += (long long)gridDim.x * PACK_SPAN) {
        const long long s1 = s0 + PACK_SPAN < total ? s0 + PACK_SPAN : total;
        int lo = 0, hi = njobs - 1;                 // last job with prefix[j] <= s0
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (prefix[mid] <= s0) lo = mid; else hi = mid - 1;
        }
        for (int j = lo; j < njobs && prefix[j] < s1; j++) {
            const vg_pack_job job = jobs[j];
            const long long b = prefix[j] > s0 ? prefix[j] : s0;
            const long long e = prefix[j] + job.total < s1 ? prefix[j] + job.total : s1;
            // runs of 8 packed elements (innermost index of every layout: 8 consecutive k, or 8 consecutive channels) share their
            // validity and read the source at one constant stride: two index decodes, eight loads, one 16-byte store per run
            if (((b - prefix[j]) | (e - prefix[j])) & 7) {
                for (long long g = b + threadIdx.x; g < e; g += 256) {
                    const size_t i = (size_t)(g - prefix[j]);
                    job.out[i] = __float2bfloat16(pack_elem(job, i));
                }
            } else {
                for (long long g = b + 8LL * threadIdx.x; g < e; g += 8 * 256) {
                    const size_t i = (size_t)(g - prefix[j]);
                    const long long s0_ = pack_src(job, i);
                    uint4 o = make_uint4(0u, 0u, 0u, 0u);
                    if (s0_ >= 0) {
                        const long long st = pack_src(job, i + 1) - s0_;
                        float f[8];
#pragma unroll
                        for (int k = 0; k < 8; k++) f[k] = job.w[s0_ + k * st];
                        o.x = pack2_bf16(f[0], f[1]); o.y = pack2_bf16(f[2], f[3]); o.z = pack2_bf16(f[4], f[5]); o.w = pack2_bf16(f[6], f[7]);
                    }
                    *reinterpret_cast<uint4*>(job.out + i) = o;
                }
            }
        }
    }
}

inline int class_taps(int K, int stride, int a) { return K <= a ? 0 : (K - a + stride - 1) / stride; }

// ------------------------------------------------------------------------------------------ direct kernels (Cin==1 / Cout==1)
// forward, Cin == 1: x fp32 [N,ID,IH,IW], w fp32 [T][Cout], y bf16 [N,OD,OH,OW,Cout].
// Register tile: a thread owns VW output voxels and 16 output channels (64 accumulators); every weight vector (16 floats,
// broadcast from shared memory) feeds 16*VW FMAs.  The VW voxels of a thread are 32 apart in the flattened (oh, ow) plane, so
// the 32 lanes of a warp always touch consecutive voxels: input loads are coalesced and each output voxel row (16 channels =
// one 32-byte sector) leaves as a single 256-bit store, 1 KB contiguous per warp instruction when Cout == 16.
template <int K, int S, bool PF = false>
__global__ void __launch_bounds__(128, PF ? 3 : 1) cin1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, bf16* __restrict__ y, int N, int ID,
                                                       int IH, int IW, int OD, int OH, int OW, int Cout) {
    constexpr int VW = 4, T = K * K * K;
    extern __shared__ float sw[];  // [T][Cout] + bias[Cout]
    for (int i = threadIdx.x; i < T * Cout; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[T * Cout + i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int ncg = Cout / 16, plane = OH * OW, nfb = (plane + 32 * VW - 1) / (32 * VW);
    const long long total = (long long)N * OD * nfb * ncg;
    for (long long i = (long long)blockIdx.x * nwarp + wid; i < total; i += (long long)gridDim.x * nwarp) {
        const int cg = (int)(i % ncg);
        long long r = i / ncg;
        const int fb = (int)(r % nfb); r /= nfb;
        const int od = (int)(r % OD);
        const int n = (int)(r / OD);
        int xoff[VW];
        bool ok[VW];
#pragma unroll
        for (int v = 0; v < VW; v++) {
            const int f = fb * 32 * VW + v * 32 + lane;
            ok[v] = f < plane;
            const int fc = ok[v] ? f : 0;
            const int oh = fc / OW, ow = fc - oh * OW;
            xoff[v] = oh * S * IW + ow * S;
        }
        float acc[VW][16];
#pragma unroll
        for (int v = 0; v < VW; v++)
#pragma unroll
            for (int k = 0; k < 16; k++) acc[v][k] = sw[T * Cout + cg * 16 + k];
        const float* xb = x + ((size_t)n * ID + od * S) * IH * IW;
        auto load_row = [&](int q, float (&xv)[VW][K]) {   // the K inputs along w of every owned voxel for tap row q = (kd, kh)
            const float* xr = xb + ((size_t)(q / K) * IH + (q % K)) * IW;
#pragma unroll
            for (int v = 0; v < VW; v++)
#pragma unroll
                for (int kw = 0; kw < K; kw++) xv[v][kw] = __ldg(xr + xoff[v] + kw);
        };
        auto fma_row = [&](int q, const float (&xv)[VW][K]) {
            const float* wr = sw + (q * K) * Cout + cg * 16;
#pragma unroll
            for (int kw = 0; kw < K; kw++) {
                float wv[16];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float4 t4 = *reinterpret_cast<const float4*>(wr + kw * Cout + j * 4);
                    wv[4 * j] = t4.x; wv[4 * j + 1] = t4.y; wv[4 * j + 2] = t4.z; wv[4 * j + 3] = t4.w;
                }
#pragma unroll
                for (int v = 0; v < VW; v++)
#pragma unroll
                    for (int k = 0; k < 16; k++) acc[v][k] = fmaf(xv[v][kw], wv[k], acc[v][k]);
            }
        };
        if (PF) {
            // software pipeline over the K*K tap rows: the loads of row q+1 are in flight while row q feeds the FMA pipe
            float xa[VW][K], xb2[VW][K];
            load_row(0, xa);
#pragma unroll 1
            for (int q = 0; q < K * K; q += 2) {
                if (q + 1 < K * K) load_row(q + 1, xb2);
                fma_row(q, xa);
                if (q + 1 < K * K) {
                    if (q + 2 < K * K) load_row(q + 2, xa);
                    fma_row(q + 1, xb2);
                }
            }
        } else {
#pragma unroll 1
            for (int q = 0; q < K * K; q++) {
                float xv[VW][K];
                load_row(q, xv);
                fma_row(q, xv);
            }
        }
        bf16* yo = y + (((size_t)n * OD + od) * plane + (size_t)fb * 32 * VW + lane) * Cout + cg * 16;
#pragma unroll
        for (int v = 0; v < VW; v++)
            if (ok[v]) {
                uint32_t o[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[v][2 * j], acc[v][2 * j + 1]);
                    o[j] = *reinterpret_cast<uint32_t*>(&h2);
                }
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(yo + (size_t)v * 32 * Cout), "r"(o[0]),
                             "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
            }
    }
}

// wgrad, Cin == 1: dw[t][co] += sum_o x[o*s+t] * dy[o][co];  dbias[co] += sum_o dy[o][co] (optional, fused).
// A task is (kd, kh, 16 output channels) and belongs to one WARP; the 32 lanes are 32 consecutive voxels of the flattened
// (oh, ow) plane, so the dy load (one 32-byte sector per lane, a 256-bit load) and the K input loads are coalesced, and the
// warps of a block -- which walk the same voxels with different tasks -- share dy through L1.  K*16 accumulators per lane,
// reduced over the lanes by shuffles once at the end, then one atomicAdd per (warp, output).
template <int K, int S>
__global__ void __launch_bounds__(K >= 4 ? 512 : 288, K >= 4 ? 1 : 2)
    cin1_wgrad_kernel(const float* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dw, float* __restrict__ dbias,
                      int N, int ID, int IH, int IW, int OD, int OH, int OW, int Cout) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int ncg = Cout / 16;
    const int task = blockIdx.y * wpb + wid;          // host guarantees gridDim.y * wpb == K*K*ncg
    const int cg = task % ncg, kh = (task / ncg) % K, kd = task / (ncg * K);
    const int rows = N * OD * OH;
    float acc[K][16], bsum[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        bsum[k] = 0.f;
#pragma unroll
        for (int q = 0; q < K; q++) acc[q][k] = 0.f;
    }
    const bool do_bias = dbias != nullptr && kd == 0 && kh == 0;
    constexpr int U = K >= 3 ? 2 : 4;   // 32-voxel chunks of one output row in flight (all loads issued before any FMA)
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int oh = row % OH, r = row / OH;
        const int od = r % OD, n = r / OD;
        const float* xr = x + (((size_t)n * ID + od * S + kd) * IH + oh * S + kh) * IW + lane * S;
        const bf16* dp = dy + ((size_t)row * OW + lane) * Cout + cg * 16;
        for (int w0 = 0; w0 < OW; w0 += 32 * U) {
            float xw[U][K];
            uint32_t raw[U][8];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int ow = w0 + u * 32;
                if (ow + lane < OW) {
#pragma unroll
                    for (int q = 0; q < K; q++) xw[u][q] = __ldg(xr + ow * S + q);
                    asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                                 : "=r"(raw[u][0]), "=r"(raw[u][1]), "=r"(raw[u][2]), "=r"(raw[u][3]), "=r"(raw[u][4]), "=r"(raw[u][5]),
                                   "=r"(raw[u][6]), "=r"(raw[u][7])
                                 : "l"(dp + (size_t)ow * Cout));
                } else {
#pragma unroll
                    for (int q = 0; q < K; q++) xw[u][q] = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; j++) raw[u][j] = 0u;
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                float g[16];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    g[2 * j] = __uint_as_float(raw[u][j] << 16);
                    g[2 * j + 1] = __uint_as_float(raw[u][j] & 0xffff0000u);
                }
#pragma unroll
                for (int q = 0; q < K; q++)
#pragma unroll
                    for (int k = 0; k < 16; k++) acc[q][k] = fmaf(xw[u][q], g[k], acc[q][k]);
                if (do_bias) {
#pragma unroll
                    for (int k = 0; k < 16; k++) bsum[k] += g[k];
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < K; q++)
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const float v = warp_sum(acc[q][k]);
            if (lane == 0) atomicAdd(dw + (size_t)((kd * K + kh) * K + q) * Cout + cg * 16 + k, v);
        }
    if (do_bias) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const float v = warp_sum(bsum[k]);
            if (lane == 0) atomicAdd(dbias + cg * 16 + k, v);
        }
    }
}

// dgrad, Cout == 1: dx[p][ci] = sum_t dy[p - t] * w[t][ci]   (stride 1 only; dy fp32 [N,OD,OH,OW], dx bf16)
__global__ void __launch_bounds__(256) cout1_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                          bf16* __restrict__ dx, int N, int ID, int IH, int IW, int OD, int OH,
                                                          int OW, int Cin, int K) {
    const int cg = Cin / 8;
    size_t total = (size_t)N * ID * IH * IW * cg;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        int c8 = (int)(i % cg);
        size_t v = i / cg;
        int pw = (int)(v % IW), ph = (int)((v / IW) % IH), pd = (int)((v / ((size_t)IW * IH)) % ID);
        int n = (int)(v / ((size_t)IW * IH * ID));
        float a[8];
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = 0.f;
        for (int kd = 0; kd < K; kd++) {
            int od = pd - kd;
            if ((unsigned)od >= (unsigned)OD) continue;
            for (int kh = 0; kh < K; kh++) {
                int oh = ph - kh;
                if ((unsigned)oh >= (unsigned)OH) continue;
                for (int kw = 0; kw < K; kw++) {
                    int ow = pw - kw;
                    if ((unsigned)ow >= (unsigned)OW) continue;
                    float g = __ldg(dy + (((size_t)n * OD + od) * OH + oh) * OW + ow);
                    float wv[8];
                    load8<float>(w + (size_t)((kd * K + kh) * K + kw) * Cin + c8 * 8, wv);   // two 128-bit loads
#pragma unroll
                    for (int k = 0; k < 8; k++) a[k] = fmaf(g, wv[k], a[k]);
                }
            }
        }
        store8<bf16>(dx + i * 8, a);
    }
}

// wgrad, Cout == 1: dw[t][ci] += sum_o x[o+t][ci] * dy[o]   (stride 1; x bf16, dy fp32)
__global__ void __launch_bounds__(256) cout1_wgrad_kernel(const bf16* __restrict__ x, const float* __restrict__ dy,
                                                          float* __restrict__ dw, int N, int ID, int IH, int IW, int OD, int OH,
                                                          int OW, int Cin, int K, int per_block) {
    // blockIdx.y = tap; threads stride over (voxel, channel-group); block-level smem reduction per channel
    extern __shared__ float sred[];  // [256/cg][Cin]
    const int t = blockIdx.y, kw = t % K, kh = (t / K) % K, kd = t / (K * K);
    const int cg = Cin / 8, nvl = 256 / cg;
    const int c8 = threadIdx.x % cg, vl = threadIdx.x / cg;
    size_t V = (size_t)N * OD * OH * OW;
    size_t v0 = (size_t)blockIdx.x * per_block, v1 = v0 + per_block < V ? v0 + per_block : V;
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = 0.f;
    if (vl < nvl)
        for (size_t v = v0 + vl; v < v1; v += nvl) {
            int ow = (int)(v % OW), oh = (int)((v / OW) % OH), od = (int)((v / ((size_t)OW * OH)) % OD);
            int n = (int)(v / ((size_t)OW * OH * OD));
            float g = __ldg(dy + v);
            float f[8];
            load8<bf16>(x + ((((size_t)n * ID + od + kd) * IH + oh + kh) * IW + ow + kw) * Cin + c8 * 8, f);
#pragma unroll
            for (int k = 0; k < 8; k++) a[k] = fmaf(g, f[k], a[k]);
        }
    if (vl < nvl) {
#pragma unroll
        for (int k = 0; k < 8; k++) sred[(size_t)vl * Cin + c8 * 8 + k] = a[k];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < Cin; c += 256) {
        float s = 0.f;
        for (int l = 0; l < nvl; l++) s += sred[(size_t)l * Cin + c];
        atomicAdd(dw + (size_t)t * Cin + c, s);
    }
}

// dbias[c] += sum over rows of dy[row][c]   (T = bf16 with C%8==0, or float with C==1)
template <typename T>
__global__ void __launch_bounds__(256) channel_sum_kernel(const T* __restrict__ dy, size_t rows, int C, float* __restrict__ out) {
    extern __shared__ float sred[];
    if (C == 1) {
        float s = 0.f;
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < rows; i += (size_t)gridDim.x * 256) s += (float)dy[i];
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tot = 0.f;
            for (int i = 0; i < 8; i++) tot += sred[i];
            atomicAdd(out, tot);
        }
        return;
    }
    const int cg = C / 8, nvl = 256 / cg;
    const int c8 = threadIdx.x % cg, vl = threadIdx.x / cg;
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = 0.f;
    if (vl < nvl) {
        // four independent 128-bit loads in flight per thread (the first version issued one per iteration: 43 % of HBM)
        const size_t step = (size_t)gridDim.x * nvl;
        size_t r = (size_t)blockIdx.x * nvl + vl;
        for (; r + 3 * step < rows; r += 4 * step) {
            float f[4][8];
#pragma unroll
            for (int u = 0; u < 4; u++) load8<T>(dy + (r + u * step) * C + c8 * 8, f[u]);
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int k = 0; k < 8; k++) a[k] += f[u][k];
        }
        for (; r < rows; r += step) {
            float f[8];
            load8<T>(dy + r * C + c8 * 8, f);
#pragma unroll
            for (int k = 0; k < 8; k++) a[k] += f[k];
        }
    }
    if (vl < nvl) {
#pragma unroll
        for (int k = 0; k < 8; k++) sred[(size_t)vl * C + c8 * 8 + k] = a[k];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float s = 0.f;
        for (int l = 0; l < nvl; l++) s += sred[(size_t)l * C + c];
        atomicAdd(out + c, s);
    }
}

inline bool desc_ok(const vg_conv3d_desc* d) {
    if (!d || d->N <= 0 || d->Cin <= 0 || d->Cout <= 0) return false;
    if (!(d->K == 1 || d->K == 3 || d->K == 4 || d->K == 7)) return false;
    if (d->K == 7 && !(d->stride == 1 && (d->Cin == 1 || d->Cout == 1))) return false;   // 'resnet' generator: conv_k7.cu
    if (!(d->stride == 1 || d->stride == 2)) return false;
    if (d->ID < d->K || d->IH < d->K || d->IW < d->K) return false;
    if (d->Cin == 1 ? d->x_dtype != VG_F32 : (d->x_dtype != VG_BF16 || d->Cin % 16)) return false;
    if (d->Cout == 1 ? d->y_dtype != VG_F32 : (d->y_dtype != VG_BF16 || d->Cout % 16)) return false;
    if (d->Cin == 1 && d->Cout == 1) return false;
    if (d->dx_lo < 0 || d->dx_hi < 0 || d->dx_lo + d->dx_hi >= d->ID) return false;
    return true;
}
inline int odim(int I, int K, int s) { return (I - K) / s + 1; }
inline int rup(int a, int b) { return (a + b - 1) / b * b; }

// VG_CONV_PATH=mma forces the warp-level mma.sync kernels everywhere (A/B testing and cross-checks)
inline bool tc_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VG_CONV_PATH");
        v = (e && e[0] == 'm') ? 0 : 1;
    }
    return v == 1;
}
// VG_WGRAD_PATH=mma forces the mma.sync weight-gradient kernel (A/B testing and cross-checks)
inline bool tc_wgrad_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VG_WGRAD_PATH");
        v = (e && e[0] == 'm') ? 0 : 1;
    }
    return v == 1 && tc_enabled();
}
// VG_SMALL=0 disables the streaming kernels of conv_small.cu (A/B testing and cross-checks)
inline bool small_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VG_SMALL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}
inline size_t mma_fwd_elems(const vg_conv3d_desc* d) {
    return d->Cin == 1 ? 0 : (size_t)d->K * d->K * d->K * rup(d->Cout, NPAD) * d->Cin;
}
inline size_t mma_dgrad_elems(const vg_conv3d_desc* d) {
    if (d->Cout == 1) return 0;
    size_t tot = 0;
    for (int a = 0; a < d->stride; a++)
        for (int b = 0; b < d->stride; b++)
            for (int c = 0; c < d->stride; c++)
                tot += (size_t)class_taps(d->K, d->stride, a) * class_taps(d->K, d->stride, b) * class_taps(d->K, d->stride, c);
    return tot * rup(d->Cin, NPAD) * d->Cout;
}
// stride 2 (K = 3 or 4): run as 8 parity classes of 2x2x2 taps over strided views of x (K = 3 is zero-padded to 4)
inline bool tc_fwd_s2(const vg_conv3d_desc* d) { return d->stride == 2 && d->K >= 3; }
inline int tc_fwd_taps(const vg_conv3d_desc* d) { return tc_fwd_s2(d) ? 8 : d->K * d->K * d->K; }
inline int tc_fwd_kdim(const vg_conv3d_desc* d) { return tc_fwd_s2(d) ? 8 * d->Cin : d->Cin; }
inline bool tc_fwd_ok(const vg_conv3d_desc* d) {
    return (d->stride == 1 || tc_fwd_s2(d)) && d->Cin % 16 == 0 && d->Cout % 16 == 0 && vg_tc_ncta(d->Cout, tc_fwd_taps(d)) > 0;
}
inline bool tc_dgrad_ok(const vg_conv3d_desc* d) { return d->Cin % 16 == 0 && d->Cout % 16 == 0; }
inline size_t tc_dgrad_class_elems(const vg_conv3d_desc* d, int a, int b, int c) {
    int T = class_taps(d->K, d->stride, a) * class_taps(d->K, d->stride, b) * class_taps(d->K, d->stride, c);
    return T ? vg_tc_pack_elems(d->Cin, d->Cout, T) : 0;
}
inline size_t rup256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" {

int vg_abi_version(void) { return 4; }

size_t vg_conv3d_packed_bytes(const vg_conv3d_desc* d, int for_dgrad) {
    if (!desc_ok(d)) return 0;
    // layout: [ mma.sync operand pack | (256-byte aligned) tcgen05 operand pack ]
    if (!for_dgrad) {
        size_t bytes = rup256(mma_fwd_elems(d) * 2);
        if (tc_fwd_ok(d)) bytes += vg_tc_pack_elems(d->Cout, tc_fwd_kdim(d), tc_fwd_taps(d)) * 2;
        return bytes;
    }
    size_t bytes = rup256(mma_dgrad_elems(d) * 2);
    if (tc_dgrad_ok(d))
        for (int a = 0; a < d->stride; a++)
            for (int b = 0; b < d->stride; b++)
                for (int c = 0; c < d->stride; c++) bytes += rup256(tc_dgrad_class_elems(d, a, b, c) * 2);
    if (tc_dgrad_ok(d) && vg_tc_s2dgrad_ok(d->K, d->stride, d->Cin, d->Cout)) bytes += rup256(vg_tc_s2dgrad_elems(d->Cin, d->Cout) * 2);
    return bytes;
}

// Expands one layer into its packing jobs (host only, no CUDA call).  Returns the number of jobs written, or a negative error.
int vg_conv3d_pack_jobs(const vg_conv3d_desc* d, const float* w, void* w_fwd, void* w_dgrad, void* jobs_out, int max_jobs) {
    VG_REQUIRE(desc_ok(d) && w && jobs_out && max_jobs > 0);
    vg_pack_job* jobs = (vg_pack_job*)jobs_out;
    int n = 0;
    const int T = d->K * d->K * d->K, s_ = d->stride;
    auto push = [&](const vg_pack_job& j) -> bool {
        if (j.total <= 0) return true;
        if (n >= max_jobs) return false;
        jobs[n++] = j;
        return true;
    };
    if (w_fwd && d->Cin != 1) {
        vg_pack_job j{};
        j.w = w; j.out = (bf16*)w_fwd; j.kind = 0; j.K = d->K; j.stride = s_; j.Cin = d->Cin; j.Cout = d->Cout;
        j.Np = rup(d->Cout, NPAD); j.T = T; j.total = (long long)T * j.Np * d->Cin;
        if (!push(j)) return VG_ERR_WORKSPACE;
    }
    if (w_dgrad && d->Cout != 1) {
        const int NpI = rup(d->Cin, NPAD);
        bf16* dst = (bf16*)w_dgrad;
        for (int a = 0; a < s_; a++)
            for (int b = 0; b < s_; b++)
                for (int c = 0; c < s_; c++) {
                    vg_pack_job j{};
                    j.w = w; j.out = dst; j.kind = 1; j.K = d->K; j.stride = s_; j.Cin = d->Cin; j.Cout = d->Cout;
                    j.ad = a; j.ah = b; j.aw = c; j.td = class_taps(d->K, s_, a); j.th = class_taps(d->K, s_, b); j.tw = class_taps(d->K, s_, c);
                    j.Np = NpI; j.total = (long long)j.td * j.th * j.tw * NpI * d->Cout;
                    if (!push(j)) return VG_ERR_WORKSPACE;
                    dst += j.total;
                }
    }
    if (w_fwd && tc_fwd_ok(d)) {
        bf16* dst = (bf16*)((char*)w_fwd + rup256(mma_fwd_elems(d) * 2));
        vg_pack_job j;
        const bool ok = tc_fwd_s2(d) ? vg_tc_pack_job(w, dst, d->K, 2, d->Cin, d->Cout, 2, 0, 0, 0, 2, 2, 2, &j)
                                     : vg_tc_pack_job(w, dst, d->K, 1, d->Cin, d->Cout, 0, 0, 0, 0, d->K, d->K, d->K, &j);
        if (!ok) return VG_ERR_CUDA;
        if (!push(j)) return VG_ERR_WORKSPACE;
    }
    if (w_dgrad && tc_dgrad_ok(d)) {
        char* dst = (char*)w_dgrad + rup256(mma_dgrad_elems(d) * 2);
        for (int a = 0; a < s_; a++)
            for (int b = 0; b < s_; b++)
                for (int c = 0; c < s_; c++) {
                    size_t el = tc_dgrad_class_elems(d, a, b, c);
                    if (!el) continue;
                    vg_pack_job j;
                    if (!vg_tc_pack_job(w, (bf16*)dst, d->K, s_, d->Cin, d->Cout, 1, a, b, c, class_taps(d->K, s_, a), class_taps(d->K, s_, b),
                                        class_taps(d->K, s_, c), &j))
                        return VG_ERR_CUDA;
                    if (!push(j)) return VG_ERR_WORKSPACE;
                    dst += rup256(el * 2);
                }
        if (vg_tc_s2dgrad_ok(d->K, s_, d->Cin, d->Cout)) {   // fused classes, after the class packs
            vg_pack_job j;
            if (!vg_tc_pack_job(w, (bf16*)dst, d->K, s_, d->Cin, d->Cout, 3, 0, 0, 0, 2, 2, 2, &j)) return VG_ERR_CUDA;
            if (!push(j)) return VG_ERR_WORKSPACE;
        }
    }
    return n;
}

size_t vg_pack_job_bytes(void) { return sizeof(vg_pack_job); }

long long vg_pack_job_total(const void* jobs, int index) { return ((const vg_pack_job*)jobs)[index].total; }

int vg_pack_run(const void* jobs_dev, const long long* prefix_dev, int njobs, long long total, void* stream) {
    VG_REQUIRE(jobs_dev && prefix_dev && njobs > 0 && total > 0);
    pack_multi_kernel<<<vg_grid_for((total + PACK_SPAN - 1) / PACK_SPAN, 1, 8), 256, 0, (cudaStream_t)stream>>>(
        (const vg_pack_job*)jobs_dev, prefix_dev, njobs, total); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_conv3d_pack_weights(const vg_conv3d_desc* d, const float* w, void* w_fwd, void* w_dgrad, void* stream) {
    vg_pack_job jobs[24];
    const int n = vg_conv3d_pack_jobs(d, w, w_fwd, w_dgrad, jobs, 24);
    if (n < 0) return n;
    for (int i = 0; i < n; i++) {
        pack_job_kernel<<<vg_grid_for(jobs[i].total, 256, 4), 256, 0, (cudaStream_t)stream>>>(jobs[i]); VG_LAUNCHED(1);
    }
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_conv3d_fwd(const vg_conv3d_desc* d, const void* x, const void* w_fwd, const float* bias, void* y, void* stream) {
    VG_REQUIRE(desc_ok(d) && x && w_fwd && y);
    cudaStream_t st = (cudaStream_t)stream;
    const int OD = odim(d->ID, d->K, d->stride), OH = odim(d->IH, d->K, d->stride), OW = odim(d->IW, d->K, d->stride);
    if (d->Cin == 1) {
        VG_REQUIRE(d->Cout % 16 == 0 && d->act == VG_ACT_NONE);
        int T = d->K * d->K * d->K;
        size_t smem = ((size_t)T * d->Cout + d->Cout) * sizeof(float);
        size_t total = (size_t)d->N * OD * ((OH * OW + 127) / 128) * (d->Cout / 16);   // warp items
        int grid = vg_grid_for(total, 4, 8);
        // k4 layers (PatchGAN d0, 1 -> 64): the software-pipelined tap loop, 0.598 -> 0.553 ms at 8x130^3; k3 (stem, 1 -> 16) measured
        // 0.549 -> 0.560 ms and keeps the plain loop.  VG_CIN1_PF=0: plain loop everywhere (A/B testing)
        const char* epf = getenv("VG_CIN1_PF");
        const bool pf = !(epf && epf[0] == '0') && d->K == 4;
#define VG_CIN1_FWD(KK, SS)                                                                                                       \
    do {                                                                                                                          \
        if (pf && KK == 4)                                                                                                        \
            cin1_fwd_kernel<KK, SS, (KK == 4)><<<grid, 128, smem, st>>>((const float*)x, (const float*)w_fwd, bias, (bf16*)y, d->N,         \
                                                                                d->ID, d->IH, d->IW, OD, OH, OW, d->Cout);            \
        else                                                                                                                      \
            cin1_fwd_kernel<KK, SS, false><<<grid, 128, smem, st>>>((const float*)x, (const float*)w_fwd, bias, (bf16*)y, d->N, d->ID, d->IH, \
                                                                  d->IW, OD, OH, OW, d->Cout);                                        \
    } while (0)
        if (d->K == 1 && d->stride == 1) VG_CIN1_FWD(1, 1);
        else if (d->K == 3 && d->stride == 1) VG_CIN1_FWD(3, 1);
        else if (d->K == 4 && d->stride == 2) VG_CIN1_FWD(4, 2);
        else if (d->K == 3 && d->stride == 2) VG_CIN1_FWD(3, 2);
        else if (d->K == 4 && d->stride == 1) VG_CIN1_FWD(4, 1);
        else if (d->K == 7 && d->stride == 1) VG_CIN1_FWD(7, 1);
        else return VG_ERR_UNSUPPORTED;
#undef VG_CIN1_FWD
        VG_LAUNCHED(1);
        VG_CHECK_LAUNCH();
        return VG_OK;
    }
    if (d->K == 7) {   // Cout == 1 (Cin == 1 returned above): direct kernel on the bf16 forward pack [T][Np][Cin], column 0
        int rc = vg_k7_many_to_one((const bf16*)x, (const bf16*)w_fwd, rup(d->Cout, NPAD) * d->Cin, bias, (float*)y, d->N, d->ID, d->IH, d->IW,
                                   d->Cin, OD, OH, OW, d->K, +1, d->act, st);
        if (rc != VG_OK) return rc;
        VG_CHECK_LAUNCH();
        return VG_OK;
    }
    if (small_enabled() && d->K == 1 && d->stride == 1) {
        int rc = VG_ERR_UNSUPPORTED;
        const long long nvox = (long long)d->N * OD * OH * OW;
        if (d->Cout == 1) rc = vg_small_cout1_k1_fwd((const bf16*)x, (const bf16*)w_fwd, bias, (float*)y, (size_t)nvox, d->Cin, d->act, st);
        else if (d->act == VG_ACT_NONE && d->y_dtype == VG_BF16)
            rc = vg_small_k1_fwd((const bf16*)x, (const bf16*)w_fwd, bias, (bf16*)y, nvox, d->Cin, d->Cout, st);
        if (rc == VG_OK) { VG_CHECK_LAUNCH(); return VG_OK; }
        if (rc != VG_ERR_UNSUPPORTED) return rc;
    }
    if (tc_enabled() && tc_fwd_ok(d) && d->y_dtype == VG_BF16) {
        const bf16* wt = (const bf16*)((const char*)w_fwd + rup256(mma_fwd_elems(d) * 2));
        const int kt = tc_fwd_s2(d) ? 2 : d->K;
        int rc = vg_tc_launch((const bf16*)x, d->N, d->ID, d->IH, d->IW, d->Cin, wt, y, bias, OD, OH, OW, d->Cout, OD, OH, OW, kt, kt, kt,
                              +1, 1, 0, 0, 0, d->act, st, 0, d->stride);
        if (rc == VG_OK) { VG_CHECK_LAUNCH(); return VG_OK; }
        if (rc != VG_ERR_UNSUPPORTED) return rc;
    }
    GConv p{};
    p.x = (const bf16*)x; p.w = (const bf16*)w_fwd; p.y = y; p.bias = bias;
    p.N = d->N; p.XD = d->ID; p.XH = d->IH; p.XW = d->IW; p.Cx = d->Cin;
    p.YD = OD; p.YH = OH; p.YW = OW; p.Cy = d->Cout;
    p.GD = OD; p.GH = OH; p.GW = OW;
    p.TD = p.TH = p.TW = d->K;
    p.so = d->stride; p.st = 1;
    p.oso = 1; p.ood = p.ooh = p.oow = 0;
    p.out_f32 = d->y_dtype == VG_F32; p.act = d->act; p.Np = rup(d->Cout, NPAD);
    int rc = launch_gconv(p, st);
    if (rc != VG_OK) return rc;
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_conv3d_dgrad(const vg_conv3d_desc* d, const void* dy, const void* w_dgrad, void* dx, void* stream) {
    VG_REQUIRE(desc_ok(d) && dy && w_dgrad && dx);
    cudaStream_t st = (cudaStream_t)stream;
    const int OD = odim(d->ID, d->K, d->stride), OH = odim(d->IH, d->K, d->stride), OW = odim(d->IW, d->K, d->stride);
    if (d->Cout == 1) {
        VG_REQUIRE(d->stride == 1 && d->Cin % 8 == 0);
        if (small_enabled() && d->K == 1) {
            int rc = vg_small_cout1_k1_dgrad((const float*)dy, (const float*)w_dgrad, (bf16*)dx, (size_t)d->N * OD * OH * OW, d->Cin, st);
            if (rc == VG_OK) { VG_CHECK_LAUNCH(); return VG_OK; }
            if (rc != VG_ERR_UNSUPPORTED) return rc;
        }
        size_t total = (size_t)d->N * d->ID * d->IH * d->IW * (d->Cin / 8);
        cout1_dgrad_kernel<<<vg_grid_for(total, 256, 16), 256, 0, st>>>((const float*)dy, (const float*)w_dgrad, (bf16*)dx, d->N,
                                                                       d->ID, d->IH, d->IW, OD, OH, OW, d->Cin, d->K); VG_LAUNCHED(1);
        VG_CHECK_LAUNCH();
        return VG_OK;
    }
    const int s = d->stride;
    if (d->K == 7) {   // Cin == 1 (Cout == 1 returned above): dx = sum over taps and channels of dy, on the bf16 dgrad pack [T][Np][Cout], row 0
        int rc = vg_k7_many_to_one((const bf16*)dy, (const bf16*)w_dgrad, rup(d->Cin, NPAD) * d->Cout, nullptr, (float*)dx, d->N, OD, OH, OW,
                                   d->Cout, d->ID, d->IH, d->IW, d->K, -1, VG_ACT_NONE, st);
        if (rc != VG_OK) return rc;
        VG_CHECK_LAUNCH();
        return VG_OK;
    }
    if (small_enabled() && d->Cin == 1 && s == 2 && d->x_dtype == VG_F32) {
        int rc = vg_small_cin1_dgrad_s2((const bf16*)dy, (const bf16*)w_dgrad, (float*)dx, d->N, d->ID, d->IH, d->IW, OD, OH, OW, d->Cout, d->K, st);
        if (rc == VG_OK) { VG_CHECK_LAUNCH(); return VG_OK; }
        if (rc != VG_ERR_UNSUPPORTED) return rc;
    }
    if (small_enabled() && d->K == 1 && s == 1 && d->x_dtype == VG_BF16 && d->y_dtype == VG_BF16 && !d->dx_lo && !d->dx_hi) {
        // 1x1x1 shortcut: dx[v][ci] = sum_co dy[v][co] w[ci][co] is the streaming forward kernel with the channel roles swapped; the
        // mma.sync dgrad pack [Np(ci)][Cout] is exactly its operand layout (one pass over dy and dx instead of K = 1-tap MMAs)
        int rc = vg_small_k1_fwd((const bf16*)dy, (const bf16*)w_dgrad, nullptr, (bf16*)dx, (long long)d->N * OD * OH * OW, d->Cout, d->Cin, st);
        if (rc == VG_OK) { VG_CHECK_LAUNCH(); return VG_OK; }
        if (rc != VG_ERR_UNSUPPORTED) return rc;
    }
    if (d->K < s) {  // k1 s2: odd-parity classes receive nothing
        size_t bytes = (size_t)d->N * d->ID * d->IH * d->IW * d->Cin * (d->x_dtype == VG_F32 ? 4 : 2);
        if (cudaMemsetAsync(dx, 0, bytes, st) != cudaSuccess) return VG_ERR_CUDA;
    }
    const int NpI = rup(d->Cin, NPAD);
    size_t woff = 0;
    const bool use_tc = tc_enabled() && tc_dgrad_ok(d) && d->x_dtype == VG_BF16;
    const char* wtc = (const char*)w_dgrad + rup256(mma_dgrad_elems(d) * 2);
    if (use_tc && vg_tc_s2dgrad_ok(d->K, s, d->Cin, d->Cout)) {
        const char* wf = wtc;
        for (int a = 0; a < s; a++)
            for (int b = 0; b < s; b++)
                for (int c = 0; c < s; c++) wf += rup256(tc_dgrad_class_elems(d, a, b, c) * 2);
        // one launch over the class grid p' (p = 2p' + a): 2x2x2 taps (zero weights where a + 2t' >= K), columns = (class, ci)
        int rc = vg_tc_launch((const bf16*)dy, d->N, OD, OH, OW, d->Cout, (const bf16*)wf, dx, nullptr, d->ID, d->IH, d->IW, d->Cin,
                              (d->ID + 1) / 2, (d->IH + 1) / 2, (d->IW + 1) / 2, 2, 2, 2, -1, 2, 0, 0, 0, VG_ACT_NONE, st, 0, 1, d->Cin);
        if (rc == VG_OK) { VG_CHECK_LAUNCH(); return VG_OK; }
        if (rc != VG_ERR_UNSUPPORTED) return rc;
    }
    for (int ad = 0; ad < s; ad++)
        for (int ah = 0; ah < s; ah++)
            for (int aw = 0; aw < s; aw++) {
                int td = class_taps(d->K, s, ad), th = class_taps(d->K, s, ah), tw = class_taps(d->K, s, aw);
                size_t cnt = (size_t)td * th * tw * NpI * d->Cout;
                if (cnt == 0) continue;
                if (use_tc) {
                    const bf16* wt = (const bf16*)wtc;
                    wtc += rup256(tc_dgrad_class_elems(d, ad, ah, aw) * 2);
                    // stride 1: only the cropped region [dx_lo, I - dx_hi) is computed (zero-padded inputs)
                    const int crop = s == 1 ? d->dx_lo + d->dx_hi : 0, goff = s == 1 ? d->dx_lo : 0;
                    int rc = vg_tc_launch((const bf16*)dy, d->N, OD, OH, OW, d->Cout, wt, dx, nullptr, d->ID, d->IH, d->IW, d->Cin,
                                          (d->ID - crop - ad + s - 1) / s, (d->IH - crop - ah + s - 1) / s, (d->IW - crop - aw + s - 1) / s, td,
                                          th, tw, -1, s, ad, ah, aw, VG_ACT_NONE, st, goff);
                    if (rc == VG_OK) { woff += cnt; continue; }
                    if (rc != VG_ERR_UNSUPPORTED) return rc;
                }
                GConv p{};
                p.x = (const bf16*)dy; p.w = (const bf16*)w_dgrad + woff; p.y = dx; p.bias = nullptr;
                p.N = d->N; p.XD = OD; p.XH = OH; p.XW = OW; p.Cx = d->Cout;
                p.YD = d->ID; p.YH = d->IH; p.YW = d->IW; p.Cy = d->Cin;
                p.GD = (d->ID - ad + s - 1) / s; p.GH = (d->IH - ah + s - 1) / s; p.GW = (d->IW - aw + s - 1) / s;
                p.TD = td; p.TH = th; p.TW = tw;
                p.so = 1; p.st = -1;
                p.oso = s; p.ood = ad; p.ooh = ah; p.oow = aw;
                p.out_f32 = d->x_dtype == VG_F32; p.act = VG_ACT_NONE; p.Np = NpI;
                int rc = launch_gconv(p, st);
                if (rc != VG_OK) return rc;
                woff += cnt;
            }
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_conv3d_wgrad(const vg_conv3d_desc* d, const void* x, const void* dy, float* dw, float* dbias, void* stream) {
    VG_REQUIRE(desc_ok(d) && x && dy && dw);
    cudaStream_t st = (cudaStream_t)stream;
    const int OD = odim(d->ID, d->K, d->stride), OH = odim(d->IH, d->K, d->stride), OW = odim(d->IW, d->K, d->stride);
    const size_t rows = (size_t)d->N * OD * OH * OW;
    if (d->K == 7) {
        if (dbias) {
            if (d->Cout == 1) {
                channel_sum_kernel<float><<<vg_grid_for(rows, 256, 2), 256, 32 * sizeof(float), st>>>((const float*)dy, rows, 1, dbias); VG_LAUNCHED(1);
            } else {
                channel_sum_kernel<bf16><<<vg_grid_for(rows, 32 * 4, 4), 256, (size_t)(256 / (d->Cout / 8)) * d->Cout * sizeof(float), st>>>(
                    (const bf16*)dy, rows, d->Cout, dbias); VG_LAUNCHED(1);
            }
        }
        int rc = d->Cin == 1 ? vg_k7_wgrad_one((const bf16*)dy, (const float*)x, dw, d->N, OD, OH, OW, d->Cout, d->ID, d->IH, d->IW, d->K, +1, 0, st)
                             : vg_k7_wgrad_one((const bf16*)x, (const float*)dy, dw, d->N, d->ID, d->IH, d->IW, d->Cin, OD, OH, OW, d->K, -1, 0, st);
        if (rc != VG_OK) return rc;
        VG_CHECK_LAUNCH();
        return VG_OK;
    }
    if (d->Cin == 1) {
        VG_REQUIRE(d->Cout % 16 == 0);
        if (small_enabled()) {
            int bias_done = 0;
            int rc = vg_small_cin1_wgrad((const float*)x, (const bf16*)dy, dw, dbias, d->N, d->ID, d->IH, d->IW, OD, OH, OW, d->Cout, d->K,
                                         d->stride, &bias_done, st);
            if (rc == VG_OK) {
                if (dbias && !bias_done) {
                    channel_sum_kernel<bf16><<<vg_grid_for(rows, 32 * 4, 4), 256, (size_t)(256 / (d->Cout / 8)) * d->Cout * sizeof(float), st>>>(
                        (const bf16*)dy, rows, d->Cout, dbias); VG_LAUNCHED(1);
                }
                VG_CHECK_LAUNCH();
                return VG_OK;
            }
            if (rc != VG_ERR_UNSUPPORTED) return rc;
        }
        const int ntask = d->K * d->K * (d->Cout / 16);
        int wpb = ntask;                       // warps per block: all tasks when they fit, else the largest divisor <= 16
        while (wpb > 16) wpb = (wpb % 2 == 0) ? wpb / 2 : 1;
        VG_REQUIRE(wpb >= 1 && ntask % wpb == 0);
        const int ntg = ntask / wpb;
        const int items = d->N * OD * OH;
        int gx = (148 * 16 / wpb + ntg - 1) / ntg;   // ~16 resident warps per SM
        if (gx > items) gx = items;
        if (gx < 1) gx = 1;
        const dim3 grid(gx, ntg);
#define VG_CIN1_WG(KK, SS)                                                                                                     \
    cin1_wgrad_kernel<KK, SS><<<grid, wpb * 32, 0, st>>>((const float*)x, (const bf16*)dy, dw, dbias, d->N, d->ID, d->IH, d->IW, OD, \
                                                         OH, OW, d->Cout)
        if (d->K == 1 && d->stride == 1) VG_CIN1_WG(1, 1);
        else if (d->K == 3 && d->stride == 1) VG_CIN1_WG(3, 1);
        else if (d->K == 4 && d->stride == 2) VG_CIN1_WG(4, 2);
        else if (d->K == 3 && d->stride == 2) VG_CIN1_WG(3, 2);
        else if (d->K == 4 && d->stride == 1) VG_CIN1_WG(4, 1);
        else return VG_ERR_UNSUPPORTED;
#undef VG_CIN1_WG
        VG_LAUNCHED(1);
        VG_CHECK_LAUNCH();
        return VG_OK;
    }
    if (dbias) {
        if (d->Cout == 1) {
            channel_sum_kernel<float><<<vg_grid_for(rows, 256, 2), 256, 32 * sizeof(float), st>>>((const float*)dy, rows, 1, dbias); VG_LAUNCHED(1);
        }
        else {
            channel_sum_kernel<bf16><<<vg_grid_for(rows, 32 * 4, 4), 256, (size_t)(256 / (d->Cout / 8)) * d->Cout * sizeof(float), st>>>(
                (const bf16*)dy, rows, d->Cout, dbias); VG_LAUNCHED(1);
        }
    }
    if (d->Cout == 1) {
        VG_REQUIRE(d->stride == 1 && d->Cin % 8 == 0 && d->Cin <= 2048);
        if (small_enabled()) {
            int rc = d->K == 1 ? vg_small_cout1_k1_wgrad((const bf16*)x, (const float*)dy, dw, rows, d->Cin, st)
                               : vg_small_cout1_wgrad((const bf16*)x, (const float*)dy, dw, d->N, d->ID, d->IH, d->IW, OD, OH, OW, d->Cin, d->K, st);
            if (rc == VG_OK) { VG_CHECK_LAUNCH(); return VG_OK; }
            if (rc != VG_ERR_UNSUPPORTED) return rc;
        }
        int T = d->K * d->K * d->K;
        int nbx = (148 * 4 + T - 1) / T;
        int per_block = (int)((rows + nbx - 1) / nbx);
        if (per_block < 32) per_block = 32;
        size_t smem = (size_t)(256 / (d->Cin / 8)) * d->Cin * sizeof(float);
        cout1_wgrad_kernel<<<dim3(vg_cdiv(rows, per_block), T), 256, smem, st>>>((const bf16*)x, (const float*)dy, dw, d->N, d->ID,
                                                                                d->IH, d->IW, OD, OH, OW, d->Cin, d->K, per_block); VG_LAUNCHED(1);
        VG_CHECK_LAUNCH();
        return VG_OK;
    }
    if (small_enabled() && d->K == 1) {
        int rc = vg_small_k1_wgrad((const bf16*)x, (const bf16*)dy, dw, d->N, d->ID, d->IH, d->IW, d->Cin, OD, OH, OW, d->Cout, d->stride, st);
        if (rc == VG_OK) { VG_CHECK_LAUNCH(); return VG_OK; }
        if (rc != VG_ERR_UNSUPPORTED) return rc;
    }
    if (tc_wgrad_enabled()) {
        int rc = vg_wg_tc_launch((const bf16*)x, (const bf16*)dy, dw, d->N, d->ID, d->IH, d->IW, d->Cin, OD, OH, OW, d->Cout, d->K, d->stride, st);
        if (rc == VG_OK) { VG_CHECK_LAUNCH(); return VG_OK; }
        if (rc != VG_ERR_UNSUPPORTED) return rc;
    }
    WGrad p{};
    p.x = (const bf16*)x; p.dy = (const bf16*)dy; p.dw = dw;
    p.N = d->N; p.XD = d->ID; p.XH = d->IH; p.XW = d->IW; p.Cx = d->Cin;
    p.OD = OD; p.OH = OH; p.OW = OW; p.Cy = d->Cout;
    p.K = d->K; p.so = d->stride;
    p.gtd = (OD + BD - 1) / BD; p.gth = (OH + BH - 1) / BH; p.gtw = (OW + BW - 1) / BW;
    p.nbricks = p.gtd * p.gth * p.gtw * d->N;
    int rc;
    if (d->K == 1) rc = (d->Cout % 32 == 0) ? launch_wgrad_t<1, 4>(p, st) : launch_wgrad_t<1, 2>(p, st);
    else if (d->K == 3) rc = (d->Cout % 32 == 0) ? launch_wgrad_t<4, 4>(p, st) : launch_wgrad_t<4, 2>(p, st);
    else rc = launch_wgrad_t<8, 2>(p, st);
    if (rc != VG_OK) return rc;
    VG_CHECK_LAUNCH();
    return VG_OK;
}

}  // extern "C"
