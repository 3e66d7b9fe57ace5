// Shared device helpers for the VAN-GAN B200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vangan_b200.h"

typedef __nv_bfloat16 bf16;

#define VG_CHECK_LAUNCH()                                   \
    do {                                                    \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return VG_ERR_CUDA;         \
    } while (0)

// launch accounting (read by bench.py through vg_launch_count): every kernel launch goes through VG_LAUNCHED(n)
extern unsigned long long g_vg_launches;
#define VG_LAUNCHED(n) (g_vg_launches += (unsigned long long)(n))

#define VG_REQUIRE(cond)                      \
    do {                                      \
        if (!(cond)) return VG_ERR_INVALID;   \
    } while (0)

static inline int vg_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// grid sizing: a multiple of the SM count (148 on B200), capped by the work available
// cudaFuncSetAttribute applies to the CURRENT device: a per-process flag would leave the second GPU of a multi-device process without
// its shared-memory opt-in.  One bit per device ordinal.
struct VgPerDevice {
    unsigned long long mask = 0;
    static unsigned long long bit() {
        int d = 0;
        cudaGetDevice(&d);
        return 1ull << (d & 63);
    }
    bool done() const { return (mask & bit()) != 0; }
    void mark() { mask |= bit(); }
};

static inline int vg_grid_for(long long work_items, int per_block, int waves = 8) {
    long long blocks = (work_items + per_block - 1) / per_block;
    long long cap = 148LL * waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum of a double; result valid in thread 0.  `sh` must hold >= 32 doubles.
__device__ __forceinline__ double block_sum_d(double v, double* sh) {
    v = warp_sum_d(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        v = lane < nw ? sh[lane] : 0.0;
        v = warp_sum_d(v);
    }
    return v;
}

// reflect index for pad-1 REFLECT padding: -1 -> 1, n -> n-2
__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// 8 x bf16 <-> 8 x float through one 128-bit access.  The packet is a plain uint4: nvcc only emits LDG.128 / STG.128 for
// the built-in vector types (a struct of four __nv_bfloat162 is copied member-wise, i.e. as four 32-bit accesses).
typedef uint4 bf16x8;
__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
    f[0] = __uint_as_float(p.x << 16); f[1] = __uint_as_float(p.x & 0xffff0000u);
    f[2] = __uint_as_float(p.y << 16); f[3] = __uint_as_float(p.y & 0xffff0000u);
    f[4] = __uint_as_float(p.z << 16); f[5] = __uint_as_float(p.z & 0xffff0000u);
    f[6] = __uint_as_float(p.w << 16); f[7] = __uint_as_float(p.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h2);
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
    return make_uint4(pack2_bf16(f[0], f[1]), pack2_bf16(f[2], f[3]), pack2_bf16(f[4], f[5]), pack2_bf16(f[6], f[7]));
}

// generic 8-wide load/store on either float or bf16 storage
template <typename T>
__device__ __forceinline__ void load8(const T* p, float* f);
template <>
__device__ __forceinline__ void load8<bf16>(const bf16* p, float* f) {
    const bf16x8 v = *reinterpret_cast<const uint4*>(p);   // plain load: some callers read-modify-write the same buffer
    unpack8(v, f);
}
template <>
__device__ __forceinline__ void load8<float>(const float* p, float* f) {
    float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float* f);
template <>
__device__ __forceinline__ void store8<bf16>(bf16* p, const float* f) {
    *reinterpret_cast<bf16x8*>(p) = pack8(f);
}
template <>
__device__ __forceinline__ void store8<float>(float* p, const float* f) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}

// Philox-4x32-10 counter RNG (one call -> 4 uniform u32); key = (seed_lo, seed_hi)
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
// two standard normals from two u32 (Box-Muller)
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    float u1 = (a + 1.0f) * 2.3283064365386963e-10f;  // (0,1]
    float u2 = b * 2.3283064365386963e-10f;
    float r = sqrtf(-2.0f * __logf(u1));
    float s, c;
    __sincosf(6.283185307179586f * u2, &s, &c);
    return make_float2(r * c, r * s);
}

// tensor-core (tcgen05) convolution path, conv_tc.cu
int vg_tc_ncta(int ncols, int T);
size_t vg_tc_pack_elems(int ncols, int K_total, int T);
int vg_tc_launch(const bf16* x, int Nb, int XD, int XH, int XW, int Cx, const bf16* wpack, void* y, const float* bias, int YD, int YH,
                 int YW, int Cy, int GD, int GH, int GW, int TD, int TH, int TW, int st, int oso, int ood, int ooh, int oow, int act,
                 cudaStream_t stream, int goff = 0, int ss = 1, int cls_cin = 0);
bool vg_tc_s2dgrad_ok(int K, int stride, int Cin, int Cout);
size_t vg_tc_s2dgrad_elems(int Cin, int Cout);
int vg_tc_pack(const float* w, bf16* out, int K, int stride, int Cin, int Cout, int dgrad, int ad, int ah, int aw, int td, int th,
               int tw, cudaStream_t st);
// tensor-core (tcgen05) weight-gradient path, wgrad_tc.cu
int vg_wg_tc_launch(const bf16* x, const bf16* dy, float* dw, int Nb, int XD, int XH, int XW, int Cx, int OD, int OH, int OW, int Cy,
                    int K, int stride, cudaStream_t stream);
// HBM-bound shapes (1x1x1 kernels, single-channel outputs, single-channel-input stride-2 dgrad), conv_small.cu
int vg_small_k1_fwd(const bf16* x, const bf16* wp, const float* bias, bf16* y, long long nvox, int Cin, int Cout, cudaStream_t st);
int vg_small_k1_wgrad(const bf16* x, const bf16* dy, float* dw, int N, int ID, int IH, int IW, int Cin, int OD, int OH, int OW, int Cout,
                      int stride, cudaStream_t st);
int vg_small_cout1_k1_fwd(const bf16* x, const bf16* wp, const float* bias, float* y, size_t nvox, int Cin, int act, cudaStream_t st);
int vg_small_cout1_k1_dgrad(const float* dy, const float* w, bf16* dx, size_t nvox, int Cin, cudaStream_t st);
int vg_small_cout1_k1_wgrad(const bf16* x, const float* dy, float* dw, size_t nvox, int Cin, cudaStream_t st);
int vg_small_cout1_wgrad(const bf16* x, const float* dy, float* dw, int N, int ID, int IH, int IW, int OD, int OH, int OW, int Cin, int K,
                         cudaStream_t st);
int vg_small_cin1_dgrad_s2(const bf16* dy, const bf16* wd, float* dx, int N, int ID, int IH, int IW, int OD, int OH, int OW, int Cout, int K,
                           cudaStream_t st);
int vg_small_cin1_wgrad(const float* x, const bf16* dy, float* dw, float* dbias, int N, int ID, int IH, int IW, int OD, int OH, int OW,
                        int Cout, int K, int stride, int* bias_done, cudaStream_t st);
