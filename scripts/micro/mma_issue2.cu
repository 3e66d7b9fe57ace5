// Micro-benchmark 2: what one tcgen05.mma (cta_group::1, kind::f16, M=128, K=16, SWIZZLE_NONE operands in shared memory) costs
// when issued back to back by one thread, as a function of
//   N            : instruction width
//   amode        : 0 = every instruction reads the SAME A tile, 1 = A cycles through `na` distinct tiles (d-march: one per source slice),
//                  2 = as 1 plus a 16-byte tap shift that changes every `na` instructions (the (th, tw) loop of the convolution)
//   sbo          : byte distance between 8-row groups of A (128 = dense core matrices, 160 = halo rows of 10 voxels)
//   dmode        : 0 = four disjoint accumulator windows round-robin, 1 = d-march sliding windows (tile m at column m*NC, window of
//                  up to 3 tiles) in the permuted slice order the convolution kernel uses
// One CTA per SM (148), all SMs busy, nothing else running in the CTA.  Prints cycles per instruction and the two models
// (math floor 128*N/256; operand fetch (4096 + 32*N)/128).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/mma_issue2.bin scripts/micro/mma_issue2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../van-gan_b200/csrc/tc_ptx.cuh"
using namespace tcp;

struct Cfg {
    int N, amode, na, sbo, dmode, NC, iters;
};

__global__ void __launch_bounds__(128, 1) k(Cfg c, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ uint4 tab[64];
    for (int i = threadIdx.x * 16; i < 200 * 1024; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(s_addr(&bar), 1); mbar_init_fence(); }
    if (warp == 0) tmem_alloc(s_addr(&slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    const uint32_t sb = s_addr(smem);
    // A region: two K-half planes of 48 KB each (LBO = 48 KB): na tiles at a pitch of 18 row groups (one d-slice of an 18x10 halo)
    const uint32_t a_lbo = 48 * 1024;
    const uint32_t tile_pitch = 18 * (uint32_t)c.sbo;
    if (threadIdx.x < 64) {
        // per-issue-position table: A offset (16-byte units), TMEM column, N of the instruction
        const int i = threadIdx.x;
        uint32_t aoff = 0, col = 0, n = (uint32_t)c.N;
        if (c.amode >= 1) aoff = (uint32_t)(i % c.na) * (tile_pitch >> 4);
        if (c.dmode == 0) col = (uint32_t)((i & 3) * c.N);
        else {
            // d-march, BD = 8, TD = 3: ED = 10 slices issued in the order s_i = (3 i) mod 10; slice s feeds tiles max(0,s-2)..min(7,s)
            const int ED = 10, s = (3 * (i % ED)) % ED;
            const int m_lo = s - 2 > 0 ? s - 2 : 0, m_hi = s < 7 ? s : 7;
            col = (uint32_t)(m_lo * c.NC);
            n = (uint32_t)((m_hi - m_lo + 1) * c.NC);
            aoff = (uint32_t)s * (tile_pitch >> 4);
        }
        tab[i] = make_uint4(aoff, col, n, 0);
    }
    __syncthreads();
    if (warp == 0) {
        const uint32_t leader = elect_one();
        const uint64_t adesc0 = make_desc(sb, a_lbo, (uint32_t)c.sbo);
        const uint64_t bdesc = make_desc(sb + 100 * 1024, (uint32_t)(c.dmode ? 3 * c.NC : c.N) * 16, 128);
        const int period = c.dmode ? 10 : (c.amode ? c.na : 4);
        long long t0 = clock64();
        if (leader) {
            int shift = 0;
            for (int i = 0; i < c.iters; i += period) {
                for (int j = 0; j < period; j++) {
                    const uint4 e = tab[j];
                    const uint32_t idesc = make_idesc_bf16(128, (int)e.z, 0, 0);
                    tc_mma(tm + e.y, adesc0 + (uint64_t)(e.x + (uint32_t)shift), bdesc, idesc, 1u);
                }
                if (c.amode == 2) shift = (shift + 1) % 3;
            }
            tc_commit(s_addr(&bar));
        }
        __syncwarp();
        mbar_wait(s_addr(&bar), 0);
        long long t1 = clock64();
        if (leader && blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 12000;
    Cfg cfgs[] = {
        // N, amode, na, sbo, dmode, NC
        {16, 0, 1, 128, 0, 0, iters},  {48, 0, 1, 128, 0, 0, iters},  {96, 0, 1, 128, 0, 0, iters},  {128, 0, 1, 128, 0, 0, iters},
        {48, 0, 1, 160, 0, 0, iters},  {128, 0, 1, 160, 0, 0, iters},
        {16, 1, 10, 128, 0, 0, iters}, {48, 1, 10, 128, 0, 0, iters}, {96, 1, 10, 128, 0, 0, iters}, {128, 1, 10, 128, 0, 0, iters},
        {16, 1, 10, 160, 0, 0, iters}, {48, 1, 10, 160, 0, 0, iters}, {96, 1, 10, 160, 0, 0, iters}, {128, 1, 10, 160, 0, 0, iters},
        {48, 2, 10, 160, 0, 0, iters}, {128, 2, 10, 160, 0, 0, iters}, {48, 2, 10, 128, 0, 0, iters},
        {48, 1, 2, 160, 0, 0, iters},  {48, 1, 4, 160, 0, 0, iters},
        {0, 1, 10, 160, 1, 16, iters}, {0, 1, 10, 160, 1, 32, iters}, {0, 1, 10, 160, 1, 48, iters}, {0, 1, 10, 128, 1, 16, iters},
    };
    for (const Cfg& c : cfgs) {
        for (int rep = 0; rep < 2; rep++) {
            k<<<148, 128, 200 * 1024>>>(c, d);
            cudaDeviceSynchronize();
        }
        long long cyc;
        cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        const int per = c.dmode ? 10 : (c.amode ? c.na : 4);
        const int done = (c.iters + per - 1) / per * per;
        if (c.dmode)
            printf("d-march NC=%2d (N = 16..%3d) sbo=%3d           : %6.1f cycles/MMA   err=%s\n", c.NC, 3 * c.NC, c.sbo, (double)cyc / done, cudaGetErrorString(e));
        else
            printf("N=%3d amode=%d na=%2d sbo=%3d                    : %6.1f cycles/MMA   (math floor %3.0f, operand-fetch model %4.1f)  err=%s\n", c.N,
                   c.amode, c.na, c.sbo, (double)cyc / done, 128.0 * c.N / 256, (4096.0 + 32.0 * c.N) / 128, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
