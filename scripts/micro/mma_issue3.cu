// Micro-benchmark 3: tcgen05.mma issue interval with a FULLY UNROLLED issue sequence whose per-instruction constants are compile-time
// immediates (mma_issue2 drives the sequence from a shared-memory table, i.e. one dependent LDS per instruction -- its 70 cycles are the
// loop, which is also what the kernel's VG_TC_DMLEAN=1 loop measured).  Separates three effects:
//   W   : number of rotating accumulator windows (an MMA that accumulates into columns a recent MMA wrote has to wait for it)
//   N   : instruction width (math floor 128*N/256, operand fetch (4096 + 32 N)/128)
//   A   : same A tile every time vs. a different 4 KB tile per instruction
// and times the exact d-march sequence of the convolution kernel (BD = 8, TD = 3, slice order (3 i) mod 10) with immediates.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/mma_issue3.bin scripts/micro/mma_issue3.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../van-gan_b200/csrc/tc_ptx.cuh"
using namespace tcp;

// MODE 0: W rotating windows of N columns, A fixed (DISTINCT = 0) or cycling through 8 tiles (DISTINCT = 1)
// MODE 1: d-march, tile m at column m*NC, BD = 8, TD = 3, permuted slice order
template <int MODE, int N, int W, int DISTINCT>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
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
    if (warp == 0) {
        const uint32_t leader = elect_one();
        const uint64_t adesc = make_desc(sb, 48 * 1024, 160);
        constexpr uint32_t TILE = 18 * 160 / 16;   // 16-byte units between A tiles (one d-slice of an 18 x 10 halo)
        long long t0 = clock64();
        if (leader) {
            if (MODE == 0) {
                const uint64_t bdesc = make_desc(sb + 100 * 1024, (uint32_t)N * 16, 128);
                const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
                for (int i = 0; i < iters; i += 16) {
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        tc_mma(tm + (uint32_t)((j % W) * N), adesc + (uint64_t)(DISTINCT ? (j & 7) * TILE : 0), bdesc, idesc, 1u);
                }
            } else {
                constexpr int NC = N, ED = 10;
                const uint64_t bdesc = make_desc(sb + 100 * 1024, (uint32_t)(3 * NC) * 16, 128);
                for (int i = 0; i < iters; i += ED) {
#pragma unroll
                    for (int j = 0; j < ED; j++) {
                        constexpr int dummy = 0;
                        const int s = (3 * j) % ED;
                        const int m_lo = s - 2 > 0 ? s - 2 : 0, m_hi = s < 7 ? s : 7;
                        tc_mma(tm + (uint32_t)(m_lo * NC), adesc + (uint64_t)(s * TILE), bdesc + (uint64_t)((2 - s + m_lo) * NC),
                               make_idesc_bf16(128, (m_hi - m_lo + 1) * NC, 0, 0), 1u);
                        (void)dummy;
                    }
                }
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

template <int MODE, int N, int W, int DISTINCT>
void run(long long* d, const char* label) {
    const int iters = MODE == 0 ? 16000 : 16000;
    cudaFuncSetAttribute(k<MODE, N, W, DISTINCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int rep = 0; rep < 2; rep++) {
        k<MODE, N, W, DISTINCT><<<148, 128, 200 * 1024>>>(iters, d);
        cudaDeviceSynchronize();
    }
    long long c;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    if (MODE == 0)
        printf("%-10s N=%3d windows=%d distinctA=%d : %6.1f cycles/MMA  (math floor %3.0f, operand fetch %4.1f)  %s\n", label, N, W, DISTINCT,
               (double)c / iters, 128.0 * N / 256, (4096.0 + 32.0 * N) / 128, e == cudaSuccess ? "" : cudaGetErrorString(e));
    else
        printf("%-10s d-march NC=%2d (N = NC..3NC, BD=8, TD=3, order (3i) mod 10) : %6.1f cycles/MMA  %s\n", label, N, (double)c / iters,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    run<0, 16, 1, 0>(d, "rotate");  run<0, 16, 2, 0>(d, "rotate");  run<0, 16, 4, 0>(d, "rotate");  run<0, 16, 8, 0>(d, "rotate");  run<0, 16, 16, 0>(d, "rotate");
    run<0, 48, 1, 0>(d, "rotate");  run<0, 48, 2, 0>(d, "rotate");  run<0, 48, 4, 0>(d, "rotate");  run<0, 48, 8, 0>(d, "rotate");
    run<0, 128, 1, 0>(d, "rotate"); run<0, 128, 2, 0>(d, "rotate"); run<0, 128, 4, 0>(d, "rotate");
    run<0, 16, 8, 1>(d, "distinctA"); run<0, 48, 8, 1>(d, "distinctA"); run<0, 96, 4, 1>(d, "distinctA"); run<0, 128, 4, 1>(d, "distinctA");
    run<0, 144, 3, 1>(d, "distinctA"); run<0, 256, 2, 1>(d, "distinctA");
    run<1, 16, 0, 1>(d, "dmarch"); run<1, 32, 0, 1>(d, "dmarch"); run<1, 48, 0, 1>(d, "dmarch");
    return 0;
}
