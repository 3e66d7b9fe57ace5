// Micro-benchmark: sustained issue interval of tcgen05.mma (cta_group::1, kind::f16, M=128, K=16) from one thread,
// as a function of N and of the amount of descriptor arithmetic between instructions.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/mma_issue.bin scripts/micro/mma_issue.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../van-gan_b200/csrc/tc_ptx.cuh"
using namespace tcp;

template <int MODE>
__global__ void __launch_bounds__(128, 1) k(int N, int iters, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x * 16; i < 160 * 1024; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(s_addr(&bar), 1); mbar_init_fence(); }
    if (warp == 0) tmem_alloc(s_addr(&slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 0) {
        const uint32_t leader = elect_one();
        const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
        const uint32_t sb = s_addr(smem);
        // A: planes of 16-byte cells, SBO 160 B (halo row pitch of 10 cells), LBO = 40 KB; B: SBO 128, LBO = N*16
        uint64_t adesc = make_desc(sb, 40 * 1024, 160);
        uint64_t bdesc = make_desc(sb + 96 * 1024, (uint32_t)N * 16, 128);
        long long t0 = clock64();
        if (leader) {
            for (int i = 0; i < iters; i += 16) {
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    if (MODE == 0) {
                        tc_mma(tm + (uint32_t)((j & 3) * N), adesc, bdesc, idesc, 1u);
                    } else {
                        // descriptor arithmetic per instruction, as the convolution issue loop does (tap shift + tile step)
                        tc_mma(tm + (uint32_t)((j & 3) * N), adesc + (uint32_t)((j >> 2) + (j & 3) * 180), bdesc + (uint32_t)((j >> 2) * 2 * N), idesc, 1u);
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

int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 16384;
    int Ns[] = {16, 32, 48, 64, 96, 128, 192, 256};
    for (int mode = 0; mode < 2; mode++)
        for (int N : Ns) {
            for (int rep = 0; rep < 2; rep++) {
                if (mode == 0) k<0><<<148, 128, 160 * 1024>>>(N, iters, d); else k<1><<<148, 128, 160 * 1024>>>(N, iters, d);
                cudaDeviceSynchronize();
            }
            long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
            printf("mode %d  M=128 N=%3d K=16: %.1f cycles/MMA  (floor 128*N/256 = %.0f)  err=%s\n", mode, N, (double)c / iters, 128.0 * N / 256, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
