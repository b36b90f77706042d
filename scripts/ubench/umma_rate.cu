// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16, SS mode) as a function of N with the
// operands resident in shared memory (no TMA traffic).  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -I../../eyediseasesegmentation_b200/csrc umma_rate.cu -o umma_rate
#include "tc_ptx.cuh"
#include <cstdio>
using namespace eds;

__global__ void __launch_bounds__(128, 1) rate_kernel(int n_mma, int bn, int iters, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1 && lane == 0) {
        const uint32_t sa = smem_u32(smem), sb = sa + 16384;
        const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint32_t phase = 0;
        long long best = 1ll << 60;
        for (int it = 0; it < iters; ++it) {
            const long long t0 = clock64();
            for (int i = 0; i < n_mma; ++i)
                umma_bf16(tmem, make_desc(sa + (i & 3) * 32, desc_hi), make_desc(sb + (i & 3) * 32, desc_hi), idesc, i > 0);
            umma_commit(&bar);
            mbar_wait(&bar, phase);
            phase ^= 1u;
            const long long t1 = clock64();
            if (t1 - t0 < best) best = t1 - t0;
        }
        if (blockIdx.x == 0) out[0] = best;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int n_mma = 2048;
    for (int grid : {1, 148})
        for (int bn : {16, 32, 64, 128, 192, 256}) {
            rate_kernel<<<grid, 128, 50 * 1024>>>(n_mma, bn, 3, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h = 0;
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("grid %3d  M128 N%3d K16: %.1f cycles/MMA (ideal tensor %d, A+B bytes %d) %s\n", grid, bn,
                   (double)h / n_mma, bn / 2, 4096 + 32 * bn, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
