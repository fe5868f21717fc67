// microbenchmark: does a tcgen05.commit between MMAs cost tensor-pipe time?  (N=128, M=128, K=16, SS mode)
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../chore_b200/csrc/tc_common.cuh"
using namespace tc;

template <int XC, int XW>
__global__ void __launch_bounds__(128, 1) k(long long *out, int iters) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, dummy, done;
    __shared__ uint32_t tmem_base_s;
    for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&dummy, 1u << 20); mbar_init(&done, 1); mbar_arrive(&done); fence_barrier_init(); }
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (warp == 0) {
        const uint32_t base = desc_lo(smem_u32(smem));
        constexpr uint32_t idesc = make_idesc(128, 128);
        long long t0 = clock64();
        uint32_t par = 0;
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
#pragma unroll 8
                for (int m = 0; m < 480; ++m) {
                    const uint32_t a = base + ((m & 3) * 16384 >> 4) + 2 * ((m >> 2) & 3);
                    const uint32_t b = base + ((65536 + (m & 1) * 16384) >> 4) + 2 * ((m >> 2) & 3);
                    umma_f16(tmem_base + ((m >> 4) & 3) * 128, a, b, idesc, 1);
                    if (XC > 0 && (m & (XC - 1)) == XC - 1) umma_commit(&dummy);
                    if (XW > 0 && (m & (XW - 1)) == XW - 1) { mbar_wait(&done, 0); tc_fence_after(); }
                }
                umma_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, par); par ^= 1;
        }
        long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory"); }
}

template <int XC, int XW>
void run(long long *d) {
    const int smem = 136 * 1024, iters = 20;
    cudaFuncSetAttribute(k<XC, XW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<XC, XW><<<148, 128, smem>>>(d, iters);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return; }
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
    double s = 0; for (auto v : h) s += v;
    printf("commit every %3d MMAs, satisfied mbarrier wait + fence every %3d MMAs: %.1f cycles per MMA\n", XC, XW, s / 148 / (iters * 480.0));
}
int main() {
    long long *d; cudaMalloc(&d, 148 * 8);
    run<0, 0>(d); run<8, 0>(d); run<4, 0>(d); run<2, 0>(d); run<1, 0>(d); run<0, 8>(d); run<0, 4>(d); run<4, 4>(d); run<4, 8>(d);
    return 0;
}
