// Probe: can a tcgen05 shared-memory descriptor (K-major, SWIZZLE_128B) address a ROW-SHIFTED view of a
// halo tile, i.e. start address = base + r0*128 B (not 1024-aligned) and a stride between 8-row groups
// that is not 1024 B?  This decides whether a 3x3 convolution can run its 9 taps from ONE staged halo tile.
//   variants: base_offset field (bits 49-51) = 0  or  (start_addr >> 7) & 7 ; halo pitch 8 / 10 / 16 pixels.
// Data are written with the absolute-address swizzle: 16-byte chunk index ^= (smem_addr >> 7) & 7.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o desc_probe tools/desc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "../chore_b200/csrc/tc_common.cuh"

using namespace tc;

constexpr int kN = 64, kRows = 18 * 16;   // up to pitch 16

__device__ __forceinline__ void umma_f16_hi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %5};\nmov.b64 db, {%2, %6};\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}"
                 ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(a_hi), "r"(b_hi) : "memory");
}

__host__ __device__ inline int aval(int R, int c) { return ((R * 7 + c * 3) % 17) - 8; }
__host__ __device__ inline int wval(int n, int k) { return ((n * 5 + k * 11) % 13) - 6; }

// one launch = one (pitch, ky, kx, mode) case; out[128][64]
__global__ void __launch_bounds__(128, 1) probe_kernel(int pitch, int ky, int kx, int mode, float *out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *halo = smem;                       // kRows x 128 B
    uint8_t *wp = smem + kRows * 128;           // 64 x 128 B (1024-aligned since kRows*128 is)
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < kRows * 64; i += 128) {
        const int R = i / 64, c = i % 64;
        const uint32_t row_addr = smem_u32(halo) + R * 128;
        const uint32_t off = R * 128 + ((((c >> 3) ^ ((row_addr >> 7) & 7))) << 4) + (c & 7) * 2;
        *reinterpret_cast<__half *>(halo + off) = __float2half((float)aval(R, c));
    }
    for (int i = tid; i < kN * 64; i += 128) {
        const int n = i / 64, k = i % 64;
        *reinterpret_cast<__half *>(wp + sw128(n, k >> 3) + (k & 7) * 2) = __float2half((float)wval(n, k));
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        if (elect_one()) {
            const uint32_t start = smem_u32(halo) + (uint32_t)((ky * pitch + kx) * 128);
            const uint32_t sbo = (uint32_t)(pitch * 128) >> 4;
            uint32_t a_hi = sbo | (1u << 14) | (2u << 29);
            if (mode == 1) a_hi |= ((start >> 7) & 7u) << 17;
            const uint32_t a_lo = desc_lo(start), b_lo = desc_lo(smem_u32(wp));
            const uint32_t idesc = make_idesc(128, kN);
            for (int ks = 0; ks < 4; ++ks) umma_f16_hi(tmem, a_lo + 2 * ks, a_hi, b_lo + 2 * ks, kDescHi, idesc, ks != 0);
            umma_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < kN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[(size_t)tid * kN + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
    }
}


// ---- second probe: L2 -> SMEM bulk-copy bandwidth per SM (what bounds the weight / activation restream) ----
__global__ void __launch_bounds__(64, 1) bulk_bw_kernel(const uint8_t *src, size_t src_bytes, int iters, uint32_t chunk,
                                                       long long *cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full[4];
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&full[i], 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        size_t off = ((size_t)blockIdx.x * 7919 * chunk) % (src_bytes - chunk);
        off &= ~(size_t)127;
        for (int i = 0; i < iters + 4; ++i) {
            const int s = i & 3;
            if (i >= 4) mbar_wait(&full[s], ((i - 4) >> 2) & 1);
            if (i < iters) {
                mbar_arrive_expect_tx(&full[s], chunk);
                bulk_g2s(smem + (size_t)s * chunk, src + off, chunk, &full[s]);
                off += chunk;
                if (off + chunk > src_bytes) off = 0;
            }
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
}

static void bulk_bw() {
    const size_t bytes = 32u << 20;
    uint8_t *src; long long *cyc;
    cudaMalloc(&src, bytes); cudaMemset(src, 1, bytes);
    cudaMalloc(&cyc, 148 * sizeof(long long));
    const uint32_t chunk = 32768;
    cudaFuncSetAttribute(bulk_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 4 * chunk);
    const int grids[5] = {1, 8, 32, 128, 148};
    for (int w = 0; w < 2; ++w)
        for (int gi = 0; gi < 5; ++gi) {
            const int iters = 512;
            bulk_bw_kernel<<<grids[gi], 64, 1024 + 4 * chunk>>>(src, bytes, iters, chunk, cyc);
            cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, cyc, grids[gi] * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < grids[gi]; ++i) mx = h[i] > mx ? h[i] : mx;
            if (w) printf("BULK grid=%3d : %.1f B/clk/SM, %.0f B/clk chip\n", grids[gi], (double)iters * chunk / mx,
                          (double)iters * chunk / mx * grids[gi]);
        }
}

int main() {
    bulk_bw();
    float *d_out;
    cudaMalloc(&d_out, 128 * kN * sizeof(float));
    const size_t smem = 1024 + kRows * 128 + kN * 128;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::vector<float> h(128 * kN);
    const int pitches[3] = {8, 10, 16};
    for (int mode = 0; mode < 2; ++mode)
        for (int pi = 0; pi < 3; ++pi) {
            const int pitch = pitches[pi];
            int bad_cases = 0;
            for (int ky = 0; ky < 3; ++ky)
                for (int kx = 0; kx < 3; ++kx) {
                    if (pitch == 8 && kx != 0) continue;   // a pitch-8 halo has no room for a horizontal shift
                    probe_kernel<<<1, 128, smem>>>(pitch, ky, kx, mode, d_out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
                    int bad = 0;
                    for (int m = 0; m < 128; ++m)
                        for (int n = 0; n < kN; ++n) {
                            const int R = (m / 8 + ky) * pitch + (m % 8) + kx;
                            float ref = 0.f;
                            for (int k = 0; k < 64; ++k) ref += (float)(aval(R, k) * wval(n, k));
                            if (h[(size_t)m * kN + n] != ref) ++bad;
                        }
                    printf("mode=%d pitch=%2d ky=%d kx=%d mismatches=%d\n", mode, pitch, ky, kx, bad);
                    if (bad) ++bad_cases;
                }
            printf("SUMMARY base_offset_mode=%d pitch=%d : %s (%d failing shifts)\n", mode, pitch, bad_cases ? "FAIL" : "OK", bad_cases);
        }
    return 0;
}
