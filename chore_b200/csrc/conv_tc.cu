// Encoder convolutions (3x3 pad 1 / 1x1, stride 1) on the 5th-generation tensor cores.
//
// Implicit GEMM per CTA: M = 128 output pixels (a 16 x 8 patch), N = Cout (<= 256), K = taps x Cin.
//   act_split_kernel   relu(groupnorm(x)) (or identity) -> two fp16 planes hi / lo, channels-last
//                      (x = hi + lo to 22 bits; fp32-faithful 3-term products as in query_tc.cu)
//   conv_tc_kernel     warp 0: TMA (cp.async.bulk.tensor.4d, 128B swizzle, zero fill = the conv's
//                      zero padding) loads the shifted 16x8x64 activation boxes of every tap and
//                      bulk-copies the pre-swizzled weight panels; warp 1 issues tcgen05.mma
//                      (kind::f16, fp32 accumulators in TMEM); warps 2-5: epilogue (bias, raw copy
//                      for the next GroupNorm, residual add, concat slice store).
// Replaces the F.conv2d + F.group_norm + ReLU triples of ConvBlock.forward (model/net_util.py:374-396)
// and the 1x1 convs of HGFilter.forward (model/HGFilters.py:173-183).
#include "common.cuh"
#include "tc_common.cuh"

#include <cuda.h>
#include <cstdlib>
#include <cstring>

using namespace tc;

namespace {

constexpr int kGroups = 32;
constexpr int kPatchH = 16, kPatchW = 8;            // 128 pixels per CTA
constexpr int kThreadsConv = 6 * 32;

// ------------------------------------------------------------------------------------------
// relu(gn(x)) -> fp16 hi / lo planes [B][H][W][Cp] (Cp = channels padded to a multiple of 64)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) act_split_kernel(const float *__restrict__ in, int ld, int off, int C, int Cp,
                                                        int HW, const double *__restrict__ sums,
                                                        const float *__restrict__ gamma, const float *__restrict__ beta,
                                                        __half *__restrict__ hi, __half *__restrict__ lo, size_t total8) {
    __shared__ float sc[256], sh[256];
    const int b = blockIdx.y;
    if (sums) {
        const int cpg = C / kGroups;
        const double inv_n = 1.0 / ((double)HW * cpg);
        for (int c = threadIdx.x; c < C; c += 256) {
            const int g = c / cpg;
            const double mean = sums[(size_t)b * kGroups * 2 + g * 2] * inv_n;
            double var = sums[(size_t)b * kGroups * 2 + g * 2 + 1] * inv_n - mean * mean;
            var = var > 0.0 ? var : 0.0;
            const float rstd = (float)(1.0 / sqrt(var + 1e-5));
            sc[c] = __ldg(gamma + c) * rstd;
            sh[c] = __ldg(beta + c) - (float)mean * sc[c];
        }
        __syncthreads();
    }
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;     // one 8-channel chunk of one pixel of image b
    if (i >= total8) return;
    const int c8n = Cp / 8;
    const int c = (int)(i % c8n) * 8;
    const size_t pix = i / c8n;
    uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
    if (c < C) {
        const float *src = in + ((size_t)b * HW + pix) * ld + off + c;
        const float4 v0 = __ldg(reinterpret_cast<const float4 *>(src)), v1 = __ldg(reinterpret_cast<const float4 *>(src) + 1);
        float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        if (sums) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(fmaf(v[j], sc[c + j], sh[c + j]), 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) split2(v[2 * j], v[2 * j + 1], h[j], l[j]);
    }
    const size_t o = (((size_t)b * HW + pix) * Cp + c) / 8;
    reinterpret_cast<uint4 *>(hi)[o] = make_uint4(h[0], h[1], h[2], h[3]);
    reinterpret_cast<uint4 *>(lo)[o] = make_uint4(l[0], l[1], l[2], l[3]);
}

// ------------------------------------------------------------------------------------------
// tensor-core implicit GEMM
// ------------------------------------------------------------------------------------------
struct ConvTcParams {
    int H, W, B;
    int KS, kblocks;            // taps per side, 64-channel k-blocks (Cp / 64)
    int N;                      // output channels of this launch (= Cout, 32..256)
    const unsigned char *wstream;   // [tap][kb][hi|lo] panels of N x 128 B
    const float *bias;
    float *out; int ld_out, off_out;
    const float *res; int ld_res, off_res;
    float *raw; int ld_raw, off_raw;
};

__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}

template <int STAGES>
__global__ void __launch_bounds__(kThreadsConv, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap map_hi,
                                                                   const __grid_constant__ CUtensorMap map_lo,
                                                                   const ConvTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t panelW = (uint32_t)p.N * 128u;                  // one weight panel (hi or lo)
    const uint32_t stage_bytes = 32768u + 2u * panelW;             // A hi | A lo | W hi | W lo
    struct Bars { uint64_t full[STAGES], empty[STAGES], acc; uint32_t tmem_base; };
    Bars *bars = reinterpret_cast<Bars *>(smem + (size_t)STAGES * stage_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (p.W + kPatchW - 1) / kPatchW;
    const int x0 = (blockIdx.x % tiles_x) * kPatchW, y0 = (blockIdx.x / tiles_x) * kPatchH;
    const int b = blockIdx.z;
    const int pad = p.KS / 2;
    const int nblk = p.KS * p.KS * p.kblocks;
    const uint32_t tmem_cols = p.N <= 32 ? 32u : (p.N <= 64 ? 64u : (p.N <= 128 ? 128u : 256u));

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
        mbar_init(&bars->acc, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, bars->tmem_base, 0);

    if (warp == 0) {
        // ---------------- producer ----------------
        for (int i = 0; i < nblk; ++i) {
            const int s = i % STAGES;
            mbar_wait(&bars->empty[s], ((i / STAGES) & 1) ^ 1);
            if (elect_one()) {
                const int tap = i / p.kblocks, kb = i - tap * p.kblocks;
                const int dy = tap / p.KS - pad, dx = tap % p.KS - pad;
                uint8_t *st = smem + (size_t)s * stage_bytes;
                mbar_arrive_expect_tx(&bars->full[s], stage_bytes);
                tma_load_4d(st, &map_hi, kb * 64, x0 + dx, y0 + dy, b, &bars->full[s]);
                tma_load_4d(st + 16384, &map_lo, kb * 64, x0 + dx, y0 + dy, b, &bars->full[s]);
                const unsigned char *w = p.wstream + (size_t)i * 2 * panelW;
                bulk_g2s(st + 32768, w, panelW, &bars->full[s]);
                bulk_g2s(st + 32768 + panelW, w + panelW, panelW, &bars->full[s]);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (converged warp, one elected lane) ----------------
        const uint32_t idesc = make_idesc(128, p.N);
        const uint32_t base_lo = desc_lo(smem_u32(smem));
        for (int i = 0; i < nblk; ++i) {
            const int s = i % STAGES;
            mbar_wait(&bars->full[s], (i / STAGES) & 1);
            tc_fence_after();
            const uint32_t a_hi = base_lo + (uint32_t)s * (stage_bytes >> 4), a_lo = a_hi + (16384u >> 4);
            const uint32_t w_hi = a_hi + (32768u >> 4), w_lo = w_hi + (panelW >> 4);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    umma_f16(tmem_base, a_hi + 2 * ks, w_hi + 2 * ks, idesc, (i | ks) != 0);
                    umma_f16(tmem_base, a_lo + 2 * ks, w_hi + 2 * ks, idesc, 1);
                    umma_f16(tmem_base, a_hi + 2 * ks, w_lo + 2 * ks, idesc, 1);
                }
                umma_commit(&bars->empty[s]);
                if (i == nblk - 1) umma_commit(&bars->acc);
            }
            __syncwarp();
        }
    } else {
        // ---------------- epilogue: one thread per output pixel ----------------
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const int gy = y0 + (r >> 3), gx = x0 + (r & 7);
        const bool valid = gy < p.H && gx < p.W;
        mbar_wait(&bars->acc, 0);
        tc_fence_after();
        const size_t pix = ((size_t)b * p.H + gy) * p.W + gx;
        for (int c0 = 0; c0 < p.N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + c0, v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    float4 o = make_float4(__uint_as_float(v[c4 * 4]), __uint_as_float(v[c4 * 4 + 1]),
                                           __uint_as_float(v[c4 * 4 + 2]), __uint_as_float(v[c4 * 4 + 3]));
                    const int c = c0 + c4 * 4;
                    if (p.bias) {
                        const float4 bb = __ldg(reinterpret_cast<const float4 *>(p.bias + c));
                        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                    }
                    if (p.raw) *reinterpret_cast<float4 *>(p.raw + pix * p.ld_raw + p.off_raw + c) = o;
                    if (p.res) {
                        const float4 rr = *reinterpret_cast<const float4 *>(p.res + pix * p.ld_res + p.off_res + c);
                        o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
                    }
                    *reinterpret_cast<float4 *>(p.out + pix * p.ld_out + p.off_out + c) = o;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// cuTensorMapEncodeTiled is resolved through the runtime (cudaGetDriverEntryPoint) so that the library
// does not link libcuda.so: it must still dlopen on a CPU-only build box.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

int make_map(CUtensorMap *map, const void *base, int Cp, int W, int H, int B) {
    const cuuint64_t gdim[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t gstride[3] = {(cuuint64_t)Cp * 2, (cuuint64_t)W * Cp * 2, (cuuint64_t)H * W * Cp * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)kPatchW, (cuuint32_t)kPatchH, 1};
    const cuuint32_t estride[4] = {1, 1, 1, 1};
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) {
        chore_set_error("cuTensorMapEncodeTiled is not available from this driver");
        return CHORE_ERR_CUDA;
    }
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(base), gdim, gstride, box, estride,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        chore_set_error("cuTensorMapEncodeTiled failed with %d (Cp=%d W=%d H=%d B=%d)", (int)r, Cp, W, H, B);
        return CHORE_ERR_CUDA;
    }
    return CHORE_OK;
}

}   // namespace

// CHORE_B200_ENCODER selects the encoder implementation: "hx" (default: conv_hx.cu + encoder_hx.cu), "tc1" (the first
// tensor-core path of this file, kept for A/B comparisons) or "simt" (fp32 CUDA-core convolutions, cross-check only)
const char *encoder_mode() {
    static const char *mode = [] {
        const char *e = getenv("CHORE_B200_ENCODER");
        if (e != nullptr && strcmp(e, "simt") == 0) return "simt";
        if (e != nullptr && strcmp(e, "tc1") == 0) return "tc1";
        return "hx";
    }();
    return mode;
}
bool encoder_use_tensor_cores() { return strcmp(encoder_mode(), "tc1") == 0; }

// weights (Cout, Cin, kh, kw) fp32 -> [tap][kb][hi|lo] panels of Cout rows x 64 k fp16, 128B swizzled.
// Packed on the device: the host-side fp16 conversions of 18 M parameters took tens of seconds.
__global__ void __launch_bounds__(256) pack_weights_kernel(const float *__restrict__ w, int cout, int cin, int kh, int kw,
                                                           int kbs, unsigned char *__restrict__ out, size_t total) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int k = (int)(i % 64);
    size_t t = i / 64;
    const int n = (int)(t % cout); t /= cout;
    const int kb = (int)(t % kbs);
    const int tap = (int)(t / kbs);
    const int ci = kb * 64 + k;
    float v = ci < cin ? w[(((size_t)n * cin + ci) * kh + tap / kw) * kw + tap % kw] : 0.f;
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    const __half hh = __float2half_rn(v);
    const __half ll = __float2half_rn(v - __half2float(hh));
    const size_t panel = (size_t)cout * 128;
    unsigned char *hi = out + ((size_t)(tap * kbs + kb) * 2) * panel, *lo = hi + panel;
    const size_t off = (size_t)n * 128 + (size_t)(((k >> 3) ^ (n & 7)) << 4) + (size_t)(k & 7) * 2;
    *reinterpret_cast<__half *>(hi + off) = hh;
    *reinterpret_cast<__half *>(lo + off) = ll;
}

int conv_tc_pack_weights(chore_handle *h, const float *w, int cout, int cin, int kh, int kw, unsigned char **dev) {
    const int cp = (cin + 63) / 64 * 64, kbs = cp / 64;
    const size_t panel = (size_t)cout * 128, bytes = (size_t)kh * kw * kbs * 2 * panel;
    const size_t nw = (size_t)cout * cin * kh * kw;
    float *tmp = nullptr;
    CHORE_CUDA(cudaMalloc(&tmp, nw * sizeof(float)));
    CHORE_CUDA(cudaMemcpy(tmp, w, nw * sizeof(float), cudaMemcpyHostToDevice));
    if (int rc = chore_dev_alloc(h, reinterpret_cast<void **>(dev), bytes)) return rc;
    const size_t total = (size_t)kh * kw * kbs * cout * 64;
    CHORE_LAUNCH(pack_weights_kernel, (unsigned)((total + 255) / 256), 256, 0, 0, tmp, cout, cin, kh, kw, kbs, *dev, total);
    CHORE_CUDA(cudaDeviceSynchronize());
    CHORE_CUDA(cudaFree(tmp));
    return CHORE_OK;
}

// launches act_split (unless planes are given) + conv_tc.  `planes` must hold 2 * B*H*W*Cp halves.
int conv_tc_launch(const ConvTcArgs &a, cudaStream_t st) {
    const int Cp = (a.Cin + 63) / 64 * 64;
    const int HW = a.H * a.W;
    __half *hi = reinterpret_cast<__half *>(a.planes), *lo = hi + (size_t)a.B * HW * Cp;
    {
        const size_t total8 = (size_t)HW * (Cp / 8);
        dim3 grid((unsigned)((total8 + 255) / 256), a.B);
        CHORE_LAUNCH(act_split_kernel, grid, 256, 0, st, a.in, a.ld_in, a.off_in, a.Cin, Cp, HW, a.gn_sums, a.gamma, a.beta, hi,
                     lo, total8);
    }
    CUtensorMap map_hi, map_lo;
    if (int rc = make_map(&map_hi, hi, Cp, a.W, a.H, a.B)) return rc;
    if (int rc = make_map(&map_lo, lo, Cp, a.W, a.H, a.B)) return rc;
    ConvTcParams p{};
    p.H = a.H; p.W = a.W; p.B = a.B; p.KS = a.KS; p.kblocks = Cp / 64; p.N = a.Cout;
    p.wstream = a.wstream; p.bias = a.bias;
    p.out = a.out; p.ld_out = a.ld_out; p.off_out = a.off_out;
    p.res = a.res; p.ld_res = a.ld_res; p.off_res = a.off_res;
    p.raw = a.raw; p.ld_raw = a.ld_raw; p.off_raw = a.off_raw;
    const int tiles = ((a.W + kPatchW - 1) / kPatchW) * ((a.H + kPatchH - 1) / kPatchH);
    dim3 grid(tiles, 1, a.B);
    const size_t stage = 32768 + 2 * (size_t)a.Cout * 128;
    if (a.Cout <= 128) {
        const size_t smem = 1024 + 3 * stage + 128;
        CHORE_ONCE_PER_DEVICE(cudaFuncSetAttribute(conv_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 3 * 65536 + 128));
        CHORE_LAUNCH(conv_tc_kernel<3>, grid, kThreadsConv, smem, st, map_hi, map_lo, p);
    } else {
        const size_t smem = 1024 + 2 * stage + 128;
        CHORE_ONCE_PER_DEVICE(cudaFuncSetAttribute(conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 2 * 98304 + 128));
        CHORE_LAUNCH(conv_tc_kernel<2>, grid, kThreadsConv, smem, st, map_hi, map_lo, p);
    }
    return CHORE_OK;
}
