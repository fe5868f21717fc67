// Shared pieces of the tensor-core query kernels (query_tc.cu: cta_group::1 forward + backward, query_tc2.cu: the
// cta_group::2 forward): tile constants, the parameter block, point source + projection (exact fp32 op order of
// model/camera.py:64-65,75-78), bilinear taps and the batched gather of one A-operand k-block.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

#include <cuda_fp16.h>

namespace {

constexpr int kTileM = 128;                  // points per tile = TMEM lanes
constexpr int kKB = 64;                      // channels per k-block (one 128 B swizzle atom of fp16)
constexpr int kPanelBytes = kTileM * kKB * 2;          // 16 KB: 128 rows x 64 fp16
constexpr int kStageA = 2 * kPanelBytes;               // hi + lo
constexpr int kNA = 2, kNACT = 3, kNW = 4;             // ring depths
constexpr int kL1Blocks = 6;                 // 4 x feat(64) | skip(64) | xyz(3, one k-step)
constexpr int kBigUnits = kL1Blocks * 4 * 2 + 2 * (4 * 2 * 2);       // 48 + 32 = 80 panels of 16 KB (layers 1-3)
constexpr int kUnitsPerTile = kBigUnits + 4;                          // + one 8 KB panel per head for the last layer
constexpr int kSmallPanelBytes = 8192;                                // [kb0 hi | kb0 lo | kb1 hi | kb1 lo] x (16 rows x 128 B)
constexpr size_t kStreamBytes = (size_t)kBigUnits * kPanelBytes + 4 * (size_t)kSmallPanelBytes;
constexpr int kThreads = 14 * 32;
constexpr int kGatherWarp0 = 2, kEpiWarp0 = 6;
constexpr size_t kSmemBytes = 1024 + (size_t)kNA * kStageA + (size_t)kNACT * kStageA + (size_t)kNW * kPanelBytes + 512;
constexpr int kBwdUnits = 12 + 4 + 4 + 4 + 4 + 12;     // backward weight stream per head: F1, F2, F3, B3, B2, B1 panels of 16 KB
constexpr int kGXLd = 384;                             // leading dimension of the gX scratch (query_bwd_tc.cu)

constexpr float kFx = static_cast<float>(979.7844 / 2048. * 2048);
constexpr float kFy = static_cast<float>(979.840 / 2048. * 2048);
constexpr float kCx = static_cast<float>(1018.952 / 2048. * 2048);
constexpr float kCy = static_cast<float>(779.486 / 2048. * 2048);

struct TcParams {
    const float *feat, *skip;
    int fh, fw;
    const float *points;
    const float *crop_center;
    int B;
    long long N, n_start, n_count;
    int grid_mode, ry, rz, batch_index;
    double step[3], bmin[3];
    unsigned head_mask;
    float *out[4];
    unsigned char *in_img;
    const unsigned char *wstream;     // kUnitsPerTile x 16 KB pre-swizzled fp16 panels
    const float *b1, *b2, *b3;        // [4][128]
    const float *w4, *b4;             // [4][16][128], [4][16] fp32
    long long tiles_per_b, total_tiles;
    unsigned long long *dbg;          // optional [grid][16] wait-cycle counters (CHORE_B200_TC_TRACE)
    // backward (query_bwd_tc_kernel): one launch per head
    // work item = (tile, slot); slot s evaluates head bwd_heads[s] into its own gX buffer (gX + s * gx_slot_stride), so
    // several heads run concurrently in ONE launch; with a single slot the heads are launched one after the other
    // and bwd_accumulate adds into the shared buffer
    int bwd_nslots;
    int bwd_heads[4];
    int bwd_accumulate;               // gX += (heads after the first; single-slot mode only)
    const float *g_heads[4];          // upstream gradient per slot (B, nout, N)
    const unsigned char *wstream_bwd; // the four per-head 40-panel streams, head-major
    long long gx_slot_stride;         // floats between the gX buffers of two slots
    float *gX;                        // (B*N, 384) gradient w.r.t. the (permuted) feature column
    float *g_points;                  // geometry kernel output (B, N, 3)
};

__host__ __device__ __forceinline__ int head_out_tc(int h) { return h == 0 ? 2 : (h == 1 ? 9 : (h == 2 ? 14 : 6)); }

using namespace tc;

// ---------------------------------------------------------------------------------------------
// point source + projection (exact fp32 op order of model/camera.py:64-65,75-78)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_point(const TcParams &q, int b, long long n, float &x, float &y, float &z) {
    if (q.grid_mode) {
        const long long g = q.n_start + n;
        long long ix, iy, iz;
        if (g < 0x7fffffffll) {          // 32-bit division is ~5x cheaper than the 64-bit sequence
            const unsigned g32 = (unsigned)g, t32 = g32 / (unsigned)q.rz;
            iz = g32 - t32 * (unsigned)q.rz;
            const unsigned x32 = t32 / (unsigned)q.ry;
            iy = t32 - x32 * (unsigned)q.ry; ix = x32;
        } else {
            iz = g % q.rz;
            const long long t = g / q.rz;
            iy = t % q.ry; ix = t / q.ry;
        }
        x = (float)__dadd_rn(__dmul_rn(q.step[0], (double)ix), q.bmin[0]);
        y = (float)__dadd_rn(__dmul_rn(q.step[1], (double)iy), q.bmin[1]);
        z = (float)__dadd_rn(__dmul_rn(q.step[2], (double)iz), q.bmin[2]);
    } else {
        const float *p = q.points + ((size_t)b * q.N + q.n_start + n) * 3;
        x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
    }
}
__device__ __forceinline__ void project_tc(float x, float y, float z, float ccx, float ccy, float &nx, float &ny) {
    float px = __fadd_rn(__fdiv_rn(__fmul_rn(kFx, x), z), kCx);
    float py = __fadd_rn(__fdiv_rn(__fmul_rn(kFy, y), z), kCy);
    px = __fsub_rn(__fadd_rn(600.0f, px), ccx);
    py = __fsub_rn(__fadd_rn(600.0f, py), ccy);
    nx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, px), 1200.0f), 1.0f);
    ny = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, py), 1200.0f), 1.0f);
}
struct TapsTc {
    int x0, y0;
    float w[4];        // nw, ne, sw, se weights (0 for out-of-range taps)
    unsigned valid;
};
__device__ __forceinline__ TapsTc make_taps_tc(float nx, float ny, int H, int W) {
    TapsTc t;
    const float ix = __fmul_rn(__fadd_rn(nx, 1.0f), 0.5f * (float)(W - 1));
    const float iy = __fmul_rn(__fadd_rn(ny, 1.0f), 0.5f * (float)(H - 1));
    t.valid = 0; t.x0 = 0; t.y0 = 0;
    t.w[0] = t.w[1] = t.w[2] = t.w[3] = 0.f;
    if (ix > -1.0f && ix < (float)W && iy > -1.0f && iy < (float)H) {
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        t.x0 = (int)fx0; t.y0 = (int)fy0;
        const float wx = ix - fx0, wy = iy - fy0;
        const bool xl = t.x0 >= 0, xr = t.x0 + 1 < W, yt = t.y0 >= 0, yb = t.y0 + 1 < H;
        t.valid = (unsigned)(xl && yt) | ((unsigned)(xr && yt) << 1) | ((unsigned)(xl && yb) << 2) | ((unsigned)(xr && yb) << 3);
        t.w[0] = (1.f - wy) * (1.f - wx); t.w[1] = (1.f - wy) * wx; t.w[2] = wy * (1.f - wx); t.w[3] = wy * wx;
    }
    return t;
}

// One 64-channel k-block of the A operand for the 32 rows of gather warp g: a half-warp owns one point
// (16 lanes x 4 channels), and the 4 bilinear taps of kGatherBatch points are issued back to back before any of
// them is consumed, so kGatherBatch * 4 independent 16-byte loads per lane cover the L2 latency (the previous
// one-point-at-a-time loop paid one full L2 round trip per point and left the MMA warp waiting on a_full).
// Arithmetic is unchanged: v = sum_k tap_k * w_k in tap order nw, ne, sw, se, invalid taps contribute exactly 0.
constexpr int kGatherBatch = 4;
// ROWS = tile rows owned by the calling warp (32 with four gather warps, 16 with eight), BATCH = points in flight per half-warp.
template <int ROWS, int BATCH>
__device__ __forceinline__ void gather_kblock_t(const float *__restrict__ base, int H, int W, int C, float my_nx, float my_ny,
                                                int g, int half, int l16, uint8_t *hi, uint8_t *lo) {
    constexpr int kGatherBatch = BATCH;
#pragma unroll 1
    for (int it0 = 0; it0 < ROWS / 2; it0 += kGatherBatch) {
        float4 v[kGatherBatch][4];
        float wgt[kGatherBatch][4];
#pragma unroll
        for (int j = 0; j < kGatherBatch; ++j) {
            const int src = (it0 + j) * 2 + half;
            const float nx = __shfl_sync(0xffffffffu, my_nx, src), ny = __shfl_sync(0xffffffffu, my_ny, src);
            const TapsTc t = make_taps_tc(nx, ny, H, W);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                wgt[j][k] = t.w[k];
                v[j][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t.valid & (1u << k))
                    v[j][k] = __ldg(reinterpret_cast<const float4 *>(base + ((size_t)(t.y0 + (k >> 1)) * W + (t.x0 + (k & 1))) * C));
            }
        }
#pragma unroll
        for (int j = 0; j < kGatherBatch; ++j) {
            const int r = g * ROWS + (it0 + j) * 2 + half;
            float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v0 = fmaf(v[j][k].x, wgt[j][k], v0); v1 = fmaf(v[j][k].y, wgt[j][k], v1);
                v2 = fmaf(v[j][k].z, wgt[j][k], v2); v3 = fmaf(v[j][k].w, wgt[j][k], v3);
            }
            uint32_t h01, l01, h23, l23;
            split2(v0, v1, h01, l01);
            split2(v2, v3, h23, l23);
            const uint32_t off = sw128(r, l16 >> 1) + (l16 & 1) * 8;
            *reinterpret_cast<uint2 *>(hi + off) = make_uint2(h01, h23);
            *reinterpret_cast<uint2 *>(lo + off) = make_uint2(l01, l23);
        }
    }
}

// ---- variant with the taps evaluated once per tile and map --------------------------------------------------------
// Lane L keeps, for the point it projects, the packed tap descriptor of a map: element offset of tap (y0, x0) from the
// map base -- a multiple of C >= 64, so the 4 validity bits ride in its low bits -- and the two fractional weights.
// gather_kblock_p broadcasts them with 3 shuffles per point instead of re-deriving them (2 shuffles + ~45 ALU ops).
struct LaneTaps { int off_valid; float wx, wy; };
__device__ __forceinline__ LaneTaps make_lane_taps(float nx, float ny, int H, int W, int C) {
    const TapsTc t = make_taps_tc(nx, ny, H, W);
    const float ix = __fmul_rn(__fadd_rn(nx, 1.0f), 0.5f * (float)(W - 1)), iy = __fmul_rn(__fadd_rn(ny, 1.0f), 0.5f * (float)(H - 1));
    LaneTaps l;
    l.off_valid = t.valid ? ((t.y0 * W + t.x0) * C) | (int)t.valid : 0;     // two's complement: the OR also works for negative offsets
    l.wx = t.valid ? ix - floorf(ix) : 0.f;
    l.wy = t.valid ? iy - floorf(iy) : 0.f;
    return l;
}
template <int ROWS, int BATCH>
__device__ __forceinline__ void gather_kblock_p(const float *__restrict__ base, int rowC, int C, const LaneTaps mine,
                                                int g, int half, int l16, uint8_t *hi, uint8_t *lo) {
#pragma unroll 1
    for (int it0 = 0; it0 < ROWS / 2; it0 += BATCH) {
        float4 v[BATCH][4];
        float wx[BATCH], wy[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
            const int src = (it0 + j) * 2 + half;
            const int ov = __shfl_sync(0xffffffffu, mine.off_valid, src);
            wx[j] = __shfl_sync(0xffffffffu, mine.wx, src);
            wy[j] = __shfl_sync(0xffffffffu, mine.wy, src);
            const float *p = base + (ov & ~15);
            v[j][0] = (ov & 1) ? __ldg(reinterpret_cast<const float4 *>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[j][1] = (ov & 2) ? __ldg(reinterpret_cast<const float4 *>(p + C)) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[j][2] = (ov & 4) ? __ldg(reinterpret_cast<const float4 *>(p + rowC)) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[j][3] = (ov & 8) ? __ldg(reinterpret_cast<const float4 *>(p + rowC + C)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
            const int r = g * ROWS + (it0 + j) * 2 + half;
            const float ux = 1.f - wx[j], uy = 1.f - wy[j];
            const float w[4] = {uy * ux, uy * wx[j], wy[j] * ux, wy[j] * wx[j]};
            float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v0 = fmaf(v[j][k].x, w[k], v0); v1 = fmaf(v[j][k].y, w[k], v1);
                v2 = fmaf(v[j][k].z, w[k], v2); v3 = fmaf(v[j][k].w, w[k], v3);
            }
            uint32_t h01, l01, h23, l23;
            split2(v0, v1, h01, l01);
            split2(v2, v3, h23, l23);
            const uint32_t off = sw128(r, l16 >> 1) + (l16 & 1) * 8;
            *reinterpret_cast<uint2 *>(hi + off) = make_uint2(h01, h23);
            *reinterpret_cast<uint2 *>(lo + off) = make_uint2(l01, l23);
        }
    }
}

__device__ __forceinline__ void gather_kblock(const float *__restrict__ base, int H, int W, int C, float my_nx, float my_ny,
                                              int g, int half, int l16, uint8_t *hi, uint8_t *lo) {
    gather_kblock_t<32, kGatherBatch>(base, H, W, C, my_nx, my_ny, g, half, l16, hi, lo);
}

}   // namespace
