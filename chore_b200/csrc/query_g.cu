// Dense-grid point query with the first MLP layer folded into the feature maps.
//
// The first layer of every head (model/chore.py:74-85, conv1d 323 -> 128) is linear, and so is grid_sample
// (model/geometry.py:12: bilinear, zeros padding): for the 320 sampled channels
//     W1 . sum_k w_k F[p_k]  =  sum_k w_k (W1 . F[p_k]).
// For a dense grid (Generator.get_grid_samples, recon/generator.py:243-267: 256^3 = 16.7 M points on a 128^2 / 256^2
// pixel map) it is ~250 x cheaper to apply W1 once per PIXEL than once per POINT:
//
//   project_maps   G_f[px][512] = W1[:, 0:256] . feat[px]      (fh x fw pixels,  4 heads x 128 hidden units)
//                  G_s[px][512] = W1[:, 259:323] . skip[px]    (2fh x 2fw pixels)
//                  two 1x1 convolutions per map on the encoder's tcgen05 kernel (conv_hx.cu: 3-term fp16 split, fp32
//                  accumulation) -- 8.6 GFLOP per image instead of 2.8 TFLOP of layer-1 MMAs per 16.7 M points
//   query_g_kernel h1 = relu(sample(G_f) + sample(G_s) + W1[:, 256:259] . (x, y, z - 2.2) + b1) evaluated by 8 gather
//                  warps straight into the A operand of layer 2 (fp16 hi / lo, 128B-swizzled K-major); layers 2, 3 and
//                  the output layer run on tcgen05 exactly as in query_tc_kernel (same weight panels, same epilogue).
//
// The MMA count per 128-point tile drops from 540 to 252 and the per-tile weight stream from 1.34 MB to 0.54 MB; the
// gather reads 8 taps x 512 B per (point, head) instead of 4 taps x 1.28 KB per point.  Results differ from the
// per-point evaluation only by the order of fp32 additions (parity tests: same 1e-4 bar, measured ~1e-6).
//
// Warp roles (19 warps, one persistent CTA per SM): 0 weight producer, 1-2 MMA issuers (heads 0,2 / 1,3), 3-10
// epilogue, 11-18 gather.  All of them walk ONE static 12-step schedule per tile,
//     L2h0 L2h1 L3h0 L2h2 L3h1 L4h0 L2h3 L3h2 L4h1 L3h3 L4h2 L4h3
// (software pipelined over the heads: the MMAs of a step only depend on epilogues of steps at least two positions back, and
// the gather delivers the heads in the order the L2 steps consume them); every wait of every role is on something an
// EARLIER step of the schedule produces, so the shared rings cannot deadlock whatever the head mask.
#include "query_tc_shared.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

constexpr int kGC = 512;                       // channels of the projected maps: head * 128 + hidden unit
constexpr int kNG = 2, kNAG = 3;               // ring depths: gathered layer-2 operand blocks / epilogue activation blocks
constexpr int kGatWarps = 8;
constexpr int kEpi0 = 3, kGat0 = 11;
constexpr int kThreadsG = (kGat0 + kGatWarps) * 32;
constexpr size_t kSmemG = 1024 + (size_t)(kNG + kNAG) * kStageA + (size_t)kNW * kPanelBytes + 512;
constexpr int kSteps = 12;
constexpr unsigned long long kSeq = 0xFEBDA7C96854ull;       // 4 bits per step: layer << 2 | head (layer 1..3 = L2, L3, L4)
constexpr int kGBatch = 1;                     // points in flight per half-warp; 2 spills and doubles the L1 footprint: 101 k instead of 57 k cycles per tile

// The "full" barriers exist once per issuer.  A parity wait is only meaningful for a waiter that is at most one phase
// ahead of the barrier; with ONE barrier per ring slot and two issuers taking turns on it, the issuer whose turn comes
// later would test a phase that is two ahead and sail through (an uninitialised operand, then a deadlock).  Producers
// therefore arrive on the barrier of the issuer that owns the step, and every issuer tracks the parity of its own uses.
struct BarsG {
    uint64_t g_full[2][kNG], g_empty[kNG];
    uint64_t w_full[2][kNW], w_empty[kNW];
    uint64_t act_full[2][kNAG], act_empty[kNAG];
    uint64_t tm_full[4], tm_empty[4];
    uint32_t tmem_base;
};

struct GParams {
    TcParams q;
    const float *gf, *gs;      // projected maps of image q.batch_index: [fh*fw][512], [4*fh*fw][512]
    const float *wz;           // [3][512]: layer-1 weights of x, y, z - 2.2
    int issuers;
};

__global__ void __launch_bounds__(kThreadsG, 1) query_g_kernel(const GParams gp) {
    const TcParams &q = gp.q;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *ringG = smem;                                      // [kNG][hi 16K | lo 16K]
    uint8_t *ringAct = ringG + (size_t)kNG * kStageA;           // [kNAG][hi 16K | lo 16K]
    uint8_t *ringW = ringAct + (size_t)kNAG * kStageA;          // [kNW][16K]
    BarsG *bars = reinterpret_cast<BarsG *>(ringW + (size_t)kNW * kPanelBytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long dbg_local[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const bool dbg_on = q.dbg != nullptr && lane == 0 && (warp == 0 || warp == 1 || warp == kGat0 || warp == kEpi0);
    const long long dbg_t0 = clock64();
#define DBG(i) (dbg_on ? &dbg_local[i] : nullptr)

    if (threadIdx.x == 0) {
        for (int o = 0; o < 2; ++o) {
            for (int i = 0; i < kNG; ++i) mbar_init(&bars->g_full[o][i], kGatWarps);
            for (int i = 0; i < kNW; ++i) mbar_init(&bars->w_full[o][i], 1);
            for (int i = 0; i < kNAG; ++i) mbar_init(&bars->act_full[o][i], 4);
        }
        for (int i = 0; i < kNG; ++i) mbar_init(&bars->g_empty[i], 1);
        for (int i = 0; i < kNW; ++i) mbar_init(&bars->w_empty[i], 1);
        for (int i = 0; i < kNAG; ++i) mbar_init(&bars->act_empty[i], 1);
        for (int i = 0; i < 4; ++i) { mbar_init(&bars->tm_full[i], 1); mbar_init(&bars->tm_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars->tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, bars->tmem_base, 0);
    const long long first_tile = blockIdx.x, tile_stride = gridDim.x;
    const unsigned owner_mask = gp.issuers == 1 ? 0u : 10u;       // heads whose steps the second issuer (warp 2) runs

    if (warp == 0) {
        // =============================== weight producer ===============================
        uint32_t u = 0;
        for (long long tile = first_tile; tile < q.total_tiles; tile += tile_stride) {
#pragma unroll 1
            for (int s = 0; s < kSteps; ++s) {
                const int layer = (int)((kSeq >> (4 * s + 2)) & 3), hd = (int)((kSeq >> (4 * s)) & 3);
                if (!((q.head_mask >> hd) & 1)) continue;
                const int n_panels = layer < 3 ? 4 : 1;
                for (int j = 0; j < n_panels; ++j, ++u) {
                    const int slot = u % kNW;
                    mbar_wait_t(&bars->w_empty[slot], ((u / kNW) & 1) ^ 1, DBG(0));
                    if (elect_one()) {
                        // the stream of query_tc_pack_weights: 48 layer-1 panels, then [L2 | L3][head][kb][hi | lo], then 4 small
                        const uint32_t bytes = layer < 3 ? kPanelBytes : kSmallPanelBytes;
                        const size_t off = layer < 3 ? (size_t)(kL1Blocks * 8 + (layer - 1) * 16 + hd * 4 + j) * kPanelBytes
                                                     : (size_t)kBigUnits * kPanelBytes + (size_t)hd * kSmallPanelBytes;
                        uint64_t *full = &bars->w_full[(owner_mask >> hd) & 1][slot];
                        mbar_arrive_expect_tx(full, bytes);
                        bulk_g2s(ringW + (size_t)slot * kPanelBytes, q.wstream + off, bytes, full);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1 || warp == 2) {
        // =============================== MMA issuers ===============================
        const int me = warp - 1;
        const unsigned my_heads = q.head_mask & (me == 0 ? ~owner_mask : owner_mask);
        uint32_t pw = 0, pg = 0, pa = 0;             // parity of this issuer's next use of every weight / gather / activation slot
        constexpr uint32_t idesc = make_idesc(kTileM, 128), idesc16 = make_idesc(kTileM, 16);
        const uint32_t ringG_lo = desc_lo(smem_u32(ringG)), ringAct_lo = desc_lo(smem_u32(ringAct)), ringW_lo = desc_lo(smem_u32(ringW));
        constexpr uint32_t kStageLo = kStageA >> 4, kPanelLo = kPanelBytes >> 4;
        uint32_t u = 0, gblk = 0, actblk = 0, tile_i = 0;
        for (long long tile = first_tile; my_heads != 0 && tile < q.total_tiles; tile += tile_stride, ++tile_i) {
#pragma unroll 1
            for (int s = 0; s < kSteps; ++s) {
                const int layer = (int)((kSeq >> (4 * s + 2)) & 3), h = (int)((kSeq >> (4 * s)) & 3);
                if (!((q.head_mask >> h) & 1)) continue;
                if (!((my_heads >> h) & 1)) {                 // the other issuer's step
                    u += layer < 3 ? 4 : 1;
                    if (layer == 1) gblk += 2; else actblk += 2;
                    continue;
                }
                const uint32_t d = tmem_base + h * 128;
                if (layer == 1) {
                    // layer 2 of head h: A operand = the two gathered blocks, consumed as they arrive
                    mbar_wait_t(&bars->tm_empty[h], (tile_i & 1) ^ 1, DBG(2));   // the previous tile's output of head h is read
                    tc_fence_after();
#pragma unroll 1
                    for (int kb = 0; kb < 2; ++kb, ++gblk) {
                        const int sg = gblk % kNG;
                        mbar_wait_t(&bars->g_full[me][sg], (pg >> sg) & 1, DBG(1));
                        pg ^= 1u << sg;
                        tc_fence_after();
                        const uint32_t a_hi = ringG_lo + sg * kStageLo, a_lo = a_hi + kPanelLo;
                        const int s0 = u % kNW, s1 = (u + 1) % kNW;
                        mbar_wait_t(&bars->w_full[me][s0], (pw >> s0) & 1, DBG(3));
                        pw ^= 1u << s0;
                        tc_fence_after();
                        const uint32_t w0 = ringW_lo + s0 * kPanelLo;
                        if (elect_one()) {
                            umma_burst_pair<4>(d, desc64(a_hi), desc64(a_lo), desc64(w0), idesc, kb != 0);
                            umma_commit(&bars->w_empty[s0]);
                        }
                        __syncwarp();
                        mbar_wait_t(&bars->w_full[me][s1], (pw >> s1) & 1, DBG(3));
                        pw ^= 1u << s1;
                        tc_fence_after();
                        const uint32_t w1 = ringW_lo + s1 * kPanelLo;
                        if (elect_one()) {
                            umma_burst_single<4>(d, desc64(a_hi), desc64(w1), idesc);
                            umma_commit(&bars->w_empty[s1]);
                            umma_commit(&bars->g_empty[sg]);
                            if (kb == 1) umma_commit(&bars->tm_full[h]);
                        }
                        __syncwarp();
                        u += 2;
                    }
                    continue;
                }
                // layers 3 and 4: both activation blocks must be complete before the accumulator is overwritten
                const uint32_t b0 = actblk, b1 = actblk + 1;
                mbar_wait_t(&bars->act_full[me][b0 % kNAG], (pa >> (b0 % kNAG)) & 1, DBG(4));
                mbar_wait_t(&bars->act_full[me][b1 % kNAG], (pa >> (b1 % kNAG)) & 1, DBG(4));
                pa ^= (1u << (b0 % kNAG)) | (1u << (b1 % kNAG));
                tc_fence_after();
                if (layer == 2) {
#pragma unroll 1
                    for (int kb = 0; kb < 2; ++kb, ++actblk) {
                        const int sa = actblk % kNAG;
                        const uint32_t a_hi = ringAct_lo + sa * kStageLo, a_lo = a_hi + kPanelLo;
                        const int s0 = u % kNW, s1 = (u + 1) % kNW;
                        mbar_wait_t(&bars->w_full[me][s0], (pw >> s0) & 1, DBG(5));
                        pw ^= 1u << s0;
                        tc_fence_after();
                        const uint32_t w0 = ringW_lo + s0 * kPanelLo;
                        if (elect_one()) {
                            umma_burst_pair<4>(d, desc64(a_hi), desc64(a_lo), desc64(w0), idesc, kb != 0);
                            umma_commit(&bars->w_empty[s0]);
                        }
                        __syncwarp();
                        mbar_wait_t(&bars->w_full[me][s1], (pw >> s1) & 1, DBG(5));
                        pw ^= 1u << s1;
                        tc_fence_after();
                        const uint32_t w1 = ringW_lo + s1 * kPanelLo;
                        if (elect_one()) {
                            umma_burst_single<4>(d, desc64(a_hi), desc64(w1), idesc);
                            umma_commit(&bars->w_empty[s1]);
                            umma_commit(&bars->act_empty[sa]);
                            if (kb == 1) umma_commit(&bars->tm_full[h]);
                        }
                        __syncwarp();
                        u += 2;
                    }
                } else {
                    // output layer (<= 14 outputs, padded to N = 16): one 8 KB panel [kb0 hi | kb0 lo | kb1 hi | kb1 lo]
                    const int s0 = u % kNW;
                    mbar_wait_t(&bars->w_full[me][s0], (pw >> s0) & 1, DBG(5));
                    pw ^= 1u << s0;
                    tc_fence_after();
                    const uint32_t w = ringW_lo + s0 * kPanelLo;
                    const int sa0 = actblk % kNAG, sa1 = (actblk + 1) % kNAG;
                    if (elect_one()) {
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb) {
                            const uint32_t a_hi = ringAct_lo + (kb == 0 ? sa0 : sa1) * kStageLo, a_lo = a_hi + kPanelLo;
                            const uint32_t w_hi = w + kb * (4096 >> 4), w_lo = w_hi + (2048 >> 4);
                            umma_burst_triple4(d, desc64(a_hi), desc64(a_lo), desc64(w_hi), desc64(w_lo), idesc16, kb != 0);
                            umma_commit(&bars->act_empty[kb == 0 ? sa0 : sa1]);
                        }
                        umma_commit(&bars->w_empty[s0]);
                        umma_commit(&bars->tm_full[h]);
                    }
                    __syncwarp();
                    u += 1;
                    actblk += 2;
                }
            }
        }
    } else if (warp >= kGat0) {
        // =============================== gather warps: layer 1 through the projected maps ===============================
        // warp g owns rows [16g, 16g + 16); a half-warp handles one point and one 64-unit block: 16 lanes x 4 hidden units
        const int g = warp - kGat0;
        const int half = lane >> 4, l16 = lane & 15;
        const int rowCf = q.fw * kGC, rowCs = 2 * q.fw * kGC;
        uint32_t gblk = 0;
        for (long long tile = first_tile; tile < q.total_tiles; tile += tile_stride) {
            const int b = q.grid_mode ? q.batch_index : (int)(tile / q.tiles_per_b);
            const long long n0 = (tile % q.tiles_per_b) * kTileM;
            const float ccx = __ldg(q.crop_center + b * 2), ccy = __ldg(q.crop_center + b * 2 + 1);
            // lane L (and L + 16) projects row 16g + (L & 15) once per tile; the block loop fetches it with shuffles
            float my_x = 0.f, my_y = 0.f, my_z = 1.f, my_nx, my_ny;
            if (n0 + g * 16 + l16 < q.n_count) load_point(q, b, n0 + g * 16 + l16, my_x, my_y, my_z);
            project_tc(my_x, my_y, my_z, ccx, ccy, my_nx, my_ny);
            const LaneTaps tapsF = make_lane_taps(my_nx, my_ny, q.fh, q.fw, kGC), tapsS = make_lane_taps(my_nx, my_ny, 2 * q.fh, 2 * q.fw, kGC);
            const float my_zr = __fsub_rn(my_z, 2.2f);                    // z_feat (model/chore.py:128-129)
#pragma unroll 1
            for (int h = 0; h < 4; ++h) {
                if (!((q.head_mask >> h) & 1)) continue;
#pragma unroll 1
                for (int kb = 0; kb < 2; ++kb, ++gblk) {
                    const int sg = gblk % kNG;
                    const int ch = h * 128 + kb * 64 + l16 * 4;
                    const float4 bias = __ldg(reinterpret_cast<const float4 *>(q.b1 + ch));
                    const float4 wzx = __ldg(reinterpret_cast<const float4 *>(gp.wz + ch));
                    const float4 wzy = __ldg(reinterpret_cast<const float4 *>(gp.wz + kGC + ch));
                    const float4 wzz = __ldg(reinterpret_cast<const float4 *>(gp.wz + 2 * kGC + ch));
                    const float *bf = gp.gf + ch, *bs = gp.gs + ch;
                    mbar_wait_t(&bars->g_empty[sg], ((gblk / kNG) & 1) ^ 1, DBG(6));
                    uint8_t *hi = ringG + (size_t)sg * kStageA, *lo = hi + kPanelBytes;
#pragma unroll 1
                    for (int it0 = 0; it0 < 8; it0 += kGBatch) {
                        float4 v[kGBatch][8];
                        float wxf[kGBatch], wyf[kGBatch], wxs[kGBatch], wys[kGBatch];
#pragma unroll
                        for (int j = 0; j < kGBatch; ++j) {
                            const int src = half * 8 + it0 + j;            // a half-warp walks 8 consecutive points: neighbours share taps
                            const int of = __shfl_sync(0xffffffffu, tapsF.off_valid, src), os = __shfl_sync(0xffffffffu, tapsS.off_valid, src);
                            wxf[j] = __shfl_sync(0xffffffffu, tapsF.wx, src); wyf[j] = __shfl_sync(0xffffffffu, tapsF.wy, src);
                            wxs[j] = __shfl_sync(0xffffffffu, tapsS.wx, src); wys[j] = __shfl_sync(0xffffffffu, tapsS.wy, src);
                            const float *pf = bf + (of & ~15), *ps = bs + (os & ~15);
                            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
                            v[j][0] = (of & 1) ? __ldg(reinterpret_cast<const float4 *>(pf)) : zero;
                            v[j][1] = (of & 2) ? __ldg(reinterpret_cast<const float4 *>(pf + kGC)) : zero;
                            v[j][2] = (of & 4) ? __ldg(reinterpret_cast<const float4 *>(pf + rowCf)) : zero;
                            v[j][3] = (of & 8) ? __ldg(reinterpret_cast<const float4 *>(pf + rowCf + kGC)) : zero;
                            v[j][4] = (os & 1) ? __ldg(reinterpret_cast<const float4 *>(ps)) : zero;
                            v[j][5] = (os & 2) ? __ldg(reinterpret_cast<const float4 *>(ps + kGC)) : zero;
                            v[j][6] = (os & 4) ? __ldg(reinterpret_cast<const float4 *>(ps + rowCs)) : zero;
                            v[j][7] = (os & 8) ? __ldg(reinterpret_cast<const float4 *>(ps + rowCs + kGC)) : zero;
                        }
#pragma unroll
                        for (int j = 0; j < kGBatch; ++j) {
                            const int src = half * 8 + it0 + j;            // a half-warp walks 8 consecutive points: neighbours share taps
                            const int r = g * 16 + src;
                            const float x = __shfl_sync(0xffffffffu, my_x, src), y = __shfl_sync(0xffffffffu, my_y, src);
                            const float zr = __shfl_sync(0xffffffffu, my_zr, src);
                            const float uxf = 1.f - wxf[j], uyf = 1.f - wyf[j], uxs = 1.f - wxs[j], uys = 1.f - wys[j];
                            const float w[8] = {uyf * uxf, uyf * wxf[j], wyf[j] * uxf, wyf[j] * wxf[j],
                                                uys * uxs, uys * wxs[j], wys[j] * uxs, wys[j] * wxs[j]};
                            float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;       // sampled maps first (largest terms), then xyz and bias
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                v0 = fmaf(v[j][k].x, w[k], v0); v1 = fmaf(v[j][k].y, w[k], v1);
                                v2 = fmaf(v[j][k].z, w[k], v2); v3 = fmaf(v[j][k].w, w[k], v3);
                            }
                            v0 = fmaf(x, wzx.x, v0); v1 = fmaf(x, wzx.y, v1); v2 = fmaf(x, wzx.z, v2); v3 = fmaf(x, wzx.w, v3);
                            v0 = fmaf(y, wzy.x, v0); v1 = fmaf(y, wzy.y, v1); v2 = fmaf(y, wzy.z, v2); v3 = fmaf(y, wzy.w, v3);
                            v0 = fmaf(zr, wzz.x, v0); v1 = fmaf(zr, wzz.y, v1); v2 = fmaf(zr, wzz.z, v2); v3 = fmaf(zr, wzz.w, v3);
                            v0 = fmaxf(v0 + bias.x, 0.f); v1 = fmaxf(v1 + bias.y, 0.f);
                            v2 = fmaxf(v2 + bias.z, 0.f); v3 = fmaxf(v3 + bias.w, 0.f);
                            uint32_t h01, l01, h23, l23;
                            split2_pos(v0, v1, h01, l01);
                            split2_pos(v2, v3, h23, l23);
                            const uint32_t off = sw128(r, l16 >> 1) + (l16 & 1) * 8;
                            *reinterpret_cast<uint2 *>(hi + off) = make_uint2(h01, h23);
                            *reinterpret_cast<uint2 *>(lo + off) = make_uint2(l01, l23);
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->g_full[(owner_mask >> h) & 1][sg]);
                }
            }
        }
    } else {
        // =============================== epilogue warps ===============================
        const int e = warp - kEpi0;                  // 0..7
        const int quarter = warp & 3;                // TMEM lane quarter this warp may access
        const int colhalf = e >> 2;                  // 0: columns 0-63, 1: columns 64-127
        const int row = quarter * 32 + lane;         // point of the tile owned by this thread
        uint32_t actblk = 0, tile_i = 0;
        for (long long tile = first_tile; tile < q.total_tiles; tile += tile_stride, ++tile_i) {
            const int b = q.grid_mode ? q.batch_index : (int)(tile / q.tiles_per_b);
            const long long n = (tile % q.tiles_per_b) * kTileM + row;
            const bool live = n < q.n_count;
            bool inimg = false;
            {
                float x = 0.f, y = 0.f, z = 1.f, nx, ny;
                if (live) load_point(q, b, n, x, y, z);
                project_tc(x, y, z, __ldg(q.crop_center + b * 2), __ldg(q.crop_center + b * 2 + 1), nx, ny);
                inimg = (nx >= -1.0f) && (nx <= 1.0f) && (ny >= -1.0f) && (ny <= 1.0f);
                if (q.in_img && live && e < 4) q.in_img[(size_t)b * q.N + q.n_start + n] = inimg;
            }
#pragma unroll 1
            for (int s = 0; s < kSteps; ++s) {
                const int layer = (int)((kSeq >> (4 * s + 2)) & 3), h = (int)((kSeq >> (4 * s)) & 3);
                if (!((q.head_mask >> h) & 1)) continue;
                mbar_wait_t(&bars->tm_full[h], (tile_i * 3 + (uint32_t)(layer - 1)) & 1, DBG(7));   // 3 completions per tile and head
                if (layer == 3 && (h & 1) != colhalf) continue;      // output layer: the heads are split between the groups
                tc_fence_after();
                if (layer < 3) {
                    // bias + ReLU + hi/lo split -> activation k-block `colhalf` of head h
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + h * 128 + colhalf * 64;
                    const float *bias = (layer == 1 ? q.b2 : q.b3) + h * 128 + colhalf * 64;
                    const uint32_t blk = actblk + colhalf;
                    const int sa = blk % kNAG;
                    mbar_wait_t(&bars->act_empty[sa], ((blk / kNAG) & 1) ^ 1, DBG(8));
                    uint8_t *hi = ringAct + (size_t)sa * kStageA, *lo = hi + kPanelBytes;
#pragma unroll
                    for (int part = 0; part < 2; ++part) {
                        uint32_t v[32];
                        tmem_ld32(taddr + part * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int c8 = 0; c8 < 4; ++c8) {     // 8 columns = one 16-byte chunk of fp16
                            uint32_t hh[4], ll[4];
                            const float4 bA = __ldg(reinterpret_cast<const float4 *>(bias + part * 32 + c8 * 8));
                            const float4 bB = __ldg(reinterpret_cast<const float4 *>(bias + part * 32 + c8 * 8 + 4));
                            const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int c = c8 * 8 + j * 2;
                                const float a0 = fmaxf(__uint_as_float(v[c]) + bb[j * 2], 0.f);
                                const float a1 = fmaxf(__uint_as_float(v[c + 1]) + bb[j * 2 + 1], 0.f);
                                split2_pos(a0, a1, hh[j], ll[j]);
                            }
                            const uint32_t off = sw128(row, part * 4 + c8);
                            *reinterpret_cast<uint4 *>(hi + off) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                            *reinterpret_cast<uint4 *>(lo + off) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
                        }
                    }
                    tc_fence_before();
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->act_full[(owner_mask >> h) & 1][sa]);
                    actblk += 2;
                } else {
                    // output layer: 16 accumulator columns -> + bias -> OUT_DIST mask -> HBM (reference layout)
                    const int nout = head_out_tc(h);
                    uint32_t v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + h * 128, v);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->tm_empty[h]);   // the next tile's layer 2 may overwrite head h
                    if (live) {
                        float *outp = q.out[h] + ((size_t)b * nout) * q.N + q.n_start + n;
#pragma unroll
                        for (int o = 0; o < 14; ++o) {
                            if (o < nout) {
                                float val = __uint_as_float(v[o]) + __ldg(q.b4 + h * 16 + o);
                                if (h == 0 && !inimg) val = 5.0f;          // model/chore.py:147-150
                                outp[(size_t)o * q.N] = val;
                            }
                        }
                    }
                }
            }
        }
    }

    if (dbg_on) {
        dbg_local[9] = (unsigned long long)(clock64() - dbg_t0);
        for (int i = 0; i < 10; ++i)
            if (dbg_local[i] && (i != 9 || warp == 1)) q.dbg[(size_t)blockIdx.x * 16 + i] = dbg_local[i];
    }
#undef DBG
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

}   // namespace

// layer-1 weights in the layouts of this path: four 1x1-convolution weight sets for conv_hx (output halves = head pairs)
// and the xyz columns as [3][512]
int query_g_pack_weights(chore_handle *h, const std::vector<float> &w1 /*[4][128][323]*/) {
    QueryGWeights &g = h->qg;
    std::vector<float> wf((size_t)256 * 256), ws((size_t)256 * 64), wz((size_t)3 * kGC);
    for (int half = 0; half < 2; ++half) {
        for (int n = 0; n < 256; ++n) {
            const float *row = w1.data() + (size_t)(half * 256 + n) * kPointC;
            for (int c = 0; c < 256; ++c) wf[(size_t)n * 256 + c] = row[c];
            for (int c = 0; c < 64; ++c) ws[(size_t)n * 64 + c] = row[259 + c];
            for (int j = 0; j < 3; ++j) wz[(size_t)j * kGC + half * 256 + n] = row[256 + j];
        }
        if (int rc = conv_hx_pack_weights(h, wf.data(), 256, 256, 1, 1, &g.wf[half])) return rc;
        if (int rc = conv_hx_pack_weights(h, ws.data(), 256, 64, 1, 1, &g.ws[half])) return rc;
    }
    if (int rc = chore_dev_alloc(h, reinterpret_cast<void **>(&g.wz), wz.size() * sizeof(float))) return rc;
    CHORE_CUDA(cudaMemcpy(g.wz, wz.data(), wz.size() * sizeof(float), cudaMemcpyHostToDevice));
    return CHORE_OK;
}

// Opt-in (CHORE_B200_QUERY_PRE=1).  Measured on B200 at 4.19 M grid points: results within 1e-6 (relative) of the per-point
// kernel at the SAME speed (8.25 ms against 8.27 ms, both under the trace build), i.e. halving the MMAs did not pay yet:
// the gather warps are busy 45 k cycles per 128-point tile (8 taps x 512 B per (point, head) = 16 KB per point through
// the LSU against 5 KB before; only 28 KB of L1 remain beside 227 KB of shared memory, so the taps that neighbouring
// points share mostly come from L2 again -- with two points in flight per half-warp the footprint doubles and the tile
// takes 101 k cycles).  Next step: keep the previous point's taps in registers (consecutive grid points move < 1 pixel).
int query_g_mode() {
    const char *e = getenv("CHORE_B200_QUERY_PRE");      // read per call: the parity test toggles it
    return (e != nullptr && e[0] == '1') ? 1 : 0;
}

// G_f / G_s of image b: two N = 256 1x1 convolutions per map (head pairs), skipped for head pairs outside the mask
static int project_maps(chore_handle *h, const float *feat, const float *skip, int fh, int fw, int b, unsigned head_mask, cudaStream_t st) {
    QueryGWeights &g = h->qg;
    const size_t px = (size_t)fh * fw;
    if (g.map_px < px) {
        // (the old maps, if any, stay owned by the handle until chore_destroy: a query in flight may still read them)
        g.gf = g.gs = nullptr;
        if (int rc = chore_dev_alloc(h, reinterpret_cast<void **>(&g.gf), px * kGC * sizeof(float))) return rc;
        if (int rc = chore_dev_alloc(h, reinterpret_cast<void **>(&g.gs), 4 * px * kGC * sizeof(float))) return rc;
        g.map_px = px;
    }
    for (int map = 0; map < 2; ++map)
        for (int half = 0; half < 2; ++half) {
            if (!((head_mask >> (2 * half)) & 3u)) continue;
            ConvHxArgs a{};
            a.in = map == 0 ? feat + (size_t)b * px * kFeatC : skip + (size_t)b * 4 * px * kSkipC;
            a.ld_in = a.Cin = map == 0 ? kFeatC : kSkipC;
            a.B = 1; a.H = map == 0 ? fh : 2 * fh; a.W = map == 0 ? fw : 2 * fw; a.KS = 1; a.N = 256;
            a.w = map == 0 ? g.wf[half] : g.ws[half];
            a.out = map == 0 ? g.gf : g.gs; a.ld_out = kGC; a.off_out = half * 256;
            ConvHxPlan pl{};
            if (int rc = conv_hx_plan(h, a, &pl)) return rc;
            if (int rc = conv_hx_launch(h, a, pl, nullptr, nullptr, st)) return rc;
        }
    return CHORE_OK;
}

int query_g_launch(chore_handle *h, const float *feat, const float *skip, int fh, int fw, const float *crop_center, long long N,
                   long long n_start, long long n_count, int batch_index, const int *res, const double *step, const double *bmin,
                   unsigned head_mask, float *const outs[4], cudaStream_t st) {
    CHORE_CHECK(h->qg.wz != nullptr, "query_g: MLP weights not loaded");
    if (int rc = project_maps(h, feat, skip, fh, fw, batch_index, head_mask, st)) return rc;
    if (const char *dbg_e = getenv("CHORE_B200_QUERY_PRE_DEBUG"))
        if (strcmp(dbg_e, "maps") == 0) return CHORE_OK;         // debugging aid: stop after the projected maps
    GParams gp{};
    TcParams &q = gp.q;
    q.feat = feat; q.skip = skip; q.fh = fh; q.fw = fw;
    q.points = nullptr; q.crop_center = crop_center;
    q.B = 1; q.N = N; q.n_start = n_start; q.n_count = n_count;
    q.grid_mode = 1; q.batch_index = batch_index;
    q.ry = res[1]; q.rz = res[2];
    for (int i = 0; i < 3; ++i) { q.step[i] = step[i]; q.bmin[i] = bmin[i]; }
    q.head_mask = head_mask;
    for (int i = 0; i < 4; ++i) q.out[i] = outs[i];
    q.in_img = nullptr;
    const MlpWeights &m = h->mlp;
    q.wstream = m.wstream; q.b1 = m.b1; q.b2 = m.b2; q.b3 = m.b3; q.w4 = m.w4; q.b4 = m.b4;
    q.tiles_per_b = (n_count + kTileM - 1) / kTileM;
    q.total_tiles = q.tiles_per_b;
    gp.gf = h->qg.gf; gp.gs = h->qg.gs; gp.wz = h->qg.wz;
    const char *iss = getenv("CHORE_B200_QUERY_PRE_ISSUERS");
    gp.issuers = (iss != nullptr && iss[0] == '1') ? 1 : 2;
    CHORE_ONCE_PER_DEVICE(cudaFuncSetAttribute(query_g_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemG));
    const long long grid = q.total_tiles < h->sm_count ? q.total_tiles : h->sm_count;
    static const bool trace = getenv("CHORE_B200_TC_TRACE") != nullptr;
    unsigned long long *dbg = nullptr;
    if (trace) {
        CHORE_CUDA(cudaMalloc(&dbg, (size_t)grid * 16 * sizeof(unsigned long long)));
        CHORE_CUDA(cudaMemsetAsync(dbg, 0, (size_t)grid * 16 * sizeof(unsigned long long), st));
        q.dbg = dbg;
    }
    CHORE_LAUNCH(query_g_kernel, (unsigned)grid, kThreadsG, kSmemG, st, gp);
    if (trace) {   // debugging aid: synchronous, prints the mean wait cycles per role
        std::vector<unsigned long long> hbuf((size_t)grid * 16);
        CHORE_CUDA(cudaMemcpyAsync(hbuf.data(), dbg, hbuf.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CHORE_CUDA(cudaStreamSynchronize(st));
        CHORE_CUDA(cudaFree(dbg));
        static const char *names[10] = {"producer:w_empty", "mma:g_full", "mma:tm_empty", "mma:w_full(L2)", "mma:act_full",
                                        "mma:w_full(L3/4)", "gather:g_empty", "epi:tm_full", "epi:act_empty", "total"};
        double sum[10] = {0};
        for (long long c = 0; c < grid; ++c)
            for (int i = 0; i < 10; ++i) sum[i] += (double)hbuf[(size_t)c * 16 + i];
        const double tiles = (double)q.total_tiles / (double)grid;
        fprintf(stderr, "[g-trace] tiles/CTA %.1f; cycles per tile:", tiles);
        for (int i = 0; i < 10; ++i) fprintf(stderr, " %s=%.0f", names[i], sum[i] / (double)grid / tiles);
        fprintf(stderr, "\n");
    }
    return CHORE_OK;
}
