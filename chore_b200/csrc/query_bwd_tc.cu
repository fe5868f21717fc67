// Gradient of the point query with respect to the points on the tensor cores (split from query_tc.cu, which keeps the forward
// kernel and the weight packing): query_bwd_tc_kernel (forward recompute + transposed layers, work item = tile x head),
// query_bwd_geom_kernel (adjoint of the bilinear gathers and of the projection) and their launch.
#include "query_tc_shared.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

// =============================================================================================
// backward to the points on the tensor cores: one launch per head with a non-zero upstream gradient
//   F1..F3  forward recompute of the three 128-wide layers (ReLU masks stay in registers)
//   gH3 = W4^T g (CUDA cores, <= 14 terms), B3, B2: gH_{l-1} = (gH_l W_l) . mask_{l-1}   (tcgen05)
//   B1      gX (+)= gH1 W1: 384 columns in three accumulator regions -> (B*N, 384) in HBM
// then query_bwd_geom_kernel turns gX into d/d(point) (bilinear + projection adjoint).
// =============================================================================================
constexpr int kBwdNW = 6;
constexpr size_t kBwdSmemBytes = 1024 + 2 * (size_t)kStageA + 2 * (size_t)kStageA + (size_t)kBwdNW * kPanelBytes + 512;

struct BarsB {
    uint64_t a_full[2], a_empty[2];
    uint64_t w_full[kBwdNW], w_empty[kBwdNW];
    uint64_t act_full[2], act_empty[2];
    uint64_t tm_full, tm_empty;
    uint32_t tmem_base;
};

// general (signed) split of 32 accumulator values -> 4 chunks of a 64-column activation block
__device__ __forceinline__ void store_act32(const float (&a)[32], uint8_t *hi, uint8_t *lo, int row, int part) {
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
        uint32_t hh[4], ll[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split2(a[c8 * 8 + j * 2], a[c8 * 8 + j * 2 + 1], hh[j], ll[j]);
        const uint32_t off = sw128(row, part * 4 + c8);
        *reinterpret_cast<uint4 *>(hi + off) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
        *reinterpret_cast<uint4 *>(lo + off) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
    }
}

__global__ void __launch_bounds__(kThreads, 1) query_bwd_tc_kernel(const TcParams q) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *ringA = smem;                                   // [2][hi 16K | lo 16K]
    uint8_t *ringAct = ringA + 2 * (size_t)kStageA;          // [2][hi 16K | lo 16K]: stage = column half
    uint8_t *ringW = ringAct + 2 * (size_t)kStageA;          // [6][16K]
    BarsB *bars = reinterpret_cast<BarsB *>(ringW + (size_t)kBwdNW * kPanelBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nslots = q.bwd_nslots;
    const long long total_items = q.total_tiles * nslots;      // item = tile * nslots + slot

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->a_full[i], 12); mbar_init(&bars->a_empty[i], 1);      // 4 gather + 8 (otherwise idle) epilogue warps
            mbar_init(&bars->act_full[i], 4); mbar_init(&bars->act_empty[i], 1);
        }
        for (int i = 0; i < kBwdNW; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], 1); }
        mbar_init(&bars->tm_full, 1); mbar_init(&bars->tm_empty, 8);
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

    if (warp == 0) {
        // ---------------- weight producer ----------------
        uint32_t u = 0;
        for (long long item = first_tile; item < total_items; item += tile_stride) {
            const unsigned char *wsrc = q.wstream_bwd + (size_t)q.bwd_heads[item % nslots] * kBwdUnits * kPanelBytes;
            for (int i = 0; i < kBwdUnits; ++i, ++u) {
                const int s = u % kBwdNW;
                mbar_wait(&bars->w_empty[s], ((u / kBwdNW) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&bars->w_full[s], kPanelBytes);
                    bulk_g2s(ringW + (size_t)s * kPanelBytes, wsrc + (size_t)i * kPanelBytes, kPanelBytes, &bars->w_full[s]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        constexpr uint32_t idesc = make_idesc(kTileM, 128);
        const uint32_t ringA_lo = desc_lo(smem_u32(ringA)), ringAct_lo = desc_lo(smem_u32(ringAct)), ringW_lo = desc_lo(smem_u32(ringW));
        constexpr uint32_t kStageLo = kStageA >> 4, kPanelLo = kPanelBytes >> 4;
        uint32_t u = 0, ablk = 0, pass = 0, tile_i = 0;
        // one (A k-block) x (hi, lo weight panel pair): 3 MMAs per k-step into accumulator d
        auto mma_block = [&](uint32_t d, uint32_t a_hi, uint32_t a_lo, int ksteps, bool first) {
            const int s0 = u % kBwdNW, s1 = (u + 1) % kBwdNW;
            mbar_wait(&bars->w_full[s0], (u / kBwdNW) & 1);
            const uint32_t w0 = ringW_lo + s0 * kPanelLo;
            if (elect_one()) {
                if (ksteps == 4) umma_burst_pair<4>(d, desc64(a_hi), desc64(a_lo), desc64(w0), idesc, !first);
                else umma_burst_pair<1>(d, desc64(a_hi), desc64(a_lo), desc64(w0), idesc, !first);
                umma_commit(&bars->w_empty[s0]);
            }
            __syncwarp();
            mbar_wait(&bars->w_full[s1], ((u + 1) / kBwdNW) & 1);
            const uint32_t w1 = ringW_lo + s1 * kPanelLo;
            if (elect_one()) {
                if (ksteps == 4) umma_burst_single<4>(d, desc64(a_hi), desc64(w1), idesc);
                else umma_burst_single<1>(d, desc64(a_hi), desc64(w1), idesc);
                umma_commit(&bars->w_empty[s1]);
            }
            __syncwarp();
            u += 2;
        };
        for (long long item = first_tile; item < total_items; item += tile_stride, ++tile_i) {
            // F1: layer 1 of this head from the gathered feature blocks
            for (int kb = 0; kb < kL1Blocks; ++kb, ++ablk) {
                const int sa = ablk % 2;
                mbar_wait(&bars->a_full[sa], (ablk / 2) & 1);
                tc_fence_after();
                const uint32_t a_hi = ringA_lo + sa * kStageLo;
                mma_block(tmem_base, a_hi, a_hi + kPanelLo, kb == kL1Blocks - 1 ? 1 : 4, kb == 0);
                if (elect_one()) {
                    umma_commit(&bars->a_empty[sa]);
                    if (kb == kL1Blocks - 1) umma_commit(&bars->tm_full);
                }
                __syncwarp();
            }
            // F2, F3, B3, B2: 128 -> 128 from the activation blocks (stage = k-block)
            for (int ph = 0; ph < 4; ++ph, ++pass) {
                mbar_wait(&bars->act_full[0], pass & 1);
                mbar_wait(&bars->act_full[1], pass & 1);
                tc_fence_after();
                for (int kb = 0; kb < 2; ++kb) {
                    const uint32_t a_hi = ringAct_lo + kb * kStageLo;
                    mma_block(tmem_base, a_hi, a_hi + kPanelLo, 4, kb == 0);
                    if (elect_one()) {
                        umma_commit(&bars->act_empty[kb]);
                        if (kb == 1) umma_commit(&bars->tm_full);
                    }
                    __syncwarp();
                }
            }
            // B1: gX = gH1 W1, three 128-column chunks into accumulator regions 1..3
            mbar_wait(&bars->act_full[0], pass & 1);
            mbar_wait(&bars->act_full[1], pass & 1);
            mbar_wait(&bars->tm_empty, (tile_i & 1) ^ 1);        // regions 1..3 drained by the previous tile's store
            tc_fence_after();
            for (int nc = 0; nc < 3; ++nc)
                for (int kb = 0; kb < 2; ++kb) {
                    const uint32_t a_hi = ringAct_lo + kb * kStageLo;
                    mma_block(tmem_base + (1 + nc) * 128, a_hi, a_hi + kPanelLo, 4, kb == 0);
                    if (nc == 2) {
                        if (elect_one()) {
                            umma_commit(&bars->act_empty[kb]);
                            if (kb == 1) umma_commit(&bars->tm_full);
                        }
                        __syncwarp();
                    }
                }
            ++pass;
        }
    } else if (warp < kEpiWarp0) {
        // ---------------- gather warps: rows [0, 64) of every feature k-block, 16 rows each (the epilogue warps, idle until
        // layer 1 is complete, gather rows [64, 128): the recompute of layer 1 for ONE head needs the whole 6-block operand, so
        // the gather is what bounds a work item) ----------------
        const int g = warp - kGatherWarp0;
        const int half = lane >> 4, l16 = lane & 15;
        uint32_t ablk = 0;
        for (long long item = first_tile; item < total_items; item += tile_stride) {
            const long long tile = item / nslots;
            const int b = (int)(tile / q.tiles_per_b);
            const long long n0 = (tile % q.tiles_per_b) * kTileM;
            const float ccx = __ldg(q.crop_center + b * 2), ccy = __ldg(q.crop_center + b * 2 + 1);
            const float *F = q.feat + (size_t)b * q.fh * q.fw * kFeatC;
            const float *S = q.skip + (size_t)b * (2 * q.fh) * (2 * q.fw) * kSkipC;
            float my_x = 0.f, my_y = 0.f, my_z = 1.f, my_nx, my_ny;
            if (n0 + g * 16 + l16 < q.n_count) load_point(q, b, n0 + g * 16 + l16, my_x, my_y, my_z);      // lanes L and L + 16: row 16 g + L
            project_tc(my_x, my_y, my_z, ccx, ccy, my_nx, my_ny);
            const LaneTaps tapsF = make_lane_taps(my_nx, my_ny, q.fh, q.fw, kFeatC), tapsS = make_lane_taps(my_nx, my_ny, 2 * q.fh, 2 * q.fw, kSkipC);
            for (int kb = 0; kb < kL1Blocks; ++kb, ++ablk) {
                const int sa = ablk % 2;
                mbar_wait(&bars->a_empty[sa], ((ablk / 2) & 1) ^ 1);
                uint8_t *hi = ringA + (size_t)sa * kStageA, *lo = hi + kPanelBytes;
                if (kb < 5) {
                    const bool is_feat = kb < 4;
                    gather_kblock_p<16, kGatherBatch>((is_feat ? F + kb * 64 : S) + l16 * 4, (is_feat ? q.fw * kFeatC : 2 * q.fw * kSkipC),
                                                      is_feat ? kFeatC : kSkipC, is_feat ? tapsF : tapsS, g, half, l16, hi, lo);
                } else if (lane < 16) {
                    const int r = g * 16 + lane;
                    uint32_t h01, l01, h23, l23;
                    split2(my_x, my_y, h01, l01);
                    split2(__fsub_rn(my_z, 2.2f), 0.f, h23, l23);
                    *reinterpret_cast<uint4 *>(hi + sw128(r, 0)) = make_uint4(h01, h23, 0u, 0u);
                    *reinterpret_cast<uint4 *>(lo + sw128(r, 0)) = make_uint4(l01, l23, 0u, 0u);
                    *reinterpret_cast<uint4 *>(hi + sw128(r, 1)) = make_uint4(0u, 0u, 0u, 0u);
                    *reinterpret_cast<uint4 *>(lo + sw128(r, 1)) = make_uint4(0u, 0u, 0u, 0u);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->a_full[sa]);
            }
        }
    } else {
        // ---------------- epilogue warps ----------------
        const int e = warp - kEpiWarp0;
        const int quarter = warp & 3, colhalf = e >> 2;
        const int row = quarter * 32 + lane;
        uint8_t *hi = ringAct + (size_t)colhalf * kStageA, *lo = hi + kPanelBytes;     // this warp's activation block
        const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + colhalf * 64;
        uint32_t pass = 0, ablk = 0;
        // CHORE_B200_TC_TRACE: cycles per phase of a work item as seen by epilogue warp 0 (14 counters per CTA)
        const bool bt_on = q.dbg != nullptr && e == 0 && lane == 0;
        __shared__ unsigned long long bt[14];          // (shared memory: the counters must not cost registers on the hot path)
        if (bt_on)
            for (int i = 0; i < 14; ++i) bt[i] = 0;
        long long bt_last = clock64();
#define BT(i) do { if (bt_on) { const long long now_ = clock64(); bt[i] += (unsigned long long)(now_ - bt_last); bt_last = now_; } } while (0)
        for (long long item = first_tile; item < total_items; item += tile_stride) {
            const long long tile = item / nslots;
            const int slot = (int)(item % nslots);
            const int hd = q.bwd_heads[slot];
            const int nout = head_out_tc(hd);
            const float *g_head = q.g_heads[slot];
            const int b = (int)(tile / q.tiles_per_b);
            const long long n = (tile % q.tiles_per_b) * kTileM + row;
            const bool live = n < q.n_count;
            {
                // ---- rows [64 + 8 e, 64 + 8 e + 8) of the six layer-1 operand blocks ----
                const long long n0 = (tile % q.tiles_per_b) * kTileM;
                const int half = lane >> 4, l16 = lane & 15, l8 = lane & 7;
                const float ccx = __ldg(q.crop_center + b * 2), ccy = __ldg(q.crop_center + b * 2 + 1);
                const float *F = q.feat + (size_t)b * q.fh * q.fw * kFeatC;
                const float *S = q.skip + (size_t)b * (2 * q.fh) * (2 * q.fw) * kSkipC;
                float hx = 0.f, hy = 0.f, hz = 1.f, hnx, hny;
                if (n0 + 64 + e * 8 + l8 < q.n_count) load_point(q, b, n0 + 64 + e * 8 + l8, hx, hy, hz);      // lane L: row 64 + 8 e + (L & 7)
                project_tc(hx, hy, hz, ccx, ccy, hnx, hny);
                const LaneTaps tapsF = make_lane_taps(hnx, hny, q.fh, q.fw, kFeatC), tapsS = make_lane_taps(hnx, hny, 2 * q.fh, 2 * q.fw, kSkipC);
                for (int kb = 0; kb < kL1Blocks; ++kb, ++ablk) {
                    const int sa = ablk % 2;
                    mbar_wait(&bars->a_empty[sa], ((ablk / 2) & 1) ^ 1);
                    uint8_t *ahi = ringA + (size_t)sa * kStageA, *alo = ahi + kPanelBytes;
                    if (kb < 5) {
                        const bool is_feat = kb < 4;
                        gather_kblock_p<8, kGatherBatch>((is_feat ? F + kb * 64 : S) + l16 * 4, (is_feat ? q.fw * kFeatC : 2 * q.fw * kSkipC),
                                                         is_feat ? kFeatC : kSkipC, is_feat ? tapsF : tapsS, 8 + e, half, l16, ahi, alo);
                    } else if (lane < 8) {
                        const int r = 64 + e * 8 + lane;
                        uint32_t h01, l01, h23, l23;
                        split2(hx, hy, h01, l01);
                        split2(__fsub_rn(hz, 2.2f), 0.f, h23, l23);
                        *reinterpret_cast<uint4 *>(ahi + sw128(r, 0)) = make_uint4(h01, h23, 0u, 0u);
                        *reinterpret_cast<uint4 *>(alo + sw128(r, 0)) = make_uint4(l01, l23, 0u, 0u);
                        *reinterpret_cast<uint4 *>(ahi + sw128(r, 1)) = make_uint4(0u, 0u, 0u, 0u);
                        *reinterpret_cast<uint4 *>(alo + sw128(r, 1)) = make_uint4(0u, 0u, 0u, 0u);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->a_full[sa]);
                }
            }
            // upstream gradient of this row (zero for dead rows and for df outside the image); loaded after the gather so that its 14
            // registers are not live across it
            float gout[14];
            {
                float x = 0.f, y = 0.f, z = 1.f, nx, ny;
                if (live) load_point(q, b, n, x, y, z);
                project_tc(x, y, z, __ldg(q.crop_center + b * 2), __ldg(q.crop_center + b * 2 + 1), nx, ny);
                const bool inimg = (nx >= -1.0f) && (nx <= 1.0f) && (ny >= -1.0f) && (ny <= 1.0f);
                const bool use = live && !(hd == 0 && !inimg);
#pragma unroll
                for (int o = 0; o < 14; ++o) gout[o] = (use && o < nout) ? __ldg(g_head + ((size_t)b * nout + o) * q.N + q.n_start + n) : 0.f;
            }
            BT(0);
            uint32_t mask[3][2];
            // ---- forward recompute: F1, F2, F3 ----
#pragma unroll 1
            for (int layer = 0; layer < 3; ++layer) {
                mbar_wait(&bars->tm_full, layer & 1);          // 6 completions per tile: parity = phase & 1
                BT(1 + layer);
                tc_fence_after();
                const float *bias = (layer == 0 ? q.b1 : (layer == 1 ? q.b2 : q.b3)) + hd * 128 + colhalf * 64;
                mbar_wait(&bars->act_empty[colhalf], (pass & 1) ^ 1);
#pragma unroll
                for (int part = 0; part < 2; ++part) {
                    uint32_t v[32];
                    tmem_ld32(tbase + part * 32, v);
                    tmem_ld_wait();
                    float a[32];
                    uint32_t m = 0;
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias + part * 32) + c4);
                        const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int c = c4 * 4 + j;
                            const float z = __uint_as_float(v[c]) + bv[j];
                            m |= (z > 0.f ? 1u : 0u) << c;
                            a[c] = fmaxf(z, 0.f);
                        }
                    }
                    mask[layer][part] = m;
                    if (layer < 2) {
                        store_act32(a, hi, lo, row, part);
                    } else {
                        // gH3 = (W4^T g) . mask3 for my 32 columns
                        const float *w4 = q.w4 + ((size_t)hd * 16) * kHidden + colhalf * 64 + part * 32;
#pragma unroll
                        for (int c = 0; c < 32; ++c) a[c] = 0.f;
#pragma unroll
                        for (int o = 0; o < 14; ++o) {
                            if (o < nout) {
#pragma unroll
                                for (int c4 = 0; c4 < 8; ++c4) {
                                    const float4 w = __ldg(reinterpret_cast<const float4 *>(w4 + (size_t)o * kHidden) + c4);
                                    a[c4 * 4] = fmaf(w.x, gout[o], a[c4 * 4]); a[c4 * 4 + 1] = fmaf(w.y, gout[o], a[c4 * 4 + 1]);
                                    a[c4 * 4 + 2] = fmaf(w.z, gout[o], a[c4 * 4 + 2]); a[c4 * 4 + 3] = fmaf(w.w, gout[o], a[c4 * 4 + 3]);
                                }
                            }
                        }
#pragma unroll
                        for (int c = 0; c < 32; ++c) a[c] = ((m >> c) & 1u) ? a[c] : 0.f;
                        store_act32(a, hi, lo, row, part);
                    }
                }
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->act_full[colhalf]);
                ++pass;
                BT(4 + layer);
            }
            // ---- backward chain: B3 -> (. mask2), B2 -> (. mask1) ----
#pragma unroll 1
            for (int step = 0; step < 2; ++step) {
                mbar_wait(&bars->tm_full, (3 + step) & 1);
                BT(7 + step);
                tc_fence_after();
                mbar_wait(&bars->act_empty[colhalf], (pass & 1) ^ 1);
#pragma unroll
                for (int part = 0; part < 2; ++part) {
                    uint32_t v[32];
                    tmem_ld32(tbase + part * 32, v);
                    tmem_ld_wait();
                    const uint32_t m = mask[1 - step][part];
                    float a[32];
#pragma unroll
                    for (int c = 0; c < 32; ++c) a[c] = ((m >> c) & 1u) ? __uint_as_float(v[c]) : 0.f;
                    store_act32(a, hi, lo, row, part);
                }
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->act_full[colhalf]);
                ++pass;
                BT(9 + step);
            }
            // ---- B1 result: accumulator regions 1..3 -> gX[point][384] ----
            mbar_wait(&bars->tm_full, 5 & 1);
            BT(11);
            tc_fence_after();
            // TMEM hands every lane one point's row; a warp store in that shape touches 32 different 128-byte lines (measured:
            // 21 k cycles for the 384 columns).  Every 32 x 32 chunk is transposed through a padded per-warp staging tile in the
            // activation ring (idle: all MMAs that read it have completed) so that a store instruction covers 4 full lines.
            uint8_t *stg = ringAct + (size_t)e * (32 * 144);
            const int rsub = lane >> 3, cq = (lane & 7) * 4;
            const long long n_warp = (tile % q.tiles_per_b) * kTileM + quarter * 32;        // first point of this warp's rows
            float *gxw = q.gX + (size_t)slot * q.gx_slot_stride + ((size_t)b * q.N + q.n_start + n_warp) * kGXLd + colhalf * 64 + cq;
#pragma unroll 1
            for (int reg = 0; reg < 3; ++reg) {
#pragma unroll
                for (int part = 0; part < 2; ++part) {
                    uint32_t v[32];
                    tmem_ld32(tbase + (1 + reg) * 128 + part * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4 *>(stg + lane * 144 + j * 16) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = 4 * i + rsub;
                        float4 o = *reinterpret_cast<const float4 *>(stg + r * 144 + (lane & 7) * 16);
                        if (n_warp + r < q.n_count) {
                            float4 *dst = reinterpret_cast<float4 *>(gxw + (size_t)r * kGXLd + reg * 128 + part * 32);
                            if (q.bwd_accumulate) {
                                const float4 old = *dst;
                                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                            }
                            *dst = o;
                        }
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tm_empty);
            BT(12);
            if (bt_on) bt[13] += 1;
        }
        if (bt_on)
            for (int i = 0; i < 14; ++i) q.dbg[(size_t)blockIdx.x * 16 + i] = bt[i];
#undef BT
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// gX (B*N, 384) in the kernel's channel order [feat 256 | skip 64 | x y z-2.2] -> d/d(point): the adjoint of the
// two bilinear gathers (grid_sampler_2d_backward semantics, zeros padding) and of the projection.  Warp per point.
__global__ void __launch_bounds__(256) query_bwd_geom_kernel(const TcParams q) {
    const int lane = threadIdx.x & 31;
    const long long p = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (p >= (long long)q.B * q.n_count) return;
    const int b = (int)(p / q.n_count);
    const long long n = p % q.n_count;
    float x, y, z, nx, ny;
    load_point(q, b, n, x, y, z);
    project_tc(x, y, z, __ldg(q.crop_center + b * 2), __ldg(q.crop_center + b * 2 + 1), nx, ny);
    const float *gx = q.gX + ((size_t)b * q.N + q.n_start + n) * kGXLd;
    float gnx = 0.f, gny = 0.f;
#pragma unroll
    for (int map = 0; map < 2; ++map) {
        const int H = map == 0 ? q.fh : 2 * q.fh, W = map == 0 ? q.fw : 2 * q.fw, C = map == 0 ? kFeatC : kSkipC;
        const float *M = map == 0 ? q.feat + (size_t)b * q.fh * q.fw * kFeatC : q.skip + (size_t)b * (2 * q.fh) * (2 * q.fw) * kSkipC;
        const float ix = __fmul_rn(__fadd_rn(nx, 1.0f), 0.5f * (float)(W - 1)), iy = __fmul_rn(__fadd_rn(ny, 1.0f), 0.5f * (float)(H - 1));
        float gix = 0.f, giy = 0.f;
        if (ix > -1.0f && ix < (float)W && iy > -1.0f && iy < (float)H) {
            const float fx0 = floorf(ix), fy0 = floorf(iy);
            const int x0 = (int)fx0, y0 = (int)fy0;
            const float w = ix - fx0, e = 1.f - w, nn = iy - fy0, s = 1.f - nn;
            const bool xl = x0 >= 0, xr = x0 + 1 < W, yt = y0 >= 0, yb = y0 + 1 < H;
            for (int c = lane * 2; c < C; c += 64) {
                float2 v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool ok = ((k & 1) ? xr : xl) && ((k >> 1) ? yb : yt);
                    v[k] = ok ? __ldg(reinterpret_cast<const float2 *>(M + ((size_t)(y0 + (k >> 1)) * W + (x0 + (k & 1))) * C + c)) : make_float2(0.f, 0.f);
                }
                float2 g = __ldg(reinterpret_cast<const float2 *>(gx + (map == 0 ? 0 : 256) + c));
                for (int sl = 1; sl < q.bwd_nslots; ++sl) {      // heads evaluated concurrently: sum their buffers in slot order
                    const float2 g2 = __ldg(reinterpret_cast<const float2 *>(gx + (size_t)sl * q.gx_slot_stride + (map == 0 ? 0 : 256) + c));
                    g.x += g2.x; g.y += g2.y;
                }
                gix += g.x * (s * (v[1].x - v[0].x) + nn * (v[3].x - v[2].x)) + g.y * (s * (v[1].y - v[0].y) + nn * (v[3].y - v[2].y));
                giy += g.x * (e * (v[2].x - v[0].x) + w * (v[3].x - v[1].x)) + g.y * (e * (v[2].y - v[0].y) + w * (v[3].y - v[1].y));
            }
        }
        gix = warp_sum(gix);
        giy = warp_sum(giy);
        gnx += gix * (0.5f * (float)(W - 1));
        gny += giy * (0.5f * (float)(H - 1));
    }
    if (lane == 0) {
        const float gpx = (gnx / 1200.0f) * 2.0f, gpy = (gny / 1200.0f) * 2.0f;
        const float ux = kFx * x, uy = kFy * y;
        float *d = q.g_points + ((size_t)b * q.N + q.n_start + n) * 3;
        float g320 = gx[320], g321 = gx[321], g322 = gx[322];
        for (int sl = 1; sl < q.bwd_nslots; ++sl) {
            const float *g2 = gx + (size_t)sl * q.gx_slot_stride;
            g320 += g2[320]; g321 += g2[321]; g322 += g2[322];
        }
        d[0] = kFx * (gpx / z) + g320;
        d[1] = kFy * (gpy / z) + g321;
        d[2] = -gpx * ((ux / z) / z) - gpy * ((uy / z) / z) + g322;
    }
}

}   // namespace

// gradient to the points on the tensor cores: one query_bwd_tc_kernel launch per head with a gradient, then the
// geometry kernel.  gX scratch: (B*N, 384) fp32 in the handle.
int query_bwd_tc_launch(chore_handle *h, const float *feat, const float *skip, int fh, int fw, const float *points,
                        const float *crop_center, int B, long long N, const float *const g_heads[4], float *g_points,
                        void *workspace, size_t workspace_bytes, cudaStream_t st) {
    TcParams q{};
    q.feat = feat; q.skip = skip; q.fh = fh; q.fw = fw;
    q.points = points; q.crop_center = crop_center;
    q.B = B; q.N = N; q.n_start = 0; q.n_count = N; q.grid_mode = 0;
    const MlpWeights &m = h->mlp;
    q.b1 = m.b1; q.b2 = m.b2; q.b3 = m.b3; q.w4 = m.w4; q.b4 = m.b4;
    q.tiles_per_b = (N + kTileM - 1) / kTileM;
    q.total_tiles = q.tiles_per_b * B;
    q.g_points = g_points;
    const size_t need = (size_t)B * N * kGXLd * sizeof(float);
    int heads[4], nheads = 0;
    for (int hd = 0; hd < 4; ++hd)
        if (g_heads[hd]) heads[nheads++] = hd;
    if (nheads == 0) {
        CHORE_CUDA(cudaMemsetAsync(g_points, 0, (size_t)B * N * 3 * sizeof(float), st));
        return CHORE_OK;
    }
    // one gX buffer per head => all heads in ONE launch (small problems are latency bound: a fit step has 54..157 tiles
    // and two heads); falls back to one launch per head accumulating into a single buffer when the scratch is too small
    int nslots = 1;
    if (workspace != nullptr) {
        // caller-owned scratch (stable under CUDA-graph capture: the pointer is baked into the graph)
        CHORE_CHECK(workspace_bytes >= need, "query backward workspace too small: %zu < %zu bytes", workspace_bytes, need);
        q.gX = static_cast<float *>(workspace);
        if (workspace_bytes >= need * nheads) nslots = nheads;
    } else {
        const size_t want = need * nheads <= ((size_t)512 << 20) ? need * nheads : need;
        if (h->bwd_ws_bytes < want) {
            if (h->bwd_ws) CHORE_CUDA(cudaFree(h->bwd_ws));
            h->bwd_ws = nullptr; h->bwd_ws_bytes = 0;
            CHORE_CUDA(cudaMalloc(&h->bwd_ws, want));
            h->bwd_ws_bytes = want;
        }
        q.gX = static_cast<float *>(h->bwd_ws);
        if (h->bwd_ws_bytes >= need * nheads) nslots = nheads;
    }
    q.gx_slot_stride = (long long)B * N * kGXLd;
    q.wstream_bwd = m.wstream_bwd;
    CHORE_ONCE_PER_DEVICE(cudaFuncSetAttribute(query_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmemBytes));
    if (nslots > 1) {
        q.bwd_nslots = nslots; q.bwd_accumulate = 0;
        for (int i = 0; i < nslots; ++i) { q.bwd_heads[i] = heads[i]; q.g_heads[i] = g_heads[heads[i]]; }
        const long long items = q.total_tiles * nslots;
        const long long grid = items < h->sm_count ? items : h->sm_count;
        static const bool trace = getenv("CHORE_B200_TC_TRACE") != nullptr;
        unsigned long long *dbg = nullptr;
        if (trace) {
            CHORE_CUDA(cudaMalloc(&dbg, (size_t)grid * 16 * sizeof(unsigned long long)));
            CHORE_CUDA(cudaMemsetAsync(dbg, 0, (size_t)grid * 16 * sizeof(unsigned long long), st));
            q.dbg = dbg;
        }
        CHORE_LAUNCH(query_bwd_tc_kernel, (unsigned)grid, kThreads, kBwdSmemBytes, st, q);
        if (trace) {   // debugging aid: synchronous, mean cycles per work item and phase (epilogue warp 0)
            std::vector<unsigned long long> hb((size_t)grid * 16);
            CHORE_CUDA(cudaMemcpyAsync(hb.data(), dbg, hb.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            CHORE_CUDA(cudaStreamSynchronize(st));
            CHORE_CUDA(cudaFree(dbg));
            q.dbg = nullptr;
            static const char *names[13] = {"start+gather", "wait:F1", "wait:F2", "wait:F3", "epi:F1", "epi:F2", "epi:F3(+W4^T g)", "wait:B3", "wait:B2",
                                            "epi:B3", "epi:B2", "wait:B1", "store:gX"};
            double sum[14] = {0};
            for (long long c = 0; c < grid; ++c)
                for (int i = 0; i < 14; ++i) sum[i] += (double)hb[(size_t)c * 16 + i];
            fprintf(stderr, "[bwd-trace] items %lld on %lld CTAs; cycles per item:", items, grid);
            double tot = 0;
            for (int i = 0; i < 13; ++i) { fprintf(stderr, " %s=%.0f", names[i], sum[i] / sum[13]); tot += sum[i] / sum[13]; }
            fprintf(stderr, " total=%.0f\n", tot);
        }
    } else {
        const long long grid = q.total_tiles < h->sm_count ? q.total_tiles : h->sm_count;
        q.bwd_nslots = 1;
        for (int i = 0; i < nheads; ++i) {
            q.bwd_heads[0] = heads[i]; q.g_heads[0] = g_heads[heads[i]]; q.bwd_accumulate = i > 0;
            CHORE_LAUNCH(query_bwd_tc_kernel, (unsigned)grid, kThreads, kBwdSmemBytes, st, q);
        }
    }
    const long long pts = (long long)B * N;
    CHORE_LAUNCH(query_bwd_geom_kernel, (unsigned)((pts + 7) / 8), 256, 0, st, q);
    return CHORE_OK;
}

