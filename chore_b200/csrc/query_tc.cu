// Fused point query on the 5th-generation tensor cores (tcgen05 + TMEM), forward only.
//
// Same contract as query_fwd_kernel (query.cu) -- projection, two bilinear gathers, 4-head MLP,
// OUT_DIST mask, reference output layouts -- but the three 128-wide layers of every head run as
// tcgen05.mma (kind::f16, M=128 points x N=128 channels x K=16 per instruction) with fp32
// accumulators in tensor memory.
//
// fp32-faithful math on fp16 tensor cores: every operand is split x = hi + lo with
// hi = fp16(x), lo = fp16(x - hi) (22 significant bits), and a product is evaluated as
// hi*hi + lo*hi + hi*lo with fp32 accumulation: three MMAs per logical MMA, error ~2^-22 per
// term, i.e. the same order as an fp32 FMA chain (the dropped lo*lo term is 2^-22 relative).
//
// One persistent CTA per SM, 14 warps, everything synchronised with mbarriers:
//   warp 0      weight producer: cp.async.bulk of pre-swizzled 16 KB weight panels (the whole
//               1.25 MB weight stream is re-read from L2 once per 128-point tile)
//   warp 1      MMA issuer (one elected lane), owns the TMEM allocation (512 columns:
//               one 128x128 fp32 accumulator per head)
//   warps 2-5   gather: projection + bilinear taps -> A operand k-blocks (64 channels, hi/lo
//               fp16, 128B-swizzled K-major) in a 2-stage ring
//   warps 6-13  epilogue: TMEM -> registers, bias + ReLU, hi/lo split -> activation k-blocks
//               (A operand of the next layer, 3-stage ring); after layer 3 the last (<=14-wide)
//               layer is evaluated in fp32 on the CUDA cores and written to HBM
//
// Per tile the issue order is L1(kb, head) for 6 k-blocks x 4 heads (all four accumulators
// live), then L2(head), L3(head); epilogues of one head overlap the MMAs of the others.
#include "query_tc_shared.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

// w_full / act_full exist once per issuer.  A parity wait is only meaningful for a waiter that is at most one phase ahead
// of the barrier; on ONE barrier per ring slot, the issuer whose turn on a slot comes second tests a phase that is two ahead
// and passes before the first issuer's panel has landed (seen as size-dependent hangs for the three-head masks 7 / 11 /
// 13 / 14, where the issuers do not alternate slot by slot).  Producers therefore arrive on the barrier of the issuer that
// owns the head, and every issuer tracks the parity of its own uses.  (a_full is waited on by both issuers for every block.)
struct Bars {
    uint64_t a_full[kNA], a_empty[kNA];
    uint64_t w_full[2][kNW], w_empty[kNW];
    uint64_t act_full[2][kNACT], act_empty[kNACT];
    uint64_t tm_full[4], tm_empty[4];
    uint32_t tmem_base;
};

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// 15 warps: warp 14 is a second MMA issuer.  The thread that issues a tcgen05.mma is held until the tensor pipe takes
// the instruction, so whatever it does between two MMAs (barrier waits, fences, descriptor moves, commits) is added to
// the math time instead of overlapping it (profiles/mma_microbench_r1.txt).  With two issuers -- warp 1 owns heads 0 and 2,
// warp 14 heads 1 and 3, same global order of weight panels and activation blocks -- one prepares its next step while the
// other's MMAs execute.
constexpr int kIssuerB = 14;
constexpr int kThreadsFwd = 15 * 32;
__global__ void __launch_bounds__(kThreadsFwd, 1) query_tc_kernel(const TcParams q) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *ringA = smem;                                      // [kNA][hi 16K | lo 16K]
    uint8_t *ringAct = ringA + (size_t)kNA * kStageA;           // [kNACT][hi 16K | lo 16K]
    uint8_t *ringW = ringAct + (size_t)kNACT * kStageA;         // [kNW][16K]
    Bars *bars = reinterpret_cast<Bars *>(ringW + (size_t)kNW * kPanelBytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tracing: one counter row per CTA, written by lane 0 of warps 0, 1, 2 and 6 only
    unsigned long long dbg_local[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const bool dbg_on = q.dbg != nullptr && lane == 0 && (warp == 0 || warp == 1 || warp == 2 || warp == 6);
    const long long dbg_t0 = clock64();
#define DBG(i) (dbg_on ? &dbg_local[i] : nullptr)

    const unsigned owner_mask = 10u;             // heads whose steps the second issuer (warp 14) runs
    if (threadIdx.x == 0) {
        // an A stage is released by every issuer that reads it
        const int n_issuers = ((q.head_mask & ~owner_mask) != 0) + ((q.head_mask & owner_mask) != 0);
        for (int i = 0; i < kNA; ++i) { mbar_init(&bars->a_full[i], 4); mbar_init(&bars->a_empty[i], n_issuers); }
        for (int i = 0; i < kNW; ++i) { mbar_init(&bars->w_full[0][i], 1); mbar_init(&bars->w_full[1][i], 1); mbar_init(&bars->w_empty[i], 1); }
        for (int i = 0; i < kNACT; ++i) { mbar_init(&bars->act_full[0][i], 4); mbar_init(&bars->act_full[1][i], 4); mbar_init(&bars->act_empty[i], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&bars->tm_full[i], 1); mbar_init(&bars->tm_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 1) {   // TMEM: all 512 columns (4 heads x 128 fp32 accumulator columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars->tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // warp-uniform by construction (lets the issue loop live in uniform registers)
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, bars->tmem_base, 0);

    const long long first_tile = blockIdx.x, tile_stride = gridDim.x;

    if (warp == 0) {
        // =============================== weight producer ===============================
        // the whole warp walks the loop (uniform control flow); one elected lane issues the copies
        uint32_t u = 0;   // running panel counter over all tiles
        for (long long tile = first_tile; tile < q.total_tiles; tile += tile_stride) {
            for (int i = 0; i < kUnitsPerTile; ++i) {
                // heads outside head_mask are not evaluated at all: their panels are never streamed
                const int hd_i = i < kL1Blocks * 8 ? (i >> 1) & 3 : (i < kBigUnits ? ((i - kL1Blocks * 8) >> 2) & 3 : i - kBigUnits);
                if (!((q.head_mask >> hd_i) & 1)) continue;
                const int s = u % kNW;
                mbar_wait_t(&bars->w_empty[s], ((u / kNW) & 1) ^ 1, DBG(0));
                if (elect_one()) {
                    const uint32_t bytes = i < kBigUnits ? kPanelBytes : kSmallPanelBytes;
                    const size_t off = i < kBigUnits ? (size_t)i * kPanelBytes
                                                     : (size_t)kBigUnits * kPanelBytes + (size_t)(i - kBigUnits) * kSmallPanelBytes;
                    uint64_t *full = &bars->w_full[(owner_mask >> hd_i) & 1][s];
                    mbar_arrive_expect_tx(full, bytes);
                    bulk_g2s(ringW + (size_t)s * kPanelBytes, q.wstream + off, bytes, full);
                }
                __syncwarp();
                ++u;
            }
        }
    } else if (warp == 1 || warp == kIssuerB) {
        // =============================== MMA issuers ===============================
        // All 32 lanes run the loop converged so that every operand is warp-uniform; only the tcgen05
        // instructions themselves are predicated on one elected lane.  Warp 1 issues for heads 0 and 2, warp 14 for heads
        // 1 and 3 (consecutive steps alternate between the issuers); both walk the same global sequence of weight panels /
        // activation blocks and skip the other's entries.
        const int me = warp == 1 ? 0 : 1;
        const unsigned my_heads = q.head_mask & (me == 0 ? ~owner_mask : owner_mask);
        uint32_t pw = 0, pa = 0;                    // parity of this issuer's next use of every weight / activation slot
        constexpr uint32_t idesc = make_idesc(kTileM, 128), idesc16 = make_idesc(kTileM, 16);
        const uint32_t ringA_lo = desc_lo(smem_u32(ringA)), ringAct_lo = desc_lo(smem_u32(ringAct)), ringW_lo = desc_lo(smem_u32(ringW));
        constexpr uint32_t kStageLo = kStageA >> 4, kPanelLo = kPanelBytes >> 4;
        uint32_t u = 0;        // weight panel counter
        uint32_t ablk = 0;     // A ring block counter
        uint32_t actblk = 0;   // activation ring block counter
        uint32_t tile_i = 0;
        const int last_head = 31 - __clz((int)my_heads);     // this issuer releases the A stage after its last active head
        for (long long tile = first_tile; my_heads != 0 && tile < q.total_tiles; tile += tile_stride, ++tile_i) {
            // ---- layer 1: 6 k-blocks x 4 heads, all four accumulators live ----
            for (int kb = 0; kb < kL1Blocks; ++kb, ++ablk) {
                const int sa = ablk % kNA;
                mbar_wait_t(&bars->a_full[sa], (ablk / kNA) & 1, DBG(1));
                tc_fence_after();
                const uint32_t a_hi = ringA_lo + sa * kStageLo, a_lo = a_hi + kPanelLo;
                const bool last = kb == kL1Blocks - 1;       // the xyz block holds 16 channels: one k-step
#pragma unroll 1
                for (int h = 0; h < 4; ++h) {
                    if (!((q.head_mask >> h) & 1)) continue;
                    if (!((my_heads >> h) & 1)) { u += 2; continue; }        // the other issuer's panels
                    if (kb == 0) {   // accumulator of head h must have been drained (previous tile)
                        mbar_wait_t(&bars->tm_empty[h], (tile_i & 1) ^ 1, DBG(2));
                        tc_fence_after();
                    }
                    const uint32_t d = tmem_base + h * 128;
                    const int s0 = u % kNW, s1 = (u + 1) % kNW;
                    mbar_wait_t(&bars->w_full[me][s0], (pw >> s0) & 1, DBG(3));   // panel w_hi: a_hi*w_hi + a_lo*w_hi
                    pw ^= 1u << s0;
                    tc_fence_after();
                    const uint32_t w0 = ringW_lo + s0 * kPanelLo;
                    if (elect_one()) {
                        if (!last) umma_burst_pair<4>(d, desc64(a_hi), desc64(a_lo), desc64(w0), idesc, kb != 0);
                        else umma_burst_pair<1>(d, desc64(a_hi), desc64(a_lo), desc64(w0), idesc, kb != 0);
                        umma_commit(&bars->w_empty[s0]);
                    }
                    __syncwarp();
                    mbar_wait_t(&bars->w_full[me][s1], (pw >> s1) & 1, DBG(3));   // panel w_lo: a_hi*w_lo
                    pw ^= 1u << s1;
                    tc_fence_after();
                    const uint32_t w1 = ringW_lo + s1 * kPanelLo;
                    if (elect_one()) {
                        if (!last) umma_burst_single<4>(d, desc64(a_hi), desc64(w1), idesc);
                        else umma_burst_single<1>(d, desc64(a_hi), desc64(w1), idesc);
                        umma_commit(&bars->w_empty[s1]);
                        if (last) umma_commit(&bars->tm_full[h]);       // layer-1 accumulator of head h complete
                        if (h == last_head) umma_commit(&bars->a_empty[sa]);
                    }
                    __syncwarp();
                    u += 2;
                }
            }
            // ---- layers 2, 3 (128 wide) and 4 (16 wide): per head, A operand = activation blocks ----
#pragma unroll 1
            for (int layer = 1; layer < 4; ++layer) {
#pragma unroll 1
                for (int h = 0; h < 4; ++h) {
                    if (!((q.head_mask >> h) & 1)) continue;
                    if (!((my_heads >> h) & 1)) { actblk += 2; u += layer < 3 ? 4 : 1; continue; }     // the other issuer's head
                    const uint32_t d = tmem_base + h * 128;
                    // both activation blocks must be complete before the accumulator is overwritten
                    const uint32_t b0 = actblk, b1 = actblk + 1;
                    mbar_wait_t(&bars->act_full[me][b0 % kNACT], (pa >> (b0 % kNACT)) & 1, DBG(4));
                    mbar_wait_t(&bars->act_full[me][b1 % kNACT], (pa >> (b1 % kNACT)) & 1, DBG(4));
                    pa ^= (1u << (b0 % kNACT)) | (1u << (b1 % kNACT));
                    tc_fence_after();
                    if (layer < 3) {
#pragma unroll 1
                        for (int kb = 0; kb < 2; ++kb, ++actblk) {
                            const int sa = actblk % kNACT;
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
                        // last layer (<= 14 outputs, padded to N = 16): one 8 KB panel [kb0 hi | kb0 lo | kb1 hi | kb1 lo]
                        const int s0 = u % kNW;
                        mbar_wait_t(&bars->w_full[me][s0], (pw >> s0) & 1, DBG(5));
                        pw ^= 1u << s0;
                        tc_fence_after();
                        const uint32_t w = ringW_lo + s0 * kPanelLo;
                        const int sa0 = actblk % kNACT, sa1 = (actblk + 1) % kNACT;
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
        }
    } else if (warp < kEpiWarp0) {
        // =============================== gather warps ===============================
        // warp g owns rows [32g, 32g+32); a half-warp handles one point: 16 lanes x 4 channels = one 64-channel k-block
        const int g = warp - kGatherWarp0;
        const int half = lane >> 4, l16 = lane & 15;
        uint32_t ablk = 0;
        for (long long tile = first_tile; tile < q.total_tiles; tile += tile_stride) {
            const int b = q.grid_mode ? q.batch_index : (int)(tile / q.tiles_per_b);
            const long long n0 = (tile % q.tiles_per_b) * kTileM;
            const float ccx = __ldg(q.crop_center + b * 2), ccy = __ldg(q.crop_center + b * 2 + 1);
            const float *F = q.feat + (size_t)b * q.fh * q.fw * kFeatC;
            const float *S = q.skip + (size_t)b * (2 * q.fh) * (2 * q.fw) * kSkipC;
            // lane L projects row 32g + L once per tile; the k-block loop fetches it with shuffles
            float my_x = 0.f, my_y = 0.f, my_z = 1.f, my_nx, my_ny;
            if (n0 + g * 32 + lane < q.n_count) load_point(q, b, n0 + g * 32 + lane, my_x, my_y, my_z);
            project_tc(my_x, my_y, my_z, ccx, ccy, my_nx, my_ny);
            const LaneTaps tapsF = make_lane_taps(my_nx, my_ny, q.fh, q.fw, kFeatC), tapsS = make_lane_taps(my_nx, my_ny, 2 * q.fh, 2 * q.fw, kSkipC);
            for (int kb = 0; kb < kL1Blocks; ++kb, ++ablk) {
                const int sa = ablk % kNA;
                mbar_wait_t(&bars->a_empty[sa], ((ablk / kNA) & 1) ^ 1, DBG(6));
                uint8_t *hi = ringA + (size_t)sa * kStageA, *lo = hi + kPanelBytes;
                if (kb < 5) {
                    const bool is_feat = kb < 4;
                    gather_kblock_p<32, kGatherBatch>((is_feat ? F + kb * 64 : S) + l16 * 4, (is_feat ? q.fw * kFeatC : 2 * q.fw * kSkipC),
                                                      is_feat ? kFeatC : kSkipC, is_feat ? tapsF : tapsS, g, half, l16, hi, lo);
                } else {
                    // z_feat = [x, y, z - 2.2] (model/chore.py:128-129) + 13 zero channels: one k-step, lane = row
                    const int r = g * 32 + lane;
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
        // =============================== epilogue warps ===============================
        const int e = warp - kEpiWarp0;              // 0..7
        const int quarter = warp & 3;                // TMEM lane quarter this warp may access
        const int colhalf = e >> 2;                  // 0: columns 0-63, 1: columns 64-127
        const int row = quarter * 32 + lane;         // point of the tile owned by this thread
        uint32_t actblk = 0;                         // activation block counter (this warp's column half adds colhalf)
        for (long long tile = first_tile; tile < q.total_tiles; tile += tile_stride) {
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
            for (int layer = 0; layer < 4; ++layer) {
#pragma unroll 1
                for (int h = 0; h < 4; ++h) {
                    if (!((q.head_mask >> h) & 1)) continue;             // head not requested: nothing was computed
                    // every warp observes EVERY completion of tm_full[h] (4 per tile: parity = layer & 1), also the group that has
                    // no output-layer work for this head: a warp that skipped the wait would test the next tile's layer-1 phase
                    // one phase early, pass on the parity of layer 3 and read a stale accumulator (seen with single-head masks
                    // on more than 148 tiles, where one group has no output-layer work at all)
                    mbar_wait_t(&bars->tm_full[h], layer & 1, DBG(7));
                    if (layer == 3 && (h & 1) != colhalf) continue;      // last layer: the heads are split between the groups
                    tc_fence_after();
                    if (layer < 3) {
                        // bias + ReLU + hi/lo split -> activation k-block `colhalf` of head h
                        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + h * 128 + colhalf * 64;
                        const float *bias = (layer == 0 ? q.b1 : (layer == 1 ? q.b2 : q.b3)) + h * 128 + colhalf * 64;
                        const uint32_t blk = actblk + colhalf;
                        const int sa = blk % kNACT;
                        mbar_wait_t(&bars->act_empty[sa], ((blk / kNACT) & 1) ^ 1, DBG(8));
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
                        // last layer: 16 accumulator columns -> + bias -> OUT_DIST mask -> HBM (reference layout)
                        const int nout = head_out_tc(h);
                        uint32_t v[16];
                        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + h * 128, v);
                        tmem_ld_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars->tm_empty[h]);   // the next tile's layer 1 may overwrite head h
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

// fp16 hi/lo panels of W[n][k] (n = output channel, k = input channel of this k-block), 128B-swizzled
void pack_panel(const float *w /* rows x 64 row-major (n, k) */, int rows, uint8_t *hi, uint8_t *lo) {
    for (int n = 0; n < rows; ++n)
        for (int k = 0; k < 64; ++k) {
            float v = w[n * 64 + k];
            v = v > 65504.f ? 65504.f : (v < -65504.f ? -65504.f : v);
            const __half h = __float2half_rn(v);
            const __half l = __float2half_rn(v - __half2float(h));
            const size_t off = (size_t)n * 128 + (size_t)(((k >> 3) ^ (n & 7)) << 4) + (size_t)(k & 7) * 2;
            const unsigned short hs = __half_as_ushort(h), ls = __half_as_ushort(l);
            memcpy(hi + off, &hs, 2);
            memcpy(lo + off, &ls, 2);
        }
}

}   // namespace

// builds the per-tile weight stream from the fp32 MLP weights
int query_tc_pack_weights(chore_handle *h, const std::vector<float> &w1 /*[4][128][323]*/,
                          const std::vector<float> &w2 /*[4][128][128]*/, const std::vector<float> &w3,
                          const std::vector<float> &w4 /*[4][16][128], rows >= n_out zero*/) {
    std::vector<uint8_t> stream(kStreamBytes, 0);
    std::vector<float> tmp(128 * 64);
    size_t u = 0;
    for (int kb = 0; kb < kL1Blocks; ++kb)
        for (int hd = 0; hd < 4; ++hd) {
            for (int n = 0; n < 128; ++n)
                for (int k = 0; k < 64; ++k) {
                    int c = -1;                       // channel of the reference's 323-vector
                    if (kb < 4) c = kb * 64 + k;      // hourglass feature
                    else if (kb == 4) c = 259 + k;    // stem skip feature
                    else if (k < 3) c = 256 + k;      // x, y, z - 2.2
                    tmp[n * 64 + k] = c >= 0 ? w1[((size_t)hd * 128 + n) * kPointC + c] : 0.f;
                }
            pack_panel(tmp.data(), 128, stream.data() + u * kPanelBytes, stream.data() + (u + 1) * kPanelBytes);
            u += 2;
        }
    for (int layer = 0; layer < 2; ++layer) {
        const std::vector<float> &w = layer == 0 ? w2 : w3;
        for (int hd = 0; hd < 4; ++hd)
            for (int kb = 0; kb < 2; ++kb) {
                for (int n = 0; n < 128; ++n)
                    for (int k = 0; k < 64; ++k) tmp[n * 64 + k] = w[((size_t)hd * 128 + n) * 128 + kb * 64 + k];
                pack_panel(tmp.data(), 128, stream.data() + u * kPanelBytes, stream.data() + (u + 1) * kPanelBytes);
                u += 2;
            }
    }
    if (u != (size_t)kBigUnits) {
        chore_set_error("internal: weight stream has %zu panels, expected %d", u, kBigUnits);
        return CHORE_ERR_INVALID;
    }
    for (int hd = 0; hd < 4; ++hd) {   // last layer: [kb0 hi | kb0 lo | kb1 hi | kb1 lo], 16 rows x 128 B each
        uint8_t *base = stream.data() + (size_t)kBigUnits * kPanelBytes + (size_t)hd * kSmallPanelBytes;
        for (int kb = 0; kb < 2; ++kb) {
            for (int n = 0; n < 16; ++n)
                for (int k = 0; k < 64; ++k) tmp[n * 64 + k] = w4[((size_t)hd * 16 + n) * 128 + kb * 64 + k];
            pack_panel(tmp.data(), 16, base + kb * 4096, base + kb * 4096 + 2048);
        }
    }
    if (int rc = chore_dev_alloc(h, reinterpret_cast<void **>(&h->mlp.wstream), stream.size())) return rc;
    CHORE_CUDA(cudaMemcpy(h->mlp.wstream, stream.data(), stream.size(), cudaMemcpyHostToDevice));

    // backward streams, one per head: F1 (6 k-blocks), F2, F3, B3 = W3^T, B2 = W2^T, B1 = W1^T (3 column chunks)
    std::vector<uint8_t> bs((size_t)4 * kBwdUnits * kPanelBytes, 0);
    auto orig_channel = [](int c) { return c < 256 ? c : (c < 320 ? 259 + (c - 256) : (c < 323 ? 256 + (c - 320) : -1)); };
    for (int hd = 0; hd < 4; ++hd) {
        uint8_t *base = bs.data() + (size_t)hd * kBwdUnits * kPanelBytes;
        size_t v = 0;
        auto emit = [&]() { pack_panel(tmp.data(), 128, base + v * kPanelBytes, base + (v + 1) * kPanelBytes); v += 2; };
        for (int kb = 0; kb < kL1Blocks; ++kb) {                       // F1
            for (int n = 0; n < 128; ++n)
                for (int k = 0; k < 64; ++k) {
                    const int c = kb < 5 ? orig_channel(kb * 64 + k) : (k < 3 ? 256 + k : -1);
                    tmp[n * 64 + k] = c >= 0 ? w1[((size_t)hd * 128 + n) * kPointC + c] : 0.f;
                }
            emit();
        }
        for (int layer = 0; layer < 2; ++layer)                          // F2, F3
            for (int kb = 0; kb < 2; ++kb) {
                const std::vector<float> &w = layer == 0 ? w2 : w3;
                for (int n = 0; n < 128; ++n)
                    for (int k = 0; k < 64; ++k) tmp[n * 64 + k] = w[((size_t)hd * 128 + n) * 128 + kb * 64 + k];
                emit();
            }
        for (int layer = 1; layer >= 0; --layer)                         // B3 (W3^T), B2 (W2^T): B[n][k] = W[k][n]
            for (int kb = 0; kb < 2; ++kb) {
                const std::vector<float> &w = layer == 0 ? w2 : w3;
                for (int n = 0; n < 128; ++n)
                    for (int k = 0; k < 64; ++k) tmp[n * 64 + k] = w[((size_t)hd * 128 + kb * 64 + k) * 128 + n];
                emit();
            }
        for (int nc = 0; nc < 3; ++nc)                                   // B1: B[n = c' - 128 nc][k = j] = W1[j][orig(c')]
            for (int kb = 0; kb < 2; ++kb) {
                for (int n = 0; n < 128; ++n) {
                    const int c = orig_channel(nc * 128 + n);
                    for (int k = 0; k < 64; ++k) tmp[n * 64 + k] = c >= 0 ? w1[((size_t)hd * 128 + kb * 64 + k) * kPointC + c] : 0.f;
                }
                emit();
            }
        if (v != (size_t)kBwdUnits) {
            chore_set_error("internal: backward weight stream has %zu panels, expected %d", v, kBwdUnits);
            return CHORE_ERR_INVALID;
        }
    }
    if (int rc = chore_dev_alloc(h, reinterpret_cast<void **>(&h->mlp.wstream_bwd), bs.size())) return rc;
    CHORE_CUDA(cudaMemcpy(h->mlp.wstream_bwd, bs.data(), bs.size(), cudaMemcpyHostToDevice));
    return CHORE_OK;
}

int query_tc_launch(chore_handle *h, const float *feat, const float *skip, int fh, int fw, const float *points,
                    const float *crop_center, int B, long long N, long long n_start, long long n_count, int grid_mode,
                    int batch_index, const int *res, const double *step, const double *bmin, unsigned head_mask,
                    float *const outs[4], unsigned char *in_img, cudaStream_t st) {
    TcParams q{};
    q.feat = feat; q.skip = skip; q.fh = fh; q.fw = fw;
    q.points = points; q.crop_center = crop_center;
    q.B = B; q.N = N; q.n_start = n_start; q.n_count = n_count;
    q.grid_mode = grid_mode; q.batch_index = batch_index;
    if (grid_mode) {
        q.ry = res[1]; q.rz = res[2];
        for (int i = 0; i < 3; ++i) { q.step[i] = step[i]; q.bmin[i] = bmin[i]; }
    }
    q.head_mask = head_mask;
    for (int i = 0; i < 4; ++i) q.out[i] = outs[i];
    q.in_img = in_img;
    const MlpWeights &m = h->mlp;
    q.wstream = m.wstream; q.b1 = m.b1; q.b2 = m.b2; q.b3 = m.b3; q.w4 = m.w4; q.b4 = m.b4;
    q.tiles_per_b = (n_count + kTileM - 1) / kTileM;
    q.total_tiles = q.tiles_per_b * (grid_mode ? 1 : B);
    CHORE_ONCE_PER_DEVICE(cudaFuncSetAttribute(query_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    const long long grid = q.total_tiles < h->sm_count ? q.total_tiles : h->sm_count;
    static const bool trace = getenv("CHORE_B200_TC_TRACE") != nullptr;
    unsigned long long *dbg = nullptr;
    if (trace) {
        CHORE_CUDA(cudaMalloc(&dbg, (size_t)grid * 16 * sizeof(unsigned long long)));
        CHORE_CUDA(cudaMemsetAsync(dbg, 0, (size_t)grid * 16 * sizeof(unsigned long long), st));
        q.dbg = dbg;
    }
    CHORE_LAUNCH(query_tc_kernel, (unsigned)grid, kThreadsFwd, kSmemBytes, st, q);
    if (trace) {   // debugging aid: synchronous, prints the mean wait cycles per role
        std::vector<unsigned long long> hbuf((size_t)grid * 16);
        CHORE_CUDA(cudaMemcpyAsync(hbuf.data(), dbg, hbuf.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CHORE_CUDA(cudaStreamSynchronize(st));
        CHORE_CUDA(cudaFree(dbg));
        static const char *names[10] = {"producer:w_empty", "mma:a_full", "mma:tm_empty", "mma:w_full(L1)", "mma:act_full",
                                        "mma:w_full(L2/3)", "gather:a_empty", "epi:tm_full", "epi:act_empty", "total"};
        double sum[10] = {0};
        for (long long c = 0; c < grid; ++c)
            for (int i = 0; i < 10; ++i) sum[i] += (double)hbuf[(size_t)c * 16 + i];
        const double tiles = (double)q.total_tiles / (double)grid;
        fprintf(stderr, "[tc-trace] tiles/CTA %.1f; cycles per tile:", tiles);
        for (int i = 0; i < 10; ++i) fprintf(stderr, " %s=%.0f", names[i], sum[i] / (double)grid / tiles);
        fprintf(stderr, "\n");
    }
    return CHORE_OK;
}
