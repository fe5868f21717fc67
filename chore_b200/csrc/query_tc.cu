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

// =============================================================================================
// backward to the points on the tensor cores: one launch per head with a non-zero upstream gradient
//   F1..F3  forward recompute of the three 128-wide layers (ReLU masks stay in registers)
//   gH3 = W4^T g (CUDA cores, <= 14 terms), B3, B2: gH_{l-1} = (gH_l W_l) . mask_{l-1}   (tcgen05)
//   B1      gX (+)= gH1 W1: 384 columns in three accumulator regions -> (B*N, 384) in HBM
// then query_bwd_geom_kernel turns gX into d/d(point) (bilinear + projection adjoint).
// =============================================================================================
constexpr int kBwdNW = 6;
constexpr int kBwdUnits = 12 + 4 + 4 + 4 + 4 + 12;     // F1, F2, F3, B3, B2, B1 panels of 16 KB
constexpr size_t kBwdSmemBytes = 1024 + 2 * (size_t)kStageA + 2 * (size_t)kStageA + (size_t)kBwdNW * kPanelBytes + 512;
constexpr int kGXLd = 384;

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
        for (long long item = first_tile; item < total_items; item += tile_stride) {
            const long long tile = item / nslots;
            const int slot = (int)(item % nslots);
            const int hd = q.bwd_heads[slot];
            const int nout = head_out_tc(hd);
            const float *g_head = q.g_heads[slot];
            const int b = (int)(tile / q.tiles_per_b);
            const long long n = (tile % q.tiles_per_b) * kTileM + row;
            const bool live = n < q.n_count;
            // upstream gradient of this row (zero for dead rows and for df outside the image)
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
            uint32_t mask[3][2];
            // ---- forward recompute: F1, F2, F3 ----
#pragma unroll 1
            for (int layer = 0; layer < 3; ++layer) {
                mbar_wait(&bars->tm_full, layer & 1);          // 6 completions per tile: parity = phase & 1
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
            }
            // ---- backward chain: B3 -> (. mask2), B2 -> (. mask1) ----
#pragma unroll 1
            for (int step = 0; step < 2; ++step) {
                mbar_wait(&bars->tm_full, (3 + step) & 1);
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
            }
            // ---- B1 result: accumulator regions 1..3 -> gX[point][384] ----
            mbar_wait(&bars->tm_full, 5 & 1);
            tc_fence_after();
            float *gx = q.gX + (size_t)slot * q.gx_slot_stride + ((size_t)b * q.N + q.n_start + (live ? n : 0)) * kGXLd + colhalf * 64;
#pragma unroll 1
            for (int reg = 0; reg < 3; ++reg) {
#pragma unroll
                for (int part = 0; part < 2; ++part) {
                    uint32_t v[32];
                    tmem_ld32(tbase + (1 + reg) * 128 + part * 32, v);
                    tmem_ld_wait();
                    if (live) {
                        float4 *dst = reinterpret_cast<float4 *>(gx + reg * 128 + part * 32);
#pragma unroll
                        for (int c4 = 0; c4 < 8; ++c4) {
                            float4 o = make_float4(__uint_as_float(v[c4 * 4]), __uint_as_float(v[c4 * 4 + 1]),
                                                   __uint_as_float(v[c4 * 4 + 2]), __uint_as_float(v[c4 * 4 + 3]));
                            if (q.bwd_accumulate) {
                                const float4 old = dst[c4];
                                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                            }
                            dst[c4] = o;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tm_empty);
        }
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
        CHORE_LAUNCH(query_bwd_tc_kernel, (unsigned)grid, kThreads, kBwdSmemBytes, st, q);
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
