// Encoder convolutions, second generation ("hx" = halo + transform): one kernel per convolution that
//   * reads the fp32 NHWC input, applies the GroupNorm affine + ReLU of the consuming layer and the fp16 hi/lo split
//     in a PROLOGUE (8 transform warps) and stages ONE (16+KS-1) x (8+KS-1) halo tile per 64-channel k-block in
//     shared memory -- the 9 taps of a 3x3 convolution are row-shifted VIEWS of that tile: the tcgen05 shared-memory
//     descriptor takes a start address that is only 128-byte aligned and a stride of `pitch` rows between 8-row groups
//     (the 128-byte swizzle is a function of the absolute shared-memory address; measured with tools/desc_probe.cu),
//     so the activation is read from L2 once instead of nine times and never exists as fp16 planes in HBM;
//   * streams the pre-swizzled weight panels with cp.async.bulk through an mbarrier ring (warp 0), issues
//     tcgen05.mma kind::f16 (3-term split: hi*hi + lo*hi + hi*lo, fp32 accumulators in TMEM, double buffered) from
//     warp 1, and drains with 4 epilogue warps: bias, raw copy, residual add, concat-slice store;
//   * accumulates the GroupNorm statistics (sum, sum of squares per image and group, fp64) of what it writes in the
//     epilogue, so no separate statistics pass over the tensor is needed;
//   * is persistent over work items (pixel tile x output-channel slice x K slice).  A tcgen05.mma with both operands in
//     shared memory costs >= 64 cycles for any N <= 128 (the 4 KB A operand is re-read per instruction), so the
//     low-resolution levels, which have fewer tiles than SMs, split the K loop (taps x 64-channel blocks) over the CTAs of
//     a thread-block cluster (rank = K part).  32-column chunk c of the tile is finished by part c mod k: the other parts
//     push their fp32 partial chunk into the owner's shared memory (st.async + mbarrier complete_tx; the owner's halo
//     buffers and weight ring are idle once its own MMAs are done, which it announces with a remote mbarrier arrive) and
//     the owner adds the partials in sender order (deterministic).
// Replaces the F.group_norm + ReLU + F.conv2d triples of ConvBlock.forward (model/net_util.py:374-396) and the 1x1
// convolutions of HGFilter.forward (model/HGFilters.py:173-183).
#include "common.cuh"
#include "tc_common.cuh"

#include <cstdlib>
#include <cstring>

using namespace tc;

namespace {

constexpr int kGroups = 32;
constexpr int kTileH = 16, kTileW = 8;                 // 128 output pixels per work item (TMEM lane = 8 * row + col)
constexpr int kTxWarps = 8;
constexpr int kFirstEpiWarp = 2, kFirstTxWarp = 6;
constexpr int kThreadsHx = (kFirstTxWarp + kTxWarps) * 32;   // 448
constexpr int kMaxSlots = 8;
constexpr uint32_t kSmemBudget = 232448;               // 227 KB opt-in maximum per CTA
constexpr uint32_t kStagePitch = 144, kStageWarpBytes = 32 * kStagePitch;   // epilogue transpose tile (padded rows: conflict free)
constexpr uint32_t kRecvOff = 40960;                   // K split: received partial chunks start here (after the helpers' staging tiles)

struct HxCtl {
    uint64_t halo_full[2], halo_empty[2], w_full[kMaxSlots], w_empty[kMaxSlots], acc_full[2], acc_empty[2];
    uint64_t peer_ready[8], data_full, pad64;                 // K split: cluster partner r can receive / all partials of my chunks landed
    uint32_t tmem_base, pad[3];
    uint2 dummy[4];                                    // target of the stores of halo slots beyond the tile
    float sc[256], sh[256];
    float stp[12][2][kGroups * 2];                     // per worker warp (3 per TMEM lane quarter) / set (0 = raw, 1 = out): this item's sums
};

struct HxParams {
    ConvHxArgs a;
    int n_splits, k_splits, kt_per, Ns, tiles_x, tiles_y, n_items, kblocks, n_slots;
    uint32_t halo_plane, w_slot;
    double inv_n;          // 1 / (H * W * channels per group of the input)
    long long *trace;      // debugging: 32 clock64 stamps of CTA 0 (null = off)
};

#define HX_STAMP(i) do { if (p.trace != nullptr && blockIdx.x == 0 && lane == 0) p.trace[i] = clock64(); } while (0)
__device__ __forceinline__ void named_bar(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
// predicated 16-byte L2 load (zero when the predicate is off): never a branch
__device__ __forceinline__ float4 ldcg_pred(const float *ptr, bool pred) {
    float4 v;
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %5, 0;\nmov.f32 %0, 0f00000000;\nmov.f32 %1, 0f00000000;\nmov.f32 %2, 0f00000000;\n"
                 "mov.f32 %3, 0f00000000;\n@p ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];\n}"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr), "r"((int)pred));
    return v;
}

// the 3 MMAs of one 16-wide k step (hi*hi, lo*hi, hi*lo) for KSN consecutive k steps; descriptors are 64-bit values
// whose low word advances by 2 (32 bytes inside the 128-byte swizzle atom) per k step.  One asm block: the issuing
// thread executes nothing but the MMAs and the descriptor increments.
template <int KSN>
__device__ __forceinline__ void umma_burst(uint32_t d_tmem, uint64_t ah, uint64_t al, uint64_t wh, uint64_t wl, uint32_t idesc,
                                           uint32_t accumulate_first) {
    asm volatile(
        "{\n.reg .pred p, q;\n.reg .b64 a, b, c, d;\n"
        "setp.ne.b32 p, %6, 0;\nsetp.eq.b32 q, %5, %5;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %3, %5, p;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %3, %5, q;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %4, %5, q;\n"
        "add.s64 a, %1, 2;\nadd.s64 b, %2, 2;\nadd.s64 c, %3, 2;\nadd.s64 d, %4, 2;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a, c, %5, q;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], b, c, %5, q;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a, d, %5, q;\n"
        "}" ::"r"(d_tmem), "l"(ah), "l"(al), "l"(wh), "l"(wl), "r"(idesc), "r"(accumulate_first) : "memory");
    if (KSN == 4)
        asm volatile(
            "{\n.reg .pred q;\n.reg .b64 a, b, c, d;\n"
            "setp.eq.b32 q, %5, %5;\n"
            "add.s64 a, %1, 4;\nadd.s64 b, %2, 4;\nadd.s64 c, %3, 4;\nadd.s64 d, %4, 4;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], a, c, %5, q;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], b, c, %5, q;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], a, d, %5, q;\n"
            "add.s64 a, %1, 6;\nadd.s64 b, %2, 6;\nadd.s64 c, %3, 6;\nadd.s64 d, %4, 6;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], a, c, %5, q;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], b, c, %5, q;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], a, d, %5, q;\n"
            "}" ::"r"(d_tmem), "l"(ah), "l"(al), "l"(wh), "l"(wl), "r"(idesc) : "memory");
}
__device__ __forceinline__ uint64_t desc64a(uint32_t saddr, uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)desc_lo(saddr); }

// GroupNorm sums of one transposed 32 x 32 chunk: this lane holds the channel quad starting at absolute channel `c` for
// 8 rows (pixv[i] < 0: row outside the image).  Rows are summed in the thread, the 4 lanes that share a quad with two
// shuffles, then channels are folded into groups of cpg (1, 2, 4, 8).  Plain stores into this warp's slot of the item
// statistics: an item touches every (warp, group) slot at most once.
__device__ __forceinline__ void stats_quad(int cpg, const float4 (&x)[8], const int (&pixv)[8], int c, float *stp, int lane) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float m = pixv[i] >= 0 ? 1.f : 0.f;
        const float4 v = make_float4(x[i].x * m, x[i].y * m, x[i].z * m, x[i].w * m);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
        s.x += __shfl_xor_sync(0xffffffffu, s.x, o); s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
        s.z += __shfl_xor_sync(0xffffffffu, s.z, o); s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
        q.x += __shfl_xor_sync(0xffffffffu, q.x, o); q.y += __shfl_xor_sync(0xffffffffu, q.y, o);
        q.z += __shfl_xor_sync(0xffffffffu, q.z, o); q.w += __shfl_xor_sync(0xffffffffu, q.w, o);
    }
    if (cpg == 1) {
        if (lane < 8) {
            float2 *d = reinterpret_cast<float2 *>(&stp[c * 2]);
            d[0] = make_float2(s.x, q.x); d[1] = make_float2(s.y, q.y); d[2] = make_float2(s.z, q.z); d[3] = make_float2(s.w, q.w);
        }
    } else if (cpg == 2) {
        if (lane < 8) {
            float2 *d = reinterpret_cast<float2 *>(&stp[(c >> 1) * 2]);
            d[0] = make_float2(s.x + s.y, q.x + q.y); d[1] = make_float2(s.z + s.w, q.z + q.w);
        }
    } else {
        float ts = (s.x + s.y) + (s.z + s.w), tq = (q.x + q.y) + (q.z + q.w);
        if (cpg == 8) { ts += __shfl_xor_sync(0xffffffffu, ts, 1); tq += __shfl_xor_sync(0xffffffffu, tq, 1); }
        if (lane < 8 && (c % cpg) == 0) *reinterpret_cast<float2 *>(&stp[(c / cpg) * 2]) = make_float2(ts, tq);
    }
}

template <int KS>
__global__ void __launch_bounds__(kThreadsHx, 1) conv_hx_kernel(const HxParams p) {
    constexpr int kTaps = KS * KS, kPitch = kTileW + KS - 1, kHaloRows = kTileH + KS - 1, kHaloPx = kHaloRows * kPitch;
    constexpr int kSlots = (kHaloPx + 15) / 16;         // halo pixels per transform thread and k-block: 12 (3x3) / 8 (1x1)
    constexpr int kPad = KS / 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t halo_stage = 2 * p.halo_plane;
    uint8_t *halo = smem;                                   // 2 stages x (hi plane | lo plane)
    uint8_t *wring = smem + 2 * halo_stage;                 // n_slots x (hi panel | lo panel)
    uint8_t *stage = wring + (size_t)p.n_slots * p.w_slot;  // 4 epilogue warps x 32 rows x (128 + 16) bytes
    HxCtl *ctl = reinterpret_cast<HxCtl *>(stage + 4 * kStageWarpBytes);
    const ConvHxArgs &a = p.a;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int nk = p.n_splits * p.k_splits;

    chore_pdl_launch_dependents();                          // the next kernel of the stream may start its own set-up
    const int tx = threadIdx.x - kFirstTxWarp * 32;         // transform thread index (0..255) or negative
    double st_s = 0.0, st_q = 0.0;
    float gam = 1.f, bet = 0.f;
    if (tx >= 0 && tx < a.Cin && a.gn_in != nullptr) { gam = __ldg(a.gamma + tx); bet = __ldg(a.beta + tx); }   // constants

    if (warp == 0) HX_STAMP(0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&ctl->halo_full[i], kTxWarps); mbar_init(&ctl->halo_empty[i], 1);
            mbar_init(&ctl->acc_full[i], 1); mbar_init(&ctl->acc_empty[i], 4);
        }
        for (int i = 0; i < kMaxSlots; ++i) { mbar_init(&ctl->w_full[i], 1); mbar_init(&ctl->w_empty[i], 1); }
        if (p.k_splits > 1) {
            // K split: the parts of one (tile, n slice) form a thread-block cluster (rank = K part) and exchange their
            // partial accumulators through distributed shared memory
            const int me = (int)(blockIdx.x % p.k_splits), n_chunks = p.Ns / 32;
            const int owned = n_chunks > me ? (n_chunks - me + p.k_splits - 1) / p.k_splits : 0;
            for (int i = 0; i < 8; ++i) mbar_init(&ctl->peer_ready[i], 1);
            mbar_init(&ctl->data_full, 1);
            // the senders' st.async stores complete the transaction count: (k_splits - 1) partials of every owned chunk
            if (owned > 0) mbar_arrive_expect_tx(&ctl->data_full, (uint32_t)((p.k_splits - 1) * owned) * 16384u);
        }
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 12 * 2 * kGroups * 2; i += kThreadsHx) (&ctl->stp[0][0][0])[i] = 0.f;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // everything above is independent of the previous kernel; from here on its outputs (activations, statistics) are read
    chore_pdl_wait();
    if (tx >= 0 && tx < a.Cin && a.gn_in != nullptr && (int)blockIdx.x < p.n_items) {
        // the GroupNorm table inputs of the first item travel while the CTA finishes its set-up
        const int b0 = ((int)blockIdx.x / nk) / tiles_per_img, g = tx / (a.Cin / kGroups);
        st_s = __ldcg(a.gn_in + (size_t)b0 * kGroups * 2 + g * 2);
        st_q = __ldcg(a.gn_in + (size_t)b0 * kGroups * 2 + g * 2 + 1);
    }
    tc_fence_before();
    __syncthreads();
    // the partners' barriers must exist before anybody signals them: arrive now, wait right before the first remote access
    if (p.k_splits > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");   // (fence.mbarrier_init released them)
    tc_fence_after();
    const uint32_t tmem_base = ctl->tmem_base;
    if (warp == 0) HX_STAMP(1);

    if (warp == 0) {
        // ---------------- weight producer ----------------
        uint32_t s = 0, ph = 0;
        const uint32_t slice = (uint32_t)p.Ns * 128u, panel = (uint32_t)a.N * 128u;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const int ks = item % p.k_splits, ns = (item / p.k_splits) % p.n_splits;
            const unsigned char *src = a.w + (size_t)(ns * p.Ns) * 128 + (size_t)(ks * p.kt_per) * 2 * panel;
            for (int kt = 0; kt < p.kt_per; ++kt, src += 2 * (size_t)panel) {
                mbar_wait(&ctl->w_empty[s], ph ^ 1u);
                if (elect_one()) {
                    uint8_t *dst = wring + (size_t)s * p.w_slot;
                    mbar_arrive_expect_tx(&ctl->w_full[s], 2 * slice);
                    bulk_g2s(dst, src, slice, &ctl->w_full[s]);
                    bulk_g2s(dst + slice, src + panel, slice, &ctl->w_full[s]);
                }
                __syncwarp();
                if (++s == (uint32_t)p.n_slots) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (converged warp, one elected lane) ----------------
        uint32_t s = 0, wph = 0, hc = 0, ic = 0;
        const uint32_t idesc = make_idesc(128, p.Ns);
        constexpr uint32_t a_hi = ((uint32_t)(kPitch * 128) >> 4) | (1u << 14) | (2u << 29);   // SBO = one halo row of pixels
        const uint64_t lo_plane = (uint64_t)(p.halo_plane >> 4), lo_panel = (uint64_t)(((uint32_t)p.Ns * 128u) >> 4);
        const uint32_t w0 = smem_u32(wring);
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++ic) {
            const int kt0 = (item % p.k_splits) * p.kt_per;
            const uint32_t abuf = ic & 1u;
            mbar_wait(&ctl->acc_empty[abuf], ((ic >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d = tmem_base + abuf * 256u;
            uint32_t first = 0;                                             // 0: the next MMA overwrites the accumulator
            int kb = kt0 / kTaps, tap = kt0 - kb * kTaps;
            int left = p.kt_per;
            while (left > 0) {
                const uint32_t hs = hc & 1u;
                mbar_wait(&ctl->halo_full[hs], (hc >> 1) & 1u);
                tc_fence_after();
                const bool full = a.Cin - kb * 64 >= 64;
                int ky = tap / KS, kx = tap - ky * KS;
                uint64_t ah = desc64a(smem_u32(halo) + hs * halo_stage + (uint32_t)((ky * kPitch + kx) * 128), a_hi);
                for (; tap < kTaps && left > 0; ++tap, --left) {
                    mbar_wait(&ctl->w_full[s], wph);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t wh = desc64a(w0 + s * p.w_slot, kDescHi);
                        if (full) umma_burst<4>(d, ah, ah + lo_plane, wh, wh + lo_panel, idesc, first);
                        else umma_burst<2>(d, ah, ah + lo_plane, wh, wh + lo_panel, idesc, first);
                        umma_commit(&ctl->w_empty[s]);
                        if (tap == kTaps - 1 || left == 1) {
                            umma_commit(&ctl->halo_empty[hs]);
                            if (left == 1) umma_commit(&ctl->acc_full[abuf]);
                        }
                    }
                    __syncwarp();
                    first = 1;
                    if (++s == (uint32_t)p.n_slots) { s = 0; wph ^= 1u; }
                    if (++kx == KS) { kx = 0; ah += (uint64_t)((kPitch - KS + 1) * 8); } else ah += 8;
                }
                if (ic == 0 && hc < 4) HX_STAMP(8 + hc);
                tap = 0; ++kb; ++hc;
            }
        }
    } else {
      if (warp >= kFirstTxWarp) {
        // ---------------- transform: fp32 halo -> relu(groupnorm) -> fp16 hi / lo planes, 128B-swizzled ----------------
        // Thread (q, p0): float4 channel chunk q of the 64-channel k-block, halo pixels p0 + 16 j.  The loads of the NEXT
        // (item, k-block) step are issued while the current one is converted: one step of latency hiding in registers.
        const int t = tx;
        const int q = t & 15, p0 = t >> 4;
        int rel[kSlots];                                    // (dy + 64) << 16 | (dx + 64) relative to the tile origin
#pragma unroll
        for (int j = 0; j < kSlots; ++j) {
            const int hp = p0 + 16 * j;
            const int hy = hp / kPitch, hx = hp - hy * kPitch;
            rel[j] = hp < kHaloPx ? (((hy - kPad) + 64) << 16) | ((hx - kPad) + 64) : -1;
        }
        const uint32_t st_off = (uint32_t)p0 * 128u + ((uint32_t)((q >> 1) ^ (p0 & 7)) << 4) + (uint32_t)(q & 1) * 8u;
        const uint32_t halo_s = smem_u32(halo), dummy_s = smem_u32(&ctl->dummy[0]);
        const float relu_floor = a.relu ? 0.f : -INFINITY;
        float4 v[kSlots];
        int off[kSlots];                                    // float offset of pixel j of the item being LOADED, -1 = zero
        uint32_t okm = 0;                                   // bit j: v[j] is an in-image pixel of the current step
        int kb_lo = 0, kb_hi = 0;                           // k-block range of the item being loaded
        auto set_item = [&](int item) -> int {              // fills off[] / kb range for `item`, returns its image index
            const int ks = item % p.k_splits, tile = item / nk;
            const int b = tile / tiles_per_img, tt = tile - b * tiles_per_img;
            const int y0 = (tt / p.tiles_x) * kTileH, x0 = (tt % p.tiles_x) * kTileW;
            kb_lo = (ks * p.kt_per) / kTaps;
            kb_hi = ((ks + 1) * p.kt_per - 1) / kTaps;
#pragma unroll
            for (int j = 0; j < kSlots; ++j) {
                const int gy = y0 + (rel[j] >> 16) - 64, gx = x0 + (rel[j] & 0xffff) - 64;
                const bool in = rel[j] >= 0 && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
                off[j] = in ? (int)((((size_t)b * a.H + gy) * a.W + gx) * a.ld_in) + a.off_in + q * 4 : -1;
            }
            return b;
        };
        uint32_t hc = 0;
        int cur_b = -1;
        int item = blockIdx.x;
        int b_item = 0;
        if (item < p.n_items) {
            b_item = set_item(item);
            const bool c_ok = kb_lo * 64 + q * 4 < a.Cin;
#pragma unroll
            for (int j = 0; j < kSlots; ++j) {
                v[j] = ldcg_pred(a.in + (off[j] >= 0 ? off[j] + kb_lo * 64 : 0), c_ok && off[j] >= 0);
                okm |= (off[j] >= 0 ? 1u : 0u) << j;
            }
        }
        if (warp == kFirstTxWarp) HX_STAMP(12);
        while (item < p.n_items) {
            if (b_item != cur_b) {
                if (cur_b >= 0) {                                 // nobody still reads the previous image's table
                    named_bar(2, kTxWarps * 32);
                    if (t < a.Cin && a.gn_in != nullptr) {
                        const int g = t / (a.Cin / kGroups);
                        st_s = __ldcg(a.gn_in + (size_t)b_item * kGroups * 2 + g * 2);
                        st_q = __ldcg(a.gn_in + (size_t)b_item * kGroups * 2 + g * 2 + 1);
                    }
                }
                if (t < a.Cin) {
                    float sc = 1.f, sh = 0.f;
                    if (a.gn_in) {
                        const double mean = st_s * p.inv_n;
                        const double var = st_q * p.inv_n - mean * mean;
                        const float rstd = 1.0f / sqrtf(fmaxf((float)var, 0.f) + 1e-5f);
                        sc = gam * rstd;
                        sh = bet - (float)mean * sc;
                    }
                    ctl->sc[t] = sc; ctl->sh[t] = sh;
                }
                if (cur_b < 0 && warp == kFirstTxWarp) HX_STAMP(13);
                named_bar(2, kTxWarps * 32);
                if (cur_b < 0 && warp == kFirstTxWarp) HX_STAMP(2);
                cur_b = b_item;
            }
            const int next_item = item + gridDim.x;
            const int my_lo = kb_lo, my_hi = kb_hi;
            for (int kb = my_lo; kb <= my_hi; ++kb, ++hc) {
                const uint32_t hs = hc & 1u;
                const int c = kb * 64 + q * 4;
                float4 scv = make_float4(1.f, 1.f, 1.f, 1.f), shv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < a.Cin) { scv = *reinterpret_cast<const float4 *>(&ctl->sc[c]); shv = *reinterpret_cast<const float4 *>(&ctl->sh[c]); }
                // what to load next: the following k-block of this item, or the first of the next item
                const bool last_kb = kb == my_hi;
                const bool more = !last_kb || next_item < p.n_items;
                const uint32_t okm_cur = okm;
                int nkb = kb + 1;
                if (last_kb && more) {
                    b_item = set_item(next_item);
                    okm = 0;
#pragma unroll
                    for (int j = 0; j < kSlots; ++j) okm |= (off[j] >= 0 ? 1u : 0u) << j;
                    nkb = kb_lo;
                }
                const bool nc_ok = more && nkb * 64 + q * 4 < a.Cin;
                mbar_wait(&ctl->halo_empty[hs], ((hc >> 1) & 1u) ^ 1u);
                if (hc == 0 && warp == kFirstTxWarp) HX_STAMP(14);
                const uint32_t hi_s = halo_s + hs * halo_stage + st_off;
#pragma unroll
                for (int j = 0; j < kSlots; ++j) {
                    float4 x = v[j];
                    const bool ok = (okm_cur >> j) & 1u;
                    x.x = fmaxf(fmaf(x.x, scv.x, shv.x), relu_floor); x.y = fmaxf(fmaf(x.y, scv.y, shv.y), relu_floor);
                    x.z = fmaxf(fmaf(x.z, scv.z, shv.z), relu_floor); x.w = fmaxf(fmaf(x.w, scv.w, shv.w), relu_floor);
                    x.x = ok ? x.x : 0.f; x.y = ok ? x.y : 0.f; x.z = ok ? x.z : 0.f; x.w = ok ? x.w : 0.f;
                    v[j] = ldcg_pred(a.in + (off[j] >= 0 ? off[j] + nkb * 64 : 0), nc_ok && off[j] >= 0);
                    uint32_t h0, l0, h1, l1;
                    split2(x.x, x.y, h0, l0);
                    split2(x.z, x.w, h1, l1);
                    const uint32_t dst = rel[j] >= 0 ? hi_s + (uint32_t)j * 2048u : dummy_s;
                    sts64(dst, h0, h1);
                    sts64(rel[j] >= 0 ? dst + p.halo_plane : dummy_s + 16u, l0, l1);
                }
                fence_proxy_async();
                __syncwarp();
                if (hc < 4 && warp == kFirstTxWarp) HX_STAMP(3 + hc);
                if (lane == 0) mbar_arrive(&ctl->halo_full[hs]);
            }
            item = next_item;
        }
      }
      {
        // ---------------- epilogue (warps 2-5: every item; transform warps 6-13: help with the CTA's last item) ----------------
        // TMEM gives every lane one pixel row of 32 channels; a warp-wide store in that layout touches 32 different
        // 128-byte lines and the L1 serialises on lines (measured: 8 float4 stores of that shape cost ~1000 cycles; the
        // 16x256b fragment shape with 8 lines per instruction is no better per byte).  Each 32 x 32 chunk is therefore
        // transposed through a padded per-warp staging tile: afterwards lane L holds the channel quad (L & 7) of the rows
        // 4 i + (L >> 3), i = 0..7, and every load / store instruction of the warp covers 4 full lines.  The per-channel
        // sums for the GroupNorm statistics fall out of that layout with two shuffles per value.
        // A CTA's last item has no successor to convert halos for, so the 8 transform warps join in (TMEM lanes are bound
        // to warp % 4: every lane quarter gets three workers, chunks are dealt round-robin; their staging tiles live in the
        // halo buffers, which are idle once the accumulator is complete).
        const bool helper = warp >= kFirstTxWarp;
        const int quarter = warp & 3;
        const int wslot = helper ? 1 + (warp - kFirstTxWarp) / 4 : 0;           // worker index inside the quarter
        const int et = threadIdx.x - kFirstEpiWarp * 32;    // 0..127 for the epilogue warps (they flush the statistics)
        const int cq = (lane & 7) * 4, rsub = lane >> 3;    // channel quad inside the chunk, row phase
        const uint32_t stg = helper ? smem_u32(halo) + (uint32_t)(warp - kFirstTxWarp) * kStageWarpBytes
                                    : smem_u32(stage) + (uint32_t)quarter * kStageWarpBytes;
        const int last_ic = (p.n_items - 1 - (int)blockIdx.x) / (int)gridDim.x;      // index of this CTA's last item
        const uint32_t stg_w = stg + (uint32_t)lane * kStagePitch, stg_r = stg + (uint32_t)rsub * kStagePitch + (uint32_t)(lane & 7) * 16u;
        float *stp_raw = &ctl->stp[quarter * 3 + wslot][0][0], *stp_out = &ctl->stp[quarter * 3 + wslot][1][0];
        const int n_chunks = p.Ns / 32;
        uint32_t ic = helper ? (uint32_t)last_ic : 0u;
        for (int item = blockIdx.x + (int)ic * (int)gridDim.x; item < p.n_items; item += gridDim.x, ++ic) {
            const bool last = (int)ic == last_ic;
            const int workers = last ? 3 : 1;
            const int ks = item % p.k_splits, tn = item / p.k_splits;        // tn = tile * n_splits + n slice
            const int tile = tn / p.n_splits, n0 = (tn - tile * p.n_splits) * p.Ns;
            const int b = tile / tiles_per_img, tt = tile - b * tiles_per_img;
            const int y0 = (tt / p.tiles_x) * kTileH, x0 = (tt % p.tiles_x) * kTileW;
            // rows of this lane: tile row r_i = quarter * 32 + 4 i + rsub -> pixel (y0 + r_i / 8, x0 + r_i % 8)
            int pixv[8];                                     // pixel index inside the tensor, or -1 outside the image
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = quarter * 32 + 4 * i + rsub;
                const int gy = y0 + (r >> 3), gx = x0 + (r & 7);
                pixv[i] = (gy < a.H && gx < a.W) ? (b * a.H + gy) * a.W + gx : -1;
            }
            const uint32_t abuf = ic & 1u;
            const uint32_t tacc = tmem_base + abuf * 256u + ((uint32_t)(quarter * 32) << 16);
            auto load_chunk = [&](int c0, float4 (&x)[8]) {   // TMEM -> staging -> transposed registers
                uint32_t u[32];
                tmem_ld32(tacc + (uint32_t)c0, u);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_w + j * 16), "r"(u[4 * j]), "r"(u[4 * j + 1]),
                                 "r"(u[4 * j + 2]), "r"(u[4 * j + 3]) : "memory");
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[i].x), "=f"(x[i].y), "=f"(x[i].z), "=f"(x[i].w)
                                 : "r"(stg_r + i * 4 * kStagePitch) : "memory");
                __syncwarp();
            };
            const bool res_on = a.res != nullptr;
            float4 rr[8];
            auto load_res = [&](int c) {                    // c = absolute first channel of this lane's quad
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    rr[i] = ldcg_pred(a.res + (size_t)(pixv[i] < 0 ? 0 : pixv[i]) * a.ld_res + a.off_res + c, pixv[i] >= 0);
            };
            // K split: chunk ci is finished by part ci % k_splits; without a split this part owns every chunk
            // ... and the chunks a part finishes are dealt to its `workers` warps of this lane quarter
            const int c_first = (ks + wslot * p.k_splits) * 32, c_step = 32 * p.k_splits * workers;
            if (res_on && c_first < p.Ns) load_res(n0 + c_first + cq);     // travels while the MMAs run
            mbar_wait(&ctl->acc_full[abuf], (ic >> 1) & 1u);
            tc_fence_after();
            if (ic == 0 && warp == kFirstEpiWarp) HX_STAMP(16);
            // K split exchange.  Every part is a receiver for the chunks it owns and a sender for the others.  A receiver
            // announces that its own MMAs are complete (its halo buffers and weight ring are idle from then on); senders
            // then push their transposed partial chunks straight into that idle shared memory of the owner (DSMEM stores)
            // and arrive on the owner's data_full barrier; the owner adds the partials from its own shared memory.
            const uint32_t recv0 = smem_u32(halo) + kRecvOff + (uint32_t)rsub * 128u + (uint32_t)(lane & 7) * 16u;
            const int owned_max = (n_chunks + p.k_splits - 1) / p.k_splits;
            if (p.k_splits > 1) {
                asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
                if (warp == kFirstEpiWarp && lane < p.k_splits && lane != ks && ks < n_chunks)
                    mbar_arrive_remote(mapa(smem_u32(&ctl->peer_ready[ks]), (uint32_t)lane));
                int parked = 0;
                for (int ci = 0; ci < n_chunks; ++ci) {
                    const int owner = ci % p.k_splits;
                    if (owner == ks) continue;
                    if ((parked++) % workers != wslot) continue;
                    float4 x[8];
                    load_chunk(ci * 32, x);
                    mbar_wait_cluster(&ctl->peer_ready[owner], 0);
                    const int sidx = ks < owner ? ks : ks - 1;
                    const uint32_t dst = mapa(recv0 + (uint32_t)(((sidx * owned_max + ci / p.k_splits) * 4 + quarter) * 4096), (uint32_t)owner);
                    const uint32_t bar = mapa(smem_u32(&ctl->data_full), (uint32_t)owner);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
                                     ::"r"(dst + i * 512), "f"(x[i].x), "f"(x[i].y), "f"(x[i].z), "f"(x[i].w), "r"(bar) : "memory");
                }
                if (ic == 0 && warp == kFirstEpiWarp) HX_STAMP(19);
                if (c_first < p.Ns) mbar_wait_cluster(&ctl->data_full, 0);
            }
            for (int c0 = c_first; c0 < p.Ns; c0 += c_step) {
                float4 x[8];
                load_chunk(c0, x);
                const int c = n0 + c0 + cq;                  // absolute first channel of this lane's quad
                const bool stamp = ic == 0 && c0 == c_first && warp == kFirstEpiWarp;
                if (stamp) HX_STAMP(20);
                for (int k2 = 0; k2 < p.k_splits - 1; ++k2) {       // partials in sender order: deterministic sums
                    const uint32_t src = recv0 + (uint32_t)(((k2 * owned_max + (c0 >> 5) / p.k_splits) * 4 + quarter) * 4096);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float4 pv;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pv.x), "=f"(pv.y), "=f"(pv.z), "=f"(pv.w)
                                     : "r"(src + i * 512) : "memory");
                        x[i].x += pv.x; x[i].y += pv.y; x[i].z += pv.z; x[i].w += pv.w;
                    }
                }
                if (a.bias) {
                    const float4 bb = __ldg(reinterpret_cast<const float4 *>(a.bias + c));
#pragma unroll
                    for (int i = 0; i < 8; ++i) { x[i].x += bb.x; x[i].y += bb.y; x[i].z += bb.z; x[i].w += bb.w; }
                }
                if (a.raw) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (pixv[i] >= 0) *reinterpret_cast<float4 *>(a.raw + (size_t)pixv[i] * a.ld_raw + a.off_raw + c) = x[i];
                }
                if (stamp) HX_STAMP(21);
                if (a.st_raw) stats_quad(a.cpg_raw, x, pixv, a.off_raw + c, stp_raw, lane);
                if (stamp) HX_STAMP(22);
                if (res_on) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { x[i].x += rr[i].x; x[i].y += rr[i].y; x[i].z += rr[i].z; x[i].w += rr[i].w; }
                    if (c0 + c_step < p.Ns) load_res(c + c_step);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (pixv[i] >= 0) *reinterpret_cast<float4 *>(a.out + (size_t)pixv[i] * a.ld_out + a.off_out + c) = x[i];
                if (stamp) HX_STAMP(23);
                if (a.st_out) stats_quad(a.cpg_out, x, pixv, a.off_out + c, stp_out, lane);
                if (stamp) HX_STAMP(25);
            }
            tc_fence_before();
            __syncwarp();
            if (ic == 0 && warp == kFirstEpiWarp) HX_STAMP(17);
            if (lane == 0 && !helper) mbar_arrive(&ctl->acc_empty[abuf]);
            if (a.st_raw || a.st_out) {
                // flush this item's statistics (they belong to image b): the worker slots summed in fixed order, fp64 atomics
                // (two barrier ids: the helpers can reach the last item's barrier while the epilogue warps are still inside
                // the previous item's, the accumulators being double buffered)
                if (last) named_bar(3, 128 * 3); else named_bar(1, 128);
                if (!helper) {
                    const int set = et >> 6, idx = et & 63;
                    double *gst = set ? a.st_out : a.st_raw;
                    double val = 0.0;
#pragma unroll
                    for (int w = 0; w < 12; ++w) {
                        // a non-last item is finished by worker 0 of every lane quarter alone; the helper slots may already hold
                        // partial sums of the CTA's LAST item (another image when B > 1), which the helpers started meanwhile
                        if (!last && (w % 3) != 0) continue;
                        val += (double)ctl->stp[w][set][idx]; ctl->stp[w][set][idx] = 0.f;
                    }
                    if (gst != nullptr && val != 0.0) atomicAdd(gst + (size_t)b * kGroups * 2 + idx, val);
                }
                if (!last) named_bar(1, 128);            // the slots are clear before the next item writes them
                if (ic == 0 && warp == kFirstEpiWarp) HX_STAMP(18);
            }
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    // K split: every remote arrival on and every store into this CTA has been waited for above (each part sends to and
    // hears from every owner), so no closing cluster barrier is needed; warps 0 / 1 complete the opening one here
    if (p.k_splits > 1 && warp < 2) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) HX_STAMP(24);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// weights (Cout, Cin, kh, kw) fp32 -> [kb][tap][hi|lo] panels of Cout rows x 64 k fp16, 128B swizzled
__global__ void __launch_bounds__(256) pack_weights_hx_kernel(const float *__restrict__ w, int cout, int cin, int kh, int kw, int kbs,
                                                              unsigned char *__restrict__ out, size_t total) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int k = (int)(i % 64);
    size_t t = i / 64;
    const int n = (int)(t % cout); t /= cout;
    const int tap = (int)(t % (kh * kw));
    const int kb = (int)(t / (kh * kw));
    const int ci = kb * 64 + k;
    float v = ci < cin ? w[(((size_t)n * cin + ci) * kh + tap / kw) * kw + tap % kw] : 0.f;
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    const __half hh = __float2half_rn(v);
    const __half ll = __float2half_rn(v - __half2float(hh));
    const size_t panel = (size_t)cout * 128;
    unsigned char *hi = out + ((size_t)(kb * kh * kw + tap) * 2) * panel, *lo = hi + panel;
    const size_t off = (size_t)n * 128 + (size_t)(((k >> 3) ^ (n & 7)) << 4) + (size_t)(k & 7) * 2;
    *reinterpret_cast<__half *>(hi + off) = hh;
    *reinterpret_cast<__half *>(lo + off) = ll;
}

}   // namespace

int conv_hx_pack_weights(chore_handle *h, const float *w, int cout, int cin, int kh, int kw, unsigned char **dev) {
    const int kbs = (cin + 63) / 64;
    const size_t panel = (size_t)cout * 128, bytes = (size_t)kh * kw * kbs * 2 * panel;
    const size_t nw = (size_t)cout * cin * kh * kw;
    float *tmp = nullptr;
    CHORE_CUDA(cudaMalloc(&tmp, nw * sizeof(float)));
    CHORE_CUDA(cudaMemcpy(tmp, w, nw * sizeof(float), cudaMemcpyHostToDevice));
    if (int rc = chore_dev_alloc(h, reinterpret_cast<void **>(dev), bytes)) return rc;
    const size_t total = (size_t)kh * kw * kbs * cout * 64;
    CHORE_LAUNCH(pack_weights_hx_kernel, (unsigned)((total + 255) / 256), 256, 0, 0, tmp, cout, cin, kh, kw, kbs, *dev, total);
    CHORE_CUDA(cudaDeviceSynchronize());
    CHORE_CUDA(cudaFree(tmp));
    return CHORE_OK;
}

// debugging aid: clock64 stamps of CTA 0 of every conv_hx launch (32 per launch, up to 256 launches) into a caller buffer
static long long *g_hx_trace = nullptr;
static int g_hx_trace_idx = 0;
extern "C" void chore_debug_hx_trace(long long *buf) { g_hx_trace = buf; g_hx_trace_idx = 0; }

int conv_hx_configure(chore_handle *h) {
    if (h->hx_configured) return CHORE_OK;
    CHORE_CUDA(cudaFuncSetAttribute(conv_hx_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    CHORE_CUDA(cudaFuncSetAttribute(conv_hx_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    h->hx_configured = true;
    return CHORE_OK;
}

static bool ksplit_enabled() {
    static const bool on = [] {
        const char *e = getenv("CHORE_B200_HX_KSPLIT");
        return !(e != nullptr && strcmp(e, "0") == 0);
    }();
    return on;
}

// work decomposition of one convolution: pixel tiles x output-channel slices x K slices (see the file header)
int conv_hx_plan(const chore_handle *h, const ConvHxArgs &a, ConvHxPlan *pl) {
    CHORE_CHECK(a.KS == 1 || a.KS == 3, "conv_hx: kernel size %d", a.KS);
    CHORE_CHECK(a.Cin % 32 == 0 && a.Cin <= 256 && a.N % 32 == 0 && a.N <= 256, "conv_hx: Cin %d / Cout %d unsupported", a.Cin, a.N);
    const int tiles = ((a.W + kTileW - 1) / kTileW) * ((a.H + kTileH - 1) / kTileH) * a.B;
    const int KT = ((a.Cin + 63) / 64) * a.KS * a.KS;
    int ks = 1;
    if (ksplit_enabled())
        for (int d = 2; d <= 6 && tiles * d <= h->sm_count; ++d)
            if (KT % d == 0 && KT / d >= 3) ks = d;
    int ns = 1;
    while (tiles * ks * ns * 2 <= h->sm_count && a.N / (ns * 2) >= 32) ns *= 2;
    pl->tiles = tiles; pl->k_splits = ks; pl->n_splits = ns;
    pl->part_floats = 0;                                   // the K parts exchange through distributed shared memory
    pl->counters = 0;
    return CHORE_OK;
}

int conv_hx_launch(chore_handle *h, const ConvHxArgs &a, const ConvHxPlan &pl, float *part, int *part_cnt, cudaStream_t st) {
    CHORE_CHECK((a.ld_in | a.off_in | a.ld_out | a.off_out | a.ld_res | a.off_res | a.ld_raw | a.off_raw) % 4 == 0, "conv_hx: unaligned channel offsets");
    CHORE_CHECK((!a.st_raw || (a.cpg_raw >= 1 && a.cpg_raw <= 8 && (a.cpg_raw & (a.cpg_raw - 1)) == 0)) &&
                (!a.st_out || (a.cpg_out >= 1 && a.cpg_out <= 8 && (a.cpg_out & (a.cpg_out - 1)) == 0)), "conv_hx: group width");
    (void)part; (void)part_cnt;
    CHORE_CHECK(pl.k_splits >= 1 && pl.k_splits <= 8 && (pl.k_splits == 1 || pl.tiles * pl.n_splits * pl.k_splits <= h->sm_count),
                "conv_hx: K split of %d parts", pl.k_splits);
    if (int rc = conv_hx_configure(h)) return rc;
    HxParams p{};
    p.a = a;
    p.tiles_x = (a.W + kTileW - 1) / kTileW;
    p.tiles_y = (a.H + kTileH - 1) / kTileH;
    p.n_splits = pl.n_splits;
    p.k_splits = pl.k_splits;
    p.Ns = a.N / pl.n_splits;
    p.n_items = pl.tiles * pl.n_splits * pl.k_splits;
    p.kblocks = (a.Cin + 63) / 64;
    p.kt_per = p.kblocks * a.KS * a.KS / pl.k_splits;
    p.halo_plane = (uint32_t)(((kTileH + a.KS - 1) * (kTileW + a.KS - 1) * 128 + 1023) / 1024 * 1024);
    p.w_slot = (uint32_t)p.Ns * 256u;
    p.inv_n = 1.0 / ((double)a.H * a.W * (a.Cin / kGroups));
    const uint32_t fixed = 1024u + 4u * p.halo_plane + 4u * kStageWarpBytes + (uint32_t)sizeof(HxCtl);
    int slots = (int)((kSmemBudget - fixed) / p.w_slot);
    p.n_slots = slots > kMaxSlots ? kMaxSlots : slots;
    CHORE_CHECK(p.n_slots >= 2, "conv_hx: weight ring does not fit (Ns %d)", p.Ns);
    const size_t smem = fixed + (size_t)p.n_slots * p.w_slot;
    int grid = p.n_items < h->sm_count ? p.n_items : h->sm_count;
    if (const char *e = getenv("CHORE_B200_HX_GRID")) {      // debugging aid: "items" = one work item per CTA (several waves)
        if (strcmp(e, "items") == 0) grid = p.n_items;
        else if (atoi(e) > 0 && atoi(e) < grid) grid = atoi(e);
    }
    p.trace = (g_hx_trace != nullptr && g_hx_trace_idx < 256) ? g_hx_trace + 32 * (g_hx_trace_idx++) : nullptr;
    if (pl.k_splits > 1) {
        // one cluster per (tile, n slice): rank = K part; the received partials live in the halo buffers + weight ring
        const int n_chunks = p.Ns / 32, owned_max = (n_chunks + pl.k_splits - 1) / pl.k_splits;
        const size_t need = kRecvOff + (size_t)(pl.k_splits - 1) * owned_max * 4 * 4096;
        CHORE_CHECK(need <= 4 * (size_t)p.halo_plane + (size_t)p.n_slots * p.w_slot && grid == p.n_items, "conv_hx: K split exchange does not fit");
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreadsHx); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)pl.k_splits; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (a.KS == 3) CHORE_CUDA(cudaLaunchKernelEx(&cfg, conv_hx_kernel<3>, p));
        else CHORE_CUDA(cudaLaunchKernelEx(&cfg, conv_hx_kernel<1>, p));
        return CHORE_OK;
    }
    if (a.KS == 3) CHORE_LAUNCH_PDL(conv_hx_kernel<3>, dim3(grid), dim3(kThreadsHx), smem, st, p);
    else CHORE_LAUNCH_PDL(conv_hx_kernel<1>, dim3(grid), dim3(kThreadsHx), smem, st, p);
    return CHORE_OK;
}
