// tcgen05 / TMEM / mbarrier / bulk-copy PTX wrappers shared by the tensor-core kernels (sm_100a).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace tc {

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    // the suspend-time hint lets the hardware park the warp instead of burning issue slots in a spin loop
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// wait that also accumulates the cycles spent blocked (tracing builds pass a non-null counter)
__device__ __forceinline__ void mbar_wait_t(uint64_t *bar, uint32_t parity, unsigned long long *acc) {
    if (acc == nullptr) { mbar_wait(bar, parity); return; }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    *acc += (unsigned long long)(clock64() - t0);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// ---- thread-block cluster PTX ----
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory object in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// bulk async copy global -> shared, completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 B apart.
// high word: stride byte offset 1024 >> 4 | version 1 (bit 46) | SWIZZLE_128B = 2 (bits 61-63)
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
// low word: start address >> 4 (14 bits) | leading byte offset field = 1 (ignored for swizzled K-major).
// Advancing K by one 16-element step (32 B) inside the swizzle atom adds 2 to this word.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
// instruction descriptor: kind::f16, A = B = F16, D = F32, both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %5};\nmov.b64 db, {%2, %5};\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}"
                 ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi) : "memory");
}
// ---- MMA bursts --------------------------------------------------------------------------------------------------------
// The thread that issues tcgen05.mma is held until the tensor pipe takes the instruction, and whatever it executes between
// two MMAs is added to the math time (profiles/mma_microbench_r1.txt).  These helpers issue a whole group of MMAs from ONE
// asm block: 64-bit descriptors come in as operands, the per-k-step advance (+2 in the low word = 32 bytes inside the
// swizzle atom) is a single add, and there is one predicate set-up per block instead of one per instruction.
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi = kDescHi) { return ((uint64_t)hi << 32) | (uint64_t)lo; }

#define CHORE_MMA(D, A, B, I, P) "tcgen05.mma.cta_group::1.kind::f16 [" D "], " A ", " B ", " I ", " P ";\n"
// for ks < KSN: D (+)= A0[ks] * W[ks];  D += A1[ks] * W[ks]      (the hi weight panel against the hi and lo activations)
template <int KSN>
__device__ __forceinline__ void umma_burst_pair(uint32_t d, uint64_t a0, uint64_t a1, uint64_t w, uint32_t idesc, uint32_t acc_first) {
    static_assert(KSN == 1 || KSN == 4, "k steps per block");
    if (KSN == 1)
        asm volatile("{\n.reg .pred p, q;\nsetp.ne.b32 p, %5, 0;\nsetp.eq.b32 q, %4, %4;\n"
                     CHORE_MMA("%0", "%1", "%3", "%4", "p") CHORE_MMA("%0", "%2", "%3", "%4", "q")
                     "}" ::"r"(d), "l"(a0), "l"(a1), "l"(w), "r"(idesc), "r"(acc_first) : "memory");
    else
        asm volatile("{\n.reg .pred p, q;\n.reg .b64 x, y, z;\nsetp.ne.b32 p, %5, 0;\nsetp.eq.b32 q, %4, %4;\n"
                     CHORE_MMA("%0", "%1", "%3", "%4", "p") CHORE_MMA("%0", "%2", "%3", "%4", "q")
                     "add.s64 x, %1, 2;\nadd.s64 y, %2, 2;\nadd.s64 z, %3, 2;\n"
                     CHORE_MMA("%0", "x", "z", "%4", "q") CHORE_MMA("%0", "y", "z", "%4", "q")
                     "add.s64 x, %1, 4;\nadd.s64 y, %2, 4;\nadd.s64 z, %3, 4;\n"
                     CHORE_MMA("%0", "x", "z", "%4", "q") CHORE_MMA("%0", "y", "z", "%4", "q")
                     "add.s64 x, %1, 6;\nadd.s64 y, %2, 6;\nadd.s64 z, %3, 6;\n"
                     CHORE_MMA("%0", "x", "z", "%4", "q") CHORE_MMA("%0", "y", "z", "%4", "q")
                     "}" ::"r"(d), "l"(a0), "l"(a1), "l"(w), "r"(idesc), "r"(acc_first) : "memory");
}
// for ks < KSN: D += A[ks] * W[ks]      (the lo weight panel against the hi activations)
template <int KSN>
__device__ __forceinline__ void umma_burst_single(uint32_t d, uint64_t a, uint64_t w, uint32_t idesc) {
    static_assert(KSN == 1 || KSN == 4, "k steps per block");
    if (KSN == 1)
        asm volatile("{\n.reg .pred q;\nsetp.eq.b32 q, %3, %3;\n" CHORE_MMA("%0", "%1", "%2", "%3", "q")
                     "}" ::"r"(d), "l"(a), "l"(w), "r"(idesc) : "memory");
    else
        asm volatile("{\n.reg .pred q;\n.reg .b64 x, z;\nsetp.eq.b32 q, %3, %3;\n"
                     CHORE_MMA("%0", "%1", "%2", "%3", "q")
                     "add.s64 x, %1, 2;\nadd.s64 z, %2, 2;\n" CHORE_MMA("%0", "x", "z", "%3", "q")
                     "add.s64 x, %1, 4;\nadd.s64 z, %2, 4;\n" CHORE_MMA("%0", "x", "z", "%3", "q")
                     "add.s64 x, %1, 6;\nadd.s64 z, %2, 6;\n" CHORE_MMA("%0", "x", "z", "%3", "q")
                     "}" ::"r"(d), "l"(a), "l"(w), "r"(idesc) : "memory");
}
// for ks < 4: D (+)= Ah[ks] * Wh[ks];  D += Al[ks] * Wh[ks];  D += Ah[ks] * Wl[ks]      (one 64-channel block, 3-term split)
__device__ __forceinline__ void umma_burst_triple4(uint32_t d, uint64_t ah, uint64_t al, uint64_t wh, uint64_t wl, uint32_t idesc,
                                                   uint32_t acc_first) {
    asm volatile("{\n.reg .pred p, q;\n.reg .b64 a, b, c, e;\nsetp.ne.b32 p, %6, 0;\nsetp.eq.b32 q, %5, %5;\n"
                 CHORE_MMA("%0", "%1", "%3", "%5", "p") CHORE_MMA("%0", "%2", "%3", "%5", "q") CHORE_MMA("%0", "%1", "%4", "%5", "q")
                 "add.s64 a, %1, 2;\nadd.s64 b, %2, 2;\nadd.s64 c, %3, 2;\nadd.s64 e, %4, 2;\n"
                 CHORE_MMA("%0", "a", "c", "%5", "q") CHORE_MMA("%0", "b", "c", "%5", "q") CHORE_MMA("%0", "a", "e", "%5", "q")
                 "add.s64 a, %1, 4;\nadd.s64 b, %2, 4;\nadd.s64 c, %3, 4;\nadd.s64 e, %4, 4;\n"
                 CHORE_MMA("%0", "a", "c", "%5", "q") CHORE_MMA("%0", "b", "c", "%5", "q") CHORE_MMA("%0", "a", "e", "%5", "q")
                 "add.s64 a, %1, 6;\nadd.s64 b, %2, 6;\nadd.s64 c, %3, 6;\nadd.s64 e, %4, 6;\n"
                 CHORE_MMA("%0", "a", "c", "%5", "q") CHORE_MMA("%0", "b", "c", "%5", "q") CHORE_MMA("%0", "a", "e", "%5", "q")
                 "}" ::"r"(d), "l"(ah), "l"(al), "l"(wh), "l"(wl), "r"(idesc), "r"(acc_first) : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// two floats -> packed fp16x2 (a in the low half), round to nearest even, saturating to +-65504 instead of inf:
// one F2FP.SATFINITE instruction replaces the clamp + two scalar conversions + pack
__device__ __forceinline__ uint32_t cvt_f16x2_sat(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

// x = hi + lo, both fp16 (hi saturates at the fp16 range); packs two consecutive values.  For |x| <= 65504 this is
// bit-identical to clamp -> __float2half_rn.
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo) {
    hi = cvt_f16x2_sat(a, b);
    const float2 back = __half22float2(*reinterpret_cast<const __half2 *>(&hi));
    lo = cvt_f16x2_sat(a - back.x, b - back.y);
}

// same for non-negative inputs (post-ReLU activations); kept as a separate name for the call sites
__device__ __forceinline__ void split2_pos(float a, float b, uint32_t &hi, uint32_t &lo) { split2(a, b, hi, lo); }

// byte offset of 16-byte chunk `c` (8 fp16) of row `r` inside a 128B-swizzled K-major panel
__device__ __forceinline__ uint32_t sw128(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }


}   // namespace tc
