// PTX wrappers shared by the tcgen05 kernels (sm_100a): mbarrier, TMA, tcgen05.mma / commit / ld, UMMA descriptors.
// Forms cross-checked against CUTLASS's cute/arch/*sm100* and cutlass/arch/barrier.h.
#pragma once
#include <cuda.h>
#include <stdint.h>
#include <cuda_runtime.h>

namespace dvm {

constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------------
// PTX wrappers (forms cross-checked against CUTLASS's cute/arch/*sm100* and cutlass/arch/barrier.h)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
// wait used by the 16 epilogue warps: back off between probes so that warps that are ahead do not take issue
// slots from the warp the CTA is waiting for
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    while (!done) {
        __nanosleep(40);
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// lean form for the issue loop: descriptors as (lo, hi) halves -- hi is a per-layout constant, lo = a precomputed base
// plus a small immediate -- so one MMA costs two integer adds and the instruction itself
template <bool kAccumulate>
__device__ __forceinline__ void tc_mma_f16_lohi(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc) {
    if (kAccumulate)
        asm volatile(
            "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
            "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
            ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc) : "memory");
    else
        asm volatile(
            "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tsetp.ne.u32 p, 1, 1;\n\t"
            "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
            ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
          "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
          "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
          "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// asynchronous TMEM load of 16 columns (one row per thread) ...
__device__ __forceinline__ void tc_ld16_issue(uint32_t taddr, float (&v)[16]) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
          "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr) : "memory");
}
// ... and the wait that makes its registers valid: they are in/out operands so that no use can be scheduled
// between the load and the wait
__device__ __forceinline__ void tc_ld16_wait(float (&v)[16]) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]),
          "+r"(u[8]), "+r"(u[9]), "+r"(u[10]), "+r"(u[11]), "+r"(u[12]), "+r"(u[13]), "+r"(u[14]), "+r"(u[15])
        :: "memory");
}
// registers of a further tcgen05.ld covered by the SAME wait: a zero-instruction barrier for the compiler, placed after
// tc_ld16_wait -- asm volatile statements keep their order, and the in/out operands make every later use depend on it
__device__ __forceinline__ void tc_ld16_after_wait(float (&v)[16]) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile(""
        : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]),
          "+r"(u[8]), "+r"(u[9]), "+r"(u[10]), "+r"(u[11]), "+r"(u[12]), "+r"(u[13]), "+r"(u[14]), "+r"(u[15])
        :: "memory");
}
// shared-memory accesses by 32-bit shared-window address (one register instead of a generic pointer pair)
__device__ __forceinline__ void sts_v2(uint32_t a, float x, float y) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory"); }
__device__ __forceinline__ float2 lds_v2(uint32_t a) { float2 r; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(a) : "memory"); return r; }
__device__ __forceinline__ void sts_f32(uint32_t a, float x) { asm volatile("st.volatile.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t a) { float r; asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(r) : "r"(a) : "memory"); return r; }

// K-major operand descriptors (version 1).  SWIZZLE_128B: rows of 128 B, 8-row atoms 1024 B apart (SBO);
// SWIZZLE_32B: rows of 32 B, 8-row atoms 256 B apart.  LBO is unused for swizzled K-major layouts (= 1).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
}

// ---- CTA-pair (cta_group::2) forms.  The pair is a cluster of 2 CTAs; the leader (cluster rank 0) issues the MMAs.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the pair's leader CTA
// the TMA of either CTA lands in its OWN shared memory but signals the LEADER's barrier
__device__ __forceinline__ void tma_load_3d_pair(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {        // arrive on the leader CTA's copy of `bar`
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {            // arrives on `bar` in BOTH CTAs once the MMAs retire
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <bool kAccumulate>
__device__ __forceinline__ void tc_mma_f16_lohi_pair(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc) {
    if (kAccumulate)
        asm volatile(
            "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
            "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
            ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc) : "memory");
    else
        asm volatile(
            "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tsetp.ne.u32 p, 1, 1;\n\t"
            "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
            ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc) : "memory");
}

// cuTensorMapEncodeTiled through the runtime (no link against libcuda)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

}  // namespace dvm
