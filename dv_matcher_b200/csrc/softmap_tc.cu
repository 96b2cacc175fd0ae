// tcgen05 candidate pass of the fused soft/hard map (sm_100a).
//
// Per CTA: a 256-row block of X (two UMMA M=128 sub-blocks, 16-bit, TMA, resident for the whole sweep)
// against a stream of 128-column tiles of Y (TMA ring); every Y tile feeds TWO tcgen05.mma chains
// (M=128, N=128, K = C + 16) whose fp32 accumulators live in TMEM (2 stages x 2 sub-blocks x 128 columns
// = all 512 columns).  Sharing the Y tile between the two sub-blocks halves the L2 -> SM operand stream
// (36 KB per 1152 tensor cycles = 31 B/clk/SM, under the ~42 B/clk/SM the L2 sustains chip-wide).
//
// The norm is folded into the GEMM: operand rows are  A = [x~, 1, 1, 1, 0..]  and  B = [-y~, h_hi, h_mid, h_lo, 0..]
// with h = |y~|^2 / 2 split into three 16-bit terms, so the accumulator IS the selection key
//     key = |y~|^2/2 - x~.y~ = (d~^2 - |x~|^2) / 2
// and the epilogue needs no FFMA / shared-memory read per entry: a min-tree over each 16-column chunk
// (0.5-1 instruction per entry) decides whether any entry of the chunk can matter (a top-16 candidate or a
// term inside the softmax window); only those chunks take the slow path.
//
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one lane),
// warps 2..17 = epilogue: TMEM lane quarter = warp % 4 (hardware rule), group = (warp - 2) / 4 selects
// (sub-block, column half).  One thread = one row x 64 columns of every tile.
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), all mbarriers; no
// CTA-wide barrier inside the sweep.
#include <cuda.h>
#include "softmap.cuh"

namespace dvm {

constexpr int TC_SUB = 128;           // rows per UMMA (M)
constexpr int TC_BM = 2 * TC_SUB;     // rows per CTA
constexpr int TC_BN = 128;            // columns per tile (UMMA N)
constexpr int TC_KBLK = 64;           // 16-bit elements per 128-byte swizzle row
constexpr int TC_KEXT = 16;           // extra K block: norm columns (one UMMA K step), 32-byte swizzle rows
constexpr int TC_EPI_WARPS = 16;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_NST = 2;             // Y ring depth
constexpr int TC_BLK_BYTES = 128 * 128;        // one 128-row x 64-element K block
constexpr int TC_EXT_BYTES = 128 * 32;         // one 128-row x 16-element K block
constexpr int TC_MAX_SPLIT = P_MAX / 2;
constexpr int TC_CHUNK = 16;               // columns per min-tree
constexpr int TC_FLUSH_AT = 4;             // flush the pending buffers when any lane holds more than this
constexpr int TC_CAP = TC_FLUSH_AT + TC_CHUNK;   // slots per lane: a chunk can append at most TC_CHUNK entries
constexpr unsigned kFull = 0xffffffffu;
constexpr int TC_PRIME_STRIDE = 16;        // priming pass: every 16th tile
constexpr int TC_PRIME_MIN_TILES = 128;     // ... when the sweep has at least this many tiles (M >= 16k)

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

// ------------------------------------------------------------------------------------------------
// operand preparation: fp32 -> 16-bit rows of pitch Ktot = Cpad + 16.
//   X rows:  [x~ (C), 0.., | 1, 1, 1, 0 x 13]                       xx[row] = |x~|^2, err_row = |x~ - x|_2
//   Y rows:  [-y~ (C), 0.., | h_hi, h_mid, h_lo, 0 x 13], h = |y~|^2/2; rows >= rows_per_b (padding up to a
//            multiple of 128): zeros with h_hi = +inf, so a padding column can never be selected.
// err_max[b] = max row rounding error of Y (certificate input), yy_max[b] = max |y~|^2.
// ------------------------------------------------------------------------------------------------
template <bool kBF16> struct Cvt16;
template <> struct Cvt16<false> {
    static __device__ __forceinline__ uint16_t bits(float v, float& back) { const __half h = __float2half_rn(v); back = __half2float(h); return __half_as_ushort(h); }
};
template <> struct Cvt16<true> {
    static __device__ __forceinline__ uint16_t bits(float v, float& back) { const __nv_bfloat16 h = __float2bfloat16_rn(v); back = __bfloat162float(h); return __bfloat16_as_ushort(h); }
};

template <bool kBF16, bool kIsY>
__global__ void __launch_bounds__(256)
tc_prep_kernel(const float* __restrict__ src, int rows_per_b, int rows_alloc, int C, int Cpad,
               uint16_t* __restrict__ dst, float* __restrict__ xx /* [B*rows_per_b], X only */,
               float* __restrict__ err_row /* X only */, float* __restrict__ err_max /* [B], Y only */,
               float* __restrict__ yy_max /* [B], Y only */) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.y;
    if (r >= rows_alloc) return;
    const int Ktot = Cpad + TC_KEXT;
    uint16_t* d = dst + ((size_t)b * rows_alloc + r) * Ktot;
    float dummy;
    if (r >= rows_per_b) {                       // Y padding row
        for (int c = lane; c < Ktot; c += 32) d[c] = (c == Cpad) ? Cvt16<kBF16>::bits(INFINITY, dummy) : (uint16_t)0;
        return;
    }
    const float* s = src + ((size_t)b * rows_per_b + r) * C;
    float n2 = 0.f, e2 = 0.f;
    for (int c = lane * 4; c < Cpad; c += 128) {             // C % 4 == 0: float4 in, 4 x 16-bit (8 bytes) out
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < C) v = __ldg(reinterpret_cast<const float4*>(s + c));
        const float in[4] = {v.x, v.y, v.z, v.w};
        uint16_t o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float x = kIsY ? -in[t] : in[t];
            float vr;
            o[t] = Cvt16<kBF16>::bits(x, vr);
            n2 = fmaf(vr, vr, n2);
            const float e = x - vr;                          // +-inf when the value overflows the 16-bit format
            e2 = fmaf(e, e, e2);
        }
        uint2 pk;
        pk.x = (uint32_t)o[0] | ((uint32_t)o[1] << 16);
        pk.y = (uint32_t)o[2] | ((uint32_t)o[3] << 16);
        *reinterpret_cast<uint2*>(d + c) = pk;
    }
    n2 = warp_sum(n2); e2 = warp_sum(e2);
    if (lane < TC_KEXT) {
        uint16_t w = 0;
        if (!kIsY) {
            if (lane < 3) w = Cvt16<kBF16>::bits(1.0f, dummy);
        } else {
            const float h = 0.5f * n2;
            float h0, h1, h2;
            const uint16_t b0 = Cvt16<kBF16>::bits(h, h0);
            const uint16_t b1 = Cvt16<kBF16>::bits(h - h0, h1);
            const uint16_t b2 = Cvt16<kBF16>::bits((h - h0) - h1, h2);
            w = lane == 0 ? b0 : lane == 1 ? b1 : lane == 2 ? b2 : (uint16_t)0;
        }
        d[Cpad + lane] = w;
    }
    if (lane == 0) {
        float e = sqrtf(e2) * 1.0001f;
        if (kIsY) {
            float hb; Cvt16<kBF16>::bits(0.5f * n2, hb);
            if (!(hb < INFINITY) || !(e < INFINITY)) e = INFINITY;          // |y|^2/2 not representable: nothing is certified
            atomicMax(reinterpret_cast<int*>(err_max + b), __float_as_int(e));   // e >= 0: int order == float order
            atomicMax(reinterpret_cast<int*>(yy_max + b), __float_as_int(n2));
        } else {
            xx[(size_t)b * rows_per_b + r] = n2;
            err_row[(size_t)b * rows_per_b + r] = e;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------------
struct TcParams {
    int N, M, KB;                // KB = Cpad / 64
    int tiles_total, tiles_per_split;
    int tile_stride;             // 1 for the sweep; > 1: the priming pass visits every tile_stride-th tile only
    int multi_split;             // column-split CTAs exchange thresholds through thr_global during the sweep
    float a2, cut_over_alpha;
    uint32_t idesc;
    const float* xx;             // [B*N]
    unsigned* thr_global;        // [B*N] per-row list threshold (true d^2 bits): written by the priming pass, refined
                                 // with atomicMin by the sweep
    CandBuffers cb;
};

// Epilogue candidate handling (per thread = one row, 64 columns of every tile):
//   * the K best (key, idx) of the row live in REGISTERS as a sorted list;
//   * per 16-column chunk a min-tree yields the chunk minimum; if it is below the row's "interesting" bound
//     thr_hi = max(list threshold, softmax-window bound) the lane appends the interesting entries of the chunk
//     to its pending buffer in shared memory (slot-major [slot][lane] -> conflict-free);
//   * when any lane holds more than TC_FLUSH_AT entries the whole warp flushes: every lane inserts ITS OWN
//     pending entries into its register list simultaneously (lane-parallel, ~10x cheaper than one divergent
//     insertion per hit); evicted / rejected entries go to the row's softmax mass;
//   * the list threshold of a row starts from the PRIMING pass (the 6th smallest key of a 1/32 column sample,
//     i.e. roughly the 150th best of the row -- no "everything is a hit" start-up phase, ~4x fewer insertions)
//     and is shared between the row's partial lists: the two column halves of a CTA via shared memory,
//     column-split CTAs via atomicMin in global memory.  Whatever the threshold was, every column a partial
//     list discarded has a key >= the list's final threshold, which is handed to finalize as the discard bound
//     `t` -- so a threshold that turns out too tight costs a rescue scan, never a wrong answer.
// Keys live in the half domain  key = (d~^2 - |x~|^2) / 2.
constexpr int KP = 6;             // list length of the priming pass

template <int K>
struct EpiState {
    TopList<K> list;
    float thr_list;        // append threshold: min(own K-th best, primed / published thresholds of the row)
    float thr_mass;        // softmax window bound in the key domain (soft mode): entries in [thr_list, thr_mass) only add mass
    float thr_hi;          // max(thr_list, thr_mass): anything below is appended to the pending buffer
    float kr, r;           // smallest key seen so far by this thread (over ALL its columns) and its distance
    float l;               // mass of the non-candidate columns relative to r
    uint32_t wr;           // write cursor into the lane's pending buffer (shared-window address, 256 B per slot)
};

__device__ __forceinline__ float ex2_approx(float x) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x)); return e; }
__device__ __forceinline__ float key_to_dist(float key, float xx) { return sqrtf(fmaxf(fmaf(2.f, key, xx), 0.f)); }

// softmax term of a non-candidate column this thread has already swept (key >= st.kr, exponent <= ~0).  These
// terms are all below the 16 exact ones, so MUFU-approximate sqrt / exp2 (2 ulp each) is ample.
__device__ __forceinline__ float epi_term(float key, float xx, float r, float a2) {
    const float x = fmaxf(fmaf(2.f, key, xx), 1e-30f);
    const float d = x * __frsqrt_rn(x);
    return ex2_approx(-a2 * (d - r));
}

// a chunk minimum below everything seen so far: move the reference point of the mass (rare: O(log M) times per
// row).  r only has to be A reference distance used consistently for l, so the MUFU approximations are fine.
template <int K>
__device__ __forceinline__ void epi_new_min(EpiState<K>& st, float m, float xx, float a2, float coa) {
    const float x = fmaxf(fmaf(2.f, m, xx), 1e-30f);
    const float rn = x * __frsqrt_rn(x);
    if (st.l != 0.f) st.l *= ex2_approx(-a2 * (st.r - rn));
    st.kr = m; st.r = rn;
    const float te = rn + coa;
    st.thr_mass = 0.5f * (te * te - xx);
    st.thr_hi = fmaxf(st.thr_list, st.thr_mass);
}

// lane-parallel flush of the pending buffers (warp-uniform trip count).  Per round every lane takes ITS next
// pending entry: a list candidate is inserted (the sorted insertion only runs in rounds where some lane has
// one), everything else -- and whatever an insertion evicts -- adds its softmax term.
template <bool kSoft, int K>
__device__ __forceinline__ void epi_flush(EpiState<K>& st, uint32_t buf_a, float xx, float a2) {
    const int cnt = (int)((st.wr - buf_a) >> 8);
    const int mx = __reduce_max_sync(kFull, cnt);
#pragma unroll 1
    for (int e = 0; e < mx; ++e) {
        const bool active = e < cnt;
        float2 kv = make_float2(INFINITY, 0.f);
        if (active) kv = lds_v2(buf_a + e * 256);
        const bool cand = kv.x < st.list.worst();
        float out = kv.x;
        if (__any_sync(kFull, cand)) {
            if (cand) out = st.list.push(kv.x, __float_as_int(kv.y));
        }
        if (kSoft && out != INFINITY) st.l += epi_term(out, xx, st.r, a2);
    }
    st.wr = buf_a;
    st.thr_list = fminf(st.thr_list, st.list.worst());
    st.thr_hi = kSoft ? fmaxf(st.thr_list, st.thr_mass) : st.thr_list;
}

__device__ __forceinline__ float min3(float a, float b, float c) { return fminf(fminf(a, b), c); }
__device__ __forceinline__ float min4(float a, float b, float c, float d) { return fminf(fminf(a, b), fminf(c, d)); }

// one 16-column chunk of one row: min-tree, slow path (append interesting entries), flush when a buffer fills up
template <bool kSoft, bool kPrime, int K>
__device__ __forceinline__ void process_chunk(EpiState<K>& st, const float (&k)[TC_CHUNK], int cbase, uint32_t buf_a,
                                              uint32_t thr_mine_a, uint32_t thr_other_a, unsigned* thr_g, bool row_ok,
                                              float xx, const TcParams& p) {
    // chunk minimum: 8 three-input min instructions
    const float m = fminf(min3(min3(k[0], k[1], k[2]), min3(k[3], k[4], k[5]), min3(k[6], k[7], k[8])),
                          min3(min3(k[9], k[10], k[11]), min3(k[12], k[13], k[14]), k[15]));
    const bool slow = m < st.thr_hi;
    if (__any_sync(kFull, slow)) {
        if (slow) {
            if (kSoft && m < st.kr) epi_new_min(st, m, xx, p.a2, p.cut_over_alpha);
            const float th = st.thr_hi;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (min4(k[q * 4], k[q * 4 + 1], k[q * 4 + 2], k[q * 4 + 3]) < th) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        if (k[q * 4 + t] < th) {       // list candidate or softmax-window term: sorted out by the flush
                            sts_v2(st.wr, k[q * 4 + t], __int_as_float(cbase + q * 4 + t));
                            st.wr += 256;
                        }
                    }
                }
            }
        }
        if (__any_sync(kFull, st.wr > buf_a + TC_FLUSH_AT * 256)) {
            epi_flush<kSoft>(st, buf_a, xx, p.a2);
            if (!kPrime && row_ok) {
                const float w = st.list.worst();
                sts_f32(thr_mine_a, w);
                const float t = lds_f32(thr_other_a);
                if (t < st.thr_list) { st.thr_list = t; st.thr_hi = kSoft ? fmaxf(t, st.thr_mass) : t; }
                if (p.multi_split && w != INFINITY) atomicMin(thr_g, __float_as_uint(fmaxf(fmaf(2.f, w, xx), 0.f)));
            }
        }
    }
}

// kPrime: priming pass -- strided tile sample, K = KP, hard mode, only output is thr_global.
template <bool kSoft, bool kPrime>
__global__ void __launch_bounds__(TC_THREADS, 1)
softmap_cand_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmXe,
                       const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmYe, const TcParams p) {
    constexpr int K = kPrime ? KP : KC;
    extern __shared__ __align__(1024) uint8_t smem[];     // swizzled operand tiles need 1024-byte alignment (checked below)
    const int unit = p.KB * TC_BLK_BYTES + TC_EXT_BYTES;  // one 128-row operand block, all of K
    uint8_t* Xs = smem;                                   // [2 sub-blocks][KB x 16 KB | 4 KB]
    uint8_t* Ys = Xs + 2 * unit;                          // [NST][KB x 16 KB | 4 KB]
    float2* cand_buf = reinterpret_cast<float2*>(Ys + TC_NST * unit);                 // [16 warps][TC_CAP][32] (key, idx)
    float* thr_sh = reinterpret_cast<float*>(cand_buf + TC_EPI_WARPS * TC_CAP * 32); // [2 halves][256 rows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(thr_sh + 2 * TC_BM);
    uint64_t* full = bars;                 // [NST]
    uint64_t* empty = bars + TC_NST;       // [NST]
    uint64_t* tfull = bars + 2 * TC_NST;   // [2]
    uint64_t* tempty = tfull + 2;          // [2]
    uint64_t* xfull = tempty + 2;          // [1]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(xfull + 1);

    const int warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);     // warp-uniform by construction: lives in uniform registers
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.z;
    const int split = blockIdx.y;
    const int row0 = blockIdx.x * TC_BM;
    const int tile0 = split * p.tiles_per_split;
    const int span = min(p.tiles_per_split, p.tiles_total - tile0);
    const int ntiles = (span + p.tile_stride - 1) / p.tile_stride;     // tiles tile0 + it * tile_stride

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (int s = 0; s < TC_NST; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, TC_EPI_WARPS); }
        mbar_init(xfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmXe);
        tma_prefetch_desc(&tmY); tma_prefetch_desc(&tmYe);
    }
    if (warp == 1) {                        // TMEM: all 512 columns (2 accumulator stages x 2 sub-blocks x 128)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            mbar_arrive_expect_tx(xfull, 2 * unit);
            for (int sb = 0; sb < 2; ++sb) {
                uint8_t* dst = Xs + sb * unit;
                for (int kb = 0; kb < p.KB; ++kb) tma_load_3d(&tmX, xfull, dst + kb * TC_BLK_BYTES, kb * TC_KBLK, row0 + sb * TC_SUB, b);
                tma_load_3d(&tmXe, xfull, dst + p.KB * TC_BLK_BYTES, 0, row0 + sb * TC_SUB, b);
            }
            for (int it = 0; it < ntiles; ++it) {
                const int s = it % TC_NST;
                const uint32_t ph = (it / TC_NST) & 1;
                mbar_wait(empty + s, ph ^ 1);
                mbar_arrive_expect_tx(full + s, unit);
                uint8_t* dst = Ys + s * unit;
                const int col0 = (tile0 + it * p.tile_stride) * TC_BN;
                for (int kb = 0; kb < p.KB; ++kb) tma_load_3d(&tmY, full + s, dst + kb * TC_BLK_BYTES, kb * TC_KBLK, col0, b);
                tma_load_3d(&tmYe, full + s, dst + p.KB * TC_BLK_BYTES, 0, col0, b);
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            mbar_wait(xfull, 0);
            for (int it = 0; it < ntiles; ++it) {
                const int s = it % TC_NST;
                const uint32_t ph = (it / TC_NST) & 1;
                const int acc = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                mbar_wait(tempty + acc, aph ^ 1);          // epilogue has drained this accumulator stage
                mbar_wait(full + s, ph);                   // Y tile landed
                tc_fence_after();
                const uint32_t ya0 = smem_u32(Ys + s * unit);
#pragma unroll
                for (int sb = 0; sb < 2; ++sb) {
                    const uint32_t d_tmem = tmem_base + (acc * 2 + sb) * TC_BN;
                    const uint32_t xa0 = smem_u32(Xs + sb * unit);
                    for (int kb = 0; kb < p.KB; ++kb) {
                        const uint32_t xa = xa0 + kb * TC_BLK_BYTES, ya = ya0 + kb * TC_BLK_BYTES;
#pragma unroll
                        for (int k = 0; k < TC_KBLK / 16; ++k)     // UMMA_K = 16 -> +32 bytes inside the swizzle row
                            tc_mma_f16(d_tmem, umma_desc_sw128(xa + k * 32), umma_desc_sw128(ya + k * 32), p.idesc, (kb | k) != 0);
                    }
                    tc_mma_f16(d_tmem, umma_desc_sw32(xa0 + p.KB * TC_BLK_BYTES), umma_desc_sw32(ya0 + p.KB * TC_BLK_BYTES), p.idesc, 1u);
                }
                tc_commit(empty + s);                      // smem slot reusable once these MMAs retire
                tc_commit(tfull + acc);                    // accumulators ready for the epilogue
            }
        }
    } else {
        // =============================== epilogue ===============================
        const int ew = warp - 2;                           // 0..15
        const int quarter = warp & 3;                      // TMEM lanes 32*quarter .. +31 are this warp's
        const int grp = ew >> 2;                           // 0..3
        const int sb = grp >> 1;                           // row sub-block
        const int half = grp & 1;                          // columns half*64 .. +63 of each tile
        const int rloc = sb * TC_SUB + quarter * 32 + lane;    // row inside the CTA block
        const int row = row0 + rloc;
        const bool row_ok = row < p.N;
        const float xx = row_ok ? __ldg(p.xx + (size_t)b * p.N + row) : 0.f;
        const uint32_t buf_a = smem_u32(cand_buf) + (uint32_t)(ew * TC_CAP * 32 + lane) * 8u;
        const uint32_t thr_mine_a = smem_u32(thr_sh) + (uint32_t)(half * TC_BM + rloc) * 4u;
        const uint32_t thr_other_a = smem_u32(thr_sh) + (uint32_t)((1 - half) * TC_BM + rloc) * 4u;
        unsigned* thr_g = p.thr_global + (size_t)b * p.N + (row_ok ? row : 0);     // true-domain d^2 bits
        sts_f32(thr_mine_a, INFINITY);
        EpiState<K> st;
        st.list.init();
        st.thr_list = -INFINITY;                           // padding rows never hit
        if (row_ok) st.thr_list = kPrime ? INFINITY : 0.5f * (__uint_as_float(__ldcg(thr_g)) - xx);
        st.thr_mass = -INFINITY;                           // no window terms until the running minimum exists
        st.thr_hi = st.thr_list;
        st.kr = INFINITY; st.r = INFINITY; st.l = 0.f; st.wr = buf_a;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + sb * TC_BN + half * 64;
#pragma unroll 1
        for (int it = 0; it < ntiles; ++it) {
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int col0 = (tile0 + it * p.tile_stride) * TC_BN + half * 64;
            // thresholds published by the row's other lists: the global one (other column splits) is fetched before
            // the wait for the accumulator and consumed after it, so its latency hides behind the MMA
            unsigned tg_bits = 0x7f800000u;
            if (!kPrime && p.multi_split && row_ok && (it & 3) == 0) tg_bits = __ldcg(thr_g);
            mbar_wait_backoff(tfull + acc, aph);
            tc_fence_after();
            if (!kPrime && row_ok) {
                const float t = fminf(lds_f32(thr_other_a), 0.5f * (__uint_as_float(tg_bits) - xx));
                if (t < st.thr_list) { st.thr_list = t; st.thr_hi = kSoft ? fmaxf(t, st.thr_mass) : t; }
            }
            const uint32_t taddr = t_lane + acc * 2 * TC_BN;
            // software-pipelined TMEM reads: chunk c+1 is in flight while chunk c is processed
            float ka[TC_CHUNK], kb[TC_CHUNK];
            tc_ld16_issue(taddr, ka);
            tc_ld16_wait(ka);
            tc_ld16_issue(taddr + TC_CHUNK, kb);
            process_chunk<kSoft, kPrime>(st, ka, col0, buf_a, thr_mine_a, thr_other_a, thr_g, row_ok, xx, p);
            tc_ld16_wait(kb);
            tc_ld16_issue(taddr + 2 * TC_CHUNK, ka);
            process_chunk<kSoft, kPrime>(st, kb, col0 + TC_CHUNK, buf_a, thr_mine_a, thr_other_a, thr_g, row_ok, xx, p);
            tc_ld16_wait(ka);
            tc_ld16_issue(taddr + 3 * TC_CHUNK, kb);
            process_chunk<kSoft, kPrime>(st, ka, col0 + 2 * TC_CHUNK, buf_a, thr_mine_a, thr_other_a, thr_g, row_ok, xx, p);
            tc_ld16_wait(kb);
            tc_fence_before();                               // all of this tile is in registers: hand the stage back
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + acc);
            process_chunk<kSoft, kPrime>(st, kb, col0 + 3 * TC_CHUNK, buf_a, thr_mine_a, thr_other_a, thr_g, row_ok, xx, p);
        }
        epi_flush<kSoft>(st, buf_a, xx, p.a2);
        if (row_ok) {
            if (kPrime) {
                const float w = st.list.worst();
                if (w != INFINITY) atomicMin(thr_g, __float_as_uint(fmaxf(fmaf(2.f, w, xx), 0.f)));
            } else {
                const size_t g_row = (size_t)b * p.N + row;
                const int pidx = split * 2 + half;
                const size_t base = (g_row * p.cb.P + pidx) * KC;
#pragma unroll
                for (int t = 0; t < K; ++t) {
                    const float k = st.list.key[t];
                    p.cb.key[base + t] = k == INFINITY ? INFINITY : fmaxf(fmaf(2.f, k, xx), 0.f);    // back to the true d^2 domain
                    p.cb.idx[base + t] = st.list.idx[t];
                }
                p.cb.l[g_row * p.cb.P + pidx] = st.l;
                p.cb.r[g_row * p.cb.P + pidx] = st.r;
                // discard bound of this partial list (true domain): everything it dropped has a key >= thr_list
                p.cb.t[g_row * p.cb.P + pidx] = st.thr_list == INFINITY ? INFINITY : fmaxf(fmaf(2.f, st.thr_list, xx), 0.f);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

__global__ void fill_u32_kernel(unsigned* p, unsigned v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// view of a [B][rows][Ktot] 16-bit operand array starting at element column k0 with `kdim` columns:
// box {box_k, 128 rows, 1}; out-of-range rows read as zero
static int make_operand_map(CUtensorMap* map, const uint16_t* base, bool bf16, int B, int rows, int Ktot, int k0, int kdim, int box_k,
                            CUtensorMapSwizzle swz) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return DVM_ERR_DEVICE; }
    cuuint64_t dims[3] = {(cuuint64_t)kdim, (cuuint64_t)rows, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)Ktot * 2, (cuuint64_t)rows * Ktot * 2};
    cuuint32_t box[3] = {(cuuint32_t)box_k, 128u, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<uint16_t*>(base + k0),
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return DVM_ERR_DEVICE; }
    return 0;
}

static int choose_split(int B, int N, int M) {
    const int row_blocks = ceil_div(N, TC_BM);
    const int tiles = ceil_div(M, TC_BN);
    int best = 1; double best_eff = -1.0;
    for (int s = 1; s <= TC_MAX_SPLIT; ++s) {
        if (s > 1 && tiles / s < 8) break;                       // keep >= 8 tiles per CTA to amortise the X load
        const int tps = ceil_div(tiles, s);
        if ((s - 1) * tps >= tiles) continue;                    // would leave an empty split
        const long long ctas = (long long)row_blocks * s * B;
        const double waves = (double)ctas / kNumSM;
        const double eff = waves / ceil(waves) - 0.02 * (s - 1); // prefer fewer partial lists on ties
        if (eff > best_eff) { best_eff = eff; best = s; }
    }
    return best;
}

int tc_num_partials(int B, int N, int M) { return 2 * choose_split(B, N, M); }

struct TcWs {
    uint16_t* Xh; uint16_t* Yh; float* xx; float* yy_max; unsigned* thr_g;
    int Cpad, Ktot, Mpad;
};

static size_t tc_ws_layout(void* base, size_t cap, int B, int N, int M, int C, TcWs* out) {
    TcWs w{};
    w.Cpad = ceil_div(C, TC_KBLK) * TC_KBLK;
    w.Ktot = w.Cpad + TC_KEXT;
    w.Mpad = ceil_div(M, TC_BN) * TC_BN;
    WsCarver ws(base, cap);
    w.Xh = ws.take<uint16_t>((size_t)B * N * w.Ktot);
    w.Yh = ws.take<uint16_t>((size_t)B * w.Mpad * w.Ktot);
    w.xx = ws.take<float>((size_t)B * N);
    w.yy_max = ws.take<float>((size_t)B);
    w.thr_g = ws.take<unsigned>((size_t)B * N);
    if (out) *out = w;
    return align_up(ws.off, 256);
}

size_t tc_workspace_bytes(int B, int N, int M, int C) { return tc_ws_layout(nullptr, 0, B, N, M, C, nullptr); }

int launch_cand_tc(const float* X, const float* Y, int B, int N, int M, int C, float alpha, bool soft, int prec,
                   CandBuffers cb, float* err_x, float* err_ymax, const float** xx_out, const float** yymax_out,
                   void* wsp, size_t ws_bytes, cudaStream_t st) {
    TcWs w;
    const size_t need = tc_ws_layout(wsp, ws_bytes, B, N, M, C, &w);
    if (!wsp || need > ws_bytes) { set_error("launch_cand_tc: workspace too small"); return DVM_ERR_WORKSPACE; }
    if (B > 65535) { set_error("launch_cand_tc: B=%d too large", B); return DVM_ERR_INVALID_ARG; }
    const bool bf16 = prec == DVM_PREC_BF16;
    if (xx_out) *xx_out = w.xx;
    if (yymax_out) *yymax_out = w.yy_max;

    DVM_CUDA(cudaMemsetAsync(err_ymax, 0, (size_t)B * sizeof(float), st));
    DVM_CUDA(cudaMemsetAsync(w.yy_max, 0, (size_t)B * sizeof(float), st));
    {
        dim3 gx(ceil_div(N, 8), B), gy(ceil_div(w.Mpad, 8), B);
        if (bf16) {
            tc_prep_kernel<true, false><<<gx, 256, 0, st>>>(X, N, N, C, w.Cpad, w.Xh, w.xx, err_x, nullptr, nullptr);
            DVM_LAUNCH_CHECK();
            tc_prep_kernel<true, true><<<gy, 256, 0, st>>>(Y, M, w.Mpad, C, w.Cpad, w.Yh, nullptr, nullptr, err_ymax, w.yy_max);
            DVM_LAUNCH_CHECK();
        } else {
            tc_prep_kernel<false, false><<<gx, 256, 0, st>>>(X, N, N, C, w.Cpad, w.Xh, w.xx, err_x, nullptr, nullptr);
            DVM_LAUNCH_CHECK();
            tc_prep_kernel<false, true><<<gy, 256, 0, st>>>(Y, M, w.Mpad, C, w.Cpad, w.Yh, nullptr, nullptr, err_ymax, w.yy_max);
            DVM_LAUNCH_CHECK();
        }
    }

    CUtensorMap tmX, tmXe, tmY, tmYe;
    int rc;
    if ((rc = make_operand_map(&tmX, w.Xh, bf16, B, N, w.Ktot, 0, w.Cpad, TC_KBLK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_operand_map(&tmXe, w.Xh, bf16, B, N, w.Ktot, w.Cpad, TC_KEXT, TC_KEXT, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
    if ((rc = make_operand_map(&tmY, w.Yh, bf16, B, w.Mpad, w.Ktot, 0, w.Cpad, TC_KBLK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_operand_map(&tmYe, w.Yh, bf16, B, w.Mpad, w.Ktot, w.Cpad, TC_KEXT, TC_KEXT, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;

    TcParams p{};
    p.N = N; p.M = M; p.KB = w.Cpad / TC_KBLK;
    p.tiles_total = ceil_div(M, TC_BN);
    const int S = cb.P / 2;
    p.tiles_per_split = ceil_div(p.tiles_total, S);
    p.a2 = alpha * kLog2e;
    // softmax window of the 16-bit pass: terms below exp(-cut) of the row maximum are dropped; the dropped mass is
    // <= M * exp(-cut) <= 1e-5 of a row sum that is >= 1
    const float cut = fminf(kExpCut, logf((float)M) + 11.6f);
    p.cut_over_alpha = alpha > 0.f ? cut / alpha : INFINITY;
    // instruction descriptor: D=f32 (bits 4-5 = 1), A/B format (0 = f16, 1 = bf16) at bits 7-9 / 10-12, K-major A and B,
    // N >> 3 at bits 17-22, M >> 4 at bits 24-28
    const uint32_t fmt = bf16 ? 1u : 0u;
    p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_SUB >> 4) << 24);
    p.xx = w.xx; p.cb = cb;
    p.thr_global = w.thr_g;
    p.multi_split = S > 1;
    p.tile_stride = 1;
    fill_u32_kernel<<<ceil_div(B * N, 256), 256, 0, st>>>(w.thr_g, 0x7f800000u, B * N);    // +inf (memset cannot write it)
    DVM_LAUNCH_CHECK();

    const size_t unit = (size_t)p.KB * TC_BLK_BYTES + TC_EXT_BYTES;
    const size_t smem = (2 + TC_NST) * unit + (size_t)TC_EPI_WARPS * TC_CAP * 32 * 8 + 2 * TC_BM * sizeof(float) + 128;
    auto kern = soft ? softmap_cand_tc_kernel<true, false> : softmap_cand_tc_kernel<false, false>;
    auto kprime = softmap_cand_tc_kernel<false, true>;
    static bool attr_done = false;
    if (!attr_done) {
        DVM_CUDA(cudaFuncSetAttribute(softmap_cand_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DVM_CUDA(cudaFuncSetAttribute(softmap_cand_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DVM_CUDA(cudaFuncSetAttribute(kprime, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = true;
    }
    if (smem > 227 * 1024) { set_error("launch_cand_tc: C=%d needs %zu bytes of shared memory", C, smem); return DVM_ERR_UNSUPPORTED; }
    prof_begin(st);
    if (p.tiles_total >= TC_PRIME_MIN_TILES) {       // priming pass over every 16th tile (6 % of the sweep's MMA work)
        TcParams pp = p;
        pp.tile_stride = TC_PRIME_STRIDE; pp.tiles_per_split = p.tiles_total; pp.multi_split = 0;
        dim3 gridp(ceil_div(N, TC_BM), 1, B);
        kprime<<<gridp, TC_THREADS, smem, st>>>(tmX, tmXe, tmY, tmYe, pp);
        DVM_LAUNCH_CHECK();
    }
    dim3 grid(ceil_div(N, TC_BM), S, B);
    kern<<<grid, TC_THREADS, smem, st>>>(tmX, tmXe, tmY, tmYe, p);
    prof_end(st);
    DVM_LAUNCH_CHECK();
    return 0;
}

}  // namespace dvm
