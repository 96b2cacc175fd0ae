// tcgen05 candidate pass of the fused soft/hard map (sm_100a).
//
// Per CTA PAIR (cluster of 2, tcgen05 cta_group::2): a 256-row block of X -- each CTA keeps ITS 128 rows (16-bit, TMA)
// resident for the whole sweep -- against a stream of 256-column tiles of Y, of which each CTA stages ITS 128 columns
// (3-stage TMA ring).  One tcgen05.mma.cta_group::2 chain per tile (M = 256 over the pair, N = 256, K = C + 16) with
// fp32 accumulators in TMEM (2 stages x 256 columns = all 512 columns of each CTA).  Per CTA the MMAs read 4 KB of A
// and 4 KB of B per 128 tensor cycles = 64 B/clk of shared-memory bandwidth (two single-CTA M=128 x N=128 chains read
// 128 B/clk, everything an SM has); the L2 -> SM operand stream is 36 KB per CTA and tile = 31 B/clk/SM, under the
// ~42 B/clk/SM the L2 sustains chip-wide.
//
// The norm is folded into the GEMM: operand rows are  A = [x~, 1, 1, 1, 0..]  and  B = [-y~, h_hi, h_mid, h_lo, 0..]
// with h = |y~|^2 / 2 split into three 16-bit terms, so the accumulator IS the selection key
//     key = |y~|^2/2 - x~.y~ = (d~^2 - |x~|^2) / 2
// and the epilogue needs no FFMA / shared-memory read per entry: a min-tree over each 16-column chunk
// (0.5-1 instruction per entry) decides whether any entry of the chunk can matter (a top-16 candidate or a
// term inside the softmax window); only those chunks take the slow path.
//
// Warp roles (832 threads per CTA): warp 0 = TMA producer, warp 1 = TMEM owner (+ MMA issuer, one lane, leader CTA),
// warps 2..17 = scanners: TMEM lane quarter = warp % 4 (hardware rule), column group = (warp - 2) / 4: one thread = one
// row x 64 columns of every tile; warps 18..25 = consumers (32 rows x column half of the tile each).
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> scanners), all mbarriers, signalled across the
// pair by multicast commits / remote arrives; no CTA-wide barrier inside the sweep.
#include <cuda.h>
#include <stdlib.h>
#include "softmap.cuh"
#include "tc_ptx.cuh"
#include "tc_prep.cuh"

namespace dvm {

constexpr int TC_SUB = 128;           // rows per CTA (its half of the UMMA M = 256 of the CTA pair)
constexpr int TC_BM = 2 * TC_SUB;     // rows per CTA PAIR (cluster of 2, tcgen05 cta_group::2)
constexpr int TC_BN = 256;            // columns per tile (UMMA N); each CTA of the pair stages 128 of them
constexpr int TC_SCAN_WARPS = 16;      // epilogue scanners
constexpr int TC_CONS_WARPS = 8;       // epilogue consumers (one per 32 rows x column half of the tile)
constexpr int TC_THREADS = 64 + 32 * (TC_SCAN_WARPS + TC_CONS_WARPS);
constexpr int TC_NST = 3;             // Y ring depth
constexpr int TC_BLK_BYTES = 128 * 128;        // one 128-row x 64-element K block
constexpr int TC_EXT_BYTES = 128 * 32;         // one 128-row x 16-element K block
constexpr int TC_MAX_SPLIT = 4;
constexpr int TC_CHUNK = 16;               // columns per min-tree
constexpr int TC_PRIME_STRIDE = 10;        // priming pass: every 10th tile.  Measured at 4 x 50k x 50k (prime + sweep, ms): stride 16: 3.24,
                                           // 12: 3.18, 10: 3.155, 8: 3.15; <= 6: thresholds so tight that rows run out of candidates (slow path)
constexpr int TC_PRIME_MIN_TILES = 16;      // ... when the sweep has at least this many tiles (M >= 4k): below, the sample is too small
constexpr int TC_PRIME_FULL_TILES = 96;     // ... and up to this many tiles (M <= 24k) the priming pass is exhaustive (measured, one wave of CTA
                                            // pairs, prime + sweep us: 20 tiles 176 -> 83, 40: 185 -> 109, 80: 218 -> 176, 200: 321 -> 379)
constexpr float TC_DENSE_ALPHA = 40.f;      // soft maps with alpha below this run the dense-window instance of the sweep


// ------------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------------
struct TcParams {
    int N, M, KB;                // KB = Cpad / 64
    int tiles_total, tiles_per_split;
    int tile_stride;             // 1 for the sweep; > 1: the priming pass visits every tile_stride-th tile only
    int prime_rank;              // priming pass: the row threshold is the prime_rank-th smallest chunk minimum it saw (per CTA)
    int multi_split;             // column-split CTAs exchange thresholds through thr_global during the sweep
    float a2, cut_over_alpha;
    uint32_t idesc;
    const float* xx;             // [B*N]
    unsigned* thr_global;        // [B*N] per-row list threshold (true d^2 bits): written by the priming pass, refined
                                 // with atomicMin by the sweep
    unsigned* rmin_global;       // [B*N] smallest sampled d^2 of the row (priming pass): initial softmax reference
    CandBuffers cb;
};

// Epilogue = SCANNER warps + CONSUMER warps.
//
// 16 scanner warps (TMEM lane quarter = warp % 4; (sub-block, column half) from the warp's group of four): one
// thread = one row x 64 columns of every tile, read with software-pipelined tcgen05.ld.  Per 16-column chunk a
// min-tree (8 three-input min instructions) gives the chunk minimum; if it is below the row's published bound thr_hi
// the lane copies the WHOLE chunk (16 keys, row, first column) into its warp's ring in shared memory -- ~20
// uniform instructions per warp and chunk, no per-entry work, no per-thread lists.  Scanner cost is therefore almost
// independent of the data.
//
// 8 consumer warps (one per 32 rows x column half of the tile: a row has one list per column half) drain the rings of
// their two scanner warps with ONE LANE PER ENTRY (full lane utilisation whatever the rows are): the entry's keys below
// the bound replace the worst entry of the row's K-entry list (shared memory) or add their softmax term
// exp2(-a2 (d - r)) to the row's mass; the lane then publishes the row's new bound thr_hi = max(list threshold,
// softmax-window bound).  Entries of the same row inside one batch are serialised (__match_any_sync).  Every scanner
// warp owns a single-producer ring (tail in a register, no atomic) and publishes its entries with one fence + tail
// store per tile; it waits for space only after publishing what it holds, the consumer never waits for a producer, so
// the protocol cannot deadlock.
//
// The list threshold of a row starts from the PRIMING pass (8th smallest chunk minimum of a 1/10 column sample ~ rank 80) and
// is shared between column-split CTAs via atomicMin in global memory.  Whatever the thresholds were, every column a
// list discarded has key >= the list's final threshold, which finalize receives as the discard bound `t`.
// Keys live in the half domain  key = (d~^2 - |x~|^2) / 2.
constexpr int KP = 8;                  // list length of the priming pass
constexpr int Q_CAP = 64;              // queue slots per consumer: two single-producer rings of Q_SUB slots
constexpr int Q_SUB = Q_CAP / 2;       // one ring per scanner warp feeding the consumer
constexpr int Q_ENTRY = 80;            // bytes: 16 keys | row, first column | sequence word, pad
constexpr int LIST_STRIDE = KC + 1;    // float2 per row (odd stride in 8-byte units: conflict-poor)
constexpr float LIST_EMPTY = 3.0e38f;  // "no entry" key (finite, so that a slot number can live in its low mantissa bits)

struct QCtl { unsigned head0, head1, tail0, tail1, done, pad0, pad1, pad2; };   // ring heads (consumer writes), published tails
                                                                                // (producers write), number of finished producers

__device__ __forceinline__ float ex2_approx(float x) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x)); return e; }
__device__ __forceinline__ float key_dist_approx(float key, float xx) {        // 2 ulp: fine for non-candidate terms
    const float x = fmaxf(fmaf(2.f, key, xx), 1e-30f);
    return x * __frsqrt_rn(x);
}
// softmax terms of one chunk against a FIXED reference (c0 = a2 * r0): sum of exp2(c0 - a2 d), d = sqrt(2 key + |x~|^2), with the
// MUFU reciprocal square root (2 ulp: these are non-candidate terms whose keys already carry the 16-bit operand rounding).
// Four independent MUFU chains in flight.
__device__ __forceinline__ float chunk_mass(const float (&k)[16], float xx, float c0, float a2) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 16; t += 4) {
        float x[4], r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = fmaxf(fmaf(2.f, k[t + u], xx), 1e-30f);
#pragma unroll
        for (int u = 0; u < 4; ++u) asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r[u]) : "f"(x[u]));
#pragma unroll
        for (int u = 0; u < 4; ++u) s[u] += ex2_approx(fmaf(-a2 * x[u], r[u], c0));
    }
    return (s[0] + s[1]) + (s[2] + s[3]);
}
__device__ __forceinline__ float min3(float a, float b, float c) { return fminf(fminf(a, b), c); }
__device__ __forceinline__ float min16(const float (&k)[16]) {                  // 8 three-input min instructions
    return fminf(min3(min3(k[0], k[1], k[2]), min3(k[3], k[4], k[5]), min3(k[6], k[7], k[8])),
                 min3(min3(k[9], k[10], k[11]), min3(k[12], k[13], k[14]), k[15]));
}
__device__ __forceinline__ unsigned lds_u32_volatile(uint32_t a) { unsigned r; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(r) : "r"(a) : "memory"); return r; }
__device__ __forceinline__ unsigned lds_u32_acquire(uint32_t a) { unsigned r; asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(r) : "r"(a) : "memory"); return r; }
__device__ __forceinline__ void sts_u32_release(uint32_t a, unsigned v) { asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t a) {
    float4 r; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a) : "memory"); return r;
}

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
__device__ __forceinline__ float max16(const float (&k)[16]) {
    return fmaxf(max3(max3(k[0], k[1], k[2]), max3(k[3], k[4], k[5]), max3(k[6], k[7], k[8])),
                 max3(max3(k[9], k[10], k[11]), max3(k[12], k[13], k[14]), k[15]));
}

// scanner: one 16-column chunk of one row.
// kPrivBound (hard mode): while the consumer has not published a list threshold yet (first tile of a CTA that starts
// without a primed threshold) the thread bounds it itself -- the largest key of any chunk it has seen is >= the row's
// 16th smallest key -- so the start-up does not flood the queue with every chunk of every row.
// makes the ring entries written so far by all lanes of the warp visible to the consumer: tail word at `tail_a`
__device__ __forceinline__ void ring_publish(uint32_t tail_a, unsigned tail, int lane) {
    __syncwarp();                                                    // orders the lanes' entry stores before lane 0's release
    if (lane == 0) sts_u32_release(tail_a, tail);
}

// kDense (dense-window instance, small alpha -- every chunk of every row lies inside the softmax window): pushing all those
// chunks through the queues makes the consumers the bottleneck (22-91 TFLOP/s at alpha = 10), so a chunk that is inside the
// window but holds no list candidate (chunk minimum >= the row's list bound) adds its 16 terms to the thread's PRIVATE
// accumulator (fixed reference r0 of the priming pass; merged with the consumers' mass at the end) and only chunks with
// candidates are pushed.  Any (possibly stale) pair of bounds is safe: a pushed chunk is settled completely by the consumer,
// a chunk summed here has no key below the list bound it was compared with, hence none below the current one.
struct DenseState { uint32_t thl_a; float xx, c0, a2; float l; };

template <bool kPrivBound, bool kDense>
__device__ __forceinline__ void scan_chunk(const float (&k)[TC_CHUNK], int cbase, uint32_t thr_hi_a, uint32_t q_a, uint32_t head_a,
                                           int row_in_q, int lane, float& priv, bool first_tile, unsigned& tail, unsigned& head_seen,
                                           unsigned& pub, DenseState& ds) {
    float th = lds_f32(thr_hi_a);
    if (kPrivBound) {
        th = fminf(th, priv);
        if (first_tile) priv = fminf(priv, max16(k));
    }
    const float cmin = min16(k);
    bool slow = cmin < th;
    if (kDense) {
        const bool win = slow && cmin >= lds_f32(ds.thl_a);          // inside the window, no candidate
        if (__any_sync(kFull, win)) {
            if (win) ds.l += chunk_mass(k, ds.xx, ds.c0, ds.a2);
            slow = slow && !win;
        }
    }
    const unsigned mask = __ballot_sync(kFull, slow);
    if (mask == 0u) return;                                          // warp-uniform
    // the ring has ONE producer (this warp): the tail is a warp-uniform register, no atomic; the consumer's head is
    // re-read only when the ring looks full.  Entries are PUBLISHED (one fence + the tail word) once per tile, at a
    // point where the warp has nothing in flight -- a release per entry drained the TMEM-load pipeline every time.
    const int n = __popc(mask);
    while ((int)(tail + (unsigned)n - head_seen) > Q_SUB) {
        if (pub != tail) { ring_publish(head_a + 8, tail, lane); pub = tail; }     // the consumer must see what it has to free
        head_seen = lds_u32_volatile(head_a);
        if ((int)(tail + (unsigned)n - head_seen) > Q_SUB) __nanosleep(20);
    }
    if (slow) {
        const unsigned g = tail + (unsigned)__popc(mask & ((1u << lane) - 1u));
        const uint32_t ea = q_a + (g % Q_SUB) * Q_ENTRY;
        sts_v4(ea, k[0], k[1], k[2], k[3]);
        sts_v4(ea + 16, k[4], k[5], k[6], k[7]);
        sts_v4(ea + 32, k[8], k[9], k[10], k[11]);
        sts_v4(ea + 48, k[12], k[13], k[14], k[15]);
        sts_v2(ea + 64, __int_as_float(row_in_q), __int_as_float(cbase));
    }
    tail += (unsigned)n;
    if (n >= 8) { ring_publish(head_a + 8, tail, lane); pub = tail; }   // flood (start-up, dense softmax windows): do not sit on the entries
}

// priming pass: the scanner thread keeps the KP smallest CHUNK MINIMA it has seen in a sorted register list (a branch-free
// insertion).  The KP-th smallest chunk minimum is a key with at least KP keys <= it, i.e. a valid bound for the row's KP-th
// smallest key -- and almost as tight as the KP-th smallest key itself (the few smallest keys of a row sit in distinct
// chunks) -- so the priming pass needs no queues and no consumers: its cost is the min-tree.
__device__ __forceinline__ void prime_chunk(const float (&k)[TC_CHUNK], float (&pl)[8]) {
    float c = min16(k);
#pragma unroll
    for (int t = 0; t < 8; ++t) { const float lo = fminf(pl[t], c); c = fmaxf(pl[t], c); pl[t] = lo; }
}

// kPrime: priming pass -- strided tile sample, hard mode, only outputs are thr_global / rmin_global.
template <bool kSoft, bool kPrime, bool kDense = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
softmap_cand_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmXe,
                       const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmYe, const TcParams p) {
    constexpr int K = kPrime ? KP : KC;
    extern __shared__ __align__(1024) uint8_t smem[];     // swizzled operand tiles need 1024-byte alignment (checked below)
    const int unit = p.KB * TC_BLK_BYTES + TC_EXT_BYTES;  // one 128-row operand block, all of K
    uint8_t* Xs = smem;                                   // [KB x 16 KB | 4 KB]: this CTA's 128 rows (its half of UMMA M = 256)
    uint8_t* Ys = Xs + unit;                              // [NST][KB x 16 KB | 4 KB]: this CTA's 128 columns of every tile (half of N)
    uint8_t* q_mem = Ys + TC_NST * unit;                  // [TC_CONS_WARPS][Q_CAP][Q_ENTRY]
    // lists and row state are indexed by  li = column half * 128 + CTA-local row  (a row has one list per column half of the tile)
    float2* lists = reinterpret_cast<float2*>(q_mem + TC_CONS_WARPS * Q_CAP * Q_ENTRY);   // [256][LIST_STRIDE] (key, idx)
    float* thr_hi_s = reinterpret_cast<float*>(lists + TC_BM * LIST_STRIDE);           // [256] bound read by the scanners
    float* thr_list_s = thr_hi_s + TC_BM;                 // [256] consumer-private row state from here on
    float* thr_mass_s = thr_list_s + TC_BM;
    float* kr_s = thr_mass_s + TC_BM;
    float* r_s = kr_s + TC_BM;
    float* l_s = r_s + TC_BM;
    float* xx_s = l_s + TC_BM;                            // [256] |x~|^2
    float* worst_s = xx_s + TC_BM;                        // [256] largest key of the row's list (slot number in its low bits)
    float* c0_s = worst_s + TC_BM;                        // [128] a2 * r0 per row (r0 = priming pass' sampled minimum): dense-window reference
    QCtl* qctl = reinterpret_cast<QCtl*>(c0_s + TC_SUB);     // [TC_CONS_WARPS]
    uint64_t* bars = reinterpret_cast<uint64_t*>(qctl + TC_CONS_WARPS);
    uint64_t* full = bars;                 // [NST]   leader's copy is used: expect_tx covers the TMA of BOTH CTAs
    uint64_t* empty = bars + TC_NST;       // [NST]   per CTA, signalled by the leader's multicast commit
    uint64_t* tfull = bars + 2 * TC_NST;   // [2]     per CTA, multicast commit
    uint64_t* tempty = tfull + 2;          // [2]     leader's copy: 2 x 16 scanner warps arrive
    uint64_t* xfull = tempty + 2;          // [1]     leader's copy
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(xfull + 1);

    const int warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);     // warp-uniform by construction: lives in uniform registers
    const int lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();             // 0 = leader (issues the MMAs), 1 = peer
    const int b = blockIdx.z;
    const int split = blockIdx.y;
    const int row0 = (blockIdx.x >> 1) * TC_BM + (int)crank * TC_SUB;   // first row of THIS CTA
    const int tile0 = split * p.tiles_per_split;
    const int span = min(p.tiles_per_split, p.tiles_total - tile0);
    const int ntiles = (span + p.tile_stride - 1) / p.tile_stride;     // tiles tile0 + it * tile_stride

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (int s = 0; s < TC_NST; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 2 * TC_SCAN_WARPS); }
        mbar_init(xfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmXe);
        tma_prefetch_desc(&tmY); tma_prefetch_desc(&tmYe);
        for (int c = 0; c < TC_CONS_WARPS; ++c) qctl[c] = QCtl{0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    }
    // row state, lists and queue sequence words (all threads)
    for (int e = threadIdx.x; e < TC_CONS_WARPS * Q_CAP; e += TC_THREADS) *reinterpret_cast<unsigned*>(q_mem + e * Q_ENTRY + 72) = 0u;
    for (int e = threadIdx.x; e < TC_BM * LIST_STRIDE; e += TC_THREADS)      // empty slots: LIST_EMPTY with the slot number in the low bits
        lists[e] = make_float2(__uint_as_float((__float_as_uint(LIST_EMPTY) & ~15u) | (unsigned)((e % LIST_STRIDE) & 15)), __int_as_float(-1));
    for (int rl = threadIdx.x; rl < TC_BM; rl += TC_THREADS) {        // rl = li: both column halves start from the same state
        const int row = row0 + (rl & (TC_SUB - 1));
        float thl = -INFINITY, thm = -INFINITY, kr = INFINITY, r = INFINITY, xx = 0.f;   // padding rows never enqueue
        if (row < p.N) {
            xx = __ldg(p.xx + (size_t)b * p.N + row);
            thl = kPrime ? INFINITY : 0.5f * (__uint_as_float(__ldcg(p.thr_global + (size_t)b * p.N + row)) - xx);
            if (kSoft) {
                // softmax reference from the priming pass (sample minimum >= row minimum: the window it gives is a superset)
                const float d2s = __uint_as_float(__ldcg(p.rmin_global + (size_t)b * p.N + row));
                thm = INFINITY;
                if (d2s < INFINITY) {
                    kr = 0.5f * (d2s - xx); r = sqrtf(fmaxf(d2s, 0.f));
                    const float te = r + p.cut_over_alpha;
                    thm = 0.5f * (te * te - xx);
                }
            }
        }
        thr_list_s[rl] = thl; thr_mass_s[rl] = thm; kr_s[rl] = kr; r_s[rl] = r; l_s[rl] = 0.f; xx_s[rl] = xx;
        if (rl < TC_SUB) c0_s[rl] = p.a2 * r;
        worst_s[rl] = __uint_as_float((__float_as_uint(LIST_EMPTY) & ~15u) | (unsigned)(K - 1));   // any empty slot: take the last
        thr_hi_s[rl] = kSoft ? fmaxf(thl, thm) : thl;
    }
    if (warp == 1) {                        // TMEM of the pair: 512 columns per CTA (2 accumulator stages x 256), same warp in both CTAs
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                     // both CTAs' barriers are initialised before anybody signals across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // =============================== TMA producer (both CTAs) ===============================
        if (lane == 0) {
            if (crank == 0) mbar_arrive_expect_tx(xfull, 2 * unit);                  // the pair's X rows: 2 x 128
            for (int kb = 0; kb < p.KB; ++kb) tma_load_3d_pair(&tmX, xfull, Xs + kb * TC_BLK_BYTES, kb * TC_KBLK, row0, b);
            tma_load_3d_pair(&tmXe, xfull, Xs + p.KB * TC_BLK_BYTES, 0, row0, b);
            for (int it = 0; it < ntiles; ++it) {
                const int s = it % TC_NST;
                const uint32_t ph = (it / TC_NST) & 1;
                mbar_wait(empty + s, ph ^ 1);                                       // own copy: the leader's commit reaches both CTAs
                if (crank == 0) mbar_arrive_expect_tx(full + s, 2 * unit);          // both halves of the tile
                uint8_t* dst = Ys + s * unit;
                const int col0 = (tile0 + it * p.tile_stride) * TC_BN + (int)crank * 128;   // this CTA's half of the tile's columns
                for (int kb = 0; kb < p.KB; ++kb) tma_load_3d_pair(&tmY, full + s, dst + kb * TC_BLK_BYTES, kb * TC_KBLK, col0, b);
                tma_load_3d_pair(&tmYe, full + s, dst + p.KB * TC_BLK_BYTES, 0, col0, b);
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer (leader CTA only) ===============================
        // One thread of the leader issues 9 tcgen05.mma.cta_group::2 per tile (M = 256 over the pair, N = 256, K = 16
        // each): every CTA feeds its own 128 rows of A and its own 128 columns of B from shared memory -- 8 KB per
        // 128 tensor cycles = 64 B/clk, half of what two single-CTA M = 128 x N = 128 chains read (the whole
        // shared-memory bandwidth of the SM, which limited the single-CTA version).  Descriptor halves are precomputed;
        // the loop body is two adds + the MMA.
        if (crank == 0 && lane == 0) {
            constexpr uint32_t HI128 = (1024u >> 4) | (1u << 14) | (2u << 29);     // SBO 1024 B, version 1, SWIZZLE_128B
            constexpr uint32_t HI32 = (256u >> 4) | (1u << 14) | (6u << 29);       // SBO 256 B, version 1, SWIZZLE_32B
            const uint32_t xlo = ((smem_u32(Xs) >> 4) & 0x3FFFu) | (1u << 16);
            const uint32_t ylo_s0 = ((smem_u32(Ys) >> 4) & 0x3FFFu) | (1u << 16);
            const uint32_t ystep = (uint32_t)unit >> 4;                            // stage stride in descriptor units
            const uint32_t ext = (uint32_t)(p.KB * TC_BLK_BYTES) >> 4;
            const uint32_t idesc = p.idesc;
            mbar_wait(xfull, 0);
            for (int it = 0; it < ntiles; ++it) {
                const int s = it % TC_NST;
                const uint32_t ph = (it / TC_NST) & 1;
                const int acc = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                mbar_wait(tempty + acc, aph ^ 1);          // the scanners of BOTH CTAs have drained this accumulator stage
                mbar_wait(full + s, ph);                   // both halves of the Y tile landed
                tc_fence_after();
                const uint32_t ylo = ylo_s0 + (uint32_t)s * ystep;
                const uint32_t d0 = tmem_base + (uint32_t)acc * TC_BN;
                tc_mma_f16_lohi_pair<false>(d0, xlo, ylo, HI128, idesc);           // k = 0 overwrites the accumulator
#pragma unroll
                for (int k = 1; k < TC_KBLK / 16; ++k)     // UMMA_K = 16 -> +32 bytes = +2 descriptor units inside the swizzle row
                    tc_mma_f16_lohi_pair<true>(d0, xlo + 2 * k, ylo + 2 * k, HI128, idesc);
                if (p.KB == 2) {
                    constexpr uint32_t kb1 = TC_BLK_BYTES >> 4;
#pragma unroll
                    for (int k = 0; k < TC_KBLK / 16; ++k)
                        tc_mma_f16_lohi_pair<true>(d0, xlo + kb1 + 2 * k, ylo + kb1 + 2 * k, HI128, idesc);
                }
                tc_mma_f16_lohi_pair<true>(d0, xlo + ext, ylo + ext, HI32, idesc); // norm block
                tc_commit_pair(empty + s);                 // smem slot reusable (both CTAs) once these MMAs retire
                tc_commit_pair(tfull + acc);               // accumulators ready for the scanners of both CTAs
            }
        }
    } else if (warp < 2 + TC_SCAN_WARPS) {
        // =============================== scanners ===============================
        const int ew = warp - 2;                           // 0..15
        const int quarter = warp & 3;                      // TMEM lanes 32*quarter .. +31 are this warp's
        const int cgp = ew >> 2;                           // column group: columns cgp*64 .. +63 of each 256-column tile
        const int ch = cgp >> 1;                           // column half: the list / consumer this warp feeds
        const int cq = ch * 4 + quarter;                   // consumer / queue of this warp's (rows, column half)
        const uint32_t thr_hi_a = smem_u32(thr_hi_s) + (uint32_t)(ch * TC_SUB + quarter * 32 + lane) * 4u;
        const uint32_t q_a = smem_u32(q_mem) + (uint32_t)(cq * Q_CAP + (cgp & 1) * Q_SUB) * Q_ENTRY;   // this warp's own ring
        const uint32_t ctl_a = smem_u32(qctl) + (uint32_t)cq * sizeof(QCtl);
        const uint32_t head_a = ctl_a + (uint32_t)(cgp & 1) * 4u;
        unsigned q_tail = 0, q_head_seen = 0, q_pub = 0;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + cgp * 64;
        float priv = INFINITY;
        DenseState ds;
        ds.thl_a = smem_u32(thr_list_s) + (uint32_t)(ch * TC_SUB + quarter * 32 + lane) * 4u;
        ds.xx = xx_s[quarter * 32 + lane]; ds.c0 = c0_s[quarter * 32 + lane]; ds.a2 = p.a2; ds.l = 0.f;
        float pl[KP];
#pragma unroll
        for (int t = 0; t < KP; ++t) pl[t] = INFINITY;
#pragma unroll 1
        for (int it = 0; it < ntiles; ++it) {
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int col0 = (tile0 + it * p.tile_stride) * TC_BN + cgp * 64;
            mbar_wait_backoff(tfull + acc, aph);
            tc_fence_after();
            const uint32_t taddr = t_lane + acc * TC_BN;
            // software-pipelined TMEM reads: chunk c+1 is in flight while chunk c is processed
            float ka[TC_CHUNK], kb[TC_CHUNK];
            tc_ld16_issue(taddr, ka);
            tc_ld16_wait(ka);
            tc_ld16_issue(taddr + TC_CHUNK, kb);
            if (kPrime) prime_chunk(ka, pl); else scan_chunk<!kSoft, kDense>(ka, col0, thr_hi_a, q_a, head_a, lane, lane, priv, it == 0, q_tail, q_head_seen, q_pub, ds);
            tc_ld16_wait(kb);
            tc_ld16_issue(taddr + 2 * TC_CHUNK, ka);
            if (kPrime) prime_chunk(kb, pl); else scan_chunk<!kSoft, kDense>(kb, col0 + TC_CHUNK, thr_hi_a, q_a, head_a, lane, lane, priv, it == 0, q_tail, q_head_seen, q_pub, ds);
            tc_ld16_wait(ka);
            tc_ld16_issue(taddr + 3 * TC_CHUNK, kb);
            if (kPrime) prime_chunk(ka, pl); else scan_chunk<!kSoft, kDense>(ka, col0 + 2 * TC_CHUNK, thr_hi_a, q_a, head_a, lane, lane, priv, it == 0, q_tail, q_head_seen, q_pub, ds);
            tc_ld16_wait(kb);
            tc_fence_before();                               // all of this tile is in registers: hand the stage back
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(tempty + acc);
            if (kPrime) prime_chunk(kb, pl); else scan_chunk<!kSoft, kDense>(kb, col0 + 3 * TC_CHUNK, thr_hi_a, q_a, head_a, lane, lane, priv, it == 0, q_tail, q_head_seen, q_pub, ds);
            if (q_pub != q_tail) { ring_publish(head_a + 8, q_tail, lane); q_pub = q_tail; }     // once per tile
        }
        if (kDense)                                          // all MMAs have retired: the X block is free.  [4 column groups][128 rows]
            reinterpret_cast<float*>(Xs)[cgp * TC_SUB + quarter * 32 + lane] = ds.l;
        if (kPrime) {                                        // hand the sorted list of this column half to the row's consumer
            float2* L = lists + (ch * TC_SUB + quarter * 32 + lane) * LIST_STRIDE + (cgp & 1) * KP;
#pragma unroll
            for (int t = 0; t < KP; ++t) L[t] = make_float2(pl[t], 0.f);
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            atomicAdd(reinterpret_cast<unsigned*>(__cvta_shared_to_generic(ctl_a + 16)), 1u);     // QCtl::done
        }
    } else {
        // =============================== consumers ===============================
        // Row lists are UNSORTED K-slot sets in shared memory; every stored key carries its slot number in its 4 low
        // mantissa bits, so "the worst entry and where it sits" is one max-tree.  Entry keys get their column offset
        // packed the same way, so "the best not yet handled key and its column" is one min-tree.  (16 ulp of
        // perturbation, covered by the certificate's E2 term; exact distances are recomputed by finalize anyway.)
        const int cw = warp - (2 + TC_SCAN_WARPS);          // consumer index: column half * 4 + quarter
        const int rl0 = cw * 32;                             // first list index li served (column half * 128 + quarter * 32); `rl` below is li
        const uint32_t q_a = smem_u32(q_mem) + (uint32_t)cw * Q_CAP * Q_ENTRY;
        const uint32_t ctl_a = smem_u32(qctl) + (uint32_t)cw * sizeof(QCtl);
        unsigned head = 0, head1 = 0;
        bool saw_done = false;
        if (kPrime) {                                        // nothing is queued: the consumers of column half 0 wait for the four
            if (cw < 4) {                                    // scanner warps of their rows (two per column half)
                const uint32_t ctl_b = ctl_a + 4u * (uint32_t)sizeof(QCtl);
                while (lds_u32_acquire(ctl_a + 16) != 2u || lds_u32_acquire(ctl_b + 16) != 2u) __nanosleep(200);
            }
        } else
        for (;;) {
            // lanes 0..15 take entries of ring 0, lanes 16..31 of ring 1: up to 16 published entries of each (ring order)
            const int sub = lane >> 4;
            const int n0 = min(16, (int)(lds_u32_acquire(ctl_a + 8) - head));
            const int n1 = min(16, (int)(lds_u32_acquire(ctl_a + 12) - head1));
            const unsigned g = (sub ? head1 : head) + (unsigned)(lane & 15);
            const uint32_t ea = q_a + (uint32_t)(sub * Q_SUB + (int)(g % Q_SUB)) * Q_ENTRY;
            if (n0 + n1 == 0) {
                if (saw_done) break;                                   // nothing was published before the producers finished
                if (lds_u32_acquire(ctl_a + 16) == 2u) { saw_done = true; continue; }   // look once more: entries precede `done`
                __nanosleep(32);
                continue;
            }
            const bool active = (lane & 15) < (sub ? n1 : n0);
            float k[TC_CHUNK];
            int rl = -1 - lane, cbase = 0;                            // inactive lanes: unique pseudo rows
            if (active) {
                const float4 k0 = lds_v4(ea), k1 = lds_v4(ea + 16), k2 = lds_v4(ea + 32), k3 = lds_v4(ea + 48);
                k[0] = k0.x; k[1] = k0.y; k[2] = k0.z; k[3] = k0.w; k[4] = k1.x; k[5] = k1.y; k[6] = k1.z; k[7] = k1.w;
                k[8] = k2.x; k[9] = k2.y; k[10] = k2.z; k[11] = k2.w; k[12] = k3.x; k[13] = k3.y; k[14] = k3.z; k[15] = k3.w;
                const float2 rc = lds_v2(ea + 64);
                rl = rl0 + __float_as_int(rc.x); cbase = __float_as_int(rc.y);
#pragma unroll
                for (int t = 0; t < TC_CHUNK; ++t) k[t] = __uint_as_float((__float_as_uint(k[t]) & ~15u) | (unsigned)t);
            } else {
#pragma unroll
                for (int t = 0; t < TC_CHUNK; ++t) k[t] = INFINITY;
            }
            // entries of the same row are processed one after the other, in queue order
            const unsigned peers = __match_any_sync(kFull, rl);
            bool todo = active;
            unsigned done_mask = ~__ballot_sync(kFull, active);       // inactive lanes count as done
            while (done_mask != kFull) {
                const bool mine = todo && (__ffs(peers & ~done_mask) - 1 == lane);
                // row state of the lanes whose turn it is
                float xx = 0.f, thl = -INFINITY, thm = -INFINITY, kr = 0.f, r = 0.f, l = 0.f, worst = -INFINITY;
                float2* L = lists + (mine ? rl : 0) * LIST_STRIDE;
                if (mine) {
                    xx = xx_s[rl]; thl = thr_list_s[rl]; thm = thr_mass_s[rl]; kr = kr_s[rl]; r = r_s[rl]; l = l_s[rl];
                    worst = worst_s[rl];
                    if (!kPrime && p.multi_split)
                        thl = fminf(thl, 0.5f * (__uint_as_float(__ldcg(p.thr_global + (size_t)b * p.N + row0 + (rl & (TC_SUB - 1)))) - xx));
                }
                bool changed = false;
                float prev = -INFINITY;                               // packed keys handled so far are <= prev
                // warp-uniform loop: every trip handles the next-best key of every lane that still has one below its bound
                for (;;) {
                    float m;
                    if (prev == -INFINITY) {                          // first trip of a lane: plain minimum
                        m = min16(k);
                    } else {
                        float cand[TC_CHUNK];
#pragma unroll
                        for (int t = 0; t < TC_CHUNK; ++t) cand[t] = k[t] > prev ? k[t] : INFINITY;
                        m = min16(cand);
                    }
                    if (kSoft && mine && m < kr) {                    // new row minimum: move the reference of the mass
                        const float rn = key_dist_approx(m, xx);
                        if (l != 0.f) l *= ex2_approx(-p.a2 * (r - rn));
                        kr = m; r = rn;
                        const float te = rn + p.cut_over_alpha;
                        thm = 0.5f * (te * te - xx);
                    }
                    const float lim = fminf(thl, worst);
                    const bool go = mine && m < (kSoft ? fmaxf(lim, thm) : lim);
                    if (!__any_sync(kFull, go)) break;
                    if (go) {
                        prev = m;
                        float out = m;
                        if (m < lim) {                                // list candidate: replace the worst entry, find the new worst
                            out = worst;
                            const int ws = (int)(__float_as_uint(worst) & 15u);
                            L[ws] = make_float2(__uint_as_float((__float_as_uint(m) & ~15u) | (unsigned)ws),
                                                __int_as_float(cbase + (int)(__float_as_uint(m) & 15u)));
                            float w = -INFINITY;
#pragma unroll
                            for (int t = 0; t < K; ++t) w = fmaxf(w, L[t].x);
                            worst = w;
                            changed = true;
                        }
                        if (kSoft && out < thm) l += ex2_approx(-p.a2 * (key_dist_approx(out, xx) - r));
                    }
                }
                if (mine) {
                    thl = fminf(thl, worst);
                    thr_list_s[rl] = thl; thr_mass_s[rl] = thm; kr_s[rl] = kr; r_s[rl] = r; l_s[rl] = l; worst_s[rl] = worst;
                    *reinterpret_cast<volatile float*>(thr_hi_s + rl) = kSoft ? fmaxf(thl, thm) : thl;
                    if (!kPrime && p.multi_split && changed && worst < LIST_EMPTY)
                        atomicMin(p.thr_global + (size_t)b * p.N + row0 + (rl & (TC_SUB - 1)), __float_as_uint(fmaxf(fmaf(2.f, worst, xx), 0.f)));
                    todo = false;
                }
                done_mask = __ballot_sync(kFull, !todo);
            }
            head += (unsigned)n0; head1 += (unsigned)n1;
            __syncwarp();
            if (lane == 0) { sts_u32_release(ctl_a, head); sts_u32_release(ctl_a + 4, head1); }    // QCtl::head0/1: frees the slots
        }
        // ---- results of the 32 rows of this consumer
        {
            const int rl = rl0 + lane;
            const int row = row0 + (rl & (TC_SUB - 1));
            if (row < p.N && (!kPrime || cw < 4)) {
                const float xx = xx_s[rl];
                const float2* L = lists + rl * LIST_STRIDE;
                if (kPrime) {
                    // merge the four sorted lists of chunk minima (column groups; two sit in the other half's list slot):
                    // KP-th smallest of the union, and the minimum
                    const float2* L2 = L + TC_SUB * LIST_STRIDE;
                    int i0 = 0, i1 = KP, i2 = 0, i3 = KP;
                    float w = INFINITY;
                    for (int t = 0; t < p.prime_rank; ++t) {
                        const float a0 = i0 < KP ? L[i0].x : INFINITY, a1 = i1 < 2 * KP ? L[i1].x : INFINITY;
                        const float a2 = i2 < KP ? L2[i2].x : INFINITY, a3 = i3 < 2 * KP ? L2[i3].x : INFINITY;
                        const float m01 = fminf(a0, a1), m23 = fminf(a2, a3);
                        w = fminf(m01, m23);
                        if (m01 <= m23) { if (a0 <= a1) ++i0; else ++i1; } else { if (a2 <= a3) ++i2; else ++i3; }
                    }
                    const float m = fminf(fminf(L[0].x, L[KP].x), fminf(L2[0].x, L2[KP].x));
                    if (w < LIST_EMPTY) atomicMin(p.thr_global + (size_t)b * p.N + row, __float_as_uint(fmaxf(fmaf(2.f, w, xx), 0.f)));
                    if (m < LIST_EMPTY) atomicMin(p.rmin_global + (size_t)b * p.N + row, __float_as_uint(fmaxf(fmaf(2.f, m, xx), 0.f)));
                } else {
                    const size_t g_row = (size_t)b * p.N + row;
                    const int part = split * 2 + (cw >> 2);                 // partial list index: (column split, column half)
                    const size_t base = (g_row * p.cb.P + part) * KC;
                    for (int t = 0; t < K; ++t) {
                        const float2 e = L[t];
                        const bool has = e.x < LIST_EMPTY;
                        p.cb.key[base + t] = has ? fmaxf(fmaf(2.f, e.x, xx), 0.f) : INFINITY;      // back to the true d^2 domain
                        p.cb.idx[base + t] = has ? __float_as_int(e.y) : -1;
                    }
                    const float thl = thr_list_s[rl];
                    float l_out = l_s[rl];
                    const float r_out = r_s[rl];
                    if (kDense) {
                        // mass the scanners of this column half summed themselves, reference r0 >= r_out
                        const float* lsc = reinterpret_cast<const float*>(Xs) + (cw >> 2) * 2 * TC_SUB + (rl & (TC_SUB - 1));
                        const float ls = lsc[0] + lsc[TC_SUB];
                        if (ls != 0.f) {
                            const float r0 = c0_s[rl & (TC_SUB - 1)] / p.a2;
                            l_out += ls * ex2_approx(-p.a2 * (r0 - r_out));
                        }
                    }
                    p.cb.l[g_row * p.cb.P + part] = l_out;
                    p.cb.r[g_row * p.cb.P + part] = r_out;
                    // discard bound of this list (true domain): everything it dropped has a key >= thr_list
                    p.cb.t[g_row * p.cb.P + part] = thl >= LIST_EMPTY ? INFINITY : fmaxf(fmaf(2.f, thl, xx), 0.f);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                     // nobody of the pair still signals barriers / reads shared memory of the other
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

__global__ void fill_u32_kernel(unsigned* p, unsigned v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int choose_split(int B, int N, int M) {
    const int row_blocks = ceil_div(N, TC_BM);
    const int tiles = ceil_div(M, TC_BN);
    int best = 1; double best_eff = -1.0;
    for (int s = 1; s <= TC_MAX_SPLIT; ++s) {
        if (s > 1 && tiles / s < 8) break;                       // keep >= 8 tiles per CTA to amortise the X load
        const int tps = ceil_div(tiles, s);
        if ((s - 1) * tps >= tiles) continue;                    // would leave an empty split
        const long long ctas = (long long)row_blocks * s * B;          // CTA pairs
        const double waves = (double)ctas / (kNumSM / 2);
        const double eff = waves / ceil(waves) - 0.02 * (s - 1); // prefer fewer partial lists on ties
        if (eff > best_eff) { best_eff = eff; best = s; }
    }
    return best;
}

int tc_num_partials(int B, int N, int M) { return 2 * choose_split(B, N, M); }   // (column split) x (column half of the tile)

struct TcWs {
    uint16_t* Xh; uint16_t* Yh; float* xx; float* yy_max; unsigned* thr_g; unsigned* rmin_g;
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
    w.thr_g = ws.take<unsigned>((size_t)2 * B * N);      // thr_g | rmin_g, filled with +inf by one launch
    w.rmin_g = w.thr_g + (size_t)B * N;
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
        dim3 gx(ceil_div(N, TC_PREP_ROWS), B), gy(ceil_div(w.Mpad, TC_PREP_ROWS), B);
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
    p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    p.xx = w.xx; p.cb = cb;
    p.thr_global = w.thr_g;
    p.rmin_global = w.rmin_g;
    p.multi_split = 1;                               // the two column halves of a tile are separate lists that share thresholds
    p.tile_stride = 1;
    fill_u32_kernel<<<ceil_div(2 * B * N, 256), 256, 0, st>>>(w.thr_g, 0x7f800000u, 2 * B * N);    // +inf (memset cannot write it)
    DVM_LAUNCH_CHECK();

    const size_t unit = (size_t)p.KB * TC_BLK_BYTES + TC_EXT_BYTES;
    const size_t smem = (1 + TC_NST) * unit + (size_t)TC_CONS_WARPS * Q_CAP * Q_ENTRY + (size_t)TC_BM * LIST_STRIDE * 8 + 8 * TC_BM * sizeof(float) + TC_SUB * sizeof(float)
                        + TC_CONS_WARPS * sizeof(QCtl) + 128;
    // dense-window instance for small alpha (needs the priming pass' reference): crossover measured between alpha = 10
    // (91 -> 300 TFLOP/s at 50k) and alpha = 100 (810 -> 390 when forced)
    // Priming: a sample of every 10th tile (threshold ~ rank 80-160 of the row) for long sweeps; for SHORT sweeps the consumers'
    // start-up (every key below the threshold is queued until the lists fill) costs ~150 us per CTA pair however few tiles
    // follow, so there the priming pass visits EVERY tile and hands over the 16th smallest chunk minimum -- an exact bound with
    // at least 16 keys below it, i.e. the lists start full.
    const bool prime_full = p.tiles_total <= TC_PRIME_FULL_TILES && p.tiles_total >= 2;
    const bool primed = prime_full || p.tiles_total >= TC_PRIME_MIN_TILES;
    const bool dense = soft && alpha < TC_DENSE_ALPHA && primed;
    auto kern = !soft ? softmap_cand_tc_kernel<false, false> : dense ? softmap_cand_tc_kernel<true, false, true> : softmap_cand_tc_kernel<true, false>;
    auto kprime = softmap_cand_tc_kernel<false, true>;
    static PerDeviceOnce attr_done;
    if (attr_done.need()) {
        DVM_CUDA(cudaFuncSetAttribute(softmap_cand_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DVM_CUDA(cudaFuncSetAttribute(softmap_cand_tc_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DVM_CUDA(cudaFuncSetAttribute(softmap_cand_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DVM_CUDA(cudaFuncSetAttribute(kprime, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done.done();
    }
    if (smem > 227 * 1024) { set_error("launch_cand_tc: C=%d needs %zu bytes of shared memory", C, smem); return DVM_ERR_UNSUPPORTED; }
    prof_begin(st);
    if (primed) {
        TcParams pp = p;
        pp.tile_stride = prime_full ? 1 : TC_PRIME_STRIDE; pp.prime_rank = prime_full ? 2 * KP : KP;
        pp.tiles_per_split = p.tiles_total; pp.multi_split = 0;
        dim3 gridp(2 * ceil_div(N, TC_BM), 1, B);                 // clusters of 2 CTAs along x
        kprime<<<gridp, TC_THREADS, smem, st>>>(tmX, tmXe, tmY, tmYe, pp);
        DVM_LAUNCH_CHECK();
    }
    dim3 grid(2 * ceil_div(N, TC_BM), S, B);
    kern<<<grid, TC_THREADS, smem, st>>>(tmX, tmXe, tmY, tmYe, p);
    prof_end(st);
    DVM_LAUNCH_CHECK();
    return 0;
}

}  // namespace dvm
