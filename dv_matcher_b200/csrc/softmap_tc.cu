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
// Warp roles (896 threads per CTA = 7 warpgroups): warpgroup 0 = warp 0 TMA producer, warp 1 TMEM owner (+ MMA issuer, one lane,
// leader CTA), warps 2-3 idle; warpgroups 1-4 (warps 4..19) = scanners: TMEM lane quarter = warp % 4 (hardware rule), column
// group = (warp - 4) / 4: one thread = one row x 64 columns of every tile; warpgroups 5-6 (warps 20..27) = consumers (32 rows x
// column half of the tile each).  Registers are redistributed after the prologue (setmaxnreg): warpgroup 0 keeps 40, the
// consumers 56, the scanners take 88 -- enough to pull their whole 64-column slice of the accumulator into registers with four
// tcgen05.ld in flight and hand the TMEM stage back BEFORE any data-dependent work (min-trees, queue pushes, ring-full waits).
// The MMA of tile t+2 therefore never waits for the slowest scanner's push path, only for the TMEM read of tile t.
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> scanners), all mbarriers, signalled across the
// pair by multicast commits / remote arrives; no CTA-wide barrier inside the sweep.
#include <cuda.h>
#include <stdlib.h>
#include "softmap.cuh"
#include "tc_ptx.cuh"

namespace dvm {

constexpr int TC_SUB = 128;           // rows per CTA (its half of the UMMA M = 256 of the CTA pair)
constexpr int TC_BM = 2 * TC_SUB;     // rows per CTA PAIR (cluster of 2, tcgen05 cta_group::2)
constexpr int TC_BN = 256;            // columns per tile (UMMA N); each CTA of the pair stages 128 of them
constexpr int TC_KBLK = 64;           // 16-bit elements per 128-byte swizzle row
constexpr int TC_KEXT = 16;           // extra K block: norm columns (one UMMA K step), 32-byte swizzle rows
constexpr int TC_SCAN_WARPS = 16;      // epilogue scanners
constexpr int TC_CONS_WARPS = 8;       // epilogue consumers (one per 32 rows x column half of the tile)
constexpr int TC_LEAD_WARPS = 4;       // warpgroup 0: TMA producer, MMA issuer, two idle warps (setmaxnreg works on whole warpgroups)
constexpr int TC_THREADS = 32 * (TC_LEAD_WARPS + TC_SCAN_WARPS + TC_CONS_WARPS);     // 896 = 7 warpgroups
constexpr int TC_REGS_LEAD = 40, TC_REGS_SCAN = 88, TC_REGS_CONS = 56;              // 128*40 + 512*88 + 256*56 = 64512 = 72*896 (the launch allocation)
constexpr int TC_NST = 3;             // Y ring depth
constexpr int TC_BLK_BYTES = 128 * 128;        // one 128-row x 64-element K block
constexpr int TC_EXT_BYTES = 128 * 32;         // one 128-row x 16-element K block
constexpr int TC_MAX_SPLIT = 4;
constexpr int TC_CHUNK = 16;               // columns per min-tree
constexpr int TC_PRIME_STRIDE = 10;        // priming pass: every 10th tile.  Measured at 4 x 50k x 50k (prime + sweep, ms): stride 16: 3.24,
                                           // 12: 3.18, 10: 3.155, 8: 3.15; <= 6: thresholds so tight that rows run out of candidates (slow path)
constexpr int TC_PRIME_MIN_TILES = 16;      // ... when the sweep has at least this many tiles (M >= 4k): below, the sample is too small
constexpr float TC_DENSE_ALPHA = 40.f;      // soft maps with alpha below this run the dense-window instance of the sweep
constexpr int TC_PREP_ROWS = 32;           // rows per block of the operand preparation (8 warps x 4 rows)     // ... when the sweep has at least this many tiles (M >= 16k)


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
    const int b = blockIdx.y;
    const int Ktot = Cpad + TC_KEXT;
    float dummy;
    float blk_err = 0.f, blk_yy = 0.f;           // Y: maxima over this warp's rows (one atomic pair per BLOCK at the end:
                                                 // a per-row atomicMax on one address serialises 200k updates)
    for (int rr = 0; rr < TC_PREP_ROWS; rr += 8) {
        const int r = blockIdx.x * TC_PREP_ROWS + rr + (threadIdx.x >> 5);
        if (r >= rows_alloc) break;
        uint16_t* d = dst + ((size_t)b * rows_alloc + r) * Ktot;
        if (r >= rows_per_b) {                       // Y padding row
            for (int c = lane; c < Ktot; c += 32) d[c] = (c == Cpad) ? Cvt16<kBF16>::bits(INFINITY, dummy) : (uint16_t)0;
            continue;
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
        {
            float e = sqrtf(e2) * 1.0001f;
            if (kIsY) {
                float hb; Cvt16<kBF16>::bits(0.5f * n2, hb);
                if (!(hb < INFINITY) || !(e < INFINITY)) e = INFINITY;          // |y|^2/2 not representable: nothing is certified
                blk_err = fmaxf(blk_err, e); blk_yy = fmaxf(blk_yy, n2);        // NaN-free: e, n2 >= 0 or +inf
            } else if (lane == 0) {
                xx[(size_t)b * rows_per_b + r] = n2;
                err_row[(size_t)b * rows_per_b + r] = e;
            }
        }
    }
    if (kIsY) {
        __shared__ float s_e[8], s_y[8];
        if (lane == 0) { s_e[threadIdx.x >> 5] = blk_err; s_y[threadIdx.x >> 5] = blk_yy; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float e = 0.f, y = 0.f;
            for (int w = 0; w < 8; ++w) { e = fmaxf(e, s_e[w]); y = fmaxf(y, s_y[w]); }
            atomicMax(reinterpret_cast<int*>(err_max + b), __float_as_int(e));   // values >= 0: int order == float order
            atomicMax(reinterpret_cast<int*>(yy_max + b), __float_as_int(y));
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
    int debug;                   // DVM_TC_DEBUG experiment switches (0 in production): 1 = scanners never push, 2 = scanners only drain
                                 // TMEM (no min-tree), 4 = accumulator wait without nanosleep back-off, 8 = consumers discard entries
    int prime_thr;               // priming pass: also publish the list threshold (0 for small problems: the sample is too small
                                 // for its 8th smallest chunk minimum to leave 16 candidates below it -- only the softmax reference)
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
// their two scanner warps with ONE LANE PER ENTRY (full lane utilisation whatever the rows are).  An entry's keys are split
// against ONE snapshot of the row's list bound `lim`:
//   * keys >= lim can never become candidates (bounds only tighten): their softmax terms exp2(c0 - a2 d) are summed by the
//     lane at once, relative to the row's CURRENT reference distance r (read with the snapshot; it only moves in the
//     serialised path below, which runs after these sums have been added): the terms of different entries of a row are
//     independent, so same-row lanes of a batch are pre-reduced with shuffles and the first of them adds the sum to the
//     row's accumulator -- no serialisation, however many entries a row with a wide softmax window pushes (such rows used
//     to serialise the whole batch: up to 16 passes of the list code per batch);
//   * keys < lim are list candidates: only these take the serialised path (same-row entries one after the other, in queue
//     order): the key replaces the worst entry of the row's K-entry list in shared memory, what it evicts (or the key itself
//     if the bound has tightened meanwhile) adds its term to the row's mass; the lane then publishes the row's new bound
//     thr_hi = max(list threshold, softmax-window bound).
// Every scanner warp owns a single-producer ring (tail in a register, no atomic) and publishes its entries with one fence +
// tail store per tile; it waits for space only after publishing what it holds, the consumer never waits for a producer, so
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
constexpr int LIST_STRIDE = KC;        // a list = 16 keys (64 B, read back as four 16-byte loads) + 16 column indices, in two arrays
constexpr float LIST_EMPTY = 3.0e38f;  // "no entry" key (finite, so that a slot number can live in its low mantissa bits)

struct QCtl { unsigned head0, head1, tail0, tail1, done, pad0, pad1, pad2; };   // ring heads (consumer writes), published tails
                                                                                // (producers write), number of finished producers

__device__ __forceinline__ float ex2_approx(float x) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x)); return e; }
__device__ __forceinline__ float key_dist_approx(float key, float xx) {        // 2 ulp: fine for non-candidate terms
    const float x = fmaxf(fmaf(2.f, key, xx), 1e-30f);
    return x * __frsqrt_rn(x);
}
// d = sqrt(2 key + |x~|^2) with the MUFU reciprocal square root (2 ulp): only used for softmax terms of NON-candidate columns,
// whose keys already carry the 16-bit operand rounding (orders of magnitude above 2 ulp)
__device__ __forceinline__ float key_dist_fast(float key, float xx) {
    const float x = fmaxf(fmaf(2.f, key, xx), 1e-30f);
    float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return x * r;
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

// makes the ring entries written so far by all lanes of the warp visible to the consumer: tail word at `tail_a`
__device__ __forceinline__ void ring_publish(uint32_t tail_a, unsigned tail, int lane) {
    __syncwarp();                                                    // orders the lanes' entry stores before lane 0's release
    if (lane == 0) sts_u32_release(tail_a, tail);
}

// A scanner warp's single-producer ring: everything lives in (warp-uniform) registers; the consumer's head is re-read only when
// the ring looks full.  Entries are PUBLISHED (one fence + the tail word) once per tile.
struct ScanRing {
    uint32_t base;           // shared-window address of slot 0
    uint32_t head_a;         // ... of the consumer's head word; the published tail sits at head_a + 8
    unsigned tail, head_seen, pub;
};

// rare: kept out of the scan loop, and by VALUE (a reference would pin the ring state to local memory)
__device__ __noinline__ uint2 ring_wait_space(uint32_t head_a, unsigned tail, unsigned pub, int n, int lane) {
    unsigned head_seen;
    for (;;) {
        if (pub != tail) { ring_publish(head_a + 8, tail, lane); pub = tail; }   // the consumer must see what it has to free
        head_seen = lds_u32_volatile(head_a);
        if ((int)(tail + (unsigned)n - head_seen) <= Q_SUB) break;
        __nanosleep(20);
    }
    return make_uint2(head_seen, pub);
}

// the lanes whose chunk can matter (`slow`) copy the WHOLE chunk (16 keys, row, first column) into the warp's ring
__device__ __forceinline__ void push_chunk(const float (&k)[TC_CHUNK], int cbase, bool slow, ScanRing& rg, int lane, unsigned lanes_below) {
    const unsigned mask = __ballot_sync(kFull, slow);
    if (mask == 0u) return;                                          // warp-uniform
    const int n = __popc(mask);
    if ((int)(rg.tail + (unsigned)n - rg.head_seen) > Q_SUB) {
        const uint2 hp = ring_wait_space(rg.head_a, rg.tail, rg.pub, n, lane);
        rg.head_seen = hp.x; rg.pub = hp.y;
    }
    if (slow) {
        const unsigned g = rg.tail + (unsigned)__popc(mask & lanes_below);
        const uint32_t ea = rg.base + (g & (unsigned)(Q_SUB - 1)) * Q_ENTRY;
        sts_v4(ea, k[0], k[1], k[2], k[3]);
        sts_v4(ea + 16, k[4], k[5], k[6], k[7]);
        sts_v4(ea + 32, k[8], k[9], k[10], k[11]);
        sts_v4(ea + 48, k[12], k[13], k[14], k[15]);
        sts_v2(ea + 64, __int_as_float(lane), __int_as_float(cbase));
    }
    rg.tail += (unsigned)n;
    if (n >= 8) { ring_publish(rg.head_a + 8, rg.tail, lane); rg.pub = rg.tail; }   // flood (start-up, dense softmax windows): do not sit on the entries
}

// softmax terms of one chunk against the fixed per-row reference (c0 = a2 * r0): exp2(c0 - a2 d), d = sqrt(2 key + |x~|^2)
__device__ __forceinline__ float chunk_mass(const float (&k)[TC_CHUNK], float xx, float c0, float a2) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < TC_CHUNK; t += 4) {                       // four independent MUFU chains in flight
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


// one tile of one scanner thread: its row x 64 columns.  Four min-trees (8 three-input min instructions each), ONE vote for
// the common "nothing below the bounds" case, then a vote + push per chunk.  thr = (list bound, softmax-window bound) of the row.
// Dense-window mode (kDense, a separate kernel instance the host selects for small alpha, where every chunk of every row lies
// inside the softmax window): pushing those chunks through the queues makes the consumers the bottleneck (22-140 TFLOP/s at
// alpha = 10), so the lanes add the 16 terms of an in-window chunk without candidates to a private accumulator (fixed
// reference r0 of the priming pass, merged with the consumers' mass at the end) and only candidate chunks are pushed.
// Measured alternatives: a per-tile decision costs ~20 instructions per warp and tile (10 % of the sweep at alpha = 100); a
// per-warp decision from the priming pass turns the mode on for most warps of a peaked 50k problem (the rank-80 list threshold
// lies inside the window there) and halves its speed -- the queues are the better path whenever they keep up.
template <bool kDense>
__device__ __forceinline__ void scan_tile(const float (&k0)[TC_CHUNK], const float (&k1)[TC_CHUNK], const float (&k2)[TC_CHUNK],
                                          const float (&k3)[TC_CHUNK], int col0, float th, float thl, ScanRing& rg, int lane, unsigned lanes_below,
                                          uint32_t xx_a, uint32_t c0_a, float a2, float& l_scan) {
    const float c_0 = min16(k0), c_1 = min16(k1), c_2 = min16(k2), c_3 = min16(k3);
    bool s0 = c_0 < th, s1 = c_1 < th, s2 = c_2 < th, s3 = c_3 < th;
    if (!__any_sync(kFull, s0 || s1 || s2 || s3)) return;
    if (kDense) {
        // in-window chunks that hold no list candidate: their 16 terms are summed here, only candidate chunks are pushed
        const bool w0 = s0 && c_0 >= thl, w1 = s1 && c_1 >= thl, w2 = s2 && c_2 >= thl, w3 = s3 && c_3 >= thl;
        const float xx = lds_f32(xx_a), c0 = lds_f32(c0_a);
        if (__any_sync(kFull, w0)) { if (w0) l_scan += chunk_mass(k0, xx, c0, a2); }
        if (__any_sync(kFull, w1)) { if (w1) l_scan += chunk_mass(k1, xx, c0, a2); }
        if (__any_sync(kFull, w2)) { if (w2) l_scan += chunk_mass(k2, xx, c0, a2); }
        if (__any_sync(kFull, w3)) { if (w3) l_scan += chunk_mass(k3, xx, c0, a2); }
        s0 = s0 && !w0; s1 = s1 && !w1; s2 = s2 && !w2; s3 = s3 && !w3;
        if (!__any_sync(kFull, s0 || s1 || s2 || s3)) return;
    }
    push_chunk(k0, col0, s0, rg, lane, lanes_below);
    push_chunk(k1, col0 + TC_CHUNK, s1, rg, lane, lanes_below);
    push_chunk(k2, col0 + 2 * TC_CHUNK, s2, rg, lane, lanes_below);
    push_chunk(k3, col0 + 3 * TC_CHUNK, s3, rg, lane, lanes_below);
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
    float* lkeys = reinterpret_cast<float*>(q_mem + TC_CONS_WARPS * Q_CAP * Q_ENTRY);   // [256][16] list keys (16-byte aligned rows)
    int* lidx = reinterpret_cast<int*>(lkeys + TC_BM * LIST_STRIDE);                    // [256][16] list column indices
    float2* thr2_s = reinterpret_cast<float2*>(lidx + TC_BM * LIST_STRIDE);             // [256] bounds read by the scanners: (max(list bound, softmax-window bound), list bound)
    float* thr_list_s = reinterpret_cast<float*>(thr2_s + TC_BM);                       // [256] consumer-private row state from here on
    float* thr_mass_s = thr_list_s + TC_BM;
    float* kr_s = thr_mass_s + TC_BM;                     // smallest key seen (tightens the softmax window)
    float* r_s = kr_s + TC_BM;                            // reference distance of the row's mass: starts at the priming pass' sampled minimum
    float* l_s = r_s + TC_BM;                             // sum of exp2(-a2 (d - r)) over the non-candidate columns
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
    for (int e = threadIdx.x; e < TC_BM * LIST_STRIDE; e += TC_THREADS) {    // empty slots: LIST_EMPTY with the slot number in the low bits
        lkeys[e] = __uint_as_float((__float_as_uint(LIST_EMPTY) & ~15u) | (unsigned)(e & 15));
        lidx[e] = -1;
    }
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
        thr2_s[rl] = (p.debug & 1) ? make_float2(-INFINITY, -INFINITY) : make_float2(kSoft ? fmaxf(thl, thm) : thl, thl);
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

    // register redistribution: every role branch starts with the setmaxnreg of its warpgroup (whole warpgroups execute the
    // same instruction; ptxas allocates the code of a branch against the budget its setmaxnreg sets)
    if (warp < TC_LEAD_WARPS) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_LEAD));
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
      }   // warps 2, 3: nothing to do (they only hold the place of a whole warpgroup for setmaxnreg)
    } else if (warp < TC_LEAD_WARPS + TC_SCAN_WARPS) {
        // =============================== scanners ===============================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REGS_SCAN));
        const int ew = warp - TC_LEAD_WARPS;               // 0..15
        const int quarter = warp & 3;                      // TMEM lanes 32*quarter .. +31 are this warp's (TC_LEAD_WARPS % 4 == 0)
        const int cgp = ew >> 2;                           // column group: columns cgp*64 .. +63 of each 256-column tile
        const int ch = cgp >> 1;                           // column half: the list / consumer this warp feeds
        const int cq = ch * 4 + quarter;                   // consumer / queue of this warp's (rows, column half)
        const uint32_t thr2_a = smem_u32(thr2_s) + (uint32_t)(ch * TC_SUB + quarter * 32 + lane) * 8u;
        const uint32_t ctl_a = smem_u32(qctl) + (uint32_t)cq * sizeof(QCtl);
        ScanRing rg;
        rg.base = smem_u32(q_mem) + (uint32_t)(cq * Q_CAP + (cgp & 1) * Q_SUB) * Q_ENTRY;   // this warp's own ring
        rg.head_a = ctl_a + (uint32_t)(cgp & 1) * 4u;
        rg.tail = rg.head_seen = rg.pub = 0u;
        const unsigned lanes_below = (1u << lane) - 1u;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + cgp * 64;
        float pl[KP];
#pragma unroll
        for (int t = 0; t < KP; ++t) pl[t] = INFINITY;
        int col0 = tile0 * TC_BN + cgp * 64;
        const int col_step = p.tile_stride * TC_BN;
        uint32_t aph = 0;
        // dense-window mode: private softmax accumulator of this thread's (row, column group), reference r0 (priming pass)
        float l_scan = 0.f;
        const uint32_t sc_xx_a = smem_u32(xx_s + quarter * 32 + lane), sc_c0_a = smem_u32(c0_s + quarter * 32 + lane);

        // one tile: wait for the accumulator stage, pull this thread's 64 columns into registers (four TMEM loads in flight,
        // one wait), hand the stage back at once, then scan
        auto tile = [&](const int acc) {
            if (p.debug & 4) mbar_wait(tfull + acc, aph); else mbar_wait_backoff(tfull + acc, aph);
            tc_fence_after();
            const uint32_t taddr = t_lane + acc * TC_BN;
            float k0[TC_CHUNK], k1[TC_CHUNK], k2[TC_CHUNK], k3[TC_CHUNK];
            tc_ld16_issue(taddr, k0);
            tc_ld16_issue(taddr + TC_CHUNK, k1);
            tc_ld16_issue(taddr + 2 * TC_CHUNK, k2);
            tc_ld16_issue(taddr + 3 * TC_CHUNK, k3);
            float th = 0.f, thl = 0.f;                               // the row's published bounds: one read per tile
            if (!kPrime) {
                if (kDense) { const float2 t2 = lds_v2(thr2_a); th = t2.x; thl = t2.y; }
                else th = lds_f32(thr2_a);
            }
            tc_ld16_wait(k0);
            tc_ld16_after_wait(k1); tc_ld16_after_wait(k2); tc_ld16_after_wait(k3);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(tempty + acc);
            if (kPrime) {
                prime_chunk(k0, pl); prime_chunk(k1, pl); prime_chunk(k2, pl); prime_chunk(k3, pl);
            } else if (p.debug & 2) {
                if (k0[0] + k1[1] + k2[2] + k3[3] == 12345.678f) pl[0] = 0.f;      // keep the loads alive
            } else {
                scan_tile<kDense>(k0, k1, k2, k3, col0, th, thl, rg, lane, lanes_below, sc_xx_a, sc_c0_a, p.a2, l_scan);
                if (rg.pub != rg.tail) { ring_publish(rg.head_a + 8, rg.tail, lane); rg.pub = rg.tail; }     // once per tile
            }
            col0 += col_step;
        };
#pragma unroll 1
        for (int it = 0; it + 1 < ntiles; it += 2) { tile(0); tile(1); aph ^= 1u; }
        if (ntiles & 1) tile(0);
        if (kDense)                                          // all MMAs have retired: the X block is free.  [4 column groups][128 rows]
            reinterpret_cast<float*>(Xs)[cgp * TC_SUB + quarter * 32 + lane] = l_scan;
        if (kPrime) {                                        // hand the sorted list of this column half to the row's consumer
            float* L = lkeys + (ch * TC_SUB + quarter * 32 + lane) * LIST_STRIDE + (cgp & 1) * KP;
#pragma unroll
            for (int t = 0; t < KP; ++t) L[t] = pl[t];
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            atomicAdd(reinterpret_cast<unsigned*>(__cvta_shared_to_generic(ctl_a + 16)), 1u);     // QCtl::done
        }
    } else {
        // =============================== consumers ===============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_CONS));
        // Row lists are UNSORTED K-slot sets in shared memory; every stored key carries its slot number in its 4 low
        // mantissa bits, so "the worst entry and where it sits" is one max-tree.  Entry keys get their column offset
        // packed the same way, so "the best not yet handled key and its column" is one min-tree.  (16 ulp of
        // perturbation, covered by the certificate's E2 term; exact distances are recomputed by finalize anyway.)
        const int cw = warp - (TC_LEAD_WARPS + TC_SCAN_WARPS);   // consumer index: column half * 4 + quarter
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
            const bool active = (lane & 15) < (sub ? n1 : n0) && !(p.debug & 8);
            float k[TC_CHUNK];
            int rl = -1 - lane, cbase = 0;                            // inactive lanes: unique pseudo rows
            float lim = -INFINITY;                                    // ONE snapshot of the row's list bound per entry
            float mass = 0.f, thm_e = -INFINITY;
            bool need_mass = false;
            if (active) {
                const float4 k0 = lds_v4(ea), k1 = lds_v4(ea + 16), k2 = lds_v4(ea + 32), k3 = lds_v4(ea + 48);
                k[0] = k0.x; k[1] = k0.y; k[2] = k0.z; k[3] = k0.w; k[4] = k1.x; k[5] = k1.y; k[6] = k1.z; k[7] = k1.w;
                k[8] = k2.x; k[9] = k2.y; k[10] = k2.z; k[11] = k2.w; k[12] = k3.x; k[13] = k3.y; k[14] = k3.z; k[15] = k3.w;
                const float2 rc = lds_v2(ea + 64);
                rl = rl0 + __float_as_int(rc.x); cbase = __float_as_int(rc.y);
                lim = fminf(thr_list_s[rl], worst_s[rl]);
                if (kSoft) {
                    // does the entry hold a NON-candidate key (>= lim) inside the softmax window?  (one min-tree; most entries
                    // of a peaked softmax only carry their candidate)
                    thm_e = thr_mass_s[rl];
                    float nc[TC_CHUNK];
#pragma unroll
                    for (int t = 0; t < TC_CHUNK; ++t) nc[t] = k[t] >= lim ? k[t] : INFINITY;
                    need_mass = min16(nc) < thm_e;
                }
            } else {
#pragma unroll
                for (int t = 0; t < TC_CHUNK; ++t) k[t] = INFINITY;
            }
            if (kSoft && __any_sync(kFull, need_mass)) {
                // terms of the keys that can never be candidates (>= lim), inside the softmax window, against the snapshot reference
                if (need_mass) {
                    const float xx = xx_s[rl], c0 = p.a2 * r_s[rl];
#pragma unroll
                    for (int t = 0; t < TC_CHUNK; ++t) {
                        const float e = ex2_approx(fmaf(-p.a2, key_dist_fast(k[t], xx), c0));
                        if (k[t] >= lim && k[t] < thm_e) mass += e;
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < TC_CHUNK; ++t)                        // candidates keep their column offset in the 4 low bits
                k[t] = k[t] < lim ? __uint_as_float((__float_as_uint(k[t]) & ~15u) | (unsigned)t) : INFINITY;
            // the two smallest candidates of every entry, once per batch
            const float m1 = min16(k);
            float m2;
            {
                float cand[TC_CHUNK];
#pragma unroll
                for (int t = 0; t < TC_CHUNK; ++t) cand[t] = k[t] > m1 ? k[t] : INFINITY;
                m2 = min16(cand);
            }
            const unsigned peers = __match_any_sync(kFull, rl);
            if (kSoft) {
                // same-row lanes of the batch: the first one collects the others' sums and adds them to the row's accumulator
                // (a row's state belongs to this warp alone: no atomic)
                const int leader = __ffs(peers) - 1;
                unsigned rest = peers & ~(1u << leader);
                while (__any_sync(kFull, rest != 0u)) {
                    const int src = rest ? __ffs(rest) - 1 : lane;
                    const float other = __shfl_sync(kFull, mass, src);
                    if (rest && lane == leader) mass += other;
                    rest &= rest - 1u;
                }
                if (active && lane == leader && mass != 0.f) l_s[rl] += mass;
            }
            // entries with candidates: same-row entries one after the other, in queue order
            bool todo = active && m1 < INFINITY;
            unsigned done_mask = ~__ballot_sync(kFull, todo);
            while (done_mask != kFull) {
                const bool mine = todo && (__ffs(peers & ~done_mask) - 1 == lane);
                // row state of the lanes whose turn it is
                float xx = 0.f, thl = -INFINITY, thm = -INFINITY, kr = 0.f, r = 0.f, l = 0.f, worst = -INFINITY;
                const int lb = (mine ? rl : 0) * LIST_STRIDE;
                if (mine) {
                    xx = xx_s[rl]; thl = thr_list_s[rl]; thm = thr_mass_s[rl]; kr = kr_s[rl]; r = r_s[rl]; l = l_s[rl];
                    worst = worst_s[rl];
                    if (!kPrime && p.multi_split)
                        thl = fminf(thl, 0.5f * (__uint_as_float(__ldcg(p.thr_global + (size_t)b * p.N + row0 + (rl & (TC_SUB - 1)))) - xx));
                }
                bool changed = false;
                float prev = -INFINITY;                               // packed keys handled so far are <= prev
                // warp-uniform loop: every trip handles the next-best candidate key of every lane that still has one.  The two
                // smallest were found once per batch (m1, m2): the common single-candidate entry costs no second min-tree.
                for (int trip = 0;; ++trip) {
                    float m;
                    if (trip == 0) m = m1;
                    else if (trip == 1) m = m2;
                    else {
                        float cand[TC_CHUNK];
#pragma unroll
                        for (int t = 0; t < TC_CHUNK; ++t) cand[t] = k[t] > prev ? k[t] : INFINITY;
                        m = min16(cand);
                    }
                    const bool go = mine && m < INFINITY;             // every key below the snapshot is settled here: list or mass
                    if (!__any_sync(kFull, go)) break;
                    if (go) {
                        prev = m;
                        if (kSoft && m < kr) {                        // new row minimum: move the reference of the row's mass
                            const float rn = key_dist_fast(m, xx);
                            if (l != 0.f) l *= ex2_approx(-p.a2 * (r - rn));
                            kr = m; r = rn;
                            const float te = rn + p.cut_over_alpha;
                            thm = 0.5f * (te * te - xx);
                        }
                        float out = m;
                        if (m < fminf(thl, worst)) {                  // still a candidate: replace the worst entry, find the new worst
                            out = worst;
                            const int ws = (int)(__float_as_uint(worst) & 15u);
                            lkeys[lb + ws] = __uint_as_float((__float_as_uint(m) & ~15u) | (unsigned)ws);
                            lidx[lb + ws] = cbase + (int)(__float_as_uint(m) & 15u);
                            const float4* L4 = reinterpret_cast<const float4*>(lkeys + lb);
                            const float4 a0 = L4[0], a1 = L4[1];
                            float w = max3(max3(a0.x, a0.y, a0.z), max3(a0.w, a1.x, a1.y), fmaxf(a1.z, a1.w));
                            if (K > 8) {
                                const float4 a2 = L4[2], a3 = L4[3];
                                w = max3(w, max3(max3(a2.x, a2.y, a2.z), max3(a2.w, a3.x, a3.y), fmaxf(a3.z, a3.w)), w);
                            }
                            worst = w;
                            changed = true;
                        }
                        // evicted / rejected key inside the softmax window (empty-slot markers are ~3e38: never)
                        if (kSoft && out < fminf(thm, 1e37f)) l += ex2_approx(-p.a2 * (key_dist_fast(out, xx) - r));
                    }
                }
                if (mine) {
                    thl = fminf(thl, worst);
                    thr_list_s[rl] = thl; thr_mass_s[rl] = thm; kr_s[rl] = kr; r_s[rl] = r; l_s[rl] = l; worst_s[rl] = worst;
                    sts_v2(smem_u32(thr2_s + rl), kSoft ? fmaxf(thl, thm) : thl, thl);    // one 8-byte store: (bound, list bound) stay a consistent pair
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
                const float* L = lkeys + rl * LIST_STRIDE;
                if (kPrime) {
                    // merge the four sorted lists of chunk minima (column groups; two sit in the other half's list slot):
                    // KP-th smallest of the union, and the minimum
                    const float* L2 = L + TC_SUB * LIST_STRIDE;
                    int i0 = 0, i1 = KP, i2 = 0, i3 = KP;
                    float w = INFINITY;
                    for (int t = 0; t < KP; ++t) {
                        const float a0 = i0 < KP ? L[i0] : INFINITY, a1 = i1 < 2 * KP ? L[i1] : INFINITY;
                        const float a2 = i2 < KP ? L2[i2] : INFINITY, a3 = i3 < 2 * KP ? L2[i3] : INFINITY;
                        const float m01 = fminf(a0, a1), m23 = fminf(a2, a3);
                        w = fminf(m01, m23);
                        if (m01 <= m23) { if (a0 <= a1) ++i0; else ++i1; } else { if (a2 <= a3) ++i2; else ++i3; }
                    }
                    const float m = fminf(fminf(L[0], L[KP]), fminf(L2[0], L2[KP]));
                    if (p.prime_thr && w < LIST_EMPTY) atomicMin(p.thr_global + (size_t)b * p.N + row, __float_as_uint(fmaxf(fmaf(2.f, w, xx), 0.f)));
                    if (m < LIST_EMPTY) atomicMin(p.rmin_global + (size_t)b * p.N + row, __float_as_uint(fmaxf(fmaf(2.f, m, xx), 0.f)));
                } else {
                    const size_t g_row = (size_t)b * p.N + row;
                    const int part = split * 2 + (cw >> 2);                 // partial list index: (column split, column half)
                    const size_t base = (g_row * p.cb.P + part) * KC;
                    for (int t = 0; t < K; ++t) {
                        const float ek = L[t];
                        const bool has = ek < LIST_EMPTY;
                        p.cb.key[base + t] = has ? fmaxf(fmaf(2.f, ek, xx), 0.f) : INFINITY;       // back to the true d^2 domain
                        p.cb.idx[base + t] = has ? lidx[rl * LIST_STRIDE + t] : -1;
                    }
                    const float thl = thr_list_s[rl];
                    float l_out = l_s[rl];
                    const float r_out = r_s[rl];
                    if (kDense) {
                        // mass the scanners of this column half summed themselves (dense-window mode), reference r0 >= r_out
                        const float* lsc = reinterpret_cast<const float*>(Xs) + (cw >> 2) * 2 * TC_SUB + (rl & (TC_SUB - 1));
                        const float ls = lsc[0] + lsc[TC_SUB];
                        if (ls != 0.f) {
                            const float r0 = sqrtf(fmaxf(__uint_as_float(__ldcg(p.rmin_global + g_row)), 0.f));
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
    { const char* e = getenv("DVM_TC_DEBUG"); p.debug = e ? atoi(e) : 0; }
    p.multi_split = 1;                               // the two column halves of a tile are separate lists that share thresholds
    p.tile_stride = 1;
    fill_u32_kernel<<<ceil_div(2 * B * N, 256), 256, 0, st>>>(w.thr_g, 0x7f800000u, 2 * B * N);    // +inf (memset cannot write it)
    DVM_LAUNCH_CHECK();

    const size_t unit = (size_t)p.KB * TC_BLK_BYTES + TC_EXT_BYTES;
    const size_t smem = (1 + TC_NST) * unit + (size_t)TC_CONS_WARPS * Q_CAP * Q_ENTRY + (size_t)TC_BM * LIST_STRIDE * 8 + 9 * TC_BM * sizeof(float) + TC_SUB * sizeof(float)
                        + TC_CONS_WARPS * sizeof(QCtl) + 128;
    // dense-window instance for small alpha (every chunk inside the softmax window): crossover measured between alpha = 10
    // (141 -> 290 TFLOP/s at 50k) and alpha = 100 (810 -> 390 when forced)
    const bool dense = soft && alpha < TC_DENSE_ALPHA;
    auto kern = !soft ? softmap_cand_tc_kernel<false, false> : dense ? softmap_cand_tc_kernel<true, false, true> : softmap_cand_tc_kernel<true, false>;
    auto kprime = softmap_cand_tc_kernel<false, true>;
    static bool attr_done = false;
    if (!attr_done) {
        DVM_CUDA(cudaFuncSetAttribute(softmap_cand_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DVM_CUDA(cudaFuncSetAttribute(softmap_cand_tc_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DVM_CUDA(cudaFuncSetAttribute(softmap_cand_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DVM_CUDA(cudaFuncSetAttribute(kprime, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = true;
    }
    if (smem > 227 * 1024) { set_error("launch_cand_tc: C=%d needs %zu bytes of shared memory", C, smem); return DVM_ERR_UNSUPPORTED; }
    prof_begin(st);
    {
        // priming pass over every 10th tile (10 % of the sweep's MMA work): list threshold + the FIXED softmax reference of every
        // row.  Small problems (< 16 tiles) sample every other tile and only take the reference from it.
        TcParams pp = p;
        const bool big = p.tiles_total >= TC_PRIME_MIN_TILES;
        pp.tile_stride = big ? TC_PRIME_STRIDE : (p.tiles_total >= 2 ? 2 : 1);
        pp.prime_thr = big ? 1 : 0;
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
