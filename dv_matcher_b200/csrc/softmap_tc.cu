// tcgen05 candidate pass of the fused soft/hard map (sm_100a).
//
// Per CTA: a 128-row block of X (16-bit, TMA, SWIZZLE_128B, resident for the whole sweep) against a
// stream of 256-column tiles of Y (TMA ring), one tcgen05.mma chain (M=128, N=256, K=C) per tile with
// the fp32 accumulator in TMEM (2 x 256 columns, double buffered), and 8 epilogue warps that pull the
// accumulator back with tcgen05.ld (32x32b: one row per thread) and run the online top-16 / softmax
// sweep of common.cuh on  d^2 = |x|^2 + |y|^2 - 2 x.y  -- the N x M matrix never leaves the SM.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one lane),
// warps 2..9 = epilogue (TMEM lane quarter = warp % 4, column half = (warp - 2) / 4).
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), all mbarriers.
#include <cuda.h>
#include "softmap.cuh"

namespace dvm {

constexpr int TC_BM = 128;            // rows per CTA  (UMMA M)
constexpr int TC_BN = 256;            // columns per tile (UMMA N)
constexpr int TC_KBLK = 64;           // 16-bit elements per 128-byte swizzle row
constexpr int TC_THREADS = 320;
constexpr int TC_EPI_THREADS = 256;
constexpr int TC_NST = 2;             // Y ring depth
constexpr int TC_X_KB_BYTES = TC_BM * 128;     // 16 KB per K block
constexpr int TC_Y_KB_BYTES = TC_BN * 128;     // 32 KB per K block
constexpr int TC_MAX_SPLIT = P_MAX / 2;

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

// K-major operand, SWIZZLE_128B: rows of 128 B, 8-row atoms 1024 B apart (SBO), LBO unused (=1), version 1
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// ------------------------------------------------------------------------------------------------
// operand preparation: fp32 -> 16-bit (K padded to a multiple of 64), norms of the ROUNDED rows,
// per-row rounding error |x~ - x|_2 (certificate input)
// ------------------------------------------------------------------------------------------------
template <bool kBF16>
__global__ void __launch_bounds__(256)
tc_prep_kernel(const float* __restrict__ src, int rows_per_b, int rows_pad, int C, int Cpad,
               uint16_t* __restrict__ dst, float* __restrict__ nrm /* [B][rows_pad] */,
               float* __restrict__ err_row /* [B*rows_per_b] or null */, float* __restrict__ err_max /* [B] or null */) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.y;
    if (r >= rows_pad) return;
    if (r >= rows_per_b) {                       // norm padding: +inf masks the column in the epilogue
        if (lane == 0) nrm[(size_t)b * rows_pad + r] = INFINITY;
        return;
    }
    const float* s = src + ((size_t)b * rows_per_b + r) * C;
    uint16_t* d = dst + ((size_t)b * rows_per_b + r) * Cpad;
    float n2 = 0.f, e2 = 0.f;
    for (int c = lane; c < Cpad; c += 32) {
        const float v = c < C ? __ldg(s + c) : 0.f;
        float vr; uint16_t bits;
        if (kBF16) { const __nv_bfloat16 h = __float2bfloat16_rn(v); vr = __bfloat162float(h); bits = __bfloat16_as_ushort(h); }
        else       { const __half h = __float2half_rn(v);           vr = __half2float(h);     bits = __half_as_ushort(h); }
        d[c] = bits;
        n2 = fmaf(vr, vr, n2);
        const float e = v - vr;
        e2 = fmaf(e, e, e2);
    }
    n2 = warp_sum(n2); e2 = warp_sum(e2);
    if (lane == 0) {
        nrm[(size_t)b * rows_pad + r] = n2;
        const float e = sqrtf(e2) * 1.0001f;
        if (err_row) err_row[(size_t)b * rows_per_b + r] = e;
        if (err_max) atomicMax(reinterpret_cast<int*>(err_max + b), __float_as_int(e));   // e >= 0: int order == float order
    }
}

// ------------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------------
struct TcParams {
    int N, M, KB;                // KB = Cpad / 64
    int Npad, Mpad;              // row strides of the norm arrays
    int tiles_total, tiles_per_split;
    float a2, cut_over_alpha;
    uint32_t idesc;
    const float* xx; const float* yy;
    unsigned* thr_global;        // [B*N] per-row threshold shared by column-split CTAs (null when S == 1)
    CandBuffers cb;
};

// Epilogue candidate handling (per thread = one row of the tile, one column half):
//   * the KC best (key, idx) of the row live in REGISTERS as a sorted list;
//   * columns whose key beats the row threshold are appended, branch-free (predicated STS.64), to a
//     per-lane buffer in shared memory, slot-major ([slot][lane] -> conflict-free whatever the slots);
//   * when any lane's buffer is more than 1/3 full the whole warp flushes: every lane inserts ITS OWN
//     pending entries into its register list at the same time (no divergence in the bootstrap phase,
//     where all 32 rows are busy), evicted / rejected entries go to the row's softmax mass;
//   * the threshold of a row is shared between its partial lists -- the two column halves of a CTA via
//     shared memory, column-split CTAs via atomicMin in global memory -- so the total number of hits per
//     row stays ~ KC * ln(M) however many partial lists there are.  Any value ever published is the
//     KC-th best of 16 real columns, hence >= the final merged KC-th best: stale reads are safe.
constexpr int TC_CAP = 24;                 // buffer slots per lane; flush when any lane holds > 8
constexpr int TC_CHUNK = 16;               // columns per tcgen05.ld (one flush check per chunk)
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
          "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct EpiState {
    TopList<KC> list;      // keys in the key domain d^2 - |x|^2
    float thr_list;        // append threshold: min(own KC-th best, thresholds published by the row's other lists)
    float thr_mass;        // softmax cut-off in the key domain (soft mode): entries in [thr_list, thr_mass) only add mass
    float r, l;            // running min distance, mass of non-candidate columns relative to r
    int cnt;               // pending entries in the lane's buffer
};

template <bool kSoft>
__device__ __forceinline__ void epi_mass_add(EpiState& st, float key, float xx, float a2) {
    if (kSoft) {
        // r is the smallest distance among EVERYTHING counted so far (list or mass): a list that adopted a
        // tighter threshold from the row's other lists may hold only far entries, so a mass-only column can be
        // closer than the list head -- rescale instead of evaluating exp2 of a large positive number.
        const float d = sqrtf(fmaxf(key + xx, 0.f));
        if (d < st.r) { if (st.l != 0.f) st.l *= exp2f(-a2 * (st.r - d)); st.r = d; }
        st.l += exp2f(-a2 * (d - st.r));
    }
}

// lane-parallel flush of the pending buffers (warp-uniform trip count)
template <bool kSoft>
__device__ __forceinline__ void epi_flush(EpiState& st, const float2* buf, int lane, float xx, float a2, float coa) {
    const int mx = __reduce_max_sync(kFull, st.cnt);
    for (int e = 0; e < mx; ++e) {
        if (e < st.cnt) {
            const float2 kv = buf[e * 32 + lane];
            const float key = kv.x;
            if (key < st.list.worst()) {
                const float ev = st.list.push(key, __float_as_int(kv.y));
                if (kSoft) {
                    const float rn = sqrtf(fmaxf(st.list.key[0] + xx, 0.f));
                    if (rn < st.r) { if (st.l != 0.f) st.l *= exp2f(-a2 * (st.r - rn)); st.r = rn; }
                    if (ev != INFINITY) epi_mass_add<kSoft>(st, ev, xx, a2);
                }
            } else {
                epi_mass_add<kSoft>(st, key, xx, a2);       // beaten since it was appended: mass only
            }
        }
    }
    st.cnt = 0;
    st.thr_list = fminf(st.thr_list, st.list.worst());
    if (kSoft && st.r != INFINITY) { const float te = st.r + coa; st.thr_mass = te * te - xx; }
}

template <bool kSoft>
__global__ void __launch_bounds__(TC_THREADS, 1)
softmap_cand_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];     // SWIZZLE_128B needs 1024-byte alignment (checked below)
    uint8_t* Xs = smem;
    uint8_t* Ys = Xs + p.KB * TC_X_KB_BYTES;
    float* yy_s = reinterpret_cast<float*>(Ys + TC_NST * p.KB * TC_Y_KB_BYTES);       // [2][256]
    float2* cand_buf = reinterpret_cast<float2*>(yy_s + 2 * TC_BN);                   // [8 warps][TC_CAP][32] (key, idx)
    float* thr_sh = reinterpret_cast<float*>(cand_buf + (TC_EPI_THREADS / 32) * TC_CAP * 32);   // [2 halves][128 rows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(thr_sh + 2 * TC_BM);
    uint64_t* full = bars;                 // [NST]
    uint64_t* empty = bars + TC_NST;       // [NST]
    uint64_t* tfull = bars + 2 * TC_NST;   // [2]
    uint64_t* tempty = tfull + 2;          // [2]
    uint64_t* xfull = tempty + 2;          // [1]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(xfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z;
    const int split = blockIdx.y;
    const int row0 = blockIdx.x * TC_BM;
    const int tile0 = split * p.tiles_per_split;
    const int ntiles = min(p.tiles_per_split, p.tiles_total - tile0);

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (int s = 0; s < TC_NST; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, TC_EPI_THREADS / 32); }
        mbar_init(xfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmY);
    }
    if (warp == 1) {                        // TMEM: all 512 columns (2 accumulator stages x 256)
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
            mbar_arrive_expect_tx(xfull, p.KB * TC_X_KB_BYTES);
            for (int kb = 0; kb < p.KB; ++kb) tma_load_3d(&tmX, xfull, Xs + kb * TC_X_KB_BYTES, kb * TC_KBLK, row0, b);
            for (int it = 0; it < ntiles; ++it) {
                const int s = it % TC_NST;
                const uint32_t ph = (it / TC_NST) & 1;
                mbar_wait(empty + s, ph ^ 1);
                mbar_arrive_expect_tx(full + s, p.KB * TC_Y_KB_BYTES);
                for (int kb = 0; kb < p.KB; ++kb)
                    tma_load_3d(&tmY, full + s, Ys + (s * p.KB + kb) * TC_Y_KB_BYTES, kb * TC_KBLK, (tile0 + it) * TC_BN, b);
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
                mbar_wait(tempty + acc, aph ^ 1);          // epilogue has drained this accumulator
                mbar_wait(full + s, ph);                   // Y tile landed
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * TC_BN;
                for (int kb = 0; kb < p.KB; ++kb) {
                    const uint32_t xa = smem_u32(Xs + kb * TC_X_KB_BYTES);
                    const uint32_t ya = smem_u32(Ys + (s * p.KB + kb) * TC_Y_KB_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_KBLK / 16; ++k)     // UMMA_K = 16 -> +32 bytes inside the swizzle row
                        tc_mma_f16(d_tmem, umma_desc_sw128(xa + k * 32), umma_desc_sw128(ya + k * 32), p.idesc, (kb | k) != 0);
                }
                tc_commit(empty + s);                      // smem slot reusable once these MMAs retire
                tc_commit(tfull + acc);                    // accumulator ready for the epilogue
            }
        }
    } else {
        // =============================== epilogue ===============================
        const int ew = warp - 2;                           // 0..7
        const int quarter = warp & 3;                      // TMEM lanes 32*quarter .. +31 are this warp's
        const int half = ew >> 2;                          // columns half*128 .. +127 of each tile
        const int etid = threadIdx.x - 64;                 // 0..255
        const int rloc = quarter * 32 + lane;              // row inside the CTA tile
        const int row = row0 + rloc;
        const bool row_ok = row < p.N;
        const float xx = row_ok ? __ldg(p.xx + (size_t)b * p.Npad + row) : 0.f;
        float2* buf = cand_buf + ew * TC_CAP * 32;
        volatile float* thr_mine = thr_sh + half * TC_BM + rloc;
        volatile float* thr_other = thr_sh + (1 - half) * TC_BM + rloc;
        unsigned* thr_g = p.thr_global ? p.thr_global + (size_t)b * p.N + row : nullptr;    // true-domain d^2 bits
        *thr_mine = INFINITY;
        EpiState st;
        st.list.init();
        st.thr_list = row_ok ? INFINITY : -INFINITY;       // padding rows never hit
        st.thr_mass = -INFINITY;                           // no mass-only entries until the running minimum exists
        st.r = INFINITY; st.l = 0.f; st.cnt = 0;
        for (int it = 0; it < ntiles; ++it) {
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int col0 = (tile0 + it) * TC_BN;
            yy_s[acc * TC_BN + etid] = __ldg(p.yy + (size_t)b * p.Mpad + col0 + etid);
            // pick up thresholds published by the row's other lists -- only once the own list is full, so that
            // the own running minimum (the reference point of the softmax mass) exists before anything is rejected
            if (row_ok && st.list.worst() != INFINITY) {
                float t = *thr_other;
                if (thr_g) t = fminf(t, __uint_as_float(__ldcg(thr_g)) - xx);
                st.thr_list = fminf(st.thr_list, t);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(tfull + acc, aph);
            tc_fence_after();
            const float* yv = yy_s + acc * TC_BN + half * 128;
#pragma unroll 1
            for (int c = 0; c < 128 / TC_CHUNK; ++c) {
                float v[TC_CHUNK];
                tc_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * TC_BN + half * 128 + c * TC_CHUNK, v);
                const int cbase = col0 + half * 128 + c * TC_CHUNK;
                const float thr_hi = kSoft ? fmaxf(st.thr_list, st.thr_mass) : st.thr_list;
                float key[TC_CHUNK];
                bool any_mass = false;
#pragma unroll
                for (int q = 0; q < TC_CHUNK / 4; ++q) {
                    const float4 y4 = *reinterpret_cast<const float4*>(yv + c * TC_CHUNK + q * 4);
                    key[q * 4 + 0] = fmaf(-2.f, v[q * 4 + 0], y4.x);
                    key[q * 4 + 1] = fmaf(-2.f, v[q * 4 + 1], y4.y);
                    key[q * 4 + 2] = fmaf(-2.f, v[q * 4 + 2], y4.z);
                    key[q * 4 + 3] = fmaf(-2.f, v[q * 4 + 3], y4.w);
                }
#pragma unroll
                for (int t = 0; t < TC_CHUNK; ++t) {
                    const bool hit = key[t] < st.thr_list;           // predicated, branch-free append
                    if (hit) buf[st.cnt * 32 + lane] = make_float2(key[t], __int_as_float(cbase + t));
                    st.cnt += hit ? 1 : 0;
                    if (kSoft) any_mass |= (!hit) && (key[t] < thr_hi);
                }
                if (kSoft) {
                    if (__any_sync(kFull, any_mass)) {              // rare in the peaked regime
#pragma unroll
                        for (int t = 0; t < TC_CHUNK; ++t)
                            if (key[t] >= st.thr_list && key[t] < thr_hi) epi_mass_add<kSoft>(st, key[t], xx, p.a2);
                    }
                }
                if (__any_sync(kFull, st.cnt > TC_CAP - TC_CHUNK)) {
                    epi_flush<kSoft>(st, buf, lane, xx, p.a2, p.cut_over_alpha);
                    if (row_ok) {
                        const float w = st.list.worst();
                        *thr_mine = w;
                        if (w != INFINITY) st.thr_list = fminf(st.thr_list, *thr_other);
                        if (thr_g && w != INFINITY) atomicMin(thr_g, __float_as_uint(fmaxf(w + xx, 0.f)));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + acc);
        }
        epi_flush<kSoft>(st, buf, lane, xx, p.a2, p.cut_over_alpha);
        if (row_ok) {
            const size_t g_row = (size_t)b * p.N + row;
            const int pidx = split * 2 + half;
            const size_t base = (g_row * p.cb.P + pidx) * KC;
#pragma unroll
            for (int t = 0; t < KC; ++t) {
                const float k = st.list.key[t];
                p.cb.key[base + t] = k == INFINITY ? INFINITY : fmaxf(k + xx, 0.f);    // back to the true d^2 domain
                p.cb.idx[base + t] = st.list.idx[t];
            }
            p.cb.l[g_row * p.cb.P + pidx] = st.l;
            p.cb.r[g_row * p.cb.P + pidx] = st.r;
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

// [B][rows][Cpad] 16-bit, box {64, box_rows, 1}, 128-byte swizzle, out-of-range rows read as zero
static int make_operand_map(CUtensorMap* map, const void* base, bool bf16, int B, int rows, int Cpad, int box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return DVM_ERR_DEVICE; }
    cuuint64_t dims[3] = {(cuuint64_t)Cpad, (cuuint64_t)rows, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)Cpad * 2, (cuuint64_t)rows * Cpad * 2};
    cuuint32_t box[3] = {(cuuint32_t)TC_KBLK, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base),
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return DVM_ERR_DEVICE; }
    return 0;
}

static int choose_split(int B, int N, int M) {
    const int row_blocks = ceil_div(N, TC_BM);
    const int tiles = ceil_div(M, TC_BN);
    int best = 1; double best_eff = -1.0;
    for (int s = 1; s <= TC_MAX_SPLIT; ++s) {
        if (s > 1 && tiles / s < 4) break;                       // keep >= 4 tiles per CTA to amortise the X load
        const int tps = ceil_div(tiles, s);
        if ((s - 1) * tps >= tiles) continue;                    // would leave an empty split
        const long long ctas = (long long)row_blocks * s * B;
        const double waves = (double)ctas / kNumSM;
        const double eff = waves / ceil(waves) - 0.01 * (s - 1); // prefer fewer partial lists on ties
        if (eff > best_eff) { best_eff = eff; best = s; }
    }
    return best;
}

int tc_num_partials(int B, int N, int M) { return 2 * choose_split(B, N, M); }

static size_t tc_ws_layout(void* base, size_t cap, int B, int N, int M, int C,
                           uint16_t** Xh, uint16_t** Yh, float** xx, float** yy, unsigned** thr_g,
                           int* Cpad_o, int* Npad_o, int* Mpad_o) {
    const int Cpad = ceil_div(C, TC_KBLK) * TC_KBLK;
    const int Npad = ceil_div(N, TC_BM) * TC_BM;
    const int Mpad = ceil_div(M, TC_BN) * TC_BN;
    WsCarver ws(base, cap);
    uint16_t* a = ws.take<uint16_t>((size_t)B * N * Cpad);
    uint16_t* bq = ws.take<uint16_t>((size_t)B * M * Cpad);
    float* c = ws.take<float>((size_t)B * Npad);
    float* d = ws.take<float>((size_t)B * Mpad);
    unsigned* tg = ws.take<unsigned>((size_t)B * N);
    if (Xh) *Xh = a; if (Yh) *Yh = bq; if (xx) *xx = c; if (yy) *yy = d; if (thr_g) *thr_g = tg;
    if (Cpad_o) *Cpad_o = Cpad; if (Npad_o) *Npad_o = Npad; if (Mpad_o) *Mpad_o = Mpad;
    return align_up(ws.off, 256);
}

size_t tc_workspace_bytes(int B, int N, int M, int C) {
    return tc_ws_layout(nullptr, 0, B, N, M, C, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int launch_cand_tc(const float* X, const float* Y, int B, int N, int M, int C, float alpha, bool soft, int prec,
                   CandBuffers cb, float* err_x, float* err_ymax, void* wsp, size_t ws_bytes, cudaStream_t st) {
    uint16_t *Xh, *Yh; float *xx, *yy; unsigned* thr_g; int Cpad, Npad, Mpad;
    const size_t need = tc_ws_layout(wsp, ws_bytes, B, N, M, C, &Xh, &Yh, &xx, &yy, &thr_g, &Cpad, &Npad, &Mpad);
    if (!wsp || need > ws_bytes) { set_error("launch_cand_tc: workspace too small"); return DVM_ERR_WORKSPACE; }
    if (B > 65535) { set_error("launch_cand_tc: B=%d too large", B); return DVM_ERR_INVALID_ARG; }
    const bool bf16 = prec == DVM_PREC_BF16;

    DVM_CUDA(cudaMemsetAsync(err_ymax, 0, (size_t)B * sizeof(float), st));
    {
        dim3 gx(ceil_div(Npad, 8), B), gy(ceil_div(Mpad, 8), B);
        if (bf16) {
            tc_prep_kernel<true><<<gx, 256, 0, st>>>(X, N, Npad, C, Cpad, Xh, xx, err_x, nullptr);
            DVM_LAUNCH_CHECK();
            tc_prep_kernel<true><<<gy, 256, 0, st>>>(Y, M, Mpad, C, Cpad, Yh, yy, nullptr, err_ymax);
            DVM_LAUNCH_CHECK();
        } else {
            tc_prep_kernel<false><<<gx, 256, 0, st>>>(X, N, Npad, C, Cpad, Xh, xx, err_x, nullptr);
            DVM_LAUNCH_CHECK();
            tc_prep_kernel<false><<<gy, 256, 0, st>>>(Y, M, Mpad, C, Cpad, Yh, yy, nullptr, err_ymax);
            DVM_LAUNCH_CHECK();
        }
    }

    CUtensorMap tmX, tmY;
    int rc;
    if ((rc = make_operand_map(&tmX, Xh, bf16, B, N, Cpad, TC_BM))) return rc;
    if ((rc = make_operand_map(&tmY, Yh, bf16, B, M, Cpad, TC_BN))) return rc;

    TcParams p{};
    p.N = N; p.M = M; p.KB = Cpad / TC_KBLK; p.Npad = Npad; p.Mpad = Mpad;
    p.tiles_total = ceil_div(M, TC_BN);
    const int S = cb.P / 2;
    p.tiles_per_split = ceil_div(p.tiles_total, S);
    p.a2 = alpha * kLog2e;
    p.cut_over_alpha = alpha > 0.f ? kExpCut / alpha : INFINITY;
    // instruction descriptor: D=f32 (bits 4-5 = 1), A/B format (0 = f16, 1 = bf16) at bits 7-9 / 10-12, K-major A and B,
    // N >> 3 at bits 17-22, M >> 4 at bits 24-28
    const uint32_t fmt = bf16 ? 1u : 0u;
    p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    p.xx = xx; p.yy = yy; p.cb = cb;
    p.thr_global = nullptr;
    if (S > 1) {                                     // +inf bit pattern: 0x7f800000 (byte-wise memset cannot write it)
        p.thr_global = thr_g;
        fill_u32_kernel<<<ceil_div(B * N, 256), 256, 0, st>>>(thr_g, 0x7f800000u, B * N);
        DVM_LAUNCH_CHECK();
    }

    const size_t smem = (size_t)p.KB * TC_X_KB_BYTES + (size_t)TC_NST * p.KB * TC_Y_KB_BYTES + 2 * TC_BN * sizeof(float)
                        + (size_t)(TC_EPI_THREADS / 32) * TC_CAP * 32 * 8 + 2 * TC_BM * sizeof(float) + 128;
    auto kern = soft ? softmap_cand_tc_kernel<true> : softmap_cand_tc_kernel<false>;
    static bool attr_done[2] = {false, false};
    if (!attr_done[soft]) {
        DVM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done[soft] = true;
    }
    if (smem > 227 * 1024) { set_error("launch_cand_tc: C=%d needs %zu bytes of shared memory", C, smem); return DVM_ERR_UNSUPPORTED; }
    dim3 grid(ceil_div(N, TC_BM), S, B);
    prof_begin(st);
    kern<<<grid, TC_THREADS, smem, st>>>(tmX, tmY, p);
    prof_end(st);
    DVM_LAUNCH_CHECK();
    return 0;
}

}  // namespace dvm
