// Soft/hard correspondence map: fp32 CUDA-core candidate pass, exact finalize, and the C entry point.
//
// Pipeline of dvm_softmap_fwd (replaces models/loss.py:91-95, 110-114, 1339-1347, 1408-1409):
//   1. candidate pass  (tcgen05 f16/bf16 in softmap_tc.cu, or fp32 SIMT below): one sweep over all M
//      columns per row keeps the KC=16 smallest squared distances and the softmax mass of the rest;
//   2. finalize        : merges the partial lists of a row, re-scores the 16 candidates EXACTLY in fp32
//      from the fp32 inputs (direct differences, the `donot_use_mm_for_euclid_dist` form), orders them
//      (ties -> lower index), emits arg-min, top-k (idx, w, d), row statistics and Pi.V, and certifies
//      that no discarded column can belong to the top-k given the candidate pass' error bound;
//   3. rows that fail the certificate are recomputed by the fp32 pass (device-side list, no host sync).
#include <cooperative_groups.h>
#include "softmap.cuh"

namespace dvm {

// =================================================================================================
// 1. fp32 CUDA-core candidate pass
// =================================================================================================
constexpr int SIMT_BM = 64;       // rows per CTA
constexpr int SIMT_BN = 64;       // columns per tile
constexpr int SIMT_CG = 4;        // column groups (threads per row) -> P = 4 partial lists per row
constexpr int SIMT_CPT = SIMT_BN / SIMT_CG;   // 16 columns per thread per tile
constexpr int SIMT_THREADS = SIMT_BM * SIMT_CG;
constexpr int ROWS_EXACT_MAX = 2048;      // up to this many uncertified rows take the one-CTA-per-row kernel

template <bool kSoft>
__global__ void __launch_bounds__(SIMT_THREADS)
softmap_cand_simt_kernel(const float* __restrict__ X, const float* __restrict__ Y, int N, int M, int C,
                         const int* __restrict__ row_list, const int* __restrict__ row_count,
                         float a2, float cut_over_alpha, CandBuffers cb) {
    extern __shared__ __align__(16) float smem[];
    const int ld = C + 4;                        // padded row stride: conflict-free LDS.128
    float* Xs = smem;
    float* Ys = smem + SIMT_BM * ld;
    __shared__ int s_rows[SIMT_BM];

    const int tid = threadIdx.x;
    const int r = tid & (SIMT_BM - 1);
    const int cg = tid >> 6;
    int n_rows;
    if (row_list) {
        n_rows = *row_count;
        if (n_rows <= ROWS_EXACT_MAX) return;                 // few rows: softmap_rows_exact_kernel handles them
        if ((int)blockIdx.x * SIMT_BM >= n_rows) return;
        if (tid < SIMT_BM) {
            const int e = blockIdx.x * SIMT_BM + tid;
            s_rows[tid] = e < n_rows ? row_list[e] : -1;
        }
    } else {
        if (tid < SIMT_BM) {
            const int i = blockIdx.x * SIMT_BM + tid;
            s_rows[tid] = i < N ? (int)blockIdx.y * N + i : -1;
        }
    }
    __syncthreads();

    // X tile: 64 rows x C, float4 granularity
    const int c4 = C >> 2;
    for (int e = tid; e < SIMT_BM * c4; e += SIMT_THREADS) {
        const int rr = e / c4, cc = e - rr * c4;
        const int g = s_rows[rr];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g >= 0) v = __ldg(reinterpret_cast<const float4*>(X + (size_t)g * C) + cc);
        *reinterpret_cast<float4*>(Xs + rr * ld + cc * 4) = v;
    }
    const int g_row = s_rows[r];
    // batch of this CTA: list mode mixes batches, so Y is addressed per row
    const int b_of_row = g_row >= 0 ? g_row / N : 0;
    // In list mode rows of one CTA may belong to different batch elements; the Y tile is shared by the
    // CTA, so the CTA sweeps once per distinct batch element present (usually one).
    int b_lo = row_list ? 0x7fffffff : (int)blockIdx.y, b_hi = row_list ? -1 : (int)blockIdx.y;
    if (row_list) {
        for (int t = 0; t < SIMT_BM; ++t) {
            const int g = s_rows[t];
            if (g >= 0) { b_lo = min(b_lo, g / N); b_hi = max(b_hi, g / N); }
        }
    }

    RowState st;
    st.init();

    for (int b = b_lo; b <= b_hi; ++b) {
        const bool mine = (g_row >= 0) && (b_of_row == b);
        const float* Yb = Y + (size_t)b * M * C;
        for (int j0 = 0; j0 < M; j0 += SIMT_BN) {
            __syncthreads();
            for (int e = tid; e < SIMT_BN * c4; e += SIMT_THREADS) {
                const int rr = e / c4, cc = e - rr * c4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j0 + rr < M) v = __ldg(reinterpret_cast<const float4*>(Yb + (size_t)(j0 + rr) * C) + cc);
                *reinterpret_cast<float4*>(Ys + rr * ld + cc * 4) = v;
            }
            __syncthreads();

            float acc[SIMT_CPT];
#pragma unroll
            for (int c = 0; c < SIMT_CPT; ++c) acc[c] = 0.f;
            const float* xr = Xs + r * ld;
            const float* yr = Ys + (cg * SIMT_CPT) * ld;
#pragma unroll 2
            for (int k = 0; k < C; k += 4) {
                const float4 xv = *reinterpret_cast<const float4*>(xr + k);
#pragma unroll
                for (int c = 0; c < SIMT_CPT; ++c) {
                    const float4 yv = *reinterpret_cast<const float4*>(yr + c * ld + k);   // warp-broadcast
                    float d;
                    d = xv.x - yv.x; acc[c] = fmaf(d, d, acc[c]);
                    d = xv.y - yv.y; acc[c] = fmaf(d, d, acc[c]);
                    d = xv.z - yv.z; acc[c] = fmaf(d, d, acc[c]);
                    d = xv.w - yv.w; acc[c] = fmaf(d, d, acc[c]);
                }
            }
            if (mine) {
#pragma unroll
                for (int c = 0; c < SIMT_CPT; ++c) {
                    const int j = j0 + cg * SIMT_CPT + c;
                    if (j < M && acc[c] < st.thr) row_state_visit<kSoft>(st, acc[c], j, a2, cut_over_alpha);
                }
            }
        }
    }

    if (g_row >= 0) {
        const size_t base = ((size_t)g_row * cb.P + cg) * KC;
#pragma unroll
        for (int t = 0; t < KC; ++t) { cb.key[base + t] = st.list.key[t]; cb.idx[base + t] = st.list.idx[t]; }
        cb.l[(size_t)g_row * cb.P + cg] = st.l;
        cb.r[(size_t)g_row * cb.P + cg] = st.r;
        cb.t[(size_t)g_row * cb.P + cg] = INFINITY;
    }
}

int launch_cand_simt(const float* X, const float* Y, int B, int N, int M, int C, float alpha, bool soft,
                     const int* row_list, const int* row_count, int max_rows, CandBuffers cb, cudaStream_t st) {
    const size_t smem = (size_t)(SIMT_BM + SIMT_BN) * (C + 4) * sizeof(float);
    static PerDeviceOnce attr_done[2];
    auto kern = soft ? softmap_cand_simt_kernel<true> : softmap_cand_simt_kernel<false>;
    if (attr_done[soft].need()) {
        DVM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_done[soft].done();
    }
    dim3 grid;
    if (row_list) grid = dim3(ceil_div(max_rows, SIMT_BM), 1, 1);
    else          grid = dim3(ceil_div(N, SIMT_BM), B, 1);
    const float a2 = alpha * kLog2e;
    const float coa = alpha > 0.f ? kExpCut / alpha : INFINITY;
    if (!row_list) prof_begin(st);
    kern<<<grid, SIMT_THREADS, smem, st>>>(X, Y, N, M, C, row_list, row_count, a2, coa, cb);
    if (!row_list) prof_end(st);
    DVM_LAUNCH_CHECK();
    return 0;
}

// =================================================================================================
// 1b. exact fp32 pass for a FEW rows (the uncertified rows of the 16-bit pass): 32 CTAs per row (see the kernel), each thread
//     striding over its share of the M columns.  Two passes over Y: (A) per-thread top-KC -> warp merge -> CTA merge -> cluster
//     merge, (B) softmax mass of the non-candidates relative to the partial list's minimum.  Writes the SIMT partial lists.
// =================================================================================================
constexpr int RE_THREADS = 256;
constexpr int RE_SLOTS = 32;               // row slots of the launch (x 32 CTAs each)
constexpr int RE_SPLIT = 8;                // CTAs (one cluster) per row; RE_SPLIT * KC = 128 entries in the cluster merge

__device__ __forceinline__ float row_d2(const float* __restrict__ xs, const float* __restrict__ y, int C) {
    float acc = 0.f;
#pragma unroll 8
    for (int k = 0; k < C; k += 4) {
        const float4 xv = *reinterpret_cast<const float4*>(xs + k);
        const float4 yv = __ldg(reinterpret_cast<const float4*>(y + k));
        float d;
        d = xv.x - yv.x; acc = fmaf(d, d, acc);
        d = xv.y - yv.y; acc = fmaf(d, d, acc);
        d = xv.z - yv.z; acc = fmaf(d, d, acc);
        d = xv.w - yv.w; acc = fmaf(d, d, acc);
    }
    return acc;
}

// Selection merge of 128 (key, idx) entries held 4 per lane by one warp: KC rounds of lexicographic warp-min; lane 0 hands
// round s to `emit(s, key, idx)`.
template <typename Emit>
__device__ __forceinline__ void re_merge128(float (&ek)[4], int (&ei)[4], Emit emit) {
    bool taken[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) taken[q] = ei[q] < 0;
    for (int s = 0; s < KC; ++s) {
        float bk = INFINITY; int bi = 0x7fffffff; int bq = -1;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (!taken[q] && kv_less(ek[q], ei[q], bk, bi)) { bk = ek[q]; bi = ei[q]; bq = q; }
        float wk = bk; int wi = bi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ok = __shfl_xor_sync(0xffffffffu, wk, o);
            const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
            if (kv_less(ok, oi, wk, wi)) { wk = ok; wi = oi; }
        }
        if (bq >= 0 && bk == wk && bi == wi) {
#pragma unroll
            for (int q = 0; q < 4; ++q) if (q == bq) taken[q] = true;
        }
        emit(s, wk, wi);
    }
}

// cb.P CLUSTERS of RE_SPLIT CTAs per row, one per partial list of the row: the columns are dealt round-robin to the
// cb.P * RE_SPLIT CTAs, so that the usual case -- a handful of rows, far fewer than SMs -- costs 1/32 of a single-SM column
// sweep (one row used to take 1.9 ms at M = 50 000 and set the step time of whichever batch contained it).  The per-CTA
// top-KC lists and masses of a cluster meet in its CTA 0 through distributed shared memory; the finalize pass merges the cb.P
// partial lists as it does for the SIMT pass.
template <bool kSoft>
__global__ void __cluster_dims__(RE_SPLIT, 1, 1) __launch_bounds__(RE_THREADS)
softmap_rows_exact_kernel(const float* __restrict__ X, const float* __restrict__ Y, int N, int M, int C,
                          const int* __restrict__ row_list, const int* __restrict__ row_count,
                          float a2, float cut_over_alpha, CandBuffers cb) {
    namespace cg = cooperative_groups;
    __shared__ __align__(16) float xs[256];
    __shared__ float s_key[(RE_THREADS / 32) * KC];
    __shared__ int s_idx[(RE_THREADS / 32) * KC];
    __shared__ float s_red[RE_THREADS / 32];
    __shared__ float s_pk[KC]; __shared__ int s_pi[KC];          // this CTA's partial list (read by CTA 0)
    __shared__ float s_wk; __shared__ int s_wi; __shared__ float s_r; __shared__ float s_l;
    const int n_rows = *row_count;
    const int part = (blockIdx.x / RE_SPLIT) % cb.P;             // which partial list of the row this cluster produces
    if (n_rows > ROWS_EXACT_MAX) return;
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // the grid holds RE_SLOTS row slots (an empty launch -- the usual case -- must cost nothing); the whole cluster walks the
    // rows slot, slot + RE_SLOTS, ... together
    for (int slot = blockIdx.x / (RE_SPLIT * cb.P); slot < n_rows; slot += gridDim.x / (RE_SPLIT * cb.P)) {
    const int g = row_list[slot];
    const int b = g / N;
    for (int c = tid; c < C; c += RE_THREADS) xs[c] = X[(size_t)g * C + c];
    __syncthreads();
    const float* Yb = Y + (size_t)b * M * C;
    // the row's columns are dealt round-robin to the cb.P * RE_SPLIT CTAs that work on it
    const int j0 = (part * RE_SPLIT + crank) * RE_THREADS + tid, jstep = cb.P * RE_SPLIT * RE_THREADS;

    // ---- pass A: per-thread sorted top-KC over this CTA's columns
    TopList<KC> list;
    list.init();
    for (int j = j0; j < M; j += jstep) {
        const float d2 = row_d2(xs, Yb + (size_t)j * C, C);
        if (d2 < list.worst()) list.push(d2, j);
    }
    // warp merge: KC rounds of lexicographic warp-min over the list heads
    int head = 0;
    for (int s = 0; s < KC; ++s) {
        float hk = INFINITY; int hi = 0x7fffffff;
#pragma unroll
        for (int t = 0; t < KC; ++t) if (t == head && list.idx[t] >= 0) { hk = list.key[t]; hi = list.idx[t]; }
        float wk = hk; int wi = hi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ok = __shfl_xor_sync(0xffffffffu, wk, o);
            const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
            if (kv_less(ok, oi, wk, wi)) { wk = ok; wi = oi; }
        }
        if (hi == wi && hi != 0x7fffffff) ++head;
        if (lane == 0) { s_key[wid * KC + s] = wk; s_idx[wid * KC + s] = wi == 0x7fffffff ? -1 : wi; }
    }
    __syncthreads();
    // CTA merge by warp 0: 8 warps x KC = 128 entries
    if (wid == 0) {
        float ek[4]; int ei[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { ek[q] = s_key[lane + 32 * q]; ei[q] = s_idx[lane + 32 * q]; }
        re_merge128(ek, ei, [&](int s, float wk, int wi) { if (lane == 0) { s_pk[s] = wk; s_pi[s] = wi == 0x7fffffff ? -1 : wi; } });
    }
    cluster.sync();
    // cluster merge by warp 0 of CTA 0: RE_SPLIT x KC = 128 entries, read from the peers' shared memory
    if (crank == 0 && wid == 0) {
        float ek[4]; int ei[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = lane + 32 * q;
            ek[q] = *cluster.map_shared_rank(&s_pk[e % KC], e / KC);
            ei[q] = *cluster.map_shared_rank(&s_pi[e % KC], e / KC);
        }
        const size_t base = ((size_t)g * cb.P + part) * KC;
        float best = INFINITY;
        re_merge128(ek, ei, [&](int s, float wk, int wi) {
            if (s == 0) best = wk;
            if (lane == 0) {
                cb.key[base + s] = wk;
                cb.idx[base + s] = wi == 0x7fffffff ? -1 : wi;
                if (s == KC - 1) { s_wk = wk; s_wi = wi; }
            }
        });
        if (lane == 0) s_r = sqrtf(best);
    }
    cluster.sync();
    // ---- pass B: mass of the non-candidates, relative to the exact row minimum
    const float r = *cluster.map_shared_rank(&s_r, 0);
    float l = 0.f;
    if (kSoft) {
        const float wk = *cluster.map_shared_rank(&s_wk, 0);
        const int wi = *cluster.map_shared_rank(&s_wi, 0);
        const float te = r + cut_over_alpha;
        const float cut2 = te * te;
        for (int j = j0; j < M; j += jstep) {
            const float d2 = row_d2(xs, Yb + (size_t)j * C, C);
            if (kv_less(wk, wi, d2, j) && d2 < cut2) l += exp2f(-a2 * (sqrtf(d2) - r));
        }
        l = warp_sum(l);
        if (lane == 0) s_red[wid] = l;
        __syncthreads();
        if (tid == 0) { l = 0.f; for (int w = 0; w < RE_THREADS / 32; ++w) l += s_red[w]; s_l = l; }
    }
    cluster.sync();
    if (crank == 0 && tid == 0) {
        l = 0.f;
        if (kSoft) for (int q = 0; q < RE_SPLIT; ++q) l += *cluster.map_shared_rank(&s_l, q);
        cb.l[(size_t)g * cb.P + part] = l; cb.r[(size_t)g * cb.P + part] = r; cb.t[(size_t)g * cb.P + part] = INFINITY;
    }
    cluster.sync();                                               // nobody exits (or starts the next row) while CTA 0 may still read its shared memory
    }
}

static int launch_rows_exact(const float* X, const float* Y, int N, int M, int C, float alpha, bool soft,
                             const int* row_list, const int* row_count, int max_rows, CandBuffers cb, cudaStream_t st) {
    const int grid = RE_SPLIT * cb.P * (max_rows < RE_SLOTS ? max_rows : RE_SLOTS);
    const float a2 = alpha * kLog2e;
    const float coa = alpha > 0.f ? kExpCut / alpha : INFINITY;
    if (soft) softmap_rows_exact_kernel<true><<<grid, RE_THREADS, 0, st>>>(X, Y, N, M, C, row_list, row_count, a2, coa, cb);
    else      softmap_rows_exact_kernel<false><<<grid, RE_THREADS, 0, st>>>(X, Y, N, M, C, row_list, row_count, a2, coa, cb);
    DVM_LAUNCH_CHECK();
    return 0;
}

// =================================================================================================
// 2. finalize: merge partial lists, exact fp32 re-scoring, ordering, weights, certificate, Pi.V
// =================================================================================================
constexpr int FIN_WARPS = 8;

struct FinalizeArgs {
    const float* X; const float* Y; const float* V;
    int B, N, M, C, Dv, topk;
    float alpha;
    CandBuffers cb;
    const int* row_list; const int* row_count;     // optional subset of rows
    const float* err_x; const float* err_ymax;     // candidate-pass distance error bounds (may be null)
    const float* tc_xx; const float* tc_yymax;     // |x~|^2 per row, max |y~|^2 per batch: scale of the tensor-core
                                                   // accumulation error of a candidate key (null for the fp32 pass)
    float rel_bound;                               // + rel_bound * d16
    int* flag_list; int* flag_count;               // rows failing the certificate (may be null)
    float* flag_thr; float* flag_r;                // per flag slot: squared scan threshold, reference distance (may be null)
    int* tie_count;                                // counts uncertified rows when flag_list is null
    int* nonfinite_count;                          // rows whose 16-bit softmax mass came out non-finite (settled exactly)
    int64_t* argmin; int* top_idx; float* top_w; float* top_d; float* row_min; float* row_sum; float* PiV;
};

// Exact fp32 re-scoring of up to 32 selected columns of row g (lane s holds sel_idx, -1 = none), ordering
// ((d^2, idx) lexicographic), weights relative to the exact row minimum, outputs and Pi.V.
// l_other = softmax mass of every column NOT among the selected ones, relative to distance r_other.
// Returns dk = exact distance of rank topk-1 (INFINITY if fewer were selected) and dmin.
template <bool kSoft>
__device__ __forceinline__ void rescore_emit(const FinalizeArgs& a, int g, int b, int lane, int sel_idx, int n_max,
                                             float l_other, float r_other, float& dk_out, float& dmin_out) {
    const float a2 = a.alpha * kLog2e;
    // d2 = sum_c (x_c - y_c)^2, lane owns channels 4*lane + 128 t, butterfly sum
    float4 xv[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int c = 4 * lane + 128 * t;
        xv[t] = (c < a.C) ? __ldg(reinterpret_cast<const float4*>(a.X + (size_t)g * a.C + c)) : make_float4(0, 0, 0, 0);
    }
    float my_d2 = INFINITY; int my_idx = 0x7fffffff;
    const float* Yb = a.Y + (size_t)b * a.M * a.C;
    for (int s0 = 0; s0 < n_max; s0 += 16) {                              // 16 candidates per round
        if (__shfl_sync(0xffffffffu, sel_idx, s0) < 0) break;             // uniform: the selection is dense from lane 0
        float acc[16];
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            const int j = __shfl_sync(0xffffffffu, sel_idx, s0 + s);
            const float* yr = Yb + (size_t)(j < 0 ? 0 : j) * a.C;        // absent: any valid row, the sum is dropped below
            float v = 0.f;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int c = 4 * lane + 128 * t;
                if (c < a.C) {
                    const float4 yv = __ldg(reinterpret_cast<const float4*>(yr + c));
                    float d;
                    d = xv[t].x - yv.x; v = fmaf(d, d, v);
                    d = xv[t].y - yv.y; v = fmaf(d, d, v);
                    d = xv[t].z - yv.z; v = fmaf(d, d, v);
                    d = xv[t].w - yv.w; v = fmaf(d, d, v);
                }
            }
            acc[s] = v;
        }
        // 16 butterfly sums at once: at every step a lane keeps one half of its values and trades the other half with
        // its partner -- the same pairing and order as warp_sum (bit-identical totals), 16 shuffles instead of 80
        const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2;
        float u8[8], u4[4], u2[2];
#pragma unroll
        for (int c = 0; c < 8; ++c) u8[c] = (b16 ? acc[c + 8] : acc[c]) + __shfl_xor_sync(0xffffffffu, b16 ? acc[c] : acc[c + 8], 16);
#pragma unroll
        for (int c = 0; c < 4; ++c) u4[c] = (b8 ? u8[c + 4] : u8[c]) + __shfl_xor_sync(0xffffffffu, b8 ? u8[c] : u8[c + 4], 8);
#pragma unroll
        for (int c = 0; c < 2; ++c) u2[c] = (b4 ? u4[c + 2] : u4[c]) + __shfl_xor_sync(0xffffffffu, b4 ? u4[c] : u4[c + 2], 4);
        float u1 = (b2 ? u2[1] : u2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? u2[0] : u2[1], 2);
        u1 += __shfl_xor_sync(0xffffffffu, u1, 1);                        // lanes 2c, 2c+1 hold candidate s0 + c
        const float mine = __shfl_sync(0xffffffffu, u1, 2 * (lane & 15));
        if ((lane & ~15) == s0 && sel_idx >= 0) { my_d2 = mine; my_idx = sel_idx; }
    }

    // ---- rank the exact scores: (d2, idx) lexicographic
    int rank = 0;
    for (int s = 0; s < n_max; ++s) {
        const float od = __shfl_sync(0xffffffffu, my_d2, s);
        const int oi = __shfl_sync(0xffffffffu, my_idx, s);
        rank += kv_less(od, oi, my_d2, my_idx) ? 1 : 0;
    }
    const bool valid = lane < n_max && my_idx != 0x7fffffff;
    if (!valid) rank = 99;
    const float my_d = valid ? sqrtf(my_d2) : INFINITY;
    float dmin = my_d;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));

    float my_w = 0.f, total = 1.f;
    if (kSoft) {
        const float e = valid ? expf(-a.alpha * (my_d - dmin)) : 0.f;
        total = warp_sum(e);
        if (l_other != 0.f) total += l_other * exp2f(-a2 * (r_other - dmin));
        my_w = e / total;
    }

    if (valid && rank < a.topk) {
        const size_t o = (size_t)g * a.topk + rank;
        a.top_idx[o] = my_idx;
        a.top_d[o] = my_d;
        if (a.top_w) a.top_w[o] = my_w;
        if (rank == 0) {
            if (a.argmin) a.argmin[g] = my_idx;
            if (a.row_min) a.row_min[g] = my_d;
            if (a.row_sum) a.row_sum[g] = total;
        }
    }
    {
        const unsigned m = __ballot_sync(0xffffffffu, rank == a.topk - 1);
        dk_out = m ? __shfl_sync(0xffffffffu, my_d, __ffs(m) - 1) : INFINITY;
        dmin_out = dmin;
    }

    // ---- Pi . V  (10-sparse gather), lanes over the Dv output channels
    if (kSoft && a.V && a.PiV) {
        float wk[DVM_TOPK_MAX]; int jk[DVM_TOPK_MAX];
#pragma unroll
        for (int k = 0; k < DVM_TOPK_MAX; ++k) {
            const unsigned m = __ballot_sync(0xffffffffu, rank == k);
            const int src = m ? __ffs(m) - 1 : 0;
            wk[k] = m ? __shfl_sync(0xffffffffu, my_w, src) : 0.f;
            jk[k] = m ? __shfl_sync(0xffffffffu, my_idx, src) : 0;
            if (k >= a.topk) wk[k] = 0.f;
        }
        const float* Vb = a.V + (size_t)b * a.M * a.Dv;
        for (int dv = lane; dv < a.Dv; dv += 32) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < DVM_TOPK_MAX; ++k)
                if (wk[k] != 0.f) acc = fmaf(wk[k], __ldg(Vb + (size_t)jk[k] * a.Dv + dv), acc);
            a.PiV[(size_t)g * a.Dv + dv] = acc;
        }
    }
}

// float <-> unsigned with the same ordering (-0 is folded onto +0 so that equal keys tie on the index)
__device__ __forceinline__ unsigned f32_ordered(float f) {
    unsigned u = __float_as_uint(f);
    if (u == 0x80000000u) u = 0u;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_from_ordered(unsigned o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

template <bool kSoft>
__global__ void __launch_bounds__(FIN_WARPS * 32) softmap_finalize_kernel(FinalizeArgs a) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * FIN_WARPS + (threadIdx.x >> 5);
    int g;
    if (a.row_list) {
        if (w >= *a.row_count) return;
        g = a.row_list[w];
    } else {
        if (w >= a.B * a.N) return;
        g = w;
    }
    const int b = g / a.N;
    const int P = a.cb.P;
    const int E = P * KC;                       // <= 128
    const float a2 = a.alpha * kLog2e;

    // ---- load the partial lists: entry e = lane + 32 q, as (order-preserving key bits, index) with "absent" = ~0
    unsigned eo[4]; int ei[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int e = lane + 32 * q;
        eo[q] = 0xffffffffu; ei[q] = 0x7fffffff;
        if (e < E) {
            const int j = a.cb.idx[(size_t)g * E + e];
            if (j >= 0) { eo[q] = f32_ordered(a.cb.key[(size_t)g * E + e]); ei[q] = j; }
        }
    }
    float l_tot = 0.f, r_star = INFINITY;
    if (kSoft) {
        float rp = INFINITY, lp = 0.f;
        if (lane < P) { rp = a.cb.r[(size_t)g * P + lane]; lp = a.cb.l[(size_t)g * P + lane]; }
        r_star = rp;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r_star = fminf(r_star, __shfl_xor_sync(0xffffffffu, r_star, o));
        float t = (lp != 0.f) ? lp * exp2f(-a2 * (rp - r_star)) : 0.f;
        l_tot = warp_sum(t);
    }

    // ---- select the KC best of the union by (key, idx): every lane keeps its (up to 4) entries sorted, one round =
    //      two warp-wide integer min reductions (redux.sync) over the lanes' heads; the owner pops its head
#define DVM_CSWAP(i, j) { const bool sw = eo[j] < eo[i] || (eo[j] == eo[i] && ei[j] < ei[i]); \
        const unsigned tk = sw ? eo[j] : eo[i]; const int ti = sw ? ei[j] : ei[i]; \
        eo[j] = sw ? eo[i] : eo[j]; ei[j] = sw ? ei[i] : ei[j]; eo[i] = tk; ei[i] = ti; }
    DVM_CSWAP(0, 1) DVM_CSWAP(2, 3) DVM_CSWAP(0, 2) DVM_CSWAP(1, 3) DVM_CSWAP(1, 2)
#undef DVM_CSWAP
    float sel_key = INFINITY; int sel_idx = -1;     // lane s (< KC) holds the s-th selected
    for (int s = 0; s < KC; ++s) {
        const unsigned wk = __reduce_min_sync(0xffffffffu, eo[0]);
        if (wk == 0xffffffffu) break;                                 // union exhausted (uniform)
        const int wi = __reduce_min_sync(0xffffffffu, eo[0] == wk ? ei[0] : 0x7fffffff);
        if (eo[0] == wk && ei[0] == wi) {                             // the owner retires it (indices are unique)
            eo[0] = eo[1]; ei[0] = ei[1]; eo[1] = eo[2]; ei[1] = ei[2]; eo[2] = eo[3]; ei[2] = ei[3]; eo[3] = 0xffffffffu; ei[3] = 0x7fffffff;
        }
        if (lane == s) { sel_key = f32_from_ordered(wk); sel_idx = wi; }
    }
    float key16 = __shfl_sync(0xffffffffu, sel_key, KC - 1);
    float tp_min;
    {                                                                 // discard bounds of the partial lists
        float tp = lane < P ? a.cb.t[(size_t)g * P + lane] : INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tp = fminf(tp, __shfl_xor_sync(0xffffffffu, tp, o));
        tp_min = tp;
        key16 = fminf(key16, tp);
    }
    const float l_base = l_tot;                                       // mass of everything outside the partial lists
    if (kSoft) {                                                      // listed but not selected: approximate terms
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (eo[q] != 0xffffffffu) t += exp2f(-a2 * (sqrtf(f32_from_ordered(eo[q])) - r_star));
        l_tot += warp_sum(t);
    }

    float dk, dmin;
    rescore_emit<kSoft>(a, g, b, lane, sel_idx, KC, l_tot, r_star, dk, dmin);

    // ---- certificate: every discarded column has candidate-pass key >= key16, i.e. a true distance of at least
    //      sqrt(key16 - E2) - (|x - x~| + |y - y~|) - rel * d16, E2 = accumulation error of the tensor-core key
    float bound = 0.f, e2 = 0.f;
    if (a.err_x) bound = a.err_x[g] + a.err_ymax[b];
    if (a.tc_xx) e2 = 4e-6f * (a.tc_xx[g] + a.tc_yymax[b]);          // MMA accumulation + 16 ulp of packed list keys
    auto certified = [&](float dkk, float kb) {
        if (kb == INFINITY) return bound < INFINITY;                  // every column is in the list (M < KC) unless the
        const float db = sqrtf(fmaxf(kb - e2, 0.f));                  // 16-bit conversion overflowed
        return dkk < db - (bound + a.rel_bound * db);
    };
    bool ok = certified(dk, key16);
    // ---- second chance (warp-uniform): the union of the partial lists holds more than the KC re-scored candidates.  Re-score the
    //      next KC of them too: the bound on what was DISCARDED then is the lists' discard bound (and the key of a 33rd entry, if
    //      any) instead of the 16th candidate's key -- about twice the gap, for 32 more row loads of a row that would otherwise
    //      cost a threshold scan over all M columns.
    if (!ok && __any_sync(0xffffffffu, eo[0] != 0xffffffffu)) {
        for (int s = KC; s < 2 * KC; ++s) {
            const unsigned wk = __reduce_min_sync(0xffffffffu, eo[0]);
            if (wk == 0xffffffffu) break;
            const int wi = __reduce_min_sync(0xffffffffu, eo[0] == wk ? ei[0] : 0x7fffffff);
            if (eo[0] == wk && ei[0] == wi) {
                eo[0] = eo[1]; ei[0] = ei[1]; eo[1] = eo[2]; ei[1] = ei[2]; eo[2] = eo[3]; ei[2] = ei[3]; eo[3] = 0xffffffffu; ei[3] = 0x7fffffff;
            }
            if (lane == s) { sel_key = f32_from_ordered(wk); sel_idx = wi; }
        }
        const unsigned rest = __reduce_min_sync(0xffffffffu, eo[0]);   // best key still unselected (a 33rd entry), if any
        const float key_rest = fminf(rest == 0xffffffffu ? INFINITY : f32_from_ordered(rest), tp_min);
        float l2 = l_base;
        if (kSoft) {
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (eo[q] != 0xffffffffu) t += exp2f(-a2 * (sqrtf(f32_from_ordered(eo[q])) - r_star));
            l2 += warp_sum(t);
        }
        rescore_emit<kSoft>(a, g, b, lane, sel_idx, 2 * KC, l2, r_star, dk, dmin);
        l_tot = l2;
        ok = certified(dk, key_rest);
    }
    if (lane == 0) {
        // the 16-bit pass sums its softmax mass against a FIXED per-row reference (the priming pass' sampled minimum): a row
        // without a finite sample, or one whose terms overflowed against it, has a non-finite mass and is settled exactly
        if (kSoft && !(l_tot < INFINITY)) { ok = false; if (a.nonfinite_count) atomicAdd(a.nonfinite_count, 1); }
        if (!ok) {
            if (a.flag_list) {
                const int e = atomicAdd(a.flag_count, 1);
                a.flag_list[e] = g;
                // every column that can still belong to the top-k has an exact distance <= dk; the rescue scan
                // evaluates d^2 in a different summation order (relative difference <= 1.6e-5), hence the slack
                if (a.flag_thr) { a.flag_thr[e] = dk * dk * (1.f + 3e-5f); a.flag_r[e] = dmin; }
            } else if (a.tie_count) atomicAdd(a.tie_count, 1);
        }
    }
}

// =================================================================================================
// 2b. rescue path for the (few) rows the 16-bit certificate rejects.  For such a row every column that can
//     still belong to the top-k has an exact distance <= dk (the exact k-th best among the 16 candidates), so
//     instead of a second top-k sweep the row needs a THRESHOLD SCAN: all columns with d^2 <= thr are listed
//     (typically k .. k+2 of them), every other column only adds its exact softmax term to the row's mass.
//     The scan is parallel over column chunks (grid = chunks x batch), deterministic (per-chunk partial
//     masses summed in chunk order) and reads Y once per 8 flagged rows.
// =================================================================================================
constexpr int RESC_MAX = 4096;       // flagged rows handled by the rescue scan; more -> fp32 candidate pass
constexpr int RESC_CAP = 32;         // listed columns per row; more (mass ties) -> fp32 pass for that row
constexpr int RESC_NCH_MAX = 2048;   // column chunks (M <= 262144 for the rescue scan; beyond: fp32 pass)
constexpr int RESC_ROWS = 32;        // flagged rows per group (a thread = 8 rows x 2 columns)

constexpr int RESC_CHUNK = 128;      // columns per CTA: the Y chunk is staged ONCE in shared memory (coalesced) and reused
                                     // for every flagged row of the batch element

template <bool kSoft>
__global__ void __launch_bounds__(256)
rescue_scan_kernel(const float* __restrict__ X, const float* __restrict__ Y, int N, int M, int C,
                   const int* __restrict__ flag_list, const int* __restrict__ flag_count,
                   const float* __restrict__ flag_thr, const float* __restrict__ flag_r, float a2,
                   int nch, int* __restrict__ resc_cnt, int* __restrict__ resc_idx, float* __restrict__ resc_mass) {
    extern __shared__ __align__(16) float rs_sm[];
    const int ld = C + 4;                                    // conflict-free float4 reads, one column per lane
    float* Ys = rs_sm;                                       // [RESC_CHUNK][ld]
    float* xs = Ys + RESC_CHUNK * ld;                        // [RESC_ROWS][C]
    __shared__ int s_slots[RESC_MAX];
    __shared__ int s_n;
    __shared__ float s_thr[RESC_ROWS], s_ref[RESC_ROWS];
    __shared__ float s_part[2][RESC_ROWS];
    const int count = *flag_count;
    if (count == 0 || count > RESC_MAX) return;
    const int b = blockIdx.y, ch = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int c0 = ch * RESC_CHUNK;
    const int ncol = min(RESC_CHUNK, M - c0);
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int e = tid; e < count; e += 256)
        if (flag_list[e] / N == b) s_slots[atomicAdd(&s_n, 1)] = e;
    __syncthreads();
    const int n = s_n;
    if (n == 0) return;
    const int c4 = C >> 2;
    const float* Yb = Y + ((size_t)b * M + c0) * C;
    for (int e = tid; e < ncol * c4; e += 256) {             // coalesced: consecutive threads read consecutive float4
        const int rr = e / c4, cc = e - rr * c4;
        *reinterpret_cast<float4*>(Ys + rr * ld + cc * 4) = __ldg(reinterpret_cast<const float4*>(Yb + (size_t)rr * C) + cc);
    }
    const int col = tid & 63;                                // this thread's columns of the chunk: col, col + 64
    const int rg = tid >> 6;                                 // ... and its quarter of the row group (rows rg*8 .. +7)
    for (int g0 = 0; g0 < n; g0 += RESC_ROWS) {
        __syncthreads();
        for (int e = tid; e < RESC_ROWS * c4; e += 256) {
            const int r = e / c4, cc = e - r * c4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g0 + r < n) v = __ldg(reinterpret_cast<const float4*>(X + (size_t)flag_list[s_slots[g0 + r]] * C) + cc);
            *reinterpret_cast<float4*>(xs + r * C + cc * 4) = v;
        }
        if (tid < RESC_ROWS) {
            const bool ok = g0 + tid < n;
            s_thr[tid] = ok ? flag_thr[s_slots[g0 + tid]] : -1.f;
            s_ref[tid] = ok ? flag_r[s_slots[g0 + tid]] : 0.f;
        }
        __syncthreads();
        if (g0 + rg * 8 >= n) continue;                      // warp-uniform (and no barrier below for hard maps) ...
        float acc[8][2];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r][0] = acc[r][1] = 0.f;
        const float* yp = Ys + col * ld;
        const float* xp = xs + (rg * 8) * C;
        for (int k = 0; k < C; k += 4) {
            const float4 y0 = *reinterpret_cast<const float4*>(yp + k);
            const float4 y1 = *reinterpret_cast<const float4*>(yp + 64 * ld + k);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float4 xv = *reinterpret_cast<const float4*>(xp + r * C + k);      // warp-broadcast
                float d;
                d = xv.x - y0.x; acc[r][0] = fmaf(d, d, acc[r][0]);
                d = xv.y - y0.y; acc[r][0] = fmaf(d, d, acc[r][0]);
                d = xv.z - y0.z; acc[r][0] = fmaf(d, d, acc[r][0]);
                d = xv.w - y0.w; acc[r][0] = fmaf(d, d, acc[r][0]);
                d = xv.x - y1.x; acc[r][1] = fmaf(d, d, acc[r][1]);
                d = xv.y - y1.y; acc[r][1] = fmaf(d, d, acc[r][1]);
                d = xv.z - y1.z; acc[r][1] = fmaf(d, d, acc[r][1]);
                d = xv.w - y1.w; acc[r][1] = fmaf(d, d, acc[r][1]);
            }
        }
        float mass[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int rr = rg * 8 + r;
            mass[r] = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (col + 64 * h < ncol && g0 + rr < n) {
                    if (acc[r][h] <= s_thr[rr]) {
                        const int slot = s_slots[g0 + rr];
                        const int q = atomicAdd(resc_cnt + slot, 1);
                        if (q < RESC_CAP) resc_idx[(size_t)slot * RESC_CAP + q] = c0 + col + 64 * h;
                    } else if (kSoft) {
                        mass[r] += exp2f(-a2 * (sqrtf(acc[r][h]) - s_ref[rr]));
                    }
                }
            }
        }
        if (kSoft) {
            // deterministic: lanes -> warp sum, then the 2 warps of a row quarter in fixed order (named barrier per quarter:
            // quarters without rows skipped the pass)
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float v = warp_sum(mass[r]);
                if (lane == 0) s_part[wid & 1][rg * 8 + r] = v;
            }
            asm volatile("bar.sync %0, 64;" ::"r"(rg + 1) : "memory");
            const int t64 = tid & 63;
            if (t64 < 8 && g0 + rg * 8 + t64 < n)
                resc_mass[(size_t)s_slots[g0 + rg * 8 + t64] * nch + ch] = s_part[0][rg * 8 + t64] + s_part[1][rg * 8 + t64];
        }
    }
}

struct RescueArgs {
    const int* flag_list; const int* flag_count; const float* flag_r;
    const int* resc_cnt; const int* resc_idx; const float* resc_mass; int nch;
    int* list2; int* count2;                       // rows that still need the fp32 candidate pass
    int active;                                    // 0: the scan did not run (M too large): hand everything to the fp32 pass
};

template <bool kSoft>
__global__ void __launch_bounds__(FIN_WARPS * 32) rescue_finalize_kernel(FinalizeArgs a, RescueArgs r) {
    const int count = *r.flag_count;
    if (count == 0) return;
    if (count > RESC_MAX || !r.active) {           // too many (or no scan): hand every flagged row to the fp32 pass
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) r.list2[e] = r.flag_list[e];
        if (blockIdx.x == 0 && threadIdx.x == 0) *r.count2 = count;
        return;
    }
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * FIN_WARPS + (threadIdx.x >> 5);
    if (w >= count) return;
    const int g = r.flag_list[w];
    const int n = r.resc_cnt[w];
    if (n > RESC_CAP) {
        if (lane == 0) r.list2[atomicAdd(r.count2, 1)] = g;
        return;
    }
    const int sel_idx = lane < n ? r.resc_idx[(size_t)w * RESC_CAP + lane] : -1;
    float l_other = 0.f;
    if (kSoft) {
        float t = 0.f;
        for (int c = lane; c < r.nch; c += 32) t += r.resc_mass[(size_t)w * r.nch + c];
        l_other = warp_sum(t);
    }
    float dk, dmin;
    rescore_emit<kSoft>(a, g, g / a.N, lane, sel_idx, RESC_CAP, l_other, r.flag_r[w], dk, dmin);
}

static int launch_finalize(const FinalizeArgs& a, bool soft, int max_rows, cudaStream_t st) {
    const int grid = ceil_div(max_rows, FIN_WARPS);
    if (soft) softmap_finalize_kernel<true><<<grid, FIN_WARPS * 32, 0, st>>>(a);
    else      softmap_finalize_kernel<false><<<grid, FIN_WARPS * 32, 0, st>>>(a);
    DVM_LAUNCH_CHECK();
    return 0;
}

static void carve_cand(WsCarver& ws, size_t rows, int P, CandBuffers& cb) {
    cb.P = P;
    cb.key = ws.take<float>(rows * P * KC);
    cb.idx = ws.take<int>(rows * P * KC);
    cb.l = ws.take<float>(rows * P);
    cb.r = ws.take<float>(rows * P);
    cb.t = ws.take<float>(rows * P);
}

struct RescueWs { float* flag_thr; float* flag_r; int* cnt; int* idx; float* mass; int* list2; };

static size_t softmap_ws_layout(void* base, size_t cap, int B, int N, int M, int C, int prec,
                                CandBuffers* simt, CandBuffers* tc, int** flag_list, int** stats_fallback,
                                float** err_x, float** err_ymax, void** tc_ws, size_t* tc_ws_bytes, RescueWs* resc = nullptr) {
    WsCarver ws(base, cap);
    const size_t rows = (size_t)B * N;
    CandBuffers c1{}, c2{};
    carve_cand(ws, rows, SIMT_CG, c1);
    int* fl = nullptr; float* ex = nullptr; float* ey = nullptr; void* tws = nullptr; size_t tb = 0;
    int* sf = ws.take<int>(4);
    if (prec != DVM_PREC_FP32) {
        carve_cand(ws, rows, tc_num_partials(B, N, M), c2);
        fl = ws.take<int>(rows);
        ex = ws.take<float>(rows);
        ey = ws.take<float>(B);
        tb = tc_workspace_bytes(B, N, M, C);
        tws = ws.take<char>(tb);
        RescueWs rw;
        rw.flag_thr = ws.take<float>(rows);
        rw.flag_r = ws.take<float>(rows);
        rw.list2 = ws.take<int>(rows);
        rw.cnt = ws.take<int>(RESC_MAX);
        rw.idx = ws.take<int>((size_t)RESC_MAX * RESC_CAP);
        { int nchw = ceil_div(M, RESC_CHUNK); if (nchw > RESC_NCH_MAX) nchw = 1; rw.mass = ws.take<float>((size_t)RESC_MAX * nchw); }
        if (resc) *resc = rw;
    }
    if (simt) *simt = c1;
    if (tc) *tc = c2;
    if (flag_list) *flag_list = fl;
    if (stats_fallback) *stats_fallback = sf;
    if (err_x) *err_x = ex;
    if (err_ymax) *err_ymax = ey;
    if (tc_ws) *tc_ws = tws;
    if (tc_ws_bytes) *tc_ws_bytes = tb;
    return align_up(ws.off, 256);
}

}  // namespace dvm

using namespace dvm;

extern "C" size_t dvm_softmap_workspace_bytes(int B, int N, int M, int C, int prec) {
    if (B <= 0 || N <= 0 || M <= 0 || C <= 0) return 0;
    if (C > 128) prec = DVM_PREC_FP32;
    return softmap_ws_layout(nullptr, 0, B, N, M, C, prec, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

static int softmap_fwd_impl(const float* X, const float* Y, const float* V,
                            int B, int N, int M, int C, int Dv, float alpha, int topk, int mode, int prec,
                            int64_t* argmin, int32_t* top_idx, float* top_w, float* top_d,
                            float* row_min, float* row_sum, float* PiV, int32_t* stats,
                            void* ws, size_t ws_bytes, cudaStream_t st) {
    DVM_CHECK_ARG(X && Y && top_idx && top_d, "dvm_softmap_fwd: X, Y, top_idx, top_d must be non-null");
    DVM_CHECK_ARG(B > 0 && N > 0 && M > 0, "dvm_softmap_fwd: empty problem (B=%d N=%d M=%d)", B, N, M);
    DVM_CHECK_ARG(C > 0 && C % 4 == 0 && C <= 256, "dvm_softmap_fwd: C=%d must be a multiple of 4 and <= 256", C);
    DVM_CHECK_ARG(topk >= 1 && topk <= DVM_TOPK_MAX && topk <= M, "dvm_softmap_fwd: topk=%d must be in [1, min(10, M=%d)]", topk, M);
    DVM_CHECK_ARG(alpha >= 0.f && isfinite(alpha), "dvm_softmap_fwd: alpha=%f must be finite and >= 0", (double)alpha);
    DVM_CHECK_ARG(mode == DVM_MODE_HARD || mode == DVM_MODE_SOFT, "dvm_softmap_fwd: bad mode %d", mode);
    DVM_CHECK_ARG(prec == DVM_PREC_FP32 || prec == DVM_PREC_F16 || prec == DVM_PREC_BF16, "dvm_softmap_fwd: bad prec %d", prec);
    DVM_CHECK_ARG((V == nullptr) == (Dv == 0) || PiV == nullptr, "dvm_softmap_fwd: V/Dv mismatch");
    DVM_CHECK_ARG((long long)B * N < 0x7fffffffLL && (long long)B * M < 0x7fffffffLL, "dvm_softmap_fwd: too many rows");
    const bool soft = mode == DVM_MODE_SOFT;
    DVM_CHECK_ARG(!soft || top_w, "dvm_softmap_fwd: soft mode needs top_w");
    // the tcgen05 pass keeps a whole K = C operand row block resident in shared memory: C <= 128.
    // Wider features take the (more precise) fp32 pass.
    if (C > 128) prec = DVM_PREC_FP32;

    CandBuffers simt{}, tc{};
    int* flag_list; int* sfb; float* err_x; float* err_ymax; void* tws; size_t tws_bytes; RescueWs rw{};
    const size_t need = softmap_ws_layout(ws, ws_bytes, B, N, M, C, prec, &simt, &tc, &flag_list, &sfb,
                                          &err_x, &err_ymax, &tws, &tws_bytes, &rw);
    if (!ws || need > ws_bytes) {
        set_error("dvm_softmap_fwd: workspace too small (%zu < %zu)", ws_bytes, need);
        return DVM_ERR_WORKSPACE;
    }
    int* st_out = stats ? stats : sfb;
    DVM_CUDA(cudaMemsetAsync(st_out, 0, 4 * sizeof(int), st));

    FinalizeArgs fa{};
    fa.X = X; fa.Y = Y; fa.V = V; fa.B = B; fa.N = N; fa.M = M; fa.C = C; fa.Dv = Dv; fa.topk = topk; fa.alpha = alpha;
    fa.argmin = argmin; fa.top_idx = top_idx; fa.top_w = top_w; fa.top_d = top_d;
    fa.row_min = row_min; fa.row_sum = row_sum; fa.PiV = PiV;
    const int rows = B * N;
    int rc;
    if (prec == DVM_PREC_FP32) {
        if ((rc = launch_cand_simt(X, Y, B, N, M, C, alpha, soft, nullptr, nullptr, rows, simt, st))) return rc;
        fa.cb = simt; fa.rel_bound = 1e-5f; fa.tie_count = st_out + 1;
        return launch_finalize(fa, soft, rows, st);
    }
    if ((rc = launch_cand_tc(X, Y, B, N, M, C, alpha, soft, prec, tc, err_x, err_ymax, &fa.tc_xx, &fa.tc_yymax, tws, tws_bytes, st))) return rc;
    fa.cb = tc; fa.rel_bound = 2e-5f; fa.err_x = err_x; fa.err_ymax = err_ymax;
    fa.flag_list = flag_list; fa.flag_count = st_out; fa.flag_thr = rw.flag_thr; fa.flag_r = rw.flag_r; fa.nonfinite_count = st_out + 3;
    DVM_CUDA(cudaMemsetAsync(rw.cnt, 0, RESC_MAX * sizeof(int), st));
    if ((rc = launch_finalize(fa, soft, rows, st))) return rc;
    // rows the certificate rejected (count lives on the device: no host sync, every kernel below exits at once
    // when it has nothing to do): threshold scan + exact emit; what even that cannot settle (> RESC_CAP columns
    // inside the threshold, or > RESC_MAX flagged rows) goes to the fp32 candidate pass.
    {
        const int nch = ceil_div(M, RESC_CHUNK);
        const float a2 = alpha * kLog2e;
        if (nch <= RESC_NCH_MAX) {
            const size_t rsm = ((size_t)RESC_CHUNK * (C + 4) + (size_t)RESC_ROWS * C) * sizeof(float);
            static PerDeviceOnce attr_done;
            if (attr_done.need()) {
                DVM_CUDA(cudaFuncSetAttribute(rescue_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
                DVM_CUDA(cudaFuncSetAttribute(rescue_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
                attr_done.done();
            }
            dim3 grid(nch, B);
            if (soft) rescue_scan_kernel<true><<<grid, 256, rsm, st>>>(X, Y, N, M, C, flag_list, st_out, rw.flag_thr, rw.flag_r, a2, nch, rw.cnt, rw.idx, rw.mass);
            else      rescue_scan_kernel<false><<<grid, 256, rsm, st>>>(X, Y, N, M, C, flag_list, st_out, rw.flag_thr, rw.flag_r, a2, nch, rw.cnt, rw.idx, rw.mass);
            DVM_LAUNCH_CHECK();
        }
        RescueArgs ra{flag_list, st_out, rw.flag_r, rw.cnt, rw.idx, rw.mass, nch, rw.list2, st_out + 2, nch <= RESC_NCH_MAX ? 1 : 0};
        FinalizeArgs fr = fa;
        fr.flag_list = nullptr; fr.flag_count = nullptr; fr.flag_thr = nullptr; fr.flag_r = nullptr; fr.tie_count = nullptr; fr.nonfinite_count = nullptr;
        const int fgrid = ceil_div(RESC_MAX, FIN_WARPS);
        if (soft) rescue_finalize_kernel<true><<<fgrid, FIN_WARPS * 32, 0, st>>>(fr, ra);
        else      rescue_finalize_kernel<false><<<fgrid, FIN_WARPS * 32, 0, st>>>(fr, ra);
        DVM_LAUNCH_CHECK();
    }
    int* list2 = rw.list2; int* count2 = st_out + 2;
    if ((rc = launch_rows_exact(X, Y, N, M, C, alpha, soft, list2, count2, rows, simt, st))) return rc;
    if ((rc = launch_cand_simt(X, Y, B, N, M, C, alpha, soft, list2, count2, rows, simt, st))) return rc;
    fa.cb = simt; fa.rel_bound = 1e-5f; fa.err_x = nullptr; fa.err_ymax = nullptr; fa.tc_xx = nullptr; fa.tc_yymax = nullptr;
    fa.row_list = list2; fa.row_count = count2; fa.flag_list = nullptr; fa.flag_count = nullptr; fa.flag_thr = nullptr; fa.flag_r = nullptr;
    fa.tie_count = st_out + 1; fa.nonfinite_count = nullptr;
    return launch_finalize(fa, soft, rows, st);
}

extern "C" int dvm_softmap_fwd(const float* X, const float* Y, const float* V,
                               int B, int N, int M, int C, int Dv, float alpha, int topk, int mode, int prec,
                               int64_t* argmin, int32_t* top_idx, float* top_w, float* top_d,
                               float* row_min, float* row_sum, float* PiV, int32_t* stats,
                               void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    prof_begin(st, 1);                 // channel 1: the whole fused op (what north_star's tensor-peak target is quoted on)
    const int rc = softmap_fwd_impl(X, Y, V, B, N, M, C, Dv, alpha, topk, mode, prec, argmin, top_idx, top_w, top_d,
                                    row_min, row_sum, PiV, stats, ws, ws_bytes, st);
    prof_end(st, 1);
    return rc;
}
