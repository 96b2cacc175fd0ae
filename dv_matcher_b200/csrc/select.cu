// Large-k selection, gathered pair distances and row gathers: the kernels of the dist loss (models/loss.py:1351-1396) and of
// knn / index_points (models/loss.py:451-473) beyond k = 10.
//
//   dvm_topk_select   per row of a score matrix: indices of the k LARGEST scores, descending (ties -> lower index), k <= 1024.
//                     `knn(a, b, k)` with k = 500 / 300: the scores 2 a.b - |b|^2 come from the tcgen05 3xTF32 GEMM
//                     (dvm_linear_act_fwd with W = b, bias = -|b|^2 / 2), the selection is a 4-pass 8-bit radix select over
//                     order-preserving keys + a bitonic sort of the k survivors -- no full row sort, one CTA per row.
//   dvm_pair_dist_*   d[b,s,t] = | feat[b, nbr[b,s,t]] - feat[b, q[s]] |_2 (direct differences) and, fused, the geodesic entries
//                     geo[b, nbr[b,s,t], q[s]] -- replaces index_points + torch.norm (a [B,S,k,C] tensor: 256 MB per shape at
//                     S = 1000, k = 500) and the Python loop over B of models/loss.py:1376-1378.  Backward: (g / d)(f_nbr - f_q)
//                     scattered with vector atomics; 0 where d = 0 (torch.norm's subgradient).
//   dvm_gather_rows_* index_points(points, idx) forward / backward for any row width.
#include "common.cuh"

namespace dvm {

// ------------------------------------------------------------------------------------------------
// top-k selection
// ------------------------------------------------------------------------------------------------
constexpr int SEL_THREADS = 256;
constexpr int SEL_KMAX = 1024;

__device__ __forceinline__ unsigned sel_ordered(float f) {          // monotone: larger float <-> larger unsigned (-0 == +0)
    unsigned u = __float_as_uint(f);
    if (u == 0x80000000u) u = 0u;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(SEL_THREADS)
topk_select_kernel(const float* __restrict__ scores, long long pitch, int N, int k, int kpad, int64_t* __restrict__ out) {
    __shared__ int hist[256];
    __shared__ unsigned s_key[SEL_KMAX];
    __shared__ int s_idx[SEL_KMAX];
    __shared__ unsigned s_prefix, s_mask;
    __shared__ int s_remaining, s_cnt, s_eq_base, s_warp_off[SEL_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float* s = scores + (size_t)blockIdx.x * pitch;
    if (tid == 0) { s_prefix = 0u; s_mask = 0u; s_remaining = k; s_cnt = 0; s_eq_base = 0; }
    // ---- 4-pass radix select of the k-th largest key
    for (int pass = 3; pass >= 0; --pass) {
        hist[tid] = 0;
        __syncthreads();
        const unsigned prefix = s_prefix, mask = s_mask;
        for (int j = tid; j < N; j += SEL_THREADS) {
            const unsigned u = sel_ordered(__ldg(s + j));
            if ((u & mask) == prefix) atomicAdd(&hist[(u >> (8 * pass)) & 255u], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int rem = s_remaining, b = 255;
            for (; b > 0; --b) { if (hist[b] >= rem) break; rem -= hist[b]; }
            s_remaining = rem;                                    // how many of bucket b are still needed
            s_prefix = prefix | ((unsigned)b << (8 * pass));
            s_mask = mask | (255u << (8 * pass));
        }
        __syncthreads();
    }
    const unsigned T = s_prefix;                                  // key of the k-th largest score
    const int need_eq = s_remaining;                              // how many entries equal to T belong to the selection
    for (int e = tid; e < kpad; e += SEL_THREADS) { s_key[e] = 0u; s_idx[e] = 0x7fffffff; }     // padding sorts last
    __syncthreads();
    // ---- collect: every key > T (any order), and the need_eq LOWEST indices among the keys == T (ordered chunk scan)
    for (int j0 = 0; j0 < N; j0 += SEL_THREADS) {
        const int j = j0 + tid;
        const unsigned u = j < N ? sel_ordered(__ldg(s + j)) : 0u;
        const bool gt = j < N && u > T, eq = j < N && u == T;
        if (gt) { const int pos = atomicAdd(&s_cnt, 1); s_key[pos] = u; s_idx[pos] = j; }
        const unsigned m = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) s_warp_off[wid] = __popc(m);
        __syncthreads();
        int off = s_eq_base;
        for (int w = 0; w < wid; ++w) off += s_warp_off[w];
        const int my = off + __popc(m & ((1u << lane) - 1u));
        if (eq && my < need_eq) { const int pos = k - 1 - my; s_key[pos] = u; s_idx[pos] = j; }    // the ties fill the tail
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < SEL_THREADS / 32; ++w) t += s_warp_off[w]; s_eq_base += t; }
        __syncthreads();
    }
    // ---- bitonic sort of kpad entries: descending key, ascending index on ties
    for (int size = 2; size <= kpad; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int e = tid; e < kpad / 2; e += SEL_THREADS) {
                const int i = 2 * e - (e & (stride - 1));         // lower element of the pair
                const int p = i + stride;
                const bool up = ((i & size) == 0);                // "up" blocks end with the better element first
                const unsigned ki = s_key[i], kp = s_key[p];
                const int ii = s_idx[i], ip = s_idx[p];
                const bool p_better = kp > ki || (kp == ki && ip < ii);
                if (p_better == up) { s_key[i] = kp; s_key[p] = ki; s_idx[i] = ip; s_idx[p] = ii; }
            }
            __syncthreads();
        }
    for (int e = tid; e < k; e += SEL_THREADS) out[(size_t)blockIdx.x * k + e] = s_idx[e];
}

// ------------------------------------------------------------------------------------------------
// gathered pair distances: 8 lanes per (query, neighbour) pair, 4 pairs per warp iteration
// ------------------------------------------------------------------------------------------------
constexpr int PD_CMAX4 = 16;          // float4 per lane: C <= 8 * 4 * 16 = 512

template <typename GeoT>
__global__ void __launch_bounds__(256)
pair_dist_fwd_kernel(const float* __restrict__ feat, const int64_t* __restrict__ qidx, const int64_t* __restrict__ nbr,
                     int B, int N, int C, int S, int k, float* __restrict__ d_out,
                     const GeoT* __restrict__ geo, float* __restrict__ geo_out) {
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5);           // (b, s)
    if (w >= B * S) return;
    const int lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
    const int b = w / S, s = w - b * S;
    const int64_t q = __ldg(qidx + s);
    const float* Fb = feat + (size_t)b * N * C;
    const int c4 = C >> 2, per = (c4 + 7) >> 3;                   // float4 per lane
    float4 fq[PD_CMAX4];
#pragma unroll
    for (int t = 0; t < PD_CMAX4; ++t)
        if (t < per) { const int c = l8 + 8 * t; fq[t] = c < c4 ? __ldg(reinterpret_cast<const float4*>(Fb + (size_t)q * C) + c) : make_float4(0, 0, 0, 0); }
    for (int t0 = 0; t0 < k; t0 += 4) {
        const int tt = t0 + sub;
        const int64_t j = tt < k ? __ldg(nbr + ((size_t)w) * k + tt) : q;
        float acc = 0.f;
#pragma unroll
        for (int t = 0; t < PD_CMAX4; ++t)
            if (t < per) {
                const int c = l8 + 8 * t;
                if (c < c4) {
                    const float4 f = __ldg(reinterpret_cast<const float4*>(Fb + (size_t)j * C) + c);
                    float e;
                    e = f.x - fq[t].x; acc = fmaf(e, e, acc); e = f.y - fq[t].y; acc = fmaf(e, e, acc);
                    e = f.z - fq[t].z; acc = fmaf(e, e, acc); e = f.w - fq[t].w; acc = fmaf(e, e, acc);
                }
            }
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (l8 == 0 && tt < k) {
            d_out[(size_t)w * k + tt] = sqrtf(acc);
            if (geo) geo_out[(size_t)w * k + tt] = (float)geo[((size_t)b * N + (size_t)j) * N + (size_t)q];     // dist[b, idx, idx_num]
        }
    }
}

__global__ void __launch_bounds__(256)
pair_dist_bwd_kernel(const float* __restrict__ feat, const int64_t* __restrict__ qidx, const int64_t* __restrict__ nbr,
                     const float* __restrict__ d, const float* __restrict__ g, int B, int N, int C, int S, int k,
                     float* __restrict__ dfeat) {
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= B * S) return;
    const int lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
    const int b = w / S, s = w - b * S;
    const int64_t q = __ldg(qidx + s);
    const float* Fb = feat + (size_t)b * N * C;
    float* Gb = dfeat + (size_t)b * N * C;
    const int c4 = C >> 2, per = (c4 + 7) >> 3;
    float4 fq[PD_CMAX4], gq[PD_CMAX4];
#pragma unroll
    for (int t = 0; t < PD_CMAX4; ++t)
        if (t < per) {
            const int c = l8 + 8 * t;
            fq[t] = c < c4 ? __ldg(reinterpret_cast<const float4*>(Fb + (size_t)q * C) + c) : make_float4(0, 0, 0, 0);
            gq[t] = make_float4(0, 0, 0, 0);
        }
    for (int t0 = 0; t0 < k; t0 += 4) {
        const int tt = t0 + sub;
        if (tt >= k) continue;
        const int64_t j = __ldg(nbr + (size_t)w * k + tt);
        const float dd = __ldg(d + (size_t)w * k + tt);
        const float coef = dd > 0.f ? __ldg(g + (size_t)w * k + tt) / dd : 0.f;        // torch.norm's backward: 0 at d = 0
        if (coef == 0.f) continue;
#pragma unroll
        for (int t = 0; t < PD_CMAX4; ++t)
            if (t < per) {
                const int c = l8 + 8 * t;
                if (c < c4) {
                    const float4 f = __ldg(reinterpret_cast<const float4*>(Fb + (size_t)j * C) + c);
                    const float4 e = make_float4(coef * (f.x - fq[t].x), coef * (f.y - fq[t].y), coef * (f.z - fq[t].z), coef * (f.w - fq[t].w));
                    atomicAdd(reinterpret_cast<float4*>(Gb + (size_t)j * C) + c, e);
                    gq[t].x -= e.x; gq[t].y -= e.y; gq[t].z -= e.z; gq[t].w -= e.w;
                }
            }
    }
#pragma unroll
    for (int t = 0; t < PD_CMAX4; ++t)
        if (t < per) {
            const int c = l8 + 8 * t;
            if (c < c4) atomicAdd(reinterpret_cast<float4*>(Gb + (size_t)q * C) + c, gq[t]);
        }
}

// ------------------------------------------------------------------------------------------------
// index_points
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_rows_fwd_kernel(const float* __restrict__ pts, const int64_t* __restrict__ idx, long long rows, int R, int N, int C,
                       float* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // one element per thread (float4 when C % 4 == 0)
    if ((C & 3) == 0) {
        const int c4 = C >> 2;
        if (e >= rows * c4) return;
        const long long r = e / c4; const int c = (int)(e - r * c4);
        const long long b = r / R;
        reinterpret_cast<float4*>(out)[e] = __ldg(reinterpret_cast<const float4*>(pts + ((size_t)b * N + (size_t)__ldg(idx + r)) * C) + c);
    } else {
        if (e >= rows * C) return;
        const long long r = e / C; const int c = (int)(e - r * C);
        const long long b = r / R;
        out[e] = __ldg(pts + ((size_t)b * N + (size_t)__ldg(idx + r)) * C + c);
    }
}

__global__ void __launch_bounds__(256)
gather_rows_bwd_kernel(const float* __restrict__ d_out, const int64_t* __restrict__ idx, long long rows, int R, int N, int C,
                       float* __restrict__ d_pts) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * C) return;
    const long long r = e / C; const int c = (int)(e - r * C);
    const long long b = r / R;
    atomicAdd(d_pts + ((size_t)b * N + (size_t)__ldg(idx + r)) * C + c, __ldg(d_out + e));
}

}  // namespace dvm

using namespace dvm;

extern "C" int dvm_topk_select(const float* scores, long long rows, int N, long long pitch, int k, int64_t* idx, void* stream) {
    DVM_CHECK_ARG(scores && idx, "dvm_topk_select: null pointer");
    DVM_CHECK_ARG(rows >= 0 && rows < (1ll << 31) && N >= 1 && pitch >= N, "dvm_topk_select: bad shape rows=%lld N=%d pitch=%lld", rows, N, pitch);
    DVM_CHECK_ARG(k >= 1 && k <= SEL_KMAX && k <= N, "dvm_topk_select: k=%d must be in [1, min(%d, N=%d)]", k, SEL_KMAX, N);
    if (rows == 0) return 0;
    int kpad = 2; while (kpad < k) kpad <<= 1;
    topk_select_kernel<<<(unsigned)rows, SEL_THREADS, 0, (cudaStream_t)stream>>>(scores, pitch, N, k, kpad, idx);
    DVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int dvm_pair_dist_fwd(const float* feat, const int64_t* qidx, const int64_t* nbr, int B, int N, int C, int S, int k,
                                 float* d, const void* geo, int geo_is_f64, float* geo_out, void* stream) {
    DVM_CHECK_ARG(feat && qidx && nbr && d, "dvm_pair_dist_fwd: null pointer");
    DVM_CHECK_ARG(B > 0 && N > 0 && S > 0 && k > 0 && C > 0 && C % 4 == 0 && C <= 32 * PD_CMAX4, "dvm_pair_dist_fwd: bad sizes (C=%d)", C);
    DVM_CHECK_ARG((geo == nullptr) == (geo_out == nullptr), "dvm_pair_dist_fwd: geo and geo_out come together");
    const int grid = ceil_div(B * S, 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (geo && geo_is_f64) pair_dist_fwd_kernel<double><<<grid, 256, 0, st>>>(feat, qidx, nbr, B, N, C, S, k, d, (const double*)geo, geo_out);
    else                   pair_dist_fwd_kernel<float><<<grid, 256, 0, st>>>(feat, qidx, nbr, B, N, C, S, k, d, (const float*)geo, geo_out);
    DVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int dvm_pair_dist_bwd(const float* feat, const int64_t* qidx, const int64_t* nbr, const float* d, const float* g,
                                 int B, int N, int C, int S, int k, float* dfeat, void* stream) {
    DVM_CHECK_ARG(feat && qidx && nbr && d && g && dfeat, "dvm_pair_dist_bwd: null pointer");
    DVM_CHECK_ARG(B > 0 && N > 0 && S > 0 && k > 0 && C > 0 && C % 4 == 0 && C <= 32 * PD_CMAX4, "dvm_pair_dist_bwd: bad sizes (C=%d)", C);
    DVM_CHECK_ARG(((uintptr_t)dfeat & 15) == 0, "dvm_pair_dist_bwd: dfeat must be 16-byte aligned");
    pair_dist_bwd_kernel<<<ceil_div(B * S, 8), 256, 0, (cudaStream_t)stream>>>(feat, qidx, nbr, d, g, B, N, C, S, k, dfeat);
    DVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int dvm_gather_rows_fwd(const float* pts, const int64_t* idx, int B, int N, int R, int C, float* out, void* stream) {
    DVM_CHECK_ARG(pts && idx && out, "dvm_gather_rows_fwd: null pointer");
    DVM_CHECK_ARG(B > 0 && N > 0 && R > 0 && C > 0, "dvm_gather_rows_fwd: bad sizes");
    const long long rows = (long long)B * R;
    const long long elems = (C & 3) == 0 ? rows * (C >> 2) : rows * C;
    DVM_CHECK_ARG(elems / 256 < (1ll << 31), "dvm_gather_rows_fwd: too large");
    gather_rows_fwd_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pts, idx, rows, R, N, C, out);
    DVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int dvm_gather_rows_bwd(const float* d_out, const int64_t* idx, int B, int N, int R, int C, float* d_pts, void* stream) {
    DVM_CHECK_ARG(d_out && idx && d_pts, "dvm_gather_rows_bwd: null pointer");
    DVM_CHECK_ARG(B > 0 && N > 0 && R > 0 && C > 0, "dvm_gather_rows_bwd: bad sizes");
    const long long rows = (long long)B * R, elems = rows * C;
    DVM_CHECK_ARG(elems / 256 < (1ll << 31), "dvm_gather_rows_bwd: too large");
    gather_rows_bwd_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_out, idx, rows, R, N, C, d_pts);
    DVM_LAUNCH_CHECK();
    return 0;
}
