// Deformation-graph forward/backward kernels and the 10-sparse gathers around the Deformer:
// rotation_6d_to_matrix (models/loss.py:39-45), DeformationGraph_geod.forward skinning + ARAP +
// rotation smoothness (lib/deformation_graph_point.py:233-261), index_points + 1x1 conv
// (models/loss.py:1252-1253 -> models/model.py:468-469) and the sparse Pi @ Y transfers
// (models/model.py:471, models/loss.py:1237).  All HBM-bound gather/scatter work: coalesced per-row
// accesses, node tables and neighbour rows served from L2.
#include "common.cuh"

namespace dvm {

// ------------------------------------------------------------------------------------------------
// rotation_6d_to_matrix
// ------------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// one warp = 32 consecutive rotations: the 24-byte inputs and 36-byte outputs are moved as contiguous float runs (coalesced
// 128-byte lines) through a per-warp shared-memory tile; strides 6 and 9 are conflict-free (6: 2-way at most).
__global__ void __launch_bounds__(256) rot6d_fwd_kernel(const float* __restrict__ d6, int n, float* __restrict__ R) {
    __shared__ float s_in[8][32 * 6], s_out[8][32 * 9];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i0 = (blockIdx.x * 8 + w) * 32;
    if (i0 >= n) return;
    const int cnt = min(32, n - i0);
    for (int e = lane; e < cnt * 6; e += 32) s_in[w][e] = __ldg(d6 + (size_t)i0 * 6 + e);
    __syncwarp();
    if (lane < cnt) {
        const float* p = s_in[w] + lane * 6;
        const V3 a1 = v3(p[0], p[1], p[2]), a2 = v3(p[3], p[4], p[5]);
        const V3 b1 = (1.f / fmaxf(sqrtf(dot(a1, a1)), 1e-12f)) * a1;        // F.normalize, eps 1e-12
        const V3 u = a2 - dot(b1, a2) * b1;
        const V3 b2 = (1.f / fmaxf(sqrtf(dot(u, u)), 1e-12f)) * u;
        const V3 b3 = cross(b1, b2);
        float* o = s_out[w] + lane * 9;
        o[0] = b1.x; o[1] = b1.y; o[2] = b1.z; o[3] = b2.x; o[4] = b2.y; o[5] = b2.z; o[6] = b3.x; o[7] = b3.y; o[8] = b3.z;
    }
    __syncwarp();
    for (int e = lane; e < cnt * 9; e += 32) R[(size_t)i0 * 9 + e] = s_out[w][e];
}

__global__ void rot6d_bwd_kernel(const float* __restrict__ d6, const float* __restrict__ dR, int n, float* __restrict__ dd6) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = d6 + (size_t)i * 6;
    const float* g = dR + (size_t)i * 9;
    const V3 a1 = v3(p[0], p[1], p[2]), a2 = v3(p[3], p[4], p[5]);
    const float n1 = fmaxf(sqrtf(dot(a1, a1)), 1e-12f);
    const V3 b1 = (1.f / n1) * a1;
    const float s = dot(b1, a2);
    const V3 u = a2 - s * b1;
    const float n2 = fmaxf(sqrtf(dot(u, u)), 1e-12f);
    const V3 b2 = (1.f / n2) * u;
    V3 g1 = v3(g[0], g[1], g[2]), g2 = v3(g[3], g[4], g[5]);
    const V3 g3 = v3(g[6], g[7], g[8]);
    g1 = g1 + cross(b2, g3);                       // b3 = b1 x b2
    g2 = g2 + cross(g3, b1);
    const V3 du = (1.f / n2) * (g2 - dot(b2, g2) * b2);          // b2 = u / |u|
    V3 da2 = du;
    const float ds = -dot(du, b1);                 // u = a2 - s b1
    g1 = g1 + (-s) * du;
    g1 = g1 + ds * a2;                             // s = b1 . a2
    da2 = da2 + ds * b1;
    const V3 da1 = (1.f / n1) * (g1 - dot(b1, g1) * b1);         // b1 = a1 / |a1|
    float* o = dd6 + (size_t)i * 6;
    o[0] = da1.x; o[1] = da1.y; o[2] = da1.z; o[3] = da2.x; o[4] = da2.y; o[5] = da2.z;
}

// ------------------------------------------------------------------------------------------------
// skinning
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ V3 ld3(const float* p) { return v3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }
__device__ __forceinline__ V3 matvec(const float* R, V3 x) {
    return v3(__ldg(R) * x.x + __ldg(R + 1) * x.y + __ldg(R + 2) * x.z,
              __ldg(R + 3) * x.x + __ldg(R + 4) * x.y + __ldg(R + 5) * x.z,
              __ldg(R + 6) * x.x + __ldg(R + 7) * x.y + __ldg(R + 8) * x.z);
}

__global__ void skin_fwd_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ nodes_idx,
                                const int64_t* __restrict__ infl, const float* __restrict__ wts,
                                const float* __restrict__ R, const float* __restrict__ t,
                                int N, int K, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float* P = xyz + (size_t)b * N * 3;
    const V3 v = ld3(P + (size_t)i * 3);
    V3 acc = v3(0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const size_t e = ((size_t)b * N + i) * 3 + k;
        const int64_t n = infl[e];
        const float w = wts[e];
        const V3 g = ld3(P + (size_t)nodes_idx[(size_t)b * K + n] * 3);
        const V3 tn = ld3(t + ((size_t)b * K + n) * 3);
        const V3 y = matvec(R + ((size_t)b * K + n) * 9, v - g) + g + tn;
        acc = acc + w * y;
    }
    float* o = out + ((size_t)b * N + i) * 3;
    o[0] = acc.x; o[1] = acc.y; o[2] = acc.z;
}

__global__ void skin_bwd_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ nodes_idx,
                                const int64_t* __restrict__ infl, const float* __restrict__ wts,
                                const float* __restrict__ dOut, int N, int K,
                                float* __restrict__ dR, float* __restrict__ dt) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float* P = xyz + (size_t)b * N * 3;
    const V3 v = ld3(P + (size_t)i * 3);
    const V3 go = ld3(dOut + ((size_t)b * N + i) * 3);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const size_t e = ((size_t)b * N + i) * 3 + k;
        const int64_t n = infl[e];
        const float w = wts[e];
        const V3 x = v - ld3(P + (size_t)nodes_idx[(size_t)b * K + n] * 3);
        const V3 gw = w * go;
        float* r = dR + ((size_t)b * K + n) * 9;
        float* tt = dt + ((size_t)b * K + n) * 3;
        atomicAdd(tt, gw.x); atomicAdd(tt + 1, gw.y); atomicAdd(tt + 2, gw.z);
        atomicAdd(r + 0, gw.x * x.x); atomicAdd(r + 1, gw.x * x.y); atomicAdd(r + 2, gw.x * x.z);
        atomicAdd(r + 3, gw.y * x.x); atomicAdd(r + 4, gw.y * x.y); atomicAdd(r + 5, gw.y * x.z);
        atomicAdd(r + 6, gw.z * x.x); atomicAdd(r + 7, gw.z * x.y); atomicAdd(r + 8, gw.z * x.z);
    }
}

// ------------------------------------------------------------------------------------------------
// ARAP + rotation smoothness: per-node partial sums -> per-block partials -> fixed-order final sum
// ------------------------------------------------------------------------------------------------
constexpr int ARAP_THREADS = 256;

__global__ void __launch_bounds__(ARAP_THREADS)
arap_fwd_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ nodes_idx, const int64_t* __restrict__ ring,
                const float* __restrict__ R, const float* __restrict__ t, int N, int K, int ring_k,
                float* __restrict__ part /* [B][gridDim.x][2] */) {
    __shared__ float s[2][ARAP_THREADS / 32];
    const int b = blockIdx.y;
    const int i = blockIdx.x * ARAP_THREADS + threadIdx.x;
    float a_sum = 0.f, r_sum = 0.f;
    if (i < K) {
        const float* P = xyz + (size_t)b * N * 3;
        const V3 gi = ld3(P + (size_t)nodes_idx[(size_t)b * K + i] * 3);
        const V3 ti = ld3(t + ((size_t)b * K + i) * 3);
        const float* Ri = R + ((size_t)b * K + i) * 9;
        float ri[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) ri[c] = __ldg(Ri + c);
        for (int q = 0; q < ring_k; ++q) {
            const int64_t j = ring[((size_t)b * K + i) * ring_k + q];
            const V3 gj = ld3(P + (size_t)nodes_idx[(size_t)b * K + j] * 3);
            const V3 tj = ld3(t + ((size_t)b * K + j) * 3);
            const V3 d = ((gi + ti) - (gj + tj)) - matvec(Ri, gi - gj);
            a_sum += dot(d, d);
            const float* Rj = R + ((size_t)b * K + j) * 9;
#pragma unroll
            for (int c = 0; c < 9; ++c) { const float e = ri[c] - __ldg(Rj + c); r_sum += e * e; }
        }
    }
    a_sum = warp_sum(a_sum); r_sum = warp_sum(r_sum);
    if ((threadIdx.x & 31) == 0) { s[0][threadIdx.x >> 5] = a_sum; s[1][threadIdx.x >> 5] = r_sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, r = 0.f;
        for (int w = 0; w < ARAP_THREADS / 32; ++w) { a += s[0][w]; r += s[1][w]; }
        float* o = part + ((size_t)b * gridDim.x + blockIdx.x) * 2;
        o[0] = a; o[1] = r;
    }
}

__global__ void arap_final_kernel(const float* __restrict__ part, int nblk, int K, int ring_k,
                                  float* __restrict__ arap, float* __restrict__ sr) {
    const int b = blockIdx.x;
    if (threadIdx.x != 0) return;
    float a = 0.f, r = 0.f;
    for (int k = 0; k < nblk; ++k) { a += part[((size_t)b * nblk + k) * 2]; r += part[((size_t)b * nblk + k) * 2 + 1]; }
    arap[b] = a / (float)K;
    if (sr) sr[b] = r / ((float)K * (float)ring_k * 9.f);
}

__global__ void arap_bwd_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ nodes_idx, const int64_t* __restrict__ ring,
                                const float* __restrict__ R, const float* __restrict__ t, const float* __restrict__ g_arap,
                                int N, int K, int ring_k, float* __restrict__ dR, float* __restrict__ dt) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const float sc = 2.f * g_arap[b] / (float)K;
    const float* P = xyz + (size_t)b * N * 3;
    const V3 gi = ld3(P + (size_t)nodes_idx[(size_t)b * K + i] * 3);
    const V3 ti = ld3(t + ((size_t)b * K + i) * 3);
    const float* Ri = R + ((size_t)b * K + i) * 9;
    V3 dti = v3(0.f, 0.f, 0.f);
    float dri[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int q = 0; q < ring_k; ++q) {
        const int64_t j = ring[((size_t)b * K + i) * ring_k + q];
        if (j == i) continue;                                  // self entry contributes exactly 0
        const V3 gj = ld3(P + (size_t)nodes_idx[(size_t)b * K + j] * 3);
        const V3 tj = ld3(t + ((size_t)b * K + j) * 3);
        const V3 e = gi - gj;
        const V3 d = sc * (((gi + ti) - (gj + tj)) - matvec(Ri, e));
        dti = dti + d;
        float* tjp = dt + ((size_t)b * K + j) * 3;
        atomicAdd(tjp, -d.x); atomicAdd(tjp + 1, -d.y); atomicAdd(tjp + 2, -d.z);
        dri[0] -= d.x * e.x; dri[1] -= d.x * e.y; dri[2] -= d.x * e.z;
        dri[3] -= d.y * e.x; dri[4] -= d.y * e.y; dri[5] -= d.y * e.z;
        dri[6] -= d.z * e.x; dri[7] -= d.z * e.y; dri[8] -= d.z * e.z;
    }
    float* tip = dt + ((size_t)b * K + i) * 3;
    atomicAdd(tip, dti.x); atomicAdd(tip + 1, dti.y); atomicAdd(tip + 2, dti.z);
    float* rip = dR + ((size_t)b * K + i) * 9;
#pragma unroll
    for (int c = 0; c < 9; ++c) atomicAdd(rip + c, dri[c]);
}

// ------------------------------------------------------------------------------------------------
// index_points + Conv2d(k -> 1, 1x1): one warp per output row, float4 lanes over the channels
// ------------------------------------------------------------------------------------------------
constexpr int GC_KMAX = 16;

__global__ void __launch_bounds__(256)
gather_conv_fwd_kernel(const float* __restrict__ feat, const int64_t* __restrict__ idx, const float* __restrict__ W,
                       const float* __restrict__ bias, int rows, int R, int N, int C, int k, float* __restrict__ out) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int b = row / R;
    const float* Fb = feat + (size_t)b * N * C;
    const float bs = bias ? __ldg(bias) : 0.f;
    for (int c = lane * 4; c < C; c += 128) {
        float4 acc = make_float4(bs, bs, bs, bs);
        for (int s = 0; s < k; ++s) {
            const int64_t j = __ldg(idx + (size_t)row * k + s);
            const float w = __ldg(W + s);
            const float4 f = __ldg(reinterpret_cast<const float4*>(Fb + (size_t)j * C + c));
            acc.x = fmaf(w, f.x, acc.x); acc.y = fmaf(w, f.y, acc.y); acc.z = fmaf(w, f.z, acc.z); acc.w = fmaf(w, f.w, acc.w);
        }
        *reinterpret_cast<float4*>(out + (size_t)row * C + c) = acc;
    }
}

__global__ void __launch_bounds__(256)
gather_conv_bwd_kernel(const float* __restrict__ feat, const int64_t* __restrict__ idx, const float* __restrict__ W,
                       const float* __restrict__ dOut, int rows, int R, int N, int C, int k,
                       float* __restrict__ dFeat, float* __restrict__ dW, float* __restrict__ dBias) {
    __shared__ float s_w[GC_KMAX + 1];
    if (threadIdx.x <= GC_KMAX) s_w[threadIdx.x] = 0.f;
    __syncthreads();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row < rows) {
        const int b = row / R;
        const float* Fb = feat + (size_t)b * N * C;
        float* dFb = dFeat + (size_t)b * N * C;
        float wacc[GC_KMAX];
#pragma unroll
        for (int s = 0; s < GC_KMAX; ++s) wacc[s] = 0.f;
        float bacc = 0.f;
        for (int c = lane * 4; c < C; c += 128) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(dOut + (size_t)row * C + c));
            bacc += (g.x + g.y) + (g.z + g.w);
#pragma unroll
            for (int s = 0; s < GC_KMAX; ++s) {
                if (s < k) {
                    const int64_t j = __ldg(idx + (size_t)row * k + s);
                    const float w = __ldg(W + s);
                    const float4 f = __ldg(reinterpret_cast<const float4*>(Fb + (size_t)j * C + c));
                    wacc[s] += g.x * f.x + g.y * f.y + g.z * f.z + g.w * f.w;
                    float* d = dFb + (size_t)j * C + c;
                    atomicAdd(d, w * g.x); atomicAdd(d + 1, w * g.y); atomicAdd(d + 2, w * g.z); atomicAdd(d + 3, w * g.w);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < GC_KMAX; ++s) {
            if (s < k) { const float v = warp_sum(wacc[s]); if (lane == 0) atomicAdd(&s_w[s], v); }
        }
        bacc = warp_sum(bacc);
        if (lane == 0) atomicAdd(&s_w[GC_KMAX], bacc);
    }
    __syncthreads();
    if (threadIdx.x < k) atomicAdd(dW + threadIdx.x, s_w[threadIdx.x]);
    if (threadIdx.x == GC_KMAX && dBias) atomicAdd(dBias, s_w[GC_KMAX]);
}

// ------------------------------------------------------------------------------------------------
// sparse transfers  out[row,:] = sum_k w[row,k] Y[b, idx[row,k], :]
// ------------------------------------------------------------------------------------------------
// wide rows (D % 4 == 0, e.g. the 128 conv-reduced feature channels of models/model.py:471): one warp per output row, float4 lanes
__global__ void __launch_bounds__(256)
sparse_transfer_fwd_kernel(const int* __restrict__ idx, const float* __restrict__ w, const float* __restrict__ Y,
                           int rows, int N, int M, int K, int D, float* __restrict__ out) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int b = row / N;
    const float* Yb = Y + (size_t)b * M * D;
    float wk[DVM_KNN_MAX]; int jk[DVM_KNN_MAX];
#pragma unroll
    for (int k = 0; k < DVM_KNN_MAX; ++k) {
        wk[k] = k < K ? __ldg(w + (size_t)row * K + k) : 0.f;
        jk[k] = k < K ? __ldg(idx + (size_t)row * K + k) : 0;
    }
    if ((D & 3) == 0) {
        for (int d = lane * 4; d < D; d += 128) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < DVM_KNN_MAX; ++k)
                if (k < K && wk[k] != 0.f) {
                    const float4 y = __ldg(reinterpret_cast<const float4*>(Yb + (size_t)jk[k] * D + d));
                    acc.x = fmaf(wk[k], y.x, acc.x); acc.y = fmaf(wk[k], y.y, acc.y); acc.z = fmaf(wk[k], y.z, acc.z); acc.w = fmaf(wk[k], y.w, acc.w);
                }
            *reinterpret_cast<float4*>(out + (size_t)row * D + d) = acc;
        }
        return;
    }
    for (int d = lane; d < D; d += 32) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < DVM_KNN_MAX; ++k)
            if (k < K && wk[k] != 0.f) acc = fmaf(wk[k], __ldg(Yb + (size_t)jk[k] * D + d), acc);
        out[(size_t)row * D + d] = acc;
    }
}

// narrow rows (D <= 4: Pi @ verts, models/loss.py:1408-1409): one THREAD per output row (a warp per row leaves 29 lanes idle)
template <int kD>
__global__ void __launch_bounds__(256)
sparse_transfer_fwd_narrow_kernel(const int* __restrict__ idx, const float* __restrict__ w, const float* __restrict__ Y,
                                  int rows, int N, int M, int K, float* __restrict__ out) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const int b = row / N;
    const float* Yb = Y + (size_t)b * M * kD;
    float acc[kD];
#pragma unroll
    for (int d = 0; d < kD; ++d) acc[d] = 0.f;
    for (int k = 0; k < K; ++k) {
        const float wk = __ldg(w + (size_t)row * K + k);
        const int j = __ldg(idx + (size_t)row * K + k);
        if (wk != 0.f) {
#pragma unroll
            for (int d = 0; d < kD; ++d) acc[d] = fmaf(wk, __ldg(Yb + (size_t)j * kD + d), acc[d]);
        }
    }
#pragma unroll
    for (int d = 0; d < kD; ++d) out[(size_t)row * kD + d] = acc[d];
}

__global__ void __launch_bounds__(256)
sparse_transfer_bwd_kernel(const int* __restrict__ idx, const float* __restrict__ w, const float* __restrict__ Y,
                           const float* __restrict__ dOut, int rows, int N, int M, int K, int D,
                           float* __restrict__ dW, float* __restrict__ dY) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int b = row / N;
    const float* Yb = Y + (size_t)b * M * D;
    float* dYb = dY ? dY + (size_t)b * M * D : nullptr;
    for (int k = 0; k < K; ++k) {
        const int j = __ldg(idx + (size_t)row * K + k);
        const float wk = __ldg(w + (size_t)row * K + k);
        float acc = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float g = __ldg(dOut + (size_t)row * D + d);
            acc = fmaf(g, __ldg(Yb + (size_t)j * D + d), acc);
            if (dYb && wk != 0.f) atomicAdd(dYb + (size_t)j * D + d, wk * g);
        }
        acc = warp_sum(acc);
        if (lane == 0 && dW) dW[(size_t)row * K + k] = acc;
    }
}

}  // namespace dvm

using namespace dvm;

extern "C" int dvm_rot6d_fwd(const float* d6, int n, float* R, void* stream) {
    DVM_CHECK_ARG(d6 && R && n > 0, "dvm_rot6d_fwd: bad arguments");
    rot6d_fwd_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(d6, n, R);       // 8 warps x 32 rotations per block
    DVM_LAUNCH_CHECK();
    return 0;
}
extern "C" int dvm_rot6d_bwd(const float* d6, const float* dR, int n, float* dd6, void* stream) {
    DVM_CHECK_ARG(d6 && dR && dd6 && n > 0, "dvm_rot6d_bwd: bad arguments");
    rot6d_bwd_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(d6, dR, n, dd6);
    DVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int dvm_skin_fwd(const float* xyz, const int64_t* nodes_idx, const int64_t* influence, const float* weights,
                            const float* R, const float* t, int B, int N, int K, float* out, void* stream) {
    DVM_CHECK_ARG(xyz && nodes_idx && influence && weights && R && t && out, "dvm_skin_fwd: null pointer");
    DVM_CHECK_ARG(B > 0 && N > 0 && K > 0 && B <= 65535, "dvm_skin_fwd: bad sizes");
    skin_fwd_kernel<<<dim3(ceil_div(N, 256), B), 256, 0, (cudaStream_t)stream>>>(xyz, nodes_idx, influence, weights, R, t, N, K, out);
    DVM_LAUNCH_CHECK();
    return 0;
}
extern "C" int dvm_skin_bwd(const float* xyz, const int64_t* nodes_idx, const int64_t* influence, const float* weights,
                            const float* dOut, int B, int N, int K, float* dR, float* dt, void* stream) {
    DVM_CHECK_ARG(xyz && nodes_idx && influence && weights && dOut && dR && dt, "dvm_skin_bwd: null pointer");
    DVM_CHECK_ARG(B > 0 && N > 0 && K > 0 && B <= 65535, "dvm_skin_bwd: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    DVM_CUDA(cudaMemsetAsync(dR, 0, (size_t)B * K * 9 * sizeof(float), st));
    DVM_CUDA(cudaMemsetAsync(dt, 0, (size_t)B * K * 3 * sizeof(float), st));
    skin_bwd_kernel<<<dim3(ceil_div(N, 256), B), 256, 0, st>>>(xyz, nodes_idx, influence, weights, dOut, N, K, dR, dt);
    DVM_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t dvm_arap_workspace_bytes(int B, int K) {
    if (B <= 0 || K <= 0) return 0;
    return align_up((size_t)B * ceil_div(K, ARAP_THREADS) * 2 * sizeof(float), 256);
}
extern "C" int dvm_arap_fwd(const float* xyz, const int64_t* nodes_idx, const int64_t* ring, const float* R, const float* t,
                            int B, int N, int K, int ring_k, float* arap, float* sr, void* ws, size_t ws_bytes, void* stream) {
    DVM_CHECK_ARG(xyz && nodes_idx && ring && R && t && arap, "dvm_arap_fwd: null pointer");
    DVM_CHECK_ARG(B > 0 && N > 0 && K > 0 && ring_k > 0 && B <= 65535, "dvm_arap_fwd: bad sizes");
    if (!ws || ws_bytes < dvm_arap_workspace_bytes(B, K)) { set_error("dvm_arap_fwd: workspace too small"); return DVM_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = ceil_div(K, ARAP_THREADS);
    arap_fwd_kernel<<<dim3(nblk, B), ARAP_THREADS, 0, st>>>(xyz, nodes_idx, ring, R, t, N, K, ring_k, (float*)ws);
    DVM_LAUNCH_CHECK();
    arap_final_kernel<<<B, 32, 0, st>>>((const float*)ws, nblk, K, ring_k, arap, sr);
    DVM_LAUNCH_CHECK();
    return 0;
}
extern "C" int dvm_arap_bwd(const float* xyz, const int64_t* nodes_idx, const int64_t* ring, const float* R, const float* t,
                            const float* g_arap, int B, int N, int K, int ring_k, float* dR, float* dt, void* stream) {
    DVM_CHECK_ARG(xyz && nodes_idx && ring && R && t && g_arap && dR && dt, "dvm_arap_bwd: null pointer");
    DVM_CHECK_ARG(B > 0 && N > 0 && K > 0 && ring_k > 0 && B <= 65535, "dvm_arap_bwd: bad sizes");
    arap_bwd_kernel<<<dim3(ceil_div(K, 256), B), 256, 0, (cudaStream_t)stream>>>(xyz, nodes_idx, ring, R, t, g_arap, N, K, ring_k, dR, dt);
    DVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int dvm_gather_conv_fwd(const float* feat, const int64_t* idx, const float* W, const float* bias,
                                   int B, int N, int R, int C, int k, float* out, void* stream) {
    DVM_CHECK_ARG(feat && idx && W && out, "dvm_gather_conv_fwd: null pointer");
    DVM_CHECK_ARG(B > 0 && N > 0 && R > 0 && C > 0 && C % 4 == 0 && k > 0 && k <= GC_KMAX, "dvm_gather_conv_fwd: bad sizes (C=%d k=%d)", C, k);
    const int rows = B * R;
    gather_conv_fwd_kernel<<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(feat, idx, W, bias, rows, R, N, C, k, out);
    DVM_LAUNCH_CHECK();
    return 0;
}
extern "C" int dvm_gather_conv_bwd(const float* feat, const int64_t* idx, const float* W, const float* dOut,
                                   int B, int N, int R, int C, int k, float* dFeat, float* dW, float* dBias, void* stream) {
    DVM_CHECK_ARG(feat && idx && W && dOut && dFeat && dW, "dvm_gather_conv_bwd: null pointer");
    DVM_CHECK_ARG(B > 0 && N > 0 && R > 0 && C > 0 && C % 4 == 0 && k > 0 && k <= GC_KMAX, "dvm_gather_conv_bwd: bad sizes");
    const int rows = B * R;
    gather_conv_bwd_kernel<<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(feat, idx, W, dOut, rows, R, N, C, k, dFeat, dW, dBias);
    DVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int dvm_sparse_transfer_fwd(const int32_t* idx, const float* w, const float* Y,
                                       int B, int N, int M, int K, int D, float* out, void* stream) {
    DVM_CHECK_ARG(idx && w && Y && out, "dvm_sparse_transfer_fwd: null pointer");
    // the wide kernel keeps the K (index, weight) pairs of a row in registers (K <= DVM_KNN_MAX); the narrow one (D <= 3) loops over K
    DVM_CHECK_ARG(B > 0 && N > 0 && M > 0 && D > 0 && K > 0 && (K <= DVM_KNN_MAX || (D <= 3 && K <= 1024)), "dvm_sparse_transfer_fwd: bad sizes");
    const int rows = B * N;
    cudaStream_t st = (cudaStream_t)stream;
    if (D == 3)      sparse_transfer_fwd_narrow_kernel<3><<<ceil_div(rows, 256), 256, 0, st>>>(idx, w, Y, rows, N, M, K, out);
    else if (D == 1) sparse_transfer_fwd_narrow_kernel<1><<<ceil_div(rows, 256), 256, 0, st>>>(idx, w, Y, rows, N, M, K, out);
    else if (D == 2) sparse_transfer_fwd_narrow_kernel<2><<<ceil_div(rows, 256), 256, 0, st>>>(idx, w, Y, rows, N, M, K, out);
    else             sparse_transfer_fwd_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(idx, w, Y, rows, N, M, K, D, out);
    DVM_LAUNCH_CHECK();
    return 0;
}
extern "C" int dvm_sparse_transfer_bwd(const int32_t* idx, const float* w, const float* Y, const float* dOut,
                                       int B, int N, int M, int K, int D, float* dW, float* dY, void* stream) {
    DVM_CHECK_ARG(idx && w && Y && dOut && (dW || dY), "dvm_sparse_transfer_bwd: null pointer");
    DVM_CHECK_ARG(B > 0 && N > 0 && M > 0 && D > 0 && K > 0 && K <= DVM_KNN_MAX, "dvm_sparse_transfer_bwd: bad sizes");
    const int rows = B * N;
    sparse_transfer_bwd_kernel<<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(idx, w, Y, dOut, rows, N, M, K, D, dW, dY);
    DVM_LAUNCH_CHECK();
    return 0;
}
