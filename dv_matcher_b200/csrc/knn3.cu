// Brute-force k-NN / Chamfer on 3-D points: shared-memory tiles of the reference cloud, a sub-warp of
// LPQ lanes per query, sorted register lists, shuffle merges.  Exact direct-difference arithmetic:
// fp32 path is (dx*dx + dy*dy) + dz*dz with NO FMA contraction (bit-identical to the oracle/reference
// restatement); fp64 path evaluates like SciPy's KD-tree (lib/deformation_graph_point.py:181-191).
#include "common.cuh"

namespace dvm {

constexpr int KNN_THREADS = 256;
constexpr int KNN_TILE = 1024;       // reference points staged per tile (16 KB as float4)

template <typename T> struct Dist3;
template <> struct Dist3<float> {
    static __device__ __forceinline__ float eval(float qx, float qy, float qz, float4 p) {
        const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
        return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    }
};
template <> struct Dist3<double> {
    static __device__ __forceinline__ double eval(float qx, float qy, float qz, float4 p) {
        const double dx = (double)qx - (double)p.x, dy = (double)qy - (double)p.y, dz = (double)qz - (double)p.z;
        return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    }
};

template <typename T, int K>
struct KList {
    T key[K]; int idx[K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int t = 0; t < K; ++t) { key[t] = (T)INFINITY; idx[t] = 0x7fffffff; }
    }
    __device__ __forceinline__ void push(T k, int j) {     // lexicographic (key, idx) insertion
        T ck = k; int ci = j;
#pragma unroll
        for (int t = 0; t < K; ++t) {
            const bool sw = ck < key[t] || (ck == key[t] && ci < idx[t]);
            const T tk = key[t]; const int ti = idx[t];
            key[t] = sw ? ck : tk; idx[t] = sw ? ci : ti;
            ck = sw ? tk : ck; ci = sw ? ti : ci;
        }
    }
};

// One query per LPQ consecutive lanes; lane s of the group scans tile entries s, s+LPQ, ...
template <typename T, int K, int LPQ>
__global__ void __launch_bounds__(KNN_THREADS)
knn3_kernel(const float* __restrict__ Q, const float* __restrict__ R, int N, int M, int k,
            int64_t* __restrict__ idx64, int32_t* __restrict__ idx32, float* __restrict__ d2f, double* __restrict__ d2d) {
    __shared__ float4 tile[KNN_TILE];
    const int b = blockIdx.y;
    const int qpb = KNN_THREADS / LPQ;
    const int q = blockIdx.x * qpb + threadIdx.x / LPQ;
    const int sub = threadIdx.x % LPQ;
    const bool live = q < N;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (live) {
        const float* qp = Q + ((size_t)b * N + q) * 3;
        qx = __ldg(qp); qy = __ldg(qp + 1); qz = __ldg(qp + 2);
    }
    const float* Rb = R + (size_t)b * M * 3;
    KList<T, K> list;
    list.init();

    for (int j0 = 0; j0 < M; j0 += KNN_TILE) {
        const int cnt = min(KNN_TILE, M - j0);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt; e += KNN_THREADS) {
            const float* p = Rb + (size_t)(j0 + e) * 3;
            tile[e] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
        }
        __syncthreads();
        if (live) {
#pragma unroll 4
            for (int e = sub; e < cnt; e += LPQ) {
                const T d = Dist3<T>::eval(qx, qy, qz, tile[e]);
                if (d < list.key[K - 1] || (d == list.key[K - 1] && j0 + e < list.idx[K - 1])) list.push(d, j0 + e);
            }
        }
    }

    // merge the LPQ sorted lists of a query: k rounds of group-wide lexicographic min
    const unsigned full = 0xffffffffu;
    int head = 0;
    for (int s = 0; s < k; ++s) {
        T hk = (T)INFINITY; int hi = 0x7fffffff;
#pragma unroll
        for (int t = 0; t < K; ++t) if (t == head) { hk = list.key[t]; hi = list.idx[t]; }
        T wk = hk; int wi = hi;
#pragma unroll
        for (int o = LPQ / 2; o > 0; o >>= 1) {
            const T ok = __shfl_xor_sync(full, wk, o);
            const int oi = __shfl_xor_sync(full, wi, o);
            if (ok < wk || (ok == wk && oi < wi)) { wk = ok; wi = oi; }
        }
        if (hi == wi && hi != 0x7fffffff) ++head;
        if (live && sub == 0) {
            const size_t o = ((size_t)b * N + q) * k + s;
            if (idx64) idx64[o] = wi;
            if (idx32) idx32[o] = wi;
            if (d2f) d2f[o] = (float)wk;
            if (d2d) d2d[o] = (double)wk;
        }
    }
}

template <typename T, int K>
static int launch_knn3_k(const float* Q, const float* R, int B, int N, int M, int k,
                         int64_t* idx64, int32_t* idx32, float* d2f, double* d2d, cudaStream_t st) {
    // few queries: 8 lanes per query so the grid still covers the 148 SMs; many queries: 1 lane each
    const long long total = (long long)B * N;
    if (total >= 148LL * 2048) {
        dim3 grid(ceil_div(N, KNN_THREADS), B);
        knn3_kernel<T, K, 1><<<grid, KNN_THREADS, 0, st>>>(Q, R, N, M, k, idx64, idx32, d2f, d2d);
    } else {
        dim3 grid(ceil_div(N, KNN_THREADS / 8), B);
        knn3_kernel<T, K, 8><<<grid, KNN_THREADS, 0, st>>>(Q, R, N, M, k, idx64, idx32, d2f, d2d);
    }
    DVM_LAUNCH_CHECK();
    return 0;
}

size_t knn3_grid_workspace_bytes(int B, int N, int M);
int launch_knn3_grid(const float* Q, const float* R, int B, int N, int M, int k, bool f64,
                     int64_t* idx64, int32_t* idx32, float* d2f, double* d2d, void* ws, size_t ws_bytes, cudaStream_t st);

constexpr int KNN_GRID_MIN_M = 1024;      // below this the brute-force sweep is as fast and one launch

// grid search when the caller provided scratch and the reference cloud is big enough, brute force otherwise;
// both give bit-identical results
int launch_knn3_auto(const float* Q, const float* R, int B, int N, int M, int k, bool f64,
                     int64_t* idx64, int32_t* idx32, float* d2f, double* d2d, void* ws, size_t ws_bytes, cudaStream_t st);

int launch_knn3(const float* Q, const float* R, int B, int N, int M, int k, bool f64,
                int64_t* idx64, int32_t* idx32, float* d2f, double* d2d, cudaStream_t st) {
#define DVM_KNN_DISPATCH(T)                                                                       \
    if (k == 1)       return launch_knn3_k<T, 1>(Q, R, B, N, M, k, idx64, idx32, d2f, d2d, st);   \
    else if (k <= 4)  return launch_knn3_k<T, 4>(Q, R, B, N, M, k, idx64, idx32, d2f, d2d, st);   \
    else if (k <= 10) return launch_knn3_k<T, 10>(Q, R, B, N, M, k, idx64, idx32, d2f, d2d, st);  \
    else              return launch_knn3_k<T, 16>(Q, R, B, N, M, k, idx64, idx32, d2f, d2d, st);
    if (f64) { DVM_KNN_DISPATCH(double) } else { DVM_KNN_DISPATCH(float) }
#undef DVM_KNN_DISPATCH
}

int launch_knn3_auto(const float* Q, const float* R, int B, int N, int M, int k, bool f64,
                     int64_t* idx64, int32_t* idx32, float* d2f, double* d2d, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (ws && M >= KNN_GRID_MIN_M && ws_bytes >= knn3_grid_workspace_bytes(B, N, M))
        return launch_knn3_grid(Q, R, B, N, M, k, f64, idx64, idx32, d2f, d2d, ws, ws_bytes, st);
    return launch_knn3(Q, R, B, N, M, k, f64, idx64, idx32, d2f, d2d, st);
}

// ------------------------------------------------------------------------------------------------
// Chamfer backward: da_i = 2 g1_i (a_i - b_idx1(i)) - sum_{j: idx2(j)=i} 2 g2_j (b_j - a_i); same for db
// ------------------------------------------------------------------------------------------------
__global__ void chamfer_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                   const int* __restrict__ idx1, const int* __restrict__ idx2,
                                   const float* __restrict__ g1, const float* __restrict__ g2,
                                   int N, int M, float* __restrict__ da, float* __restrict__ db) {
    const int bi = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const float* ab = a + (size_t)bi * N * 3; const float* bb = b + (size_t)bi * M * 3;
    float* dab = da + (size_t)bi * N * 3;     float* dbb = db + (size_t)bi * M * 3;
    if (t < N) {
        const int j = idx1[(size_t)bi * N + t];
        const float g = 2.f * g1[(size_t)bi * N + t];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = g * (ab[t * 3 + c] - bb[j * 3 + c]);
            atomicAdd(dab + t * 3 + c, v);
            atomicAdd(dbb + j * 3 + c, -v);
        }
    }
    if (t < M) {
        const int i = idx2[(size_t)bi * M + t];
        const float g = 2.f * g2[(size_t)bi * M + t];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = g * (bb[t * 3 + c] - ab[i * 3 + c]);
            atomicAdd(dbb + t * 3 + c, v);
            atomicAdd(dab + i * 3 + c, -v);
        }
    }
}

}  // namespace dvm

using namespace dvm;

extern "C" size_t dvm_knn3_workspace_bytes(int B, int N, int M) {
    if (B <= 0 || N <= 0 || M <= 0) return 0;
    return M >= KNN_GRID_MIN_M ? knn3_grid_workspace_bytes(B, N, M) : 0;
}

extern "C" size_t dvm_chamfer_workspace_bytes(int B, int N, int M) {
    if (B <= 0 || N <= 0 || M <= 0) return 0;
    const size_t a = dvm_knn3_workspace_bytes(B, N, M), b = dvm_knn3_workspace_bytes(B, M, N);
    return a > b ? a : b;
}

extern "C" int dvm_knn3(const float* Q, const float* R, int B, int N, int M, int k, int use_f64,
                        int64_t* idx, int32_t* idx32, float* d2, double* d2_f64, void* ws, size_t ws_bytes, void* stream) {
    DVM_CHECK_ARG(Q && R, "dvm_knn3: null input");
    DVM_CHECK_ARG(B > 0 && N > 0 && M > 0, "dvm_knn3: empty problem (B=%d N=%d M=%d)", B, N, M);
    DVM_CHECK_ARG(k >= 1 && k <= DVM_KNN_MAX && k <= M, "dvm_knn3: k=%d must be in [1, min(16, M=%d)]", k, M);
    DVM_CHECK_ARG(B <= 65535, "dvm_knn3: B=%d too large", B);
    return launch_knn3_auto(Q, R, B, N, M, k, use_f64 != 0, idx, idx32, d2, d2_f64, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int dvm_chamfer_fwd(const float* a, const float* b, int B, int N, int M,
                               float* dist1, float* dist2, int32_t* idx1, int32_t* idx2, void* ws, size_t ws_bytes, void* stream) {
    DVM_CHECK_ARG(a && b && dist1 && dist2 && idx1 && idx2, "dvm_chamfer_fwd: null pointer");
    DVM_CHECK_ARG(B > 0 && N > 0 && M > 0 && B <= 65535, "dvm_chamfer_fwd: bad sizes (B=%d N=%d M=%d)", B, N, M);
    int rc = launch_knn3_auto(a, b, B, N, M, 1, false, nullptr, idx1, dist1, nullptr, ws, ws_bytes, (cudaStream_t)stream);
    if (rc) return rc;
    return launch_knn3_auto(b, a, B, M, N, 1, false, nullptr, idx2, dist2, nullptr, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int dvm_chamfer_bwd(const float* a, const float* b, const int32_t* idx1, const int32_t* idx2,
                               const float* g1, const float* g2, int B, int N, int M,
                               float* da, float* db, void* stream) {
    DVM_CHECK_ARG(a && b && idx1 && idx2 && g1 && g2 && da && db, "dvm_chamfer_bwd: null pointer");
    DVM_CHECK_ARG(B > 0 && N > 0 && M > 0 && B <= 65535, "dvm_chamfer_bwd: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    DVM_CUDA(cudaMemsetAsync(da, 0, (size_t)B * N * 3 * sizeof(float), st));
    DVM_CUDA(cudaMemsetAsync(db, 0, (size_t)B * M * 3 * sizeof(float), st));
    dim3 grid(ceil_div(max(N, M), 256), B);
    chamfer_bwd_kernel<<<grid, 256, 0, st>>>(a, b, idx1, idx2, g1, g2, N, M, da, db);
    DVM_LAUNCH_CHECK();
    return 0;
}
