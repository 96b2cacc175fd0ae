// Deformation-graph construction on the GPU: farthest point sampling and the graph tensors of
// construct_graph_euclidean (lib/deformation_graph_point.py:18-33, 177-201).
#include "common.cuh"

namespace dvm {

int launch_knn3_auto(const float* Q, const float* R, int B, int N, int M, int k, bool f64,
                     int64_t* idx64, int32_t* idx32, float* d2f, double* d2d, void* ws, size_t ws_bytes, cudaStream_t st);
size_t knn3_grid_workspace_bytes(int B, int N, int M);

// ------------------------------------------------------------------------------------------------
// FPS: one CTA per cloud, K dependent iterations.  Each iteration: masked min-update of the running
// distance with the unfused squared distance to the newest centroid, then a block-wide arg-max with
// first-index tie-break (== torch.max(distance, -1)[1]).  Points are strided over the threads; up to
// FPS_PPT points per thread live in registers, larger clouds keep the running distance in `ws`.
// ------------------------------------------------------------------------------------------------
constexpr int FPS_THREADS = 1024;
constexpr int FPS_PPT = 8;

__device__ __forceinline__ float fps_d2(float x, float y, float z, float cx, float cy, float cz) {
    const float dx = __fsub_rn(x, cx), dy = __fsub_rn(y, cy), dz = __fsub_rn(z, cz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// (value, index) arg-max with lower index on ties
__device__ __forceinline__ void argmax_combine(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

template <bool kRegs>
__global__ void __launch_bounds__(FPS_THREADS)
fps_kernel(const float* __restrict__ xyz, int N, int K, const int64_t* __restrict__ start,
           int64_t* __restrict__ out, float* __restrict__ dist_ws) {
    __shared__ float s_v[32];
    __shared__ int s_i[32];
    __shared__ int s_far;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float* P = xyz + (size_t)b * N * 3;
    float* dist_g = kRegs ? nullptr : dist_ws + (size_t)b * N;

    float px[FPS_PPT], py[FPS_PPT], pz[FPS_PPT], pd[FPS_PPT];
    if (kRegs) {
#pragma unroll
        for (int q = 0; q < FPS_PPT; ++q) {
            const int p = tid + q * FPS_THREADS;
            px[q] = py[q] = pz[q] = 0.f; pd[q] = -1.f;          // padding never wins the arg-max
            if (p < N) { px[q] = P[p * 3]; py[q] = P[p * 3 + 1]; pz[q] = P[p * 3 + 2]; pd[q] = 1e10f; }
        }
    } else {
        for (int p = tid; p < N; p += FPS_THREADS) dist_g[p] = 1e10f;
    }
    int far = (int)start[b];
    for (int it = 0; it < K; ++it) {
        if (tid == 0) out[(size_t)b * K + it] = far;
        const float cx = __ldg(P + far * 3), cy = __ldg(P + far * 3 + 1), cz = __ldg(P + far * 3 + 2);
        float bv = -1.f; int bi = 0x7fffffff;
        if (kRegs) {
#pragma unroll
            for (int q = 0; q < FPS_PPT; ++q) {
                const int p = tid + q * FPS_THREADS;
                if (p < N) {
                    const float d = fps_d2(px[q], py[q], pz[q], cx, cy, cz);
                    if (d < pd[q]) pd[q] = d;
                    if (pd[q] > bv) { bv = pd[q]; bi = p; }     // ascending p: strict '>' keeps the first
                }
            }
        } else {
            for (int p = tid; p < N; p += FPS_THREADS) {
                const float d = fps_d2(__ldg(P + p * 3), __ldg(P + p * 3 + 1), __ldg(P + p * 3 + 2), cx, cy, cz);
                float cur = dist_g[p];
                if (d < cur) { cur = d; dist_g[p] = d; }
                if (cur > bv) { bv = cur; bi = p; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            argmax_combine(bv, bi, __shfl_xor_sync(0xffffffffu, bv, o), __shfl_xor_sync(0xffffffffu, bi, o));
        if (lane == 0) { s_v[wid] = bv; s_i[wid] = bi; }
        __syncthreads();
        if (wid == 0) {
            bv = s_v[lane]; bi = s_i[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                argmax_combine(bv, bi, __shfl_xor_sync(0xffffffffu, bv, o), __shfl_xor_sync(0xffffffffu, bi, o));
            if (lane == 0) s_far = bi;
        }
        __syncthreads();
        far = s_far;
    }
}

// ------------------------------------------------------------------------------------------------
// FPS on a thread-block CLUSTER: one cluster (CS CTAs x 1024 threads) per cloud.  The points of the cloud are
// spread over the CTAs and live in registers together with their running distance (<= FPSC_PPT per thread), so an
// iteration is: local min-update + arg-max, block reduce, exchange of the CS block results through distributed
// shared memory (every CTA writes its (value, index, x, y, z) into slot [parity][rank] of EVERY CTA of the
// cluster), ONE cluster barrier, redundant final selection.  Slots are double-buffered by iteration parity, which
// is what makes a single barrier per iteration sufficient.  Selection rule ((value desc, index asc)) and arithmetic
// are those of fps_kernel, so the node lists stay bit-identical to the reference.
// ------------------------------------------------------------------------------------------------
constexpr int FPSC_SMEM_PPT = 14;  // shared-memory variant: 14 x 1024 x 16 B = 224 KB per CTA
constexpr int FPSC_PPT = 8;       // 8 x 4 registers of point state per thread (1024-thread CTAs have 64 registers)

struct __align__(16) FpsSlot { float v; int i; float x, y, z; float pad[3]; };   // 32 bytes: 16-byte aligned remote stores

__device__ __forceinline__ uint32_t fps_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t fps_mapa(uint32_t saddr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void fps_st_remote(uint32_t addr, const FpsSlot& s) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(__float_as_uint(s.v)), "r"((unsigned)s.i),
                 "r"(__float_as_uint(s.x)), "r"(__float_as_uint(s.y)) : "memory");
    asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(addr + 16), "r"(__float_as_uint(s.z)) : "memory");
}
__device__ __forceinline__ void fps_cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// kSmem: the CTA's points and running distances live in dynamic shared memory ([ppt][1024] per array, conflict-free)
// instead of registers: up to 14 points per thread = 229376 points per 16-CTA cluster (a 200k cloud: 3.1 s -> see DESIGN).
// Same arithmetic and the same ascending visiting order per thread: identical node lists.
template <int CS, bool kSmem>
__global__ void __launch_bounds__(FPS_THREADS)
fps_cluster_kernel(const float* __restrict__ xyz, int N, int K, const int64_t* __restrict__ start, int64_t* __restrict__ out, int ppt) {
    extern __shared__ float fps_dyn[];
    __shared__ float s_v[32];
    __shared__ int s_i[32];
    __shared__ FpsSlot s_slot[2][16];                        // [parity][rank]
    const int b = blockIdx.x / CS;
    const uint32_t rank = fps_cluster_rank();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float* P = xyz + (size_t)b * N * 3;
    // point p of the cloud belongs to CTA (p / 1024) % CS, thread p % 1024, register (p / 1024) / CS
    float px[FPSC_PPT], py[FPSC_PPT], pz[FPSC_PPT], pd[FPSC_PPT];
    float* sx = fps_dyn; float* sy = sx + ppt * FPS_THREADS; float* sz = sy + ppt * FPS_THREADS; float* sd = sz + ppt * FPS_THREADS;
    if (kSmem) {
        for (int q = 0; q < ppt; ++q) {
            const int p = (q * CS + (int)rank) * FPS_THREADS + tid;
            const int e = q * FPS_THREADS + tid;
            sx[e] = sy[e] = sz[e] = 0.f; sd[e] = -1.f;
            if (p < N) { sx[e] = P[p * 3]; sy[e] = P[p * 3 + 1]; sz[e] = P[p * 3 + 2]; sd[e] = 1e10f; }
        }
    } else {
#pragma unroll
        for (int q = 0; q < FPSC_PPT; ++q) {
            const int p = (q * CS + (int)rank) * FPS_THREADS + tid;
            px[q] = py[q] = pz[q] = 0.f; pd[q] = -1.f;            // padding never wins the arg-max
            if (p < N) { px[q] = P[p * 3]; py[q] = P[p * 3 + 1]; pz[q] = P[p * 3 + 2]; pd[q] = 1e10f; }
        }
    }
    int far = (int)start[b];
    float cx = __ldg(P + far * 3), cy = __ldg(P + far * 3 + 1), cz = __ldg(P + far * 3 + 2);
    fps_cluster_barrier();                                    // all CTAs of the cluster are running before remote stores
    for (int it = 0; it < K; ++it) {
        if (rank == 0 && tid == 0) out[(size_t)b * K + it] = far;
        float bv = -1.f; int bi = 0x7fffffff; float bx = 0.f, by = 0.f, bz = 0.f;
        if (kSmem) {
            for (int q = 0; q < ppt; ++q) {
                const int p = (q * CS + (int)rank) * FPS_THREADS + tid;
                const int e = q * FPS_THREADS + tid;
                if (p < N) {
                    const float x = sx[e], y = sy[e], z = sz[e];
                    float dd = sd[e];
                    const float d = fps_d2(x, y, z, cx, cy, cz);
                    if (d < dd) { dd = d; sd[e] = d; }
                    if (dd > bv) { bv = dd; bi = p; bx = x; by = y; bz = z; }
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < FPSC_PPT; ++q) {
                const int p = (q * CS + (int)rank) * FPS_THREADS + tid;
                if (p < N) {
                    const float d = fps_d2(px[q], py[q], pz[q], cx, cy, cz);
                    if (d < pd[q]) pd[q] = d;
                    if (pd[q] > bv) { bv = pd[q]; bi = p; bx = px[q]; by = py[q]; bz = pz[q]; }     // ascending p per thread: strict '>' keeps the first
                }
            }
        }
        // warp arg-max (value desc, index asc), carrying the coordinates
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const float ox = __shfl_xor_sync(0xffffffffu, bx, o), oy = __shfl_xor_sync(0xffffffffu, by, o), oz = __shfl_xor_sync(0xffffffffu, bz, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bx = ox; by = oy; bz = oz; }
        }
        if (lane == 0) { s_v[wid] = bv; s_i[wid] = bi; }
        __shared__ float s_x[32], s_y[32], s_z[32];
        if (lane == 0) { s_x[wid] = bx; s_y[wid] = by; s_z[wid] = bz; }
        __syncthreads();
        if (wid == 0) {
            bv = s_v[lane]; bi = s_i[lane]; bx = s_x[lane]; by = s_y[lane]; bz = s_z[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                const float ox = __shfl_xor_sync(0xffffffffu, bx, o), oy = __shfl_xor_sync(0xffffffffu, by, o), oz = __shfl_xor_sync(0xffffffffu, bz, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bx = ox; by = oy; bz = oz; }
            }
            if (lane < CS) {                                   // lane r publishes this CTA's result into CTA r
                FpsSlot s{bv, bi, bx, by, bz, {0.f, 0.f, 0.f}};
                const uint32_t local = (uint32_t)__cvta_generic_to_shared(&s_slot[it & 1][rank]);
                fps_st_remote(fps_mapa(local, (uint32_t)lane), s);
            }
        }
        fps_cluster_barrier();                                 // also orders the block (all threads take part)
        // final selection over the CS block results (same order in every CTA)
        float gv = -2.f; int gi = 0x7fffffff;
#pragma unroll
        for (int r = 0; r < CS; ++r) {
            const FpsSlot s = s_slot[it & 1][r];
            if (s.v > gv || (s.v == gv && s.i < gi)) { gv = s.v; gi = s.i; cx = s.x; cy = s.y; cz = s.z; }
        }
        far = gi;
    }
    fps_cluster_barrier();                                     // nobody exits while a peer may still write into its shared memory
}

template <int CS, bool kSmem>
static int launch_fps_cluster(const float* xyz, int B, int N, int K, const int64_t* start, int64_t* out, cudaStream_t st) {
    auto kern = fps_cluster_kernel<CS, kSmem>;
    if (CS > 8) DVM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    const int ppt = kSmem ? ceil_div(N, CS * FPS_THREADS) : 0;
    const size_t dyn = (size_t)ppt * FPS_THREADS * 4 * sizeof(float);
    if (kSmem) DVM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(B * CS); cfg.blockDim = dim3(FPS_THREADS); cfg.dynamicSmemBytes = dyn; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    DVM_CUDA(cudaLaunchKernelEx(&cfg, kern, xyz, N, K, start, out, ppt));
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// graph tensors
// ------------------------------------------------------------------------------------------------
__global__ void gather_nodes_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ nodes_idx,
                                    int N, int K, float* __restrict__ nodes_xyz) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const int64_t v = nodes_idx[(size_t)b * K + i];
    const float* p = xyz + ((size_t)b * N + v) * 3;
    float* o = nodes_xyz + ((size_t)b * K + i) * 3;
    o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
}

// sigma[b] = 20 * mean_i sqrt(d2nn[b,i,1])   (fp64, deterministic: one CTA per cloud, fixed order)
__global__ void __launch_bounds__(1024) sigma_kernel(const double* __restrict__ d2nn, int N, double* __restrict__ sigma) {
    __shared__ double s[32];
    const int b = blockIdx.x;
    double acc = 0.0;
    for (int i = threadIdx.x; i < N; i += 1024) acc += sqrt(d2nn[((size_t)b * N + i) * 2 + 1]);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = warp_sum(s[threadIdx.x]);
        if (threadIdx.x == 0) sigma[b] = 20.0 * acc / (double)N;
    }
}

// dists = sqrt(d2); w = exp(-(dists^2) / (float)(2 sigma^2)); w /= sum   (fp32, as the reference's
// fp32-tensor / 0-dim-f64-tensor arithmetic evaluates it)
__global__ void weights_kernel(const float* __restrict__ d2, const double* __restrict__ sigma, int N,
                               float* __restrict__ dists, float* __restrict__ weights) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double sg = sigma[b];
    const float denom = (float)(2.0 * sg * sg);
    float w[3], s = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float d = sqrtf(d2[((size_t)b * N + i) * 3 + k]);
        dists[((size_t)b * N + i) * 3 + k] = d;
        w[k] = expf(-(d * d) / denom);
    }
    s = (w[0] + w[1]) + w[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) weights[((size_t)b * N + i) * 3 + k] = w[k] / s;
}

}  // namespace dvm

using namespace dvm;

extern "C" size_t dvm_fps_workspace_bytes(int B, int N) {
    if (B <= 0 || N <= 0) return 0;
    return N > FPS_THREADS * FPS_PPT ? align_up((size_t)B * N * sizeof(float), 256) : 256;
}

extern "C" int dvm_fps(const float* xyz, int B, int N, int K, const int64_t* start, int64_t* out,
                       void* ws, size_t ws_bytes, void* stream) {
    DVM_CHECK_ARG(xyz && start && out, "dvm_fps: null pointer");
    DVM_CHECK_ARG(B > 0 && N > 0 && K > 0 && K <= N, "dvm_fps: bad sizes (B=%d N=%d K=%d)", B, N, K);
    cudaStream_t st = (cudaStream_t)stream;
    // few big clouds: a cluster of CTAs per cloud (points in registers, results exchanged through distributed shared
    // memory); many small clouds: one CTA each already fills the machine
    if (N > FPS_THREADS * FPS_PPT && (long long)B * 8 <= 4 * kNumSM) {      // one CTA holds <= 8192 points in registers (1.1 us / iteration)
        if (N <= 8 * FPS_THREADS * FPSC_PPT) return launch_fps_cluster<8, false>(xyz, B, N, K, start, out, st);
        if (N <= 16 * FPS_THREADS * FPSC_PPT) return launch_fps_cluster<16, false>(xyz, B, N, K, start, out, st);
        if (N <= 16 * FPS_THREADS * FPSC_SMEM_PPT && (long long)B * 16 <= kNumSM)     // points in shared memory: <= 229376 per cloud
            return launch_fps_cluster<16, true>(xyz, B, N, K, start, out, st);
    }
    if (N <= FPS_THREADS * FPS_PPT) {
        fps_kernel<true><<<B, FPS_THREADS, 0, st>>>(xyz, N, K, start, out, nullptr);
    } else {
        if (!ws || ws_bytes < dvm_fps_workspace_bytes(B, N)) {
            set_error("dvm_fps: workspace too small");
            return DVM_ERR_WORKSPACE;
        }
        fps_kernel<false><<<B, FPS_THREADS, 0, st>>>(xyz, N, K, start, out, (float*)ws);
    }
    DVM_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t dvm_graph_workspace_bytes(int B, int N, int K) {
    if (B <= 0 || N <= 0 || K <= 0) return 0;
    WsCarver ws(nullptr, 0);
    ws.take<float>((size_t)B * K * 3);      // node coordinates
    ws.take<float>((size_t)B * N * 3);      // squared distances vertex -> 3 nodes
    ws.take<double>((size_t)B * N * 2);     // fp64 squared NN distances
    ws.take<char>(knn3_grid_workspace_bytes(B, N, N));   // grid scratch of the three k-NN queries (largest: N points)
    return align_up(ws.off, 256);
}

extern "C" int dvm_graph_weights(const float* xyz, const int64_t* nodes_idx, int B, int N, int K,
                                 int64_t* influence, float* dists, float* weights, int64_t* ring, double* sigma,
                                 void* wsp, size_t ws_bytes, void* stream) {
    DVM_CHECK_ARG(xyz && nodes_idx && influence && dists && weights && ring && sigma, "dvm_graph_weights: null pointer");
    DVM_CHECK_ARG(B > 0 && N >= 2 && K >= 9 && K <= N && B <= 65535, "dvm_graph_weights: bad sizes (B=%d N=%d K=%d)", B, N, K);
    if (!wsp || ws_bytes < dvm_graph_workspace_bytes(B, N, K)) {
        set_error("dvm_graph_weights: workspace too small");
        return DVM_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    WsCarver ws(wsp, ws_bytes);
    float* nodes_xyz = ws.take<float>((size_t)B * K * 3);
    float* d2 = ws.take<float>((size_t)B * N * 3);
    double* d2nn = ws.take<double>((size_t)B * N * 2);
    const size_t gbytes = knn3_grid_workspace_bytes(B, N, N);
    void* gws = ws.take<char>(gbytes);
    gather_nodes_kernel<<<dim3(ceil_div(K, 256), B), 256, 0, st>>>(xyz, nodes_idx, N, K, nodes_xyz);
    DVM_LAUNCH_CHECK();
    int rc;
    // 3 nearest nodes per vertex, exact fp32 form (topk(3) of -geod[nodes].T, :186-188)
    if ((rc = launch_knn3_auto(xyz, nodes_xyz, B, N, K, 3, false, influence, nullptr, d2, nullptr, gws, gbytes, st))) return rc;
    // node ring: KDTree(nodes).query(nodes, 9), fp64 (:181-183)
    if ((rc = launch_knn3_auto(nodes_xyz, nodes_xyz, B, K, K, 9, true, ring, nullptr, nullptr, nullptr, gws, gbytes, st))) return rc;
    // sigma = 20 * mean NN spacing: KDTree(vertices).query(vertices, 2)[:,1], fp64 (:190-192)
    if ((rc = launch_knn3_auto(xyz, xyz, B, N, N, 2, true, nullptr, nullptr, nullptr, d2nn, gws, gbytes, st))) return rc;
    sigma_kernel<<<B, 1024, 0, st>>>(d2nn, N, sigma);
    DVM_LAUNCH_CHECK();
    weights_kernel<<<dim3(ceil_div(N, 256), B), 256, 0, st>>>(d2, sigma, N, dists, weights);
    DVM_LAUNCH_CHECK();
    return 0;
}
