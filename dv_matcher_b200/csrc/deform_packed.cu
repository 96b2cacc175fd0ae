// HBM-shaped form of the deformation-graph forward/backward (lib/deformation_graph_point.py:233-261, models/loss.py:39-45,
// 1257-1264): the graph is static per shape, so everything that only depends on the graph is laid out ONCE for coalesced,
// vectorised streaming (built by dvm_graph_pack-side host code in deformation_graph.build_graphs):
//
//   node record  float4[4] = 64 B, 64-byte aligned:  q0 = R0..R3 | q1 = R4..R7 | q2 = R8, t0, t1, t2 | q3 = g0, g1, g2, 0
//                (t and g share one 32-byte sector: ARAP without the smoothness term reads ONE sector per ring neighbour);
//   NODES ARE RENUMBERED in Morton order of their positions (node_perm[new] = old; the Deformer is asked for its rows in that
//   order, so the per-step node table comes out Morton-ordered for free): spatial neighbours are memory neighbours, the
//   records a warp of neighbouring vertices gathers share sectors and stay in L1;
//   vorder  int32[B][N]     vertices in Morton order of their coordinates, s_xyz f32[B][N][3] the same vertices' coordinates
//                           (static per shape: streamed, not gathered); the warped points are scattered back through vorder;
//   s_infl  int32[B][3][N], s_w f32[B][3][N]   influence lists (new node numbers) of vertex vorder[i], slot-major: every load
//                           of a warp is one contiguous 128-byte line;
//   s_ring int32[B][9][K]   ring of (new) node i, new node numbers, slot-major;
//   csr_ptr int32[B][K+1], csr_vert int32[B][3N], csr_w f32[B][3N]   vertex lists per (new) node for the skinning backward: one
//                           thread owns a node and sums its vertices in ascending vertex order -- no atomics, deterministic.
//
// DRAM bytes per vertex (skinning): vorder 4 + s_infl 12 + s_w 12 + xyz 12 + out 12 = 52, + node records 64 B / node = 32 B /
// vertex at K = N/2: 84 B (SURVEY 8d counts 78 with 60-byte unpadded records).  ARAP per node: record 64 + norder 4 + ring 36.
#include "common.cuh"

namespace dvm {

namespace {
struct P3 { float x, y, z; };
__device__ __forceinline__ P3 p3(float x, float y, float z) { return P3{x, y, z}; }
__device__ __forceinline__ P3 operator+(P3 a, P3 b) { return p3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ P3 operator-(P3 a, P3 b) { return p3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ P3 operator*(float s, P3 a) { return p3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot3(P3 a, P3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ P3 cross3(P3 a, P3 b) { return p3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ P3 ldp3(const float* p) { return p3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }
}  // namespace

// ------------------------------------------------------------------------------------------------
// node table.  One warp = 32 consecutive nodes: inputs are read as contiguous float runs (coalesced), transposed through a
// per-warp shared-memory tile, records leave as sixteen coalesced 512-byte float4 stores.
//   kFromD9 = false: R[n][9], t[n][3] given (the reference's forward(vertices, R, t) signature)
//   kFromD9 = true : d9[n][9] = Deformer output: t = d9[0:3], 6D = d9[3:9] + [1,0,0,0,1,0], R = rotation_6d_to_matrix(6D)
//                    (models/loss.py:1257-1264 fused with 39-45; same arithmetic as rot6d_fwd_kernel)
// ------------------------------------------------------------------------------------------------
constexpr int NT_WARPS = 8;

// perm (may be null): input row of node i is perm[i] within its cloud of K nodes (inputs in the caller's node order, table in
// Morton order); with a permutation the 48-byte inputs are gathered per node instead of streamed.
template <bool kFromD9>
__global__ void __launch_bounds__(NT_WARPS * 32)
node_table_kernel(const float* __restrict__ Rin, const float* __restrict__ tin, const float* __restrict__ g, int n,
                  const int* __restrict__ perm, int K,
                  float4* __restrict__ table, float* __restrict__ R_out, float* __restrict__ t_out) {
    __shared__ float s_in[NT_WARPS][32 * 9 + 32 * 3 + 32 * 3];
    __shared__ __align__(16) float s_rec[NT_WARPS][32 * 16];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int node0 = (blockIdx.x * NT_WARPS + w) * 32;
    if (node0 >= n) return;
    const int cnt = min(32, n - node0);
    float* in = s_in[w];
    {
        if (perm == nullptr) {
            const float* src = Rin + (size_t)node0 * 9;                   // R rows or d9 rows: 9 floats per node
            for (int e = lane; e < cnt * 9; e += 32) in[e] = __ldg(src + e);
            if (!kFromD9) {
                const float* ts = tin + (size_t)node0 * 3;
                for (int e = lane; e < cnt * 3; e += 32) in[288 + e] = __ldg(ts + e);
            }
        } else if (lane < cnt) {
            const int i = node0 + lane;
            const size_t srow = (size_t)(i / K) * K + __ldg(perm + i);
#pragma unroll
            for (int c = 0; c < 9; ++c) in[lane * 9 + c] = __ldg(Rin + srow * 9 + c);
            if (!kFromD9) {
#pragma unroll
                for (int c = 0; c < 3; ++c) in[288 + lane * 3 + c] = __ldg(tin + srow * 3 + c);
            }
        }
        const float* gs = g + (size_t)node0 * 3;
        for (int e = lane; e < cnt * 3; e += 32) in[384 + e] = __ldg(gs + e);
    }
    __syncwarp();
    float r[9], t[3], gg[3];
    if (lane < cnt) {
        const float* p = in + lane * 9;
        if (kFromD9) {
            t[0] = p[0]; t[1] = p[1]; t[2] = p[2];
            const P3 a1 = p3(p[3] + 1.0f, p[4] + 0.0f, p[5] + 0.0f), a2 = p3(p[6] + 0.0f, p[7] + 1.0f, p[8] + 0.0f);
            const P3 b1 = (1.f / fmaxf(sqrtf(dot3(a1, a1)), 1e-12f)) * a1;        // F.normalize, eps 1e-12
            const P3 u = a2 - dot3(b1, a2) * b1;
            const P3 b2 = (1.f / fmaxf(sqrtf(dot3(u, u)), 1e-12f)) * u;
            const P3 b3 = cross3(b1, b2);
            r[0] = b1.x; r[1] = b1.y; r[2] = b1.z; r[3] = b2.x; r[4] = b2.y; r[5] = b2.z; r[6] = b3.x; r[7] = b3.y; r[8] = b3.z;
        } else {
#pragma unroll
            for (int c = 0; c < 9; ++c) r[c] = p[c];
            t[0] = in[288 + lane * 3]; t[1] = in[288 + lane * 3 + 1]; t[2] = in[288 + lane * 3 + 2];
        }
        gg[0] = in[384 + lane * 3]; gg[1] = in[384 + lane * 3 + 1]; gg[2] = in[384 + lane * 3 + 2];
        // record -> shared tile; float4 slot q of node `lane` is stored at slot q ^ ((lane >> 1) & 3): conflict-free v4 stores
        float4* rec = reinterpret_cast<float4*>(s_rec[w]) + lane * 4;
        const int sw = (lane >> 1) & 3;
        rec[0 ^ sw] = make_float4(r[0], r[1], r[2], r[3]);
        rec[1 ^ sw] = make_float4(r[4], r[5], r[6], r[7]);
        rec[2 ^ sw] = make_float4(r[8], t[0], t[1], t[2]);
        rec[3 ^ sw] = make_float4(gg[0], gg[1], gg[2], 0.f);
    }
    __syncwarp();
    {
        const float4* rec = reinterpret_cast<const float4*>(s_rec[w]);
        float4* dst = table + (size_t)node0 * 4;
        for (int f = lane; f < cnt * 4; f += 32) {
            const int nd = f >> 2, q = f & 3;
            dst[f] = rec[nd * 4 + (q ^ ((nd >> 1) & 3))];
        }
    }
    if (kFromD9 && R_out) {                      // optional plain copies (callers that want R, t as tensors)
        __syncwarp();
        if (lane < cnt) {
#pragma unroll
            for (int c = 0; c < 9; ++c) in[lane * 9 + c] = r[c];
            in[288 + lane * 3] = t[0]; in[288 + lane * 3 + 1] = t[1]; in[288 + lane * 3 + 2] = t[2];
        }
        __syncwarp();
        for (int e = lane; e < cnt * 9; e += 32) R_out[(size_t)node0 * 9 + e] = in[e];
        for (int e = lane; e < cnt * 3; e += 32) t_out[(size_t)node0 * 3 + e] = in[288 + e];
    }
}

// ------------------------------------------------------------------------------------------------
// skinning forward on the packed layout: thread = one vertex in Morton order
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
skin_fwd_packed_kernel(const float* __restrict__ s_xyz, const int* __restrict__ vorder, const int* __restrict__ s_infl,
                       const float* __restrict__ s_w, const float4* __restrict__ table, int N, int K, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int vid = __ldg(vorder + (size_t)b * N + i);
    const P3 v = ldp3(s_xyz + ((size_t)b * N + i) * 3);
    const float4* tb = table + (size_t)b * K * 4;
    int nk[3]; float wk[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        nk[k] = __ldg(s_infl + ((size_t)b * 3 + k) * N + i);
        wk[k] = __ldg(s_w + ((size_t)b * 3 + k) * N + i);
    }
    P3 acc = p3(0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float4* rec = tb + (size_t)nk[k] * 4;
        const float4 q0 = __ldg(rec), q1 = __ldg(rec + 1), q2 = __ldg(rec + 2), q3 = __ldg(rec + 3);
        const P3 g = p3(q3.x, q3.y, q3.z), tn = p3(q2.y, q2.z, q2.w);
        const P3 x = v - g;
        const P3 y = p3(q0.x * x.x + q0.y * x.y + q0.z * x.z, q0.w * x.x + q1.x * x.y + q1.y * x.z,
                        q1.z * x.x + q1.w * x.y + q2.x * x.z) + g + tn;
        acc = acc + wk[k] * y;
    }
    float* o = out + ((size_t)b * N + vid) * 3;
    o[0] = acc.x; o[1] = acc.y; o[2] = acc.z;
}

// ------------------------------------------------------------------------------------------------
// skinning backward, node-major: thread = one node, sums over the vertices it influences (ascending vertex id)
//   dt_n = sum w dOut_v,   dR_n = sum w dOut_v (x) (v - g_n)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
skin_bwd_csr_kernel(const float* __restrict__ xyz, const float* __restrict__ nodes_xyz, const int* __restrict__ csr_ptr,
                    const int* __restrict__ csr_vert, const float* __restrict__ csr_w, const float* __restrict__ dOut,
                    int N, int K, float* __restrict__ dR, float* __restrict__ dt) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= K) return;
    const int* ptr = csr_ptr + (size_t)b * (K + 1);
    const int e0 = __ldg(ptr + n), e1 = __ldg(ptr + n + 1);
    const P3 g = ldp3(nodes_xyz + ((size_t)b * K + n) * 3);
    const float* P = xyz + (size_t)b * N * 3;
    const float* G = dOut + (size_t)b * N * 3;
    float r[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    P3 tt = p3(0.f, 0.f, 0.f);
    for (int e = e0; e < e1; ++e) {
        const int vid = __ldg(csr_vert + (size_t)b * 3 * N + e);
        const float w = __ldg(csr_w + (size_t)b * 3 * N + e);
        const P3 x = ldp3(P + (size_t)vid * 3) - g;
        const P3 gw = w * ldp3(G + (size_t)vid * 3);
        tt = tt + gw;
        r[0] += gw.x * x.x; r[1] += gw.x * x.y; r[2] += gw.x * x.z;
        r[3] += gw.y * x.x; r[4] += gw.y * x.y; r[5] += gw.y * x.z;
        r[6] += gw.z * x.x; r[7] += gw.z * x.y; r[8] += gw.z * x.z;
    }
    float* ro = dR + ((size_t)b * K + n) * 9;
#pragma unroll
    for (int c = 0; c < 9; ++c) ro[c] = r[c];
    float* to = dt + ((size_t)b * K + n) * 3;
    to[0] = tt.x; to[1] = tt.y; to[2] = tt.z;
}

// ------------------------------------------------------------------------------------------------
// ARAP (+ optional rotation smoothness) on the packed layout: thread = one node (nodes are numbered in Morton order);
// per-block partial sums, fixed-order final sum (deterministic)
// ------------------------------------------------------------------------------------------------
constexpr int AP_THREADS = 256;

template <bool kSr>
__global__ void __launch_bounds__(AP_THREADS)
arap_fwd_packed_kernel(const int* __restrict__ s_ring, const float4* __restrict__ table,
                       int K, int ring_k, float* __restrict__ part /* [B][gridDim.x][2] */) {
    __shared__ float s[2][AP_THREADS / 32];
    const int b = blockIdx.y;
    const int i = blockIdx.x * AP_THREADS + threadIdx.x;
    float a_sum = 0.f, r_sum = 0.f;
    if (i < K) {
        const float4* tb = table + (size_t)b * K * 4;
        const float4* rec = tb + (size_t)i * 4;
        const float4 q0 = __ldg(rec), q1 = __ldg(rec + 1), q2 = __ldg(rec + 2), q3 = __ldg(rec + 3);
        const P3 gi = p3(q3.x, q3.y, q3.z), ti = p3(q2.y, q2.z, q2.w);
        for (int q = 0; q < ring_k; ++q) {
            const int j = __ldg(s_ring + ((size_t)b * ring_k + q) * K + i);
            const float4* rj = tb + (size_t)j * 4;
            const float4 j2 = __ldg(rj + 2), j3 = __ldg(rj + 3);
            const P3 gj = p3(j3.x, j3.y, j3.z), tj = p3(j2.y, j2.z, j2.w);
            const P3 e = gi - gj;
            const P3 d = ((gi + ti) - (gj + tj)) - p3(q0.x * e.x + q0.y * e.y + q0.z * e.z, q0.w * e.x + q1.x * e.y + q1.y * e.z,
                                                       q1.z * e.x + q1.w * e.y + q2.x * e.z);
            a_sum += dot3(d, d);
            if (kSr) {
                const float4 j0 = __ldg(rj), j1 = __ldg(rj + 1);
                float dd;
                dd = q0.x - j0.x; r_sum += dd * dd; dd = q0.y - j0.y; r_sum += dd * dd; dd = q0.z - j0.z; r_sum += dd * dd;
                dd = q0.w - j0.w; r_sum += dd * dd; dd = q1.x - j1.x; r_sum += dd * dd; dd = q1.y - j1.y; r_sum += dd * dd;
                dd = q1.z - j1.z; r_sum += dd * dd; dd = q1.w - j1.w; r_sum += dd * dd; dd = q2.x - j2.x; r_sum += dd * dd;
            }
        }
    }
    a_sum = warp_sum(a_sum);
    if (kSr) r_sum = warp_sum(r_sum);
    if ((threadIdx.x & 31) == 0) { s[0][threadIdx.x >> 5] = a_sum; s[1][threadIdx.x >> 5] = r_sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, r = 0.f;
        for (int w = 0; w < AP_THREADS / 32; ++w) { a += s[0][w]; r += s[1][w]; }
        float* o = part + ((size_t)b * gridDim.x + blockIdx.x) * 2;
        o[0] = a; o[1] = r;
    }
}

__global__ void arap_final_packed_kernel(const float* __restrict__ part, int nblk, int K, int ring_k,
                                         float* __restrict__ arap, float* __restrict__ sr) {
    const int b = blockIdx.x;
    // fixed order: lane l sums blocks l, l+32, ... then a butterfly -- independent of launch timing
    float a = 0.f, r = 0.f;
    for (int k = threadIdx.x; k < nblk; k += 32) { a += part[((size_t)b * nblk + k) * 2]; r += part[((size_t)b * nblk + k) * 2 + 1]; }
    a = warp_sum(a); r = warp_sum(r);
    if (threadIdx.x == 0) {
        arap[b] = a / (float)K;
        if (sr) sr[b] = r / ((float)K * (float)ring_k * 9.f);
    }
}

}  // namespace dvm

using namespace dvm;

extern "C" int dvm_node_table(const float* R, const float* t, const float* nodes_xyz, const int32_t* node_perm, int B, int K, float* table, void* stream) {
    DVM_CHECK_ARG(R && t && nodes_xyz && table, "dvm_node_table: null pointer");
    DVM_CHECK_ARG(B > 0 && K > 0 && (long long)B * K < 0x7fffffffLL, "dvm_node_table: bad sizes");
    DVM_CHECK_ARG(((uintptr_t)table & 63) == 0, "dvm_node_table: table must be 64-byte aligned");
    const int n = B * K;
    node_table_kernel<false><<<ceil_div(n, NT_WARPS * 32), NT_WARPS * 32, 0, (cudaStream_t)stream>>>(
        R, t, nodes_xyz, n, node_perm, K, reinterpret_cast<float4*>(table), nullptr, nullptr);
    DVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int dvm_node_table_from_d9(const float* d9, const float* nodes_xyz, const int32_t* node_perm, int B, int K, float* table,
                                      float* R_out, float* t_out, void* stream) {
    DVM_CHECK_ARG(d9 && nodes_xyz && table, "dvm_node_table_from_d9: null pointer");
    DVM_CHECK_ARG((R_out == nullptr) == (t_out == nullptr), "dvm_node_table_from_d9: R_out and t_out come together");
    DVM_CHECK_ARG(B > 0 && K > 0 && (long long)B * K < 0x7fffffffLL, "dvm_node_table_from_d9: bad sizes");
    DVM_CHECK_ARG(((uintptr_t)table & 63) == 0, "dvm_node_table_from_d9: table must be 64-byte aligned");
    const int n = B * K;
    node_table_kernel<true><<<ceil_div(n, NT_WARPS * 32), NT_WARPS * 32, 0, (cudaStream_t)stream>>>(
        d9, nullptr, nodes_xyz, n, node_perm, K, reinterpret_cast<float4*>(table), R_out, t_out);
    DVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int dvm_skin_fwd_packed(const float* s_xyz, const int32_t* vorder, const int32_t* s_infl, const float* s_w,
                                   const float* table, int B, int N, int K, float* out, void* stream) {
    DVM_CHECK_ARG(s_xyz && vorder && s_infl && s_w && table && out, "dvm_skin_fwd_packed: null pointer");
    DVM_CHECK_ARG(B > 0 && N > 0 && K > 0 && B <= 65535, "dvm_skin_fwd_packed: bad sizes");
    skin_fwd_packed_kernel<<<dim3(ceil_div(N, 256), B), 256, 0, (cudaStream_t)stream>>>(
        s_xyz, vorder, s_infl, s_w, reinterpret_cast<const float4*>(table), N, K, out);
    DVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int dvm_skin_bwd_csr(const float* xyz, const float* nodes_xyz, const int32_t* csr_ptr, const int32_t* csr_vert,
                                const float* csr_w, const float* dOut, int B, int N, int K, float* dR, float* dt, void* stream) {
    DVM_CHECK_ARG(xyz && nodes_xyz && csr_ptr && csr_vert && csr_w && dOut && dR && dt, "dvm_skin_bwd_csr: null pointer");
    DVM_CHECK_ARG(B > 0 && N > 0 && K > 0 && B <= 65535, "dvm_skin_bwd_csr: bad sizes");
    skin_bwd_csr_kernel<<<dim3(ceil_div(K, 128), B), 128, 0, (cudaStream_t)stream>>>(xyz, nodes_xyz, csr_ptr, csr_vert, csr_w, dOut, N, K, dR, dt);
    DVM_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t dvm_arap_packed_workspace_bytes(int B, int K) {
    if (B <= 0 || K <= 0) return 0;
    return align_up((size_t)B * ceil_div(K, AP_THREADS) * 2 * sizeof(float), 256);
}

extern "C" int dvm_arap_fwd_packed(const int32_t* s_ring, const float* table, int B, int K, int ring_k,
                                   float* arap, float* sr, void* ws, size_t ws_bytes, void* stream) {
    DVM_CHECK_ARG(s_ring && table && arap, "dvm_arap_fwd_packed: null pointer");
    DVM_CHECK_ARG(B > 0 && K > 0 && ring_k > 0 && B <= 65535, "dvm_arap_fwd_packed: bad sizes");
    if (!ws || ws_bytes < dvm_arap_packed_workspace_bytes(B, K)) { set_error("dvm_arap_fwd_packed: workspace too small"); return DVM_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = ceil_div(K, AP_THREADS);
    const float4* tb = reinterpret_cast<const float4*>(table);
    if (sr) arap_fwd_packed_kernel<true><<<dim3(nblk, B), AP_THREADS, 0, st>>>(s_ring, tb, K, ring_k, (float*)ws);
    else    arap_fwd_packed_kernel<false><<<dim3(nblk, B), AP_THREADS, 0, st>>>(s_ring, tb, K, ring_k, (float*)ws);
    DVM_LAUNCH_CHECK();
    arap_final_packed_kernel<<<B, 32, 0, st>>>((const float*)ws, nblk, K, ring_k, arap, sr);
    DVM_LAUNCH_CHECK();
    return 0;
}
