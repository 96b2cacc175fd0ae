// Backward of the top-k soft map w.r.t. the features (autograd of models/loss.py:110-114 + 1339-1347).
//
// Forward:  P_ij = exp(-alpha (d_ij - rmin_i)) / Z_i  over ALL columns j,  w_ik = P_i,j(k) for the kept top-k.
// With g_ij = dW_ik on the kept entries and 0 elsewhere, c_i = sum_k dW_ik w_ik:
//     dL/dd_ij = -alpha P_ij (g_ij - c_i)            (softmax Jacobian; every column of the row takes part)
//     dX_i =  sum_j G_ij (x_i - y_j),   dY_j = sum_i G_ij (y_j - x_i),   G_ij = (dL/dd_ij) / d_ij   (0 where d_ij = 0,
//     as torch.cdist's backward does).
// Split as G = G_dense + G_topk with  G_dense_ij = alpha c_i P_ij / d_ij  (all j)  and
// G_topk_ik = -alpha w_ik dW_ik / d_ik  (10 entries per row):
//   * two dense tiled passes (fp32 CUDA cores, exact direct-difference distances like the forward's exact form),
//     one owning 64-row blocks of X (-> dX), one owning 64-column blocks of Y (-> dY): no atomics, deterministic;
//     tiles whose every P_ij underflows the softmax window skip their accumulation phase;
//   * a sparse pass for the top-k part (dX by the row's warp, dY with atomicAdd: 10 per row).
// This is the parity-first version; the tcgen05 flash-style backward is listed as next in DESIGN.md.
#include "softmap.cuh"

namespace dvm {

constexpr int BW_T = 64;              // tile: 64 "owner" items x 64 swept items
constexpr int BW_THREADS = 256;

// kOwnerIsRow: owner side = X rows (stats indexed by owner), swept side = Y columns; else owner = Y columns, swept = X rows
template <bool kOwnerIsRow>
__global__ void __launch_bounds__(BW_THREADS)
softmap_bwd_dense_kernel(const float* __restrict__ X, const float* __restrict__ Y, int N, int M, int C,
                         float alpha, float a2, float cut2,
                         const float* __restrict__ cvec, const float* __restrict__ rmin, const float* __restrict__ rsum,
                         float* __restrict__ dOwner) {
    extern __shared__ __align__(16) float sm[];
    const int ld = C + 4;
    float* As = sm;                          // [64][ld] owner tile
    float* Bs = As + BW_T * ld;              // [64][ld] swept tile
    float* Gs = Bs + BW_T * ld;              // [64][65]
    float* st_c = Gs + BW_T * (BW_T + 1);    // [64] stats of the ROW side of the current tile pair: alpha * c_i / Z_i
    float* st_m = st_c + BW_T;               // [64] rmin_i
    float* gsum = st_m + BW_T;               // [64] sum_b G_ab per owner item

    const int b = blockIdx.y;
    const int nOwner = kOwnerIsRow ? N : M, nSwept = kOwnerIsRow ? M : N;
    const float* Aown = (kOwnerIsRow ? X : Y) + (size_t)b * nOwner * C;
    const float* Bswp = (kOwnerIsRow ? Y : X) + (size_t)b * nSwept * C;
    const int a0 = blockIdx.x * BW_T;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int r = tid & 63, cg = tid >> 6;             // phase 1: owner item r, swept group cg (16 items)
    const int c4 = C >> 2;

    for (int e = tid; e < BW_T * c4; e += BW_THREADS) {
        const int rr = e / c4, cc = e - rr * c4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a0 + rr < nOwner) v = __ldg(reinterpret_cast<const float4*>(Aown + (size_t)(a0 + rr) * C) + cc);
        *reinterpret_cast<float4*>(As + rr * ld + cc * 4) = v;
    }
    if (tid < BW_T) gsum[tid] = 0.f;
    if (kOwnerIsRow && tid < BW_T) {
        const int i = a0 + tid;
        const bool ok = i < N;
        st_c[tid] = ok ? alpha * cvec[(size_t)b * N + i] / rsum[(size_t)b * N + i] : 0.f;
        st_m[tid] = ok ? rmin[(size_t)b * N + i] : 0.f;
    }
    // phase 2 ownership: warp wid -> owner rows wid*8 .. +7, lane -> channels 4*lane (+128)
    float acc[8][2][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][h][c] = 0.f;
    float gpart = 0.f;                                   // partial sum_b G for owner item r (this thread's 16 swept items per tile)

    for (int s0 = 0; s0 < nSwept; s0 += BW_T) {
        __syncthreads();
        for (int e = tid; e < BW_T * c4; e += BW_THREADS) {
            const int rr = e / c4, cc = e - rr * c4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s0 + rr < nSwept) v = __ldg(reinterpret_cast<const float4*>(Bswp + (size_t)(s0 + rr) * C) + cc);
            *reinterpret_cast<float4*>(Bs + rr * ld + cc * 4) = v;
        }
        if (!kOwnerIsRow && tid < BW_T) {
            const int i = s0 + tid;
            const bool ok = i < N;
            st_c[tid] = ok ? alpha * cvec[(size_t)b * N + i] / rsum[(size_t)b * N + i] : 0.f;
            st_m[tid] = ok ? rmin[(size_t)b * N + i] : 0.f;
        }
        __syncthreads();
        // ---- phase 1: squared distances (direct differences), G tile
        float d2[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) d2[c] = 0.f;
        const float* ar = As + r * ld;
        const float* br = Bs + (cg * 16) * ld;
#pragma unroll 2
        for (int k = 0; k < C; k += 4) {
            const float4 av = *reinterpret_cast<const float4*>(ar + k);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float4 bv = *reinterpret_cast<const float4*>(br + c * ld + k);   // warp-broadcast
                float d;
                d = av.x - bv.x; d2[c] = fmaf(d, d, d2[c]);
                d = av.y - bv.y; d2[c] = fmaf(d, d, d2[c]);
                d = av.z - bv.z; d2[c] = fmaf(d, d, d2[c]);
                d = av.w - bv.w; d2[c] = fmaf(d, d, d2[c]);
            }
        }
        bool any = false;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const int sw = cg * 16 + c;                  // swept item inside the tile
            const int irow = kOwnerIsRow ? r : sw;       // tile-local index of the ROW (X) side of the pair
            const bool valid = (a0 + r < nOwner) && (s0 + sw < nSwept);
            const float d = sqrtf(d2[c]);
            const float ex = -a2 * (d - st_m[irow]);
            float G = 0.f;
            if (valid && d > 0.f && ex > -cut2) G = st_c[irow] * exp2f(ex) / d;
            Gs[r * (BW_T + 1) + sw] = G;
            gpart += G;
            any |= (G != 0.f);
        }
        // ---- phase 2: acc[a][ch] += sum_b G[a][b] * B[b][ch]   (skipped when the whole tile is outside the window)
        if (__syncthreads_or(any)) {
#pragma unroll 4
            for (int bb = 0; bb < BW_T; ++bb) {
                float4 bv[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int ch = 4 * lane + 128 * h;
                    bv[h] = ch < C ? *reinterpret_cast<const float4*>(Bs + bb * ld + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float g = Gs[(wid * 8 + i) * (BW_T + 1) + bb];               // warp-broadcast
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        acc[i][h][0] = fmaf(g, bv[h].x, acc[i][h][0]);
                        acc[i][h][1] = fmaf(g, bv[h].y, acc[i][h][1]);
                        acc[i][h][2] = fmaf(g, bv[h].z, acc[i][h][2]);
                        acc[i][h][3] = fmaf(g, bv[h].w, acc[i][h][3]);
                    }
                }
            }
        }
    }
    // owner item r: sum the four column-group partials in a fixed order (deterministic)
    __syncthreads();
    float* gp = Gs;                                       // reuse: [4][64]
    gp[cg * BW_T + r] = gpart;
    __syncthreads();
    if (tid < BW_T) gsum[tid] = (gp[tid] + gp[BW_T + tid]) + (gp[2 * BW_T + tid] + gp[3 * BW_T + tid]);
    __syncthreads();
    // dOwner_a = A_a * gsum_a - acc_a
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = wid * 8 + i;
        if (a0 + rr >= nOwner) continue;
        const float gs = gsum[rr];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ch = 4 * lane + 128 * h;
            if (ch < C) {
                const float4 av = *reinterpret_cast<const float4*>(As + rr * ld + ch);
                float4 o;
                o.x = fmaf(av.x, gs, -acc[i][h][0]); o.y = fmaf(av.y, gs, -acc[i][h][1]);
                o.z = fmaf(av.z, gs, -acc[i][h][2]); o.w = fmaf(av.w, gs, -acc[i][h][3]);
                *reinterpret_cast<float4*>(dOwner + ((size_t)b * nOwner + a0 + rr) * C + ch) = o;
            }
        }
    }
}

// c_i = sum_k dW_ik w_ik
__global__ void softmap_bwd_rowstat_kernel(const float* __restrict__ top_w, const float* __restrict__ dW, int rows, int topk,
                                           float* __restrict__ cvec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    float c = 0.f;
    for (int k = 0; k < topk; ++k) c = fmaf(dW[(size_t)i * topk + k], top_w[(size_t)i * topk + k], c);
    cvec[i] = c;
}

// top-k part: G_ik = -alpha w_ik dW_ik / d_ik;  dX_i += G (x_i - y_j),  dY_j -= G (x_i - y_j)   (one warp per row)
__global__ void __launch_bounds__(256)
softmap_bwd_topk_kernel(const float* __restrict__ X, const float* __restrict__ Y, int N, int M, int C, float alpha, int topk,
                        const int* __restrict__ top_idx, const float* __restrict__ top_w, const float* __restrict__ top_d,
                        const float* __restrict__ dW, int rows, float* __restrict__ dX, float* __restrict__ dY) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= rows) return;
    const int b = g / N;
    const float* xr = X + (size_t)g * C;
    float* dxr = dX + (size_t)g * C;
    for (int k = 0; k < topk; ++k) {
        const float d = top_d[(size_t)g * topk + k];
        const float G = d > 0.f ? -alpha * top_w[(size_t)g * topk + k] * dW[(size_t)g * topk + k] / d : 0.f;
        if (G == 0.f) continue;                                    // warp-uniform
        const int j = top_idx[(size_t)g * topk + k];
        const float* yr = Y + ((size_t)b * M + j) * C;
        float* dyr = dY + ((size_t)b * M + j) * C;
        for (int c = lane; c < C; c += 32) {
            const float v = G * (xr[c] - yr[c]);
            dxr[c] += v;                                           // this warp owns row g
            atomicAdd(dyr + c, -v);
        }
    }
}

int launch_softmap_bwd_topk(const float* X, const float* Y, int B, int N, int M, int C, float alpha, int topk, const int32_t* top_idx,
                            const float* top_w, const float* top_d, const float* dW, float* dX, float* dY, cudaStream_t st) {
    const int rows = B * N;
    softmap_bwd_topk_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(X, Y, N, M, C, alpha, topk, top_idx, top_w, top_d, dW, rows, dX, dY);
    DVM_LAUNCH_CHECK();
    return 0;
}

}  // namespace dvm

using namespace dvm;

extern "C" size_t dvm_softmap_bwd_workspace_bytes(int B, int N, int M, int C) {
    if (B <= 0 || N <= 0 || M <= 0 || C <= 0) return 0;
    return align_up((size_t)B * N * sizeof(float), 256);
}

extern "C" int dvm_softmap_bwd(const float* X, const float* Y, int B, int N, int M, int C, float alpha, int topk,
                               const int32_t* top_idx, const float* top_w, const float* top_d,
                               const float* row_min, const float* row_sum, const float* dW,
                               float* dX, float* dY, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    DVM_CHECK_ARG(X && Y && top_idx && top_w && top_d && row_min && row_sum && dW && dX && dY, "dvm_softmap_bwd: null pointer");
    DVM_CHECK_ARG(B > 0 && N > 0 && M > 0 && B <= 65535, "dvm_softmap_bwd: bad sizes (B=%d N=%d M=%d)", B, N, M);
    DVM_CHECK_ARG(C > 0 && C % 4 == 0 && C <= 256, "dvm_softmap_bwd: C=%d must be a multiple of 4 and <= 256", C);
    DVM_CHECK_ARG(topk >= 1 && topk <= DVM_TOPK_MAX, "dvm_softmap_bwd: bad topk %d", topk);
    DVM_CHECK_ARG(alpha >= 0.f && isfinite(alpha), "dvm_softmap_bwd: bad alpha");
    if (!ws || ws_bytes < dvm_softmap_bwd_workspace_bytes(B, N, M, C)) {
        set_error("dvm_softmap_bwd: workspace too small");
        return DVM_ERR_WORKSPACE;
    }
    float* cvec = (float*)ws;
    const int rows = B * N;
    softmap_bwd_rowstat_kernel<<<ceil_div(rows, 256), 256, 0, st>>>(top_w, dW, rows, topk, cvec);
    DVM_LAUNCH_CHECK();
    const size_t smem = ((size_t)2 * BW_T * (C + 4) + BW_T * (BW_T + 1) + 3 * BW_T) * sizeof(float);
    static PerDeviceOnce attr_done;
    if (attr_done.need()) {
        DVM_CUDA(cudaFuncSetAttribute(softmap_bwd_dense_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        DVM_CUDA(cudaFuncSetAttribute(softmap_bwd_dense_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_done.done();
    }
    const float a2 = alpha * kLog2e;
    // window: terms below exp(-cut) of the row maximum are dropped (<= M * 1e-14 of the row's gradient mass)
    const float cut2 = kExpCut * kLog2e;
    softmap_bwd_dense_kernel<true><<<dim3(ceil_div(N, BW_T), B), BW_THREADS, smem, st>>>(X, Y, N, M, C, alpha, a2, cut2, cvec, row_min, row_sum, dX);
    DVM_LAUNCH_CHECK();
    softmap_bwd_dense_kernel<false><<<dim3(ceil_div(M, BW_T), B), BW_THREADS, smem, st>>>(X, Y, N, M, C, alpha, a2, cut2, cvec, row_min, row_sum, dY);
    DVM_LAUNCH_CHECK();
    softmap_bwd_topk_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(X, Y, N, M, C, alpha, topk, top_idx, top_w, top_d, dW, rows, dX, dY);
    DVM_LAUNCH_CHECK();
    return 0;
}
