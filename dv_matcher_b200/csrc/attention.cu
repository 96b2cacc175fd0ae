// Row softmax of a score chunk, written TRANSPOSED (LG-Net's SA_Layer global attention,
// models/model.py:113-119: energy = x_q x_k; attention = softmax(energy, -1); attention /= 1e-9 + attention.sum(1); x_r = x_v attention).
//
// The N x N attention matrix never exists: dv_matcher_b200/lgnet.py walks row chunks.  For a chunk of `rows` query rows
//   E  [rows][N]  = Q_chunk K^T                        (dvm_linear_act_fwd: tcgen05, 3xTF32, fp32-equivalent)
//   Pt [N][rows]  = softmax over each row of E, transposed                                    (this kernel)
//   U  [N][C+1]  += Pt [x_v_chunk | 1]^T                (dvm_linear_act_fwd again: the appended row of ones makes the column sums
//                                                        of the attention matrix fall out of the same GEMM, deterministically)
// and x_r[c][j] = U[j][c] / (1e-9 + U[j][C]).  The transposed store is what lets the second GEMM read P with the contraction
// index (the chunk's rows) contiguous, i.e. K-major like every other operand of that kernel.
//
// Two launches: (1) one warp per row reduces (max, sum of exp) with coalesced reads; (2) a 2-D grid of (32 rows x 256 columns)
// blocks: every warp loads 4 rows x 32 columns, the exponentials go through a padded shared-memory tile and leave as 128-byte row
// segments of Pt.  E is read three times (the last two from L2), Pt written once.
#include "common.cuh"

namespace dvm {

constexpr int SM_ROWS = 32;
constexpr int SM_COLS = 256;
constexpr int SM_THREADS = 256;

__global__ void __launch_bounds__(256)
softmax_row_stats_kernel(const float* __restrict__ E, int rows, int N, long long pitch, float* __restrict__ stats /* [rows][2] */) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float* e = E + (size_t)r * pitch;
    float m = -INFINITY, z = 0.f;                                       // exact max, then sum of exp(e - max): torch.softmax's form
    for (int j = lane; j < N; j += 32) m = fmaxf(m, __ldg(e + j));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    for (int j = lane; j < N; j += 32) z += expf(__ldg(e + j) - m);
    z = warp_sum(z);
    if (lane == 0) { stats[2 * r] = m; stats[2 * r + 1] = z > 0.f ? 1.f / z : 0.f; }
}

__global__ void __launch_bounds__(SM_THREADS)
softmax_rows_transposed_kernel(const float* __restrict__ E, int rows, int N, long long pitch, const float* __restrict__ stats,
                               float* __restrict__ Pt, long long pt_pitch) {
    __shared__ float s_m[SM_ROWS], s_iz[SM_ROWS];
    __shared__ float tile[SM_ROWS][33];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int r0 = blockIdx.x * SM_ROWS;
    const int c_lo = blockIdx.y * SM_COLS, c_hi = min(N, c_lo + SM_COLS);
    if (threadIdx.x < SM_ROWS) {
        const int r = r0 + threadIdx.x;
        s_m[threadIdx.x] = r < rows ? stats[2 * r] : 0.f;
        s_iz[threadIdx.x] = r < rows ? stats[2 * r + 1] : 0.f;
    }
    __syncthreads();
    for (int j0 = c_lo; j0 < c_hi; j0 += 32) {
#pragma unroll
        for (int q = 0; q < SM_ROWS / 8; ++q) {
            const int rl = wid * (SM_ROWS / 8) + q, r = r0 + rl, j = j0 + lane;
            float p = 0.f;
            if (r < rows && j < c_hi) p = expf(__ldg(E + (size_t)r * pitch + j) - s_m[rl]) * s_iz[rl];
            tile[rl][lane] = p;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int jl = wid * 4 + q, j = j0 + jl;          // warp w writes columns 4w .. 4w+3 of the tile as rows of Pt
            if (j < c_hi && r0 + lane < rows) Pt[(size_t)j * pt_pitch + r0 + lane] = tile[lane][jl];
        }
        __syncthreads();
    }
}

// ---- backward of the same attention (training at sizes where the N x N matrices do not fit) -----------------------------------
// With P = softmax_rows(E), s_j = sum_i P_ij, t_j = 1 / (1e-9 + s_j), A_ij = P_ij t_j, x_r = x_v A and G = dL/dx_r:
//   dA_ij = sum_c x_v[c,i] G[c,j]                       (GEMM per row chunk)
//   w_j   = sum_i dA_ij A_ij = sum_c G[c,j] x_r[c,j]    (no N x N sum needed)
//   dP_ij = t_j (dA_ij - w_j)                           (through A = P t and s_j)
//   dE_ij = P_ij (dP_ij - sum_j' P_ij' dP_ij')          (row softmax)
// softmax_rows_inplace_kernel turns an energy chunk into P; attn_bwd_rowdot_kernel / attn_bwd_write_kernel turn (P, dA) into dE,
// in place over dA and transposed into dEt (the operand layout of the dK GEMM).
__global__ void __launch_bounds__(256)
softmax_rows_inplace_kernel(float* __restrict__ E, int rows, int N, long long pitch, const float* __restrict__ stats) {
    const int r = blockIdx.y;
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j < N) {
        float* e = E + (size_t)r * pitch + j;
        *e = expf(*e - __ldg(stats + 2 * r)) * __ldg(stats + 2 * r + 1);
    }
}

__global__ void __launch_bounds__(256)
attn_bwd_rowdot_kernel(const float* __restrict__ P, const float* __restrict__ dA, const float* __restrict__ t, const float* __restrict__ w,
                       int rows, int N, long long pitch, float* __restrict__ rowdot) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float* p = P + (size_t)r * pitch;
    const float* d = dA + (size_t)r * pitch;
    float acc = 0.f;
    for (int j = lane; j < N; j += 32) acc = fmaf(__ldg(p + j), __ldg(t + j) * (__ldg(d + j) - __ldg(w + j)), acc);
    acc = warp_sum(acc);
    if (lane == 0) rowdot[r] = acc;
}

__global__ void __launch_bounds__(SM_THREADS)
attn_bwd_write_kernel(const float* __restrict__ P, float* __restrict__ dA_dE, const float* __restrict__ t, const float* __restrict__ w,
                      const float* __restrict__ rowdot, int rows, int N, long long pitch, float* __restrict__ dEt, long long dEt_pitch) {
    __shared__ float s_rd[SM_ROWS];
    __shared__ float tile[SM_ROWS][33];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int r0 = blockIdx.x * SM_ROWS;
    const int c_lo = blockIdx.y * SM_COLS, c_hi = min(N, c_lo + SM_COLS);
    if (threadIdx.x < SM_ROWS) s_rd[threadIdx.x] = r0 + threadIdx.x < rows ? rowdot[r0 + threadIdx.x] : 0.f;
    __syncthreads();
    for (int j0 = c_lo; j0 < c_hi; j0 += 32) {
#pragma unroll
        for (int q = 0; q < SM_ROWS / 8; ++q) {
            const int rl = wid * (SM_ROWS / 8) + q, r = r0 + rl, j = j0 + lane;
            float v = 0.f;
            if (r < rows && j < c_hi) {
                const size_t o = (size_t)r * pitch + j;
                v = __ldg(P + o) * (__ldg(t + j) * (dA_dE[o] - __ldg(w + j)) - s_rd[rl]);
                dA_dE[o] = v;
            }
            tile[rl][lane] = v;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int jl = wid * 4 + q, j = j0 + jl;
            if (j < c_hi && r0 + lane < rows) dEt[(size_t)j * dEt_pitch + r0 + lane] = tile[lane][jl];
        }
        __syncthreads();
    }
}

}  // namespace dvm

using namespace dvm;

extern "C" int dvm_softmax_rows_transposed(const float* E, int rows, int N, long long pitch, float* Pt, long long pt_pitch,
                                           float* stats, void* stream) {
    DVM_CHECK_ARG(E && Pt && stats, "dvm_softmax_rows_transposed: null pointer");
    DVM_CHECK_ARG(rows > 0 && N > 0 && pitch >= N && pt_pitch >= rows, "dvm_softmax_rows_transposed: bad sizes");
    DVM_CHECK_ARG(ceil_div(N, SM_COLS) <= 65535, "dvm_softmax_rows_transposed: N too large");
    cudaStream_t st = (cudaStream_t)stream;
    softmax_row_stats_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(E, rows, N, pitch, stats);
    DVM_LAUNCH_CHECK();
    softmax_rows_transposed_kernel<<<dim3(ceil_div(rows, SM_ROWS), ceil_div(N, SM_COLS)), SM_THREADS, 0, st>>>(E, rows, N, pitch, stats, Pt, pt_pitch);
    DVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int dvm_softmax_rows_inplace(float* E, int rows, int N, long long pitch, float* stats, void* stream) {
    DVM_CHECK_ARG(E && stats, "dvm_softmax_rows_inplace: null pointer");
    DVM_CHECK_ARG(rows > 0 && rows <= 65535 && N > 0 && pitch >= N, "dvm_softmax_rows_inplace: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    softmax_row_stats_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(E, rows, N, pitch, stats);
    DVM_LAUNCH_CHECK();
    softmax_rows_inplace_kernel<<<dim3(ceil_div(N, 256), rows), 256, 0, st>>>(E, rows, N, pitch, stats);
    DVM_LAUNCH_CHECK();
    return 0;
}

extern "C" int dvm_attn_softmax_bwd(const float* P, float* dA_dE, const float* t, const float* w, int rows, int N, long long pitch,
                                    float* dEt, long long dEt_pitch, float* rowdot, void* stream) {
    DVM_CHECK_ARG(P && dA_dE && t && w && dEt && rowdot, "dvm_attn_softmax_bwd: null pointer");
    DVM_CHECK_ARG(rows > 0 && N > 0 && pitch >= N && dEt_pitch >= rows && ceil_div(N, SM_COLS) <= 65535, "dvm_attn_softmax_bwd: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    attn_bwd_rowdot_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(P, dA_dE, t, w, rows, N, pitch, rowdot);
    DVM_LAUNCH_CHECK();
    attn_bwd_write_kernel<<<dim3(ceil_div(rows, SM_ROWS), ceil_div(N, SM_COLS)), SM_THREADS, 0, st>>>(P, dA_dE, t, w, rowdot, rows, N, pitch, dEt, dEt_pitch);
    DVM_LAUNCH_CHECK();
    return 0;
}
