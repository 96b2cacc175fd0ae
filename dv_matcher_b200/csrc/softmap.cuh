// Internal interface between the soft-map candidate passes (SIMT fp32 / tcgen05) and finalize.
#pragma once
#include "common.cuh"

namespace dvm {

// Partial candidate lists produced by a candidate pass for `rows` rows, P lists per row.
struct CandBuffers {
    float* key;     // [rows][P][KC]  squared distances (true domain), ascending, INFINITY-padded
    int*   idx;     // [rows][P][KC]  column indices, -1 padded
    float* l;       // [rows][P]      sum over non-candidate columns of exp(-alpha (d - r))
    float* r;       // [rows][P]      reference distance of l (running min of the partial sweep)
    float* t;       // [rows][P]      discard bound: every column the partial list dropped has d^2 >= t (INFINITY: none beyond key[KC-1])
    int    P;
};

// fp32 CUDA-core candidate pass.  row_list == nullptr: all B*N rows; otherwise the global row ids
// (b*N + i) in row_list[0 .. *row_count) are processed and candidates are written at the same row id.
int launch_cand_simt(const float* X, const float* Y, int B, int N, int M, int C, float alpha, bool soft,
                     const int* row_list, const int* row_count, int max_rows,
                     CandBuffers cb, cudaStream_t st);

// tcgen05 candidate pass (softmap_tc.cu): operands converted to 16-bit, K padded to a multiple of 64 plus a
// 16-wide block that folds |y~|^2/2 into the GEMM; candidate keys come back in the true d^2 domain.
size_t tc_workspace_bytes(int B, int N, int M, int C);
// xx_out / yymax_out receive device pointers (inside ws) to |x~|^2 per row [B*N] and max |y~|^2 per batch [B].
int launch_cand_tc(const float* X, const float* Y, int B, int N, int M, int C, float alpha, bool soft, int prec,
                   CandBuffers cb, float* err_x, float* err_ymax, const float** xx_out, const float** yymax_out,
                   void* ws, size_t ws_bytes, cudaStream_t st);
int tc_num_partials(int B, int N, int M);

}  // namespace dvm
