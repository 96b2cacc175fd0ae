/*
 * dvm_b200.h -- C ABI of libdvm_b200.so: the B200 (sm_100a) hot path of DV-Matcher's dense
 * correspondence + deformation pipeline.
 *
 * Conventions (every entry point):
 *   - all data pointers are DEVICE pointers owned by the caller (PyTorch), contiguous, row-major;
 *   - `stream` is a cudaStream_t passed as void*; work is only enqueued, never synchronised;
 *   - no allocation inside: scratch comes from the caller (`ws`, sized by *_workspace_bytes);
 *   - return value: 0 on success, a positive cudaError_t, or a negative DVM_ERR_* argument code;
 *     dvm_last_error_string() describes the last failure on the calling thread;
 *   - thread-safe per stream (no global mutable state besides cached function attributes).
 *
 * Each entry point cites the reference (rqhuang88/DV-Matcher) interface it replaces.
 */
#ifndef DVM_B200_H
#define DVM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVM_VERSION 100

#define DVM_ERR_INVALID_ARG   (-1)
#define DVM_ERR_WORKSPACE     (-2)
#define DVM_ERR_UNSUPPORTED   (-3)
#define DVM_ERR_DEVICE        (-4)

/* similarity precision of dvm_softmap_fwd's candidate pass (final top-k distances are always
 * re-scored exactly in fp32 from the fp32 inputs, whatever the candidate pass used) */
#define DVM_PREC_FP32   0   /* CUDA-core fp32 direct differences (exact form, parity mode)        */
#define DVM_PREC_F16    1   /* tcgen05 kind::f16, fp16 operands, fp32 accumulation in TMEM       */
#define DVM_PREC_BF16   2   /* tcgen05 kind::f16, bf16 operands, fp32 accumulation in TMEM       */

#define DVM_MODE_HARD   0   /* arg-min / top-k only (knnsearch_t)                                */
#define DVM_MODE_SOFT   1   /* + row softmax statistics and top-k weights (knnsearch_t_grad+topk_pi) */

#define DVM_TOPK_MAX    10  /* models/loss.py:1340 hard-codes k = 10                             */
#define DVM_KNN_MAX     16

int         dvm_version(void);
const char* dvm_last_error_string(void);
/* 0 if the current device is compute capability 10.x, DVM_ERR_DEVICE otherwise */
int         dvm_device_check(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long   dvm_launch_count(void);
/* CUDA-event brackets around every launch of the dominant kernel (the similarity candidate pass of
 * dvm_softmap_fwd) on its own stream: enable(1) resets and starts recording, read() synchronises the
 * recorded events and returns the summed duration and the number of launches bracketed. */
int         dvm_profile_enable(int on);
int         dvm_profile_read(double* total_ms, int* brackets);
/* same for a named channel: 0 = the candidate pass (priming + sweep), 1 = the whole fused op dvm_softmap_fwd
 * (operand preparation + priming + sweep + finalize + rescue), the span north_star's tensor-peak target is quoted on */
int         dvm_profile_read_channel(int channel, double* total_ms, int* brackets);

/* ------------------------------------------------------------------------------------------
 * Fused similarity -> row softmax -> top-k soft map -> Pi.V -> arg-min, never materialising N x M.
 * Replaces, in one call:  knnsearch_t        models/loss.py:91-95   (argmin, int64, 0-based)
 *                         knnsearch_t_grad   models/loss.py:110-114 (softmax(-alpha*cdist))
 *                         topk_pi            models/loss.py:1339-1347 (top-10, un-renormalised)
 *                         Pi_12 @ verts2     models/loss.py:1408-1409
 * X[B,N,C], Y[B,M,C], V[B,M,Dv] (may be NULL with Dv=0), fp32.  C % 4 == 0, C <= 256, topk <= 10 <= M.
 * Outputs (any may be NULL except top_idx/top_d):
 *   argmin  int64[B,N]      top_idx int32[B,N,topk] (ascending distance; ties -> lower index)
 *   top_w   f32[B,N,topk]   top_d   f32[B,N,topk] (exact fp32 Euclidean distances)
 *   row_min f32[B,N] (= d of the nearest column)   row_sum f32[B,N] (= sum_j exp(-alpha (d_j - row_min)))
 *   PiV     f32[B,N,Dv]
 *   stats   int32[4]: [0] rows whose top-k the 16-bit candidate pass could not certify (settled exactly by the
 *                     rescue scan), [1] exact fp32 near-ties at rank topk / topk+1, [2] rows that needed the full
 *                     fp32 candidate pass, [3] rows whose 16-bit softmax mass was non-finite (counted in [0])
 * ------------------------------------------------------------------------------------------ */
size_t dvm_softmap_workspace_bytes(int B, int N, int M, int C, int prec);
int dvm_softmap_fwd(const float* X, const float* Y, const float* V,
                    int B, int N, int M, int C, int Dv, float alpha, int topk, int mode, int prec,
                    int64_t* argmin, int32_t* top_idx, float* top_w, float* top_d,
                    float* row_min, float* row_sum, float* PiV, int32_t* stats,
                    void* ws, size_t ws_bytes, void* stream);

/* Backward of the soft map w.r.t. the features (autograd of models/loss.py:110-114 + 1339-1347).
 * dW[B,N,topk] is the gradient w.r.t. the kept (un-renormalised) weights; produces dX[B,N,C] and
 * ACCUMULATES into dY[B,M,C] (caller zeroes it).  Uses the saved top_idx/top_w/top_d/row_min/row_sum. */
size_t dvm_softmap_bwd_workspace_bytes(int B, int N, int M, int C);
int dvm_softmap_bwd(const float* X, const float* Y, int B, int N, int M, int C, float alpha, int topk,
                    const int32_t* top_idx, const float* top_w, const float* top_d,
                    const float* row_min, const float* row_sum, const float* dW,
                    float* dX, float* dY, void* ws, size_t ws_bytes, void* stream);

/* The same backward with the dense part on tensor cores (tcgen05, f16 operands, fp32 accumulation in TMEM): per tile
 * S = X~ Y~^T -> G = alpha c_i P_ij / d_ij (f16, written to shared memory as the next MMA's operand) -> dX += G Y~, dY += G^T X~;
 * the ten kept entries of every row (most of its mass) are then treated EXACTLY in fp32: their exact dense + top-k gradient is added
 * and what the tensor-core passes contributed for them is removed.  C <= 128.  dX, dY are OVERWRITTEN.  Stated bound on the
 * gradient, relative to its largest entry: 5e-3 for alpha <= 60, 5e-2 with cosine >= 0.999 at alpha = 100 (measured 7e-4 / 2.9e-2,
 * cosine 0.9997; the fp32 version above keeps 2e-4). */
size_t dvm_softmap_bwd_tc_workspace_bytes(int B, int N, int M, int C);
int dvm_softmap_bwd_tc(const float* X, const float* Y, int B, int N, int M, int C, float alpha, int topk,
                       const int32_t* top_idx, const float* top_w, const float* top_d,
                       const float* row_min, const float* row_sum, const float* dW,
                       float* dX, float* dY, void* ws, size_t ws_bytes, void* stream);

/* 10-sparse transfers Pi @ Y for any row width D: torch.matmul(Pi_12, feat2) models/model.py:471,
 * einsum('bij,bjkm->bikm') models/loss.py:1237 (with Y viewed as [B,M,k*m]).
 * out[b,i,:] = sum_k w[b,i,k] * Y[b, idx[b,i,k], :].   bwd: dW = <dOut, Y[idx]>, dY += w * dOut.
 * K <= DVM_KNN_MAX; the forward also takes K <= 1024 when D <= 3 (the top-40 reconstructions of test_partial.py:73-96). */
int dvm_sparse_transfer_fwd(const int32_t* idx, const float* w, const float* Y,
                            int B, int N, int M, int K, int D, float* out, void* stream);
int dvm_sparse_transfer_bwd(const int32_t* idx, const float* w, const float* Y, const float* dOut,
                            int B, int N, int M, int K, int D, float* dW, float* dY, void* stream);

/* ------------------------------------------------------------------------------------------
 * Brute-force k-NN on 3-D points with the exact direct-difference squared distance
 * ((dx*dx + dy*dy) + dz*dz, no FMA contraction), ascending, ties -> lower index.
 * Replaces knn_grad(x,y,k) models/loss.py:97-101 / deform.py:24-28 (k=10 on xyz) and the two SciPy
 * KD-tree queries of lib/deformation_graph_point.py:181-191 (use_f64=1 evaluates in fp64 like SciPy).
 * Q[B,N,3], R[B,M,3] -> idx int64[B,N,k] (or idx32 int32, either may be NULL), d2 f32[B,N,k] (NULL ok;
 * with use_f64 the squared distance is rounded to fp32 on output, d2_f64 receives the fp64 value).
 * ws (dvm_knn3_workspace_bytes) enables the uniform-grid search (O(N) work, bit-identical results); with
 * ws == NULL, or for reference clouds below 1024 points, the shared-memory-tiled brute-force sweep runs.
 * ------------------------------------------------------------------------------------------ */
size_t dvm_knn3_workspace_bytes(int B, int N, int M);
int dvm_knn3(const float* Q, const float* R, int B, int N, int M, int k, int use_f64,
             int64_t* idx, int32_t* idx32, float* d2, double* d2_f64, void* ws, size_t ws_bytes, void* stream);

/* chamfer_3DDist forward/backward (ThibaultGROUEIX/ChamferDistancePytorch chamfer3D, call sites
 * models/loss.py:1099,1223 and 750,874): squared distances, int32 arg-mins, both directions. */
size_t dvm_chamfer_workspace_bytes(int B, int N, int M);
int dvm_chamfer_fwd(const float* a, const float* b, int B, int N, int M,
                    float* dist1, float* dist2, int32_t* idx1, int32_t* idx2, void* ws, size_t ws_bytes, void* stream);
/* da[B,N,3], db[B,M,3] are OVERWRITTEN. */
int dvm_chamfer_bwd(const float* a, const float* b, const int32_t* idx1, const int32_t* idx2,
                    const float* g1, const float* g2, int B, int N, int M,
                    float* da, float* db, void* stream);

/* ------------------------------------------------------------------------------------------
 * Deformation graph.
 * dvm_fps: farthest_point_sample lib/deformation_graph_point.py:18-33 with the random start index
 *   injected (start[B] int64): distance init 1e10, masked min update, first-index arg-max, unfused
 *   fp32 arithmetic -> node lists bit-identical to the reference.  xyz[B,N,3] -> out int64[B,K].
 * ------------------------------------------------------------------------------------------ */
size_t dvm_fps_workspace_bytes(int B, int N);
int dvm_fps(const float* xyz, int B, int N, int K, const int64_t* start, int64_t* out,
            void* ws, size_t ws_bytes, void* stream);

/* construct_graph_euclidean lib/deformation_graph_point.py:177-201 minus node selection:
 * given nodes_idx[B,K] (vertex indices) computes
 *   influence int64[B,N,3]  3 nearest nodes per vertex (node index space), dists f32[B,N,3],
 *   weights   f32[B,N,3]    exp(-d^2 / 2 sigma^2) row-normalised,
 *   ring      int64[B,K,9]  9-NN node ring incl. self (fp64 evaluation, as SciPy),
 *   sigma     f64[B]        20 * mean distance to the nearest other vertex. */
size_t dvm_graph_workspace_bytes(int B, int N, int K);
int dvm_graph_weights(const float* xyz, const int64_t* nodes_idx, int B, int N, int K,
                      int64_t* influence, float* dists, float* weights, int64_t* ring, double* sigma,
                      void* ws, size_t ws_bytes, void* stream);

/* rotation_6d_to_matrix models/loss.py:39-45: d6[n,6] -> R[n,9] (rows b1,b2,b3); bwd: dR -> dd6. */
int dvm_rot6d_fwd(const float* d6, int n, float* R, void* stream);
int dvm_rot6d_bwd(const float* d6, const float* dR, int n, float* dd6, void* stream);

/* DeformationGraph_geod.forward lib/deformation_graph_point.py:233-261, batched over B clouds:
 * skinning warp  out[b,i] = sum_k w[b,i,k] (R_n (v_i - g_n) + g_n + t_n),  n = influence[b,i,k],
 * g = xyz[nodes_idx].  R[B,K,9], t[B,K,3].  bwd: dOut[B,N,3] -> dR[B,K,9], dt[B,K,3] (OVERWRITTEN). */
int dvm_skin_fwd(const float* xyz, const int64_t* nodes_idx, const int64_t* influence, const float* weights,
                 const float* R, const float* t, int B, int N, int K, float* out, void* stream);
int dvm_skin_bwd(const float* xyz, const int64_t* nodes_idx, const int64_t* influence, const float* weights,
                 const float* dOut, int B, int N, int K, float* dR, float* dt, void* stream);
/* ARAP + rotation smoothness of the same forward: arap[b] = sum_i sum_{j in ring(i)} ||(g_i+t_i) -
 * (g_j+t_j) - R_i (g_i - g_j)||^2 / K,  sr[b] = mean (R_i - R_j)^2.  Deterministic two-stage reduction.
 * bwd ACCUMULATES g_arap[b] * d arap / d(R,t) into dR, dt. */
size_t dvm_arap_workspace_bytes(int B, int K);
int dvm_arap_fwd(const float* xyz, const int64_t* nodes_idx, const int64_t* ring, const float* R, const float* t,
                 int B, int N, int K, int ring_k, float* arap, float* sr, void* ws, size_t ws_bytes, void* stream);
int dvm_arap_bwd(const float* xyz, const int64_t* nodes_idx, const int64_t* ring, const float* R, const float* t,
                 const float* g_arap, int B, int N, int K, int ring_k, float* dR, float* dt, void* stream);

/* ------------------------------------------------------------------------------------------
 * HBM-shaped layout of the same forward (lib/deformation_graph_point.py:233-261): the product path.  The graph is
 * static per shape, so the host lays it out once (deformation_graph.pack_graph).  NODES ARE RENUMBERED in Morton order of
 * their positions: node_perm int32[B][K], node_perm[new] = old (the reference's FPS order).
 *   table   f32[B][K][16]   node records in the NEW order, 64-byte aligned: R0..R8, t0..t2, g0..g2, 0   (g = node position)
 *   vorder  int32[B][N]     vertices in Morton order;  s_xyz f32[B][N][3] their coordinates;  s_infl int32[B][3][N] (new node
 *                           numbers), s_w f32[B][3][N]: influence lists of vertex vorder[i], slot-major;
 *   s_ring  int32[B][ring_k][K]: ring of new node i (new numbers), slot-major;
 *   csr_ptr int32[B][K+1], csr_vert int32[B][3N], csr_w f32[B][3N]: vertices influenced by each new node (ascending vertex id).
 * dvm_node_table packs (R, t, g): R[B][K][9], t[B][K][3] rows are read through node_perm when it is non-NULL (inputs in the
 * reference's node order), directly otherwise (inputs already in the new order); nodes_xyz[B][K][3] is in the new order.
 * dvm_node_table_from_d9 fuses models/loss.py:1257-1264 + rotation_6d_to_matrix (39-45): d9[B][K][9] = Deformer output,
 *   t = d9[0:3], R = rot6d(d9[3:9] + [1,0,0,0,1,0]); R_out/t_out optional plain copies (new order).
 * dvm_skin_fwd_packed == dvm_skin_fwd (out[B][N][3] in the ORIGINAL vertex order), dvm_arap_fwd_packed == dvm_arap_fwd (sr may
 * be NULL: smoothness skipped), dvm_skin_bwd_csr == dvm_skin_bwd without atomics (deterministic; dR, dt OVERWRITTEN, new order).
 * ------------------------------------------------------------------------------------------ */
int dvm_node_table(const float* R, const float* t, const float* nodes_xyz, const int32_t* node_perm, int B, int K, float* table, void* stream);
int dvm_node_table_from_d9(const float* d9, const float* nodes_xyz, const int32_t* node_perm, int B, int K, float* table,
                           float* R_out, float* t_out, void* stream);
int dvm_skin_fwd_packed(const float* s_xyz, const int32_t* vorder, const int32_t* s_infl, const float* s_w,
                        const float* table, int B, int N, int K, float* out, void* stream);
int dvm_skin_bwd_csr(const float* xyz, const float* nodes_xyz, const int32_t* csr_ptr, const int32_t* csr_vert,
                     const float* csr_w, const float* dOut, int B, int N, int K, float* dR, float* dt, void* stream);
size_t dvm_arap_packed_workspace_bytes(int B, int K);
int dvm_arap_fwd_packed(const int32_t* s_ring, const float* table, int B, int K, int ring_k,
                        float* arap, float* sr, void* ws, size_t ws_bytes, void* stream);

/* index_points + Conv2d(k->1, 1x1) over the neighbour axis, fused (models/loss.py:1252-1253 feeding
 * models/model.py:468-469):  out[b,r,c] = bias + sum_s W[s] * feat[b, idx[b,r,s], c]  for feat[B,N,C],
 * idx[B,R,k] (R output rows per cloud: all N vertices, or only the K graph nodes).
 * bwd: dFeat (ACCUMULATED), dW[k], dBias[1] (ACCUMULATED). */
int dvm_gather_conv_fwd(const float* feat, const int64_t* idx, const float* W, const float* bias,
                        int B, int N, int R, int C, int k, float* out, void* stream);
int dvm_gather_conv_bwd(const float* feat, const int64_t* idx, const float* W, const float* dOut,
                        int B, int N, int R, int C, int k, float* dFeat, float* dW, float* dBias, void* stream);

/* ------------------------------------------------------------------------------------------
 * Large-k feature-space k-NN, the gathered distances of the dist loss and index_points.
 * dvm_topk_select: per row of scores[rows][pitch] (N valid columns) the indices of the k largest scores, descending score,
 *   ties -> lower index, k <= 1024.  Second half of knn(a, b, k) models/loss.py:451-462 for k = 500 / 300 (:1367, :1380): the
 *   scores 2 a.b - |b|^2 come from dvm_linear_act_fwd (x = a, W = b, bias = -|b|^2 / 2; tcgen05, fp32-equivalent).
 * dvm_pair_dist_fwd: d[b,s,t] = | feat[b, nbr[b,s,t]] - feat[b, qidx[s]] |_2  == torch.norm(index_points(feat, idx) -
 *   f1[:, :, None, :], dim=-1) of models/loss.py:1368-1369 without the [B,S,k,C] tensor; geo (may be NULL; fp32 or fp64
 *   [B][N][N], geo_is_f64) additionally gathers geo_out[b,s,t] = geo[b, nbr[b,s,t], qidx[s]] (:1370-1378, the Python loop over B).
 * dvm_pair_dist_bwd: ACCUMULATES d/dfeat of sum g*d into dfeat[B][N][C] (vector atomics; 0 where d = 0 like torch.norm).
 * dvm_gather_rows_fwd/bwd: index_points models/loss.py:464-473: out[b,r,:] = pts[b, idx[b,r], :] (idx flattened to [B][R]);
 *   bwd ACCUMULATES into d_pts.
 * ------------------------------------------------------------------------------------------ */
int dvm_topk_select(const float* scores, long long rows, int N, long long pitch, int k, int64_t* idx, void* stream);
int dvm_pair_dist_fwd(const float* feat, const int64_t* qidx, const int64_t* nbr, int B, int N, int C, int S, int k,
                      float* d, const void* geo, int geo_is_f64, float* geo_out, void* stream);
int dvm_pair_dist_bwd(const float* feat, const int64_t* qidx, const int64_t* nbr, const float* d, const float* g,
                      int B, int N, int C, int S, int k, float* dfeat, void* stream);
int dvm_gather_rows_fwd(const float* pts, const int64_t* idx, int B, int N, int R, int C, float* out, void* stream);
int dvm_gather_rows_bwd(const float* d_out, const int64_t* idx, int B, int N, int R, int C, float* d_pts, void* stream);

/* Row softmax of a score chunk E[rows][pitch] (N valid columns), written transposed: Pt[j * pt_pitch + i] = softmax_i(E[i, :])[j].
 * Middle step of LG-Net's SA_Layer attention without the N x N matrix (models/model.py:113-119; dv_matcher_b200/lgnet.py):
 * scores and the weighted sum on either side are dvm_linear_act_fwd GEMMs.  stats: scratch of 2 * rows floats (row max, 1 / row sum). */
int dvm_softmax_rows_transposed(const float* E, int rows, int N, long long pitch, float* Pt, long long pt_pitch, float* stats, void* stream);
/* Backward of that attention, per row chunk (csrc/attention.cu has the derivation): dvm_softmax_rows_inplace turns an energy chunk
 * into P (rows <= 65535); dvm_attn_softmax_bwd turns (P, dA) into dE = P (t (dA - w) - rowdot), in place over dA and transposed
 * into dEt[N][dEt_pitch].  t[N] = 1 / (1e-9 + column sums), w[N] = sum_c G[c,:] x_r[c,:]; rowdot: scratch of `rows` floats. */
int dvm_softmax_rows_inplace(float* E, int rows, int N, long long pitch, float* stats, void* stream);
int dvm_attn_softmax_bwd(const float* P, float* dA_dE, const float* t, const float* w, int rows, int N, long long pitch,
                         float* dEt, long long dEt_pitch, float* rowdot, void* stream);

/* One layer of the Deformer's decoder MLP (models/model.py:433-452: nn.Linear + nn.ELU; called at :476-477):
 *   out[r, n] = act( sum_k x[r,k] W[n,k] + bias[n] ),  x[rows][x_pitch] (K used), W[N][w_pitch] (nn.Linear layout),
 *   act 0 = identity, 1 = ELU(alpha=1).  tcgen05 tensor cores with 3xTF32 operand splitting (fp32-equivalent: relative
 *   error ~1e-6 of |x||W| per output, like an fp32 SGEMM).  x_pitch, w_pitch: multiples of 4 floats; pointers 16-byte
 *   aligned; bias may be NULL.  Forward only (training keeps torch's autograd Linear). */
int dvm_linear_act_fwd(const float* x, long long rows, int K, int x_pitch, const float* W, int w_pitch, const float* bias,
                       int N, int act, float* out, int out_pitch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DVM_B200_H */
