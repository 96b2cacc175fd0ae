// tcgen05 backward of the top-k soft map w.r.t. the features (autograd of models/loss.py:110-114 + 1339-1347): the dense part
//     G_ij = alpha c_i P_ij / d_ij,   dX_i = sum_j G_ij (x_i - y_j),   dY_j = sum_i G_ij (y_j - x_i)
// (softmap_bwd.cu has the derivation and the exact fp32 CUDA-core version) as a flash-attention-style pair of GEMMs per tile:
//
//   MMA1   S = A_owner . B_swept^T  on 16-bit operands with the norm folded in (the forward's operand format: the accumulator is
//          key_ij = |y~_j|^2/2 - x~_i.y~_j)                                    -> TMEM, two stages of 128 columns
//   epi    d = sqrt(2 key + |x~_i|^2),  G = coef_i exp2(-a2 (d - rmin_i)) / d   -> f16, written to shared memory in the K-major
//          128-byte-swizzled operand layout (the epilogue warps are the "producer" of MMA2's A operand)
//   MMA2   acc[owner][c] += G[owner][swept] . V[swept][c]   with V = the SAME swept tile in shared memory, read as an MN-major
//          operand (points along K, channels along N): no transposed copy of the features exists anywhere
//   final  dOwner = gsum * owner~ -+ acc  (gsum = sum of the ROUNDED G values, owner~ = the 16-bit-rounded owner row, so the
//          difference x~_i - y~_j cancels exactly as it does inside the GEMM)
//
// Two launches: owners = X rows (-> dX) and owners = Y columns (-> dY); no atomics, deterministic.  The 10-sparse top-k part
// (exact fp32) stays in softmap_bwd_topk_kernel.  Per tile the epilogue needs 2 MUFU per entry (rsqrt, ex2): 2048 clocks for
// 128 x 128 entries against 2 x 512 tensor clocks -- the kernel is MUFU-bound by construction, like the dense-window forward.
//
// Range of the 16-bit G: with owners = rows the factor coef_i = alpha c_i / Z_i is applied in the final epilogue (G' = e/d only);
// with owners = columns it varies along K and is applied per entry, scaled by 1 / max|coef| (computed on the device).
#include <cuda.h>
#include "softmap.cuh"
#include "tc_ptx.cuh"
#include "tc_prep.cuh"

namespace dvm {

constexpr int BT_M = 128;             // owners per CTA (UMMA M)
constexpr int BT_N = 128;             // swept items per tile (UMMA N of MMA1, K of MMA2)
constexpr int BT_NST = 3;             // swept ring depth
constexpr int BT_EPI_WARPS = 8;       // (TMEM lane quarter) x (column half of the S tile)
constexpr int BT_THREADS = 64 + 32 * BT_EPI_WARPS;
constexpr int BT_BLK = 128 * 128;     // one 128-row x 64-element K block (bytes)
constexpr int BT_EXT = 128 * 32;      // one 128-row x 16-element K block (bytes)

struct BtParams {
    int N, M, C, Cpad, KB;
    int nOwner, nSwept, tiles;
    float a2;
    uint32_t idesc1, idesc2;
    const float* xx;                  // [B*N] |x~_i|^2
    const float* coef;                // [B*N] alpha c_i / Z_i
    const float* rmin;                // [B*N]
    const float* coef_max;            // [1]   max |coef| (owners = columns: per-entry scaling)
    const float* own_feat;            // fp32 owner-side features [B][nOwner][C]
    float* dOwn;                      // [B][nOwner][C]
    int bf16;
};

__device__ __forceinline__ float bt_ex2(float x) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x)); return e; }
__device__ __forceinline__ float bt_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ void bt_mma(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, bool acc) {
    asm volatile(
        "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tsetp.ne.u32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"((uint32_t)acc) : "memory");
}
__device__ __forceinline__ void bt_named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <bool kOwnerIsRow>
__global__ void __launch_bounds__(BT_THREADS, 1)
softmap_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmOe,
                      const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmSe, const BtParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int unit = p.KB * BT_BLK + BT_EXT;              // one 128-row operand block, all of K
    uint8_t* Os = smem;                                   // owner rows (resident)
    uint8_t* Ss = Os + unit;                              // [NST] swept tiles
    uint8_t* Gs = Ss + BT_NST * unit;                     // [2][2 x 16 KB] G tiles: 128 owners x 128 swept, f16, K-major SW128
    float* st_coef = reinterpret_cast<float*>(Gs + 2 * 2 * BT_BLK);     // [2][128] row-side stats of the swept tile (owners = columns)
    float* st_rmin = st_coef + 2 * BT_N;
    float* st_xx = st_rmin + 2 * BT_N;
    float* s_gsum = st_xx + 2 * BT_N;                     // [2][128] per column half
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_gsum + 2 * BT_M);
    uint64_t* full = bars;                // [NST]
    uint64_t* empty = full + BT_NST;      // [NST]
    uint64_t* sfull = empty + BT_NST;     // [2]
    uint64_t* sfree = sfull + 2;          // [2]  8 epilogue warps arrive
    uint64_t* gfull = sfree + 2;          // [2]  8 epilogue warps arrive
    uint64_t* gfree = gfull + 2;          // [2]
    uint64_t* ofull = gfree + 2;          // [1]
    uint64_t* accfull = ofull + 1;        // [1]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(accfull + 1);

    const int warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int own0 = blockIdx.x * BT_M;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (int s = 0; s < BT_NST; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(sfull + s, 1); mbar_init(sfree + s, BT_EPI_WARPS); mbar_init(gfull + s, BT_EPI_WARPS); mbar_init(gfree + s, 1); }
        mbar_init(ofull, 1); mbar_init(accfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&tmO); tma_prefetch_desc(&tmOe); tma_prefetch_desc(&tmS); tma_prefetch_desc(&tmSe);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const int T = p.tiles;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            mbar_arrive_expect_tx(ofull, unit);
            for (int kb = 0; kb < p.KB; ++kb) tma_load_3d(&tmO, ofull, Os + kb * BT_BLK, kb * TC_KBLK, own0, b);
            tma_load_3d(&tmOe, ofull, Os + p.KB * BT_BLK, 0, own0, b);
            for (int t = 0; t < T; ++t) {
                const int s = t % BT_NST;
                const uint32_t ph = (t / BT_NST) & 1;
                mbar_wait(empty + s, ph ^ 1);
                mbar_arrive_expect_tx(full + s, unit);
                uint8_t* dst = Ss + s * unit;
                for (int kb = 0; kb < p.KB; ++kb) tma_load_3d(&tmS, full + s, dst + kb * BT_BLK, kb * TC_KBLK, t * BT_N, b);
                tma_load_3d(&tmSe, full + s, dst + p.KB * BT_BLK, 0, t * BT_N, b);
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            constexpr uint32_t HI128 = (1024u >> 4) | (1u << 14) | (2u << 29);     // SBO 1024 B, version 1, SWIZZLE_128B
            constexpr uint32_t HI32 = (256u >> 4) | (1u << 14) | (6u << 29);       // SBO 256 B, version 1, SWIZZLE_32B
            const uint32_t olo = ((smem_u32(Os) >> 4) & 0x3FFFu) | (1u << 16);
            const uint32_t ext = (uint32_t)(p.KB * BT_BLK) >> 4;
            const uint32_t acc_t = tmem_base + 256u;
            auto mma2 = [&](int u) {
                const int gs = u & 1, s = u % BT_NST;
                mbar_wait(gfull + gs, (u >> 1) & 1);               // the epilogue has written G(u) (and fenced it to the async proxy)
                tc_fence_after();
                const uint32_t glo = ((smem_u32(Gs + gs * 2 * BT_BLK) >> 4) & 0x3FFFu) | (1u << 16);
                // V = swept tile, MN-major: channels (N) contiguous in 64-element blocks 16 KB apart (LBO), points (K) in 8-row groups
                // 1024 B apart (SBO); 16 points per MMA = +2048 B
                const uint32_t vlo = ((smem_u32(Ss + s * unit) >> 4) & 0x3FFFu) | ((uint32_t)(BT_BLK >> 4) << 16);
#pragma unroll
                for (int k = 0; k < BT_N / 16; ++k) {
                    const uint32_t a = glo + (uint32_t)((k >> 2) * (BT_BLK >> 4) + (k & 3) * 2);      // K block of 64 swept, +32 B per step
                    bt_mma(acc_t, a, HI128, vlo + (uint32_t)k * (2048u >> 4), HI128, p.idesc2, u > 0 || k > 0);
                }
                tc_commit(gfree + gs);                             // G stage reusable
                tc_commit(empty + s);                              // swept stage reusable (MMA1 and MMA2 of this tile have read it)
            };
            mbar_wait(ofull, 0);
            for (int t = 0; t < T; ++t) {
                const int s = t % BT_NST, acc = t & 1;
                mbar_wait(sfree + acc, ((t >> 1) & 1) ^ 1);
                mbar_wait(full + s, (t / BT_NST) & 1);
                tc_fence_after();
                const uint32_t slo = ((smem_u32(Ss + s * unit) >> 4) & 0x3FFFu) | (1u << 16);
                const uint32_t d0 = tmem_base + (uint32_t)acc * BT_N;
                for (int kb = 0; kb < p.KB; ++kb)
#pragma unroll
                    for (int k = 0; k < TC_KBLK / 16; ++k)
                        bt_mma(d0, olo + (uint32_t)kb * (BT_BLK >> 4) + 2 * k, HI128, slo + (uint32_t)kb * (BT_BLK >> 4) + 2 * k, HI128, p.idesc1, kb > 0 || k > 0);
                bt_mma(d0, olo + ext, HI32, slo + ext, HI32, p.idesc1, true);       // norm block
                tc_commit(sfull + acc);
                if (t > 0) mma2(t - 1);                            // one tile behind: the epilogue of t-1 overlaps MMA1(t)
            }
            mma2(T - 1);
            tc_commit(accfull);
        }
    } else {
        // =============================== epilogue: S -> G (f16, shared memory), row sums ===============================
        const int e = warp - 2;
        const int quarter = warp & 3, half = e >> 2;
        const int r = quarter * 32 + lane;                        // owner row of this thread (TMEM lane)
        const int own = own0 + r;
        const bool own_ok = own < p.nOwner;
        float o_coef = 0.f, o_rmin = 0.f, o_xx = 0.f;
        if (kOwnerIsRow && own_ok) {
            o_coef = __ldg(p.coef + (size_t)b * p.N + own); o_rmin = __ldg(p.rmin + (size_t)b * p.N + own); o_xx = __ldg(p.xx + (size_t)b * p.N + own);
        }
        const float inv_scale = kOwnerIsRow ? 1.f : 1.f / fmaxf(__ldg(p.coef_max), 1e-30f);
        float gsum = 0.f;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + half * 64;
        for (int t = 0; t < T; ++t) {
            const int acc = t & 1;
            if (!kOwnerIsRow) {                                    // row-side stats of this swept tile
                if (e < 4) {
                    const int i = t * BT_N + r;
                    const bool ok = i < p.N;
                    st_coef[acc * BT_N + r] = ok ? __ldg(p.coef + (size_t)b * p.N + i) * inv_scale : 0.f;
                    st_rmin[acc * BT_N + r] = ok ? __ldg(p.rmin + (size_t)b * p.N + i) : 0.f;
                    st_xx[acc * BT_N + r] = ok ? __ldg(p.xx + (size_t)b * p.N + i) : 0.f;
                }
                bt_named_bar(1, 32 * BT_EPI_WARPS);
            }
            mbar_wait_backoff(sfull + acc, (t >> 1) & 1);
            tc_fence_after();
            mbar_wait_backoff(gfree + acc, ((t >> 1) & 1) ^ 1);   // MMA2 of tile t-2 has finished reading this G stage
            const uint32_t taddr = t_lane + acc * BT_N;
            uint8_t* grow = Gs + acc * 2 * BT_BLK + half * BT_BLK + (r >> 3) * 1024 + (r & 7) * 128;    // this thread's 128-byte row of its K block
            float ka[16], kb[16];
            tc_ld16_issue(taddr, ka);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float (&k)[16] = (c & 1) ? kb : ka;
                tc_ld16_wait(k);
                if (c < 3) tc_ld16_issue(taddr + (c + 1) * 16, (c & 1) ? ka : kb);
                uint32_t pk[8];
#pragma unroll
                for (int u = 0; u < 16; u += 2) {
                    float g2[2];
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        const int j = half * 64 + c * 16 + u + v;              // swept item inside the tile
                        const float key = k[u + v];
                        const float xx = kOwnerIsRow ? o_xx : st_xx[acc * BT_N + j];
                        const float rmin = kOwnerIsRow ? o_rmin : st_rmin[acc * BT_N + j];
                        const float cf = kOwnerIsRow ? 1.f : st_coef[acc * BT_N + j];
                        const float d2 = fmaf(2.f, key, xx);
                        // padding / out-of-range items, and pairs closer than the 16-bit operands resolve (cdist's backward is 0 at d = 0)
                        const bool ok = key < 1e30f && (t * BT_N + j) < p.nSwept && d2 > 1e-6f * xx;
                        const float d2c = fmaxf(d2, 1e-30f);
                        const float rinv = bt_rsqrt(d2c);
                        const float g = cf * bt_ex2(-p.a2 * fmaf(d2c, rinv, -rmin)) * rinv;
                        g2[v] = ok ? g : 0.f;
                    }
                    const __half2 h2 = __floats2half2_rn(g2[0], g2[1]);
                    gsum += __low2float(h2) + __high2float(h2);                // sum of the ROUNDED values: what the GEMM sees
                    pk[u >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
                }
                // two 16-byte chunks (8 values each) of the row: logical chunk 2c, 2c+1 -> physical chunk ^ (row & 7)
                const int c0 = (2 * c) ^ (r & 7), c1 = (2 * c + 1) ^ (r & 7);
                *reinterpret_cast<uint4*>(grow + c0 * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(grow + c1 * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes of G -> visible to the tensor core
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(gfull + acc); mbar_arrive(sfree + acc); }
        }
        // =============================== final: dOwner = gsum * owner~ -+ acc ===============================
        s_gsum[half * BT_M + r] = gsum;
        bt_named_bar(1, 32 * BT_EPI_WARPS);
        const float gs = s_gsum[r] + s_gsum[BT_M + r];
        mbar_wait_backoff(accfull, 0);
        tc_fence_after();
        const float scale = kOwnerIsRow ? o_coef : __ldg(p.coef_max);
        const int cw = p.Cpad / 2;                                 // channels per column half
        const uint32_t aaddr = tmem_base + 256u + ((uint32_t)(quarter * 32) << 16) + half * cw;
        for (int c0 = 0; c0 < cw; c0 += 16) {
            float v[16];
            tc_ld16_issue(aaddr + c0, v);
            tc_ld16_wait(v);
            if (own_ok) {
#pragma unroll
                for (int u = 0; u < 16; u += 4) {
                    const int ch = half * cw + c0 + u;
                    if (ch < p.C) {
                        const float4 f = __ldg(reinterpret_cast<const float4*>(p.own_feat + ((size_t)b * p.nOwner + own) * p.C + ch));
                        const float fr[4] = {__half2float(__float2half_rn(f.x)), __half2float(__float2half_rn(f.y)),
                                             __half2float(__float2half_rn(f.z)), __half2float(__float2half_rn(f.w))};
                        float4 o;
                        if (kOwnerIsRow) {     // acc = sum_j G'(-y~_j): dX = coef (gsum x~ + acc)
                            o.x = scale * fmaf(gs, fr[0], v[u]); o.y = scale * fmaf(gs, fr[1], v[u + 1]);
                            o.z = scale * fmaf(gs, fr[2], v[u + 2]); o.w = scale * fmaf(gs, fr[3], v[u + 3]);
                        } else {               // acc = sum_i G x~_i: dY = s (gsum y~ - acc)
                            o.x = scale * fmaf(gs, fr[0], -v[u]); o.y = scale * fmaf(gs, fr[1], -v[u + 1]);
                            o.z = scale * fmaf(gs, fr[2], -v[u + 2]); o.w = scale * fmaf(gs, fr[3], -v[u + 3]);
                        }
                        *reinterpret_cast<float4*>(p.dOwn + ((size_t)b * p.nOwner + own) * p.C + ch) = o;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// coef_i = alpha c_i / Z_i,  c_i = sum_k dW_ik w_ik;  max |coef| for the per-entry scaling of the column pass
__global__ void softmap_bwd_tc_rowstat_kernel(const float* __restrict__ top_w, const float* __restrict__ dW, const float* __restrict__ rsum,
                                              int rows, int topk, float alpha, float* __restrict__ coef, float* __restrict__ coef_max) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float cf = 0.f;
    if (i < rows) {
        float c = 0.f;
        for (int k = 0; k < topk; ++k) c = fmaf(dW[(size_t)i * topk + k], top_w[(size_t)i * topk + k], c);
        cf = alpha * c / rsum[i];
        coef[i] = cf;
    }
    float m = fabsf(cf);
    if (!(m < INFINITY)) m = 0.f;
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(coef_max), __float_as_int(m));
}

// Top-k entries, exactly.  The dense tensor-core passes carry the 16-bit rounding of the operands into every P_ij (relative
// error alpha * delta_d: percent level at alpha = 100) -- harmless for the long tail, not for the ten largest terms of a row,
// which hold most of its mass.  For those entries this kernel (one warp per row, like the fp32 top-k pass it replaces)
//   * adds the EXACT dense + top-k gradient  G_e = (alpha c_i - alpha dW_ik) w_ik / d_ik  with fp32 direct differences, and
//   * removes what the tensor-core passes added for the same entry: G~ recomputed from the 16-bit-rounded operands with the same
//     formula, rounding and MUFU instructions (the only difference is the summation order of the 128 products: ~1e-3 of the
//     term at alpha = 100), applied to the rounded difference x~_i - y~_j exactly as the GEMMs applied it.
__global__ void __launch_bounds__(256)
softmap_bwd_tc_topk_kernel(const float* __restrict__ X, const float* __restrict__ Y, int N, int M, int C, float alpha, float a2, int topk,
                           const int* __restrict__ top_idx, const float* __restrict__ top_w, const float* __restrict__ top_d,
                           const float* __restrict__ dW, const float* __restrict__ rmin, const float* __restrict__ rsum,
                           const float* __restrict__ coef, const float* __restrict__ coef_max, const float* __restrict__ xx,
                           int rows, float* __restrict__ dX, float* __restrict__ dY) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= rows) return;
    const int b = g / N;
    const float cf = coef[g], zi = rsum[g], rm = rmin[g], xxi = xx[g];
    const float s = fmaxf(*coef_max, 1e-30f), cfs = cf / s;
    const float* xr = X + (size_t)g * C;
    float* dxr = dX + (size_t)g * C;
    float xv[4], xt[4];                                            // channels lane, lane + 32, ... (C <= 128)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c = lane + 32 * q;
        xv[q] = c < C ? xr[c] : 0.f;
        xt[q] = __half2float(__float2half_rn(xv[q]));
    }
    for (int k = 0; k < topk; ++k) {
        const int j = top_idx[(size_t)g * topk + k];
        const float d = top_d[(size_t)g * topk + k], w = top_w[(size_t)g * topk + k];
        const float* yr = Y + ((size_t)b * M + j) * C;
        float* dyr = dY + ((size_t)b * M + j) * C;
        float yv[4], yt[4], dot = 0.f, yy = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = lane + 32 * q;
            yv[q] = c < C ? yr[c] : 0.f;
            yt[q] = __half2float(__float2half_rn(yv[q]));
            dot = fmaf(xt[q], yt[q], dot); yy = fmaf(yt[q], yt[q], yy);
        }
        dot = warp_sum(dot); yy = warp_sum(yy);
        // what the tensor-core passes added for this entry
        const float d2 = fmaf(-2.f, dot, yy) + xxi;
        const bool ok = d2 > 1e-6f * xxi;
        const float d2c = fmaxf(d2, 1e-30f);
        const float rinv = bt_rsqrt(d2c);
        const float ge = bt_ex2(-a2 * fmaf(d2c, rinv, -rm)) * rinv;
        const float gA = ok ? cf * __half2float(__float2half_rn(ge)) : 0.f;
        const float gB = ok ? s * __half2float(__float2half_rn(cfs * ge)) : 0.f;
        // the exact value
        const float gE = d > 0.f ? (cf * zi - alpha * dW[(size_t)g * topk + k]) * w / d : 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = lane + 32 * q;
            if (c < C) {
                const float ex = xv[q] - yv[q], ap = xt[q] - yt[q];
                dxr[c] += gE * ex - gA * ap;                       // this warp owns row g
                atomicAdd(dyr + c, gB * ap - gE * ex);
            }
        }
    }
}

struct BtWs { uint16_t* Xh; uint16_t* Yh; float* xx; float* coef; float* coef_max; float* scratch; int Cpad, Ktot, Npad, Mpad; };

static size_t bt_ws_layout(void* base, size_t cap, int B, int N, int M, int C, BtWs* out) {
    BtWs w{};
    w.Cpad = ceil_div(C, TC_KBLK) * TC_KBLK;
    w.Ktot = w.Cpad + TC_KEXT;
    w.Npad = N; w.Mpad = ceil_div(M, BT_N) * BT_N;
    WsCarver ws(base, cap);
    w.Xh = ws.take<uint16_t>((size_t)B * N * w.Ktot);
    w.Yh = ws.take<uint16_t>((size_t)B * w.Mpad * w.Ktot);
    w.xx = ws.take<float>((size_t)B * N);
    w.coef = ws.take<float>((size_t)B * N);
    w.coef_max = ws.take<float>(64);
    w.scratch = ws.take<float>((size_t)B * N + 2 * B + 64);      // prep by-products nobody reads here (rounding errors, maxima)
    if (out) *out = w;
    return align_up(ws.off, 256);
}

}  // namespace dvm

using namespace dvm;

extern "C" size_t dvm_softmap_bwd_tc_workspace_bytes(int B, int N, int M, int C) {
    if (B <= 0 || N <= 0 || M <= 0 || C <= 0) return 0;
    return bt_ws_layout(nullptr, 0, B, N, M, C, nullptr);
}

extern "C" int dvm_softmap_bwd_tc(const float* X, const float* Y, int B, int N, int M, int C, float alpha, int topk,
                                  const int32_t* top_idx, const float* top_w, const float* top_d,
                                  const float* row_min, const float* row_sum, const float* dW,
                                  float* dX, float* dY, void* wsp, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    DVM_CHECK_ARG(X && Y && top_idx && top_w && top_d && row_min && row_sum && dW && dX && dY, "dvm_softmap_bwd_tc: null pointer");
    DVM_CHECK_ARG(B > 0 && N > 0 && M > 0 && B <= 65535, "dvm_softmap_bwd_tc: bad sizes (B=%d N=%d M=%d)", B, N, M);
    DVM_CHECK_ARG(C > 0 && C % 4 == 0 && C <= 128, "dvm_softmap_bwd_tc: C=%d must be a multiple of 4 and <= 128 (wider features: dvm_softmap_bwd)", C);
    DVM_CHECK_ARG(topk >= 1 && topk <= DVM_TOPK_MAX, "dvm_softmap_bwd_tc: bad topk %d", topk);
    DVM_CHECK_ARG(alpha >= 0.f && isfinite(alpha), "dvm_softmap_bwd_tc: bad alpha");
    BtWs w;
    const size_t need = bt_ws_layout(wsp, ws_bytes, B, N, M, C, &w);
    if (!wsp || need > ws_bytes) { set_error("dvm_softmap_bwd_tc: workspace too small"); return DVM_ERR_WORKSPACE; }
    const int rows = B * N;
    DVM_CUDA(cudaMemsetAsync(w.coef_max, 0, 64 * sizeof(float), st));
    DVM_CUDA(cudaMemsetAsync(w.scratch, 0, ((size_t)B * N + 2 * B + 64) * sizeof(float), st));
    softmap_bwd_tc_rowstat_kernel<<<ceil_div(rows, 256), 256, 0, st>>>(top_w, dW, row_sum, rows, topk, alpha, w.coef, w.coef_max);
    DVM_LAUNCH_CHECK();
    {
        dim3 gx(ceil_div(N, TC_PREP_ROWS), B), gy(ceil_div(w.Mpad, TC_PREP_ROWS), B);
        tc_prep_kernel<false, false><<<gx, 256, 0, st>>>(X, N, N, C, w.Cpad, w.Xh, w.xx, w.scratch, nullptr, nullptr);
        DVM_LAUNCH_CHECK();
        tc_prep_kernel<false, true><<<gy, 256, 0, st>>>(Y, M, w.Mpad, C, w.Cpad, w.Yh, nullptr, nullptr, w.scratch + (size_t)B * N, w.scratch + (size_t)B * N + B);
        DVM_LAUNCH_CHECK();
    }
    CUtensorMap tmX, tmXe, tmY, tmYe;
    int rc;
    if ((rc = make_operand_map(&tmX, w.Xh, false, B, N, w.Ktot, 0, w.Cpad, TC_KBLK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_operand_map(&tmXe, w.Xh, false, B, N, w.Ktot, w.Cpad, TC_KEXT, TC_KEXT, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
    if ((rc = make_operand_map(&tmY, w.Yh, false, B, w.Mpad, w.Ktot, 0, w.Cpad, TC_KBLK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_operand_map(&tmYe, w.Yh, false, B, w.Mpad, w.Ktot, w.Cpad, TC_KEXT, TC_KEXT, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;

    BtParams p{};
    p.N = N; p.M = M; p.C = C; p.Cpad = w.Cpad; p.KB = w.Cpad / TC_KBLK;
    p.a2 = alpha * kLog2e;
    // D = f32 (bits 4-5 = 1), A/B = f16 (0) at bits 7-9 / 10-12, b_major (bit 16): 0 = K-major, 1 = MN-major, N >> 3 at 17-22, M >> 4 at 24-28
    p.idesc1 = (1u << 4) | ((uint32_t)(BT_N >> 3) << 17) | ((uint32_t)(BT_M >> 4) << 24);
    p.idesc2 = (1u << 4) | (1u << 16) | ((uint32_t)(w.Cpad >> 3) << 17) | ((uint32_t)(BT_M >> 4) << 24);
    p.xx = w.xx; p.coef = w.coef; p.rmin = row_min; p.coef_max = w.coef_max;
    const size_t unit = (size_t)p.KB * BT_BLK + BT_EXT;
    const size_t smem = (1 + BT_NST) * unit + 2 * 2 * BT_BLK + (3 * 2 * BT_N + 2 * BT_M) * sizeof(float) + 256;
    static PerDeviceOnce attr_done;
    if (attr_done.need()) {
        DVM_CUDA(cudaFuncSetAttribute(softmap_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DVM_CUDA(cudaFuncSetAttribute(softmap_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done.done();
    }
    if (smem > 227 * 1024) { set_error("dvm_softmap_bwd_tc: needs %zu bytes of shared memory", smem); return DVM_ERR_UNSUPPORTED; }
    // owners = X rows, swept = Y columns  -> dX
    p.nOwner = N; p.nSwept = M; p.tiles = ceil_div(M, BT_N); p.own_feat = X; p.dOwn = dX;
    softmap_bwd_tc_kernel<true><<<dim3(ceil_div(N, BT_M), B), BT_THREADS, smem, st>>>(tmX, tmXe, tmY, tmYe, p);
    DVM_LAUNCH_CHECK();
    // owners = Y columns, swept = X rows  -> dY
    p.nOwner = M; p.nSwept = N; p.tiles = ceil_div(N, BT_N); p.own_feat = Y; p.dOwn = dY;
    softmap_bwd_tc_kernel<false><<<dim3(ceil_div(M, BT_M), B), BT_THREADS, smem, st>>>(tmY, tmYe, tmX, tmXe, p);
    DVM_LAUNCH_CHECK();
    softmap_bwd_tc_topk_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(X, Y, N, M, C, alpha, p.a2, topk, top_idx, top_w, top_d, dW, row_min, row_sum,
                                                                  w.coef, w.coef_max, w.xx, rows, dX, dY);
    DVM_LAUNCH_CHECK();
    return 0;
}
