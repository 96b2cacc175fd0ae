// Operand preparation shared by the tcgen05 soft-map kernels (forward candidate pass, backward): fp32 -> 16-bit rows with the
// norm folded in as an extra 16-wide K block, and the TMA views of those arrays.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace dvm {

constexpr int TC_KBLK = 64;           // 16-bit elements per 128-byte swizzle row
constexpr int TC_KEXT = 16;           // extra K block: norm columns (one UMMA K step), 32-byte swizzle rows
constexpr int TC_PREP_ROWS = 32;      // rows per block of the operand preparation (8 warps x 4 rows)

// ------------------------------------------------------------------------------------------------
// operand preparation: fp32 -> 16-bit rows of pitch Ktot = Cpad + 16.
//   X rows:  [x~ (C), 0.., | 1, 1, 1, 0 x 13]                       xx[row] = |x~|^2, err_row = |x~ - x|_2
//   Y rows:  [-y~ (C), 0.., | h_hi, h_mid, h_lo, 0 x 13], h = |y~|^2/2; rows >= rows_per_b (padding up to a
//            multiple of 128): zeros with h_hi = +inf, so a padding column can never be selected.
// err_max[b] = max row rounding error of Y (certificate input), yy_max[b] = max |y~|^2.
// ------------------------------------------------------------------------------------------------
template <bool kBF16> struct Cvt16;
template <> struct Cvt16<false> {
    static __device__ __forceinline__ uint16_t bits(float v, float& back) { const __half h = __float2half_rn(v); back = __half2float(h); return __half_as_ushort(h); }
};
template <> struct Cvt16<true> {
    static __device__ __forceinline__ uint16_t bits(float v, float& back) { const __nv_bfloat16 h = __float2bfloat16_rn(v); back = __bfloat162float(h); return __bfloat16_as_ushort(h); }
};

template <bool kBF16, bool kIsY>
__global__ void __launch_bounds__(256)
tc_prep_kernel(const float* __restrict__ src, int rows_per_b, int rows_alloc, int C, int Cpad,
               uint16_t* __restrict__ dst, float* __restrict__ xx /* [B*rows_per_b], X only */,
               float* __restrict__ err_row /* X only */, float* __restrict__ err_max /* [B], Y only */,
               float* __restrict__ yy_max /* [B], Y only */) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int Ktot = Cpad + TC_KEXT;
    float dummy;
    float blk_err = 0.f, blk_yy = 0.f;           // Y: maxima over this warp's rows (one atomic pair per BLOCK at the end:
                                                 // a per-row atomicMax on one address serialises 200k updates)
    for (int rr = 0; rr < TC_PREP_ROWS; rr += 8) {
        const int r = blockIdx.x * TC_PREP_ROWS + rr + (threadIdx.x >> 5);
        if (r >= rows_alloc) break;
        uint16_t* d = dst + ((size_t)b * rows_alloc + r) * Ktot;
        if (r >= rows_per_b) {                       // Y padding row
            for (int c = lane; c < Ktot; c += 32) d[c] = (c == Cpad) ? Cvt16<kBF16>::bits(INFINITY, dummy) : (uint16_t)0;
            continue;
        }
        const float* s = src + ((size_t)b * rows_per_b + r) * C;
        float n2 = 0.f, e2 = 0.f;
        for (int c = lane * 4; c < Cpad; c += 128) {             // C % 4 == 0: float4 in, 4 x 16-bit (8 bytes) out
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < C) v = __ldg(reinterpret_cast<const float4*>(s + c));
            const float in[4] = {v.x, v.y, v.z, v.w};
            uint16_t o[4];
    #pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float x = kIsY ? -in[t] : in[t];
                float vr;
                o[t] = Cvt16<kBF16>::bits(x, vr);
                n2 = fmaf(vr, vr, n2);
                const float e = x - vr;                          // +-inf when the value overflows the 16-bit format
                e2 = fmaf(e, e, e2);
            }
            uint2 pk;
            pk.x = (uint32_t)o[0] | ((uint32_t)o[1] << 16);
            pk.y = (uint32_t)o[2] | ((uint32_t)o[3] << 16);
            *reinterpret_cast<uint2*>(d + c) = pk;
        }
        n2 = warp_sum(n2); e2 = warp_sum(e2);
        if (lane < TC_KEXT) {
            uint16_t w = 0;
            if (!kIsY) {
                if (lane < 3) w = Cvt16<kBF16>::bits(1.0f, dummy);
            } else {
                const float h = 0.5f * n2;
                float h0, h1, h2;
                const uint16_t b0 = Cvt16<kBF16>::bits(h, h0);
                const uint16_t b1 = Cvt16<kBF16>::bits(h - h0, h1);
                const uint16_t b2 = Cvt16<kBF16>::bits((h - h0) - h1, h2);
                w = lane == 0 ? b0 : lane == 1 ? b1 : lane == 2 ? b2 : (uint16_t)0;
            }
            d[Cpad + lane] = w;
        }
        {
            float e = sqrtf(e2) * 1.0001f;
            if (kIsY) {
                float hb; Cvt16<kBF16>::bits(0.5f * n2, hb);
                if (!(hb < INFINITY) || !(e < INFINITY)) e = INFINITY;          // |y|^2/2 not representable: nothing is certified
                blk_err = fmaxf(blk_err, e); blk_yy = fmaxf(blk_yy, n2);        // NaN-free: e, n2 >= 0 or +inf
            } else if (lane == 0) {
                xx[(size_t)b * rows_per_b + r] = n2;
                err_row[(size_t)b * rows_per_b + r] = e;
            }
        }
    }
    if (kIsY) {
        __shared__ float s_e[8], s_y[8];
        if (lane == 0) { s_e[threadIdx.x >> 5] = blk_err; s_y[threadIdx.x >> 5] = blk_yy; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float e = 0.f, y = 0.f;
            for (int w = 0; w < 8; ++w) { e = fmaxf(e, s_e[w]); y = fmaxf(y, s_y[w]); }
            atomicMax(reinterpret_cast<int*>(err_max + b), __float_as_int(e));   // values >= 0: int order == float order
            atomicMax(reinterpret_cast<int*>(yy_max + b), __float_as_int(y));
        }
    }
}

// view of a [B][rows][Ktot] 16-bit operand array starting at element column k0 with `kdim` columns:
// box {box_k, 128 rows, 1}; out-of-range rows read as zero
static inline int make_operand_map(CUtensorMap* map, const uint16_t* base, bool bf16, int B, int rows, int Ktot, int k0, int kdim, int box_k,
                            CUtensorMapSwizzle swz) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return DVM_ERR_DEVICE; }
    cuuint64_t dims[3] = {(cuuint64_t)kdim, (cuuint64_t)rows, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)Ktot * 2, (cuuint64_t)rows * Ktot * 2};
    cuuint32_t box[3] = {(cuuint32_t)box_k, 128u, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<uint16_t*>(base + k0),
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return DVM_ERR_DEVICE; }
    return 0;
}


}  // namespace dvm
