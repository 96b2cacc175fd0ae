// Linear layer  out = act(x W^T + b)  of the Deformer's decoder MLP (models/model.py:433-452, 466) on tcgen05
// tensor cores with fp32-equivalent accuracy (3xTF32).
//
// Every fp32 operand value v is split on chip into  hi = tf32(v)  and  lo = v - hi  (exact in fp32); the product is
// accumulated as  lo_a*hi_b + hi_a*lo_b + hi_a*hi_b  by three kind::tf32 MMAs into one fp32 TMEM accumulator.  What is
// dropped (lo*lo and the tensor core's truncation of lo to 11 bits) is <= 2^-21 relative per product, the size of the
// fp32 rounding an SGEMM commits anyway -- so the layer stays inside the 1e-4 bound on deformed coordinates with two
// orders of magnitude to spare, at 1/3 of the TF32 tensor rate instead of the SIMT fp32 rate.
//
// Persistent kernel: one CTA per SM walks the tiles (128 rows of x = UMMA M, BN <= 256 output features = UMMA N; the
// N tiles of a row block are neighbours in the walk, so x is read from DRAM once).  K streams in blocks of 16 fp32
// (64-byte swizzle rows) through a 4-stage TMA ring that runs across tile boundaries.  Stage layout
// [x hi | W hi | x lo | W lo]: TMA writes the raw fp32 into the hi half, four converter warps split it in place (hi)
// and into the lo half and hand the stage to the MMA issuer (fence.proxy.async + mbarrier).  The accumulator is
// double-buffered in TMEM (2 x BN columns): eight epilogue warps (TMEM -> bias -> ELU -> global) drain tile i while
// the MMAs of tile i+1 run.
// Warp roles (448 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = converters,
// warps 6..13 = epilogue (TMEM lane quarter = warp % 4, column half = (warp - 6) / 4).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace dvm {

constexpr int LIN_BM = 128;
constexpr int LIN_BK = 16;                 // fp32 per K block = one 64-byte swizzle row
constexpr int LIN_NST = 4;
constexpr int LIN_THREADS = 448;
constexpr int LIN_ROW_BYTES = LIN_BK * 4;

// K-major SWIZZLE_64B operand descriptor: rows of 64 B, 8-row atoms 512 B apart
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float tf32_rn(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}

struct LinParams {
    int rows, K, N, BN;
    int n_tiles, tiles_total;      // N tiles per row block, all tiles
    int out_pitch, act;            // act: 0 = identity, 1 = ELU(alpha = 1)
    uint32_t idesc, tmem_cols;
    const float* bias;             // [N] or nullptr
    float* out;                    // [rows][out_pitch]
};

template <int BN>
__global__ void __launch_bounds__(LIN_THREADS, 1)
linear_tf32x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const LinParams p) {
    extern __shared__ uint8_t lin_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(lin_smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = (p.K + LIN_BK - 1) / LIN_BK;
    constexpr uint32_t half_bytes = (uint32_t)(LIN_BM + BN) * LIN_ROW_BYTES;      // hi (or lo) part of one stage
    constexpr uint32_t stage_bytes = 2 * half_bytes;
    constexpr int ACC_COLS = BN < 32 ? 32 : BN;                                   // TMEM columns per accumulator stage
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LIN_NST * stage_bytes);
    uint64_t* full = bars;                  // [NST] TMA -> converters
    uint64_t* conv = bars + LIN_NST;        // [NST] converters -> MMA
    uint64_t* empty = bars + 2 * LIN_NST;   // [NST] MMA -> TMA
    uint64_t* tfull = bars + 3 * LIN_NST;   // [2]   MMA -> epilogue
    uint64_t* tempty = tfull + 2;           // [2]   epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* bias_s = reinterpret_cast<float*>(tempty + 4);                         // [n_tiles * BN]
    for (int i = threadIdx.x; i < p.n_tiles * BN; i += LIN_THREADS) bias_s[i] = (p.bias != nullptr && i < p.N) ? __ldg(p.bias + i) : 0.f;

    if (threadIdx.x == 0) {
        for (int s = 0; s < LIN_NST; ++s) { mbar_init(&full[s], 1); mbar_init(&conv[s], 4); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int g = 0;                                                            // running K-block counter: ring slot and phase
            for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
                const int row0 = (tile / p.n_tiles) * LIN_BM, n0 = (tile % p.n_tiles) * BN;
                for (int kb = 0; kb < KB; ++kb, ++g) {
                    const int s = g % LIN_NST;
                    mbar_wait(&empty[s], ((g / LIN_NST) & 1) ^ 1);
                    uint8_t* st = smem + s * stage_bytes;
                    mbar_arrive_expect_tx(&full[s], half_bytes);
                    tma_load_3d(&tmX, &full[s], st, kb * LIN_BK, row0, 0);
                    tma_load_3d(&tmW, &full[s], st + LIN_BM * LIN_ROW_BYTES, kb * LIN_BK, n0, 0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int g = 0, it = 0;
            for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);                     // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(acc * ACC_COLS);
                for (int kb = 0; kb < KB; ++kb, ++g) {
                    const int s = g % LIN_NST;
                    mbar_wait(&conv[s], (g / LIN_NST) & 1);
                    tc_fence_after();
                    const uint32_t xa = smem_u32(smem + s * stage_bytes);
                    const uint32_t wa = xa + LIN_BM * LIN_ROW_BYTES;
#pragma unroll
                    for (int k = 0; k < LIN_BK / 8; ++k) {
                        const uint64_t xh = umma_desc_sw64(xa + k * 32), wh = umma_desc_sw64(wa + k * 32);
                        const uint64_t xl = umma_desc_sw64(xa + half_bytes + k * 32), wl = umma_desc_sw64(wa + half_bytes + k * 32);
                        tc_mma_tf32(d, xl, wh, p.idesc, (kb | k) != 0);
                        tc_mma_tf32(d, xh, wl, p.idesc, 1u);
                        tc_mma_tf32(d, xh, wh, p.idesc, 1u);
                    }
                    tc_commit(&empty[s]);
                }
                tc_commit(&tfull[acc]);
            }
        }
    } else if (warp < 6) {
        const int t = threadIdx.x - 64;                       // 0..127
        constexpr int n_vec = (int)(half_bytes >> 4);
        constexpr int n_it = (n_vec + 127) / 128;
        int g = 0;
        for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
            for (int kb = 0; kb < KB; ++kb, ++g) {
                const int s = g % LIN_NST;
                mbar_wait(&full[s], (g / LIN_NST) & 1);
                float4* hi = reinterpret_cast<float4*>(smem + s * stage_bytes);
                float4* lo = reinterpret_cast<float4*>(smem + s * stage_bytes + half_bytes);
                float4 v[n_it];
#pragma unroll
                for (int i = 0; i < n_it; ++i) if (n_vec % 128 == 0 || t + i * 128 < n_vec) v[i] = hi[t + i * 128];       // the split is element-wise: the swizzle is irrelevant here
#pragma unroll
                for (int i = 0; i < n_it; ++i) {
                    if (n_vec % 128 != 0 && t + i * 128 >= n_vec) break;
                    float4 h, l;
                    h.x = tf32_rn(v[i].x); h.y = tf32_rn(v[i].y); h.z = tf32_rn(v[i].z); h.w = tf32_rn(v[i].w);
                    l.x = v[i].x - h.x; l.y = v[i].y - h.y; l.z = v[i].z - h.z; l.w = v[i].w - h.w;
                    hi[t + i * 128] = h; lo[t + i * 128] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes -> visible to the MMA's async-proxy reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&conv[s]);
            }
        }
    } else {
        // ---- epilogue: a warp owns TMEM lanes (warp % 4) * 32 .. + 31 = rows of the block, and one column half
        const int q = warp & 3;
        const int hf = (warp - 6) >> 2;
        constexpr int CH = BN >= 32 ? BN / 2 : BN;                 // columns per epilogue warp
        const bool has_cols = BN >= 32 || hf == 0;
        const int cbeg = (BN >= 32) ? hf * CH : 0;
        const bool vec_ok = (p.out_pitch & 3) == 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const int row0 = (tile / p.n_tiles) * LIN_BM, n0 = (tile % p.n_tiles) * BN;
            mbar_wait_backoff(&tfull[acc], (it >> 1) & 1);
            tc_fence_after();
            if (has_cols) {
                const int row = row0 + q * 32 + lane;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS);
                float* orow = p.out + (size_t)row * p.out_pitch;
                float v[2][16];
                tc_ld16_issue(taddr + cbeg, v[0]);
#pragma unroll
                for (int ci = 0; ci < CH / 16; ++ci) {
                    const int c0 = cbeg + ci * 16;
                    float (&cur)[16] = v[ci & 1];
                    tc_ld16_wait(cur);
                    if (ci + 1 < CH / 16) tc_ld16_issue(taddr + c0 + 16, v[(ci + 1) & 1]);
                    const int n = n0 + c0;
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(bias_s + n + j);
                        cur[j] += b4.x; cur[j + 1] += b4.y; cur[j + 2] += b4.z; cur[j + 3] += b4.w;
                    }
                    if (p.act == 1) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) cur[j] = cur[j] > 0.f ? cur[j] : expf(cur[j]) - 1.f;     // torch's CUDA ELU formula
                    }
                    if (row < p.rows) {
                        if (vec_ok && n + 16 <= p.N) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(orow + n + j) = make_float4(cur[j], cur[j + 1], cur[j + 2], cur[j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) if (n + j < p.N) orow[n + j] = cur[j];
                        }
                    }
                }
            }
            tc_fence_before();                                 // this warp's TMEM reads are complete (wait::ld above)
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

static int make_f32_map(CUtensorMap* map, const float* base, long long rows, int K, int pitch, int box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return DVM_ERR_DEVICE; }
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, 1};
    cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)rows * pitch * 4};
    cuuint32_t box[3] = {(cuuint32_t)LIN_BK, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (fp32) failed with CUresult %d", (int)r); return DVM_ERR_DEVICE; }
    return 0;
}

}  // namespace dvm

using namespace dvm;

extern "C" int dvm_linear_act_fwd(const float* x, long long rows, int K, int x_pitch, const float* W, int w_pitch, const float* bias,
                                  int N, int act, float* out, int out_pitch, void* stream) {
    DVM_CHECK_ARG(x && W && out, "dvm_linear_act_fwd: null pointer");
    DVM_CHECK_ARG(rows >= 0 && rows < (1ll << 31) && K >= 1 && N >= 1, "dvm_linear_act_fwd: bad shape rows=%lld K=%d N=%d", rows, K, N);
    DVM_CHECK_ARG(x_pitch >= K && w_pitch >= K && out_pitch >= N, "dvm_linear_act_fwd: pitch smaller than the row length");
    DVM_CHECK_ARG((x_pitch & 3) == 0 && (w_pitch & 3) == 0, "dvm_linear_act_fwd: x_pitch and w_pitch must be multiples of 4 floats (TMA row stride)");
    DVM_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)out & 15) == 0, "dvm_linear_act_fwd: pointers must be 16-byte aligned");
    DVM_CHECK_ARG(act == 0 || act == 1, "dvm_linear_act_fwd: act must be 0 (identity) or 1 (ELU)");
    if (rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    LinParams p{};
    p.rows = (int)rows; p.K = K; p.N = N;
    p.BN = N >= 256 ? 256 : N > 64 ? 128 : N > 16 ? 64 : 16;
    p.out_pitch = out_pitch; p.act = act; p.bias = bias; p.out = out;
    p.n_tiles = ceil_div(N, p.BN);
    const long long tiles = (long long)p.n_tiles * ceil_div((int)rows, LIN_BM);
    if (tiles >= (1ll << 31)) { set_error("dvm_linear_act_fwd: rows=%lld too large", rows); return DVM_ERR_INVALID_ARG; }
    p.tiles_total = (int)tiles;
    p.tmem_cols = 32; while ((int)p.tmem_cols < 2 * (p.BN < 32 ? 32 : p.BN)) p.tmem_cols *= 2;      // two accumulator stages
    // instruction descriptor: D = f32 (bits 4-5 = 1), A/B = tf32 (2) at bits 7-9 / 10-12, K-major, N >> 3 at 17-22, M >> 4 at 24-28
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(LIN_BM >> 4) << 24);
    CUtensorMap tmX, tmW;
    int rc;
    if ((rc = make_f32_map(&tmX, x, rows, K, x_pitch, LIN_BM))) return rc;
    if ((rc = make_f32_map(&tmW, W, N, K, w_pitch, p.BN))) return rc;
    const size_t smem = (size_t)LIN_NST * 2 * (LIN_BM + p.BN) * LIN_ROW_BYTES + 1024 + 256 + (size_t)p.n_tiles * p.BN * 4;
    if (smem > 227 * 1024) { set_error("dvm_linear_act_fwd: N=%d needs %zu bytes of shared memory", N, smem); return DVM_ERR_UNSUPPORTED; }
    static PerDeviceOnce attr_done;
    if (attr_done.need()) {
        DVM_CUDA(cudaFuncSetAttribute(linear_tf32x3_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DVM_CUDA(cudaFuncSetAttribute(linear_tf32x3_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DVM_CUDA(cudaFuncSetAttribute(linear_tf32x3_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DVM_CUDA(cudaFuncSetAttribute(linear_tf32x3_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done.done();
    }
    const int grid = p.tiles_total < kNumSM ? p.tiles_total : kNumSM;          // persistent: one CTA per SM
    switch (p.BN) {
        case 256: linear_tf32x3_kernel<256><<<grid, LIN_THREADS, smem, st>>>(tmX, tmW, p); break;
        case 128: linear_tf32x3_kernel<128><<<grid, LIN_THREADS, smem, st>>>(tmX, tmW, p); break;
        case 64: linear_tf32x3_kernel<64><<<grid, LIN_THREADS, smem, st>>>(tmX, tmW, p); break;
        default: linear_tf32x3_kernel<16><<<grid, LIN_THREADS, smem, st>>>(tmX, tmW, p); break;
    }
    DVM_LAUNCH_CHECK();
    return 0;
}
