// Shared device/host helpers for libdvm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/dvm_b200.h"

namespace dvm {

constexpr int kNumSM = 148;            // B200: 2 dies x 74 SMs; grids are sized in multiples of this
constexpr int KC = 16;                 // candidates kept per row (or per partial row) before exact re-scoring
constexpr int P_MAX = 8;               // max partial candidate lists per row handed to finalize
constexpr float kExpCut = 32.0f;       // terms below exp(-32) of the row max are skipped: the dropped mass is
                                       // <= M * 1.3e-14 of a row sum that is >= 1 (M <= 1e6 -> < 1.3e-8 relative)
constexpr float kLog2e = 1.4426950408889634f;

void set_error(const char* fmt, ...);
void count_launch();                       // every kernel launch of the library is counted (dvm_launch_count)
void prof_begin(cudaStream_t st, int ch = 0);   // optional CUDA-event brackets: channel 0 = the dominant (candidate-pass)
void prof_end(cudaStream_t st, int ch = 0);     // kernels, channel 1 = the whole fused op dvm_softmap_fwd

#define DVM_CHECK_ARG(cond, ...)                                     \
    do { if (!(cond)) { dvm::set_error(__VA_ARGS__); return DVM_ERR_INVALID_ARG; } } while (0)

#define DVM_CUDA(call)                                                                    \
    do { cudaError_t e__ = (call);                                                        \
         if (e__ != cudaSuccess) {                                                        \
             dvm::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
             return (int)e__; } } while (0)

#define DVM_LAUNCH_CHECK()                                                                \
    do { dvm::count_launch(); cudaError_t e__ = cudaGetLastError();                                            \
         if (e__ != cudaSuccess) {                                                        \
             dvm::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
             return (int)e__; } } while (0)

// Per-device one-time setup (cudaFuncSetAttribute is a per-device property: a process that drives several GPUs must repeat it
// on each).  `if (once.need()) { ...; once.done(); }`
struct PerDeviceOnce {
    unsigned long long mask = 0;                       // bit d = done on device d (races only repeat an idempotent call)
    static int dev() { int d = 0; cudaGetDevice(&d); return d & 63; }
    bool need() const { return !((mask >> dev()) & 1ull); }
    void done() { mask |= 1ull << dev(); }
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Carves aligned sub-buffers out of the caller's workspace.
struct WsCarver {
    char* base; size_t off = 0; size_t cap;
    WsCarver(void* p, size_t bytes) : base((char*)p), cap(bytes) {}
    template <typename T> T* take(size_t n) {
        off = align_up(off, 256);
        T* r = (T*)(base ? base + off : nullptr);
        off += n * sizeof(T);
        return r;
    }
    bool ok() const { return off <= cap; }
};

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// (key, idx) lexicographic "less": smaller key first, lower index on ties
__device__ __forceinline__ bool kv_less(float ka, int ia, float kb, int ib) {
    return ka < kb || (ka == kb && ia < ib);
}

// Sorted (ascending) fixed-size candidate list living in registers.  `push` assumes the caller has
// already checked key < keys[K-1] (or wants the full bubble anyway); returns the evicted key.
template <int K>
struct TopList {
    float key[K];
    int   idx[K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int t = 0; t < K; ++t) { key[t] = INFINITY; idx[t] = -1; }
    }
    __device__ __forceinline__ float worst() const { return key[K - 1]; }
    // strict '<' against existing entries: with an ascending column scan equal keys keep the lower index
    __device__ __forceinline__ float push(float k, int j) {
        float ck = k; int ci = j;
#pragma unroll
        for (int t = 0; t < K; ++t) {
            const bool sw = ck < key[t];
            const float tk = key[t]; const int ti = idx[t];
            key[t] = sw ? ck : tk;  idx[t] = sw ? ci : ti;
            ck = sw ? tk : ck;      ci = sw ? ti : ci;
        }
        return ck;
    }
};

// Per-(partial-)row running state of the fused softmax/top-k sweep.
//   list : KC smallest squared distances seen so far (candidates for exact re-scoring)
//   l    : sum over NON-candidate columns of exp(-alpha (d_j - r)), r = sqrt(list.key[0]) (running min)
// Terms enter `l` either directly (column not good enough for the list) or when evicted from the list.
struct RowState {
    TopList<KC> list;
    float l, r, thr;   // thr: squared-distance threshold below which a column needs the slow path
    __device__ __forceinline__ void init() { list.init(); l = 0.f; r = INFINITY; thr = INFINITY; }
};

// slow path for one column with squared distance d2 (true domain), index j.  a2 = alpha*log2(e),
// cut_over_alpha = kExpCut/alpha (INFINITY when alpha == 0), soft = accumulate softmax terms.
template <bool kSoft>
__device__ __forceinline__ void row_state_visit(RowState& s, float d2, int j, float a2, float cut_over_alpha) {
    const float d2c = fmaxf(d2, 0.f);
    if (d2c < s.list.worst()) {
        const float ev = s.list.push(d2c, j);
        if (kSoft) {
            const float rn = sqrtf(s.list.key[0]);
            if (rn < s.r) {                                    // new running minimum: rescale l
                if (s.l != 0.f) s.l *= exp2f(-a2 * (s.r - rn));
                s.r = rn;
            }
            if (ev != INFINITY) s.l += exp2f(-a2 * (sqrtf(ev) - s.r));
            const float te = s.r + cut_over_alpha;
            s.thr = fmaxf(s.list.worst(), te * te);
        } else {
            s.thr = s.list.worst();
        }
    } else if (kSoft) {
        s.l += exp2f(-a2 * (sqrtf(d2c) - s.r));
    }
}

}  // namespace dvm
