// Exact k-NN / Chamfer on 3-D points through a uniform grid: O(N) work instead of the O(N*M) brute force.
//
// Results are bit-identical to the brute-force kernel of knn3.cu (same unfused fp32 / fp64 distance
// arithmetic, (d^2, index) lexicographic order): the grid only decides WHICH reference points a query looks
// at, and the search of a query stops only when every unvisited cell is provably farther than its current
// k-th neighbour (with a safety margin that covers the fp32 rounding of the cell assignment).
//
// Per reference cloud (batched over B): bounding box -> cell size h = sqrt(box surface / M) (about 4 points
// per occupied cell for surface samples), capped so that cells <= 4 M -> counting sort of the points by cell
// (x fastest, so a run of cells along x is one contiguous span of points).  Per query: a growing box of cells
// around its (clamped) cell, visited as thick shells of span look-ups; queries on the surface finish after
// 2-3 steps (~25 look-ups), queries far from the reference cloud skip the empty space through a coarse grid.
#include "common.cuh"

namespace dvm {

constexpr int GRID_THREADS = 256;
constexpr int GRID_QTHREADS = 128;       // query kernel: small blocks -- the work per query varies a lot (dense regions, far queries), and
                                         // the last wave of big blocks left half of the SMs idle

constexpr int GRID_COARSE = 8;    // a coarse cell = 8 x 8 x 8 fine cells: used to skip empty space around far queries
constexpr int GRID_COARSE_MAX = 1 << 16;
// cell edge = GRID_H_SCALE * sqrt(bounding-box surface / M): ~4 points per occupied cell of a surface cloud.  Measured at
// N = 50k (xyz 10-NN | Chamfer | match+deform step, ms): scale 1.0: 0.38 | 0.54 | 7.56, 2.0: 0.29 | 0.49 | 7.34, 3.0: 0.34 | 0.48 | 7.55
// -- with one point per cell the first 3x3x3 box rarely holds 10 points and every row of cells costs ~40 instructions.
constexpr float GRID_H_SCALE = 2.0f;

struct GridHeader {          // one per cloud, written by grid_setup_kernel
    float x0, y0, z0, h, inv_h;
    int nx, ny, nz, ncell;
    int cnx, cny, cnz, cncell;   // coarse grid
};

__global__ void __launch_bounds__(1024)
grid_setup_kernel(const float* __restrict__ R, int M, int ncell_max, float h_scale, GridHeader* __restrict__ hdr) {
    __shared__ float s_lo[3][32], s_hi[3][32];
    const int b = blockIdx.x;
    const float* P = R + (size_t)b * M * 3;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = threadIdx.x; i < M; i += 1024) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { const float v = P[i * 3 + c]; lo[c] = fminf(lo[c], v); hi[c] = fmaxf(hi[c], v); }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if (lane == 0) { s_lo[c][wid] = lo[c]; s_hi[c][wid] = hi[c]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int c = 0; c < 3; ++c) {
            lo[c] = s_lo[c][0]; hi[c] = s_hi[c][0];
            for (int w = 1; w < 32; ++w) { lo[c] = fminf(lo[c], s_lo[c][w]); hi[c] = fmaxf(hi[c], s_hi[c][w]); }
        }
        float ext[3];
        for (int c = 0; c < 3; ++c) { ext[c] = hi[c] - lo[c]; if (!(ext[c] > 0.f) || !(ext[c] < INFINITY)) ext[c] = 0.f; }
        const float emax = fmaxf(ext[0], fmaxf(ext[1], ext[2]));
        float area = 2.f * (ext[0] * ext[1] + ext[1] * ext[2] + ext[0] * ext[2]);
        float h = h_scale * sqrtf(area / (float)M);
        if (!(h > emax * 1e-4f)) h = emax * 1e-4f;        // degenerate (collinear / coincident) clouds
        if (!(h > 0.f)) h = 1.f;
        int nx, ny, nz;
        for (;;) {                                        // coarsen until the cell budget fits
            nx = (int)fminf(ext[0] / h, 2e6f) + 1; ny = (int)fminf(ext[1] / h, 2e6f) + 1; nz = (int)fminf(ext[2] / h, 2e6f) + 1;
            const double cn = (double)((nx + GRID_COARSE - 1) / GRID_COARSE) * ((ny + GRID_COARSE - 1) / GRID_COARSE) * ((nz + GRID_COARSE - 1) / GRID_COARSE);
            if ((double)nx * ny * nz <= (double)ncell_max && cn <= (double)GRID_COARSE_MAX) break;
            h *= 1.26f;
        }
        GridHeader g;
        g.x0 = lo[0]; g.y0 = lo[1]; g.z0 = lo[2]; g.h = h; g.inv_h = 1.f / h;
        g.nx = nx; g.ny = ny; g.nz = nz; g.ncell = nx * ny * nz;
        g.cnx = (nx + GRID_COARSE - 1) / GRID_COARSE; g.cny = (ny + GRID_COARSE - 1) / GRID_COARSE; g.cnz = (nz + GRID_COARSE - 1) / GRID_COARSE;
        g.cncell = g.cnx * g.cny * g.cnz;
        hdr[b] = g;
    }
}

__device__ __forceinline__ int cell_coord(float v, float v0, float inv_h, int n) {
    const float t = (v - v0) * inv_h;
    int c = (int)floorf(t);
    if (!(t >= 0.f)) c = 0;                               // also catches NaN
    return c < n ? c : n - 1;
}

__global__ void __launch_bounds__(GRID_THREADS)
grid_count_kernel(const float* __restrict__ R, int M, const GridHeader* __restrict__ hdr, int stride,
                  int* __restrict__ cell_of, int* __restrict__ count, int* __restrict__ ccount) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * GRID_THREADS + threadIdx.x;
    if (i >= M) return;
    const GridHeader g = hdr[b];
    const float* p = R + ((size_t)b * M + i) * 3;
    const int cx = cell_coord(p[0], g.x0, g.inv_h, g.nx), cy = cell_coord(p[1], g.y0, g.inv_h, g.ny), cz = cell_coord(p[2], g.z0, g.inv_h, g.nz);
    const int c = (cz * g.ny + cy) * g.nx + cx;
    cell_of[(size_t)b * M + i] = c;
    atomicAdd(count + (size_t)b * stride + c, 1);
    const int cc = ((cz / GRID_COARSE) * g.cny + cy / GRID_COARSE) * g.cnx + cx / GRID_COARSE;
    atomicAdd(ccount + (size_t)b * (GRID_COARSE_MAX + 1) + cc, 1);
}

// Exclusive scan of count[0..n) -> start[0..n] in place (cursor = copy of start for the scatter), three phases so
// that it scales over the SMs: (A) sums of 4096-cell blocks, (B) one block per cloud scans the block sums,
// (C) every block rescans its cells with its offset.  blockIdx.z: 0 = fine grid, 1 = coarse grid.
constexpr int SCAN_IT = 4;
constexpr int SCAN_BLOCK = 1024 * SCAN_IT;

struct ScanArgs {
    const GridHeader* hdr; int stride; int* start; int* cursor; int cstride; int* cstart;
    int* bsum; int nblk_max;       // [2][B][nblk_max]
};

__device__ __forceinline__ int scan_n(const ScanArgs& a, int b, bool coarse) { return coarse ? a.hdr[b].cncell : a.hdr[b].ncell; }
__device__ __forceinline__ int* scan_ptr(const ScanArgs& a, int b, bool coarse) {
    return coarse ? a.cstart + (size_t)b * a.cstride : a.start + (size_t)b * a.stride;
}

__global__ void __launch_bounds__(1024) grid_scan_a_kernel(ScanArgs a, int B) {
    __shared__ int s_w[32];
    const int b = blockIdx.y; const bool coarse = blockIdx.z == 1;
    const int n = scan_n(a, b, coarse);
    const int base = blockIdx.x * SCAN_BLOCK;
    if (base >= n) return;
    const int* st = scan_ptr(a, b, coarse);
    int sum = 0;
#pragma unroll
    for (int q = 0; q < SCAN_IT; ++q) { const int i = base + threadIdx.x * SCAN_IT + q; sum += i < n ? st[i] : 0; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x < 32) {
        int w = s_w[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (threadIdx.x == 0) a.bsum[((size_t)blockIdx.z * B + b) * a.nblk_max + blockIdx.x] = w;
    }
}

__global__ void __launch_bounds__(1024) grid_scan_b_kernel(ScanArgs a, int B) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int b = blockIdx.x; const bool coarse = blockIdx.y == 1;
    const int n = scan_n(a, b, coarse);
    const int nblk = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    int* bs = a.bsum + ((size_t)blockIdx.y * B + b) * a.nblk_max;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nblk; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nblk ? bs[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int w = s_w[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            s_w[lane] = w;
        }
        __syncthreads();
        const int incl = s_carry + (wid ? s_w[wid - 1] : 0) + inc;
        if (i < nblk) bs[i] = incl - v;                    // exclusive offset of block i
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) scan_ptr(a, b, coarse)[n] = s_carry;     // total = number of points
}

__global__ void __launch_bounds__(1024) grid_scan_c_kernel(ScanArgs a, int B) {
    __shared__ int s_w[32];
    const int b = blockIdx.y; const bool coarse = blockIdx.z == 1;
    const int n = scan_n(a, b, coarse);
    const int base = blockIdx.x * SCAN_BLOCK;
    if (base >= n) return;
    int* st = scan_ptr(a, b, coarse);
    int* cu = coarse ? st : a.cursor + (size_t)b * a.stride;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int v[SCAN_IT], sum = 0;
#pragma unroll
    for (int q = 0; q < SCAN_IT; ++q) { const int i = base + threadIdx.x * SCAN_IT + q; v[q] = i < n ? st[i] : 0; sum += v[q]; }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = s_w[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        s_w[lane] = w;
    }
    __syncthreads();
    int excl = a.bsum[((size_t)blockIdx.z * B + b) * a.nblk_max + blockIdx.x] + (wid ? s_w[wid - 1] : 0) + inc - sum;
#pragma unroll
    for (int q = 0; q < SCAN_IT; ++q) {
        const int i = base + threadIdx.x * SCAN_IT + q;
        if (i < n) { st[i] = excl; cu[i] = excl; }
        excl += v[q];
    }
}

__global__ void __launch_bounds__(GRID_THREADS)
grid_scatter_kernel(const float* __restrict__ R, int M, int stride, const int* __restrict__ cell_of,
                    int* __restrict__ cursor, float4* __restrict__ sorted) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * GRID_THREADS + threadIdx.x;
    if (i >= M) return;
    const float* p = R + ((size_t)b * M + i) * 3;
    const int c = cell_of[(size_t)b * M + i];
    const int pos = atomicAdd(cursor + (size_t)b * stride + c, 1);
    sorted[(size_t)b * M + pos] = make_float4(p[0], p[1], p[2], __int_as_float(i));
}

// queries sorted by the cell of the reference grid they fall into (clamped): warps then hold spatial neighbours, which walk
// the same rows and spans (little divergence, L1-friendly).  Same count / scan / scatter machinery as the points.
__global__ void __launch_bounds__(GRID_THREADS)
grid_qcount_kernel(const float* __restrict__ Q, int N, const GridHeader* __restrict__ hdr, int stride,
                   int* __restrict__ qcell, int* __restrict__ qcount) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * GRID_THREADS + threadIdx.x;
    if (i >= N) return;
    const GridHeader g = hdr[b];
    const float* p = Q + ((size_t)b * N + i) * 3;
    const int cx = cell_coord(p[0], g.x0, g.inv_h, g.nx), cy = cell_coord(p[1], g.y0, g.inv_h, g.ny), cz = cell_coord(p[2], g.z0, g.inv_h, g.nz);
    const int c = (cz * g.ny + cy) * g.nx + cx;
    qcell[(size_t)b * N + i] = c;
    atomicAdd(qcount + (size_t)b * stride + c, 1);
}

__global__ void __launch_bounds__(GRID_THREADS)
grid_qscatter_kernel(int N, int stride, const int* __restrict__ qcell, int* __restrict__ qcursor, int* __restrict__ qorder) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * GRID_THREADS + threadIdx.x;
    if (i >= N) return;
    const int pos = atomicAdd(qcursor + (size_t)b * stride + qcell[(size_t)b * N + i], 1);
    qorder[(size_t)b * N + pos] = i;
}

// occupied x-extent of every (z, y) row of cells: rowx[row] = (first, last) non-empty cell, (1, 0) when the row is
// empty.  ny * nz entries per cloud -- small enough to stay in L1, so that the (mostly empty) rows a far query walks
// through are rejected without touching the cell-start array in L2.
__global__ void __launch_bounds__(GRID_THREADS)
grid_rowinfo_kernel(const GridHeader* __restrict__ hdr, int stride, const int* __restrict__ start, int2* __restrict__ rowx) {
    const int b = blockIdx.y;
    const GridHeader g = hdr[b];
    const int row = blockIdx.x * GRID_THREADS + threadIdx.x;
    if (row >= g.ny * g.nz) return;
    const int* st = start + (size_t)b * stride + (size_t)row * g.nx;
    int lo = 1, hi = 0;
    int prev = st[0];
    for (int x = 0; x < g.nx; ++x) {
        const int nxt = st[x + 1];
        if (nxt != prev) { if (hi < lo) lo = x; hi = x; }
        prev = nxt;
    }
    rowx[(size_t)b * stride + row] = make_int2(lo, hi);
}

template <typename T> struct GDist;
template <> struct GDist<float> {
    static __device__ __forceinline__ float eval(float qx, float qy, float qz, float4 p) {
        const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
        return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    }
};
template <> struct GDist<double> {
    static __device__ __forceinline__ double eval(float qx, float qy, float qz, float4 p) {
        const double dx = (double)qx - (double)p.x, dy = (double)qy - (double)p.y, dz = (double)qz - (double)p.z;
        return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    }
};

template <typename T, int K>
struct GList {
    T key[K]; int idx[K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int t = 0; t < K; ++t) { key[t] = (T)INFINITY; idx[t] = 0x7fffffff; }
    }
    __device__ __forceinline__ void push(T k, int j) {     // lexicographic (key, idx) insertion
        T ck = k; int ci = j;
#pragma unroll
        for (int t = 0; t < K; ++t) {
            const bool sw = ck < key[t] || (ck == key[t] && ci < idx[t]);
            const T tk = key[t]; const int ti = idx[t];
            key[t] = sw ? ck : tk; idx[t] = sw ? ci : ti;
            ck = sw ? tk : ck; ci = sw ? ti : ci;
        }
    }
};

template <typename T, int K>
__device__ __forceinline__ void scan_span(GList<T, K>& list, const float4* __restrict__ pts, int s, int e, float qx, float qy, float qz) {
    for (int i = s; i < e; ++i) {
        const float4 p = __ldg(pts + i);
        const T d = GDist<T>::eval(qx, qy, qz, p);
        const int j = __float_as_int(p.w);
        if (d < list.key[K - 1] || (d == list.key[K - 1] && j < list.idx[K - 1])) list.push(d, j);
    }
}

// GRID_LPQ lanes per query: each lane takes every GRID_LPQ-th (y, z) row of the box and the lists are merged at
// the end.  8 for Chamfer (k = 1: queries are typically OFF the reference surface and walk hundreds of mostly
// empty rows -> 8 independent chains of dependent loads), 1 for the self / node k-NN (queries on the surface).

// kSelf: Q is the reference cloud itself -> thread t handles the t-th point in CELL-SORTED order, so the lanes of a
// warp are spatial neighbours (same rows, same spans: little divergence, L1-friendly) and write to their original row.
template <typename T, int K, int GRID_LPQ, bool kSelf>
__global__ void __launch_bounds__(GRID_QTHREADS)
knn3_grid_kernel(const float* __restrict__ Q, int N, int M, int k, const GridHeader* __restrict__ hdr, int stride,
                 const int* __restrict__ start, const int* __restrict__ cstart, const int2* __restrict__ rowx_all,
                 const float4* __restrict__ sorted, const int* __restrict__ qorder, int64_t* __restrict__ idx64, int32_t* __restrict__ idx32, float* __restrict__ d2f, double* __restrict__ d2d) {
    const int b = blockIdx.y;
    int q = blockIdx.x * (GRID_QTHREADS / GRID_LPQ) + threadIdx.x / GRID_LPQ;
    const int sub = threadIdx.x % GRID_LPQ;
    const bool live = q < N;
    const GridHeader g = hdr[b];
    const int* st = start + (size_t)b * stride;
    const int2* rowx = rowx_all + (size_t)b * stride;
    const float4* pts = sorted + (size_t)b * M;
    float qx = g.x0, qy = g.y0, qz = g.z0;
    if (live) {
        if (kSelf) {
            const float4 p = __ldg(pts + q);
            qx = p.x; qy = p.y; qz = p.z; q = __float_as_int(p.w);
        } else {
            if (qorder) q = __ldg(qorder + (size_t)b * N + q);         // cell-sorted query order
            const float* qp = Q + ((size_t)b * N + q) * 3;
            qx = __ldg(qp); qy = __ldg(qp + 1); qz = __ldg(qp + 2);
        }
    }
    const int cx = cell_coord(qx, g.x0, g.inv_h, g.nx), cy = cell_coord(qy, g.y0, g.inv_h, g.ny), cz = cell_coord(qz, g.z0, g.inv_h, g.nz);
    GList<T, K> list;
    list.init();
    const float margin = 1e-3f * g.h;                    // covers the fp32 rounding of the points' cell assignment
    const unsigned full = 0xffffffffu;
    // The visited region is always a box [c - sp, c + sp]^3 of cells; every step visits the "thick shell" between the
    // old box and a larger one (rows that crossed the old box contribute their two end spans only).  Step sizes:
    //   * while fewer than k points are known: grow geometrically; the coarse grid (8^3 fine cells per coarse cell)
    //     tells once how far the first populated region is, so empty space around far queries is skipped;
    //   * once k points are known: jump straight to the box that contains the ball of the largest known distance.
    // Correctness never depends on the step heuristic: the loop ends only when k known points are strictly closer
    // than the nearest non-boundary face of the visited box (or the box covers the grid).  All lanes of a team run
    // the loop in lock step (the step decisions are team-uniform).
    int sp = -1, s = GRID_LPQ > 1 ? 1 : 0;
    bool coarse_done = false;
    T team_ub = (T)INFINITY;          // upper bound of the team's k-th distance: min over lanes that know k points
    for (;;) {
        const int zlo = max(cz - s, 0), zhi = min(cz + s, g.nz - 1);
        const int ylo = max(cy - s, 0), yhi = min(cy + s, g.ny - 1);
        const int xlo = max(cx - s, 0), xhi = min(cx + s, g.nx - 1);
        const int wy = yhi - ylo + 1;
        const int nrows = (zhi - zlo + 1) * wy;
        // (z, y) of row r = sub + i * GRID_LPQ, advanced without a division per row
        const int step_z = GRID_LPQ / wy, step_y = GRID_LPQ % wy;
        int z = zlo + sub / wy - step_z, y = ylo + sub % wy - step_y;
        for (int r = sub; r < nrows; r += GRID_LPQ) {
            z += step_z; y += step_y;
            if (y > yhi) { y -= wy; ++z; }
            const int2 ext = __ldg(rowx + z * g.ny + y);       // occupied cells of the row (L1-resident table)
            int x0 = max(xlo, ext.x), x1 = min(xhi, ext.y);
            if (x0 > x1) continue;                             // nothing of the row inside the box
            // ball pruning with the lane's own k-th distance (an upper bound of the final one): a row whose (y, z)
            // slab is farther than that is skipped, the others are clipped to the chord of the ball.  Margins keep
            // every point that could tie (equal distance, lower index) inside.
            T kth_l = team_ub;
#pragma unroll
            for (int t = 0; t < K; ++t) if (t == k - 1) kth_l = list.key[t] < kth_l ? list.key[t] : kth_l;
            if (kth_l < (T)INFINITY) {
                const float ylo_f = g.y0 + (float)y * g.h, zlo_f = g.z0 + (float)z * g.h;
                const float dy = fmaxf(fmaxf(ylo_f - qy, qy - (ylo_f + g.h)) - margin, 0.f);
                const float dz = fmaxf(fmaxf(zlo_f - qz, qz - (zlo_f + g.h)) - margin, 0.f);
                const float lb2 = (dy * dy + dz * dz) * (1.f - 1e-5f);
                const float best = (float)kth_l * (1.f + 1e-5f) + 1e-30f;
                if (lb2 > best) continue;
                const float rx = sqrtf(best - lb2) * (1.f + 1e-5f) + margin;
                const float fa = (qx - rx - g.x0) * g.inv_h - 1.f, fb = (qx + rx - g.x0) * g.inv_h + 1.f;
                if (fa > (float)x0) x0 = (int)fminf(fa, 4e6f);
                if (fb < (float)x1) x1 = (int)fmaxf(fb, -1.f);
                if (x0 > x1) continue;
            }
            const int row = (z * g.ny + y) * g.nx;
            if (z >= cz - sp && z <= cz + sp && y >= cy - sp && y <= cy + sp) {   // row crossed the old box: two end spans
                const int l1 = min(cx - sp - 1, x1);          // left span  [x0, l1]
                if (l1 >= x0) scan_span<T, K>(list, pts, st[row + x0], st[row + l1 + 1], qx, qy, qz);
                const int r0 = max(cx + sp + 1, x0);          // right span [r0, x1]
                if (r0 <= x1) scan_span<T, K>(list, pts, st[row + r0], st[row + x1 + 1], qx, qy, qz);
            } else {
                scan_span<T, K>(list, pts, st[row + x0], st[row + x1 + 1], qx, qy, qz);
            }
        }
        sp = s;
        // every unvisited point lies beyond one of the box faces that are not on the grid boundary
        float bound = INFINITY;
        if (cx - s > 0)        bound = fminf(bound, qx - (g.x0 + (float)(cx - s) * g.h));
        if (cx + s < g.nx - 1) bound = fminf(bound, (g.x0 + (float)(cx + s + 1) * g.h) - qx);
        if (cy - s > 0)        bound = fminf(bound, qy - (g.y0 + (float)(cy - s) * g.h));
        if (cy + s < g.ny - 1) bound = fminf(bound, (g.y0 + (float)(cy + s + 1) * g.h) - qy);
        if (cz - s > 0)        bound = fminf(bound, qz - (g.z0 + (float)(cz - s) * g.h));
        if (cz + s < g.nz - 1) bound = fminf(bound, (g.z0 + (float)(cz + s + 1) * g.h) - qz);
        if (bound == INFINITY) break;                     // the box covers the whole grid (team-uniform)
        bound = fmaxf(bound - margin, 0.f);
        // team-wide: how many known points are strictly inside the safe radius, how many are known at all, and the
        // largest known distance (an upper bound of the k-th distance once k points are known)
        const T safe = (T)bound * (T)bound * (T)(1.0 - 1e-6);
        int n_safe = 0, n_known = 0; T far = (T)0;
#pragma unroll
        for (int t = 0; t < K; ++t) {
            n_safe += list.key[t] < safe ? 1 : 0;
            if (list.key[t] < (T)INFINITY) { ++n_known; far = list.key[t]; }
            if (t == k - 1 && list.key[t] < team_ub) team_ub = list.key[t];
        }
#pragma unroll
        for (int o = GRID_LPQ / 2; o > 0; o >>= 1) {
            n_safe += __shfl_xor_sync(full, n_safe, o);
            n_known += __shfl_xor_sync(full, n_known, o);
            const T of = __shfl_xor_sync(full, far, o);
            far = of > far ? of : far;
            const T ou = __shfl_xor_sync(full, team_ub, o);
            team_ub = ou < team_ub ? ou : team_ub;
        }
        if (n_safe >= k) break;
        if (team_ub < (T)INFINITY) far = team_ub;
        if (n_known >= k) {
            // half-width whose box contains the ball of radius sqrt(far) around q (q may sit anywhere in its cell,
            // or outside the grid next to it)
            const float rad = sqrtf((float)far) + margin;
            const float ox = fmaxf(fmaxf(g.x0 - qx, qx - (g.x0 + (float)g.nx * g.h)), 0.f);
            const float oy = fmaxf(fmaxf(g.y0 - qy, qy - (g.y0 + (float)g.ny * g.h)), 0.f);
            const float oz = fmaxf(fmaxf(g.z0 - qz, qz - (g.z0 + (float)g.nz * g.h)), 0.f);
            const float reach = fmaxf(rad - fminf(ox, fminf(oy, oz)), 0.f);
            const int need = (int)fminf(reach * g.inv_h, 4e6f) + 1;
            s = max(s + 1, need);
        } else {
            int jump = 2 * s + 1;
            if (!coarse_done) {
                coarse_done = true;
                const int* cst = cstart + (size_t)b * (GRID_COARSE_MAX + 1);
                const int Cx = cx / GRID_COARSE, Cy = cy / GRID_COARSE, Cz = cz / GRID_COARSE;
                int S = 0;
                for (;; ++S) {
                    bool found = false;
                    const int zl = max(Cz - S, 0), zh = min(Cz + S, g.cnz - 1), yl = max(Cy - S, 0), yh = min(Cy + S, g.cny - 1);
                    const int xl = max(Cx - S, 0), xh = min(Cx + S, g.cnx - 1);
                    const int cw = yh - yl + 1, cn = (zh - zl + 1) * cw;
                    for (int r = sub; r < cn; r += GRID_LPQ) {
                        const int row = ((zl + r / cw) * g.cny + yl + r % cw) * g.cnx;
                        if (cst[row + xh + 1] != cst[row + xl]) { found = true; break; }
                    }
#pragma unroll
                    for (int o = GRID_LPQ / 2; o > 0; o >>= 1) found |= (bool)__shfl_xor_sync(full, (int)found, o);
                    if (found || (xl == 0 && yl == 0 && zl == 0 && xh == g.cnx - 1 && yh == g.cny - 1 && zh == g.cnz - 1)) break;
                }
                // the coarse box of half-width S-1 is empty, hence so are the fine shells 0 .. (S-1)*8
                if (S > 1) jump = max(jump, (S - 1) * GRID_COARSE + 1);
            }
            s = jump;
        }
    }
    // merge the GRID_LPQ sorted lists of the team: k rounds of team-wide lexicographic min
    int head = 0;
    for (int r = 0; r < k; ++r) {
        T hk = (T)INFINITY; int hi = 0x7fffffff;
#pragma unroll
        for (int t = 0; t < K; ++t) if (t == head) { hk = list.key[t]; hi = list.idx[t]; }
        T wk = hk; int wi = hi;
#pragma unroll
        for (int o = GRID_LPQ / 2; o > 0; o >>= 1) {
            const T ok = __shfl_xor_sync(full, wk, o);
            const int oi = __shfl_xor_sync(full, wi, o);
            if (ok < wk || (ok == wk && oi < wi)) { wk = ok; wi = oi; }
        }
        if (hi == wi && hi != 0x7fffffff) ++head;
        if (live && sub == 0) {
            const size_t o = ((size_t)b * N + q) * k + r;
            if (idx64) idx64[o] = wi;
            if (idx32) idx32[o] = wi;
            if (d2f) d2f[o] = (float)wk;
            if (d2d) d2d[o] = (double)wk;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct GridWs { GridHeader* hdr; int* start; int* cursor; int* cstart; int* cell_of; float4* sorted; int* bsum; int2* rowx;
                int* qstart; int* qcursor; int* qcell; int* qorder; int stride; int ncell_max; int nblk_max; };

static size_t grid_ws_layout(void* base, size_t cap, int B, int N, int M, GridWs* out) {
    GridWs w{};
    long long nc = 4LL * M + 64;
    if (nc > (1LL << 24)) nc = 1LL << 24;
    w.ncell_max = (int)nc;
    w.stride = w.ncell_max + 1;
    WsCarver ws(base, cap);
    w.hdr = ws.take<GridHeader>(B);
    w.start = ws.take<int>((size_t)B * w.stride);
    w.cursor = ws.take<int>((size_t)B * w.stride);
    w.cstart = ws.take<int>((size_t)B * (GRID_COARSE_MAX + 1));
    w.cell_of = ws.take<int>((size_t)B * M);
    w.sorted = ws.take<float4>((size_t)B * M);
    w.nblk_max = ceil_div(w.ncell_max > GRID_COARSE_MAX ? w.ncell_max : GRID_COARSE_MAX, SCAN_BLOCK);
    w.bsum = ws.take<int>((size_t)2 * B * w.nblk_max);
    w.rowx = ws.take<int2>((size_t)B * w.stride);
    w.qstart = ws.take<int>((size_t)B * w.stride);          // queries sorted by reference-grid cell (non-self queries)
    w.qcursor = ws.take<int>((size_t)B * w.stride);
    w.qcell = ws.take<int>((size_t)B * N);
    w.qorder = ws.take<int>((size_t)B * N);
    if (out) *out = w;
    return align_up(ws.off, 256);
}

size_t knn3_grid_workspace_bytes(int B, int N, int M) { return grid_ws_layout(nullptr, 0, B, N, M, nullptr); }

template <typename T, int K, int LPQ = 1>
static void launch_query(const float* Q, bool self, const int* qorder, int B, int N, int M, int k, const GridWs& w,
                         int64_t* idx64, int32_t* idx32, float* d2f, double* d2d, cudaStream_t st) {
    dim3 grid(ceil_div(N, GRID_QTHREADS / LPQ), B);
    if (self) knn3_grid_kernel<T, K, LPQ, true><<<grid, GRID_QTHREADS, 0, st>>>(Q, N, M, k, w.hdr, w.stride, w.start, w.cstart, w.rowx, w.sorted, qorder, idx64, idx32, d2f, d2d);
    else      knn3_grid_kernel<T, K, LPQ, false><<<grid, GRID_QTHREADS, 0, st>>>(Q, N, M, k, w.hdr, w.stride, w.start, w.cstart, w.rowx, w.sorted, qorder, idx64, idx32, d2f, d2d);
}

// k nearest neighbours of Q[B,N,3] in R[B,M,3] through a grid built on R (inside ws)
int launch_knn3_grid(const float* Q, const float* R, int B, int N, int M, int k, bool f64,
                     int64_t* idx64, int32_t* idx32, float* d2f, double* d2d, void* wsp, size_t ws_bytes, cudaStream_t st) {
    GridWs w;
    const size_t need = grid_ws_layout(wsp, ws_bytes, B, N, M, &w);
    if (!wsp || need > ws_bytes) { set_error("knn3 grid: workspace too small (%zu < %zu)", ws_bytes, need); return DVM_ERR_WORKSPACE; }
    grid_setup_kernel<<<B, 1024, 0, st>>>(R, M, w.ncell_max, GRID_H_SCALE, w.hdr);
    DVM_LAUNCH_CHECK();
    DVM_CUDA(cudaMemsetAsync(w.start, 0, (size_t)B * w.stride * sizeof(int), st));
    DVM_CUDA(cudaMemsetAsync(w.cstart, 0, (size_t)B * (GRID_COARSE_MAX + 1) * sizeof(int), st));
    dim3 gp(ceil_div(M, GRID_THREADS), B);
    grid_count_kernel<<<gp, GRID_THREADS, 0, st>>>(R, M, w.hdr, w.stride, w.cell_of, w.start, w.cstart);
    DVM_LAUNCH_CHECK();
    ScanArgs sa{w.hdr, w.stride, w.start, w.cursor, GRID_COARSE_MAX + 1, w.cstart, w.bsum, w.nblk_max};
    grid_scan_a_kernel<<<dim3(w.nblk_max, B, 2), 1024, 0, st>>>(sa, B);
    DVM_LAUNCH_CHECK();
    grid_scan_b_kernel<<<dim3(B, 2), 1024, 0, st>>>(sa, B);
    DVM_LAUNCH_CHECK();
    grid_scan_c_kernel<<<dim3(w.nblk_max, B, 2), 1024, 0, st>>>(sa, B);
    DVM_LAUNCH_CHECK();
    grid_scatter_kernel<<<gp, GRID_THREADS, 0, st>>>(R, M, w.stride, w.cell_of, w.cursor, w.sorted);
    DVM_LAUNCH_CHECK();
    grid_rowinfo_kernel<<<dim3(ceil_div(w.ncell_max, GRID_THREADS), B), GRID_THREADS, 0, st>>>(w.hdr, w.stride, w.start, w.rowx);
    DVM_LAUNCH_CHECK();
    const bool self = (Q == R) && (N == M);
    const int* qorder = nullptr;
    if (!self) {
        DVM_CUDA(cudaMemsetAsync(w.qstart, 0, (size_t)B * w.stride * sizeof(int), st));
        dim3 gq(ceil_div(N, GRID_THREADS), B);
        grid_qcount_kernel<<<gq, GRID_THREADS, 0, st>>>(Q, N, w.hdr, w.stride, w.qcell, w.qstart);
        DVM_LAUNCH_CHECK();
        ScanArgs sq{w.hdr, w.stride, w.qstart, w.qcursor, GRID_COARSE_MAX + 1, w.cstart, w.bsum, w.nblk_max};
        grid_scan_a_kernel<<<dim3(w.nblk_max, B, 1), 1024, 0, st>>>(sq, B);
        DVM_LAUNCH_CHECK();
        grid_scan_b_kernel<<<dim3(B, 1), 1024, 0, st>>>(sq, B);
        DVM_LAUNCH_CHECK();
        grid_scan_c_kernel<<<dim3(w.nblk_max, B, 1), 1024, 0, st>>>(sq, B);
        DVM_LAUNCH_CHECK();
        grid_qscatter_kernel<<<gq, GRID_THREADS, 0, st>>>(N, w.stride, w.qcell, w.qcursor, w.qorder);
        DVM_LAUNCH_CHECK();
        qorder = w.qorder;
    }
#define DVM_GRID_DISPATCH(T)                                                                         \
    /* lanes per Chamfer query: 8 for big clouds, 4 below 16k points (measured, both directions + grid builds, ms: 4 x 50k 0.333    \
       vs 0.354; 32 x 5k 0.279 vs 0.245) */                                                                                      \
    if (k == 1) { if (N <= 16384) launch_query<T, 1, 4>(Q, self, qorder, B, N, M, k, w, idx64, idx32, d2f, d2d, st);             \
                  else            launch_query<T, 1, 8>(Q, self, qorder, B, N, M, k, w, idx64, idx32, d2f, d2d, st); }           \
    else if (k <= 4)  launch_query<T, 4>(Q, self, qorder, B, N, M, k, w, idx64, idx32, d2f, d2d, st);               \
    else if (k <= 10) launch_query<T, 10>(Q, self, qorder, B, N, M, k, w, idx64, idx32, d2f, d2d, st);              \
    else              launch_query<T, 16>(Q, self, qorder, B, N, M, k, w, idx64, idx32, d2f, d2d, st);
    if (f64) { DVM_GRID_DISPATCH(double) } else { DVM_GRID_DISPATCH(float) }
#undef DVM_GRID_DISPATCH
    DVM_LAUNCH_CHECK();
    return 0;
}

}  // namespace dvm
