// Neighbour queries: ball query, 3-NN, k-NN.
//
// Reference kernels: ball_query_kernel_fast (pointnet2/src/ball_query_gpu.cu:9-45),
// three_nn_kernel_fast (pointnet2/src/interpolate_gpu.cu:9-52) -- one thread per query, each
// streaming all N candidates from global memory with stride-3 scalar loads -- and the
// matmul + topk formulation of knn_point (src/projectPN/utils.py:344-380), which
// materialises the full (B,S,N) distance matrix in HBM.
// This design: candidates are staged once per block into shared memory as float4 tiles by
// coalesced loads and consumed by LDS.128 broadcasts (ball query, 3-NN: one thread per query,
// block-wide early exit) or conflict-free per-lane reads (k-NN: one warp per query, the
// running k-best list distributed over the lanes, ballot + shuffle insertion).  Nothing but
// the inputs and the final indices touches HBM.
#include "common.cuh"

namespace i2p {

constexpr int Q_THREADS = 256;
constexpr int Q_TILE = 1024;  // candidates per shared-memory tile (16 KB as float4)

// Cooperative, coalesced load of candidates [base, base+count) of a (N,3) array into a float4
// tile; .w receives |x|^2 in the (x^2 + y^2) + z^2 order of torch.sum(x**2, -1) when WITH_NORM.
template <bool WITH_NORM>
__device__ __forceinline__ void load_tile(float4 *tile, const float *pts, int base, int count) {
    float *flat = reinterpret_cast<float *>(tile);
    const float *src = pts + (size_t)base * 3;
    for (int f = threadIdx.x; f < count * 3; f += blockDim.x) {
        const int p = f / 3, c = f - p * 3;
        flat[p * 4 + c] = __ldg(src + f);
    }
    if (WITH_NORM) {
        __syncthreads();
        for (int p = threadIdx.x; p < count; p += blockDim.x) {
            const float4 v = tile[p];
            tile[p].w = __fadd_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)), __fmul_rn(v.z, v.z));
        }
    }
}

// ---------------------------------------------------------------------------------------
// ball query
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(Q_THREADS) ball_query_kernel(int n, int m, float radius2, int nsample,
                                                               const float *__restrict__ new_xyz,
                                                               const float *__restrict__ xyz,
                                                               int32_t *__restrict__ idx) {
    __shared__ float4 tile[Q_TILE];
    const int b = blockIdx.y;
    const int q = blockIdx.x * Q_THREADS + threadIdx.x;
    const bool active = q < m;
    const float *pts = xyz + (size_t)b * n * 3;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        const float *c = new_xyz + ((size_t)b * m + q) * 3;
        qx = c[0]; qy = c[1]; qz = c[2];
    }
    int32_t *out = idx + ((size_t)b * m + (active ? q : 0)) * nsample;
    int cnt = 0, first = 0;
    for (int base = 0; base < n; base += Q_TILE) {
        const int count = min(Q_TILE, n - base);
        __syncthreads();
        load_tile<false>(tile, pts, base, count);
        __syncthreads();
        if (active && cnt < nsample) {
            for (int i = 0; i < count; ++i) {
                const float4 p = tile[i];
                const float d2 = sqlen(__fsub_rn(qx, p.x), __fsub_rn(qy, p.y), __fsub_rn(qz, p.z));  // :33
                if (d2 < radius2) {  // :34 strict
                    if (cnt == 0) first = base + i;
                    out[cnt] = base + i;
                    if (++cnt >= nsample) break;  // :42
                }
            }
        }
        if (__syncthreads_and(!active || cnt >= nsample)) break;
    }
    // :35-39 on the first hit the reference pre-fills every slot with it; the later hits then
    // overwrite slots [1, cnt) -- same final content as filling the tail here
    if (active && cnt > 0)
        for (int l = cnt; l < nsample; ++l) out[l] = first;
}

// Warp-cooperative form for small clouds: one thread per query leaves the GPU almost empty when B * M is a few
// thousand (N <= 8192 in the operator sweep: 2 warps per SM).  Here a warp owns BQ_QPW queries and its 32 lanes test 32
// consecutive candidates of the shared tile at once; the lanes inside the radius are compacted in index order with a
// ballot (rank = popcount of the lower lanes), so the slots fill exactly as the reference's sequential scan fills them.
// Default for N <= 32768 (measured, see i2p_ball_query); I2P_BALL_WARP=0 / 1 forces the thread / warp form.
constexpr int BQ_QPW = 4;

__global__ void __launch_bounds__(Q_THREADS) ball_query_warp_kernel(int n, int m, float radius2, int nsample,
                                                                    const float *__restrict__ new_xyz,
                                                                    const float *__restrict__ xyz,
                                                                    int32_t *__restrict__ idx) {
    __shared__ float4 tile[Q_TILE];
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q0 = (blockIdx.x * (Q_THREADS / 32) + warp) * BQ_QPW;
    const float *pts = xyz + (size_t)b * n * 3;
    float qx[BQ_QPW], qy[BQ_QPW], qz[BQ_QPW];
    int cnt[BQ_QPW], first[BQ_QPW];
#pragma unroll
    for (int j = 0; j < BQ_QPW; ++j) {
        const bool valid = q0 + j < m;
        const float *c = new_xyz + ((size_t)b * m + (valid ? q0 + j : 0)) * 3;
        qx[j] = c[0]; qy[j] = c[1]; qz[j] = c[2];
        cnt[j] = valid ? 0 : nsample;          // a padding query counts as finished
        first[j] = 0;
    }
    const unsigned lower = (1u << lane) - 1u;
    for (int base = 0; base < n; base += Q_TILE) {
        const int count = min(Q_TILE, n - base);
        __syncthreads();
        load_tile<false>(tile, pts, base, count);
        __syncthreads();
        bool done = true;
#pragma unroll
        for (int j = 0; j < BQ_QPW; ++j) done = done && cnt[j] >= nsample;
        for (int i0 = 0; i0 < count && !done; i0 += 32) {
            const int i = i0 + lane;
            const bool in = i < count;
            const float4 p = tile[in ? i : 0];
            done = true;
#pragma unroll
            for (int j = 0; j < BQ_QPW; ++j) {
                if (cnt[j] >= nsample) continue;                 // warp-uniform
                const float d2 = sqlen(__fsub_rn(qx[j], p.x), __fsub_rn(qy[j], p.y), __fsub_rn(qz[j], p.z));  // :33
                const bool hit = in && d2 < radius2;             // :34 strict
                const unsigned mask = __ballot_sync(FULL, hit);
                if (mask != 0u) {
                    if (cnt[j] == 0) first[j] = base + i0 + __ffs(mask) - 1;
                    const int slot = cnt[j] + __popc(mask & lower);
                    if (hit && slot < nsample) idx[((size_t)b * m + q0 + j) * nsample + slot] = base + i;
                    cnt[j] = min(nsample, cnt[j] + __popc(mask));
                }
                done = done && cnt[j] >= nsample;
            }
        }
        if (__syncthreads_and(done)) break;
    }
    // :35-39 the slots behind the last hit hold the first hit; a query without any hit leaves the caller's zeros
#pragma unroll
    for (int j = 0; j < BQ_QPW; ++j) {
        if (q0 + j >= m || cnt[j] == 0) continue;
        for (int l = cnt[j] + lane; l < nsample; l += 32) idx[((size_t)b * m + q0 + j) * nsample + l] = first[j];
    }
}

// ---------------------------------------------------------------------------------------
// three nearest neighbours
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(Q_THREADS) three_nn_kernel(int n, int m, const float *__restrict__ unknown,
                                                             const float *__restrict__ known,
                                                             float *__restrict__ dist2, int32_t *__restrict__ idx) {
    __shared__ float4 tile[Q_TILE];
    const int b = blockIdx.y;
    const int q = blockIdx.x * Q_THREADS + threadIdx.x;
    const bool active = q < n;
    const float *pts = known + (size_t)b * m * 3;
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (active) {
        const float *c = unknown + ((size_t)b * n + q) * 3;
        ux = c[0]; uy = c[1]; uz = c[2];
    }
    // The reference keeps best* as double initialised to 1e40 and compares the f32 distance
    // against it (:30,:37-46); +inf in f32 orders identically and converts to the same output.
    float best1 = INFINITY, best2 = INFINITY, best3 = INFINITY;
    int i1 = 0, i2 = 0, i3 = 0;
    for (int base = 0; base < m; base += Q_TILE) {
        const int count = min(Q_TILE, m - base);
        __syncthreads();
        load_tile<false>(tile, pts, base, count);
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int i = 0; i < count; ++i) {
                const float4 p = tile[i];
                const float d = sqlen(__fsub_rn(ux, p.x), __fsub_rn(uy, p.y), __fsub_rn(uz, p.z));  // :36
                const int k = base + i;
                if (d < best1) {
                    best3 = best2; i3 = i2; best2 = best1; i2 = i1; best1 = d; i1 = k;
                } else if (d < best2) {
                    best3 = best2; i3 = i2; best2 = d; i2 = k;
                } else if (d < best3) {
                    best3 = d; i3 = k;
                }
            }
        }
    }
    if (active) {
        float *od = dist2 + ((size_t)b * n + q) * 3;
        int32_t *oi = idx + ((size_t)b * n + q) * 3;
        od[0] = best1; od[1] = best2; od[2] = best3;
        oi[0] = i1; oi[1] = i2; oi[2] = i3;
    }
}

// ---------------------------------------------------------------------------------------
// k nearest neighbours, k <= 32: one warp per query, sorted k-best list across the lanes
// ---------------------------------------------------------------------------------------
constexpr int KNN_WARPS = Q_THREADS / 32;
constexpr int KNN_QPW = 4;  // queries per warp, so a staged tile serves 32 queries per block

// GEMM_FORM: dist = (-2 (q.x) + |q|^2) + |x|^2 as square_distance builds it
// (src/projectPN/utils.py:344-365: matmul, then two in-place adds); otherwise the direct
// (q-x)^2 form of the pointnet2 kernels.
template <bool GEMM_FORM, typename IdxT>
__global__ void __launch_bounds__(Q_THREADS) knn_kernel(int n, int s, int k, const float *__restrict__ xyz,
                                                        const float *__restrict__ new_xyz,
                                                        IdxT *__restrict__ idx_out, float *__restrict__ dist_out) {
    __shared__ float4 tile[Q_TILE];
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q0 = (blockIdx.x * KNN_WARPS + warp) * KNN_QPW;
    const float *pts = xyz + (size_t)b * n * 3;

    float qx[KNN_QPW], qy[KNN_QPW], qz[KNN_QPW], qn[KNN_QPW];
    float ld[KNN_QPW], thresh[KNN_QPW];  // lane i holds the i-th smallest distance so far
    int li[KNN_QPW];
#pragma unroll
    for (int u = 0; u < KNN_QPW; ++u) {
        const int q = min(q0 + u, s - 1);
        const float *c = new_xyz + ((size_t)b * s + q) * 3;
        qx[u] = c[0]; qy[u] = c[1]; qz[u] = c[2];
        qn[u] = __fadd_rn(__fadd_rn(__fmul_rn(qx[u], qx[u]), __fmul_rn(qy[u], qy[u])), __fmul_rn(qz[u], qz[u]));
        ld[u] = INFINITY; thresh[u] = INFINITY; li[u] = 0;
    }

    for (int base = 0; base < n; base += Q_TILE) {
        const int count = min(Q_TILE, n - base);
        __syncthreads();
        load_tile<GEMM_FORM>(tile, pts, base, count);
        __syncthreads();
#pragma unroll
        for (int u = 0; u < KNN_QPW; ++u) {
            if (q0 + u >= s) continue;  // warp-uniform
            for (int i0 = 0; i0 < count; i0 += 32) {
                const int i = i0 + lane;
                float d = INFINITY;
                if (i < count) {
                    const float4 p = tile[i];
                    if (GEMM_FORM) {
                        const float dot = __fmaf_rn(qz[u], p.z, __fmaf_rn(qy[u], p.y, __fmul_rn(qx[u], p.x)));
                        d = __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, dot), qn[u]), p.w);
                    } else {
                        d = sqlen(__fsub_rn(qx[u], p.x), __fsub_rn(qy[u], p.y), __fsub_rn(qz[u], p.z));
                    }
                }
                unsigned cand = __ballot_sync(FULL, d < thresh[u]);
                while (cand) {  // ascending candidate index: earlier index wins distance ties
                    const int src = __ffs(cand) - 1;
                    cand &= cand - 1;
                    const float dn = __shfl_sync(FULL, d, src);
                    if (!(dn < thresh[u])) continue;  // the threshold moved since the ballot
                    const int in = base + i0 + src;
                    const int pos = __popc(__ballot_sync(FULL, ld[u] <= dn));  // after equal entries
                    const float up_d = __shfl_up_sync(FULL, ld[u], 1);
                    const int up_i = __shfl_up_sync(FULL, li[u], 1);
                    if (lane == pos) { ld[u] = dn; li[u] = in; }
                    else if (lane > pos) { ld[u] = up_d; li[u] = up_i; }
                    thresh[u] = __shfl_sync(FULL, ld[u], k - 1);
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < KNN_QPW; ++u) {
        const int q = q0 + u;
        if (q < s && lane < k) {
            const size_t o = ((size_t)b * s + q) * k + lane;
            idx_out[o] = (IdxT)li[u];
            if (dist_out != nullptr) dist_out[o] = ld[u];
        }
    }
}

}  // namespace i2p

namespace i2p {
// ---------------------------------------------------------------------------------------
// neighbour queries over a uniform grid (cell list): O(N + M * candidates) instead of O(N * M)
// ---------------------------------------------------------------------------------------
// The reference kernels scan all N points per query.  For larger clouds the points are binned into the cells of a
// uniform grid over the cloud's bounding box (counting sort: bounding box, count, scan, scatter -- four small kernels;
// the sorted copy carries the original index in .w), and a query visits only nearby cells:
//   ball query  cells no smaller than 1.002 r, the 27 cells around the query's own, keeping the nsample SMALLEST
//               indices inside the radius -- the reference's "first nsample in index order";
//   k-NN, 3-NN  shells of cells of growing Chebyshev radius rho around the query's cell until the k-th best distance
//               is below ((rho - 0.01) * smallest cell width)^2: every point not yet visited is further than that.
//               The k best are kept ordered by (distance, index), the order of the brute-force scans.
// Distances are computed by the same expression as the brute-force kernels, so outputs are bit-identical.  The 0.2 % /
// 1 % margins cover the rounding of the cell computation (at most ~1e-4 of a cell: up to 1024 cells per axis in f32).
constexpr int BG_MAXDIM = 1024, BG_MAXCELLS = 1 << 18;
constexpr int BG_MAXNS = 64;    // ball query: nsample kept per thread
constexpr int BG_MAXK = 32;

struct GridBox { unsigned lo[3], hi[3]; };     // order-preserving unsigned images of the float bounds (lo: its complement, see bg_bbox_kernel)

__device__ __forceinline__ unsigned f2ord(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

struct GridDims {
    float lo[3], hi[3], inv[3];
    float cs_min;        // smallest cell width over the axes with more than one cell (+inf when there is one cell)
    int n[3], ncell;
};

__device__ __forceinline__ int grid_coord(const GridDims &g, int d, float v) {
    const int c = (int)floorf((v - g.lo[d]) * g.inv[d]);
    return c < 0 ? 0 : (c >= g.n[d] ? g.n[d] - 1 : c);
}

// Bounding box of every cloud.  box is zero-initialised by the memset that clears the counts: lo holds the COMPLEMENT of the
// ordered image, so that both bounds are running maxima over zero.  One atomic per bound and block.
__global__ void __launch_bounds__(256) bg_bbox_kernel(int n, const float *__restrict__ xyz, GridBox *box) {
    __shared__ unsigned red[8][6];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *pts = xyz + (size_t)b * n * 3;
    unsigned v[6] = {0u, 0u, 0u, 0u, 0u, 0u};     // ~lo[3], hi[3]
    // the cloud as a flat array of 3n floats, coalesced: element f belongs to axis f mod 3
    for (int f0 = blockIdx.x * 256 * 3; f0 < 3 * n; f0 += gridDim.x * 256 * 3) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int f = f0 + j * 256 + threadIdx.x;
            if (f < 3 * n) {
                const unsigned u = f2ord(__ldg(pts + f));
                const int d = (j * 256 + threadIdx.x) % 3;      // f0 is a multiple of 3
                if (d == 0) { v[0] = max(v[0], ~u); v[3] = max(v[3], u); }
                else if (d == 1) { v[1] = max(v[1], ~u); v[4] = max(v[4], u); }
                else { v[2] = max(v[2], ~u); v[5] = max(v[5], u); }
            }
        }
    }
#pragma unroll
    for (int d = 0; d < 6; ++d) {
        v[d] = __reduce_max_sync(FULL, v[d]);
        if (lane == 0) red[warp][d] = v[d];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        unsigned m = 0u;
#pragma unroll
        for (int w = 0; w < 8; ++w) m = max(m, red[w][threadIdx.x]);
        unsigned *dst = threadIdx.x < 3 ? &box[b].lo[threadIdx.x] : &box[b].hi[threadIdx.x - 3];
        atomicMax(dst, m);
    }
}

// Cell sizes of one cloud: cells at least `min_cell` wide (ball query: 1.001 r) and, when per_cell > 0, sized for that
// many points per cell on average (k-NN); at most BG_MAXDIM per axis and max_cells in all.
__device__ GridDims grid_dims(const GridBox &box, int n, int max_cells, float min_cell, float per_cell) {
    GridDims g;
    float ext[3], measure = 1.f;
    int live = 0;
    for (int d = 0; d < 3; ++d) {
        g.lo[d] = ord2f(~box.lo[d]);
        g.hi[d] = ord2f(box.hi[d]);
        ext[d] = g.hi[d] - g.lo[d];
        if (ext[d] > 0.f && isfinite(ext[d])) { measure *= ext[d]; ++live; } else ext[d] = 0.f;
    }
    float cs = min_cell;
    if (per_cell > 0.f && live > 0) cs = fmaxf(cs, powf(measure * per_cell / (float)n, 1.f / (float)live));
    if (!(cs > 0.f)) cs = 1.f;
    for (int it = 0; it < 64; ++it) {
        long long prod = 1;
        for (int d = 0; d < 3; ++d) {
            int c = ext[d] > 0.f ? (int)fminf(ext[d] / cs, (float)BG_MAXDIM) : 1;
            g.n[d] = c < 1 ? 1 : c;
            prod *= g.n[d];
        }
        if (prod <= max_cells) break;
        cs *= 1.26f;
    }
    g.cs_min = INFINITY;
    for (int d = 0; d < 3; ++d) {
        g.inv[d] = ext[d] > 0.f ? (float)g.n[d] / ext[d] : 0.f;
        if (g.n[d] > 1) g.cs_min = fminf(g.cs_min, ext[d] / (float)g.n[d]);
    }
    g.ncell = g.n[0] * g.n[1] * g.n[2];
    return g;
}

// cell of every point + histogram; counts (b, max_cells + 1) pre-zeroed.  Every block derives the cloud's grid from the
// bounding box (one thread, a few hundred cycles); block 0 of each cloud records it for the kernels that follow.
__global__ void __launch_bounds__(256) bg_count_kernel(int n, int max_cells, float min_cell, float per_cell, const float *__restrict__ xyz,
                                                      const GridBox *box, GridDims *dims, int *__restrict__ cell_of, int *counts) {
    __shared__ GridDims gs;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) {
        if (box != nullptr) {
            gs = grid_dims(box[b], n, max_cells, min_cell, per_cell);
            if (blockIdx.x == 0) dims[b] = gs;
        } else {
            gs = dims[b];        // a second point set (the queries) binned over an existing grid
        }
    }
    __syncthreads();
    const GridDims g = gs;
    const float *pts = xyz + (size_t)b * n * 3;
    for (int k = blockIdx.x * 256 + threadIdx.x; k < n; k += gridDim.x * 256) {
        const int cx = grid_coord(g, 0, __ldg(pts + (size_t)k * 3)), cy = grid_coord(g, 1, __ldg(pts + (size_t)k * 3 + 1)),
                  cz = grid_coord(g, 2, __ldg(pts + (size_t)k * 3 + 2));
        const int c = (cz * g.n[1] + cy) * g.n[0] + cx;
        cell_of[(size_t)b * n + k] = c;
        atomicAdd(counts + (size_t)b * (max_cells + 1) + c, 1);
    }
}

// Exclusive scan of the cell counts: counts -> starts (in place, starts[ncell] = n), fill = a copy of the starts for the
// scatter.  4096 cells per block; a block takes its chunk from a ticket counter (so that every chunk it waits for is
// already running), publishes the chunk's total and sums the totals of the chunks before it (at most 64 of them).
constexpr int BG_CHUNK = 4096, BG_MAXCHUNKS = BG_MAXCELLS / BG_CHUNK;
struct ScanState { unsigned ticket; int total[BG_MAXCHUNKS]; };    // zero-initialised; total holds sum + 1 once published

__global__ void __launch_bounds__(1024) bg_scan_kernel(int max_cells, const GridDims *dims, int *counts, int *fill, ScanState *state) {
    __shared__ int wsum[32];
    __shared__ int chunk_s, before_s;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncell = dims[b].ncell;
    ScanState &st = state[b];
    if (tid == 0) chunk_s = (int)atomicAdd(&st.ticket, 1u);
    __syncthreads();
    const int chunk = chunk_s, base = chunk * BG_CHUNK;
    if (base >= ncell) return;
    int *cnt = counts + (size_t)b * (max_cells + 1), *fl = fill + (size_t)b * max_cells;
    const int c0 = base + tid * 4;
    int v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = c0 + j < ncell ? cnt[c0 + j] : 0;
    const int local = v[0] + v[1] + v[2] + v[3];
    int incl = local;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(FULL, incl, off); if (lane >= off) incl += t; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int w = wsum[lane];
        int wi = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(FULL, wi, off); if (lane >= off) wi += t; }
        wsum[lane] = wi - w;     // exclusive prefix of the warp sums
        if (lane == 31) atomicExch(&st.total[chunk], wi + 1);        // publish this chunk's total
        int before = 0;
        for (int j = lane; j < chunk; j += 32) {
            int t;
            while ((t = atomicAdd(&st.total[j], 0)) == 0) __nanosleep(20);
            before += t - 1;
        }
        before = __reduce_add_sync(FULL, before);
        if (lane == 0) before_s = before;
    }
    __syncthreads();
    int run = before_s + wsum[warp] + incl - local;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (c0 + j < ncell) { cnt[c0 + j] = run; fl[c0 + j] = run; }
        run += v[j];
        if (c0 + j == ncell - 1) cnt[ncell] = run;
    }
}

__global__ void __launch_bounds__(256) bg_scatter_kernel(int n, int max_cells, const float *__restrict__ xyz, const int *__restrict__ cell_of,
                                                        int *fill, float4 *__restrict__ sorted) {
    const int b = blockIdx.y;
    const float *pts = xyz + (size_t)b * n * 3;
    for (int k = blockIdx.x * 256 + threadIdx.x; k < n; k += gridDim.x * 256) {
        const int c = cell_of[(size_t)b * n + k];
        const int pos = atomicAdd(fill + (size_t)b * max_cells + c, 1);
        sorted[(size_t)b * n + pos] = make_float4(__ldg(pts + (size_t)k * 3), __ldg(pts + (size_t)k * 3 + 1), __ldg(pts + (size_t)k * 3 + 2),
                                                  __int_as_float(k));
    }
}

// A point set binned over a grid, in stream-ordered scratch memory
struct CellList {
    unsigned char *ws = nullptr;
    GridDims *dims = nullptr;
    int max_cells = 0;           // no more cells than twice the points (and at most BG_MAXCELLS)
    int *starts = nullptr;       // (b, max_cells + 1)
    float4 *sorted = nullptr;    // (b, n): x, y, z, original index
};

// over == nullptr: grid from the bounding box of xyz; otherwise xyz (the queries) is binned over that list's grid, which
// puts queries of one cell -- same candidate rows -- next to each other.
static int cell_list_build(CellList &cl, int b, int n, const float *xyz, float min_cell, float per_cell, const CellList *over,
                           cudaStream_t s, const char *what) {
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    // zero-initialised head: bounding boxes | scan state | counts (b, max_cells + 1); then dims | fill | cell_of | sorted
    const int mc = cl.max_cells = over != nullptr ? over->max_cells : (2 * n < 64 ? 64 : (2 * n < BG_MAXCELLS ? 2 * n : BG_MAXCELLS));
    const size_t box_b = al(sizeof(GridBox) * b), state_b = al(sizeof(ScanState) * b), cnt_b = al(sizeof(int) * (size_t)b * (mc + 1));
    const size_t dims_b = al(sizeof(GridDims) * b), fill_b = al(sizeof(int) * (size_t)b * mc);
    const size_t cell_b = al(sizeof(int) * (size_t)b * n), sort_b = sizeof(float4) * (size_t)b * n;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&cl.ws), box_b + state_b + cnt_b + dims_b + fill_b + cell_b + sort_b, s);
    if (e != cudaSuccess) { set_error("%s (cell list): %s", what, cudaGetErrorString(e)); return I2P_ERR_CUDA; }
    unsigned char *p = cl.ws;
    GridBox *box = reinterpret_cast<GridBox *>(p); p += box_b;
    ScanState *state = reinterpret_cast<ScanState *>(p); p += state_b;
    cl.starts = reinterpret_cast<int *>(p); p += cnt_b;
    cl.dims = over != nullptr ? over->dims : reinterpret_cast<GridDims *>(p); p += dims_b;
    int *fill = reinterpret_cast<int *>(p); p += fill_b;
    int *cell_of = reinterpret_cast<int *>(p); p += cell_b;
    cl.sorted = reinterpret_cast<float4 *>(p);
    cudaMemsetAsync(cl.ws, 0, box_b + state_b + cnt_b, s);
    const int gx = ceil_div(n, 256) < 148 ? ceil_div(n, 256) : 148;
    if (over == nullptr) bg_bbox_kernel<<<dim3(ceil_div(n, 256 * 4) < 148 ? ceil_div(n, 256 * 4) : 148, b), 256, 0, s>>>(n, xyz, box);
    bg_count_kernel<<<dim3(gx, b), 256, 0, s>>>(n, mc, min_cell, per_cell, xyz, over == nullptr ? box : nullptr, cl.dims, cell_of, cl.starts);
    bg_scan_kernel<<<dim3(ceil_div(mc, BG_CHUNK), b), 1024, 0, s>>>(mc, cl.dims, cl.starts, fill, state);
    bg_scatter_kernel<<<dim3(gx, b), 256, 0, s>>>(n, mc, xyz, cell_of, fill, cl.sorted);
    return I2P_OK;
}

// Queries visited in cell order: lanes of a warp then read the same few candidate rows (L1 hits, broadcast loads) instead of
// 32 unrelated places of L2.  Binning the queries costs ~25 us; it pays once queries * k^2 >= 2^19 (k-NN from 2048 queries
// at k = 16, 3-NN from ~60k; never for the ball query's 27 cells: profiles/r2_sweep_ops.md).  I2P_CELL_SORT_QUERIES=0 / 1 forces.
static bool sort_queries(int queries, int k) {
    static int form = -1;
    if (form < 0) { const char *e = getenv("I2P_CELL_SORT_QUERIES"); form = e ? atoi(e) : 2; }
    return form == 1 || (form == 2 && k > 0 && (long long)queries * k * k >= (1ll << 19));
}

__global__ void __launch_bounds__(128) bg_ball_kernel(int n, int m, int max_cells, float radius, float radius2, int nsample,
                                                     const float *__restrict__ new_xyz, const GridDims *dims,
                                                     const int *__restrict__ starts, const float4 *__restrict__ sorted,
                                                     const float4 *__restrict__ qsorted, int32_t *__restrict__ idx) {
    const int b = blockIdx.y, t = blockIdx.x * 128 + threadIdx.x;
    if (t >= m) return;
    const GridDims g = dims[b];
    int q = t;
    float qv[3];
    if (qsorted != nullptr) {
        const float4 v = qsorted[(size_t)b * m + t];
        qv[0] = v.x; qv[1] = v.y; qv[2] = v.z; q = __float_as_int(v.w);
    } else {
        const float *c = new_xyz + ((size_t)b * m + q) * 3;
        qv[0] = c[0]; qv[1] = c[1]; qv[2] = c[2];
    }
    const int *st = starts + (size_t)b * (max_cells + 1);
    const float4 *pts = sorted + (size_t)b * n;
    int best[BG_MAXNS];       // the nsample smallest indices inside the radius, ascending
    int cnt = 0;
    bool any = true;          // a query further than the radius outside the bounding box has no neighbour at all
    int c0[3], c1[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (qv[d] < g.lo[d] - 1.001f * radius || qv[d] > g.hi[d] + 1.001f * radius) any = false;
        const int cc = grid_coord(g, d, qv[d]);
        c0[d] = max(cc - 1, 0);
        c1[d] = min(cc + 1, g.n[d] - 1);
    }
    if (any) {
        for (int cz = c0[2]; cz <= c1[2]; ++cz)
            for (int cy = c0[1]; cy <= c1[1]; ++cy) {
                // the cells of one x-run are consecutive: one contiguous range of the sorted array
                const int cb = (cz * g.n[1] + cy) * g.n[0];
                const int pb = st[cb + c0[0]], pe = st[cb + c1[0] + 1];
                for (int p = pb; p < pe; ++p) {
                    const float4 v = __ldg(pts + p);
                    const float d2 = sqlen(__fsub_rn(qv[0], v.x), __fsub_rn(qv[1], v.y), __fsub_rn(qv[2], v.z));   // ball_query_gpu.cu:33
                    if (d2 < radius2) {                                                                          // :34 strict
                        const int k = __float_as_int(v.w);
                        if (cnt < nsample || k < best[cnt - 1]) {
                            int i = cnt < nsample ? cnt++ : nsample - 1;
                            while (i > 0 && best[i - 1] > k) { best[i] = best[i - 1]; --i; }
                            best[i] = k;
                        }
                    }
                }
            }
    }
    if (cnt > 0) {
        int32_t *out = idx + ((size_t)b * m + q) * nsample;
        for (int l = 0; l < nsample; ++l) out[l] = l < cnt ? best[l] : best[0];     // :35-39: unfilled slots hold the first hit
    }
}

static int ball_query_grid(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int32_t *idx,
                           cudaStream_t s) {
    CellList cl, ql;
    int rc = cell_list_build(cl, b, n, xyz, radius * 1.002f, 0.f, nullptr, s, "ball_query");
    if (rc != I2P_OK) return rc;
    if (sort_queries(m, 0) && (rc = cell_list_build(ql, b, m, new_xyz, 0.f, 0.f, &cl, s, "ball_query")) != I2P_OK) { cudaFreeAsync(cl.ws, s); return rc; }
    bg_ball_kernel<<<dim3(ceil_div(m, 128), b), 128, 0, s>>>(n, m, cl.max_cells, radius, radius * radius, nsample, new_xyz, cl.dims, cl.starts,
                                                             cl.sorted, ql.sorted, idx);
    rc = check_launch("ball_query(grid)");
    cudaFreeAsync(cl.ws, s);
    if (ql.ws != nullptr) cudaFreeAsync(ql.ws, s);
    return rc;
}

// k nearest of `cand` (b, n, 3) for every query of `qry` (b, s, 3), direct (q - x)^2 distances, ordered by (distance, index).
// Slots beyond the number of candidates keep (+inf, 0) -- the 3-NN kernel's result for fewer than three known points.
template <int KMAX>
__global__ void __launch_bounds__(128, 4) bg_knn_kernel(int n, int s, int k, int max_cells, const float *__restrict__ qry, const GridDims *dims,
                                                    const int *__restrict__ starts, const float4 *__restrict__ sorted,
                                                    const float4 *__restrict__ qsorted, int32_t *__restrict__ idx_out,
                                                    float *__restrict__ dist_out) {
    const int b = blockIdx.y, t = blockIdx.x * 128 + threadIdx.x;
    if (t >= s) return;
    const GridDims g = dims[b];
    int q = t;
    float qx, qy, qz;
    if (qsorted != nullptr) {
        const float4 v = qsorted[(size_t)b * s + t];
        qx = v.x; qy = v.y; qz = v.z; q = __float_as_int(v.w);
    } else {
        const float *c = qry + ((size_t)b * s + q) * 3;
        qx = c[0]; qy = c[1]; qz = c[2];
    }
    const int *st = starts + (size_t)b * (max_cells + 1);
    const float4 *pts = sorted + (size_t)b * n;
    // the KMAX best as a sorted list in registers (every index below is a compile-time constant): lists in local memory
    // indexed per lane cost one L1 wavefront per lane and access, which bounded the first version of this kernel
    float bd[KMAX];
    int bi[KMAX];
#pragma unroll
    for (int i = 0; i < KMAX; ++i) { bd[i] = INFINITY; bi[i] = 0x7fffffff; }
    float kth = INFINITY;      // the k-th best so far: (distance, index) a candidate has to beat
    int kth_i = 0x7fffffff;
    const int cx = grid_coord(g, 0, qx), cy = grid_coord(g, 1, qy), cz = grid_coord(g, 2, qz);
    const int reach = max(max(max(cx, g.n[0] - 1 - cx), max(cy, g.n[1] - 1 - cy)), max(cz, g.n[2] - 1 - cz));
    for (int rho = 0; rho <= reach; ++rho) {
        if (rho > 0) {
            const float safe = ((float)rho - 1.01f) * g.cs_min;      // everything within `safe` has been visited after shell rho - 1
            if (safe > 0.f && kth < safe * safe) break;
        }
        const int z0 = max(cz - rho, 0), z1 = min(cz + rho, g.n[2] - 1), y0 = max(cy - rho, 0), y1 = min(cy + rho, g.n[1] - 1);
        const int x0 = max(cx - rho, 0), x1 = min(cx + rho, g.n[0] - 1);
        for (int z = z0; z <= z1; ++z)
            for (int y = y0; y <= y1; ++y) {
                const int cb = (z * g.n[1] + y) * g.n[0];
                // a whole row of the shell is one contiguous range of the sorted array; inside the shell's faces only the
                // row's two end cells belong to it
                const bool row = abs(z - cz) == rho || abs(y - cy) == rho;
                for (int part = 0; part < (row ? 1 : 2); ++part) {
                    int pb, pe;
                    if (row) { pb = st[cb + x0]; pe = st[cb + x1 + 1]; }
                    else {
                        const int x = part == 0 ? cx - rho : cx + rho;
                        if (x < 0 || x >= g.n[0]) continue;
                        pb = st[cb + x]; pe = st[cb + x + 1];
                    }
                    for (int p = pb; p < pe; ++p) {
                        const float4 v = __ldg(pts + p);
                        const float d = sqlen(__fsub_rn(qx, v.x), __fsub_rn(qy, v.y), __fsub_rn(qz, v.z));      // interpolate_gpu.cu:36
                        const int id = __float_as_int(v.w);
                        if (d < kth || (d == kth && id < kth_i)) {
                            float cd = d;
                            int ci = id;
#pragma unroll
                            for (int i = 0; i < KMAX; ++i) {       // the displaced entry moves on down the list
                                const bool lt = cd < bd[i] || (cd == bd[i] && ci < bi[i]);
                                const float td = lt ? bd[i] : cd;
                                const int ti = lt ? bi[i] : ci;
                                bd[i] = lt ? cd : bd[i];
                                bi[i] = lt ? ci : bi[i];
                                cd = td; ci = ti;
                            }
#pragma unroll
                            for (int i = 0; i < KMAX; ++i)      // entry k - 1, without a run-time index (which would put the list in local memory)
                                if (i < k) { kth = bd[i]; kth_i = bi[i]; }
                        }
                    }
                }
            }
    }
    const size_t o = ((size_t)b * s + q) * k;
#pragma unroll
    for (int l = 0; l < KMAX; ++l) {
        if (l < k) {
            idx_out[o + l] = bi[l] == 0x7fffffff ? 0 : bi[l];
            if (dist_out != nullptr) dist_out[o + l] = bd[l];
        }
    }
}

static int knn_grid(int b, int n, int s, int k, const float *qry, const float *cand, float *dist2, int32_t *idx, cudaStream_t st,
                    const char *what) {
    CellList cl, ql;
    static float per_cell = -1.f;   // I2P_CELL_POINTS: tuning override of the average points per cell (measurements in profiles/)
    if (per_cell < 0.f) { const char *e = getenv("I2P_CELL_POINTS"); per_cell = e ? (float)atof(e) : 2.f; }
    int rc = cell_list_build(cl, b, n, cand, 0.f, per_cell, nullptr, st, what);
    if (rc != I2P_OK) return rc;
    if (sort_queries(s, k) && (rc = cell_list_build(ql, b, s, qry, 0.f, 0.f, &cl, st, what)) != I2P_OK) { cudaFreeAsync(cl.ws, st); return rc; }
    const dim3 grid(ceil_div(s, 128), b);
    if (k <= 4) bg_knn_kernel<4><<<grid, 128, 0, st>>>(n, s, k, cl.max_cells, qry, cl.dims, cl.starts, cl.sorted, ql.sorted, idx, dist2);
    else if (k <= 8) bg_knn_kernel<8><<<grid, 128, 0, st>>>(n, s, k, cl.max_cells, qry, cl.dims, cl.starts, cl.sorted, ql.sorted, idx, dist2);
    else if (k <= 16) bg_knn_kernel<16><<<grid, 128, 0, st>>>(n, s, k, cl.max_cells, qry, cl.dims, cl.starts, cl.sorted, ql.sorted, idx, dist2);
    else bg_knn_kernel<BG_MAXK><<<grid, 128, 0, st>>>(n, s, k, cl.max_cells, qry, cl.dims, cl.starts, cl.sorted, ql.sorted, idx, dist2);
    rc = check_launch(what);
    cudaFreeAsync(cl.ws, st);
    if (ql.ws != nullptr) cudaFreeAsync(ql.ws, st);
    return rc;
}

// I2P_CELL_LIST: 0 = brute force only, 1 = cell list whenever it applies, 2 (default) = cell list when the brute-force scan
// would compute at least 4M distances over at least 1024 candidates (where the ~40 us of building the list are earned
// back: profiles/r2_sweep_ops.md)
static bool use_cell_list(int queries, int candidates) {
    static int form = -1;
    if (form < 0) { const char *e = getenv("I2P_CELL_LIST"); form = e ? atoi(e) : 2; }
    if (queries < 1 || candidates < 1) return false;
    return form == 1 || (form == 2 && candidates >= 1024 && (long long)queries * candidates >= (4ll << 20));
}
}  // namespace i2p

extern "C" {

int i2p_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                   int32_t *idx, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(b >= 0 && n >= 0 && m >= 0 && nsample >= 1, "ball_query: bad sizes");
    I2P_REQUIRE(b <= 65535, "ball_query: batch > 65535");
    if (b == 0 || m == 0 || n == 0) return I2P_OK;
    const float radius2 = radius * radius;  // ball_query_gpu.cu:24
    // cell list for larger clouds (I2P_CELL_LIST=0 turns it off, =1 forces it); brute force below
    if (nsample <= BG_MAXNS && radius > 0.f && use_cell_list(m, n))
        return ball_query_grid(b, n, m, radius, nsample, new_xyz, xyz, idx, as_stream(stream));
    // The warp-cooperative kernel fills the GPU when there are few queries and wins up to N = 32k in the operator sweep
    // (profiles/r2_sweep_ops.md: 67 vs 279 us at N = 4096, 1746 vs 2235 us at 32k; 6.6 vs 5.4 ms at 65k), the
    // thread-per-query kernel beyond.  I2P_BALL_WARP=0 / 1 forces one of them.
    static int warp_form = -1;
    if (warp_form < 0) { const char *e = getenv("I2P_BALL_WARP"); warp_form = e ? atoi(e) : 2; }
    if (warp_form == 1 || (warp_form == 2 && n <= 32768)) {
        dim3 grid(ceil_div(m, (Q_THREADS / 32) * BQ_QPW), b);
        ball_query_warp_kernel<<<grid, Q_THREADS, 0, as_stream(stream)>>>(n, m, radius2, nsample, new_xyz, xyz, idx);
        return check_launch("ball_query(warp)");
    }
    dim3 grid(ceil_div(m, Q_THREADS), b);
    ball_query_kernel<<<grid, Q_THREADS, 0, as_stream(stream)>>>(n, m, radius2, nsample, new_xyz, xyz, idx);
    return check_launch("ball_query");
}

int i2p_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int32_t *idx,
                 void *stream) {
    using namespace i2p;
    I2P_REQUIRE(b >= 0 && n >= 0 && m >= 0 && b <= 65535, "three_nn: bad sizes");
    if (b == 0 || n == 0) return I2P_OK;
    if (m >= 3 && use_cell_list(n, m)) return knn_grid(b, m, n, 3, unknown, known, dist2, idx, as_stream(stream), "three_nn(grid)");
    dim3 grid(ceil_div(n, Q_THREADS), b);
    three_nn_kernel<<<grid, Q_THREADS, 0, as_stream(stream)>>>(n, m, unknown, known, dist2, idx);
    return check_launch("three_nn");
}

int i2p_knn_point(int b, int n, int s, int nsample, const float *xyz, const float *new_xyz, int64_t *group_idx,
                  float *dist_out, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(b >= 0 && n >= 0 && s >= 0 && nsample >= 1 && b <= 65535, "knn_point: bad sizes");
    I2P_REQUIRE(nsample <= n, "knn_point: nsample=%d > n=%d (torch.topk raises here too)", nsample, n);
    if (nsample > 32) {
        set_error("knn_point: nsample=%d > 32 is not built (the reference's configs use 4..32)", nsample);
        return I2P_ERR_UNSUPPORTED;
    }
    if (b == 0 || s == 0) return I2P_OK;
    dim3 grid(ceil_div(s, KNN_WARPS * KNN_QPW), b);
    knn_kernel<true, int64_t><<<grid, Q_THREADS, 0, as_stream(stream)>>>(n, s, nsample, xyz, new_xyz, group_idx, dist_out);
    return check_launch("knn_point");
}

int i2p_knn(int b, int n, int m, int k, const float *unknown, const float *known, float *dist2, int32_t *idx,
            void *stream) {
    using namespace i2p;
    I2P_REQUIRE(b >= 0 && n >= 0 && m >= 0 && k >= 1 && b <= 65535, "knn: bad sizes");
    I2P_REQUIRE(k <= m, "knn: k=%d > m=%d", k, m);
    if (k > 32) {
        set_error("knn: k=%d > 32 is not built", k);
        return I2P_ERR_UNSUPPORTED;
    }
    if (b == 0 || n == 0) return I2P_OK;
    if (use_cell_list(n, m)) return knn_grid(b, m, n, k, unknown, known, dist2, idx, as_stream(stream), "knn(grid)");
    dim3 grid(ceil_div(n, KNN_WARPS * KNN_QPW), b);
    knn_kernel<false, int32_t><<<grid, Q_THREADS, 0, as_stream(stream)>>>(m, n, k, known, unknown, idx, dist2);
    return check_launch("knn");
}
}
