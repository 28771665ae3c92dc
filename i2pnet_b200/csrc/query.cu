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

extern "C" {

int i2p_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                   int32_t *idx, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(b >= 0 && n >= 0 && m >= 0 && nsample >= 1, "ball_query: bad sizes");
    I2P_REQUIRE(b <= 65535, "ball_query: batch > 65535");
    if (b == 0 || m == 0 || n == 0) return I2P_OK;
    const float radius2 = radius * radius;  // ball_query_gpu.cu:24
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
    dim3 grid(ceil_div(n, KNN_WARPS * KNN_QPW), b);
    knn_kernel<false, int32_t><<<grid, Q_THREADS, 0, as_stream(stream)>>>(m, n, k, known, unknown, idx, dist2);
    return check_launch("knn");
}
}
