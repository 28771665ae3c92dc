// Cost-volume glue (src/projectPN/PPBackbone_center.py:354-490, CostVolume.forward) as four kernels.
//
// 1. cv_build: the operand of the first shared-MLP layer,
//        X[b,n,k,:] = [ xyz1[b,n] (3) | xyz2[b,j] (3) | pi[b,n,:] * qi[b,j,:] (C) | maxc[b,j,:] (C, optional) ],
//    j = k (every pixel, cost volume 1) or j = idx[b,n,k] (the point's nearest pixels, cost volume 2), written
//    in one pass from the small per-point / per-pixel tensors.  The reference materialises the two
//    broadcast operands with `repeat`, multiplies, masks (3 more passes), max-reduces, expands and
//    concatenates: ~750 MB of traffic per step for cost volume 1 at batch 8, against 153 MB written here.
//    A second output holds the 6 coordinate channels alone (input of the position encoding).
// 2. cv_build_bwd: dX read once by blocks of 8 points x 16 pixels; every block adds its partial sums (over its
//    pixels for d pi / d xyz1, over its points for d qi / d maxc / d xyz2) with one red.add per element.
// 3. softmax_wsum: out[b,n,c] = sum_k softmax_k(l[b,n,k,c]) v[b,n,k,c] with an optional validity mask
//    (l * m - 1e10 (1 - m), as the reference spells it), replacing softmax + multiply + sum (+ 4 mask ops);
// 4. its backward, which recomputes the softmax from the saved logits:
//        dv = p g,   dl = p g (v - out) [* m].
// All four are HBM-bound streaming kernels: consecutive threads own consecutive channels.
#include <math.h>

#include "common.cuh"

namespace i2p {

constexpr int CVB_NT = 8, CVB_KT = 16;   // points x pixels per block in cv_build_bwd

struct CvGeom {
    int B, N, K, N2, C, has_max, Cx;   // Cx = 6 + C + (has_max ? C : 0)
};

// grid (N, B), block Cx - 6 threads (one per feature channel); threads 0..5 also write the coordinate channels.
// ncu of the first version: 37 instructions per output element (64-bit index arithmetic, three branches per pixel) at
// 42 % issue utilisation -- instruction-bound at 1.35 TB/s.  Here every per-pixel address is a pointer increment, the
// index / every-pixel forms are separate instantiations and the loop is unrolled eight pixels deep.
template <bool IDX>
__global__ void __launch_bounds__(256) cv_build_kernel(CvGeom g, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                                                      const float *__restrict__ pi, const float *__restrict__ qi,
                                                      const float *__restrict__ maxc, const int32_t *__restrict__ idx,
                                                      float *__restrict__ X, float *__restrict__ xyz6) {
    const int n = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
    const bool is_max = t >= g.C;
    const int c = is_max ? t - g.C : t;
    const size_t bn = (size_t)b * g.N + n;
    const float pv = is_max ? 1.f : __ldg(pi + bn * g.C + c);                       // the max branch is copied unscaled
    const float *src = (is_max ? maxc : qi) + (size_t)b * g.N2 * g.C + c;           // + j * C per pixel
    const int32_t *ip = IDX ? idx + bn * g.K : nullptr;
    float *dst = X + bn * g.K * g.Cx + 6 + t;                                       // + Cx per pixel
    const int C = g.C, Cx = g.Cx, K = g.K;
    int k = 0;
    for (; k + 8 <= K; k += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (size_t)(IDX ? __ldg(ip + k + u) : k + u) * C);
#pragma unroll
        for (int u = 0; u < 8; ++u) dst[(size_t)(k + u) * Cx] = is_max ? v[u] : __fmul_rn(pv, v[u]);
    }
    for (; k < K; ++k) {
        const float v = __ldg(src + (size_t)(IDX ? __ldg(ip + k) : k) * C);
        dst[(size_t)k * Cx] = is_max ? v : __fmul_rn(pv, v);
    }
    if (t < 6) {      // the coordinate channels, inside X and as the separate 6-channel output
        const float xv = t < 3 ? __ldg(xyz1 + bn * 3 + t) : 0.f;
        float *row = X + bn * K * Cx + t, *row6 = xyz6 + bn * K * 6 + t;
        const float *x2 = xyz2 + (size_t)b * g.N2 * 3 + (t - 3);
        for (int kk = 0; kk < K; ++kk) {
            const float w = t < 3 ? xv : __ldg(x2 + (size_t)(IDX ? __ldg(ip + kk) : kk) * 3);
            row[(size_t)kk * Cx] = w;
            row6[(size_t)kk * 6] = w;
        }
    }
}

// grid (ceil(N / 8), ceil(K / 16), B), block Cx - 6 threads: 8 points x 16 pixels per block.  dxyz6 (B,N,K,6) is the
// gradient of the separate coordinate output (may be NULL).  Every output is accumulated with atomics (a block
// holds a partial sum over its pixels for d pi / d xyz1 and over its points for d qi / d maxc / d xyz2) and must
// be zero on entry.
__global__ void __launch_bounds__(256) cv_build_bwd_kernel(CvGeom g, const float *__restrict__ dX, const float *__restrict__ dxyz6,
                                                          const float *__restrict__ pi, const float *__restrict__ qi,
                                                          const int32_t *__restrict__ idx, float *dxyz1, float *dxyz2,
                                                          float *dpi, float *dqi, float *dmaxc) {
    const int n0 = blockIdx.x * CVB_NT, k0 = blockIdx.y * CVB_KT, b = blockIdx.z, t = threadIdx.x;
    const int nn = min(CVB_NT, g.N - n0), k1 = min(g.K, k0 + CVB_KT);
    const bool is_max = t >= g.C;
    const int c = is_max ? t - g.C : t;
    float pv[CVB_NT], accp[CVB_NT];
#pragma unroll
    for (int i = 0; i < CVB_NT; ++i) {
        pv[i] = (!is_max && i < nn) ? __ldg(pi + ((size_t)b * g.N + n0 + i) * g.C + c) : 0.f;
        accp[i] = 0.f;
    }
    for (int k = k0; k < k1; ++k) {
        float aq = 0.f;   // partial sum over the block's points for (pixel k, this channel), every-pixel form
        float v[CVB_NT];
#pragma unroll
        for (int i = 0; i < CVB_NT; ++i)   // the block's eight rows of this pixel: independent loads, all in flight
            v[i] = i < nn ? __ldg(dX + (((size_t)b * g.N + n0 + i) * g.K + k) * g.Cx + 6 + t) : 0.f;
#pragma unroll
        for (int i = 0; i < CVB_NT; ++i) {
            if (i >= nn) break;
            const size_t r = ((size_t)b * g.N + n0 + i) * g.K + k;
            const int j = idx != nullptr ? __ldg(idx + r) : k;
            if (is_max) {
                if (idx != nullptr) atomicAdd(dmaxc + ((size_t)b * g.N2 + j) * g.C + c, v[i]);
                else aq += v[i];
            } else {
                accp[i] = __fmaf_rn(v[i], __ldg(qi + ((size_t)b * g.N2 + j) * g.C + c), accp[i]);
                if (idx != nullptr) atomicAdd(dqi + ((size_t)b * g.N2 + j) * g.C + c, v[i] * pv[i]);
                else aq = __fmaf_rn(v[i], pv[i], aq);
            }
        }
        if (idx == nullptr) {
            atomicAdd((is_max ? dmaxc : dqi) + ((size_t)b * g.N2 + k) * g.C + c, aq);
        }
    }
#pragma unroll
    for (int i = 0; i < CVB_NT; ++i) {
        if (i >= nn) break;
        if (!is_max) atomicAdd(dpi + ((size_t)b * g.N + n0 + i) * g.C + c, accp[i]);
    }
}

// Every-pixel form (idx == NULL, cost volume 1: 153 MB of dX at batch 8).  ncu of the kernel above on this shape:
// 120 registers -> 2 blocks per SM, 11 % of the warp slots, 6.2 warps stalled on the long scoreboard per issue, 8 % of
// the DRAM peak: latency-bound with eight loads in flight per thread.  Here a block owns KT = 4 pixels and a slice of
// the points; a thread (= channel) walks the slice four points at a time (16 independent loads in flight, ~40
// registers), keeps the sums over the points for its four pixels in registers (d qi / d maxc / d xyz2: ONE atomic per
// element and block at the end) and adds each point's sum over the four pixels to d pi / d xyz1 (a coalesced red per
// point).  Same arithmetic per element as the kernel above.
constexpr int CVD_KT = 4;

__global__ void __launch_bounds__(256) cv_build_bwd_dense_kernel(CvGeom g, int n_per_block, const float *__restrict__ dX,
                                                                const float *__restrict__ dxyz6, const float *__restrict__ pi,
                                                                const float *__restrict__ qi, float *dxyz1, float *dxyz2,
                                                                float *dpi, float *dqi, float *dmaxc) {
    const int k0 = blockIdx.x * CVD_KT, b = blockIdx.z, t = threadIdx.x;
    const int nb = blockIdx.y * n_per_block, ne = min(g.N, nb + n_per_block);
    const int kn = min(CVD_KT, g.K - k0);
    const bool is_max = t >= g.C;
    const int c = is_max ? t - g.C : t;
    float qv[CVD_KT], aq[CVD_KT];
#pragma unroll
    for (int kk = 0; kk < CVD_KT; ++kk) {
        qv[kk] = (!is_max && kk < kn) ? __ldg(qi + ((size_t)b * g.N2 + k0 + kk) * g.C + c) : 0.f;
        aq[kk] = 0.f;
    }
    const float *base = dX + ((size_t)b * g.N * g.K + k0) * g.Cx + 6 + t;
    // The K / 4 blocks of a slice all add into the same d pi rows: each starts its walk at a different point, so that
    // they do not hit the same addresses at the same time (same-address reds serialise in L2).
    const int nit = (ne - nb + 3) / 4, rot = nit > 0 ? (int)((blockIdx.x * 7u) % (unsigned)nit) : 0;
    for (int it = 0; it < nit; ++it) {
        const int n = nb + 4 * ((it + rot) % nit);
        float v[4][CVD_KT], pv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {       // 16 independent loads
            const bool in = n + i < ne;
            pv[i] = (in && !is_max) ? __ldg(pi + ((size_t)b * g.N + n + i) * g.C + c) : 0.f;
#pragma unroll
            for (int kk = 0; kk < CVD_KT; ++kk)
                v[i][kk] = (in && kk < kn) ? __ldg(base + ((size_t)(n + i) * g.K + kk) * g.Cx) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (n + i >= ne) break;
            if (is_max) {
#pragma unroll
                for (int kk = 0; kk < CVD_KT; ++kk) aq[kk] += v[i][kk];
            } else {
                float accp = 0.f;
#pragma unroll
                for (int kk = 0; kk < CVD_KT; ++kk) {
                    accp = __fmaf_rn(v[i][kk], qv[kk], accp);
                    aq[kk] = __fmaf_rn(v[i][kk], pv[i], aq[kk]);
                }
                atomicAdd(dpi + ((size_t)b * g.N + n + i) * g.C + c, accp);
            }
        }
    }
    for (int kk = 0; kk < kn; ++kk) atomicAdd((is_max ? dmaxc : dqi) + ((size_t)b * g.N2 + k0 + kk) * g.C + c, aq[kk]);
}

// The six coordinate channels (the copy inside X plus the separate 6-channel output), both forms: a block
// per point, a thread per pixel.  d xyz1[n] = sum over the pixels (block reduction, plain store); d xyz2[k] = sum over the
// points (one red per element).  In the kernels above these channels were a serial chain of dependent loads on six
// threads of warp 0 that the whole block waited for (~7 us per iteration: the bulk of their 245 us).
__global__ void __launch_bounds__(128) cv_coord_bwd_kernel(CvGeom g, const float *__restrict__ dX, const float *__restrict__ dxyz6,
                                                          const int32_t *__restrict__ idx, float *dxyz1, float *dxyz2) {
    __shared__ float red[4][3];
    const int n = blockIdx.x, b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float s[3] = {0.f, 0.f, 0.f};
    for (int k = threadIdx.x; k < g.K; k += 128) {
        const size_t r = ((size_t)b * g.N + n) * g.K + k;
        float w[6];
#pragma unroll
        for (int t = 0; t < 6; ++t) w[t] = __ldg(dX + r * g.Cx + t) + (dxyz6 != nullptr ? __ldg(dxyz6 + r * 6 + t) : 0.f);
#pragma unroll
        const int j = idx != nullptr ? __ldg(idx + r) : k;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            s[t] += w[t];
            atomicAdd(dxyz2 + ((size_t)b * g.N2 + j) * 3 + t, w[3 + t]);
        }
    }
#pragma unroll
    for (int t = 0; t < 3; ++t) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s[t] += __shfl_xor_sync(FULL, s[t], off);
        if (lane == 0) red[warp][t] = s[t];
    }
    __syncthreads();
    if (threadIdx.x < 3) dxyz1[((size_t)b * g.N + n) * 3 + threadIdx.x] += red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
}

__device__ __forceinline__ float masked_logit(float l, float m) { return l * m + -1e10f * (1.f - m); }

// thread per (group, channel); logits / values (G, K, C), mask (G, K) or NULL -> out (G, C)
__global__ void __launch_bounds__(256) softmax_wsum_kernel(long long total, int K, int C, const float *__restrict__ logit,
                                                          const float *__restrict__ value, const float *__restrict__ mask,
                                                          float *__restrict__ out) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long grp = e / C;
        const int c = (int)(e - grp * C);
        const float *lp = logit + (size_t)grp * K * C + c, *vp = value + (size_t)grp * K * C + c;
        const float *mp = mask != nullptr ? mask + (size_t)grp * K : nullptr;
        float mx = -INFINITY;
#pragma unroll 8
        for (int k = 0; k < K; ++k) {
            float l = __ldg(lp + (size_t)k * C);
            if (mp != nullptr) l = masked_logit(l, __ldg(mp + k));
            mx = fmaxf(mx, l);
        }
        float s = 0.f, acc = 0.f;
#pragma unroll 8
        for (int k = 0; k < K; ++k) {
            float l = __ldg(lp + (size_t)k * C);
            if (mp != nullptr) l = masked_logit(l, __ldg(mp + k));
            const float p = expf(l - mx);
            s += p;
            acc = __fmaf_rn(p, __ldg(vp + (size_t)k * C), acc);
        }
        out[e] = acc / s;
    }
}

__global__ void __launch_bounds__(256) softmax_wsum_bwd_kernel(long long total, int K, int C, const float *__restrict__ logit,
                                                              const float *__restrict__ value, const float *__restrict__ mask,
                                                              const float *__restrict__ out, const float *__restrict__ gout,
                                                              float *__restrict__ dlogit, float *__restrict__ dvalue) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long grp = e / C;
        const int c = (int)(e - grp * C);
        const size_t base = (size_t)grp * K * C + c;
        const float *mp = mask != nullptr ? mask + (size_t)grp * K : nullptr;
        float mx = -INFINITY;
#pragma unroll 8
        for (int k = 0; k < K; ++k) {
            float l = __ldg(logit + base + (size_t)k * C);
            if (mp != nullptr) l = masked_logit(l, __ldg(mp + k));
            mx = fmaxf(mx, l);
        }
        float s = 0.f;
#pragma unroll 8
        for (int k = 0; k < K; ++k) {
            float l = __ldg(logit + base + (size_t)k * C);
            if (mp != nullptr) l = masked_logit(l, __ldg(mp + k));
            s += expf(l - mx);
        }
        const float o = __ldg(out + e), gv = __ldg(gout + e), inv = 1.f / s;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            float l = __ldg(logit + base + (size_t)k * C);
            const float m = mp != nullptr ? __ldg(mp + k) : 1.f;
            if (mp != nullptr) l = masked_logit(l, m);
            const float p = expf(l - mx) * inv;
            const float pg = p * gv;
            dvalue[base + (size_t)k * C] = pg;
            dlogit[base + (size_t)k * C] = pg * (__ldg(value + base + (size_t)k * C) - o) * m;
        }
    }
}

static bool cv_geom(CvGeom &g, int B, int N, int K, int N2, int C, int has_max) {
    g.B = B; g.N = N; g.K = K; g.N2 = N2; g.C = C; g.has_max = has_max ? 1 : 0;
    g.Cx = 6 + C + (has_max ? C : 0);
    const int threads = g.Cx - 6;
    return B >= 1 && N >= 1 && K >= 1 && N2 >= 1 && C >= 8 && threads <= 256 && B <= 65535;
}

}  // namespace i2p

extern "C" {

int i2p_cv_build(int B, int N, int K, int N2, int C, const float *xyz1, const float *xyz2, const float *pi, const float *qi,
                 const float *maxc, const int32_t *idx, float *X, float *xyz6, void *stream) {
    using namespace i2p;
    CvGeom g;
    I2P_REQUIRE(cv_geom(g, B, N, K, N2, C, maxc != nullptr), "cv_build: bad sizes (C + optional C must be <= 256)");
    I2P_REQUIRE(idx != nullptr || K == N2, "cv_build: without an index every point sees every pixel (K == N2)");
    if (idx != nullptr) cv_build_kernel<true><<<dim3(N, B), g.Cx - 6, 0, as_stream(stream)>>>(g, xyz1, xyz2, pi, qi, maxc, idx, X, xyz6);
    else cv_build_kernel<false><<<dim3(N, B), g.Cx - 6, 0, as_stream(stream)>>>(g, xyz1, xyz2, pi, qi, maxc, idx, X, xyz6);
    return check_launch("cv_build");
}

int i2p_cv_build_bwd(int B, int N, int K, int N2, int C, int has_max, const float *dX, const float *dxyz6, const float *pi,
                     const float *qi, const int32_t *idx, float *dxyz1, float *dxyz2, float *dpi, float *dqi, float *dmaxc,
                     void *stream) {
    using namespace i2p;
    CvGeom g;
    I2P_REQUIRE(cv_geom(g, B, N, K, N2, C, has_max), "cv_build_bwd: bad sizes");
    I2P_REQUIRE(idx != nullptr || K == N2, "cv_build_bwd: K must equal N2 without an index");
    I2P_REQUIRE(!has_max || dmaxc != nullptr, "cv_build_bwd: dmaxc missing");
    if (idx == nullptr) {
        // about two blocks per SM: K / 4 pixel groups x point slices x B
        const int kb = ceil_div(K, CVD_KT);
        int splits = ceil_div(2 * 148, kb * B);
        if (splits < 1) splits = 1;
        int npb = ceil_div(N, splits);
        npb = ((npb + 3) / 4) * 4;
        cv_build_bwd_dense_kernel<<<dim3(kb, ceil_div(N, npb), B), g.Cx - 6, 0, as_stream(stream)>>>(g, npb, dX, dxyz6, pi, qi, dxyz1,
                                                                                                   dxyz2, dpi, dqi, dmaxc);
        int rc = check_launch("cv_build_bwd(dense)");
        if (rc != I2P_OK) return rc;
        cv_coord_bwd_kernel<<<dim3(N, B), 128, 0, as_stream(stream)>>>(g, dX, dxyz6, nullptr, dxyz1, dxyz2);
        return check_launch("cv_build_bwd(coordinates)");
    }
    cv_build_bwd_kernel<<<dim3(ceil_div(N, CVB_NT), ceil_div(K, CVB_KT), B), g.Cx - 6, 0, as_stream(stream)>>>(g, dX, dxyz6, pi, qi, idx, dxyz1, dxyz2,
                                                                                          dpi, dqi, dmaxc);
    int rc = check_launch("cv_build_bwd");
    if (rc != I2P_OK) return rc;
    cv_coord_bwd_kernel<<<dim3(N, B), 128, 0, as_stream(stream)>>>(g, dX, dxyz6, idx, dxyz1, dxyz2);
    return check_launch("cv_build_bwd(coordinates)");
}

int i2p_softmax_wsum(long long groups, int K, int C, const float *logit, const float *value, const float *mask, float *out,
                     void *stream) {
    using namespace i2p;
    I2P_REQUIRE(groups >= 0 && K >= 1 && C >= 1, "softmax_wsum: bad sizes");
    const long long total = groups * C;
    if (total == 0) return I2P_OK;
    const long long gr = (total + 255) / 256;
    softmax_wsum_kernel<<<(int)(gr < 148 * 16 ? gr : 148 * 16), 256, 0, as_stream(stream)>>>(total, K, C, logit, value, mask, out);
    return check_launch("softmax_wsum");
}

int i2p_softmax_wsum_bwd(long long groups, int K, int C, const float *logit, const float *value, const float *mask,
                         const float *out, const float *gout, float *dlogit, float *dvalue, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(groups >= 0 && K >= 1 && C >= 1, "softmax_wsum_bwd: bad sizes");
    const long long total = groups * C;
    if (total == 0) return I2P_OK;
    const long long gr = (total + 255) / 256;
    softmax_wsum_bwd_kernel<<<(int)(gr < 148 * 16 ? gr : 148 * 16), 256, 0, as_stream(stream)>>>(total, K, C, logit, value, mask, out,
                                                                                                 gout, dlogit, dvalue);
    return check_launch("softmax_wsum_bwd");
}
}
