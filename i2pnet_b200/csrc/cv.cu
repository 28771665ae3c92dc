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

// grid (N, B), block Cx - 6 threads (one per feature channel); threads 0..5 also write the coordinate channels
__global__ void __launch_bounds__(256) cv_build_kernel(CvGeom g, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                                                      const float *__restrict__ pi, const float *__restrict__ qi,
                                                      const float *__restrict__ maxc, const int32_t *__restrict__ idx,
                                                      float *__restrict__ X, float *__restrict__ xyz6) {
    const int n = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
    const bool is_max = t >= g.C;
    const int c = is_max ? t - g.C : t;
    const float pv = is_max ? 0.f : __ldg(pi + ((size_t)b * g.N + n) * g.C + c);
    const float xv = t < 3 ? __ldg(xyz1 + ((size_t)b * g.N + n) * 3 + t) : 0.f;
    float *row = X + ((size_t)b * g.N + n) * g.K * g.Cx;
    float *row6 = xyz6 + ((size_t)b * g.N + n) * g.K * 6;
#pragma unroll 4
    for (int k = 0; k < g.K; ++k) {
        const int j = idx != nullptr ? __ldg(idx + ((size_t)b * g.N + n) * g.K + k) : k;
        const float v = is_max ? __ldg(maxc + ((size_t)b * g.N2 + j) * g.C + c) : __fmul_rn(pv, __ldg(qi + ((size_t)b * g.N2 + j) * g.C + c));
        row[(size_t)k * g.Cx + 6 + t] = v;
        if (t < 6) {
            const float w = t < 3 ? xv : __ldg(xyz2 + ((size_t)b * g.N2 + j) * 3 + (t - 3));
            row[(size_t)k * g.Cx + t] = w;
            row6[(size_t)k * 6 + t] = w;
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
    float pv[CVB_NT], accp[CVB_NT], acc1[CVB_NT];
#pragma unroll
    for (int i = 0; i < CVB_NT; ++i) {
        pv[i] = (!is_max && i < nn) ? __ldg(pi + ((size_t)b * g.N + n0 + i) * g.C + c) : 0.f;
        accp[i] = 0.f;
        acc1[i] = 0.f;
    }
    for (int k = k0; k < k1; ++k) {
        float aq = 0.f, a2 = 0.f;   // partial sums over the block's points for (pixel k, this channel), every-pixel form
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
            if (t < 6) {   // coordinate channels: both the copy inside X and the separate 6-channel output
                float w = __ldg(dX + r * g.Cx + t);
                if (dxyz6 != nullptr) w += __ldg(dxyz6 + r * 6 + t);
                if (t < 3) acc1[i] += w;
                else if (idx != nullptr) atomicAdd(dxyz2 + ((size_t)b * g.N2 + j) * 3 + (t - 3), w);
                else a2 += w;
            }
        }
        if (idx == nullptr) {
            atomicAdd((is_max ? dmaxc : dqi) + ((size_t)b * g.N2 + k) * g.C + c, aq);
            if (t >= 3 && t < 6) atomicAdd(dxyz2 + ((size_t)b * g.N2 + k) * 3 + (t - 3), a2);
        }
    }
#pragma unroll
    for (int i = 0; i < CVB_NT; ++i) {
        if (i >= nn) break;
        if (!is_max) atomicAdd(dpi + ((size_t)b * g.N + n0 + i) * g.C + c, accp[i]);
        if (t < 3) atomicAdd(dxyz1 + ((size_t)b * g.N + n0 + i) * 3 + t, acc1[i]);
    }
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
        for (int k = 0; k < K; ++k) {
            float l = __ldg(lp + (size_t)k * C);
            if (mp != nullptr) l = masked_logit(l, __ldg(mp + k));
            mx = fmaxf(mx, l);
        }
        float s = 0.f, acc = 0.f;
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
        for (int k = 0; k < K; ++k) {
            float l = __ldg(logit + base + (size_t)k * C);
            if (mp != nullptr) l = masked_logit(l, __ldg(mp + k));
            mx = fmaxf(mx, l);
        }
        float s = 0.f;
        for (int k = 0; k < K; ++k) {
            float l = __ldg(logit + base + (size_t)k * C);
            if (mp != nullptr) l = masked_logit(l, __ldg(mp + k));
            s += expf(l - mx);
        }
        const float o = __ldg(out + e), gv = __ldg(gout + e), inv = 1.f / s;
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
    cv_build_kernel<<<dim3(N, B), g.Cx - 6, 0, as_stream(stream)>>>(g, xyz1, xyz2, pi, qi, maxc, idx, X, xyz6);
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
    cv_build_bwd_kernel<<<dim3(ceil_div(N, CVB_NT), ceil_div(K, CVB_KT), B), g.Cx - 6, 0, as_stream(stream)>>>(g, dX, dxyz6, pi, qi, idx, dxyz1, dxyz2,
                                                                                          dpi, dqi, dmaxc);
    return check_launch("cv_build_bwd");
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
