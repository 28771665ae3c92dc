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

// K split over four lanes: a warp owns 8 channels of one group, lane = 8 * (k mod 4) + channel; every lane runs an online
// softmax over its quarter of the K axis (running maximum, rescaled sum and weighted sum: one pass, one exponential pair
// per element), the quarters are merged by shuffles.  The thread-per-(group, channel) form below ran 116 736 threads of
// 160 dependent iterations each at cost volume 1 (3 blocks per SM): 49 us for 75 MB.  Also leaves (max, 1 / sum) for the
// backward pass, which then needs no reduction at all.
__global__ void __launch_bounds__(256) softmax_wsum_split_kernel(long long warps_total, int K, int C, const float *__restrict__ logit,
                                                                const float *__restrict__ value, const float *__restrict__ mask,
                                                                float *__restrict__ out, float2 *__restrict__ stat) {
    const int lane = threadIdx.x & 31, ks = lane >> 3, cl = lane & 7;
    const int c8 = C >> 3;
    for (long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); w < warps_total; w += (long long)gridDim.x * 8) {
        const long long grp = w / c8;
        const int c = (int)(w - grp * c8) * 8 + cl;
        const float *lp = logit + (size_t)grp * K * C + c, *vp = value + (size_t)grp * K * C + c;
        const float *mp = mask != nullptr ? mask + (size_t)grp * K : nullptr;
        float m = -INFINITY, s = 0.f, acc = 0.f;
#pragma unroll 4
        for (int k = ks; k < K; k += 4) {
            float l = __ldg(lp + (size_t)k * C);
            const float v = __ldg(vp + (size_t)k * C);
            if (mp != nullptr) l = masked_logit(l, __ldg(mp + k));
            const float mn = fmaxf(m, l);
            const float sc = expf(m - mn), p = expf(l - mn);     // m = -inf at the start: sc = 0
            s = __fmaf_rn(s, sc, p);
            acc = __fmaf_rn(acc, sc, p * v);
            m = mn;
        }
#pragma unroll
        for (int off = 8; off < 32; off <<= 1) {
            const float m2 = __shfl_xor_sync(FULL, m, off), s2 = __shfl_xor_sync(FULL, s, off), a2 = __shfl_xor_sync(FULL, acc, off);
            const float mn = fmaxf(m, m2);
            const float f1 = m == -INFINITY ? 0.f : expf(m - mn), f2 = m2 == -INFINITY ? 0.f : expf(m2 - mn);
            s = s * f1 + s2 * f2;
            acc = acc * f1 + a2 * f2;
            m = mn;
        }
        if (ks == 0) {
            out[grp * C + c] = acc / s;
            if (stat != nullptr) stat[grp * C + c] = make_float2(m, 1.f / s);
        }
    }
}

// backward from the saved (max, 1 / sum): element-wise over (G, K, C) in float4
__global__ void __launch_bounds__(256) softmax_wsum_bwd_flat_kernel(long long total4, int K, int C4, const float4 *__restrict__ logit,
                                                                   const float4 *__restrict__ value, const float *__restrict__ mask,
                                                                   const float4 *__restrict__ out, const float4 *__restrict__ gout,
                                                                   const float4 *__restrict__ stat, float4 *__restrict__ dlogit,
                                                                   float4 *__restrict__ dvalue) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total4; e += (long long)gridDim.x * 256) {
        const long long gk = e / C4;               // group * K + k
        const int c4 = (int)(e - gk * C4);
        const long long grp = gk / K;
        const float mk = mask != nullptr ? __ldg(mask + gk) : 1.f;
        const float4 l = __ldg(logit + e), v = __ldg(value + e), o = __ldg(out + grp * C4 + c4), g = __ldg(gout + grp * C4 + c4);
        const float4 st0 = __ldg(stat + (grp * C4 + c4) * 2), st1 = __ldg(stat + (grp * C4 + c4) * 2 + 1);   // (mx, inv) x 4 channels
        const float ll[4] = {l.x, l.y, l.z, l.w}, vv[4] = {v.x, v.y, v.z, v.w}, oo[4] = {o.x, o.y, o.z, o.w}, gg[4] = {g.x, g.y, g.z, g.w};
        const float mx[4] = {st0.x, st0.z, st1.x, st1.z}, inv[4] = {st0.y, st0.w, st1.y, st1.w};
        float dv[4], dl[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float lj = mask != nullptr ? masked_logit(ll[j], mk) : ll[j];
            const float pg = expf(lj - mx[j]) * inv[j] * gg[j];
            dv[j] = pg;
            dl[j] = pg * (vv[j] - oo[j]) * mk;
        }
        dvalue[e] = make_float4(dv[0], dv[1], dv[2], dv[3]);
        dlogit[e] = make_float4(dl[0], dl[1], dl[2], dl[3]);
    }
}

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


// ---------------------------------------------------------------------------------------------
// cost-volume operand preparation (PPBackbone_center.py:379-397) as ONE kernel per direction
// ---------------------------------------------------------------------------------------------
// Everything between the pyramids' outputs and cv_build is arithmetic on (B, N, C) / (B, N2, C) tensors of ~1 MB:
// depth restoration xyz = uv * z, the row-wise standardisation (x - mean) / max(std, 1e-12) (unbiased std) of the point
// and pixel features, and for the backward-validation channel the largest / smallest standardised point feature over the
// valid points and maxc[k,c] = qi[k,c] * (qi > 0 ? hi[c] : lo[c]).  Through ATen that is ~40 launches forward and more
// backward, every one of them on the step's critical path (CUPTI timeline: 215 us before cost volume 1, one kernel in
// flight).  One block per cloud does it all; the extrema carry the index of the first point that attains them, which
// is where the backward sends their gradient (torch.max's convention).
// (128-thread blocks: these kernels run beside persistent weight-gradient kernels that fill the SMs; a 512-thread block
// with ~90 registers per thread needs 3/4 of an SM's register file at once and waited up to 90 us for it)
constexpr int PREP_THREADS = 128, PREP_WARPS = PREP_THREADS / 32, PREP_MAXJ = 8;    // C <= 256

struct PrepArgs {
    int N, N2, C, has_max;
    const float *uv, *z, *pf, *qf;                   // (B,N,3) (B,N,1) (B,N,C) (B,N2,C)
    float *xyz, *pi, *qi, *den_p, *den_q, *maxc;     // (B,N,3) (B,N,C) (B,N2,C) (B,N) (B,N2) (B,N2,C)
    float *hi, *lo;                                  // (B,C) each; 0 when the cloud has no valid point
    int *arg_hi, *arg_lo;                            // (B,C); -1 when the cloud has no valid point
    float *scratch;                                  // (B, S, 2, C) per-block extrema, S = blocks per cloud
    int *scratch_arg;
    unsigned *tickets;                               // (B) zero between launches
};

// standardise one row held as v[j] = x[lane + 32 j]; -> the clipped denominator
template <int MJ>
__device__ __forceinline__ float standardise_row(float (&v)[MJ], int C, int lane) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < MJ; ++j) s += (lane + 32 * j < C) ? v[j] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < MJ; ++j) {
        v[j] -= mean;
        q += (lane + 32 * j < C) ? v[j] * v[j] : 0.f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(FULL, q, o);
    const float den = fmaxf(sqrtf(q / (float)(C - 1)), 1e-12f);
#pragma unroll
    for (int j = 0; j < MJ; ++j) v[j] /= den;
    return den;
}

template <int MJ>     // MJ = channels per lane: C <= 32 MJ
__global__ void __launch_bounds__(PREP_THREADS) cv_prep_fwd_kernel(const PrepArgs a) {
    __shared__ float hi_s[32 * PREP_MAXJ], lo_s[32 * PREP_MAXJ];
    __shared__ int ahi_s[32 * PREP_MAXJ], alo_s[32 * PREP_MAXJ];
    __shared__ unsigned ticket_s;
    const int b = blockIdx.y, S = gridDim.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = a.N, N2 = a.N2, C = a.C;
    float hi[MJ], lo[MJ];
    int ahi[MJ], alo[MJ];
#pragma unroll
    for (int j = 0; j < MJ; ++j) { hi[j] = -INFINITY; lo[j] = INFINITY; ahi[j] = alo[j] = -1; }
    // the cloud's rows (points, then pixels) are dealt to the S blocks of the cloud and their warps
    for (int r = blockIdx.x * PREP_WARPS + warp; r < N + N2; r += S * PREP_WARPS) {
        const bool point = r < N;
        const size_t row = point ? (size_t)b * N + r : (size_t)b * N2 + (r - N);
        const float *src = (point ? a.pf : a.qf) + row * C;
        float v[MJ];
#pragma unroll
        for (int j = 0; j < MJ; ++j) v[j] = (lane + 32 * j < C) ? __ldg(src + lane + 32 * j) : 0.f;
        const float den = standardise_row(v, C, lane);
        float *dst = (point ? a.pi : a.qi) + row * C;
#pragma unroll
        for (int j = 0; j < MJ; ++j)
            if (lane + 32 * j < C) dst[lane + 32 * j] = v[j];
        if (lane == 0) (point ? a.den_p : a.den_q)[row] = den;
        if (point) {
            const float zz = __ldg(a.z + row);
            const float x = lane < 3 ? __ldg(a.uv + row * 3 + lane) * zz : 0.f;     // restore depth (:379)
            if (lane < 3) a.xyz[row * 3 + lane] = x;
            const bool valid = __any_sync(FULL, x != 0.f);
            if (valid && a.has_max) {
#pragma unroll
                for (int j = 0; j < MJ; ++j) {       // rows come in increasing order: strict comparisons keep the first
                    if (v[j] > hi[j]) { hi[j] = v[j]; ahi[j] = r; }
                    if (v[j] < lo[j]) { lo[j] = v[j]; alo[j] = r; }
                }
            }
        }
    }
    if (!a.has_max) return;
    // merge the warps' extrema, one warp at a time (value first, then the smaller point index)
    auto merge = [&](int c, float h, int ah, float l, int al, bool first) {
        if (first) { hi_s[c] = h; ahi_s[c] = ah; lo_s[c] = l; alo_s[c] = al; return; }
        if (ah >= 0 && (ahi_s[c] < 0 || h > hi_s[c] || (h == hi_s[c] && ah < ahi_s[c]))) { hi_s[c] = h; ahi_s[c] = ah; }
        if (al >= 0 && (alo_s[c] < 0 || l < lo_s[c] || (l == lo_s[c] && al < alo_s[c]))) { lo_s[c] = l; alo_s[c] = al; }
    };
    for (int w = 0; w < PREP_WARPS; ++w) {
        if (warp == w) {
#pragma unroll
            for (int j = 0; j < MJ; ++j) merge(lane + 32 * j, hi[j], ahi[j], lo[j], alo[j], w == 0);
        }
        __syncthreads();
    }
    // the block's extrema go to the cloud's scratch rows; the block that finishes last merges them (its ticket says so)
    float *sc_v = a.scratch + ((size_t)b * S + blockIdx.x) * 2 * C;
    int *sc_i = a.scratch_arg + ((size_t)b * S + blockIdx.x) * 2 * C;
    for (int c = threadIdx.x; c < C; c += PREP_THREADS) {
        sc_v[c] = hi_s[c]; sc_v[C + c] = lo_s[c]; sc_i[c] = ahi_s[c]; sc_i[C + c] = alo_s[c];
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) ticket_s = atomicAdd(a.tickets + b, 1u);
    __syncthreads();
    if (ticket_s != (unsigned)(S - 1)) return;
    __threadfence();
    if (threadIdx.x == 0) a.tickets[b] = 0u;            // ready for the next launch
    for (int c = threadIdx.x; c < C; c += PREP_THREADS) {
        for (int s2 = 0; s2 < S; ++s2) {
            const float *pv = a.scratch + ((size_t)b * S + s2) * 2 * C;
            const int *pa = a.scratch_arg + ((size_t)b * S + s2) * 2 * C;
            merge(c, __ldcg(pv + c), __ldcg(pa + c), __ldcg(pv + C + c), __ldcg(pa + C + c), s2 == 0);
        }
    }
    __syncthreads();
    const bool any_valid = ahi_s[0] >= 0;      // a valid point sets every channel
    for (int c = threadIdx.x; c < C; c += PREP_THREADS) {
        a.hi[(size_t)b * C + c] = any_valid ? hi_s[c] : 0.f;
        a.lo[(size_t)b * C + c] = any_valid ? lo_s[c] : 0.f;
        a.arg_hi[(size_t)b * C + c] = any_valid ? ahi_s[c] : -1;
        a.arg_lo[(size_t)b * C + c] = any_valid ? alo_s[c] : -1;
    }
}

// maxc = max over the valid points of pi[n,c] * qi[k,c] = qi times the largest (qi > 0) or smallest valid pi[:,c]; -1e10 for a
// cloud without valid points.  (Element-wise over (B, N2, C); inside the last block of cv_prep_fwd it was a serial 25 us.)
__global__ void __launch_bounds__(256) cv_maxc_kernel(long long total, int N2C, int C, const float *__restrict__ qi, const float *__restrict__ hi,
                                                     const float *__restrict__ lo, const int *__restrict__ arg_hi, float *__restrict__ maxc) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long b = e / N2C;
        const int c = (int)(e % C);
        const float q = qi[e];
        maxc[e] = __ldg(arg_hi + b * C) >= 0 ? (q > 0.f ? q * __ldg(hi + b * C + c) : q * __ldg(lo + b * C + c)) : -1e10f;
    }
}

struct PrepBwdArgs {
    int N, N2, C, has_max;
    const float *uv, *z, *pi, *qi, *den_p, *den_q, *hi, *lo;
    const int *arg_hi, *arg_lo;
    const float *d_xyz, *d_pi, *d_qi, *d_maxc;      // any of them may be null (no gradient arrived)
    float *d_uv, *d_z, *d_pf, *d_qf;
};

// backward of one standardised row: y (saved output), g = dL/dy  ->  dL/dx, in place in g
template <int MJ>
__device__ __forceinline__ void standardise_row_bwd(float (&g)[MJ], const float (&y)[MJ], float den, int C, int lane) {
    float sg = 0.f, sgy = 0.f;
#pragma unroll
    for (int j = 0; j < MJ; ++j) {
        const bool in = lane + 32 * j < C;
        sg += in ? g[j] : 0.f;
        sgy += in ? g[j] * y[j] : 0.f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sg += __shfl_xor_sync(FULL, sg, o); sgy += __shfl_xor_sync(FULL, sgy, o); }
    const float mg = sg / (float)C;
    // the clipped denominator is a constant: only the mean's term remains
    const float k = den > 1e-12f ? sgy / (float)(C - 1) : 0.f;
#pragma unroll
    for (int j = 0; j < MJ; ++j) g[j] = (g[j] - mg - y[j] * k) / den;
}

template <int MJ>
__global__ void __launch_bounds__(PREP_THREADS) cv_prep_bwd_kernel(const PrepBwdArgs a) {
    __shared__ float dhi_s[32 * PREP_MAXJ], dlo_s[32 * PREP_MAXJ];
    const int b = blockIdx.y, S = gridDim.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = a.N, N2 = a.N2, C = a.C;
    const bool use_max = a.has_max && a.d_maxc != nullptr && a.arg_hi[(size_t)b * C] >= 0;
    for (int c = threadIdx.x; c < 32 * PREP_MAXJ; c += PREP_THREADS) dhi_s[c] = dlo_s[c] = 0.f;
    __syncthreads();
    if (use_max) {
        // d hi[c] = sum over the pixels with qi > 0 of d maxc * qi, d lo[c] over the others (every block of the cloud
        // computes them: N2 rows, cheaper than a second kernel)
        float dh[MJ], dl[MJ];
#pragma unroll
        for (int j = 0; j < MJ; ++j) dh[j] = dl[j] = 0.f;
#pragma unroll 4
        for (int k = warp; k < N2; k += PREP_WARPS) {
            const size_t row = ((size_t)b * N2 + k) * C;
#pragma unroll
            for (int j = 0; j < MJ; ++j) {
                const int c = lane + 32 * j;
                if (c < C) {
                    const float q = __ldg(a.qi + row + c), t = __ldg(a.d_maxc + row + c) * q;
                    if (q > 0.f) dh[j] += t; else dl[j] += t;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < MJ; ++j) { atomicAdd(&dhi_s[lane + 32 * j], dh[j]); atomicAdd(&dlo_s[lane + 32 * j], dl[j]); }
    }
    __syncthreads();
    for (int r = blockIdx.x * PREP_WARPS + warp; r < N + N2; r += S * PREP_WARPS) {
        const bool point = r < N;
        const size_t row = point ? (size_t)b * N + r : (size_t)b * N2 + (r - N);
        const float *ysrc = (point ? a.pi : a.qi) + row * C;
        const float *gsrc = point ? a.d_pi : a.d_qi;
        float y[MJ], g[MJ];
#pragma unroll
        for (int j = 0; j < MJ; ++j) {
            const int c = lane + 32 * j;
            const bool in = c < C;
            y[j] = in ? __ldg(ysrc + c) : 0.f;
            g[j] = (in && gsrc != nullptr) ? __ldg(gsrc + row * C + c) : 0.f;
            if (in && use_max) {
                if (point) {
                    if (r == __ldg(a.arg_hi + (size_t)b * C + c)) g[j] += dhi_s[c];
                    if (r == __ldg(a.arg_lo + (size_t)b * C + c)) g[j] += dlo_s[c];
                } else {
                    const float sel = y[j] > 0.f ? __ldg(a.hi + (size_t)b * C + c) : __ldg(a.lo + (size_t)b * C + c);
                    g[j] += __ldg(a.d_maxc + row * C + c) * sel;
                }
            }
        }
        standardise_row_bwd(g, y, __ldg((point ? a.den_p : a.den_q) + row), C, lane);
        float *dst = (point ? a.d_pf : a.d_qf) + row * C;
#pragma unroll
        for (int j = 0; j < MJ; ++j)
            if (lane + 32 * j < C) dst[lane + 32 * j] = g[j];
        if (point) {
            const float zz = __ldg(a.z + row);
            float gz = 0.f;
            if (lane < 3) {
                const float gx = a.d_xyz != nullptr ? __ldg(a.d_xyz + row * 3 + lane) : 0.f;
                a.d_uv[row * 3 + lane] = gx * zz;
                gz = gx * __ldg(a.uv + row * 3 + lane);
            }
            gz += __shfl_down_sync(FULL, gz, 2);
            gz += __shfl_down_sync(FULL, gz, 1);     // lane 0: g0 u0 + g2 u2 + g1 u1
            if (lane == 0) a.d_z[row] = gz;
        }
    }
}

// Pixel centres of an (h, w) feature map on the normalised camera plane: K'^-1 [u, v, 1] with K' = the intrinsic rescaled
// to the map (rows 0 / 1 times sx / sy), inverted by the adjugate (modellearn_proj_center.py:275-287; 57 ATen launches).
__global__ void __launch_bounds__(256) pixel_rays_kernel(int h, int w, float sx, float sy, const float *__restrict__ intrinsic,
                                                        float *__restrict__ rays) {
    const int b = blockIdx.y, p = blockIdx.x * 256 + threadIdx.x;
    if (p >= h * w) return;
    const float *K = intrinsic + (size_t)b * 9;
    // every product and sum rounded separately, in the order of the element-wise formulation it replaces (inverse3x3's
    // adjugate, then (grid * K^-1).sum(-1)): the rays feed a nearest-pixel selection, where one ulp can flip a near-tie
    auto mul = [](float x, float y) { return __fmul_rn(x, y); };
    auto sub = [](float x, float y) { return __fsub_rn(x, y); };
    auto add = [](float x, float y) { return __fadd_rn(x, y); };
    const float a_ = mul(K[0], sx), b_ = mul(K[1], sx), c_ = mul(K[2], sx), d_ = mul(K[3], sy), e_ = mul(K[4], sy), f_ = mul(K[5], sy);
    const float g_ = K[6], h_ = K[7], i_ = K[8];
    const float A = sub(mul(e_, i_), mul(f_, h_)), Bc = -sub(mul(d_, i_), mul(f_, g_)), Cc = sub(mul(d_, h_), mul(e_, g_));
    const float det = add(add(mul(a_, A), mul(b_, Bc)), mul(c_, Cc));
    const float adj[9] = {A, -sub(mul(b_, i_), mul(c_, h_)), sub(mul(b_, f_), mul(c_, e_)),
                          Bc, sub(mul(a_, i_), mul(c_, g_)), -sub(mul(a_, f_), mul(c_, d_)),
                          Cc, -sub(mul(a_, h_), mul(b_, g_)), sub(mul(a_, e_), mul(b_, d_))};
    const float u = (float)(p % w), v = (float)(p / w);
    float *o = rays + ((size_t)b * h * w + p) * 3;
#pragma unroll
    for (int i = 0; i < 3; ++i)
        o[i] = add(add(mul(u, __fdiv_rn(adj[3 * i], det)), mul(v, __fdiv_rn(adj[3 * i + 1], det))), __fdiv_rn(adj[3 * i + 2], det));
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
                     float *stat, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(groups >= 0 && K >= 1 && C >= 1, "softmax_wsum: bad sizes");
    const long long total = groups * C;
    if (total == 0) return I2P_OK;
    if (C % 8 == 0) {
        const long long warps = groups * (C / 8), gr = (warps + 7) / 8;
        softmax_wsum_split_kernel<<<(int)(gr < 148 * 32 ? gr : 148 * 32), 256, 0, as_stream(stream)>>>(warps, K, C, logit, value, mask, out,
                                                                                                    reinterpret_cast<float2 *>(stat));
        return check_launch("softmax_wsum");
    }
    I2P_REQUIRE(stat == nullptr, "softmax_wsum: the (max, 1 / sum) output needs a channel count that is a multiple of 8");
    const long long gr = (total + 255) / 256;
    softmax_wsum_kernel<<<(int)(gr < 148 * 16 ? gr : 148 * 16), 256, 0, as_stream(stream)>>>(total, K, C, logit, value, mask, out);
    return check_launch("softmax_wsum");
}

int i2p_softmax_wsum_bwd(long long groups, int K, int C, const float *logit, const float *value, const float *mask,
                         const float *out, const float *gout, const float *stat, float *dlogit, float *dvalue, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(groups >= 0 && K >= 1 && C >= 1, "softmax_wsum_bwd: bad sizes");
    const long long total = groups * C;
    if (total == 0) return I2P_OK;
    auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (stat != nullptr && C % 4 == 0 && al16(logit) && al16(value) && al16(out) && al16(gout) && al16(stat) && al16(dlogit) && al16(dvalue)) {
        const long long total4 = groups * K * (C / 4), gr = (total4 + 255) / 256;
        softmax_wsum_bwd_flat_kernel<<<(int)(gr < 148 * 32 ? gr : 148 * 32), 256, 0, as_stream(stream)>>>(
            total4, K, C / 4, reinterpret_cast<const float4 *>(logit), reinterpret_cast<const float4 *>(value), mask,
            reinterpret_cast<const float4 *>(out), reinterpret_cast<const float4 *>(gout), reinterpret_cast<const float4 *>(stat),
            reinterpret_cast<float4 *>(dlogit), reinterpret_cast<float4 *>(dvalue));
        return check_launch("softmax_wsum_bwd");
    }
    const long long gr = (total + 255) / 256;
    softmax_wsum_bwd_kernel<<<(int)(gr < 148 * 16 ? gr : 148 * 16), 256, 0, as_stream(stream)>>>(total, K, C, logit, value, mask, out,
                                                                                                 gout, dlogit, dvalue);
    return check_launch("softmax_wsum_bwd");
}
// blocks per cloud: enough that a warp sees two or three rows (the kernels are latency-bound chains of row reductions)
static int prep_blocks(int rows) {
    const int s = i2p::ceil_div(rows, 2 * i2p::PREP_WARPS);
    return s < 1 ? 1 : (s > 32 ? 32 : s);
}

int i2p_cv_prep_scratch_floats(int B, int N, int N2, int C) {
    /* per-block extrema (value + index) and one ticket per cloud, in 4-byte words */
    return B * prep_blocks(N + N2) * 4 * C + B;
}

int i2p_cv_prep_fwd(int B, int N, int N2, int C, int has_max, const float *uv, const float *z, const float *pf, const float *qf,
                    float *xyz, float *pi, float *qi, float *den_p, float *den_q, float *maxc, float *hi, float *lo, int32_t *arg_hi,
                    int32_t *arg_lo, float *scratch, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(B >= 0 && N >= 1 && N2 >= 1 && C >= 2 && C <= 32 * PREP_MAXJ && B <= 65535, "cv_prep: bad sizes (2 <= C <= 256)");
    I2P_REQUIRE(!has_max || (maxc != nullptr && hi != nullptr && lo != nullptr && arg_hi != nullptr && arg_lo != nullptr && scratch != nullptr),
                "cv_prep: outputs of the backward-validation channel missing");
    if (B == 0) return I2P_OK;
    const int S = prep_blocks(N + N2);
    PrepArgs a{N, N2, C, has_max ? 1 : 0, uv, z, pf, qf, xyz, pi, qi, den_p, den_q, maxc, hi, lo, arg_hi, arg_lo, nullptr, nullptr, nullptr};
    if (has_max) {
        a.scratch = scratch;
        a.scratch_arg = reinterpret_cast<int *>(scratch) + (size_t)B * S * 2 * C;
        a.tickets = reinterpret_cast<unsigned *>(scratch) + (size_t)B * S * 4 * C;
    }
    if (C <= 64) cv_prep_fwd_kernel<2><<<dim3(S, B), PREP_THREADS, 0, as_stream(stream)>>>(a);
    else if (C <= 128) cv_prep_fwd_kernel<4><<<dim3(S, B), PREP_THREADS, 0, as_stream(stream)>>>(a);
    else cv_prep_fwd_kernel<PREP_MAXJ><<<dim3(S, B), PREP_THREADS, 0, as_stream(stream)>>>(a);
    int rc = check_launch("cv_prep_fwd");
    if (rc != I2P_OK || !has_max) return rc;
    const long long total = (long long)B * N2 * C, gr = (total + 255) / 256;
    cv_maxc_kernel<<<(int)(gr < 148 * 8 ? gr : 148 * 8), 256, 0, as_stream(stream)>>>(total, N2 * C, C, qi, hi, lo, arg_hi, maxc);
    return check_launch("cv_prep_fwd(maxc)");
}

int i2p_cv_prep_bwd(int B, int N, int N2, int C, int has_max, const float *uv, const float *z, const float *pi, const float *qi,
                    const float *den_p, const float *den_q, const float *hi, const float *lo, const int32_t *arg_hi,
                    const int32_t *arg_lo, const float *d_xyz, const float *d_pi, const float *d_qi, const float *d_maxc, float *d_uv,
                    float *d_z, float *d_pf, float *d_qf, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(B >= 0 && N >= 1 && N2 >= 1 && C >= 2 && C <= 32 * PREP_MAXJ && B <= 65535, "cv_prep_bwd: bad sizes (2 <= C <= 256)");
    if (B == 0) return I2P_OK;
    PrepBwdArgs a{N, N2, C, has_max ? 1 : 0, uv, z, pi, qi, den_p, den_q, hi, lo, arg_hi, arg_lo, d_xyz, d_pi, d_qi, d_maxc, d_uv, d_z, d_pf, d_qf};
    const dim3 grid(prep_blocks(N + N2), B);
    if (C <= 64) cv_prep_bwd_kernel<2><<<grid, PREP_THREADS, 0, as_stream(stream)>>>(a);
    else if (C <= 128) cv_prep_bwd_kernel<4><<<grid, PREP_THREADS, 0, as_stream(stream)>>>(a);
    else cv_prep_bwd_kernel<PREP_MAXJ><<<grid, PREP_THREADS, 0, as_stream(stream)>>>(a);
    return check_launch("cv_prep_bwd");
}

int i2p_pixel_rays(int B, int h, int w, float sx, float sy, const float *intrinsic, float *rays, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(B >= 0 && h >= 1 && w >= 1 && B <= 65535, "pixel_rays: bad sizes");
    if (B == 0) return I2P_OK;
    pixel_rays_kernel<<<dim3(ceil_div(h * w, 256), B), 256, 0, as_stream(stream)>>>(h, w, sx, sy, intrinsic, rays);
    return check_launch("pixel_rays");
}
}
