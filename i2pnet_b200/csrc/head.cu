// Pose head and pose loss as single kernels (forward and backward each).
//
// PoseHead (src/projectPN/PPBackbone_center.py:503-560): softmax of the mask over the POINTS (per channel), mask-weighted
// pooling of the prediction, a 1x1 "hidden" layer (C -> 256), dropout, and the two linear heads (256 -> 4 quaternion,
// 256 -> 3 translation), the quaternion normalised.  The reference runs it as ~16 ATen / cuDNN launches forward and ~34
// backward per head on tensors of a few KB; here one CTA per sample does the whole chain out of shared memory.
// Get_loss (compute_loss.py:102-133): the learnably balanced L1 / L2 pose loss of the refined and the coarse pose,
// ~47 launches forward in the reference formulation, one CTA here.
#include <math.h>

#include "common.cuh"

namespace i2p {

constexpr int HEAD_THREADS = 1024;    // the forward is three dependent passes over the points of ONE sample per block: latency-bound, so as many row groups as a block holds (256 threads: 50 us per head, twice on the serial chain)
constexpr int HEAD_MAXC = 128, HEAD_MAXH = 256;

struct HeadArgs {
    int B, N, C, Hd;
    const float *pred, *mask;          // (B, N, C)
    const float *w1, *b1;              // (Hd, C), (Hd)
    const float *wq, *bq, *wt, *bt;    // (4, Hd), (4), (3, Hd), (3)
    const float *drop;                 // (B, Hd) dropout multipliers (0 or 1 / (1 - p)), or null
    float *mask_p;                     // (B, N, C) softmax over the points
    float *pooled, *hidden, *q_raw;    // (B, C), (B, Hd) after dropout, (B, 4) before normalisation
    float *q, *t;                      // (B, 4), (B, 3)
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    return v;
}

__global__ void __launch_bounds__(HEAD_THREADS) pose_head_fwd_kernel(const HeadArgs a) {
    __shared__ float part[HEAD_THREADS / 32][HEAD_MAXC];   // per row-group partials (max, then sum, then weighted sum)
    __shared__ float cmax[HEAD_MAXC], csum[HEAD_MAXC], pooled[HEAD_MAXC], hidden[HEAD_MAXH], out7[8];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int C = a.C, N = a.N, Hd = a.Hd;
    const int G = HEAD_THREADS / C;                         // row groups (C divides the block size: 64 or 128 channels)
    const int c = tid % C, g = tid / C;
    const float *mk = a.mask + (size_t)b * N * C, *pr = a.pred + (size_t)b * N * C;
    float *mp = a.mask_p + (size_t)b * N * C;
    // softmax over the points, per channel: max, sum of exponentials, weighted sum
    float m = -INFINITY;
    if (g < G) {
#pragma unroll 8
        for (int n = g; n < N; n += G) m = fmaxf(m, __ldg(mk + (size_t)n * C + c));
    }
    if (g < G) part[g][c] = m;
    __syncthreads();
    if (tid < C) {
        float mm = part[0][tid];
        for (int i = 1; i < G; ++i) mm = fmaxf(mm, part[i][tid]);
        cmax[tid] = mm;
    }
    __syncthreads();
    float s = 0.f;
    if (g < G) {
        const float mx = cmax[c];
#pragma unroll 8
        for (int n = g; n < N; n += G) s += __expf(__ldg(mk + (size_t)n * C + c) - mx);
    }
    if (g < G) part[g][c] = s;
    __syncthreads();
    if (tid < C) {
        float ss = 0.f;
        for (int i = 0; i < G; ++i) ss += part[i][tid];
        csum[tid] = ss;
    }
    __syncthreads();
    float acc = 0.f;
    if (g < G) {
        const float inv = 1.f / csum[c], mx = cmax[c];
#pragma unroll 8
        for (int n = g; n < N; n += G) {
            const float p = __expf(__ldg(mk + (size_t)n * C + c) - mx) * inv;
            mp[(size_t)n * C + c] = p;
            acc = __fmaf_rn(__ldg(pr + (size_t)n * C + c), p, acc);
        }
        part[g][c] = acc;
    }
    __syncthreads();
    if (tid < C) {
        float ss = 0.f;
        for (int i = 0; i < G; ++i) ss += part[i][tid];
        pooled[tid] = ss;
        a.pooled[(size_t)b * C + tid] = ss;
    }
    __syncthreads();
    // hidden layer: a warp per output (lanes run along the weight row: coalesced), dropout
    {
        const int warp = tid >> 5, lane = tid & 31;
        for (int h = warp; h < Hd; h += HEAD_THREADS / 32) {
            const float *w = a.w1 + (size_t)h * C;
            float v = 0.f;
            for (int k = lane; k < C; k += 32) v = __fmaf_rn(__ldg(w + k), pooled[k], v);
            v = warp_sum(v);
            if (lane == 0) {
                v += a.b1[h];
                if (a.drop != nullptr) v *= a.drop[(size_t)b * Hd + h];
                hidden[h] = v;
                a.hidden[(size_t)b * Hd + h] = v;
            }
        }
    }
    __syncthreads();
    // the seven head outputs: one warp each
    const int warp = tid >> 5, lane = tid & 31;
    if (warp < 7) {
        const float *w = warp < 4 ? a.wq + (size_t)warp * Hd : a.wt + (size_t)(warp - 4) * Hd;
        float v = 0.f;
        for (int k = lane; k < Hd; k += 32) v = __fmaf_rn(w[k], hidden[k], v);
        v = warp_sum(v);
        if (lane == 0) out7[warp] = v + (warp < 4 ? a.bq[warp] : a.bt[warp - 4]);
    }
    __syncthreads();
    if (tid == 0) {
        const float s2 = out7[0] * out7[0] + out7[1] * out7[1] + out7[2] * out7[2] + out7[3] * out7[3];
        const float den = sqrtf(s2 + 1e-10f) + 1e-10f;
        for (int i = 0; i < 4; ++i) {
            a.q_raw[b * 4 + i] = out7[i];
            a.q[b * 4 + i] = out7[i] / den;
        }
        for (int i = 0; i < 3; ++i) a.t[b * 3 + i] = out7[4 + i];
    }
}

struct HeadBwdArgs {
    int B, N, C, Hd;
    const float *pred, *mask_p, *pooled, *hidden, *q_raw, *drop;
    const float *w1, *wq, *wt;
    const float *dq, *dt;                       // (B, 4), (B, 3); either may be null (no gradient)
    float *dpred, *dmask;                       // (B, N, C)
    float *dw1, *db1, *dwq, *dbq, *dwt, *dbt;   // accumulated with atomics (sum over the batch)
};

// grid (B, HEAD_SPLIT): the blocks of a sample all rebuild the (tiny) head gradients, then share the expensive parts --
// the Hd x C outer product for dW1 and the N x C element-wise tail -- by slices.
constexpr int HEAD_SPLIT = 4;
constexpr int HEAD_BWD_THREADS = 1024;   // (same reasoning as the forward: dependent passes of a few KB per block)

__global__ void __launch_bounds__(HEAD_BWD_THREADS) pose_head_bwd_kernel(const HeadBwdArgs a) {
    __shared__ float d7[8], dhid[HEAD_MAXH], dpool[HEAD_MAXC], pooled[HEAD_MAXC], part[HEAD_BWD_THREADS / 32][HEAD_MAXC];
    const int b = blockIdx.x, sp = blockIdx.y, tid = threadIdx.x;
    const int C = a.C, N = a.N, Hd = a.Hd;
    const bool lead = sp == 0;      // the block that adds the small parameter gradients
    if (tid == 0) {
        // q = r / (s + e), s = sqrt(|r|^2 + 1e-10):  dr = dq / (s + e) - r (dq . r) / ((s + e)^2 s)
        float r[4], dq[4];
        for (int i = 0; i < 4; ++i) { r[i] = a.q_raw[b * 4 + i]; dq[i] = a.dq != nullptr ? a.dq[b * 4 + i] : 0.f; }
        const float s = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3] + 1e-10f), den = s + 1e-10f;
        const float dot = dq[0] * r[0] + dq[1] * r[1] + dq[2] * r[2] + dq[3] * r[3];
        for (int i = 0; i < 4; ++i) d7[i] = dq[i] / den - r[i] * dot / (den * den * s);
        for (int i = 0; i < 3; ++i) d7[4 + i] = a.dt != nullptr ? a.dt[b * 3 + i] : 0.f;
        if (lead) {
            for (int i = 0; i < 4; ++i) atomicAdd(a.dbq + i, d7[i]);
            for (int i = 0; i < 3; ++i) atomicAdd(a.dbt + i, d7[4 + i]);
        }
    }
    for (int k = tid; k < C; k += HEAD_BWD_THREADS) pooled[k] = a.pooled[(size_t)b * C + k];
    __syncthreads();
    // d hidden (after dropout) = Wq^T dq_raw + Wt^T dt ; head weight gradients ; through the dropout
    for (int h = tid; h < Hd; h += HEAD_BWD_THREADS) {
        const float hv = a.hidden[(size_t)b * Hd + h];
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v = __fmaf_rn(a.wq[(size_t)i * Hd + h], d7[i], v);
            if (lead) atomicAdd(a.dwq + (size_t)i * Hd + h, d7[i] * hv);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            v = __fmaf_rn(a.wt[(size_t)i * Hd + h], d7[4 + i], v);
            if (lead) atomicAdd(a.dwt + (size_t)i * Hd + h, d7[4 + i] * hv);
        }
        if (a.drop != nullptr) v *= a.drop[(size_t)b * Hd + h];
        dhid[h] = v;
        if (lead) atomicAdd(a.db1 + h, v);
    }
    __syncthreads();
    // dW1 = dhid (x) pooled: this block's slice of the hidden units
    const int h_per = (Hd + HEAD_SPLIT - 1) / HEAD_SPLIT, h0 = sp * h_per, h1 = min(Hd, h0 + h_per);
    for (int e = h0 * C + tid; e < h1 * C; e += HEAD_BWD_THREADS) {
        const int h = e / C, k = e - h * C;
        atomicAdd(a.dw1 + e, dhid[h] * pooled[k]);
    }
    // dpooled = W1^T dhid: threads = (channel, slice of the hidden units), partial sums meet in shared memory
    {
        const int G = HEAD_BWD_THREADS / C, k = tid % C, gq = tid / C;
        float v = 0.f;
        if (gq < G) {
#pragma unroll 8
            for (int h = gq; h < Hd; h += G) v = __fmaf_rn(__ldg(a.w1 + (size_t)h * C + k), dhid[h], v);
        }
        if (gq < G) part[gq][k] = v;
        __syncthreads();
        if (tid < C) {
            float t = 0.f;
            for (int i = 0; i < G; ++i) t += part[i][tid];
            dpool[tid] = t;
        }
        __syncthreads();
    }
    // pooled = sum_n pred p:  dpred = dpooled p ;  dmask = p dpooled (pred - pooled)   (softmax over the points)
    const float *pr = a.pred + (size_t)b * N * C, *mp = a.mask_p + (size_t)b * N * C;
    float *dp = a.dpred + (size_t)b * N * C, *dm = a.dmask + (size_t)b * N * C;
    const int e_per = (N * C + HEAD_SPLIT - 1) / HEAD_SPLIT, e0 = sp * e_per, e1 = min(N * C, e0 + e_per);
    for (int e = e0 + tid; e < e1; e += HEAD_BWD_THREADS) {
        const int k = e % C;
        const float p = mp[e], g = dpool[k];
        dp[e] = g * p;
        dm[e] = p * g * (pr[e] - pooled[k]);
    }
}

// ---- Get_loss: total = 1.6 L(out4) + 0.8 L(out3), L = lx exp(-sx) + sx + lq exp(-sq) + sq,
// lq = mean_b sqrt(|q_gt - q|^2 + 1e-10), lx = mean |t - t_gt| (l1) or mean_b sqrt(|t - t_gt|^2 + 1e-10)
struct LossArgs {
    int B, l1;
    const float *out3, *out4, *q_gt, *t_gt, *sx, *sq;
    float *loss;         // (3): total, rotation part, translation part
    const float *dloss;  // backward: (1) gradient of the total
    float *dout3, *dout4, *dsx, *dsq;
};

__device__ __forceinline__ void pose_terms(const LossArgs &a, const float *o, int b, float &lq, float &lx) {
    float s = 1e-10f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float d = a.q_gt[b * 4 + i] - o[b * 7 + i]; s = __fmaf_rn(d, d, s); }
    lq = sqrtf(s);
    if (a.l1) {
        lx = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) lx += fabsf(o[b * 7 + 4 + i] - a.t_gt[b * 3 + i]);
    } else {
        float u = 1e-10f;
#pragma unroll
        for (int i = 0; i < 3; ++i) { const float d = o[b * 7 + 4 + i] - a.t_gt[b * 3 + i]; u = __fmaf_rn(d, d, u); }
        lx = sqrtf(u);
    }
}

template <bool BWD>
__global__ void __launch_bounds__(128) pose_loss_kernel(const LossArgs a) {
    __shared__ float red[4][4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float ex = __expf(-a.sx[0]), eq = __expf(-a.sq[0]);
    const float nq = 1.f / (float)a.B, nx = a.l1 ? 1.f / (float)(3 * a.B) : 1.f / (float)a.B;
    float q3 = 0.f, x3 = 0.f, q4 = 0.f, x4 = 0.f;
    for (int b = tid; b < a.B; b += 128) {
        float lq, lx;
        pose_terms(a, a.out3, b, lq, lx);
        q3 += lq; x3 += lx;
        if (BWD) {
            const float g = a.dloss[0] * 0.8f;
            for (int i = 0; i < 4; ++i) a.dout3[b * 7 + i] = g * eq * nq * (a.out3[b * 7 + i] - a.q_gt[b * 4 + i]) / lq;
            for (int i = 0; i < 3; ++i) {
                const float d = a.out3[b * 7 + 4 + i] - a.t_gt[b * 3 + i];
                a.dout3[b * 7 + 4 + i] = g * ex * nx * (a.l1 ? (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) : d / lx);
            }
        }
        pose_terms(a, a.out4, b, lq, lx);
        q4 += lq; x4 += lx;
        if (BWD) {
            const float g = a.dloss[0] * 1.6f;
            for (int i = 0; i < 4; ++i) a.dout4[b * 7 + i] = g * eq * nq * (a.out4[b * 7 + i] - a.q_gt[b * 4 + i]) / lq;
            for (int i = 0; i < 3; ++i) {
                const float d = a.out4[b * 7 + 4 + i] - a.t_gt[b * 3 + i];
                a.dout4[b * 7 + 4 + i] = g * ex * nx * (a.l1 ? (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) : d / lx);
            }
        }
    }
    q3 = warp_sum(q3); x3 = warp_sum(x3); q4 = warp_sum(q4); x4 = warp_sum(x4);
    if (lane == 0) { red[warp][0] = q3; red[warp][1] = x3; red[warp][2] = q4; red[warp][3] = x4; }
    __syncthreads();
    if (tid == 0) {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        for (int w = 0; w < 4; ++w)
            for (int i = 0; i < 4; ++i) t[i] += red[w][i];
        const float lq3 = t[0] * nq, lx3 = t[1] * nx, lq4 = t[2] * nq, lx4 = t[3] * nx;
        if (!BWD) {
            const float sx = a.sx[0], sq = a.sq[0];
            const float l3 = lx3 * ex + sx + lq3 * eq + sq, l4 = lx4 * ex + sx + lq4 * eq + sq;
            a.loss[0] = 1.6f * l4 + 0.8f * l3;
            a.loss[1] = 1.6f * lq4 + 0.8f * lq3;
            a.loss[2] = 1.6f * lx4 + 0.8f * lx3;
        } else {
            const float g = a.dloss[0];
            a.dsx[0] = g * (1.6f * (1.f - lx4 * ex) + 0.8f * (1.f - lx3 * ex));
            a.dsq[0] = g * (1.6f * (1.f - lq4 * eq) + 0.8f * (1.f - lq3 * eq));
        }
    }
}

}  // namespace i2p

extern "C" {

int i2p_pose_head_fwd(int B, int N, int C, int Hd, const float *pred, const float *mask, const float *w1, const float *b1,
                      const float *wq, const float *bq, const float *wt, const float *bt, const float *drop, float *mask_p,
                      float *pooled, float *hidden, float *q_raw, float *q, float *t, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(B >= 1 && N >= 1 && C >= 1 && C <= HEAD_MAXC && HEAD_THREADS % C == 0 && Hd >= 1 && Hd <= HEAD_MAXH,
                "pose_head_fwd: C must divide 256 and be <= 128, hidden <= 256");
    HeadArgs a{B, N, C, Hd, pred, mask, w1, b1, wq, bq, wt, bt, drop, mask_p, pooled, hidden, q_raw, q, t};
    pose_head_fwd_kernel<<<B, HEAD_THREADS, 0, as_stream(stream)>>>(a);
    return check_launch("pose_head_fwd");
}

int i2p_pose_head_bwd(int B, int N, int C, int Hd, const float *pred, const float *mask_p, const float *pooled,
                      const float *hidden, const float *q_raw, const float *drop, const float *w1, const float *wq,
                      const float *wt, const float *dq, const float *dt, float *dpred, float *dmask, float *dw1, float *db1,
                      float *dwq, float *dbq, float *dwt, float *dbt, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(B >= 1 && N >= 1 && C >= 1 && C <= HEAD_MAXC && HEAD_BWD_THREADS % C == 0 && Hd >= 1 && Hd <= HEAD_MAXH, "pose_head_bwd: bad sizes");
    HeadBwdArgs a{B, N, C, Hd, pred, mask_p, pooled, hidden, q_raw, drop, w1, wq, wt, dq, dt, dpred, dmask,
                  dw1, db1, dwq, dbq, dwt, dbt};
    pose_head_bwd_kernel<<<dim3(B, HEAD_SPLIT), HEAD_BWD_THREADS, 0, as_stream(stream)>>>(a);
    return check_launch("pose_head_bwd");
}

int i2p_pose_loss_fwd(int B, int l1, const float *out3, const float *out4, const float *q_gt, const float *t_gt,
                      const float *sx, const float *sq, float *loss3, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(B >= 1, "pose_loss_fwd: bad batch");
    LossArgs a{B, l1, out3, out4, q_gt, t_gt, sx, sq, loss3, nullptr, nullptr, nullptr, nullptr, nullptr};
    pose_loss_kernel<false><<<1, 128, 0, as_stream(stream)>>>(a);
    return check_launch("pose_loss_fwd");
}

int i2p_pose_loss_bwd(int B, int l1, const float *out3, const float *out4, const float *q_gt, const float *t_gt,
                      const float *sx, const float *sq, const float *dloss, float *dout3, float *dout4, float *dsx,
                      float *dsq, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(B >= 1, "pose_loss_bwd: bad batch");
    LossArgs a{B, l1, out3, out4, q_gt, t_gt, sx, sq, nullptr, dloss, dout3, dout4, dsx, dsq};
    pose_loss_kernel<true><<<1, 128, 0, as_stream(stream)>>>(a);
    return check_launch("pose_loss_bwd");
}
}
