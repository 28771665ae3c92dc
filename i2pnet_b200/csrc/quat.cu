// Hamilton product of quaternion fields with broadcasting over the point axis, with optional conjugation of
// either operand -- the one primitive behind the reference's quaternion warp and pose composition
// (src/modules/warp_utils.py:25-60 mul_q, :78-94 warp_quat_xyz; src/modellearn_proj_center.py:414-421).
// The reference spells one product as 16 multiplies + 12 adds + 8 slices + a stack on (B,N) tensors, i.e.
// ~30 kernel launches forward and ~60 in autograd's backward; five products per forward made these
// ~450 launches the largest launch-count item of the training step.  Here a product is one launch,
// and its backward is two more launches of the same kernel:  c = a (x) b  =>  da = dc (x) conj(b),
// db = conj(a) (x) dc.  Operation order inside a component follows the reference (left-to-right sums).
#include "common.cuh"

namespace i2p {

// out[b,n,:] = A (x) B,  A = a[b, na == 1 ? 0 : n, :] (conjugated if conj_a), likewise B
__global__ void __launch_bounds__(256) quat_mul_kernel(long long total, int N, int na, int nb, int conj_a, int conj_b,
                                                       const float4 *__restrict__ a, const float4 *__restrict__ b,
                                                       float4 *__restrict__ out) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long bi = e / N;
        const int n = (int)(e - bi * N);
        float4 p = __ldg(a + bi * na + (na == 1 ? 0 : n)), q = __ldg(b + bi * nb + (nb == 1 ? 0 : n));
        if (conj_a) { p.y = -p.y; p.z = -p.z; p.w = -p.w; }
        if (conj_b) { q.y = -q.y; q.z = -q.z; q.w = -q.w; }
        float4 r;   // (x, y, z, w) = components (0, 1, 2, 3)
        r.x = __fsub_rn(__fsub_rn(__fsub_rn(__fmul_rn(p.x, q.x), __fmul_rn(p.y, q.y)), __fmul_rn(p.z, q.z)), __fmul_rn(p.w, q.w));
        r.y = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.x, q.y), __fmul_rn(p.y, q.x)), __fmul_rn(p.z, q.w)), __fmul_rn(p.w, q.z));
        r.z = __fadd_rn(__fadd_rn(__fsub_rn(__fmul_rn(p.x, q.z), __fmul_rn(p.y, q.w)), __fmul_rn(p.z, q.x)), __fmul_rn(p.w, q.y));
        r.w = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(p.x, q.w), __fmul_rn(p.y, q.z)), __fmul_rn(p.z, q.y)), __fmul_rn(p.w, q.x));
        out[e] = r;
    }
}

// ---------------------------------------------------------------------------------------------
// rigid warp of a point set by a pose: out[b,n] = (q (x) [0, p] (x) q^-1)[1:4] + t,  q^-1 = conj(q) / (|q|^2 + 1e-10)
// ---------------------------------------------------------------------------------------------
// warp_quat_xyz (src/modules/warp_utils.py:78-94) and the pose composition t = R(q3) t_in + t3
// (src/modellearn_proj_center.py:414-421) are this one operation; through mul_q / inv_q it is 2 product launches +
// ~12 element-wise ones forward and ~35 backward (the quaternion inverse, the sums over the point axis of the broadcast
// operands), all on the serial chain between the coarse pose and the second cost volume.  One kernel per direction;
// the forward follows the operation order of the product kernel above.
__device__ __forceinline__ float4 qmul(float4 p, float4 q) {
    float4 r;
    r.x = __fsub_rn(__fsub_rn(__fsub_rn(__fmul_rn(p.x, q.x), __fmul_rn(p.y, q.y)), __fmul_rn(p.z, q.z)), __fmul_rn(p.w, q.w));
    r.y = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.x, q.y), __fmul_rn(p.y, q.x)), __fmul_rn(p.z, q.w)), __fmul_rn(p.w, q.z));
    r.z = __fadd_rn(__fadd_rn(__fsub_rn(__fmul_rn(p.x, q.z), __fmul_rn(p.y, q.w)), __fmul_rn(p.z, q.x)), __fmul_rn(p.w, q.y));
    r.w = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(p.x, q.w), __fmul_rn(p.y, q.z)), __fmul_rn(p.z, q.y)), __fmul_rn(p.w, q.x));
    return r;
}
__device__ __forceinline__ float4 qconj(float4 q) { return make_float4(q.x, -q.y, -q.z, -q.w); }

struct WarpArgs {
    int N, mask_invalid;          // mask_invalid: all-zero input points (empty range-image cells) stay zero
    const float *p, *q, *t;       // (B,N,3) (B,4) (B,3)
    const float *g;               // backward: dL/d out (B,N,3)
    float *out;                   // forward: (B,N,3)
    float *dp, *dq, *dt;          // backward: (B,N,3) (B,4) (B,3)
};

__device__ __forceinline__ float4 quat_inverse(float4 q, float &s) {
    s = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q.x, q.x), __fmul_rn(q.y, q.y)), __fmul_rn(q.z, q.z)), __fmul_rn(q.w, q.w)), 1e-10f);
    return make_float4(__fdiv_rn(q.x, s), __fdiv_rn(-q.y, s), __fdiv_rn(-q.z, s), __fdiv_rn(-q.w, s));
}

__global__ void __launch_bounds__(256) quat_warp_fwd_kernel(const WarpArgs a) {
    const int b = blockIdx.y;
    const float4 q = *reinterpret_cast<const float4 *>(a.q + (size_t)b * 4);
    float s;
    const float4 qi = quat_inverse(q, s);
    const float tx = a.t[(size_t)b * 3], ty = a.t[(size_t)b * 3 + 1], tz = a.t[(size_t)b * 3 + 2];
    for (int n = blockIdx.x * 256 + threadIdx.x; n < a.N; n += gridDim.x * 256) {
        const float *pp = a.p + ((size_t)b * a.N + n) * 3;
        const float4 p = make_float4(0.f, pp[0], pp[1], pp[2]);
        const float4 r = qmul(qmul(q, p), qi);
        const bool keep = !a.mask_invalid || p.y != 0.f || p.z != 0.f || p.w != 0.f;
        float *o = a.out + ((size_t)b * a.N + n) * 3;
        o[0] = keep ? __fadd_rn(r.y, tx) : 0.f;
        o[1] = keep ? __fadd_rn(r.z, ty) : 0.f;
        o[2] = keep ? __fadd_rn(r.w, tz) : 0.f;
    }
}

// one block per cloud: the gradients of the pose are sums over its points
__global__ void __launch_bounds__(256) quat_warp_bwd_kernel(const WarpArgs a) {
    __shared__ float red[8][11];
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4 q = *reinterpret_cast<const float4 *>(a.q + (size_t)b * 4);
    float s;
    const float4 qi = quat_inverse(q, s);
    float acc[11];     // d q through the first product (4), d q^-1 (4), d t (3)
#pragma unroll
    for (int i = 0; i < 11; ++i) acc[i] = 0.f;
    for (int n = threadIdx.x; n < a.N; n += 256) {
        const size_t o = ((size_t)b * a.N + n) * 3;
        const float4 p = make_float4(0.f, a.p[o], a.p[o + 1], a.p[o + 2]);
        const bool keep = !a.mask_invalid || p.y != 0.f || p.z != 0.f || p.w != 0.f;
        const float4 dr = keep ? make_float4(0.f, a.g[o], a.g[o + 1], a.g[o + 2]) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 u = qmul(q, p);
        const float4 du = qmul(dr, qconj(qi));            // c = a (x) b:  da = dc (x) conj(b),  db = conj(a) (x) dc
        const float4 dqi = qmul(qconj(u), dr);
        const float4 dpq = qmul(qconj(q), du);
        const float4 dq1 = qmul(du, qconj(p));
        if (a.dp != nullptr) { a.dp[o] = dpq.y; a.dp[o + 1] = dpq.z; a.dp[o + 2] = dpq.w; }
        acc[0] += dq1.x; acc[1] += dq1.y; acc[2] += dq1.z; acc[3] += dq1.w;
        acc[4] += dqi.x; acc[5] += dqi.y; acc[6] += dqi.z; acc[7] += dqi.w;
        acc[8] += dr.y; acc[9] += dr.z; acc[10] += dr.w;
    }
#pragma unroll
    for (int i = 0; i < 11; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(FULL, acc[i], o);
        if (lane == 0) red[warp][i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t[11];
        for (int i = 0; i < 11; ++i) {
            t[i] = 0.f;
            for (int w = 0; w < 8; ++w) t[i] += red[w][i];
        }
        // q^-1 = conj(q) / s, s = |q|^2 + 1e-10:  d conj(q) = d q^-1 / s,  d s = -(d q^-1 . conj(q)) / s^2,  d q += 2 q d s
        const float ds = -(t[4] * q.x - t[5] * q.y - t[6] * q.z - t[7] * q.w) / (s * s);
        if (a.dq != nullptr) {
            a.dq[(size_t)b * 4] = t[0] + t[4] / s + 2.f * q.x * ds;
            a.dq[(size_t)b * 4 + 1] = t[1] - t[5] / s + 2.f * q.y * ds;
            a.dq[(size_t)b * 4 + 2] = t[2] - t[6] / s + 2.f * q.z * ds;
            a.dq[(size_t)b * 4 + 3] = t[3] - t[7] / s + 2.f * q.w * ds;
        }
        if (a.dt != nullptr) { a.dt[(size_t)b * 3] = t[8]; a.dt[(size_t)b * 3 + 1] = t[9]; a.dt[(size_t)b * 3 + 2] = t[10]; }
    }
}

}  // namespace i2p

extern "C" int i2p_quat_warp_fwd(int B, int N, int mask_invalid, const float *p, const float *q, const float *t, float *out, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(B >= 0 && N >= 1 && B <= 65535, "quat_warp: bad sizes");
    I2P_REQUIRE(((uintptr_t)q & 15) == 0, "quat_warp: the quaternions must be 16-byte aligned");
    if (B == 0) return I2P_OK;
    WarpArgs a{N, mask_invalid, p, q, t, nullptr, out, nullptr, nullptr, nullptr};
    const int gx = ceil_div(N, 256);
    quat_warp_fwd_kernel<<<dim3(gx < 64 ? gx : 64, B), 256, 0, as_stream(stream)>>>(a);
    return check_launch("quat_warp_fwd");
}

extern "C" int i2p_quat_warp_bwd(int B, int N, int mask_invalid, const float *p, const float *q, const float *g, float *dp, float *dq,
                                 float *dt, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(B >= 0 && N >= 1 && B <= 65535, "quat_warp_bwd: bad sizes");
    I2P_REQUIRE(((uintptr_t)q & 15) == 0, "quat_warp_bwd: the quaternions must be 16-byte aligned");
    if (B == 0) return I2P_OK;
    WarpArgs a{N, mask_invalid, p, q, nullptr, g, nullptr, dp, dq, dt};
    quat_warp_bwd_kernel<<<B, 256, 0, as_stream(stream)>>>(a);
    return check_launch("quat_warp_bwd");
}

extern "C" int i2p_quat_mul(int B, int N, int na, int nb, int conj_a, int conj_b, const float *a, const float *b, float *out,
                            void *stream) {
    using namespace i2p;
    I2P_REQUIRE(B >= 0 && N >= 1 && (na == 1 || na == N) && (nb == 1 || nb == N), "quat_mul: operands must be (B,1,4) or (B,N,4)");
    I2P_REQUIRE(((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0 && ((uintptr_t)out & 15) == 0, "quat_mul: 16-byte alignment");
    const long long total = (long long)B * N;
    if (total == 0) return I2P_OK;
    const long long g = (total + 255) / 256;
    quat_mul_kernel<<<(int)(g < 148 * 8 ? g : 148 * 8), 256, 0, as_stream(stream)>>>(
        total, N, na, nb, conj_a, conj_b, reinterpret_cast<const float4 *>(a), reinterpret_cast<const float4 *>(b),
        reinterpret_cast<float4 *>(out));
    return check_launch("quat_mul");
}
