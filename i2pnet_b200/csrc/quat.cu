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

}  // namespace i2p

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
