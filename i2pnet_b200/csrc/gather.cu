// Index-driven gathers and their scatter-add backward passes.
//
// Reference kernels: gather_points(_grad) (pointnet2/src/sampling_gpu.cu:8-79),
// group_points(_grad) (group_points_gpu.cu:8-86), three_interpolate(_grad)
// (interpolate_gpu.cu:77-160) -- one thread per OUTPUT ELEMENT with the channel in
// blockIdx.y, so every index is re-read C times -- and the torch.gather formulation of
// gather_torch (src/projectPN/utils.py:36-60), which materialises a (B,M,C) int64 index.
// This design: a thread owns one output POSITION, reads its index (and weights) once and
// walks a slab of channels, so index traffic drops by the slab size while stores stay
// coalesced along the position axis; the channels-last row gather moves whole rows with
// 128-bit loads/stores and its backward uses vector red.global.add.v4.f32.
#include "common.cuh"

namespace i2p {

constexpr int G_THREADS = 256;
constexpr int G_CSLAB = 8;  // channels per thread

// out[b,c,j] = points[b,c,idx[b,j]]     j over J = npoints*nsample positions
__global__ void __launch_bounds__(G_THREADS) gather_cn_kernel(int c, int n, int J, const float *__restrict__ points,
                                                              const int32_t *__restrict__ idx,
                                                              float *__restrict__ out) {
    const int b = blockIdx.z;
    const int j = blockIdx.x * G_THREADS + threadIdx.x;
    if (j >= J) return;
    const int c0 = blockIdx.y * G_CSLAB;
    const int i = __ldg(idx + (size_t)b * J + j);
    const float *src = points + ((size_t)b * c + c0) * n + i;
    float *dst = out + ((size_t)b * c + c0) * J + j;
    const int cc = min(G_CSLAB, c - c0);
    float v[G_CSLAB];
#pragma unroll
    for (int u = 0; u < G_CSLAB; ++u)
        if (u < cc) v[u] = __ldg(src + (size_t)u * n);
#pragma unroll
    for (int u = 0; u < G_CSLAB; ++u)
        if (u < cc) dst[(size_t)u * J] = v[u];
}

// grad_points[b,c,idx[b,j]] += grad_out[b,c,j]
__global__ void __launch_bounds__(G_THREADS) scatter_cn_kernel(int c, int n, int J,
                                                               const float *__restrict__ grad_out,
                                                               const int32_t *__restrict__ idx,
                                                               float *__restrict__ grad_points) {
    const int b = blockIdx.z;
    const int j = blockIdx.x * G_THREADS + threadIdx.x;
    if (j >= J) return;
    const int c0 = blockIdx.y * G_CSLAB;
    const int i = __ldg(idx + (size_t)b * J + j);
    const float *src = grad_out + ((size_t)b * c + c0) * J + j;
    float *dst = grad_points + ((size_t)b * c + c0) * n + i;
    const int cc = min(G_CSLAB, c - c0);
#pragma unroll
    for (int u = 0; u < G_CSLAB; ++u)
        if (u < cc) atomicAdd(dst + (size_t)u * n, __ldg(src + (size_t)u * J));
}

// out[b,c,j] = w0*p[i0] + w1*p[i1] + w2*p[i2], contracted as nvcc -O2 contracts the reference
// expression (interpolate_gpu.cu:96), SASS FMUL w1*p1; FFMA w0*p0+t; FFMA w2*p2+t:
// fma(w2,p2, fma(w0,p0, w1*p1))
__global__ void __launch_bounds__(G_THREADS) interp_kernel(int c, int m, int n, const float *__restrict__ points,
                                                           const int32_t *__restrict__ idx,
                                                           const float *__restrict__ weight,
                                                           float *__restrict__ out) {
    const int b = blockIdx.z;
    const int j = blockIdx.x * G_THREADS + threadIdx.x;
    if (j >= n) return;
    const int c0 = blockIdx.y * G_CSLAB;
    const int32_t *id = idx + ((size_t)b * n + j) * 3;
    const float *w = weight + ((size_t)b * n + j) * 3;
    const int i0 = __ldg(id), i1 = __ldg(id + 1), i2 = __ldg(id + 2);
    const float w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    const int cc = min(G_CSLAB, c - c0);
#pragma unroll
    for (int u = 0; u < G_CSLAB; ++u)
        if (u < cc) {
            const float *p = points + ((size_t)b * c + c0 + u) * m;
            out[((size_t)b * c + c0 + u) * n + j] =
                __fmaf_rn(w2, __ldg(p + i2), __fmaf_rn(w0, __ldg(p + i0), __fmul_rn(w1, __ldg(p + i1))));
        }
}

__global__ void __launch_bounds__(G_THREADS) interp_grad_kernel(int c, int n, int m,
                                                                const float *__restrict__ grad_out,
                                                                const int32_t *__restrict__ idx,
                                                                const float *__restrict__ weight,
                                                                float *__restrict__ grad_points) {
    const int b = blockIdx.z;
    const int j = blockIdx.x * G_THREADS + threadIdx.x;
    if (j >= n) return;
    const int c0 = blockIdx.y * G_CSLAB;
    const int32_t *id = idx + ((size_t)b * n + j) * 3;
    const float *w = weight + ((size_t)b * n + j) * 3;
    const int i0 = __ldg(id), i1 = __ldg(id + 1), i2 = __ldg(id + 2);
    const float w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    const int cc = min(G_CSLAB, c - c0);
#pragma unroll
    for (int u = 0; u < G_CSLAB; ++u)
        if (u < cc) {
            const float g = __ldg(grad_out + ((size_t)b * c + c0 + u) * n + j);
            float *gp = grad_points + ((size_t)b * c + c0 + u) * m;
            atomicAdd(gp + i0, __fmul_rn(g, w0));  // interpolate_gpu.cu:139-141
            atomicAdd(gp + i1, __fmul_rn(g, w1));
            atomicAdd(gp + i2, __fmul_rn(g, w2));
        }
}

// channels-last row gather: out[b,j,:] = feature[b,idx[b,j],:]
template <typename V>
__global__ void __launch_bounds__(G_THREADS) gather_rows_kernel(long long total, int hw, int cv, int m,
                                                                const V *__restrict__ feature,
                                                                const int32_t *__restrict__ idx,
                                                                V *__restrict__ out) {
    for (long long e = (long long)blockIdx.x * G_THREADS + threadIdx.x; e < total;
         e += (long long)gridDim.x * G_THREADS) {
        const long long row = e / cv;  // = b*m + j
        const int col = (int)(e - row * cv);
        const long long b = row / m;
        const int i = __ldg(idx + row);
        out[e] = __ldg(feature + (b * hw + i) * cv + col);
    }
}

// First-level set-abstraction operand (ProjectPointNet.forward_center, src/projectPN/PPBackbone_center.py:150-178):
// out[b,j,k,:] = [ g - c (3) | centre (3) | g (3) | |g - c| (1) ],  g = src[b, idx[b,j,k]], c = ctr[b,j], centre = cen[b,j]:
// gather, subtraction, broadcast, vector norm and concatenation (five launches, four passes over the 37 MB result at
// batch 8) as one pass.  Thread per (b, j, k) row: ten consecutive floats.
__global__ void __launch_bounds__(G_THREADS) sa_geometry_kernel(long long rows, int hw, int n, int K, const float *__restrict__ src,
                                                                const float *__restrict__ ctr, const float *__restrict__ cen,
                                                                const int32_t *__restrict__ idx, float *__restrict__ out) {
    for (long long r = (long long)blockIdx.x * G_THREADS + threadIdx.x; r < rows; r += (long long)gridDim.x * G_THREADS) {
        const long long bj = r / K, b = bj / n;
        const float *g = src + (b * hw + __ldg(idx + r)) * 3, *c = ctr + bj * 3, *e = cen + bj * 3;
        const float gx = __ldg(g), gy = __ldg(g + 1), gz = __ldg(g + 2);
        const float dx = __fsub_rn(gx, __ldg(c)), dy = __fsub_rn(gy, __ldg(c + 1)), dz = __fsub_rn(gz, __ldg(c + 2));
        const float d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        float2 *o = reinterpret_cast<float2 *>(out + r * 10);      // 40-byte rows: 8-byte aligned
        o[0] = make_float2(dx, dy);
        o[1] = make_float2(dz, __ldg(e));
        o[2] = make_float2(__ldg(e + 1), __ldg(e + 2));
        o[3] = make_float2(gx, gy);
        o[4] = make_float2(gz, d);
    }
}

__device__ __forceinline__ void red_add(float *p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(float4 *p, float4 v) { atomicAdd(p, v); }  // red.global.add.v4.f32

template <typename V>
__global__ void __launch_bounds__(G_THREADS) scatter_rows_kernel(long long total, int hw, int cv, int m,
                                                                 const V *__restrict__ grad_out,
                                                                 const int32_t *__restrict__ idx,
                                                                 V *__restrict__ grad_feature) {
    for (long long e = (long long)blockIdx.x * G_THREADS + threadIdx.x; e < total;
         e += (long long)gridDim.x * G_THREADS) {
        const long long row = e / cv;
        const int col = (int)(e - row * cv);
        const long long b = row / m;
        const int i = __ldg(idx + row);
        red_add(grad_feature + (b * hw + i) * cv + col, __ldg(grad_out + e));
    }
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline int rows_grid(long long total) {
    long long g = (total + G_THREADS - 1) / G_THREADS;
    const long long cap = 148LL * 16;  // 16 resident 256-thread CTAs per SM worth of grid-stride work
    return (int)(g < cap ? g : cap);
}

}  // namespace i2p

extern "C" {

#define I2P_GRID3(J, c, b) dim3(i2p::ceil_div((J), i2p::G_THREADS), i2p::ceil_div((c), i2p::G_CSLAB), (b))

int i2p_gather_points(int b, int c, int n, int npoints, const float *points, const int32_t *idx, float *out,
                      void *stream) {
    using namespace i2p;
    I2P_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && b <= 65535, "gather_points: bad sizes");
    if (b == 0 || c == 0 || npoints == 0) return I2P_OK;
    gather_cn_kernel<<<I2P_GRID3(npoints, c, b), G_THREADS, 0, as_stream(stream)>>>(c, n, npoints, points, idx, out);
    return check_launch("gather_points");
}

int i2p_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int32_t *idx,
                           float *grad_points, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && b <= 65535, "gather_points_grad: bad sizes");
    if (b == 0 || c == 0 || npoints == 0) return I2P_OK;
    scatter_cn_kernel<<<I2P_GRID3(npoints, c, b), G_THREADS, 0, as_stream(stream)>>>(c, n, npoints, grad_out, idx,
                                                                                    grad_points);
    return check_launch("gather_points_grad");
}

int i2p_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int32_t *idx,
                     float *out, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0 && b <= 65535, "group_points: bad sizes");
    const long long J = (long long)npoints * nsample;
    I2P_REQUIRE(J < (1LL << 31), "group_points: npoints*nsample overflows int32");
    if (b == 0 || c == 0 || J == 0) return I2P_OK;
    gather_cn_kernel<<<I2P_GRID3(J, c, b), G_THREADS, 0, as_stream(stream)>>>(c, n, (int)J, points, idx, out);
    return check_launch("group_points");
}

int i2p_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                          const int32_t *idx, float *grad_points, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0 && b <= 65535,
                "group_points_grad: bad sizes");
    const long long J = (long long)npoints * nsample;
    I2P_REQUIRE(J < (1LL << 31), "group_points_grad: npoints*nsample overflows int32");
    if (b == 0 || c == 0 || J == 0) return I2P_OK;
    scatter_cn_kernel<<<I2P_GRID3(J, c, b), G_THREADS, 0, as_stream(stream)>>>(c, n, (int)J, grad_out, idx,
                                                                              grad_points);
    return check_launch("group_points_grad");
}

int i2p_three_interpolate(int b, int c, int m, int n, const float *points, const int32_t *idx,
                          const float *weight, float *out, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(b >= 0 && c >= 0 && n >= 0 && m >= 0 && b <= 65535, "three_interpolate: bad sizes");
    if (b == 0 || c == 0 || n == 0) return I2P_OK;
    interp_kernel<<<I2P_GRID3(n, c, b), G_THREADS, 0, as_stream(stream)>>>(c, m, n, points, idx, weight, out);
    return check_launch("three_interpolate");
}

int i2p_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int32_t *idx,
                               const float *weight, float *grad_points, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(b >= 0 && c >= 0 && n >= 0 && m >= 0 && b <= 65535, "three_interpolate_grad: bad sizes");
    if (b == 0 || c == 0 || n == 0) return I2P_OK;
    interp_grad_kernel<<<I2P_GRID3(n, c, b), G_THREADS, 0, as_stream(stream)>>>(c, n, m, grad_out, idx, weight,
                                                                               grad_points);
    return check_launch("three_interpolate_grad");
}

int i2p_gather_rows(int b, int hw, int c, int m, const float *feature, const int32_t *flat_idx, float *out,
                    void *stream) {
    using namespace i2p;
    I2P_REQUIRE(b >= 0 && hw >= 0 && c >= 0 && m >= 0, "gather_rows: bad sizes");
    if (b == 0 || c == 0 || m == 0) return I2P_OK;
    if (c % 4 == 0 && aligned16(feature) && aligned16(out)) {
        const long long total = (long long)b * m * (c / 4);
        gather_rows_kernel<float4><<<rows_grid(total), G_THREADS, 0, as_stream(stream)>>>(
            total, hw, c / 4, m, reinterpret_cast<const float4 *>(feature), flat_idx, reinterpret_cast<float4 *>(out));
    } else {
        const long long total = (long long)b * m * c;
        gather_rows_kernel<float><<<rows_grid(total), G_THREADS, 0, as_stream(stream)>>>(total, hw, c, m, feature,
                                                                                        flat_idx, out);
    }
    return check_launch("gather_rows");
}

int i2p_sa_geometry(int b, int hw, int n, int k, const float *src, const float *ctr, const float *cen, const int32_t *flat_idx,
                    float *out, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(b >= 0 && hw >= 1 && n >= 0 && k >= 1, "sa_geometry: bad sizes");
    I2P_REQUIRE((reinterpret_cast<uintptr_t>(out) & 7) == 0, "sa_geometry: out must be 8-byte aligned");
    const long long rows = (long long)b * n * k;
    if (rows == 0) return I2P_OK;
    sa_geometry_kernel<<<rows_grid(rows), G_THREADS, 0, as_stream(stream)>>>(rows, hw, n, k, src, ctr, cen, flat_idx, out);
    return check_launch("sa_geometry");
}

int i2p_gather_rows_grad(int b, int hw, int c, int m, const float *grad_out, const int32_t *flat_idx,
                         float *grad_feature, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(b >= 0 && hw >= 0 && c >= 0 && m >= 0, "gather_rows_grad: bad sizes");
    if (b == 0 || c == 0 || m == 0) return I2P_OK;
    if (c % 4 == 0 && aligned16(grad_out) && aligned16(grad_feature)) {
        const long long total = (long long)b * m * (c / 4);
        scatter_rows_kernel<float4><<<rows_grid(total), G_THREADS, 0, as_stream(stream)>>>(
            total, hw, c / 4, m, reinterpret_cast<const float4 *>(grad_out), flat_idx,
            reinterpret_cast<float4 *>(grad_feature));
    } else {
        const long long total = (long long)b * m * c;
        scatter_rows_kernel<float><<<rows_grid(total), G_THREADS, 0, as_stream(stream)>>>(total, hw, c, m, grad_out,
                                                                                         flat_idx, grad_feature);
    }
    return check_launch("gather_rows_grad");
}
}
