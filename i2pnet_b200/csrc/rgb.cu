// Tail of one RGB feature-pyramid block: BatchNorm2d -> LeakyReLU(0.1) -> MaxPool2d(3, stride, pad 1)
// over the NCHW f32 output of the block's 3x3 convolution, forward and backward.
//
// Reference formulation (src/modules/basicConv.py:6-20, fifteen blocks per forward): three separate
// ATen / cuDNN passes forward and three backward over tensors of up to 42 MB (batch 8, 16 x 160 x 512).
// The library batch-norm kernels assign ONE thread block per channel, so the 16- and 32-channel
// full-resolution layers run on 16-32 of the 148 SMs (measured: 283 us per backward call, 15x off the
// HBM time of the bytes they move).
//
// Here the three ops are one pass each way over the convolution output y, parallel over
// (chunk of a plane) x channel x sample:
//   forward : per-chunk (count, mean, M2) -> Chan merge per channel (also updates the running
//             statistics) -> out = maxpool(leaky(y * scale + shift)), arg-max offset kept as int8;
//   backward: the gradient of an input position is gathered from the <= 9 pooling windows that cover
//             it (no atomics, no zero-filled scatter target); pass 1 reduces the two batch-norm
//             sums S1 = sum dz, S2 = sum dz * yhat, pass 2 writes dy = scale (dz - S1/n - yhat S2/n).
// Neither the normalised nor the activated tensor ever exists in HBM.  Algorithmic bytes per
// element of y: forward 4 (stats) + 4 + 5/s^2 (pool), backward 2 x (4 + 5/s^2) + 4.
#include <math.h>

#include "common.cuh"

namespace i2p {

constexpr int RGB_CHUNK = 4096;   // elements of one (sample, channel) plane per block: 16 per thread
constexpr int RGB_THREADS = 256;

__device__ __forceinline__ float block_sum(float v, float *red) {   // red[8]; result broadcast to all threads
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    const int warp = threadIdx.x >> 5;
    __syncthreads();   // red may still be read from a previous call
    if ((threadIdx.x & 31) == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < RGB_THREADS / 32; ++w) t += red[w];
    return t;
}

// tile_stats (C, B * chunks, 3) = (count, mean, M2) of y[b, c, chunk]
__global__ void __launch_bounds__(RGB_THREADS) rgb_bn_stats_kernel(int C, int HW, const float *__restrict__ y,
                                                                   float *__restrict__ tile_stats) {
    __shared__ float red[RGB_THREADS / 32];
    const int chunk = blockIdx.x, c = blockIdx.y, b = blockIdx.z, chunks = gridDim.x;
    const int start = chunk * RGB_CHUNK, stop = min(HW, start + RGB_CHUNK);
    const float *p = y + ((size_t)b * C + c) * HW;
    float v[RGB_CHUNK / RGB_THREADS];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < RGB_CHUNK / RGB_THREADS; ++j) {
        const int i = start + threadIdx.x + RGB_THREADS * j;
        v[j] = i < stop ? __ldg(p + i) : 0.f;
        s += v[j];
    }
    const float n = (float)(stop - start);
    const float mean = block_sum(s, red) / n;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < RGB_CHUNK / RGB_THREADS; ++j) {
        const float d = v[j] - mean;
        q += (start + threadIdx.x + RGB_THREADS * j < stop) ? d * d : 0.f;
    }
    q = block_sum(q, red);
    if (threadIdx.x == 0) {
        float *t = tile_stats + ((size_t)c * (gridDim.z * chunks) + (size_t)b * chunks + chunk) * 3;
        t[0] = n; t[1] = mean; t[2] = q;
    }
}

__device__ __forceinline__ void chan3(float &n, float &mu, float &m2, float n2, float mu2, float s2) {
    if (n2 > 0.f) {
        const float tot = n + n2, d = mu2 - mu, f = n2 / tot;
        mu += d * f;
        m2 += s2 + d * d * n * f;
        n = tot;
    }
}

// stats (4, C) = mean, rstd, scale, shift; running statistics updated as nn.BatchNorm2d does in training
// (momentum, unbiased variance); s12 (2, C) f64 zeroed for the backward pass.
__global__ void __launch_bounds__(128) rgb_bn_finalize_kernel(int C, int ntiles, const float *tile_stats,
                                                              const float *gamma, const float *beta, float eps,
                                                              float momentum, float *running_mean, float *running_var,
                                                              long long *num_batches_tracked, float *stats,
                                                              double *s12) {
    __shared__ float sn[4], smu[4], sm2[4];
    const int c = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float n = 0.f, mu = 0.f, m2 = 0.f;
    for (int t = threadIdx.x; t < ntiles; t += 128) {
        const float *ts = tile_stats + ((size_t)c * ntiles + t) * 3;
        chan3(n, mu, m2, ts[0], ts[1], ts[2]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        chan3(n, mu, m2, __shfl_down_sync(FULL, n, off), __shfl_down_sync(FULL, mu, off), __shfl_down_sync(FULL, m2, off));
    if (lane == 0) { sn[warp] = n; smu[warp] = mu; sm2[warp] = m2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 4; ++w) chan3(n, mu, m2, sn[w], smu[w], sm2[w]);
        const double var = (double)m2 / (double)n;
        const float r = (float)(1.0 / sqrt(var + (double)eps));
        const float g = gamma != nullptr ? gamma[c] : 1.f, bt = beta != nullptr ? beta[c] : 0.f;
        stats[c] = mu;
        stats[C + c] = r;
        stats[2 * C + c] = g * r;
        stats[3 * C + c] = bt - mu * g * r;
        if (running_mean != nullptr) {
            const double unbiased = n > 1.f ? (double)m2 / ((double)n - 1.0) : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
            if (c == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;
        }
        if (s12 != nullptr) { s12[c] = 0.0; s12[C + c] = 0.0; }
    }
}

// eval mode: the same constants from the running statistics
__global__ void rgb_bn_from_running_kernel(int C, const float *gamma, const float *beta, float eps,
                                           const float *running_mean, const float *running_var, float *stats,
                                           double *s12) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float mu = running_mean[c];
    const float r = (float)(1.0 / sqrt((double)running_var[c] + (double)eps));
    const float g = gamma != nullptr ? gamma[c] : 1.f, bt = beta != nullptr ? beta[c] : 0.f;
    stats[c] = mu;
    stats[C + c] = r;
    stats[2 * C + c] = g * r;
    stats[3 * C + c] = bt - mu * g * r;
    if (s12 != nullptr) { s12[c] = 0.0; s12[C + c] = 0.0; }
}

struct PoolGeom {
    int C, H, W, Ho, Wo, s;
};

__device__ __forceinline__ float leaky(float z, float slope) { return z > 0.f ? z : z * slope; }

// out[b,c,ho,wo] = max over the 3x3 window of leaky(y * scale + shift); arg = kh * 3 + kw of the first maximum
// in scan order (ATen's rule: `val > maxval || isnan(val)`).
__global__ void __launch_bounds__(256) rgb_bn_act_pool_fwd_kernel(long long total, PoolGeom g, const float *__restrict__ y,
                                                                  const float *__restrict__ stats, float slope,
                                                                  float *__restrict__ out, int8_t *__restrict__ arg) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int wo = (int)(e % g.Wo);
        const long long t = e / g.Wo;
        const int ho = (int)(t % g.Ho);
        const long long plane = t / g.Ho;
        const int c = (int)(plane % g.C);
        const float sc = __ldg(stats + 2 * g.C + c), sh = __ldg(stats + 3 * g.C + c);
        const float *p = y + (size_t)plane * g.H * g.W;
        const int h0 = ho * g.s - 1, w0 = wo * g.s - 1;
        float best = -INFINITY;
        int bi = -1;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int h = h0 + kh;
            if (h < 0 || h >= g.H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int w = w0 + kw;
                if (w < 0 || w >= g.W) continue;
                const float v = leaky(__fmaf_rn(__ldg(p + (size_t)h * g.W + w), sc, sh), slope);
                if (bi < 0 || v > best || isnan(v)) { best = v; bi = kh * 3 + kw; }
            }
        }
        out[e] = best;
        arg[e] = (int8_t)bi;
    }
}

// Gradient w.r.t. the activated value at (h, w) of one plane: the sum of dout over the windows whose
// arg-max is this position.
__device__ __forceinline__ float pooled_grad(const PoolGeom &g, int h, int w, const float *__restrict__ dout,
                                             const int8_t *__restrict__ arg) {
    const int ho_lo = h >= 1 ? (h + g.s - 2) / g.s : 0, ho_hi = min(g.Ho - 1, (h + 1) / g.s);
    const int wo_lo = w >= 1 ? (w + g.s - 2) / g.s : 0, wo_hi = min(g.Wo - 1, (w + 1) / g.s);
    float acc = 0.f;
    for (int ho = ho_lo; ho <= ho_hi; ++ho) {
        const int kh = h - (ho * g.s - 1);
        for (int wo = wo_lo; wo <= wo_hi; ++wo) {
            const int kw = w - (wo * g.s - 1);
            const int o = ho * g.Wo + wo;
            if (__ldg(arg + o) == kh * 3 + kw) acc += __ldg(dout + o);
        }
    }
    return acc;
}

// s12[0][c] += sum dz, s12[1][c] += sum dz * yhat over one chunk of one plane (dz = pooled gradient * act')
__global__ void __launch_bounds__(RGB_THREADS) rgb_bwd_reduce_kernel(PoolGeom g, const float *__restrict__ y,
                                                                     const float *__restrict__ stats, float slope,
                                                                     const float *__restrict__ dout,
                                                                     const int8_t *__restrict__ arg, double *s12) {
    __shared__ float red[RGB_THREADS / 32];
    const int chunk = blockIdx.x, c = blockIdx.y, b = blockIdx.z;
    const int HW = g.H * g.W;
    const int start = chunk * RGB_CHUNK, stop = min(HW, start + RGB_CHUNK);
    const size_t plane = (size_t)b * g.C + c;
    const float *p = y + plane * HW;
    const float *dp = dout + plane * g.Ho * g.Wo;
    const int8_t *ap = arg + plane * g.Ho * g.Wo;
    const float mu = __ldg(stats + c), rs = __ldg(stats + g.C + c), sc = __ldg(stats + 2 * g.C + c),
                sh = __ldg(stats + 3 * g.C + c);
    float s1 = 0.f, s2 = 0.f;
    for (int i = start + threadIdx.x; i < stop; i += RGB_THREADS) {
        const int h = i / g.W, w = i - h * g.W;
        const float gsum = pooled_grad(g, h, w, dp, ap);
        const float yv = __ldg(p + i);
        const float dz = gsum * (__fmaf_rn(yv, sc, sh) > 0.f ? 1.f : slope);
        s1 += dz;
        s2 += dz * ((yv - mu) * rs);
    }
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) {
        atomicAdd(s12 + c, (double)s1);
        atomicAdd(s12 + g.C + c, (double)s2);
    }
}

// dy = scale * (dz - S1/n - yhat * S2/n)   (batch statistics)   or   scale * dz   (running statistics);
// the first block also emits dgamma = S2, dbeta = S1 in f32.
__global__ void __launch_bounds__(256) rgb_bwd_dx_kernel(long long total, PoolGeom g, int batch_stats, double inv_n,
                                                         const float *__restrict__ y, const float *__restrict__ stats,
                                                         float slope, const float *__restrict__ dout,
                                                         const int8_t *__restrict__ arg, const double *__restrict__ s12,
                                                         float *__restrict__ dy, float *dgamma, float *dbeta) {
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < g.C; c += 256) {
            if (dbeta != nullptr) dbeta[c] = (float)s12[c];
            if (dgamma != nullptr) dgamma[c] = (float)s12[g.C + c];
        }
    }
    const int HW = g.H * g.W;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int i = (int)(e % HW);
        const long long plane = e / HW;
        const int c = (int)(plane % g.C);
        const int h = i / g.W, w = i - h * g.W;
        const float gsum = pooled_grad(g, h, w, dout + (size_t)plane * g.Ho * g.Wo, arg + (size_t)plane * g.Ho * g.Wo);
        const float yv = __ldg(y + e);
        const float sc = __ldg(stats + 2 * g.C + c), sh = __ldg(stats + 3 * g.C + c);
        float dz = gsum * (__fmaf_rn(yv, sc, sh) > 0.f ? 1.f : slope);
        if (batch_stats) {
            const float mu = __ldg(stats + c), rs = __ldg(stats + g.C + c);
            const float s1n = (float)(s12[c] * inv_n), s2n = (float)(s12[g.C + c] * inv_n);
            dz = dz - s1n - ((yv - mu) * rs) * s2n;
        }
        dy[e] = sc * dz;
    }
}

static int pool_out(int n, int s) { return (n + 2 - 3) / s + 1; }

static bool pool_geom(PoolGeom &g, int C, int H, int W, int stride) {
    g.C = C; g.H = H; g.W = W; g.s = stride;
    g.Ho = pool_out(H, stride); g.Wo = pool_out(W, stride);
    return C >= 1 && H >= 1 && W >= 1 && (stride == 1 || stride == 2);
}

static int grid_for(long long total) {
    const long long g = (total + 255) / 256;
    return (int)(g < 148 * 32 ? g : 148 * 32);
}

}  // namespace i2p

extern "C" {

int i2p_rgb_num_chunks(int hw) { return (hw + i2p::RGB_CHUNK - 1) / i2p::RGB_CHUNK; }
int i2p_rgb_pool_out(int n, int stride) { return i2p::pool_out(n, stride); }

int i2p_rgb_bn_stats(int B, int C, int H, int W, const float *y, float *tile_stats, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(B >= 1 && C >= 1 && H >= 1 && W >= 1 && B <= 65535 && C <= 65535, "rgb_bn_stats: bad sizes");
    const int HW = H * W;
    dim3 grid(ceil_div(HW, RGB_CHUNK), C, B);
    rgb_bn_stats_kernel<<<grid, RGB_THREADS, 0, as_stream(stream)>>>(C, HW, y, tile_stats);
    return check_launch("rgb_bn_stats");
}

int i2p_rgb_bn_finalize(int C, int ntiles, const float *tile_stats, const float *gamma, const float *beta, float eps,
                        float momentum, float *running_mean, float *running_var, long long *num_batches_tracked,
                        float *stats, double *s12, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(C >= 1 && ntiles >= 1, "rgb_bn_finalize: bad sizes");
    rgb_bn_finalize_kernel<<<C, 128, 0, as_stream(stream)>>>(C, ntiles, tile_stats, gamma, beta, eps, momentum,
                                                             running_mean, running_var, num_batches_tracked, stats, s12);
    return check_launch("rgb_bn_finalize");
}

int i2p_rgb_bn_from_running(int C, const float *gamma, const float *beta, float eps, const float *running_mean,
                            const float *running_var, float *stats, double *s12, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(C >= 1 && running_mean != nullptr && running_var != nullptr, "rgb_bn_from_running: bad arguments");
    rgb_bn_from_running_kernel<<<ceil_div(C, 128), 128, 0, as_stream(stream)>>>(C, gamma, beta, eps, running_mean,
                                                                                running_var, stats, s12);
    return check_launch("rgb_bn_from_running");
}

int i2p_rgb_bn_act_pool_fwd(int B, int C, int H, int W, int stride, const float *y, const float *stats, float slope,
                            float *out, int8_t *arg, void *stream) {
    using namespace i2p;
    PoolGeom g;
    I2P_REQUIRE(pool_geom(g, C, H, W, stride) && B >= 1, "rgb_bn_act_pool_fwd: bad sizes (stride must be 1 or 2)");
    const long long total = (long long)B * C * g.Ho * g.Wo;
    rgb_bn_act_pool_fwd_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(total, g, y, stats, slope, out, arg);
    return check_launch("rgb_bn_act_pool_fwd");
}

int i2p_rgb_bn_act_pool_bwd(int B, int C, int H, int W, int stride, int batch_stats, const float *y, const float *stats,
                            float slope, const float *dout, const int8_t *arg, double *s12, float *dy, float *dgamma,
                            float *dbeta, void *stream) {
    using namespace i2p;
    PoolGeom g;
    I2P_REQUIRE(pool_geom(g, C, H, W, stride) && B >= 1 && B <= 65535 && C <= 65535,
                "rgb_bn_act_pool_bwd: bad sizes (stride must be 1 or 2)");
    const int HW = H * W;
    dim3 grid(ceil_div(HW, RGB_CHUNK), C, B);
    rgb_bwd_reduce_kernel<<<grid, RGB_THREADS, 0, as_stream(stream)>>>(g, y, stats, slope, dout, arg, s12);
    int rc = check_launch("rgb_bwd_reduce");
    if (rc != I2P_OK) return rc;
    const long long total = (long long)B * C * HW;
    rgb_bwd_dx_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(total, g, batch_stats, 1.0 / (double)((long long)B * HW),
                                                                      y, stats, slope, dout, arg, s12, dy, dgamma, dbeta);
    return check_launch("rgb_bwd_dx");
}
}
