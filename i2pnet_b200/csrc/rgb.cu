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
//             statistics) -> out = maxpool(leaky((y - mean) * scale + beta)) from a shared-memory tile;
//   backward: each block re-derives the arg-max of the pooling windows touching its tile and scatters
//             their gradients in shared memory (no global atomics, no arg-max tensor, no zero-filled
//             scatter target); pass 0 writes dz = pooled gradient * act' and reduces the two batch-norm
//             sums S1 = sum dz, S2 = sum dz * yhat, pass 1 streams dy = scale (dz - S1/n - yhat S2/n) in place.
// Neither the normalised nor the activated tensor ever exists in HBM.  Algorithmic bytes per
// element of y: forward 4 (stats) + 4 + 4/s^2 (pool), backward (4 + 4/s^2 + 4) + (4 + 4 + 4).
#include <math.h>

#include "common.cuh"

namespace i2p {

constexpr int RGB_SPREAD = 16;   // the backward sums s12 are spread over 16 slots per channel: (2 * 16, C) f64
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

// stats (4, C) = mean, rstd, scale = gamma * rstd, beta; running statistics updated as nn.BatchNorm2d does in training
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
        stats[3 * C + c] = bt;
        if (running_mean != nullptr) {
            const double unbiased = n > 1.f ? (double)m2 / ((double)n - 1.0) : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
            if (c == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;
        }
        if (s12 != nullptr)
            for (int k = 0; k < 2 * RGB_SPREAD; ++k) s12[(size_t)k * C + c] = 0.0;
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
    stats[3 * C + c] = bt;
    if (s12 != nullptr)
        for (int k = 0; k < 2 * RGB_SPREAD; ++k) s12[(size_t)k * C + c] = 0.0;
}

struct PoolGeom {
    int C, H, W, Ho, Wo, s;
};

__device__ __forceinline__ float leaky(float z, float slope) { return z > 0.f ? z : z * slope; }
// (y - mean) * (gamma * rstd) + beta.  Not y * scale + shift with a pre-folded shift: the raw RGB input is
// not normalised, |mean| >> std in the first blocks, and the folded form's rounding error (relative to
// |mean * scale|) flips several times more LeakyReLU / arg-max decisions against f64 than this one.
__device__ __forceinline__ float norm(float y, float mean, float scale, float beta) { return __fmaf_rn(y - mean, scale, beta); }

// ---- tiled pooling kernels ------------------------------------------------------------------------
// One block owns a TH x TW tile of one (sample, channel) plane.  Forward: the normalised + activated
// tile (one halo row / column) goes to shared memory once, every output takes its 3x3 maximum from
// there.  Backward: the block rebuilds the pre-activations of its tile with a halo of two, re-derives
// the arg-max of every pooling window that touches the tile (identical arithmetic to the forward, so
// identical decisions: no arg-max tensor is stored) and adds that window's gradient to the arg-max
// position in a shared-memory accumulator -- windows of neighbouring tiles that reach into this one are
// recomputed here instead of exchanging anything between blocks.
constexpr int RGB_TH = 32, RGB_TW = 128;   // 4096 elements per block: 16 per thread amortise the block launch

// out[b,c,ho,wo] = max over the 3x3 window of leaky((y - mean) * scale + beta); first maximum in scan order wins
// (ATen: `val > maxval || isnan(val)`), which only matters for the backward pass.
template <int S>
__global__ void __launch_bounds__(256) rgb_pool_fwd_kernel(PoolGeom g, const float *__restrict__ y,
                                                          const float *__restrict__ stats, float slope,
                                                          float *__restrict__ out) {
    constexpr int ZH = RGB_TH + 2, ZW = RGB_TW + 2, OH = RGB_TH / S, OW = RGB_TW / S;
    __shared__ float zt[ZH][ZW + 1];
    const int plane = blockIdx.z, c = plane % g.C, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h0 = blockIdx.y * RGB_TH, w0 = blockIdx.x * RGB_TW;
    const float mu = __ldg(stats + c), sc = __ldg(stats + 2 * g.C + c), bt = __ldg(stats + 3 * g.C + c);
    const float *p = y + (size_t)plane * g.H * g.W;
    // Loops cover the part of the tile that lies inside the plane (the pyramid's coarse levels are 10 x 32 planes in a
    // 32 x 128 tile), a warp per row and a lane per column: no index division anywhere (the first version spent a
    // third of its instructions on idx / width, idx % width -- these kernels are instruction-bound, not HBM-bound).
    const int zh = min(ZH, g.H - h0 + 2), zw = min(ZW, g.W - w0 + 2);
    for (int hh = warp; hh < zh; hh += 8) {
        const int h = h0 - 1 + hh;
        const bool row_in = h >= 0 && h < g.H;
        const float *pr = p + (size_t)h * g.W + (w0 - 1);
        for (int ww = lane; ww < zw; ww += 32) {
            const int w = w0 - 1 + ww;
            zt[hh][ww] = (row_in && w >= 0 && w < g.W) ? leaky(norm(__ldg(pr + ww), mu, sc, bt), slope) : -INFINITY;
        }
    }
    __syncthreads();
    const int ho0 = h0 / S, wo0 = w0 / S;
    float *o = out + (size_t)plane * g.Ho * g.Wo;
    const int oh_n = min(OH, g.Ho - ho0), ow_n = min(OW, g.Wo - wo0);
    for (int oh = warp; oh < oh_n; oh += 8) {
        float *orow = o + (size_t)(ho0 + oh) * g.Wo + wo0;
        for (int ow = lane; ow < ow_n; ow += 32) {
            float best = -INFINITY;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float v = zt[oh * S + kh][ow * S + kw];
                    best = (v > best || v != v) ? v : best;
                }
            orow[ow] = best;
        }
    }
}

// Pass 0: dz = pooled gradient * act' for every element of the tile, written to dy; s12[slot][0][c] += sum dz,
// s12[slot][1][c] += sum dz * yhat.  Pass 1 (rgb_bn_bwd_apply_kernel) then streams dy <- scale * (dz - S1/n - yhat S2/n).
template <int S>
__global__ void __launch_bounds__(256, 4) rgb_pool_bwd_kernel(PoolGeom g, const float *__restrict__ y, const float *__restrict__ stats,
                                                          float slope, const float *__restrict__ dout, double *s12,
                                                          float *__restrict__ dy) {
    constexpr int ZH = RGB_TH + 4, ZW = RGB_TW + 4;
    constexpr int NWH = RGB_TH / S + (S == 1 ? 2 : 1), NWW = RGB_TW / S + (S == 1 ? 2 : 1);   // windows touching the tile
    __shared__ float zt[ZH][ZW + 1];          // ACTIVATED value, origin (h0 - 2, w0 - 2), -inf outside the plane
    __shared__ float dzt[RGB_TH][RGB_TW];     // gradient w.r.t. the activated value
    __shared__ float red[RGB_THREADS / 32];
    const int plane = blockIdx.z, c = plane % g.C, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h0 = blockIdx.y * RGB_TH, w0 = blockIdx.x * RGB_TW;
    const float mu = __ldg(stats + c), rs = __ldg(stats + g.C + c), sc = __ldg(stats + 2 * g.C + c),
                bt = __ldg(stats + 3 * g.C + c);
    const float *p = y + (size_t)plane * g.H * g.W;
    // (as in the forward kernel: only the part of the tile inside the plane is touched; a warp per row, a lane per column)
    const int th_n = min(RGB_TH, g.H - h0), tw_n = min(RGB_TW, g.W - w0);
    const int zh = min(ZH, th_n + 4), zw = min(ZW, tw_n + 4);
    for (int hh = warp; hh < zh; hh += 8) {
        const int h = h0 - 2 + hh;
        const bool row_in = h >= 0 && h < g.H;
        const float *pr = p + (size_t)h * g.W + (w0 - 2);
        for (int ww = lane; ww < zw; ww += 32) {
            const int w = w0 - 2 + ww;
            // the activation is monotonic: the arg-max of the activated window is taken on the same values as the forward
            zt[hh][ww] = (row_in && w >= 0 && w < g.W) ? leaky(norm(__ldg(pr + ww), mu, sc, bt), slope) : -INFINITY;
        }
    }
    for (int r = warp; r < th_n; r += 8)
        for (int cc = lane; cc < tw_n; cc += 32) dzt[r][cc] = 0.f;
    __syncthreads();
    // every pooling window that overlaps the tile: window (i, j) is output (ho_first + i, wo_first + j) and its
    // top-left input element sits at tile coordinates (S == 1 ? i : 2 i + 1, ...)
    const int ho_first = S == 1 ? h0 - 1 : h0 / 2, wo_first = S == 1 ? w0 - 1 : w0 / 2;
    const float *dp = dout + (size_t)plane * g.Ho * g.Wo;
    const int nwh = min(NWH, th_n / S + 2), nww = min(NWW, tw_n / S + 2);
    for (int i = warp; i < nwh; i += 8) {
        const int ho = ho_first + i;
        if (ho < 0 || ho >= g.Ho) continue;
        const int th = S == 1 ? i : 2 * i + 1;
        const float *drow = dp + (size_t)ho * g.Wo + wo_first;
        for (int j = lane; j < nww; j += 32) {
            const int wo = wo_first + j;
            if (wo < 0 || wo >= g.Wo) continue;
            const int tw = S == 1 ? j : 2 * j + 1;
            float best = -INFINITY;
            int bh = 0, bw = 0;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float v = zt[th + kh][tw + kw];        // padding is -inf: it never wins (v > -inf is false for it)
                    if (v > best || v != v) { best = v; bh = kh; bw = kw; }
                }
            const int r = th + bh - 2, cc = tw + bw - 2;   // arg-max position in core-tile coordinates
            if (r >= 0 && r < RGB_TH && cc >= 0 && cc < RGB_TW) atomicAdd(&dzt[r][cc], __ldg(drow + j));
        }
    }
    __syncthreads();
    float s1 = 0.f, s2 = 0.f;
    for (int r = warp; r < th_n; r += 8) {
        const size_t row = (size_t)plane * g.H * g.W + (size_t)(h0 + r) * g.W + w0;
        for (int cc = lane; cc < tw_n; cc += 32) {
            const float dz = dzt[r][cc] * (zt[r + 2][cc + 2] > 0.f ? 1.f : slope);     // sign of the activated value = sign of z
            const float yhat = (__ldg(y + row + cc) - mu) * rs;
            s1 += dz;
            s2 += dz * yhat;
            dy[row + cc] = dz;
        }
    }
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) {
        const int slot = (blockIdx.x + blockIdx.y * gridDim.x + blockIdx.z) % RGB_SPREAD;
        atomicAdd(s12 + (size_t)(2 * slot) * g.C + c, (double)s1);
        atomicAdd(s12 + (size_t)(2 * slot + 1) * g.C + c, (double)s2);
    }
}

// dy (holding dz) <- scale * (dz - S1/n - yhat * S2/n) (batch statistics) or scale * dz (running statistics), in place;
// one (sample, channel) plane per blockIdx.y, float4 where the plane size allows; block (0, c) also emits
// dgamma = S2, dbeta = S1 in f32.
__global__ void __launch_bounds__(256) rgb_bn_bwd_apply_kernel(int C, int HW, int batch_stats, double inv_n,
                                                              const float *__restrict__ y, const float *__restrict__ stats,
                                                              const double *__restrict__ s12, float *dy, float *dgamma,
                                                              float *dbeta) {
    __shared__ float sums[2];
    const int plane = blockIdx.y, c = plane % C;
    if (threadIdx.x == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int k = 0; k < RGB_SPREAD; ++k) { t1 += s12[(size_t)(2 * k) * C + c]; t2 += s12[(size_t)(2 * k + 1) * C + c]; }
        sums[0] = (float)(t1 * inv_n);
        sums[1] = (float)(t2 * inv_n);
        if (blockIdx.x == 0 && plane < C) {
            if (dbeta != nullptr) dbeta[c] = (float)t1;
            if (dgamma != nullptr) dgamma[c] = (float)t2;
        }
    }
    __syncthreads();
    const float mu = __ldg(stats + c), rs = __ldg(stats + C + c), sc = __ldg(stats + 2 * C + c);
    const float s1n = batch_stats ? sums[0] : 0.f, s2n = batch_stats ? sums[1] : 0.f;
    const float *yp = y + (size_t)plane * HW;
    float *dp = dy + (size_t)plane * HW;
    if ((HW & 3) == 0) {
        for (int i = blockIdx.x * 256 + threadIdx.x; i < HW / 4; i += gridDim.x * 256) {
            const float4 yv = __ldg(reinterpret_cast<const float4 *>(yp) + i);
            float4 d = reinterpret_cast<float4 *>(dp)[i];
            d.x = sc * (d.x - s1n - ((yv.x - mu) * rs) * s2n);
            d.y = sc * (d.y - s1n - ((yv.y - mu) * rs) * s2n);
            d.z = sc * (d.z - s1n - ((yv.z - mu) * rs) * s2n);
            d.w = sc * (d.w - s1n - ((yv.w - mu) * rs) * s2n);
            reinterpret_cast<float4 *>(dp)[i] = d;
        }
    } else {
        for (int i = blockIdx.x * 256 + threadIdx.x; i < HW; i += gridDim.x * 256)
            dp[i] = sc * (dp[i] - s1n - ((__ldg(yp + i) - mu) * rs) * s2n);
    }
}

static int pool_out(int n, int s) { return (n + 2 - 3) / s + 1; }

static bool pool_geom(PoolGeom &g, int C, int H, int W, int stride) {
    g.C = C; g.H = H; g.W = W; g.s = stride;
    g.Ho = pool_out(H, stride); g.Wo = pool_out(W, stride);
    return C >= 1 && H >= 1 && W >= 1 && (stride == 1 || stride == 2);
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

int i2p_rgb_s12_slots(void) { return 2 * i2p::RGB_SPREAD; }

int i2p_rgb_bn_act_pool_fwd(int B, int C, int H, int W, int stride, const float *y, const float *stats, float slope,
                            float *out, void *stream) {
    using namespace i2p;
    PoolGeom g;
    I2P_REQUIRE(pool_geom(g, C, H, W, stride) && B >= 1 && (long long)B * C <= 65535,
                "rgb_bn_act_pool_fwd: bad sizes (stride must be 1 or 2, B * C <= 65535)");
    dim3 grid(ceil_div(W, RGB_TW), ceil_div(H, RGB_TH), B * C);
    if (stride == 1) rgb_pool_fwd_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(g, y, stats, slope, out);
    else rgb_pool_fwd_kernel<2><<<grid, 256, 0, as_stream(stream)>>>(g, y, stats, slope, out);
    return check_launch("rgb_bn_act_pool_fwd");
}

int i2p_rgb_bn_act_pool_bwd(int B, int C, int H, int W, int stride, int batch_stats, const float *y, const float *stats,
                            float slope, const float *dout, double *s12, float *dy, float *dgamma, float *dbeta,
                            void *stream) {
    using namespace i2p;
    PoolGeom g;
    I2P_REQUIRE(pool_geom(g, C, H, W, stride) && B >= 1 && (long long)B * C <= 65535,
                "rgb_bn_act_pool_bwd: bad sizes (stride must be 1 or 2, B * C <= 65535)");
    dim3 grid(ceil_div(W, RGB_TW), ceil_div(H, RGB_TH), B * C);
    const double inv_n = 1.0 / (double)((long long)B * H * W);
    cudaStream_t s = as_stream(stream);
    if (stride == 1) rgb_pool_bwd_kernel<1><<<grid, 256, 0, s>>>(g, y, stats, slope, dout, s12, dy);
    else rgb_pool_bwd_kernel<2><<<grid, 256, 0, s>>>(g, y, stats, slope, dout, s12, dy);
    int rc = check_launch("rgb_pool_bwd(reduce)");
    if (rc != I2P_OK) return rc;
    const int HW = H * W;
    const int per_plane = (HW / 4 + 255) / 256;
    dim3 grid2(per_plane < 1 ? 1 : (per_plane > 64 ? 64 : per_plane), B * C);
    rgb_bn_bwd_apply_kernel<<<grid2, 256, 0, s>>>(C, HW, batch_stats, inv_n, y, stats, s12, dy, dgamma, dbeta);
    return check_launch("rgb_pool_bwd(dx)");
}
}
