// Gradient clipping + Adam on flat buffers (the reference step: torch.nn.utils.clip_grad_norm_(parameters, 10) then
// torch.optim.Adam(lr 1e-3, weight_decay 1e-4).step(), train20v2learn_wandb_proj.py:198-205, 481-483).
//
// The stock path costs ~560 launches per step when it is captured into a CUDA graph (the capturable Adam computes its
// bias corrections with one tiny pow kernel per parameter tensor) plus the clip's four.  Here all parameters, gradients
// and both moments are flat f32 buffers, and the step is two launches:
//   grad_sumsq : sum of squares of the (summed-over-ranks) gradient in f64, and the step counter / bias corrections
//   adam_step  : scale = (1 / world) * min(1, max_norm / (|g| / world + 1e-6)); g' = scale * g + wd * p;
//                m = b1 m + (1 - b1) g'; v = b2 v + (1 - b2) g'^2; p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// -- the formulas of torch.optim.Adam (L2 weight decay, not AdamW) and of clip_grad_norm_.
#include <math.h>

#include "common.cuh"

namespace i2p {

// state[0] = sum of squares (f64), state[1] = step count t (as f64), state[2] = 1 - b1^t, state[3] = 1 - b2^t,
// then the arrival ticket (u32 at byte 32) and the learning rate (f32 at byte 48): a launch argument would be frozen
// into a captured CUDA graph, whereas the reference trainer decays it every epoch (ExponentialLR(0.99),
// train20v2learn_wandb_proj.py:205, 524) -- replays read the current value from here.
__global__ void __launch_bounds__(256) grad_sumsq_kernel(long long n, const float *g, double *state, double b1, double b2,
                                                         unsigned *ticket) {
    double acc = 0.0;
    const long long n4 = n / 4;
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const float4 v = __ldg(g4 + i);
        acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4 * 4)) {
        const float v = g[n4 * 4 + threadIdx.x];
        acc += (double)v * v;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(FULL, acc, off);
    __shared__ double part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += part[w];
        atomicAdd(state, s);
        __threadfence();
        // the last block to arrive advances the step counter and publishes the bias corrections
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            const double t = state[1] + 1.0;
            state[1] = t;
            state[2] = 1.0 - pow(b1, t);
            state[3] = 1.0 - pow(b2, t);
            *ticket = 0;
        }
    }
}

__global__ void __launch_bounds__(256) adam_step_kernel(long long n, float *p, const float *g, float *m, float *v,
                                                        const double *state, float lr_arg, float b1, float b2, float eps,
                                                        float wd, float max_norm, float inv_world) {
    const float lr = lr_arg > 0.f ? lr_arg : *reinterpret_cast<const float *>(reinterpret_cast<const char *>(state) + 48);
    const float total_norm = (float)sqrt(state[0]) * inv_world;
    const float coef = fminf(max_norm / (total_norm + 1e-6f), 1.0f);
    const float gscale = max_norm > 0.f ? inv_world * coef : inv_world;
    const float step_size = lr / (float)state[2];
    const float inv_bc2_sqrt = (float)(1.0 / sqrt(state[3]));
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float pv = p[i];
        const float gv = __fmaf_rn(wd, pv, gscale * g[i]);
        const float mv = __fmaf_rn(b1, m[i], (1.f - b1) * gv);
        const float vv = __fmaf_rn(b2, v[i], (1.f - b2) * gv * gv);
        m[i] = mv;
        v[i] = vv;
        p[i] = pv - step_size * (mv / (sqrtf(vv) * inv_bc2_sqrt + eps));
    }
}

}  // namespace i2p

extern "C" {

int i2p_optim_state_bytes(void) { return 64; }
int i2p_optim_lr_offset(void) { return 48; }

int i2p_clip_adam_step(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, void *state,
                       float lr, float beta1, float beta2, float eps, float weight_decay, float max_norm, int world,
                       void *stream) {
    using namespace i2p;
    I2P_REQUIRE(n >= 0 && world >= 1 && state != nullptr, "clip_adam_step: bad arguments");
    I2P_REQUIRE((reinterpret_cast<uintptr_t>(grad) & 15) == 0 && (reinterpret_cast<uintptr_t>(state) & 7) == 0,
                "clip_adam_step: grad must be 16-byte and state 8-byte aligned");
    if (n == 0) return I2P_OK;
    cudaStream_t s = as_stream(stream);
    double *st = reinterpret_cast<double *>(state);
    unsigned *ticket = reinterpret_cast<unsigned *>(st + 4);
    cudaError_t e = cudaMemsetAsync(st, 0, sizeof(double), s);    // the sum of squares only: the step count persists
    if (e != cudaSuccess) { set_error("clip_adam_step: %s", cudaGetErrorString(e)); return I2P_ERR_CUDA; }
    const long long blocks = (n / 4 + 255) / 256;
    const int grid = (int)(blocks < 1 ? 1 : (blocks > 4 * 148 ? 4 * 148 : blocks));
    grad_sumsq_kernel<<<grid, 256, 0, s>>>(n, grad, st, (double)beta1, (double)beta2, ticket);
    int rc = check_launch("grad_sumsq");
    if (rc != I2P_OK) return rc;
    const long long b2 = (n + 255) / 256;
    adam_step_kernel<<<(int)(b2 > 8 * 148 ? 8 * 148 : b2), 256, 0, s>>>(n, param, grad, exp_avg, exp_avg_sq, st, lr, beta1, beta2,
                                                                         eps, weight_decay, max_norm, 1.0f / (float)world);
    return check_launch("adam_step");
}
}
