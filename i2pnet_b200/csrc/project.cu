// Spherical range-image projection (reference: project_seq with rank=False,
// src/projectPN/utils.py:111-187 -- a dozen elementwise torch kernels, three zero-fills and a
// Python loop over the batch issuing three index_put_ per sample, with an unordered winner
// when two points share a cell).
// This design: two launches for the whole batch.  (1) every point computes its cell with the
// reference's f32 formulas and claims it with atomicMax on its own index, which makes the
// winner deterministic (highest index = the CPU index_put_ result); (2) every CELL pulls its
// winner's coordinates and features, or writes zeros -- so the images are written exactly once
// and never zero-filled separately.
#include <math.h>

#include "common.cuh"

namespace i2p {

constexpr int P_THREADS = 256;
constexpr int P_MAX_FEATS = 4;

struct ProjArgs {
    int n, H, W;
    float pi, inv_az, inv_vres, voff;
    const float *xyz;
    int nfeat;
    const float *feat[P_MAX_FEATS];
    int fdim[P_MAX_FEATS];
    float *xyz_proj;
    float *feat_proj[P_MAX_FEATS];
    int32_t *owner;
};

__device__ __forceinline__ long long clampll(long long v, long long lo, long long hi) {
    return v < lo ? lo : (v > hi ? hi : v);
}

__global__ void __launch_bounds__(P_THREADS) project_claim_kernel(const ProjArgs a, long long total) {
    const long long g = (long long)blockIdx.x * P_THREADS + threadIdx.x;
    if (g >= total) return;
    const int b = (int)(g / a.n), i = (int)(g % a.n);
    const float *p = a.xyz + g * 3;
    const float x = p[0], y = p[1], z = p[2];
    const float r = sqrtf(__fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x))));       // utils.py:144
    // torch divides by a Python scalar as a multiplication by its f32 reciprocal on CUDA
    const long long col = (long long)__fmul_rn(__fsub_rn(a.pi, atan2f(y, x)), a.inv_az);   // :147
    const float beta = asinf(__fdiv_rn(z, r));                                             // :150
    const long long row = (long long)a.H - (long long)__fadd_rn(__fmul_rn(beta, a.inv_vres), a.voff);  // :152
    const int rr = (int)clampll(row, 0, a.H - 1), cc = (int)clampll(col, 0, a.W - 1);      // :154-155
    atomicMax(a.owner + ((size_t)b * a.H + rr) * a.W + cc, i);
}

__global__ void __launch_bounds__(P_THREADS) project_fill_kernel(const ProjArgs a, long long cells) {
    const long long g = (long long)blockIdx.x * P_THREADS + threadIdx.x;
    if (g >= cells) return;
    const long long b = g / ((long long)a.H * a.W);
    const int o = a.owner[g];
    const long long src = b * a.n + o;
    float *q = a.xyz_proj + g * 3;
    if (o >= 0) {
        const float *p = a.xyz + src * 3;
        q[0] = p[0]; q[1] = p[1]; q[2] = p[2];
    } else {
        q[0] = 0.f; q[1] = 0.f; q[2] = 0.f;
    }
#pragma unroll
    for (int f = 0; f < P_MAX_FEATS; ++f) {
        if (f < a.nfeat) {
            const int d = a.fdim[f];
            float *fo = a.feat_proj[f] + g * d;
            const float *fi = a.feat[f] + src * d;
            for (int c = 0; c < d; ++c) fo[c] = o >= 0 ? fi[c] : 0.f;
        }
    }
}

}  // namespace i2p

extern "C" int i2p_project_seq(int b, int n, int H, int W, float fup_deg, float fdown_deg, const float *xyz,
                               int nfeat, const float *const *feats, const int *fdims, float *xyz_proj,
                               float *const *feat_projs, int32_t *owner, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(b >= 0 && n >= 0 && H >= 2 && W >= 1, "project_seq: bad sizes");
    I2P_REQUIRE(nfeat >= 0 && nfeat <= P_MAX_FEATS, "project_seq: at most %d feature arrays", P_MAX_FEATS);
    if (b == 0) return I2P_OK;
    // scalar set-up in double like the Python code (utils.py:125-139), cast to f32 where the
    // scalars meet f32 tensors
    const double deg2rad = M_PI / 180.0;
    const double az = 360.0 / W * deg2rad;
    const double down = (double)fdown_deg * deg2rad, up = (double)fup_deg * deg2rad;
    const double vres = (up - down) / (H - 1);
    const double voff = -down / vres;
    ProjArgs a{};
    a.n = n; a.H = H; a.W = W;
    a.pi = (float)M_PI;
    a.inv_az = 1.0f / (float)az;
    a.inv_vres = 1.0f / (float)vres;
    a.voff = (float)voff;
    a.xyz = xyz; a.nfeat = nfeat; a.xyz_proj = xyz_proj; a.owner = owner;
    for (int f = 0; f < nfeat; ++f) {
        a.feat[f] = feats[f]; a.fdim[f] = fdims[f]; a.feat_proj[f] = feat_projs[f];
    }
    cudaStream_t s = as_stream(stream);
    const long long cells = (long long)b * H * W, total = (long long)b * n;
    cudaError_t e = cudaMemsetAsync(owner, 0xff, cells * sizeof(int32_t), s);  // -1 = empty
    if (e != cudaSuccess) {
        set_error("project_seq: %s", cudaGetErrorString(e));
        return I2P_ERR_CUDA;
    }
    if (total > 0) {
        project_claim_kernel<<<ceil_div(total, P_THREADS), P_THREADS, 0, s>>>(a, total);
        int rc = check_launch("project_seq(claim)");
        if (rc) return rc;
    }
    project_fill_kernel<<<ceil_div(cells, P_THREADS), P_THREADS, 0, s>>>(a, cells);
    return check_launch("project_seq(fill)");
}
