// Shared per-point MLP: y = BN_batchstats(x W^T + b) -> activation, chained, forward and backward.
//
// Reference formulation (src/projectPN/PPBackbone_center.py:10-46, used ~35 times per forward):
// permute to NCHW, cuDNN/cuBLAS 1x1 conv, BatchNorm2d over the batch statistics, in-place
// (Leaky)ReLU, permute back -- each a separate pass over a (B, N, K, C) tensor, and as many again
// in autograd's backward (plus a bias-gradient reduction that is identically zero under BN).
//
// This design keeps only the RAW pre-normalisation outputs y_l in HBM and folds everything else
// into the GEMM kernels:
//   forward, one launch per layer:  y_l = act(y_{l-1} * scale + shift) W_l^T + b_l, where the
//       previous layer's normalise + activation is applied while the A tile is staged into shared
//       memory; the epilogue emits per-tile (mean, M2) of y_l, merged exactly (Chan) by a
//       finalize kernel into mean / rstd / scale / shift -- no E[y^2]-E[y]^2 cancellation;
//   consumer:  act(y_L * scale + shift), optionally max-reduced over the K neighbours with the
//       arg-max kept for the backward pass;
//   backward, two launches per layer: dW_l (split over row chunks) and dX_l, both rebuilding
//       dY_l = gamma rstd (dz - S1/n - yhat S2/n) on the fly from (g_l, y_l); the dX epilogue already
//       accumulates the S1, S2 sums the NEXT (earlier) layer's batch-norm backward needs.
// All arithmetic is f32 FMA (the parity configuration: 1e-4 relative rules out TF32); tiles are
// 128 x {16,32,64} x 16 with 8 x {1,2,4} register blocking.
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace i2p {

constexpr int MLP_THREADS = 256;
constexpr int MLP_BM = 128;  // rows per tile
constexpr int MLP_BK = 16;
constexpr int MLP_MAXC = 512;  // widest layer supported by the per-channel shared tables

__device__ __forceinline__ float act_fwd(float z, float slope) { return z > 0.f ? z : z * slope; }
__device__ __forceinline__ float act_grad(float z, float slope) { return z > 0.f ? 1.f : slope; }

// -------------------------------------------------------------------------------------------
// forward GEMM with fused input transform and tile statistics
// -------------------------------------------------------------------------------------------
struct FwdArgs {
    int rows, cin, cout;
    const float *x, *in_scale, *in_shift;  // in_scale == nullptr: x is used as is
    float in_slope;
    const float *w, *bias;
    float *y, *tile_stats;  // tile_stats (ntiles, cout, 2) or nullptr
    int dbg;                // profiling aid (I2P_TC_DBG): bit0 no proxy fence, bit1 no MMA, bit2 no loads, bit3 no stores
};

template <int TN>
__device__ __forceinline__ void load_bvec(float (&bv)[TN], const float *p) {
    if (TN == 4) {
        const float4 v = *reinterpret_cast<const float4 *>(p);
        bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
    } else if (TN == 2) {
        const float2 v = *reinterpret_cast<const float2 *>(p);
        bv[0] = v.x; bv[1] = v.y;
    } else {
        bv[0] = p[0];
    }
}

template <int BN>
__global__ void __launch_bounds__(MLP_THREADS, BN == 16 ? 4 : (BN == 32 ? 3 : 2)) pw_linear_fwd_kernel(const FwdArgs a) {
    constexpr int TN = BN / 16, TM = 8, LDA = MLP_BM + 4, LDB = BN + 4;
    __shared__ __align__(16) float As[MLP_BK][LDA];
    __shared__ __align__(16) float Bs[MLP_BK][LDB];
    __shared__ float red[16][BN];
    __shared__ float colmean[BN];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int r0 = blockIdx.x * MLP_BM, n0 = blockIdx.y * BN;
    const int lk = tid & 15, lm = tid >> 4;  // loader coordinates: 16 consecutive k per row
    const bool has_tf = a.in_scale != nullptr;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // Software pipeline: fetch() only ISSUES the global loads of the next k-tile (raw values, no
    // dependent arithmetic); stage() applies the input transform and writes shared memory one
    // iteration later, so the load latency is covered by the FMAs of the current tile.
    float pa[MLP_BM / 16], pb[BN / 16], psc = 1.f, psh = 0.f;
    auto fetch = [&](int k0) {
        const int k = k0 + lk;
        const bool kin = k < a.cin;
        if (has_tf && kin) { psc = __ldg(a.in_scale + k); psh = __ldg(a.in_shift + k); }
#pragma unroll
        for (int i = 0; i < MLP_BM / 16; ++i) {
            const int r = r0 + lm + 16 * i;
            pa[i] = (kin && r < a.rows) ? __ldg(a.x + (size_t)r * a.cin + k) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < BN / 16; ++i) {
            const int n = n0 + lm + 16 * i;
            pb[i] = (kin && n < a.cout) ? __ldg(a.w + (size_t)n * a.cin + k) : 0.f;
        }
    };
    auto stage = [&](int k0) {
        const bool kin = k0 + lk < a.cin;
#pragma unroll
        for (int i = 0; i < MLP_BM / 16; ++i) {
            float v = pa[i];
            if (has_tf) v = (kin && r0 + lm + 16 * i < a.rows) ? act_fwd(__fmaf_rn(v, psc, psh), a.in_slope) : 0.f;
            As[lk][lm + 16 * i] = v;
        }
#pragma unroll
        for (int i = 0; i < BN / 16; ++i) Bs[lk][lm + 16 * i] = pb[i];
    };

    fetch(0);
    for (int k0 = 0; k0 < a.cin; k0 += MLP_BK) {
        __syncthreads();
        stage(k0);
        __syncthreads();
        if (k0 + MLP_BK < a.cin) fetch(k0 + MLP_BK);
#pragma unroll
        for (int k = 0; k < MLP_BK; ++k) {
            float av[TM], bv[TN];
            *reinterpret_cast<float4 *>(&av[0]) = *reinterpret_cast<const float4 *>(&As[k][ty * TM]);
            *reinterpret_cast<float4 *>(&av[4]) = *reinterpret_cast<const float4 *>(&As[k][ty * TM + 4]);
            load_bvec<TN>(bv, &Bs[k][tx * TN]);
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = __fmaf_rn(av[i], bv[j], acc[i][j]);
        }
    }

    // epilogue: bias, store, tile statistics
    const int vr = min(MLP_BM, a.rows - r0);  // valid rows in this tile
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx * TN + j;
        const float b = (n < a.cout && a.bias != nullptr) ? __ldg(a.bias + n) : 0.f;
#pragma unroll
        for (int i = 0; i < TM; ++i) acc[i][j] += b;
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int r = r0 + ty * TM + i;
        if (r < a.rows) {
            float *dst = a.y + (size_t)r * a.cout + n0 + tx * TN;
            if (TN == 4 && n0 + tx * TN + 3 < a.cout && (a.cout & 3) == 0) {
                *reinterpret_cast<float4 *>(dst) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j)
                    if (n0 + tx * TN + j < a.cout) dst[j] = acc[i][j];
            }
        }
    }
    if (a.tile_stats == nullptr) return;
    // two-pass within the tile: column mean, then sum of squared deviations
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < TM; ++i) s += (ty * TM + i < vr) ? acc[i][j] : 0.f;
        red[ty][tx * TN + j] = s;
    }
    __syncthreads();
    if (tid < BN) {
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < 16; ++t) s += red[t][tid];
        colmean[tid] = s / (float)vr;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const float mu = colmean[tx * TN + j];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const float d = acc[i][j] - mu;
            s += (ty * TM + i < vr) ? d * d : 0.f;
        }
        red[ty][tx * TN + j] = s;
    }
    __syncthreads();
    if (tid < BN && n0 + tid < a.cout) {
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < 16; ++t) s += red[t][tid];
        float *ts = a.tile_stats + ((size_t)blockIdx.x * a.cout + n0 + tid) * 2;
        ts[0] = colmean[tid];
        ts[1] = s;
    }
}

// -------------------------------------------------------------------------------------------
// forward GEMM on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulator in TMEM)
// -------------------------------------------------------------------------------------------
// Same contract as pw_linear_fwd_kernel.  One CTA owns a 128-row x BN-column output tile; the K
// loop runs in chunks of 32 through a two-stage shared-memory ring.  All 256 threads fetch the raw
// operands of chunk c+1 from global memory, apply the fused input transform, split every element
// into tf32 (hi, lo) and store both into the UMMA canonical layout (umma.cuh); one elected thread
// then issues 4 k-steps x 3 MMAs (hi*hi, lo*hi, hi*lo) and commits them to the stage's mbarrier, so
// the tensor core works on chunk c while the CTA stages chunk c+1.  The epilogue pulls the
// accumulator out of TMEM (tcgen05.ld), adds the bias, parks the tile in shared memory, writes it
// with coalesced 128-bit stores and reduces the per-tile batch-norm statistics from it.
constexpr int TC_BK = 32;

// Byte stride between the k-column blocks of a [ROWS x 32] operand tile: ROWS * 16 plus one 16-byte
// pad, so that 32 lanes writing 32 consecutive k of one row hit 32 different banks (the staging
// loads are coalesced along k; without the pad every store would be an 8-way bank conflict).
template <int ROWS>
struct TcTile {
    static constexpr uint32_t LBO = ROWS * 16 + 16, SBO = 128, BYTES = (TC_BK / 4) * LBO;
};

constexpr int TC_THREADS = 256;   // 8 warps: lane = k within the chunk, warp = row phase

// Occupancy is the latency-hiding mechanism here, exactly as in the FMA kernels: a CTA is light
// (one 50 KB shared-memory stage, 64 TMEM columns, <= 64 registers per thread) so that four of them
// share an SM -- while one waits for its global loads or drains its epilogue, the others stage
// operands and keep the tensor core busy.
template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 4) pw_linear_fwd_tc_kernel(const FwdArgs a) {
    using TA = TcTile<MLP_BM>;
    using TB = TcTile<BN>;
    constexpr int NA = MLP_BM / 8, NB = BN / 8;
    extern __shared__ __align__(1024) unsigned char tc_smem[];   // A hi | A lo | B hi | B lo
    __shared__ uint64_t mma_bar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * BN, r0 = blockIdx.y * MLP_BM;   // column tiles of one row tile are adjacent CTAs
    const bool has_tf = a.in_scale != nullptr;
    if (warp == 0) umma::tmem_alloc(&tmem_slot, BN < 32 ? 32 : BN);
    if (tid == 0) {
        umma::mbar_init(&mma_bar, 1);
        umma::mbar_fence_init();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_d = tmem_slot;
    const uint32_t smem_base = umma::smem_u32(tc_smem);
    constexpr uint32_t idesc = umma::instr_desc_tf32(BN);

    // staging coordinates: row = warp + 8 i  ->  (row >> 3) * 128 + (row & 7) * 16 = i * 128 + warp * 16
    const uint32_t koff_a = (uint32_t)(lane >> 2) * TA::LBO + (uint32_t)(lane & 3) * 4u + (uint32_t)warp * 16u;
    const uint32_t koff_b = (uint32_t)(lane >> 2) * TB::LBO + (uint32_t)(lane & 3) * 4u + (uint32_t)warp * 16u;
    const int nchunks = (a.cin + TC_BK - 1) / TC_BK;
    const float *xbase = a.x + (size_t)(r0 + warp) * a.cin + lane;
    const float *wbase = a.w + (size_t)(n0 + warp) * a.cin + lane;

    // Interior tiles (all 128 rows, all BN channels valid) take unguarded loads; the activation is
    // max(z, slope * z), valid for the slopes in use (0 <= slope <= 1: ReLU, LeakyReLU(0.1), identity).
    const bool interior = r0 + MLP_BM <= a.rows && n0 + BN <= a.cout;
    const int astride = 8 * a.cin;   // elements between this thread's consecutive rows
    float ra[NA], rb[NB], psc = 1.f, psh = 0.f;
    auto fetch = [&](int c) {   // issue the global loads of chunk c; nothing depends on them yet
        const int k = c * TC_BK + lane;
        const bool kin = k < a.cin;
        const float *xp = xbase + c * TC_BK, *wp = wbase + c * TC_BK;
        if (has_tf && kin) { psc = __ldg(a.in_scale + k); psh = __ldg(a.in_shift + k); }
        if (a.dbg & 4) {
#pragma unroll
            for (int i = 0; i < NA; ++i) ra[i] = (float)(i + lane);
#pragma unroll
            for (int i = 0; i < NB; ++i) rb[i] = (float)(i - lane);
        } else if (interior && kin) {
#pragma unroll
            for (int i = 0; i < NA; ++i) ra[i] = __ldg(xp + i * astride);
#pragma unroll
            for (int i = 0; i < NB; ++i) rb[i] = __ldg(wp + i * astride);
        } else {
#pragma unroll
            for (int i = 0; i < NA; ++i) ra[i] = (kin && r0 + warp + 8 * i < a.rows) ? __ldg(xp + i * astride) : 0.f;
#pragma unroll
            for (int i = 0; i < NB; ++i) rb[i] = (kin && n0 + warp + 8 * i < a.cout) ? __ldg(wp + i * astride) : 0.f;
        }
    };

    fetch(0);
    for (int c = 0; c < nchunks; ++c) {
        if (c >= 1) umma::mbar_wait(&mma_bar, (uint32_t)(c - 1) & 1u);   // the tensor core is done with the stage
        const bool kin = c * TC_BK + lane < a.cin;
        if (has_tf) {
            if (interior && kin) {
#pragma unroll
                for (int i = 0; i < NA; ++i) {
                    const float z = __fmaf_rn(ra[i], psc, psh);
                    ra[i] = fmaxf(z, z * a.in_slope);
                }
            } else {
#pragma unroll
                for (int i = 0; i < NA; ++i) {
                    const float z = __fmaf_rn(ra[i], psc, psh);
                    ra[i] = (kin && r0 + warp + 8 * i < a.rows) ? fmaxf(z, z * a.in_slope) : 0.f;
                }
            }
        }
        if (!(a.dbg & 8))
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            uint32_t hi, lo;
            umma::split_tf32(ra[i], hi, lo);
            *reinterpret_cast<uint32_t *>(tc_smem + koff_a + i * 128) = hi;
            *reinterpret_cast<uint32_t *>(tc_smem + TA::BYTES + koff_a + i * 128) = lo;
        }
        if (!(a.dbg & 8))
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            uint32_t hi, lo;
            umma::split_tf32(rb[i], hi, lo);
            *reinterpret_cast<uint32_t *>(tc_smem + 2 * TA::BYTES + koff_b + i * 128) = hi;
            *reinterpret_cast<uint32_t *>(tc_smem + 2 * TA::BYTES + TB::BYTES + koff_b + i * 128) = lo;
        }
        if (!(a.dbg & 1)) umma::fence_smem_to_async();
        __syncthreads();
        if (tid == 0) {
            umma::fence_after_sync();
            const uint32_t sa = smem_base, sb = sa + 2 * TA::BYTES;
#pragma unroll
            for (int j = 0; j < TC_BK / 8; ++j) {
                const uint64_t a_hi = umma::smem_desc(sa + j * 2 * TA::LBO, TA::LBO, TA::SBO);
                const uint64_t a_lo = umma::smem_desc(sa + TA::BYTES + j * 2 * TA::LBO, TA::LBO, TA::SBO);
                const uint64_t b_hi = umma::smem_desc(sb + j * 2 * TB::LBO, TB::LBO, TB::SBO);
                const uint64_t b_lo = umma::smem_desc(sb + TB::BYTES + j * 2 * TB::LBO, TB::LBO, TB::SBO);
                if (a.dbg & 2) continue;
                umma::mma_tf32(tmem_d, a_hi, b_hi, idesc, c > 0 || j > 0);
                umma::mma_tf32(tmem_d, a_lo, b_hi, idesc, true);
                umma::mma_tf32(tmem_d, a_hi, b_lo, idesc, true);
            }
            umma::commit(&mma_bar);
        }
        if (c + 1 < nchunks) fetch(c + 1);   // global loads in flight while the tensor core runs
    }
    umma::mbar_wait(&mma_bar, (uint32_t)(nchunks - 1) & 1u);
    umma::fence_after_sync();
    __syncthreads();   // every thread is past its last shared-memory operand write / wait

    // epilogue: TMEM -> registers (+bias) -> shared tile [128][BN+4]
    float *tile = reinterpret_cast<float *>(tc_smem);
    constexpr int LDT = BN + 4;
    {
        const int row = (warp & 3) * 32 + lane;                 // a warp can only read its own 32 TMEM lanes
        constexpr int CQ = BN / 2;                              // warps 0-3: first half of the columns, 4-7: second
        const int cbase = (warp >> 2) * CQ;
#pragma unroll
        for (int c0 = 0; c0 < CQ; c0 += 16) {
            float v[16];
            umma::tmem_ld16(tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cbase + c0), v);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = n0 + cbase + c0 + j;
                tile[row * LDT + cbase + c0 + j] = v[j] + ((n < a.cout && a.bias != nullptr) ? __ldg(a.bias + n) : 0.f);
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_d, BN < 32 ? 32 : BN);

    const int vr = min(MLP_BM, a.rows - r0);
    for (int e = tid; e < MLP_BM * (BN / 4); e += TC_THREADS) {   // coalesced 128-bit stores
        const int row = e / (BN / 4), c4 = e % (BN / 4);
        if (row < vr && n0 + c4 * 4 + 3 < a.cout)
            *reinterpret_cast<float4 *>(a.y + (size_t)(r0 + row) * a.cout + n0 + c4 * 4) =
                *reinterpret_cast<const float4 *>(&tile[row * LDT + c4 * 4]);
    }
    if (a.tile_stats == nullptr) return;
    // per-tile (mean, M2): 4 threads per column, 32 rows each, merged with Chan's formula
    {
        const int col = tid >> 2, part = tid & 3;
        float n = 0.f, mu = 0.f, m2 = 0.f;
        if (col < BN) {
            const int rb0 = part * 32, re0 = min(vr, rb0 + 32);
            float sum = 0.f;
            for (int r = rb0; r < re0; ++r) sum += tile[r * LDT + col];
            n = (float)max(re0 - rb0, 0);
            mu = n > 0.f ? sum / n : 0.f;
            for (int r = rb0; r < re0; ++r) { const float d = tile[r * LDT + col] - mu; m2 += d * d; }
        }
        // BN * 4 = 256 threads: the four parts of a column are adjacent lanes
        for (int off = 1; off < 4; off <<= 1) {
            const float n2 = __shfl_xor_sync(FULL, n, off), mu2 = __shfl_xor_sync(FULL, mu, off),
                        s2 = __shfl_xor_sync(FULL, m2, off);
            const float tot = n + n2;
            if (tot > 0.f) {
                const float d = mu2 - mu;
                const float nmu = mu + d * (n2 / tot);
                m2 = m2 + s2 + d * d * n * (n2 / tot);
                mu = nmu;
                n = tot;
            }
        }
        if (col < BN && part == 0 && n0 + col < a.cout) {
            float *ts = a.tile_stats + ((size_t)blockIdx.y * a.cout + n0 + col) * 2;
            ts[0] = mu;
            ts[1] = m2;
        }
    }
}

// One 128-thread block per channel merges the per-tile (count, mean, M2) triples (Chan et al.).
// The merge formula only ever ADDS non-negative terms to M2, so f32 is accurate here; the final
// variance -> rstd step is done in f64.
__device__ __forceinline__ void chan_merge(float &n, float &mu, float &m2, float n2, float mu2, float s2) {
    if (n2 > 0.f) {
        const float tot = n + n2, d = mu2 - mu, f = n2 / tot;
        mu += d * f;
        m2 += s2 + d * d * n * f;
        n = tot;
    }
}

// Merge of the per-tile (mean, M2) pairs of one channel: two passes over the tile statistics (they sit in L2) --
// the row-weighted mean first, then M2 = sum_t M2_t + n_t (mean_t - mean)^2 -- with f64 accumulators.  No serial chain
// of Chan merges (a division each): 45 of these kernels sit between the layers of the step.
constexpr int FIN_THREADS = 512;

__global__ void __launch_bounds__(FIN_THREADS) bn_finalize_kernel(int rows, int cout, int ntiles, const float *tile_stats,
                                                                  const float *gamma, const float *beta, float eps,
                                                                  float *mean, float *rstd, float *scale, float *shift) {
    __shared__ double part[FIN_THREADS / 32];
    __shared__ double bcast;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x;
    const float2 *ts = reinterpret_cast<const float2 *>(tile_stats);
    auto block_sum = [&](double v) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(FULL, v, off);
        if (lane == 0) part[warp] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < FIN_THREADS / 32; ++w) t += part[w];
            bcast = t;
        }
        __syncthreads();
        return bcast;
    };
    // (four independent loads per thread in flight: up to 7200 tiles per channel at the first set-abstraction level)
    double s = 0.0;
    for (int t0 = threadIdx.x; t0 < ntiles; t0 += 4 * FIN_THREADS) {
        float m[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + u * FIN_THREADS;
            m[u] = t < ntiles ? (float)min(MLP_BM, rows - t * MLP_BM) * __ldg(ts + (size_t)t * cout + c).x : 0.f;
        }
        s += ((double)m[0] + (double)m[1]) + ((double)m[2] + (double)m[3]);
    }
    const double mu_d = block_sum(s) / (double)rows;
    const float mu = (float)mu_d;
    double q = 0.0;
    for (int t0 = threadIdx.x; t0 < ntiles; t0 += 4 * FIN_THREADS) {
        float m[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + u * FIN_THREADS;
            if (t < ntiles) {
                const float2 v = __ldg(ts + (size_t)t * cout + c);
                const float d = v.x - mu;
                m[u] = __fmaf_rn((float)min(MLP_BM, rows - t * MLP_BM) * d, d, v.y);
            } else {
                m[u] = 0.f;
            }
        }
        q += ((double)m[0] + (double)m[1]) + ((double)m[2] + (double)m[3]);
    }
    const double m2 = block_sum(q);
    if (threadIdx.x == 0) {
        const double var = m2 / (double)rows;  // biased, as BatchNorm normalises
        const float r = (float)(1.0 / sqrt(var + (double)eps));
        const float g = gamma != nullptr ? gamma[c] : 1.f, b = beta != nullptr ? beta[c] : 0.f;
        mean[c] = mu;
        rstd[c] = r;
        scale[c] = g * r;
        shift[c] = b - mu * g * r;
    }
}

// out = act(y * scale + shift)
__global__ void __launch_bounds__(256) bn_act_kernel(long long total4, int c4, const float4 *y, const float4 *scale,
                                                     const float4 *shift, float slope, float4 *out) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total4; e += (long long)gridDim.x * 256) {
        const int c = (int)(e % c4);
        const float4 v = __ldg(y + e), sc = __ldg(scale + c), sh = __ldg(shift + c);
        out[e] = make_float4(act_fwd(__fmaf_rn(v.x, sc.x, sh.x), slope), act_fwd(__fmaf_rn(v.y, sc.y, sh.y), slope),
                             act_fwd(__fmaf_rn(v.z, sc.z, sh.z), slope), act_fwd(__fmaf_rn(v.w, sc.w, sh.w), slope));
    }
}

// out[g,c] = max_k act(y[g,k,c] * scale + shift); arg[g,c] = first k attaining it
__global__ void __launch_bounds__(256) bn_act_maxk_kernel(long long total, int k, int c, const float *y,
                                                          const float *scale, const float *shift, float slope,
                                                          float *out, int32_t *arg) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long g = e / c;
        const int ch = (int)(e - g * c);
        const float sc = __ldg(scale + ch), sh = __ldg(shift + ch);
        const float *p = y + (size_t)g * k * c + ch;
        float best = -INFINITY;
        int bi = 0;
        for (int j = 0; j < k; ++j) {
            const float v = act_fwd(__fmaf_rn(__ldg(p + (size_t)j * c), sc, sh), slope);
            if (v > best) { best = v; bi = j; }
        }
        out[e] = best;
        arg[e] = bi;
    }
}

// -------------------------------------------------------------------------------------------
// backward
// -------------------------------------------------------------------------------------------
// Where the gradient w.r.t. a layer's ACTIVATED output comes from: a dense (rows, c) tensor, or the
// (groups, c) gradient of a max-over-k output routed through its arg-max.
struct GradSrc {
    const float *dense;  // (rows, c) or nullptr
    const float *dout;   // (groups, c)
    const int32_t *arg;  // (groups, c)
    int k;
    int kshift;          // log2(k) when k is a power of two (every group size of the model), else -1
    // group of row r (rows fit 31 bits): a shift, or one 32-bit division -- never the emulated 64-bit one, which at one
    // division per staged element dominated the skinny dW / dX kernels
    __device__ __forceinline__ unsigned group_of(long long r) const {
        return kshift >= 0 ? ((unsigned)r >> kshift) : ((unsigned)r / (unsigned)k);
    }
    __device__ __forceinline__ float at(long long r, int ch, int c) const {
        if (dense != nullptr) return __ldg(dense + (size_t)r * c + ch);
        const unsigned g = group_of(r);
        const int kk = (int)((unsigned)r - g * (unsigned)k);
        return __ldg(arg + (size_t)g * c + ch) == kk ? __ldg(dout + (size_t)g * c + ch) : 0.f;
    }
};

struct BnRef {  // a layer's raw output and its batch-norm constants
    const float *y, *mean, *rstd, *scale, *shift;
    float slope;
};

// s12[0][c] = sum_r dz, s12[1][c] = sum_r dz * yhat   (dz = g * act'(z)); f64 accumulators, pre-zeroed.
// Dense source: a thread owns FOUR consecutive channels (128-bit loads of g and y) and every (256 / (cw / 4))-th row of
// the block's slice, four rows per iteration (eight independent 128-bit loads in flight); the first version read one
// float per thread and iteration with two loads in flight and was latency-bound at a fifth of the HBM rate.
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(long long rows, int c, GradSrc gs, BnRef bn,
                                                            int rows_per_block, double *s12) {
    __shared__ float r1[256 * 4], r2[256 * 4];
    const int cw = c < 64 ? c : 64;  // columns per block (c is a multiple of 16)
    const long long rb = (long long)blockIdx.x * rows_per_block;
    const long long re = min(rows, rb + rows_per_block);
    if (gs.dense != nullptr) {
        const int cq = cw / 4, lanes = 256 / cq;               // threads per row, rows in flight per pass
        const int q = threadIdx.x % cq, rl = threadIdx.x / cq;
        const int ch = blockIdx.y * cw + q * 4;
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        if (rl < lanes) {
            const float4 mu = *reinterpret_cast<const float4 *>(bn.mean + ch), rs = *reinterpret_cast<const float4 *>(bn.rstd + ch);
            const float4 sc = *reinterpret_cast<const float4 *>(bn.scale + ch), sh = *reinterpret_cast<const float4 *>(bn.shift + ch);
            const float4 *yp = reinterpret_cast<const float4 *>(bn.y + ch), *gp = reinterpret_cast<const float4 *>(gs.dense + ch);
            const size_t ld4 = (size_t)c / 4;
            for (long long r0 = rb + rl; r0 < re; r0 += 4LL * lanes) {
                float4 yv[4], gv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const long long r = r0 + (long long)u * lanes;
                    const bool in = r < re;
                    yv[u] = in ? __ldg(yp + (size_t)r * ld4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    gv[u] = in ? __ldg(gp + (size_t)r * ld4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float dz;
                    dz = gv[u].x * act_grad(__fmaf_rn(yv[u].x, sc.x, sh.x), bn.slope); s1[0] += dz; s2[0] += dz * ((yv[u].x - mu.x) * rs.x);
                    dz = gv[u].y * act_grad(__fmaf_rn(yv[u].y, sc.y, sh.y), bn.slope); s1[1] += dz; s2[1] += dz * ((yv[u].y - mu.y) * rs.y);
                    dz = gv[u].z * act_grad(__fmaf_rn(yv[u].z, sc.z, sh.z), bn.slope); s1[2] += dz; s2[2] += dz * ((yv[u].z - mu.z) * rs.z);
                    dz = gv[u].w * act_grad(__fmaf_rn(yv[u].w, sc.w, sh.w), bn.slope); s1[3] += dz; s2[3] += dz * ((yv[u].w - mu.w) * rs.w);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { r1[threadIdx.x * 4 + j] = s1[j]; r2[threadIdx.x * 4 + j] = s2[j]; }
        __syncthreads();
        if (threadIdx.x < cw && blockIdx.y * cw + threadIdx.x < c) {
            const int qq = threadIdx.x / 4, j = threadIdx.x % 4;
            float t1 = 0.f, t2 = 0.f;
            for (int l = 0; l < lanes; ++l) { t1 += r1[(l * cq + qq) * 4 + j]; t2 += r2[(l * cq + qq) * 4 + j]; }
            atomicAdd(s12 + blockIdx.y * cw + threadIdx.x, (double)t1);
            atomicAdd(s12 + c + blockIdx.y * cw + threadIdx.x, (double)t2);
        }
        return;
    }
    const int lanes = 256 / cw;
    const int col = threadIdx.x % cw, rl = threadIdx.x / cw;
    const int ch = blockIdx.y * cw + col;
    float s1 = 0.f, s2 = 0.f;
    if (ch < c && rl < lanes) {
        const float mu = bn.mean[ch], rs = bn.rstd[ch], sc = bn.scale[ch], sh = bn.shift[ch];
        // only the arg-max element of each group carries gradient
        const long long gb = rb / gs.k, ge = (re + gs.k - 1) / gs.k;
        for (long long g = gb + rl; g < ge; g += lanes) {
            const long long r = g * gs.k + __ldg(gs.arg + (size_t)g * c + ch);
            if (r < rb || r >= re) continue;
            const float yv = __ldg(bn.y + (size_t)r * c + ch);
            const float dz = __ldg(gs.dout + (size_t)g * c + ch) * act_grad(__fmaf_rn(yv, sc, sh), bn.slope);
            s1 += dz;
            s2 += dz * ((yv - mu) * rs);
        }
    }
    r1[threadIdx.x] = s1;
    r2[threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.x < cw && ch < c) {
        float t1 = 0.f, t2 = 0.f;
        for (int l = 0; l < lanes; ++l) { t1 += r1[l * cw + threadIdx.x]; t2 += r2[l * cw + threadIdx.x]; }
        atomicAdd(s12 + ch, (double)t1);
        atomicAdd(s12 + c + ch, (double)t2);
    }
}

// Per-channel constants that turn (g, y) into dY, staged once per block.
struct DyTables {
    float scale[MLP_MAXC], shift[MLP_MAXC], mean[MLP_MAXC], rstd[MLP_MAXC], s1n[MLP_MAXC], s2n[MLP_MAXC];
};

__device__ __forceinline__ void load_dy_tables(DyTables &t, const BnRef &bn, const double *s12, int c, long long rows) {
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        t.scale[i] = bn.scale[i]; t.shift[i] = bn.shift[i]; t.mean[i] = bn.mean[i]; t.rstd[i] = bn.rstd[i];
        t.s1n[i] = (float)(s12[i] / (double)rows);
        t.s2n[i] = (float)(s12[c + i] / (double)rows);
    }
}

__device__ __forceinline__ float make_dy(const DyTables &t, float g, float yv, int ch, float slope) {
    const float dz = g * act_grad(__fmaf_rn(yv, t.scale[ch], t.shift[ch]), slope);
    const float yhat = (yv - t.mean[ch]) * t.rstd[ch];
    return t.scale[ch] * (dz - t.s1n[ch] - yhat * t.s2n[ch]);
}

// Raw operands of one dY element, as fetched (no dependent arithmetic until stage time).
struct RawDy {
    float g, y;
    int arg;  // max-k source: the arg-max row offset of this (group, channel); dense source: unused
};

__device__ __forceinline__ RawDy fetch_dy(const GradSrc &gs, const float *y, long long r, int ch, int c) {
    RawDy v;
    v.y = __ldg(y + (size_t)r * c + ch);
    if (gs.dense != nullptr) {
        v.g = __ldg(gs.dense + (size_t)r * c + ch);
        v.arg = INT_MIN;  // dense source: always taken
    } else {
        const unsigned g = gs.group_of(r);
        v.g = __ldg(gs.dout + (size_t)g * c + ch);
        v.arg = __ldg(gs.arg + (size_t)g * c + ch) - (int)((unsigned)r - g * (unsigned)gs.k);  // 0 <=> this row is the arg-max
    }
    return v;
}

__device__ __forceinline__ float finish_dy(const DyTables &t, const RawDy &v, int ch, float slope) {
    const float g = (v.arg == 0 || v.arg == INT_MIN) ? v.g : 0.f;
    return make_dy(t, g, v.y, ch, slope);
}

struct DxArgs {
    int rows, cin, cout;
    GradSrc gs;     // gradient w.r.t. this layer's activated output (rows, cout)
    BnRef bn;       // this layer
    const double *s12;
    const float *w;  // (cout, cin)
    float *dx;       // (rows, cin): gradient w.r.t. this layer's input (the previous activation)
    BnRef prev;      // previous layer (prev.y == nullptr: the input is not a BN output)
    double *prev_s12;
};

// dx[r,i] = sum_o dY[r,o] W[o,i]; epilogue accumulates the previous layer's S1/S2.
template <int BN>
__global__ void __launch_bounds__(MLP_THREADS, BN == 16 ? 4 : (BN == 32 ? 3 : 2)) pw_linear_bwd_dx_kernel(const DxArgs a) {
    constexpr int TN = BN / 16, TM = 8, LDA = MLP_BM + 4, LDB = BN + 4;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    DyTables &tab = *reinterpret_cast<DyTables *>(dyn_smem);
    __shared__ __align__(16) float As[MLP_BK][LDA];
    __shared__ __align__(16) float Bs[MLP_BK][LDB];
    __shared__ float red1[16][BN], red2[16][BN];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int r0 = blockIdx.x * MLP_BM, n0 = blockIdx.y * BN;
    const int lk = tid & 15, lm = tid >> 4;     // A loader: 16 consecutive k (= cout) per row
    const int bn_ = tid % BN, bk = tid / BN;    // B loader: consecutive n (= cin), 256/BN k per pass
    constexpr int BPASS = MLP_BK * BN / MLP_THREADS;
    load_dy_tables(tab, a.bn, a.s12, a.cout, a.rows);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    RawDy pa[MLP_BM / 16];
    float pb[BPASS];
    auto fetch = [&](int k0) {
        const int k = k0 + lk;
#pragma unroll
        for (int i = 0; i < MLP_BM / 16; ++i) {
            const int r = r0 + lm + 16 * i;
            if (k < a.cout && r < a.rows) pa[i] = fetch_dy(a.gs, a.bn.y, r, k, a.cout);
        }
#pragma unroll
        for (int i = 0; i < BPASS; ++i) {
            const int kk = k0 + bk + i * (MLP_THREADS / BN), n = n0 + bn_;
            pb[i] = (kk < a.cout && n < a.cin) ? __ldg(a.w + (size_t)kk * a.cin + n) : 0.f;
        }
    };
    auto stage = [&](int k0) {
        const int k = k0 + lk;
#pragma unroll
        for (int i = 0; i < MLP_BM / 16; ++i)
            As[lk][lm + 16 * i] = (k < a.cout && r0 + lm + 16 * i < a.rows) ? finish_dy(tab, pa[i], k, a.bn.slope) : 0.f;
#pragma unroll
        for (int i = 0; i < BPASS; ++i) Bs[bk + i * (MLP_THREADS / BN)][bn_] = pb[i];
    };

    fetch(0);
    for (int k0 = 0; k0 < a.cout; k0 += MLP_BK) {
        __syncthreads();   // (first iteration: also publishes the dY tables)
        stage(k0);
        __syncthreads();
        if (k0 + MLP_BK < a.cout) fetch(k0 + MLP_BK);
#pragma unroll
        for (int k = 0; k < MLP_BK; ++k) {
            float av[TM], bv[TN];
            *reinterpret_cast<float4 *>(&av[0]) = *reinterpret_cast<const float4 *>(&As[k][ty * TM]);
            *reinterpret_cast<float4 *>(&av[4]) = *reinterpret_cast<const float4 *>(&As[k][ty * TM + 4]);
            load_bvec<TN>(bv, &Bs[k][tx * TN]);
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = __fmaf_rn(av[i], bv[j], acc[i][j]);
        }
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int r = r0 + ty * TM + i;
        if (r < a.rows)
#pragma unroll
            for (int j = 0; j < TN; ++j)
                if (n0 + tx * TN + j < a.cin) a.dx[(size_t)r * a.cin + n0 + tx * TN + j] = acc[i][j];
    }
    if (a.prev.y == nullptr) return;
    // the previous layer's batch-norm backward sums, from the tile still in registers
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const int ch = n0 + tx * TN + j;
        float s1 = 0.f, s2 = 0.f;
        if (ch < a.cin) {
            const float mu = a.prev.mean[ch], rs = a.prev.rstd[ch], sc = a.prev.scale[ch], sh = a.prev.shift[ch];
            float yv[TM];
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const int r = r0 + ty * TM + i;
                yv[i] = r < a.rows ? __ldg(a.prev.y + (size_t)r * a.cin + ch) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                if (r0 + ty * TM + i < a.rows) {
                    const float dz = acc[i][j] * act_grad(__fmaf_rn(yv[i], sc, sh), a.prev.slope);
                    s1 += dz;
                    s2 += dz * ((yv[i] - mu) * rs);
                }
            }
        }
        red1[ty][tx * TN + j] = s1;
        red2[ty][tx * TN + j] = s2;
    }
    __syncthreads();
    if (tid < BN && n0 + tid < a.cin) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int t = 0; t < 16; ++t) { t1 += red1[t][tid]; t2 += red2[t][tid]; }
        atomicAdd(a.prev_s12 + n0 + tid, (double)t1);
        atomicAdd(a.prev_s12 + a.cin + n0 + tid, (double)t2);
    }
}

struct DwArgs {
    int rows, cin, cout, rows_per_block;
    GradSrc gs;
    BnRef bn;
    const double *s12;
    const float *x;  // (rows, cin): raw input, or the previous layer's raw output when prev.scale != nullptr
    BnRef prev;
    float *dw;       // (cout, cin), pre-zeroed, accumulated with atomics
};

// dW[o,i] += sum_{r in chunk} dY[r,o] * A_prev[r,i].  BMo x BNi output tile, 16 x 16 threads with a
// (BMo/16) x (BNi/16) register block, BK rows per step (more rows per step for skinny tiles, so that
// every thread keeps several independent loads in flight).
template <int BMo, int BNi, int BK>
__global__ void __launch_bounds__(MLP_THREADS, 2) pw_linear_bwd_dw_kernel(const DwArgs a) {
    constexpr int TMo = BMo / 16, TNi = BNi / 16, LDA = BMo + 4, LDB = BNi + 4;
    constexpr int APT = BK * BMo / MLP_THREADS, BPT = BK * BNi / MLP_THREADS;  // elements per thread per step
    constexpr int AROWS = MLP_THREADS / BMo, BROWS = MLP_THREADS / BNi;       // rows covered per loader pass
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    DyTables &tab = *reinterpret_cast<DyTables *>(dyn_smem);
    __shared__ __align__(16) float As[BK][LDA];  // [row r][o]
    __shared__ __align__(16) float Bs[BK][LDB];  // [row r][i]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    // tile indices fastest: CTAs sharing a chunk of rows run together and re-read it from L2
    const int m0 = blockIdx.x * BMo, n0 = blockIdx.y * BNi;
    const long long rb = (long long)blockIdx.z * a.rows_per_block;
    const long long re = min((long long)a.rows, rb + a.rows_per_block);
    const int ac = tid % BMo, ar = tid / BMo, bc = tid % BNi, br = tid / BNi;
    const bool has_tf = a.prev.scale != nullptr;
    load_dy_tables(tab, a.bn, a.s12, a.cout, a.rows);
    float bsc = 1.f, bsh = 0.f;
    if (has_tf && n0 + bc < a.cin) { bsc = __ldg(a.prev.scale + n0 + bc); bsh = __ldg(a.prev.shift + n0 + bc); }

    float acc[TMo][TNi];
#pragma unroll
    for (int i = 0; i < TMo; ++i)
#pragma unroll
        for (int j = 0; j < TNi; ++j) acc[i][j] = 0.f;

    RawDy pa[APT];
    float pb[BPT];
    auto fetch = [&](long long rr0) {
#pragma unroll
        for (int i = 0; i < APT; ++i) {
            const long long r = rr0 + ar + AROWS * i;
            if (r < re && m0 + ac < a.cout) pa[i] = fetch_dy(a.gs, a.bn.y, r, m0 + ac, a.cout);
        }
#pragma unroll
        for (int i = 0; i < BPT; ++i) {
            const long long r = rr0 + br + BROWS * i;
            pb[i] = (r < re && n0 + bc < a.cin) ? __ldg(a.x + (size_t)r * a.cin + n0 + bc) : 0.f;
        }
    };
    auto stage = [&](long long rr0) {
#pragma unroll
        for (int i = 0; i < APT; ++i)
            As[ar + AROWS * i][ac] = (rr0 + ar + AROWS * i < re && m0 + ac < a.cout)
                                         ? finish_dy(tab, pa[i], m0 + ac, a.bn.slope) : 0.f;
#pragma unroll
        for (int i = 0; i < BPT; ++i) {
            float v = pb[i];
            if (has_tf) v = (rr0 + br + BROWS * i < re && n0 + bc < a.cin) ? act_fwd(__fmaf_rn(v, bsc, bsh), a.prev.slope) : 0.f;
            Bs[br + BROWS * i][bc] = v;
        }
    };

    fetch(rb);
    for (long long rr0 = rb; rr0 < re; rr0 += BK) {
        __syncthreads();
        stage(rr0);
        __syncthreads();
        if (rr0 + BK < re) fetch(rr0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float av[TMo], bv[TNi];
            load_bvec<TMo>(av, &As[k][ty * TMo]);
            load_bvec<TNi>(bv, &Bs[k][tx * TNi]);
#pragma unroll
            for (int i = 0; i < TMo; ++i)
#pragma unroll
                for (int j = 0; j < TNi; ++j) acc[i][j] = __fmaf_rn(av[i], bv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < TMo; ++i) {
        const int o = m0 + ty * TMo + i;
        if (o < a.cout)
#pragma unroll
            for (int j = 0; j < TNi; ++j) {
                const int ci = n0 + tx * TNi + j;
                if (ci < a.cin) atomicAdd(a.dw + (size_t)o * a.cin + ci, acc[i][j]);
            }
    }
}

template <int BMo, int BNi, int BK>
static int launch_dw(DwArgs a, cudaStream_t s) {
    const int tiles = ceil_div(a.cout, BMo) * ceil_div(a.cin, BNi);
    // about four waves of CTAs (2 resident per SM) over the 148 SMs, at least 8 steps per CTA
    const int chunks = (8 * 148 + tiles - 1) / tiles;
    int rpb = (a.rows + chunks - 1) / chunks;
    rpb = ((rpb + BK - 1) / BK) * BK;
    if (rpb < 8 * BK) rpb = 8 * BK;
    a.rows_per_block = rpb;
    dim3 grid(ceil_div(a.cout, BMo), ceil_div(a.cin, BNi), ceil_div(a.rows, rpb));
    pw_linear_bwd_dw_kernel<BMo, BNi, BK><<<grid, MLP_THREADS, wgrad_smem(pw_linear_bwd_dw_kernel<BMo, BNi, BK>, (int)sizeof(DyTables)), s>>>(a);
    return check_launch("pw_linear_bwd_dw");
}

// dW for skinny layers (cin <= 16; cout tile CO = 16 / 32 / 64): the output tile is so small that the kernel above
// spends two shared-memory loads per two FMAs.  Here a thread owns a 4 x 4 block of the tile (two LDS.128 per 16 FMAs),
// CO threads cover the tile, and the 256 / CO groups of the CTA take every (256 / CO)-th staged row; the groups' partial
// tiles meet in shared memory at the end.  Same operands, staging and epilogue (atomics into dW) as the general kernel.
template <int CO>
__global__ void __launch_bounds__(MLP_THREADS, CO == 64 ? 2 : 3) pw_linear_bwd_dw_skinny_kernel(const DwArgs a) {
    constexpr int CI = 16, BK = 64, T = CO, KG = MLP_THREADS / T, LDA = CO + 4, LDB = CI + 4;
    constexpr int APT = BK * CO / MLP_THREADS, BPT = BK * CI / MLP_THREADS;
    constexpr int AROWS = MLP_THREADS / CO, BROWS = MLP_THREADS / CI;
    constexpr int STAGE_FLOATS = BK * LDA + BK * LDB, RED_FLOATS = KG * CO * CI;
    constexpr int BUF_FLOATS = STAGE_FLOATS > RED_FLOATS ? STAGE_FLOATS : RED_FLOATS;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    DyTables &tab = *reinterpret_cast<DyTables *>(dyn_smem);
    __shared__ __align__(16) float buf[BUF_FLOATS];
    float (*As)[LDA] = reinterpret_cast<float (*)[LDA]>(buf);             // [row][o]
    float (*Bs)[LDB] = reinterpret_cast<float (*)[LDB]>(buf + BK * LDA);  // [row][i]
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * CO;
    const long long rb = (long long)blockIdx.y * a.rows_per_block;
    const long long re = min((long long)a.rows, rb + a.rows_per_block);
    const int ac = tid % CO, ar = tid / CO, bc = tid % CI, br = tid / CI;
    const int t = tid % T, kg = tid / T, to = t / (CI / 4), ti = t % (CI / 4);
    const bool has_tf = a.prev.scale != nullptr;
    load_dy_tables(tab, a.bn, a.s12, a.cout, a.rows);
    float bsc = 1.f, bsh = 0.f;
    if (has_tf && bc < a.cin) { bsc = __ldg(a.prev.scale + bc); bsh = __ldg(a.prev.shift + bc); }

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    RawDy pa[APT];
    float pb[BPT];
    auto fetch = [&](long long rr0) {
#pragma unroll
        for (int i = 0; i < APT; ++i) {
            const long long r = rr0 + ar + AROWS * i;
            if (r < re && m0 + ac < a.cout) pa[i] = fetch_dy(a.gs, a.bn.y, r, m0 + ac, a.cout);
        }
#pragma unroll
        for (int i = 0; i < BPT; ++i) {
            const long long r = rr0 + br + BROWS * i;
            pb[i] = (r < re && bc < a.cin) ? __ldg(a.x + (size_t)r * a.cin + bc) : 0.f;
        }
    };
    auto stage = [&](long long rr0) {
#pragma unroll
        for (int i = 0; i < APT; ++i)
            As[ar + AROWS * i][ac] = (rr0 + ar + AROWS * i < re && m0 + ac < a.cout)
                                         ? finish_dy(tab, pa[i], m0 + ac, a.bn.slope) : 0.f;
#pragma unroll
        for (int i = 0; i < BPT; ++i) {
            float v = pb[i];
            if (has_tf) v = (rr0 + br + BROWS * i < re && bc < a.cin) ? act_fwd(__fmaf_rn(v, bsc, bsh), a.prev.slope) : 0.f;
            Bs[br + BROWS * i][bc] = v;
        }
    };

    fetch(rb);
    for (long long rr0 = rb; rr0 < re; rr0 += BK) {
        __syncthreads();
        stage(rr0);
        __syncthreads();
        if (rr0 + BK < re) fetch(rr0 + BK);
#pragma unroll
        for (int k = kg; k < BK; k += KG) {
            const float4 av = *reinterpret_cast<const float4 *>(&As[k][4 * to]);
            const float4 bv = *reinterpret_cast<const float4 *>(&Bs[k][4 * ti]);
            const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(a4[i], b4[j], acc[i][j]);
        }
    }
    __syncthreads();   // staging buffers are dead: they become the partial-tile exchange
    float *red = buf;  // [kg][o][i]
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4 *>(&red[(kg * CO + 4 * to + i) * CI + 4 * ti]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    __syncthreads();
    for (int e = tid; e < CO * CI; e += MLP_THREADS) {
        float sum = 0.f;
#pragma unroll
        for (int g = 0; g < KG; ++g) sum += red[g * CO * CI + e];
        const int o = m0 + e / CI, ci = e % CI;
        if (o < a.cout && ci < a.cin) atomicAdd(a.dw + (size_t)o * a.cin + ci, sum);
    }
}

template <int CO>
static int launch_dw_skinny(DwArgs a, cudaStream_t s) {
    constexpr int BK = 64;
    const int tiles = ceil_div(a.cout, CO);
    const int chunks = (4 * 148 + tiles - 1) / tiles;     // two waves of CTAs at 2 per SM
    int rpb = (a.rows + chunks - 1) / chunks;
    rpb = ((rpb + BK - 1) / BK) * BK;
    if (rpb < 4 * BK) rpb = 4 * BK;
    a.rows_per_block = rpb;
    dim3 grid(tiles, ceil_div(a.rows, rpb));
    pw_linear_bwd_dw_skinny_kernel<CO><<<grid, MLP_THREADS, wgrad_smem(pw_linear_bwd_dw_skinny_kernel<CO>, (int)sizeof(DyTables)), s>>>(a);
    return check_launch("pw_linear_bwd_dw(skinny)");
}

static int pick_bn(int n) { return n <= 16 ? 16 : (n <= 32 ? 32 : 64); }

// Which shared-MLP GEMMs run on the tensor cores (tcgen05, 3xTF32; kernels in mlp_tc.cu), as a bit mask:
// 1 forward, 2 dX, 4 dW, 8 the first-generation forward kernel of this file (kept for A/B measurements).
// 16 = SWIZZLE_128B operand tiles + bulk-copied weights in the forward / dX kernels.  Default 23 (= 7 | 16);
// environment variable I2P_MLP_TC overrides; 0 = f32 FMA kernels everywhere.
static int g_mlp_tc = -1;
static int mlp_tc_mask() {
    if (g_mlp_tc < 0) {
        const char *e = getenv("I2P_MLP_TC");
        g_mlp_tc = (e == nullptr) ? 119 : atoi(e);
    }
    return g_mlp_tc;
}

}  // namespace i2p

extern "C" {

void i2p_set_mlp_tensor_cores(int mask) { i2p::g_mlp_tc = mask < 0 ? 0 : mask; }
int i2p_get_mlp_tensor_cores(void) { return i2p::mlp_tc_mask(); }

int i2p_pw_num_tiles(int rows) { return (rows + i2p::MLP_BM - 1) / i2p::MLP_BM; }

int i2p_pw_linear_fwd(int rows, int cin, int cout, const float *x, const float *in_scale, const float *in_shift,
                      float in_slope, const float *w, const float *bias, float *y, float *tile_stats, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(rows >= 0 && cin >= 1 && cout >= 1, "pw_linear_fwd: bad sizes");
    if (rows == 0) return I2P_OK;
    static int dbg = -1;
    if (dbg < 0) { const char *e = getenv("I2P_TC_DBG"); dbg = e ? atoi(e) : 0; }
    FwdArgs a{rows, cin, cout, x, in_scale, in_shift, in_slope, w, bias, y, tile_stats, dbg};
    cudaStream_t s = as_stream(stream);
    if ((mlp_tc_mask() & 8) && cout >= 64 && cout % 64 == 0 && cin <= MLP_MAXC) {
        // first-generation tcgen05 path: 128 x 64 tiles, scalar staging, four CTAs per SM
        constexpr int smem = 2 * TcTile<MLP_BM>::BYTES + 2 * TcTile<64>::BYTES;
        static bool once = false;
        if (!once) { cudaFuncSetAttribute(pw_linear_fwd_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); once = true; }
        pw_linear_fwd_tc_kernel<64><<<dim3(cout / 64, ceil_div(rows, MLP_BM)), TC_THREADS, smem, s>>>(a);
        return check_launch("pw_linear_fwd(tcgen05)");
    }
    const int bn = pick_bn(cout);
    dim3 grid(ceil_div(rows, MLP_BM), ceil_div(cout, bn));
    if (bn == 16) pw_linear_fwd_kernel<16><<<grid, MLP_THREADS, 0, s>>>(a);
    else if (bn == 32) pw_linear_fwd_kernel<32><<<grid, MLP_THREADS, 0, s>>>(a);
    else pw_linear_fwd_kernel<64><<<grid, MLP_THREADS, 0, s>>>(a);
    return check_launch("pw_linear_fwd");
}

int i2p_bn_finalize(int rows, int cout, const float *tile_stats, const float *gamma, const float *beta, float eps,
                    float *mean, float *rstd, float *scale, float *shift, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(rows >= 1 && cout >= 1, "bn_finalize: bad sizes");
    bn_finalize_kernel<<<cout, FIN_THREADS, 0, as_stream(stream)>>>(rows, cout, ceil_div(rows, MLP_BM), tile_stats,
                                                                        gamma, beta, eps, mean, rstd, scale, shift);
    return check_launch("bn_finalize");
}

int i2p_bn_act(long long rows, int c, const float *y, const float *scale, const float *shift, float slope, float *out,
               void *stream) {
    using namespace i2p;
    I2P_REQUIRE(rows >= 0 && c >= 4 && (c & 3) == 0, "bn_act: channel count must be a multiple of 4");
    if (rows == 0) return I2P_OK;
    const long long total4 = rows * (c / 4);
    const long long g = (total4 + 255) / 256;
    bn_act_kernel<<<(int)(g < 148 * 16 ? g : 148 * 16), 256, 0, as_stream(stream)>>>(
        total4, c / 4, reinterpret_cast<const float4 *>(y), reinterpret_cast<const float4 *>(scale),
        reinterpret_cast<const float4 *>(shift), slope, reinterpret_cast<float4 *>(out));
    return check_launch("bn_act");
}

int i2p_bn_act_maxk(long long groups, int k, int c, const float *y, const float *scale, const float *shift,
                    float slope, float *out, int32_t *arg, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(groups >= 0 && k >= 1 && c >= 1, "bn_act_maxk: bad sizes");
    if (groups == 0) return I2P_OK;
    const long long total = groups * c, g = (total + 255) / 256;
    bn_act_maxk_kernel<<<(int)(g < 148 * 16 ? g : 148 * 16), 256, 0, as_stream(stream)>>>(total, k, c, y, scale, shift,
                                                                                         slope, out, arg);
    return check_launch("bn_act_maxk");
}

static i2p::GradSrc make_gs(const float *g_dense, const float *dout, const int32_t *arg, int k) {
    i2p::GradSrc gs;
    gs.dense = g_dense; gs.dout = dout; gs.arg = arg; gs.k = k > 0 ? k : 1;
    gs.kshift = -1;
    if ((gs.k & (gs.k - 1)) == 0) { gs.kshift = 0; while ((1 << gs.kshift) < gs.k) ++gs.kshift; }
    return gs;
}

int i2p_bn_bwd_reduce(long long rows, int c, const float *g_dense, const float *dout, const int32_t *arg, int k,
                      const float *y, const float *mean, const float *rstd, const float *scale, const float *shift,
                      float slope, double *s12, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(rows >= 0 && c >= 16 && c % 16 == 0 && c <= MLP_MAXC, "bn_bwd_reduce: c must be a multiple of 16, <= 512");
    I2P_REQUIRE(g_dense != nullptr || (dout != nullptr && arg != nullptr && k >= 1), "bn_bwd_reduce: no gradient source");
    I2P_REQUIRE(rows <= 0x7fffffffLL, "bn_bwd_reduce: at most 2^31 - 1 rows");
    if (rows == 0) return I2P_OK;
    const int cw = c < 64 ? c : 64;
    // about four waves of blocks; each block owns whole groups of k rows
    long long rpb = rows / (4 * 148 / ceil_div(c, cw) + 1) + 1;
    if (rpb < 64) rpb = 64;
    const int kk = g_dense == nullptr ? k : 1;
    rpb = ((rpb + kk - 1) / kk) * kk;
    dim3 grid(ceil_div(rows, rpb), ceil_div(c, cw));
    BnRef bn{y, mean, rstd, scale, shift, slope};
    bn_bwd_reduce_kernel<<<grid, 256, 0, as_stream(stream)>>>(rows, c, make_gs(g_dense, dout, arg, k), bn, (int)rpb, s12);
    return check_launch("bn_bwd_reduce");
}

int i2p_pw_linear_bwd_dx(int rows, int cin, int cout, const float *g_dense, const float *dout, const int32_t *arg,
                         int k, const float *y, const float *mean, const float *rstd, const float *scale,
                         const float *shift, float slope, const double *s12, const float *w, float *dx,
                         const float *prev_y, const float *prev_mean, const float *prev_rstd, const float *prev_scale,
                         const float *prev_shift, float prev_slope, double *prev_s12, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(rows >= 0 && cin >= 1 && cout >= 1 && cout <= MLP_MAXC, "pw_linear_bwd_dx: bad sizes");
    if (rows == 0) return I2P_OK;
    DxArgs a;
    a.rows = rows; a.cin = cin; a.cout = cout; a.gs = make_gs(g_dense, dout, arg, k);
    a.bn = BnRef{y, mean, rstd, scale, shift, slope};
    a.s12 = s12; a.w = w; a.dx = dx;
    a.prev = BnRef{prev_y, prev_mean, prev_rstd, prev_scale, prev_shift, prev_slope};
    a.prev_s12 = prev_s12;
    const int bn = pick_bn(cin);
    dim3 grid(ceil_div(rows, MLP_BM), ceil_div(cin, bn));
    const size_t smem = sizeof(DyTables);
    cudaStream_t s = as_stream(stream);
    if (bn == 16) pw_linear_bwd_dx_kernel<16><<<grid, MLP_THREADS, smem, s>>>(a);
    else if (bn == 32) pw_linear_bwd_dx_kernel<32><<<grid, MLP_THREADS, smem, s>>>(a);
    else pw_linear_bwd_dx_kernel<64><<<grid, MLP_THREADS, smem, s>>>(a);
    return check_launch("pw_linear_bwd_dx");
}

int i2p_pw_linear_bwd_dw(int rows, int cin, int cout, const float *g_dense, const float *dout, const int32_t *arg,
                         int k, const float *y, const float *mean, const float *rstd, const float *scale,
                         const float *shift, float slope, const double *s12, const float *x, const float *prev_scale,
                         const float *prev_shift, float prev_slope, float *dw, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(rows >= 0 && cin >= 1 && cout >= 1 && cout <= MLP_MAXC, "pw_linear_bwd_dw: bad sizes");
    if (rows == 0) return I2P_OK;
    DwArgs a;
    a.rows = rows; a.cin = cin; a.cout = cout; a.gs = make_gs(g_dense, dout, arg, k);
    a.bn = BnRef{y, mean, rstd, scale, shift, slope};
    a.s12 = s12; a.x = x;
    a.prev = BnRef{nullptr, nullptr, nullptr, prev_scale, prev_shift, prev_slope};
    a.dw = dw;
    cudaStream_t st = as_stream(stream);
    static int skinny = -1;   // I2P_DW_SKINNY=0: the general tile kernel for cin <= 16 as well (A/B measurements)
    if (skinny < 0) { const char *e = getenv("I2P_DW_SKINNY"); skinny = e ? atoi(e) : 1; }
    if (cin <= 16 && skinny) {
        if (cout <= 16) return launch_dw_skinny<16>(a, st);
        if (cout <= 32) return launch_dw_skinny<32>(a, st);
        return launch_dw_skinny<64>(a, st);
    }
    if (cout <= 16) return cin <= 16 ? launch_dw<16, 16, 64>(a, st) : launch_dw<16, 64, 32>(a, st);
    if (cout <= 32) return cin <= 16 ? launch_dw<32, 16, 32>(a, st) : launch_dw<32, 64, 32>(a, st);
    return cin <= 16 ? launch_dw<64, 16, 32>(a, st) : launch_dw<64, 64, 16>(a, st);
}
}

