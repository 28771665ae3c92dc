// Projection-window K-nearest select (reference kernel: fused_conv_select_k_gpu,
// src/projectPN/fused_conv_select/fused_conv_go.cu:11-240; launcher :243-264).
//
// Reference design: B blocks x 512 threads, one THREAD per centre, three 150-entry
// per-thread arrays in local memory, an O(K * kH*kW) scalar selection sort.
// This design: one WARP per centre, grid over all B*npoints centres.  The kH*kW window
// slots live in registers, slot t in lane t%32 register t/32 (<= 5 registers for the
// reference's 150-slot limit); an empty centre exits after a single load.  Valid slots are compacted (ballot + popc) into a per-warp
// shared list and every candidate computes its output position as its RANK in that list --
// independent compares, no serial chain.  Only when two valid distances are bit-equal does
// the emitted order depend on the reference's swap history; that (rare) case replays its
// selection sort step by step with two redux.sync reductions and three shuffles per step.
// Either way the result is bit-identical to the reference.  Results are staged in shared
// memory and written with coalesced stores.
#include "common.cuh"

namespace i2p {

constexpr int SEL_WARPS = 4;      // warps (centres) per block
constexpr int SEL_MAX_SLOTS = 160;  // 5 registers x 32 lanes >= the reference's 150
constexpr int SEL_MAX_K = 150;      // fused_conv_go.cu:52-53 arrays are [150]

struct SelectArgs {
    int batch, H, W, npoints, kH, kW, K, flag;
    float distance;
    int stride_h, stride_w, small_h, small_w;
    const float *xyz1, *xyz2;
    const int32_t *idx_n2, *random_hw;
    int out_w, stride_ch, stride_cw;  // regular centre grid when idx_n2 == nullptr
    unsigned kw_magic, ow_magic, sh_magic, sw_magic;  // ceil(2^32 / d) for d = kW, out_w, stride_h, stride_w:
                                                      // q = umulhi(n, magic) is exact for n, d < 2^16
    int64_t *sel_b, *sel_h, *sel_w;   // drop-in outputs (partial writes)
    float *sel_mask;
    int32_t *flat_idx;                // compact outputs (full writes)
    float *flat_mask;
};

// n / d through the precomputed magic (0 encodes d == 1)
__device__ __forceinline__ unsigned fastdiv(unsigned n, unsigned magic) { return magic == 0u ? n : __umulhi(n, magic); }

template <int R>
__device__ __forceinline__ float pick(const float (&a)[R], int r) {
    float v = a[0];
#pragma unroll
    for (int i = 1; i < R; ++i) v = (r == i) ? a[i] : v;
    return v;
}
template <int R>
__device__ __forceinline__ int pick(const int (&a)[R], int r) {
    int v = a[0];
#pragma unroll
    for (int i = 1; i < R; ++i) v = (r == i) ? a[i] : v;
    return v;
}

template <int R, bool FLAT>
__global__ void __launch_bounds__(SEL_WARPS * 32) select_k_kernel(const SelectArgs a) {
    __shared__ int s_hw[SEL_WARPS][SEL_MAX_K];               // packed (h << 16 | w), -1 = slot not written
    __shared__ unsigned s_d[SEL_WARPS][SEL_MAX_SLOTS];       // compacted valid candidates: distance bits
    __shared__ int s_c[SEL_WARPS][SEL_MAX_SLOTS + 1];        // ... their (h << 16 | w); [last] = the nearest

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.y;                                 // no integer division anywhere on this path:
    const int cn = blockIdx.x * SEL_WARPS + warp;             // four of them cost more than the rest of an
    if (cn >= a.npoints) return;                              // empty centre's work
    const int total = a.kH * a.kW;
    const int K = a.K;

    int sH, sW;
    if (a.idx_n2 != nullptr) {
        const int32_t *c = a.idx_n2 + ((size_t)b * a.npoints + cn) * 2;  // fused_conv_go.cu:65-66
        sH = __ldg(c);
        sW = __ldg(c + 1);
    } else {
        const unsigned ch = fastdiv((unsigned)cn, a.ow_magic);
        sH = (int)ch * a.stride_ch;
        sW = (cn - (int)ch * a.out_w) * a.stride_cw;
    }
    for (int k = lane; k < K; k += 32) s_hw[warp][k] = -1;

    const float *pc = a.xyz1 + (((size_t)b * a.H + sH) * a.W + sW) * 3;
    const float xc = __ldg(pc), yc = __ldg(pc + 1), zc = __ldg(pc + 2);
    const float dist_c = fmaxf(sqlen(xc, yc, zc), 1e-10f);  // :72
    const bool centre_valid = !(dist_c <= 1e-10f);           // :74-78 empty centre -> nothing written

    float qx[R], qy[R], qz[R];
    int hw[R];
    bool inside[R];
    if (centre_valid) {  // warp-uniform: an empty centre costs one load and the output stores
        const int half_H = a.kH / 2, half_W = a.kW / 2;
        const int base_h = (int)fastdiv((unsigned)sH, a.sh_magic) - half_H;              // :89-92  sH / stride_h
        const int base_w = (int)fastdiv((unsigned)sW, a.sw_magic) - half_W;
        const float *x2 = a.xyz2 + (size_t)b * a.small_h * a.small_w * 3;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int t = r * 32 + lane;
            inside[r] = false;
            hw[r] = 0;
            qx[r] = qy[r] = qz[r] = 0.f;
            if (t < total) {
                const unsigned khw = a.random_hw != nullptr ? (unsigned)__ldg(a.random_hw + t) : (unsigned)t;
                const unsigned q = fastdiv(khw, a.kw_magic);  // khw / kW, exact for khw, kW < 2^16
                int kh = base_h + (int)q;
                int kw = base_w + (int)(khw - q * (unsigned)a.kW);
                if (a.flag & I2P_FLAG_SHIFT) {  // :96-113 the range image is circular in width
                    if (kw < 0) kw += a.small_w;
                    if (kw >= a.small_w) kw -= a.small_w;
                }
                // (a window wider than the image can still fall outside after one wrap; the reference
                // would read out of bounds there, this kernel drops the slot)
                if (kh >= 0 && kh < a.small_h && kw >= 0 && kw < a.small_w) {
                    const float *q3 = x2 + (unsigned)(kh * a.small_w + kw) * 3u;
                    qx[r] = __ldg(q3); qy[r] = __ldg(q3 + 1); qz[r] = __ldg(q3 + 2);
                    hw[r] = (kh << 16) | kw;
                    inside[r] = true;
                }
            }
        }
    }

    if (centre_valid) {
        const float dist_square = __fmul_rn(a.distance, a.distance);  // :26
        float dist[R];
        int nvalid = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float dq0 = sqlen(qx[r], qy[r], qz[r]);  // :140 empty cell
            const float dq = fmaxf(sqlen(__fsub_rn(xc, qx[r]), __fsub_rn(yc, qy[r]), __fsub_rn(zc, qz[r])), 1e-10f);
            const bool ok = inside[r] && !(dq0 <= 1e-10f) && !(dq > dist_square);  // :146, :156
            dist[r] = ok ? dq : 1e10f;
            hw[r] = ok ? hw[r] : 0;
            // compact the valid slots, in slot order, into this warp's shared list
            const unsigned vmask = __ballot_sync(FULL, ok);
            if (ok) {
                const int pos = nvalid + __popc(vmask & ((1u << lane) - 1u));
                s_d[warp][pos] = __float_as_uint(dq);
                s_c[warp][pos] = hw[r];
            }
            nvalid += __popc(vmask);
        }
        __syncwarp();

        // Fast path: with pairwise-distinct distances the reference's selection sort emits the
        // valid slots in ascending distance, so each candidate's output position is its RANK --
        // nvalid independent compares per candidate, no serial dependency chain.
        bool tie = false;
        for (int c = lane; c < nvalid; c += 32) {
            const unsigned mine = s_d[warp][c];
            int rank = 0;
            for (int j = 0; j < nvalid; ++j) {
                const unsigned other = s_d[warp][j];
                rank += other < mine;
                tie |= (other == mine) && (j != c);
            }
            if (rank < K) s_hw[warp][rank] = s_c[warp][c];
            if (rank == 0) s_c[warp][SEL_MAX_SLOTS] = s_c[warp][c];  // the nearest, for FLAG_COPY
        }
        tie = __any_sync(FULL, tie);
        __syncwarp();
        if (!tie) {
            if (a.flag & I2P_FLAG_COPY) {  // :211-222 pad with the nearest; hw = 0 when nothing is valid
                const int first = nvalid > 0 ? s_c[warp][SEL_MAX_SLOTS] : 0;
                for (int k = lane; k < K; k += 32)
                    if (k >= nvalid) s_hw[warp][k] = first;
            }
        } else {
            // Exact distances tie: the reference's emitted order then depends on its swap
            // history.  Replay the K steps of its selection sort (:183-208), one warp-wide step each.
            for (int k = lane; k < K; k += 32) s_hw[warp][k] = -1;
            __syncwarp();
            const int steps = K < total ? K : total;
            for (int s = 0; s < steps; ++s) {
                unsigned best_d = 0xffffffffu, best_t = 0xffffffffu;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const unsigned t = r * 32 + lane;
                    const unsigned bits = __float_as_uint(dist[r]);  // all distances are positive floats
                    if (t >= (unsigned)s && t < (unsigned)total && bits < best_d) {
                        best_d = bits;
                        best_t = t;
                    }
                }
                const unsigned m = __reduce_min_sync(FULL, best_d);
                const unsigned mi = __reduce_min_sync(FULL, best_d == m ? best_t : 0xffffffffu);
                const int hw_m = __shfl_sync(FULL, pick<R>(hw, mi >> 5), mi & 31);
                const int hw_s = __shfl_sync(FULL, pick<R>(hw, s >> 5), s & 31);
                const float d_s = __shfl_sync(FULL, pick<R>(dist, s >> 5), s & 31);
                if (mi != (unsigned)s && lane == (int)(mi & 31)) {  // the swap: slot mi takes slot s's entry
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (r == (int)(mi >> 5)) {
                            dist[r] = d_s;
                            hw[r] = hw_s;
                        }
                }
                const bool valid = __uint_as_float(m) < 1e10f;
                if (s == 0 && (a.flag & I2P_FLAG_COPY)) {  // :211-222, unconditional on validity
                    for (int k = lane; k < K; k += 32) s_hw[warp][k] = hw_m;
                    __syncwarp();
                }
                if (valid && lane == 0) s_hw[warp][s] = hw_m;  // :225-233
                if (!valid) break;  // every remaining slot is 1e10: nothing more is written
            }
        }
    }
    __syncwarp();

    const size_t o = ((size_t)b * a.npoints + cn) * K;
    for (int k = lane; k < K; k += 32) {
        const int v = s_hw[warp][k];
        if (FLAT) {
            a.flat_idx[o + k] = v < 0 ? 0 : (v >> 16) * a.small_w + (v & 0xffff);
            a.flat_mask[o + k] = v < 0 ? 0.f : 1.f;
        } else if (v >= 0) {
            a.sel_b[o + k] = b;
            a.sel_h[o + k] = v >> 16;
            a.sel_w[o + k] = v & 0xffff;
            a.sel_mask[o + k] = 1.0f;
        }
    }
}

template <bool FLAT>
static int launch_select(const SelectArgs &a, cudaStream_t stream) {
    const int total = a.kH * a.kW;
    I2P_REQUIRE(a.batch >= 0 && a.npoints >= 0 && a.K >= 1, "select: bad batch/npoints/K");
    I2P_REQUIRE(total >= 1 && total <= SEL_MAX_K && a.K <= SEL_MAX_K,
                "select: kH*kW=%d or K=%d exceeds 150, the reference's per-thread array size "
                "(fused_conv_go.cu:52-53)", total, a.K);
    I2P_REQUIRE(a.small_w < 65536 && a.small_h < 32768 && (long long)a.small_h * a.small_w < (1LL << 30) &&
                    a.stride_h >= 1 && a.stride_w >= 1, "select: image too large or stride < 1");
    const long long centres = (long long)a.batch * a.npoints;
    if (centres == 0) return I2P_OK;
    I2P_REQUIRE(centres < (1LL << 31), "select: batch * npoints exceeds 2^31");
    I2P_REQUIRE(a.H < 65536 && a.W < 65536 && a.batch <= 65535 && (a.idx_n2 != nullptr || a.npoints < 65536),
                "select: image or centre grid too large");
    SelectArgs b = a;
    auto magic = [](int d) { return d <= 1 ? 0u : (unsigned)((0x100000000ULL + (unsigned)d - 1) / (unsigned)d); };
    b.kw_magic = magic(a.kW);
    b.ow_magic = magic(a.out_w);
    b.sh_magic = magic(a.stride_h);
    b.sw_magic = magic(a.stride_w);
    const dim3 grid(ceil_div(a.npoints, SEL_WARPS), a.batch);
    const int block = SEL_WARPS * 32;
    switch ((total + 31) / 32) {
        case 1: select_k_kernel<1, FLAT><<<grid, block, 0, stream>>>(b); break;
        case 2: select_k_kernel<2, FLAT><<<grid, block, 0, stream>>>(b); break;
        case 3: select_k_kernel<3, FLAT><<<grid, block, 0, stream>>>(b); break;
        case 4: select_k_kernel<4, FLAT><<<grid, block, 0, stream>>>(b); break;
        default: select_k_kernel<5, FLAT><<<grid, block, 0, stream>>>(b); break;
    }
    return check_launch("fused_conv_select_k");
}

}  // namespace i2p

extern "C" {

int i2p_fused_conv_select_k(int batch, int H, int W, int npoints, int kH, int kW, int K, int flag,
                            float distance, int stride_h, int stride_w, const float *xyz1, const float *xyz2,
                            const int32_t *idx_n2, const int32_t *random_hw, int64_t *selected_b_idx,
                            int64_t *selected_h_idx, int64_t *selected_w_idx, float *selected_mask,
                            int small_h, int small_w, void *stream) {
    i2p::SelectArgs a{};
    a.batch = batch; a.H = H; a.W = W; a.npoints = npoints; a.kH = kH; a.kW = kW; a.K = K; a.flag = flag;
    a.distance = distance; a.stride_h = stride_h; a.stride_w = stride_w; a.small_h = small_h; a.small_w = small_w;
    a.xyz1 = xyz1; a.xyz2 = xyz2; a.idx_n2 = idx_n2; a.random_hw = random_hw;
    a.sel_b = selected_b_idx; a.sel_h = selected_h_idx; a.sel_w = selected_w_idx; a.sel_mask = selected_mask;
    I2P_REQUIRE(idx_n2 != nullptr && random_hw != nullptr, "fused_conv_select_k: null index input");
    return i2p::launch_select<false>(a, i2p::as_stream(stream));
}

int i2p_select_k_flat(int batch, int H, int W, int npoints, int kH, int kW, int K, int flag, float distance,
                      int stride_h, int stride_w, const float *xyz1, const float *xyz2, const int32_t *idx_n2,
                      int out_w, int stride_ch, int stride_cw, int32_t *flat_idx, float *mask, int small_h,
                      int small_w, void *stream) {
    i2p::SelectArgs a{};
    a.batch = batch; a.H = H; a.W = W; a.npoints = npoints; a.kH = kH; a.kW = kW; a.K = K; a.flag = flag;
    a.distance = distance; a.stride_h = stride_h; a.stride_w = stride_w; a.small_h = small_h; a.small_w = small_w;
    a.xyz1 = xyz1; a.xyz2 = xyz2; a.idx_n2 = idx_n2; a.random_hw = nullptr;
    a.out_w = out_w; a.stride_ch = stride_ch; a.stride_cw = stride_cw;
    a.flat_idx = flat_idx; a.flat_mask = mask;
    I2P_REQUIRE(idx_n2 != nullptr || out_w >= 1, "select_k_flat: need idx_n2 or a centre grid");
    return i2p::launch_select<true>(a, i2p::as_stream(stream));
}
}
