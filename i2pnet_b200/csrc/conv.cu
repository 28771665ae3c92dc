// 3x3 convolutions of the RGB feature pyramid (src/modules/basicConv.py:6-20: fifteen conv3x3 + BN + LeakyReLU +
// MaxPool blocks; the reference runs them through ATen / cuDNN) as own kernels: forward and data gradient as an
// im2col-free implicit GEMM on the 5th-generation tensor cores (tcgen05.mma kind::tf32 with the 3-term hi/lo split of
// umma.cuh, f32-accurate, accumulator in tensor memory), weight gradient as a register-blocked f32 FMA kernel.
//
// Implicit GEMM without im2col.  One CTA computes 128 consecutive positions of ONE image's "virtually padded" plane
// for BN output channels: the plane is indexed flat with row pitch Wp = W + 4, the four extra positions of every row
// (and everything above / below the image) read as zero, so that the zero padding of the convolution is part of the
// data and a 3x3 tap (dy, dx) is a pure SHIFT of the flat index by dy * Wp + dx.  For a chunk of 16 input channels the
// CTA stages three bands of 130 positions (rows y - 1, y, y + 1 of the tile, one halo position left and right) in
// shared memory as four channel-group planes per band set: plane g holds, per position, the 16 bytes of channels
// 4g .. 4g + 3.  That IS the canonical K-major no-swizzle UMMA layout (8 x 16-byte core matrices, rows 16 bytes apart,
// LBO = plane pitch), and because a row is exactly 16 bytes -- the granularity of a descriptor's start address -- the
// A operand of tap (dy, dx) is the same planes read from start address + ((dy + 1) * 130 + dx + 1) * 16: nine shifted
// descriptors over one staged tile instead of nine staged tiles.  K per chunk = 9 taps x 16 channels = 18 k-steps of
// 8, each three MMAs (hi*hi + lo*hi + hi*lo).  The weights are split and laid out once per step by a pack kernel and
// arrive by one bulk copy (TMA engine) per chunk.  Epilogue: tcgen05.ld (lane = position), + bias, coalesced NCHW
// stores, and per-tile (count, mean, M2) of every output channel for the batch statistics of the BatchNorm that
// follows (the format csrc/rgb.cu's finalize merges) -- the separate statistics pass over y disappears.
//
// The data gradient is the same kernel: dX = conv(dY, W') with W'[ci][co][tap] = W[co][ci][8 - tap].
//
// Layout is NCHW f32 like the reference's tensors (lanes = consecutive x of one channel plane: coalesced loads while
// staging and coalesced stores from the accumulator's lanes); no layout conversion anywhere on the image branch.
#include <math.h>

#include "common.cuh"
#include "umma.cuh"

namespace i2p {
namespace conv {

constexpr int BM = 128, THREADS = 256;
constexpr int BAND = BM + 2, SLOTS = 3 * BAND;            // positions per band / per plane
constexpr uint32_t PLANE_BYTES = SLOTS * 16;              // 6240
constexpr int PAD = 4;                                    // Wp = W + PAD

__host__ __device__ constexpr int ck_of(int ki) { return ki < 16 ? 8 : 16; }                 // channels per chunk
__host__ __device__ constexpr int bn_of(int no) { return no <= 16 ? 16 : 32; }               // output channels per CTA
__host__ __device__ inline int chunks_of(int ki) { return (ki + ck_of(ki) - 1) / ck_of(ki); }
__host__ __device__ inline int ntiles_of(int no) { return (no + bn_of(no) - 1) / bn_of(no); }
// floats of one (n-tile, chunk) weight block: [hi | lo][tap 9][k-group CK/4][n BN][4]
__host__ __device__ inline long long block_floats(int ki, int no) { return 2LL * 9 * (ck_of(ki) / 4) * bn_of(no) * 4; }

// W (no_src..): forward  pack[n = co][k = ci][tap] = w[co][ci][tap]            (w is (cout, cin, 3, 3))
//               dgrad    pack[n = ci][k = co][tap] = w[co][ci][8 - tap]
__global__ void __launch_bounds__(256) pack_kernel(int cin, int cout, int dgrad, const float *__restrict__ w,
                                                   float *__restrict__ pack) {
    const int ki = dgrad ? cout : cin, no = dgrad ? cin : cout;
    const int ck = ck_of(ki), bn = bn_of(no), kgs = ck / 4;
    const int nch = chunks_of(ki), nt = ntiles_of(no);
    const long long half = 9LL * kgs * bn * 4, total = (long long)nt * nch * half;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long blk = e / half;
        int r = (int)(e - blk * half);
        const int t = (int)(blk / nch), c = (int)(blk - (long long)t * nch);
        const int el = r & 3; r >>= 2;
        const int nl = r % bn; r /= bn;
        const int kg = r % kgs, tap = r / kgs;
        const int n = t * bn + nl, k = c * ck + kg * 4 + el;
        float v = 0.f;
        if (n < no && k < ki) v = dgrad ? __ldg(w + ((size_t)k * cin + n) * 9 + (8 - tap)) : __ldg(w + ((size_t)n * cin + k) * 9 + tap);
        uint32_t hi, lo;
        umma::split_tf32(v, hi, lo);
        float *base = pack + blk * 2 * half;
        const long long off = e - blk * half;
        base[off] = __uint_as_float(hi);
        base[half + off] = __uint_as_float(lo);
    }
}

struct Args {
    int B, ki, no, H, W, Wp, tiles, nchunks;
    const float *x;        // (B, ki, H, W)
    const float *wpack;
    const float *bias;     // (no) or null
    float *y;              // (B, no, H, W)
    float *tile_stats;     // (no, B * tiles, 3) = (count, mean, M2) or null
};

__host__ __device__ constexpr uint32_t idesc(int n) {     // kind::tf32, f32 accumulate, K-major A and B, M = 128
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int CK, int BN>
__host__ __device__ constexpr int smem_bytes() { return 2 * (CK / 4) * (int)PLANE_BYTES + 2 * 9 * (CK / 4) * BN * 16; }

template <int NC>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[NC]) {
    static_assert(NC == 8 || NC == 16, "8 or 16 columns");
    uint32_t r[NC];
    if constexpr (NC == 16) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
    } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = __uint_as_float(r[i]);
}

template <int CK, int BN>
__global__ void __launch_bounds__(THREADS, BN == 16 ? 3 : 2) conv3x3_tc_kernel(const Args a) {
    constexpr int KG = CK / 4;                                   // channel-group planes per chunk
    constexpr uint32_t A_HALF = KG * PLANE_BYTES;                // hi (or lo) planes of the activation tile
    constexpr uint32_t B_HALF = 9 * KG * BN * 16;                // hi (or lo) weights of one chunk
    constexpr int NC = BN / 2;                                   // accumulator columns per thread in the epilogue
    extern __shared__ __align__(128) unsigned char smem[];       // A hi | A lo | B hi | B lo
    __shared__ uint64_t mma_bar, b_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float red_sum[4][BN], red_m2[4][BN];
    __shared__ int red_cnt[4];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, nt = blockIdx.y, b = blockIdx.z;
    const int i0 = tile * BM, n0 = nt * BN;
    const int H = a.H, W = a.W, Wp = a.Wp;

    if (warp == 0) umma::tmem_alloc(&tmem_slot, 4 * BN);
    if (tid == 0) {
        umma::mbar_init(&mma_bar, 1);
        umma::mbar_init(&b_bar, 1);
        umma::mbar_fence_init();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_d = tmem_slot;
    const uint32_t sbase = umma::smem_u32(smem);
    const uint32_t sB = sbase + 2 * A_HALF;
    const float *wblk = a.wpack + (size_t)nt * a.nchunks * (2 * B_HALF / 4);
    const size_t plane = (size_t)H * W;
    const float *xb = a.x + (size_t)b * a.ki * plane;

    for (int c = 0; c < a.nchunks; ++c) {
        if (c >= 1) umma::mbar_wait(&mma_bar, (uint32_t)(c - 1) & 1u);     // the tensor core is done with the stage
        if (tid == 0) {     // this chunk's weights: one bulk copy, landing through the async proxy
            umma::mbar_expect_tx(&b_bar, 2 * B_HALF);
            umma::bulk_g2s(sB, wblk + (size_t)c * (2 * B_HALF / 4), 2 * B_HALF, &b_bar);
        }
        // ---- stage the activation bands of channels c * CK .. + CK - 1: item = (plane g, band, slot)
        const int ch0 = c * CK;
        for (int it = tid; it < KG * SLOTS; it += THREADS) {
            const int g = it / SLOTS, s = it - g * SLOTS;          // slot inside the plane: band * BAND + position
            const int band = s / BAND, pos = s - band * BAND;
            // flat padded index of this slot, shifted by two rows so that it is never negative
            const int qq = i0 + (band + 1) * Wp + pos - 1;
            const int yy = qq / Wp - 2, xx = qq - (yy + 2) * Wp;
            float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
            if (yy >= 0 && yy < H && xx < W) {
                const int ch = ch0 + g * 4;
                const float *p = xb + (size_t)ch * plane + (size_t)yy * W + xx;
                if (ch < a.ki) v0 = __ldg(p);
                if (ch + 1 < a.ki) v1 = __ldg(p + plane);
                if (ch + 2 < a.ki) v2 = __ldg(p + 2 * plane);
                if (ch + 3 < a.ki) v3 = __ldg(p + 3 * plane);
            }
            uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
            umma::split_tf32(v0, h0, l0);
            umma::split_tf32(v1, h1, l1);
            umma::split_tf32(v2, h2, l2);
            umma::split_tf32(v3, h3, l3);
            unsigned char *dst = smem + (size_t)g * PLANE_BYTES + (size_t)s * 16;
            *reinterpret_cast<uint4 *>(dst) = make_uint4(h0, h1, h2, h3);
            *reinterpret_cast<uint4 *>(dst + A_HALF) = make_uint4(l0, l1, l2, l3);
        }
        umma::fence_smem_to_async();
        __syncthreads();
        if (tid == 0) {
            umma::fence_after_sync();
            umma::mbar_wait(&b_bar, (uint32_t)c & 1u);
            constexpr uint32_t id = idesc(BN);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const uint32_t shift = (uint32_t)((tap / 3) * BAND + (tap % 3)) * 16u;     // band dy + 1, position dx + 1
#pragma unroll
                for (int j = 0; j < CK / 8; ++j) {
                    const uint32_t ao = sbase + shift + (uint32_t)(2 * j) * PLANE_BYTES;
                    const uint32_t bo = sB + (uint32_t)(tap * KG + 2 * j) * (BN * 16);
                    const uint64_t ah = umma::smem_desc(ao, PLANE_BYTES, 128), al = umma::smem_desc(ao + A_HALF, PLANE_BYTES, 128);
                    const uint64_t bh = umma::smem_desc(bo, BN * 16, 128), bl = umma::smem_desc(bo + B_HALF, BN * 16, 128);
                    // Four accumulators: the tensor core adds into the accumulator with truncation, a bias that grows with
                    // the number of accumulation steps times |accumulator| (measured 4e-6 of the output at K = 576 with
                    // one accumulator and same-signed products, i.e. post-activation inputs).  The large hi*hi terms go to
                    // one accumulator per kernel row (K / 24 steps each), the 2^-11-sized correction terms to a fourth; the
                    // epilogue adds the four in f32 with round-to-nearest.
                    umma::mma_tf32(tmem_d + (uint32_t)((tap / 3) * BN), ah, bh, id, !(c == 0 && tap % 3 == 0 && j == 0));
                    umma::mma_tf32(tmem_d + 3 * BN, al, bh, id, !(c == 0 && tap == 0 && j == 0));
                    umma::mma_tf32(tmem_d + 3 * BN, ah, bl, id, true);
                }
            }
            umma::commit(&mma_bar);
        }
    }
    umma::mbar_wait(&mma_bar, (uint32_t)(a.nchunks - 1) & 1u);
    umma::fence_after_sync();

    // ---- epilogue: lane = position i0 + row; warps 0-3 take the first half of the channels, 4-7 the second
    const int row = (warp & 3) * 32 + lane;
    const int cbase = (warp >> 2) * NC;
    const int q = i0 + row;
    const int yy = q / Wp, xx = q - yy * Wp;
    const bool valid = yy < H && xx < W;
    float v[NC];
    {
        float u0[NC], u1[NC], u2[NC];
        const uint32_t t0 = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)cbase;
        tmem_ld<NC>(t0, v);
        tmem_ld<NC>(t0 + BN, u0);
        tmem_ld<NC>(t0 + 2 * BN, u1);
        tmem_ld<NC>(t0 + 3 * BN, u2);
#pragma unroll
        for (int j = 0; j < NC; ++j) v[j] = (v[j] + u0[j]) + (u1[j] + u2[j]);
    }
    float *yb = a.y + ((size_t)b * a.no) * plane + (size_t)yy * W + xx;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        const int n = n0 + cbase + j;
        if (a.bias != nullptr && n < a.no) v[j] += __ldg(a.bias + n);
        if (valid && n < a.no) yb[(size_t)n * plane] = v[j];
    }
    umma::fence_before_sync();
    if (a.tile_stats != nullptr) {
        // (count, mean, M2) of the tile's valid positions per channel: warp sums -> CTA mean -> warp sums of squares
        const unsigned vmask = __ballot_sync(FULL, valid);
        if (lane == 0 && warp < 4) red_cnt[warp] = __popc(vmask);
        float s[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            s[j] = valid ? v[j] : 0.f;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s[j] += __shfl_xor_sync(FULL, s[j], off);
        }
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < NC; ++j) red_sum[warp & 3][cbase + j] = s[j];
        }
        __syncthreads();
        const float cnt = (float)(red_cnt[0] + red_cnt[1] + red_cnt[2] + red_cnt[3]);
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const float mean = cnt > 0.f ? (red_sum[0][cbase + j] + red_sum[1][cbase + j] + red_sum[2][cbase + j] + red_sum[3][cbase + j]) / cnt : 0.f;
            const float d = valid ? v[j] - mean : 0.f;
            float m2 = d * d;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) m2 += __shfl_xor_sync(FULL, m2, off);
            s[j] = m2;
        }
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < NC; ++j) red_m2[warp & 3][cbase + j] = s[j];
        }
        __syncthreads();
        if (tid < BN && n0 + tid < a.no) {
            const float mean = cnt > 0.f ? (red_sum[0][tid] + red_sum[1][tid] + red_sum[2][tid] + red_sum[3][tid]) / cnt : 0.f;
            float *ts = a.tile_stats + ((size_t)(n0 + tid) * ((size_t)a.B * a.tiles) + (size_t)b * a.tiles + tile) * 3;
            ts[0] = cnt;
            ts[1] = mean;
            ts[2] = red_m2[0][tid] + red_m2[1][tid] + red_m2[2][tid] + red_m2[3][tid];
        }
    } else {
        __syncthreads();
    }
    if (warp == 0) umma::tmem_dealloc(tmem_d, 4 * BN);
}

// ---------------------------------------------------------------------------------------------
// weight gradient: dW[co][ci][ky][kx] += sum_{b, y, x} dY[b][co][y][x] * X[b][ci][y + ky - 1][x + kx - 1]
// One CTA = one image tile of TR rows x TW columns (about 512 positions) and one (16 output channels) x (CI input
// channels, 16 or 4) sub-problem.  x (with a one-position halo, zero outside the image) and dY are staged in shared
// memory position-major ([position][channel], 20-float records: conflict-free 128-bit stores from the plane-major
// loads and conflict-free 128-bit loads below).  A thread owns a 4 x 4 block of (co, ci) pairs with all nine taps in
// registers (144 accumulators) and every 16th position of its rows: per position one LDS.128 of dY and nine of x feed
// 144 FMAs.  The 16 position-partitions of a block are the lanes of a half-warp: shuffle reduction, then one
// red.global.add per element and CTA.
// ---------------------------------------------------------------------------------------------
struct WArgs {
    int B, cin, cout, H, W, TR, TW, tiles_x, tiles_y;
    const float *x, *dy;
    float *dw;
};

template <int CI>      // input channels per CTA: 16, or 4 for the three-channel first layer
__global__ void __launch_bounds__(THREADS, 1) wgrad_kernel(const WArgs a) {
    constexpr int XS = CI == 16 ? 20 : 4;        // floats per staged x position
    constexpr int DS = 20;                       // floats per staged dY position (16 channels + 4 padding)
    constexpr int CIB = CI / 4, NBLK = 4 * CIB;  // 4 x 4 register blocks: 4 along co, CIB along ci
    constexpr int GROUPS = THREADS / (16 * NBLK);
    extern __shared__ __align__(16) float sm[];
    const int TR = a.TR, TW = a.TW, XW = TW + 2;
    float *xs = sm;                                   // (TR + 2) x XW x XS
    float *ds = sm + (size_t)(TR + 2) * XW * XS;      // TR x TW x DS
    const int tid = threadIdx.x;
    const int tx = blockIdx.x % a.tiles_x, ty = blockIdx.x / a.tiles_x;
    const int cig = blockIdx.y % ((a.cin + CI - 1) / CI), cog = blockIdx.y / ((a.cin + CI - 1) / CI);
    const int b = blockIdx.z;
    const int y0 = ty * TR, x0 = tx * TW, ci0 = cig * CI, co0 = cog * 16;
    const size_t plane = (size_t)a.H * a.W;
    const float *xb = a.x + (size_t)b * a.cin * plane, *dyb = a.dy + (size_t)b * a.cout * plane;

    // ---- stage x: item = (4-channel group, position); lanes run along x (coalesced plane reads)
    for (int it = tid; it < CIB * (TR + 2) * XW; it += THREADS) {
        const int g = it / ((TR + 2) * XW), p = it - g * ((TR + 2) * XW);
        const int r = p / XW, cx = p - r * XW;
        const int yy = y0 + r - 1, xx = x0 + cx - 1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yy >= 0 && yy < a.H && xx >= 0 && xx < a.W) {
            const int ch = ci0 + g * 4;
            const float *src = xb + (size_t)ch * plane + (size_t)yy * a.W + xx;
            if (ch < a.cin) v.x = __ldg(src);
            if (ch + 1 < a.cin) v.y = __ldg(src + plane);
            if (ch + 2 < a.cin) v.z = __ldg(src + 2 * plane);
            if (ch + 3 < a.cin) v.w = __ldg(src + 3 * plane);
        }
        *reinterpret_cast<float4 *>(xs + (size_t)p * XS + g * 4) = v;
    }
    for (int it = tid; it < 4 * TR * TW; it += THREADS) {
        const int g = it / (TR * TW), p = it - g * (TR * TW);
        const int r = p / TW, cx = p - r * TW;
        const int yy = y0 + r, xx = x0 + cx;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yy < a.H && xx < a.W) {
            const int ch = co0 + g * 4;
            const float *src = dyb + (size_t)ch * plane + (size_t)yy * a.W + xx;
            if (ch < a.cout) v.x = __ldg(src);
            if (ch + 1 < a.cout) v.y = __ldg(src + plane);
            if (ch + 2 < a.cout) v.z = __ldg(src + 2 * plane);
            if (ch + 3 < a.cout) v.w = __ldg(src + 3 * plane);
        }
        *reinterpret_cast<float4 *>(ds + (size_t)p * DS + g * 4) = v;
    }
    __syncthreads();

    const int part = tid & 15, blk = (tid >> 4) % NBLK, grp = tid / (16 * NBLK);
    const int cob = blk / CIB, cib = blk - cob * CIB;
    float acc[4][4][9];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int t = 0; t < 9; ++t) acc[i][j][t] = 0.f;

    for (int r = grp; r < TR; r += GROUPS) {
        for (int cx = part; cx < TW; cx += 16) {
            const float4 d = *reinterpret_cast<const float4 *>(ds + (size_t)(r * TW + cx) * DS + cob * 4);
            const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float4 xv = *reinterpret_cast<const float4 *>(xs + (size_t)((r + t / 3) * XW + cx + t % 3) * XS + cib * 4);
                const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j][t] = __fmaf_rn(dv[i], xa[j], acc[i][j][t]);
            }
        }
    }
    // ---- reduce the 16 position-partitions (lanes of a half-warp), then one atomic per element
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                float s = acc[i][j][t];
#pragma unroll
                for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(FULL, s, off);
                acc[i][j][t] = s;
            }
    if (part == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int co = co0 + cob * 4 + i;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ci = ci0 + cib * 4 + j;
                if (co < a.cout && ci < a.cin) {
#pragma unroll
                    for (int t = 0; t < 9; ++t) atomicAdd(a.dw + ((size_t)co * a.cin + ci) * 9 + t, acc[i][j][t]);
                }
            }
        }
    }
}

template <typename K>
static void allow_smem(K kernel, int bytes) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

}  // namespace conv
}  // namespace i2p

extern "C" {

/* floats of the packed (tf32 hi | lo, UMMA layout) weights of a 3x3 convolution; dgrad: the transposed, flipped form */
long long i2p_conv3x3_pack_floats(int cin, int cout, int dgrad) {
    using namespace i2p::conv;
    const int ki = dgrad ? cout : cin, no = dgrad ? cin : cout;
    return (long long)ntiles_of(no) * chunks_of(ki) * block_floats(ki, no);
}

int i2p_conv3x3_tiles(int H, int W) { return (H * (W + i2p::conv::PAD) + i2p::conv::BM - 1) / i2p::conv::BM; }

int i2p_conv3x3_pack(int cin, int cout, int dgrad, const float *w, float *pack, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(cin >= 1 && cout >= 1 && w != nullptr && pack != nullptr, "conv3x3_pack: bad arguments");
    const long long total = i2p_conv3x3_pack_floats(cin, cout, dgrad) / 2;
    const long long blocks = (total + 255) / 256;
    conv::pack_kernel<<<(int)(blocks < 592 ? blocks : 592), 256, 0, as_stream(stream)>>>(cin, cout, dgrad, w, pack);
    return check_launch("conv3x3_pack");
}

/* y (B, no, H, W) = conv3x3(x (B, ki, H, W), pad 1) [+ bias]; wpack from i2p_conv3x3_pack (forward: cin = ki, cout = no;
 * data gradient: dgrad = 1 with cin = no, cout = ki).  tile_stats (no, B * i2p_conv3x3_tiles(H, W), 3) or null. */
int i2p_conv3x3_tc(int B, int ki, int no, int H, int W, const float *x, const float *wpack, const float *bias, float *y,
                   float *tile_stats, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(B >= 1 && B <= 65535 && ki >= 1 && no >= 1 && H >= 1 && W >= 1, "conv3x3_tc: bad sizes");
    I2P_REQUIRE((long long)(H + 3) * (W + conv::PAD) < (1LL << 30), "conv3x3_tc: image too large");
    I2P_REQUIRE((reinterpret_cast<uintptr_t>(wpack) & 15) == 0, "conv3x3_tc: wpack must be 16-byte aligned");
    conv::Args a;
    a.B = B; a.ki = ki; a.no = no; a.H = H; a.W = W; a.Wp = W + conv::PAD;
    a.tiles = i2p_conv3x3_tiles(H, W);
    a.nchunks = conv::chunks_of(ki);
    a.x = x; a.wpack = wpack; a.bias = bias; a.y = y; a.tile_stats = tile_stats;
    dim3 grid(a.tiles, conv::ntiles_of(no), B);
    cudaStream_t s = as_stream(stream);
#define I2P_CONV(CK_, BN_)                                                                                          \
    do {                                                                                                            \
        static bool once = false;                                                                                   \
        if (!once) { conv::allow_smem(conv::conv3x3_tc_kernel<CK_, BN_>, conv::smem_bytes<CK_, BN_>()); once = true; } \
        conv::conv3x3_tc_kernel<CK_, BN_><<<grid, conv::THREADS, conv::smem_bytes<CK_, BN_>(), s>>>(a);              \
    } while (0)
    const int ck = conv::ck_of(ki), bn = conv::bn_of(no);
    if (ck == 8) { if (bn == 16) I2P_CONV(8, 16); else I2P_CONV(8, 32); }
    else { if (bn == 16) I2P_CONV(16, 16); else I2P_CONV(16, 32); }
#undef I2P_CONV
    return check_launch("conv3x3_tc");
}

/* dw (cout, cin, 3, 3) += weight gradient of y = conv3x3(x, w, pad 1) from dy (B, cout, H, W) and x (B, cin, H, W) */
int i2p_conv3x3_wgrad(int B, int cin, int cout, int H, int W, const float *x, const float *dy, float *dw, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(B >= 1 && B <= 65535 && cin >= 1 && cout >= 1 && H >= 1 && W >= 1, "conv3x3_wgrad: bad sizes");
    conv::WArgs a;
    a.B = B; a.cin = cin; a.cout = cout; a.H = H; a.W = W;
    a.TW = W < 128 ? W : 128;
    a.TR = 512 / a.TW;
    if (a.TR > H) a.TR = H;
    if (a.TR < 1) a.TR = 1;
    a.tiles_x = ceil_div(W, a.TW);
    a.tiles_y = ceil_div(H, a.TR);
    a.x = x; a.dy = dy; a.dw = dw;
    const int ci_t = cin <= 4 ? 4 : 16;
    dim3 grid(a.tiles_x * a.tiles_y, ceil_div(cin, ci_t) * ceil_div(cout, 16), B);
    const int xs = ci_t == 16 ? 20 : 4;
    const int bytes = ((a.TR + 2) * (a.TW + 2) * xs + a.TR * a.TW * 20) * 4;
    cudaStream_t s = as_stream(stream);
    if (ci_t == 16) {
        static bool once = false;
        if (!once) { conv::allow_smem(conv::wgrad_kernel<16>, 128 * 1024); once = true; }
        conv::wgrad_kernel<16><<<grid, conv::THREADS, bytes, s>>>(a);
    } else {
        static bool once = false;
        if (!once) { conv::allow_smem(conv::wgrad_kernel<4>, 128 * 1024); once = true; }
        conv::wgrad_kernel<4><<<grid, conv::THREADS, bytes, s>>>(a);
    }
    return check_launch("conv3x3_wgrad");
}
}
