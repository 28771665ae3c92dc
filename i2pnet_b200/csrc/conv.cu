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
constexpr uint32_t PLANE_BYTES = SLOTS * 16 + 48;         // 6288 = 16 (mod 128): consecutive planes are 4 banks apart (see the staging)
constexpr int PAD = 4;                                    // Wp = W + PAD

__host__ __device__ constexpr int ck_of(int ki) { return ki < 16 ? 8 : 16; }                 // channels per chunk
__host__ __device__ constexpr int bn_of(int no) { return no <= 16 ? 16 : 32; }               // output channels per CTA
__host__ __device__ inline int chunks_of(int ki) { return (ki + ck_of(ki) - 1) / ck_of(ki); }
__host__ __device__ inline int ntiles_of(int no) { return (no + bn_of(no) - 1) / bn_of(no); }
// floats of one (n-tile, chunk) weight block: [tap 9][k-group CK/4][hi n BN | lo n BN][4]
__host__ __device__ inline long long block_floats(int ki, int no) { return 2LL * 9 * (ck_of(ki) / 4) * bn_of(no) * 4; }

// W (no_src..): forward  pack[n = co][k = ci][tap] = w[co][ci][tap]            (w is (cout, cin, 3, 3))
//               dgrad    pack[n = ci][k = co][tap] = w[co][ci][8 - tap]
__device__ __forceinline__ void pack_body(int cin, int cout, int dgrad, const float *__restrict__ w, float *__restrict__ pack, int block,
                                          int nblocks) {
    const int ki = dgrad ? cout : cin, no = dgrad ? cin : cout;
    const int ck = ck_of(ki), bn = bn_of(no), kgs = ck / 4;
    const int nch = chunks_of(ki), nt = ntiles_of(no);
    const long long half = 9LL * kgs * bn * 4, total = (long long)nt * nch * half;
    for (long long e = (long long)block * 256 + threadIdx.x; e < total; e += (long long)nblocks * 256) {
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
        // [tap][k-group][hi rows 0 .. bn-1 | lo rows 0 .. bn-1][4]: the hi and lo halves of one (tap, k-group) are ONE
        // K-major operand of 2 bn rows, so that A_hi x [B_hi | B_lo] is a single MMA of N = 2 bn
        float *base = pack + blk * 2 * half + ((size_t)(tap * kgs + kg) * 2 * bn) * 4;
        base[nl * 4 + el] = __uint_as_float(hi);
        base[(bn + nl) * 4 + el] = __uint_as_float(lo);
    }
}

__global__ void __launch_bounds__(256) pack_kernel(int cin, int cout, int dgrad, const float *__restrict__ w,
                                                   float *__restrict__ pack) {
    pack_body(cin, cout, dgrad, w, pack, blockIdx.x, gridDim.x);
}

// every layer and orientation of a pyramid in one launch: table[l] = { cin | cout << 32, dgrad, weight pointer, pack pointer }
__global__ void __launch_bounds__(256) pack_multi_kernel(const long long *__restrict__ table) {
    const long long *t = table + (size_t)blockIdx.y * 4;
    pack_body((int)(t[0] & 0xffffffffll), (int)(t[0] >> 32), (int)t[1], reinterpret_cast<const float *>(t[2]),
              reinterpret_cast<float *>(t[3]), blockIdx.x, gridDim.x);
}

struct Args {
    int B, ki, no, H, W, Wp, tiles, nchunks;
    float inv_wp;
    const float *x;        // (B, ki, H, W)
    const float *wpack;
    const float *bias;     // (no) or null
    float *y;              // (B, no, H, W)
    float *tile_stats;     // (no, gridDim.x, 3) = (count, mean, M2) per CTA, or null
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

// Persistent: a CTA owns one n-tile (BN output channels) and loops over (image, 128-position tile) work items.  Per
// item: [chunk loop: operand loads -> (tensor core done with the stage) -> split + store -> MMAs], then the global loads
// of the NEXT item's first chunk are issued into registers, and only then the epilogue of this item runs (wait for the
// MMAs, TMEM -> registers, stores, statistics): the next item's memory latency hides behind it.  TMEM, the barriers and
// -- for layers of at most 16 input channels -- the weights are set up once per CTA.
//
// The kernel is bound by instruction issue (ncu, profiles/r2_conv_full.md), so the staging is vectorised (VEC: W % 4 == 0
// and 16-byte aligned planes): an item is 4 consecutive positions x the 4 channels of one plane = four LDG.128, sixteen
// splits, eight STS.128.  A quad never straddles a row end (Wp, the tile origin and the quad offset are multiples of 4),
// so it is entirely inside or outside the image.  Lanes of a quarter-warp hold the KG planes x consecutive quads; planes
// are 16 (mod 128) bytes apart, so that the eight 16-byte stores of a phase fall into eight different bank groups.
// The batch statistics are accumulated in registers over ALL items of the CTA as sums shifted by a per-channel reference
// (the first item's mean, so that the cancellation in M2 = S2 - S1^2 / n stays mild) and reduced once per CTA.
template <int CK, int BN, bool VEC>
__global__ void __launch_bounds__(THREADS, 2) conv3x3_tc_kernel(const Args a) {
    constexpr int KG = CK / 4;                                   // channel-group planes per chunk
    constexpr uint32_t A_HALF = KG * PLANE_BYTES;                // hi (or lo) planes of the activation tile
    constexpr uint32_t B_HALF = 9 * KG * BN * 16;                // half the bytes of one chunk's weights (hi and lo interleaved per k-group)
    constexpr int NC = BN / 2;                                   // accumulator columns per thread in the epilogue
    constexpr int SITEMS = (KG * SLOTS + THREADS - 1) / THREADS; // scalar path: items = (plane, slot)
    constexpr int QPW = 32 / KG;                                 // vector path: quads per warp-item (lanes = KG planes x QPW quads)
    constexpr int WITEMS = 3 * 32 / QPW;                         // warp-items per chunk: 3 bands x 32 quads
    constexpr int VROUNDS = (WITEMS + 7) / 8;
    constexpr int NLD = VEC ? 4 * VROUNDS + 1 : SITEMS;
    constexpr uint32_t TCOLS = BN == 16 ? 128 : 256;             // six accumulators of BN columns, rounded up to a power of two
    extern __shared__ __align__(128) unsigned char smem[];       // A hi | A lo | B hi | B lo
    __shared__ uint64_t mma_bar, b_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float red_sum[4][BN], red_m2[4][BN], ref_s[THREADS / 32][NC];
    __shared__ int red_cnt[4];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nt = blockIdx.y, n0 = nt * BN;
    const int H = a.H, W = a.W, Wp = a.Wp;
    const int nwork = a.B * a.tiles;

    if (warp == 0) umma::tmem_alloc(&tmem_slot, TCOLS);
    if (tid == 0) {
        umma::mbar_init(&mma_bar, 3);      // three issuing warps commit per chunk
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
    uint32_t mma_commits = 0, b_loads = 0;     // phase counters of the two barriers (uniform across the CTA)

    // row / column of a flat padded index shifted by two rows (never negative); exact: the quotient's fraction is at
    // least 0.5 / Wp away from an integer and the float error is below that for indices < 2^20
    auto rowcol = [&](int qq, int &yy, int &xx) {
        const int y2 = (int)(((float)qq + 0.5f) * a.inv_wp);
        yy = y2 - 2;
        xx = qq - y2 * Wp;
    };

    float4 ld[NLD];
    auto fetch = [&](int T, int c) {
        const int b = T / a.tiles, i0 = (T - b * a.tiles) * BM;
        const float *xb = a.x + (size_t)b * a.ki * plane;
        const int ch0 = c * CK;
        if constexpr (VEC) {
            const int g = lane % KG, kk = lane / KG;
            const int ch = ch0 + g * 4;
#pragma unroll
            for (int r = 0; r < VROUNDS; ++r) {
                const int wi = warp + 8 * r;                         // warp-item: (band, group of QPW quads)
                const int band = wi / (32 / QPW), k = (wi % (32 / QPW)) * QPW + kk;
                int yy, xx;
                rowcol(i0 + (band + 1) * Wp + 4 * k, yy, xx);
                const bool ok = wi < WITEMS && yy >= 0 && yy < H && xx < W;
                const float4 *p = reinterpret_cast<const float4 *>(xb + (size_t)ch * plane + (size_t)yy * W + xx);
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                ld[4 * r + 0] = ok && ch < a.ki ? __ldg(p) : z;
                ld[4 * r + 1] = ok && ch + 1 < a.ki ? __ldg(p + plane / 4) : z;
                ld[4 * r + 2] = ok && ch + 2 < a.ki ? __ldg(p + 2 * (plane / 4)) : z;
                ld[4 * r + 3] = ok && ch + 3 < a.ki ? __ldg(p + 3 * (plane / 4)) : z;
            }
            // the two halo positions of every band and plane (slot 0 and slot 129): threads 0 .. 6 KG - 1
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tid < 6 * KG) {
                const int g2 = tid % KG, e = tid / KG, band = e >> 1, pos = (e & 1) * (BAND - 1);
                int yy, xx;
                rowcol(i0 + (band + 1) * Wp + pos - 1, yy, xx);
                if (yy >= 0 && yy < H && xx < W) {
                    const int ch2 = ch0 + g2 * 4;
                    const float *p = xb + (size_t)ch2 * plane + (size_t)yy * W + xx;
                    if (ch2 < a.ki) v.x = __ldg(p);
                    if (ch2 + 1 < a.ki) v.y = __ldg(p + plane);
                    if (ch2 + 2 < a.ki) v.z = __ldg(p + 2 * plane);
                    if (ch2 + 3 < a.ki) v.w = __ldg(p + 3 * plane);
                }
            }
            ld[4 * VROUNDS] = v;
        } else {
#pragma unroll
            for (int i = 0; i < SITEMS; ++i) {
                const int it = tid + i * THREADS;
                const int g = it / SLOTS, s = it - g * SLOTS;          // slot inside the plane: band * BAND + position
                const int band = s / BAND, pos = s - band * BAND;
                int yy, xx;
                rowcol(i0 + (band + 1) * Wp + pos - 1, yy, xx);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (it < KG * SLOTS && yy >= 0 && yy < H && xx < W) {
                    const int ch = ch0 + g * 4;
                    const float *p = xb + (size_t)ch * plane + (size_t)yy * W + xx;
                    if (ch < a.ki) v.x = __ldg(p);
                    if (ch + 1 < a.ki) v.y = __ldg(p + plane);
                    if (ch + 2 < a.ki) v.z = __ldg(p + 2 * plane);
                    if (ch + 3 < a.ki) v.w = __ldg(p + 3 * plane);
                }
                ld[i] = v;
            }
        }
    };
    auto put = [&](unsigned char *dst, float v0, float v1, float v2, float v3) {   // split and store one position's 4 channels
        uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
        umma::split_tf32(v0, h0, l0);
        umma::split_tf32(v1, h1, l1);
        umma::split_tf32(v2, h2, l2);
        umma::split_tf32(v3, h3, l3);
        *reinterpret_cast<uint4 *>(dst) = make_uint4(h0, h1, h2, h3);
        *reinterpret_cast<uint4 *>(dst + A_HALF) = make_uint4(l0, l1, l2, l3);
    };
    auto store = [&]() {
        if constexpr (VEC) {
            const int g = lane % KG, kk = lane / KG;
#pragma unroll
            for (int r = 0; r < VROUNDS; ++r) {
                const int wi = warp + 8 * r;
                if (wi < WITEMS) {
                    const int band = wi / (32 / QPW), k = (wi % (32 / QPW)) * QPW + kk;
                    unsigned char *dst = smem + (size_t)g * PLANE_BYTES + (size_t)(band * BAND + 1 + 4 * k) * 16;
                    const float4 c0 = ld[4 * r], c1 = ld[4 * r + 1], c2 = ld[4 * r + 2], c3 = ld[4 * r + 3];
                    put(dst, c0.x, c1.x, c2.x, c3.x);
                    put(dst + 16, c0.y, c1.y, c2.y, c3.y);
                    put(dst + 32, c0.z, c1.z, c2.z, c3.z);
                    put(dst + 48, c0.w, c1.w, c2.w, c3.w);
                }
            }
            if (tid < 6 * KG) {
                const int g2 = tid % KG, e = tid / KG, band = e >> 1, pos = (e & 1) * (BAND - 1);
                const float4 v = ld[4 * VROUNDS];
                put(smem + (size_t)g2 * PLANE_BYTES + (size_t)(band * BAND + pos) * 16, v.x, v.y, v.z, v.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < SITEMS; ++i) {
                const int it = tid + i * THREADS;
                if (it < KG * SLOTS) {
                    const int g = it / SLOTS, s = it - g * SLOTS;
                    put(smem + (size_t)g * PLANE_BYTES + (size_t)s * 16, ld[i].x, ld[i].y, ld[i].z, ld[i].w);
                }
            }
        }
    };

    // epilogue constants and the statistics accumulators of this thread's NC channels
    const int row = (warp & 3) * 32 + lane;
    const int cbase = (warp >> 2) * NC;
    float s1[NC], s2[NC];     // sums of (y - ref), (y - ref)^2 over this thread's positions; ref (shared memory): per warp and channel
#pragma unroll
    for (int j = 0; j < NC; ++j) s1[j] = s2[j] = 0.f;
    float *ref = ref_s[warp];
    if (lane < NC) ref[lane] = 0.f;
    __syncwarp();
    const float *bias_p = a.bias != nullptr ? a.bias + n0 + cbase : nullptr;    // (read per item: L1-resident broadcast loads)
    int my_count = 0;
    bool have_ref = false;

    int T = blockIdx.x;
    if (T < nwork) fetch(T, 0);
    for (; T < nwork; T += gridDim.x) {
        const int b = T / a.tiles, tile = T - b * a.tiles, i0 = tile * BM;
        for (int c = 0; c < a.nchunks; ++c) {
            if (c >= 1) {
                fetch(T, c);
                umma::mbar_wait(&mma_bar, (mma_commits - 1) & 1u);          // the tensor core is done with the stage
            }   // (c == 0: the previous item's epilogue has waited for its last MMAs)
            const bool load_b = a.nchunks > 1 || b_loads == 0;
            if (load_b && tid == 0) {   // this chunk's weights: one bulk copy, landing through the async proxy
                umma::mbar_expect_tx(&b_bar, 2 * B_HALF);
                umma::bulk_g2s(sB, wblk + (size_t)c * (2 * B_HALF / 4), 2 * B_HALF, &b_bar);
            }
            store();
            umma::fence_smem_to_async();
            umma::fence_before_sync();       // (orders the previous item's tcgen05.ld before the MMAs that overwrite TMEM)
            __syncthreads();
            // MMA issue.  Measured on B200 (tools/micro/mma_rate*.cu): ONE thread needs ~120 cycles per tcgen05.mma whatever
            // N, kind or layout, but issuers on different warps overlap (2 issuers: 60, 4: ~45 cycles per MMA and SM).  With
            // N = 16 / 32 an MMA is nowhere near 120 cycles of tensor-core work, so the 54 MMAs of a chunk are issued by
            // three warps, one kernel row each, into their own accumulators: row r adds its hi*hi terms into accumulator 2r
            // and its two 2^-11-sized correction terms into accumulator 2r + 1 (the tensor core accumulates with truncation:
            // keeping the number of accumulation steps per large accumulator small keeps that bias at 2e-7, see the
            // tests); the epilogue adds the six in f32.  hi*hi and hi*lo share their A operand, so they are ONE MMA of
            // N = 2 BN over the weight rows [hi | lo]: 36 MMAs per chunk instead of 54.  The barrier expects three commits.
            if (warp < 3) {
                if (lane == 0) {
                    umma::fence_after_sync();
                    if (load_b) umma::mbar_wait(&b_bar, b_loads & 1u);
                    const int ky = warp;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int tap = ky * 3 + kx;
                        const uint32_t shift = (uint32_t)(ky * BAND + kx) * 16u;     // band dy + 1, position dx + 1
#pragma unroll
                        for (int j = 0; j < CK / 8; ++j) {
                            const uint32_t ao = sbase + shift + (uint32_t)(2 * j) * PLANE_BYTES;
                            const uint32_t bo = sB + (uint32_t)(tap * KG + 2 * j) * (2 * BN * 16);
                            const uint64_t ah = umma::smem_desc(ao, PLANE_BYTES, 128), al = umma::smem_desc(ao + A_HALF, PLANE_BYTES, 128);
                            const uint64_t bd = umma::smem_desc(bo, 2 * BN * 16, 128);     // rows 0 .. BN-1: hi, BN .. 2BN-1: lo
                            const bool first = c == 0 && kx == 0 && j == 0;
                            // A_hi x [B_hi | B_lo] -> [main | correction] in one MMA of N = 2 BN; A_lo x B_hi -> correction
                            umma::mma_tf32(tmem_d + (uint32_t)(2 * ky * BN), ah, bd, idesc(2 * BN), !first);
                            umma::mma_tf32(tmem_d + (uint32_t)((2 * ky + 1) * BN), al, bd, idesc(BN), true);
                        }
                    }
                    umma::commit(&mma_bar);
                }
                __syncwarp();     // the issuing warps stay converged: their other lanes do not run ahead into the barrier spin
            }
            ++mma_commits;
            if (load_b) ++b_loads;
        }
        if (T + (int)gridDim.x < nwork) fetch(T + gridDim.x, 0);     // next item's loads fly during this item's epilogue
        umma::mbar_wait(&mma_bar, (mma_commits - 1) & 1u);
        umma::fence_after_sync();

        // ---- epilogue: lane = position i0 + row; warps 0-3 take the first half of the channels, 4-7 the second
        const int q = i0 + row;
        const int yy = q / Wp, xx = q - yy * Wp;
        const bool valid = yy < H && xx < W;
        // 8 columns at a time (keeps the live registers of the epilogue small); the six accumulators are added in f32.
        // Statistics: the first item with valid positions sets the warp's per-channel reference (its mean over the warp's
        // positions); from then on every valid position adds (y - ref) and (y - ref)^2 to this thread's sums.
        const uint32_t t0 = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)cbase;
        float *yb = a.y + ((size_t)b * a.no + n0 + cbase) * plane + (size_t)yy * W + xx;
        const bool stats = a.tile_stats != nullptr;
        const int nvalid = __popc(__ballot_sync(FULL, valid));
        const bool set_ref = stats && !have_ref && nvalid > 0;
#pragma unroll
        for (int h = 0; h < NC; h += 8) {
            float acc[8];
            {
                uint32_t r[6][8];
#pragma unroll
                for (int k = 0; k < 6; ++k)      // six loads in flight, ONE wait (a wait per load costs ~300 cycles each)
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                                 : "=r"(r[k][0]), "=r"(r[k][1]), "=r"(r[k][2]), "=r"(r[k][3]), "=r"(r[k][4]), "=r"(r[k][5]),
                                   "=r"(r[k][6]), "=r"(r[k][7])
                                 : "r"(t0 + (uint32_t)(k * BN + h)));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 8; ++j)      // mains (even blocks) and corrections (odd blocks) separately, then together
                    acc[j] = ((__uint_as_float(r[0][j]) + __uint_as_float(r[2][j])) + __uint_as_float(r[4][j])) +
                             ((__uint_as_float(r[1][j]) + __uint_as_float(r[3][j])) + __uint_as_float(r[5][j]));
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool in = n0 + cbase + h + j < a.no;
                const float val = acc[j] + (bias_p != nullptr && in ? __ldg(bias_p + h + j) : 0.f);
                if (valid && in) yb[(size_t)(h + j) * plane] = val;
                if (set_ref) {
                    float m = valid ? val : 0.f;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) m += __shfl_xor_sync(FULL, m, off);
                    if (lane == 0) ref[h + j] = m / (float)nvalid;
                    __syncwarp();
                }
                if (stats && valid) {
                    const float d = val - ref[h + j];
                    s1[h + j] += d;
                    s2[h + j] = __fmaf_rn(d, d, s2[h + j]);
                }
            }
        }
        if (set_ref) have_ref = true;
        if (stats && valid) ++my_count;
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_d, TCOLS);

    if (a.tile_stats != nullptr) {
        // per warp and channel: n, mean = ref + S1 / n, M2 = S2 - S1^2 / n; the four row-warps of a channel are merged
        // with Chan's formula: one (count, mean, M2) per CTA and channel
        int cnt_w = my_count;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) cnt_w += __shfl_xor_sync(FULL, cnt_w, off);
        const float nw = (float)cnt_w;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            float t1 = s1[j], t2 = s2[j];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                t1 += __shfl_xor_sync(FULL, t1, off);
                t2 += __shfl_xor_sync(FULL, t2, off);
            }
            if (lane == 0) {
                red_sum[warp & 3][cbase + j] = nw > 0.f ? ref[j] + t1 / nw : 0.f;
                red_m2[warp & 3][cbase + j] = nw > 0.f ? fmaxf(t2 - t1 * t1 / nw, 0.f) : 0.f;
            }
        }
        if (lane == 0 && warp < 4) red_cnt[warp] = cnt_w;
        __syncthreads();
        if (tid < BN && n0 + tid < a.no) {
            float n = 0.f, mu = 0.f, m2 = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const float n2 = (float)red_cnt[w];
                if (n2 > 0.f) {
                    const float tot = n + n2, d = red_sum[w][tid] - mu, f = n2 / tot;
                    mu += d * f;
                    m2 += red_m2[w][tid] + d * d * n * f;
                    n = tot;
                }
            }
            float *ts = a.tile_stats + ((size_t)(n0 + tid) * gridDim.x + blockIdx.x) * 3;
            ts[0] = n;
            ts[1] = mu;
            ts[2] = m2;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// weight gradient: dW[co][ci][ky][kx] += sum_{b, y, x} dY[b][co][y][x] * X[b][ci][y + ky - 1][x + kx - 1]
//
// Persistent CTAs, one (16 output channels) x (CI input channels: 16, or 4 for the three-channel first layer)
// sub-problem each, looping over image tiles of RL rows x TW columns.  A thread owns a 4 x 4 block of (co, ci) pairs
// with all nine taps in registers (144 accumulators, kept across tiles) and ONE ROW of the tile: it sweeps the row's TW
// columns with a 3 x 3 window per input channel in registers, so that a position costs 12 new x values + 4 dY values
// from shared memory for 144 FMAs.  Tiles are staged plane-major ([channel][row][column], exactly the global layout) by
// 16-byte cp.async with zero fill outside the image (the column origin of the x tile is 4 to the left of the tile, so
// every copy is an aligned quad that is entirely inside or outside), double-buffered: tile t + 1 streams in while tile t
// is being reduced.  Row pitches are 4 (mod 8) floats: the lanes of a warp (= rows) hit 8 different banks.
// One shuffle reduction over the row-lanes and one red.global.add per element at the very end.
// ---------------------------------------------------------------------------------------------
struct WArgs {
    int B, cin, cout, H, W, tiles_x, tiles_y, ntiles, vec;
    const float *x, *dy;
    float *dw;
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void *src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async4_zfill(uint32_t dst, const void *src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(dst), "l"(src), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

template <int CI, int TW>
struct WGeom {
    static constexpr int CIB = CI / 4, NBLK = 4 * CIB, RL = THREADS / NBLK;      // row-lanes = rows of a tile
    static constexpr int XC = TW + 8;                                             // staged x columns: origin x0 - 4
    static constexpr int XP = XC + ((4 - XC % 8) + 8) % 8, DP = TW + ((4 - TW % 8) + 8) % 8;   // pitches = 4 (mod 8)
    static constexpr int X_FLOATS = CI * (RL + 2) * XP, D_FLOATS = 16 * RL * DP;
    static constexpr int BYTES = 2 * (X_FLOATS + D_FLOATS) * 4;
};

template <int CI, int TW>
__global__ void __launch_bounds__(THREADS, 1) wgrad_kernel(const WArgs a) {
    using G = WGeom<CI, TW>;
    constexpr int CIB = G::CIB, NBLK = G::NBLK, RL = G::RL, XP = G::XP, DP = G::DP;
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    const int ncig = (a.cin + CI - 1) / CI;
    const int cig = blockIdx.y % ncig, cog = blockIdx.y / ncig;
    const int ci0 = cig * CI, co0 = cog * 16;
    const size_t plane = (size_t)a.H * a.W;
    const int per_image = a.tiles_x * a.tiles_y;

    auto stage = [&](int t, int buf) {
        float *xs = sm + (size_t)buf * (G::X_FLOATS + G::D_FLOATS), *ds = xs + G::X_FLOATS;
        const int b = t / per_image, r_ = t - b * per_image;
        const int ty = r_ / a.tiles_x, tx = r_ - ty * a.tiles_x;
        const int y0 = ty * RL, x0 = tx * TW;
        const float *xb = a.x + (size_t)b * a.cin * plane, *dyb = a.dy + (size_t)b * a.cout * plane;
        if (a.vec) {
            constexpr int XQ = G::XC / 4, DQ = TW / 4;
            for (int it = tid; it < CI * (RL + 2) * XQ; it += THREADS) {
                const int q = it % XQ, r = (it / XQ) % (RL + 2), ch = it / (XQ * (RL + 2));
                const int yy = y0 + r - 1, xx = x0 - 4 + 4 * q;
                const bool ok = ci0 + ch < a.cin && yy >= 0 && yy < a.H && xx >= 0 && xx < a.W;
                const float *src = ok ? xb + (size_t)(ci0 + ch) * plane + (size_t)yy * a.W + xx : a.x;
                cp_async16_zfill(umma::smem_u32(xs + (ch * (RL + 2) + r) * XP + 4 * q), src, ok);
            }
            for (int it = tid; it < 16 * RL * DQ; it += THREADS) {
                const int q = it % DQ, r = (it / DQ) % RL, ch = it / (DQ * RL);
                const int yy = y0 + r, xx = x0 + 4 * q;
                const bool ok = co0 + ch < a.cout && yy < a.H && xx < a.W;
                const float *src = ok ? dyb + (size_t)(co0 + ch) * plane + (size_t)yy * a.W + xx : a.dy;
                cp_async16_zfill(umma::smem_u32(ds + (ch * RL + r) * DP + 4 * q), src, ok);
            }
        } else {       // any width / alignment: element-wise copies
            for (int it = tid; it < CI * (RL + 2) * G::XC; it += THREADS) {
                const int c = it % G::XC, r = (it / G::XC) % (RL + 2), ch = it / (G::XC * (RL + 2));
                const int yy = y0 + r - 1, xx = x0 - 4 + c;
                const bool ok = ci0 + ch < a.cin && yy >= 0 && yy < a.H && xx >= 0 && xx < a.W;
                const float *src = ok ? xb + (size_t)(ci0 + ch) * plane + (size_t)yy * a.W + xx : a.x;
                cp_async4_zfill(umma::smem_u32(xs + (ch * (RL + 2) + r) * XP + c), src, ok);
            }
            for (int it = tid; it < 16 * RL * TW; it += THREADS) {
                const int c = it % TW, r = (it / TW) % RL, ch = it / (TW * RL);
                const int yy = y0 + r, xx = x0 + c;
                const bool ok = co0 + ch < a.cout && yy < a.H && xx < a.W;
                const float *src = ok ? dyb + (size_t)(co0 + ch) * plane + (size_t)yy * a.W + xx : a.dy;
                cp_async4_zfill(umma::smem_u32(ds + (ch * RL + r) * DP + c), src, ok);
            }
        }
    };

    const int row = tid % RL, blk = tid / RL;
    const int cob = blk / CIB, cib = blk - cob * CIB;
    float acc[4][4][9];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int t = 0; t < 9; ++t) acc[i][j][t] = 0.f;

    int t = blockIdx.x, it_ = 0;
    if (t < a.ntiles) stage(t, 0);
    cp_async_commit();
    for (; t < a.ntiles; t += gridDim.x, ++it_) {
        const int buf = it_ & 1;
        if (t + (int)gridDim.x < a.ntiles) stage(t + gridDim.x, buf ^ 1);     // the other buffer was released by the barrier below
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const float *xs = sm + (size_t)buf * (G::X_FLOATS + G::D_FLOATS), *ds = xs + G::X_FLOATS;
        // rows beyond the image hold zeros in both operands: no need to mask the sweep
        const float *xr = xs + ((cib * 4) * (RL + 2) + row) * XP;       // channel cib*4, tile row `row` - 1 (the window's top)
        const float *dr = ds + ((cob * 4) * RL + row) * DP;
        float w[4][3][3];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                w[j][ky][1] = xr[(j * (RL + 2) + ky) * XP + 3];          // column x0 - 1
                w[j][ky][2] = xr[(j * (RL + 2) + ky) * XP + 4];          // column x0
            }
#pragma unroll 4
        for (int c = 0; c < TW; ++c) {
            float d[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) d[i] = dr[i * RL * DP + c];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    w[j][ky][0] = w[j][ky][1];
                    w[j][ky][1] = w[j][ky][2];
                    w[j][ky][2] = xr[(j * (RL + 2) + ky) * XP + c + 5];
                }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int k = 0; k < 9; ++k) acc[i][j][k] = __fmaf_rn(d[i], w[j][k / 3][k % 3], acc[i][j][k]);
        }
        __syncthreads();
    }
    cp_async_wait<0>();
    // ---- reduce over the row-lanes of a block that share a warp, then one atomic per element and warp
    constexpr int LANES = RL < 32 ? RL : 32;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                float v = acc[i][j][k];
#pragma unroll
                for (int off = LANES / 2; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
                acc[i][j][k] = v;
            }
    if (tid % LANES == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int co = co0 + cob * 4 + i;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ci = ci0 + cib * 4 + j;
                if (co < a.cout && ci < a.cin) {
#pragma unroll
                    for (int k = 0; k < 9; ++k) atomicAdd(a.dw + ((size_t)co * a.cin + ci) * 9 + k, acc[i][j][k]);
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

static int conv_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

/* number of persistent CTAs per n-tile of i2p_conv3x3_tc = number of (count, mean, M2) statistics slots per channel */
int i2p_conv3x3_stat_slots(int B, int no, int H, int W) {
    const int nn = i2p::conv::ntiles_of(no);
    long long gx = (2LL * conv_sms() + nn - 1) / nn;          // two resident CTAs per SM over all n-tiles
    const long long work = (long long)i2p_conv3x3_tiles(H, W) * B;
    return (int)(gx > work ? work : gx);
}

int i2p_conv3x3_pack(int cin, int cout, int dgrad, const float *w, float *pack, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(cin >= 1 && cout >= 1 && w != nullptr && pack != nullptr, "conv3x3_pack: bad arguments");
    const long long total = i2p_conv3x3_pack_floats(cin, cout, dgrad) / 2;
    const long long blocks = (total + 255) / 256;
    conv::pack_kernel<<<(int)(blocks < 592 ? blocks : 592), 256, 0, as_stream(stream)>>>(cin, cout, dgrad, w, pack);
    return check_launch("conv3x3_pack");
}

int i2p_conv3x3_pack_multi(int n, const long long *table, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(n >= 0 && n <= 65535 && (n == 0 || table != nullptr), "conv3x3_pack_multi: bad table");
    if (n == 0) return I2P_OK;
    conv::pack_multi_kernel<<<dim3(16, n), 256, 0, as_stream(stream)>>>(table);
    return check_launch("conv3x3_pack_multi");
}

/* y (B, no, H, W) = conv3x3(x (B, ki, H, W), pad 1) [+ bias]; wpack from i2p_conv3x3_pack (forward: cin = ki, cout = no;
 * data gradient: dgrad = 1 with cin = no, cout = ki).  tile_stats (no, B * i2p_conv3x3_tiles(H, W), 3) or null. */
int i2p_conv3x3_tc(int B, int ki, int no, int H, int W, const float *x, const float *wpack, const float *bias, float *y,
                   float *tile_stats, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(B >= 1 && B <= 65535 && ki >= 1 && no >= 1 && H >= 1 && W >= 1, "conv3x3_tc: bad sizes");
    I2P_REQUIRE((long long)(H + 3) * (W + conv::PAD) < (1LL << 20), "conv3x3_tc: image too large (flat padded index must stay below 2^20)");
    I2P_REQUIRE((reinterpret_cast<uintptr_t>(wpack) & 15) == 0, "conv3x3_tc: wpack must be 16-byte aligned");
    conv::Args a;
    a.B = B; a.ki = ki; a.no = no; a.H = H; a.W = W; a.Wp = W + conv::PAD;
    a.inv_wp = 1.0f / (float)a.Wp;
    a.tiles = i2p_conv3x3_tiles(H, W);
    a.nchunks = conv::chunks_of(ki);
    a.x = x; a.wpack = wpack; a.bias = bias; a.y = y; a.tile_stats = tile_stats;
    const int nn = conv::ntiles_of(no);
    dim3 grid((unsigned)i2p_conv3x3_stat_slots(B, no, H, W), nn);
    cudaStream_t s = as_stream(stream);
    const bool vec = W % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
#define I2P_CONV(CK_, BN_, V_)                                                                                            \
    do {                                                                                                                  \
        static bool once = false;                                                                                         \
        if (!once) { conv::allow_smem(conv::conv3x3_tc_kernel<CK_, BN_, V_>, conv::smem_bytes<CK_, BN_>()); once = true; } \
        conv::conv3x3_tc_kernel<CK_, BN_, V_><<<grid, conv::THREADS, conv::smem_bytes<CK_, BN_>(), s>>>(a);                \
    } while (0)
#define I2P_CONV_V(CK_, BN_) do { if (vec) I2P_CONV(CK_, BN_, true); else I2P_CONV(CK_, BN_, false); } while (0)
    const int ck = conv::ck_of(ki), bn = conv::bn_of(no);
    if (ck == 8) { if (bn == 16) I2P_CONV_V(8, 16); else I2P_CONV_V(8, 32); }
    else { if (bn == 16) I2P_CONV_V(16, 16); else I2P_CONV_V(16, 32); }
#undef I2P_CONV_V
#undef I2P_CONV
    return check_launch("conv3x3_tc");
}

/* dw (cout, cin, 3, 3) += weight gradient of y = conv3x3(x, w, pad 1) from dy (B, cout, H, W) and x (B, cin, H, W) */
int i2p_conv3x3_wgrad(int B, int cin, int cout, int H, int W, const float *x, const float *dy, float *dw, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(B >= 1 && B <= 65535 && cin >= 1 && cout >= 1 && H >= 1 && W >= 1, "conv3x3_wgrad: bad sizes");
    conv::WArgs a;
    a.B = B; a.cin = cin; a.cout = cout; a.H = H; a.W = W;
    a.x = x; a.dy = dy; a.dw = dw;
    a.vec = (W % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(dy) % 16 == 0);
    cudaStream_t s = as_stream(stream);
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
#define I2P_WGRAD(CI_, TW_)                                                                                    \
    do {                                                                                                       \
        using G = conv::WGeom<CI_, TW_>;                                                                       \
        static bool once = false;                                                                              \
        if (!once) { conv::allow_smem(conv::wgrad_kernel<CI_, TW_>, G::BYTES); once = true; }                  \
        a.tiles_x = ceil_div(W, TW_);                                                                          \
        a.tiles_y = ceil_div(H, G::RL);                                                                        \
        a.ntiles = a.tiles_x * a.tiles_y * B;                                                                  \
        const int groups = ceil_div(cin, CI_) * ceil_div(cout, 16);                                            \
        int gx = sms / groups;                                                                                 \
        if (gx < 1) gx = 1;                                                                                    \
        if (gx > a.ntiles) gx = a.ntiles;                                                                      \
        dim3 grid(gx, groups);                                                                                 \
        conv::wgrad_kernel<CI_, TW_><<<grid, conv::THREADS, wgrad_smem(conv::wgrad_kernel<CI_, TW_>, G::BYTES), s>>>(a); \
    } while (0)
    if (cin <= 4) I2P_WGRAD(4, 8);
    else I2P_WGRAD(16, 32);
#undef I2P_WGRAD
    return check_launch("conv3x3_wgrad");
}
}
