// Shared per-point MLP on the 5th-generation tensor cores: forward, dX and dW GEMMs of mlp.cu's layer
// algebra as tcgen05.mma kind::tf32 with the 3-term hi/lo split (f32-accurate, see umma.cuh) and the
// accumulator in tensor memory.
//
// Why the operands are staged by the CTA's threads and not by TMA: every activation operand needs a
// per-element transform on its way in -- the previous layer's batch-norm + activation (forward, dW) or
// the reconstruction of dY from (g, y) (dX, dW) -- and the split into (hi, lo) tf32 halves.  What TMA
// would buy (no register round trip) does not apply; what matters instead is the instruction count of
// that round trip, which bounded the first version of the forward kernel (ncu: 19 instructions per
// element, MIO-throttle stalls from scalar STS).  Here
//   * a thread moves 4 consecutive k of one row: one LDG.128 (or .64 / .32 when the row stride does not
//     allow it), the transform, one STS.128 for hi and one for lo -- the UMMA canonical no-swizzle
//     layouts keep 16 bytes of K (K-major) or of M/N (MN-major) contiguous, and the lane -> (row, k)
//     mapping makes every 8-lane store phase cover 128 contiguous bytes (no bank conflicts);
//   * the weight operand is split and laid out ONCE per layer and step by a tiny pack kernel; the GEMM
//     CTAs then fetch it with cp.async 16-byte copies (no registers, no arithmetic);
//   * dW reads both operands in their natural row-major form as MN-major UMMA operands (K = rows), so
//     nothing is transposed through shared memory.  (For 32-bit elements the tensor core takes MN-major
//     operands in one layout only, SWIZZLE_128B_BASE32B; with the no-swizzle layout it silently
//     accumulates zeros -- measured, and CUTLASS's sm100 builder asserts the same.)
// One CTA = one 128-row (forward, dX) or 128-channel (dW) tile, 256 threads, one shared-memory stage;
// three to four CTAs share an SM so that staging, tensor-core work and epilogues of different tiles
// overlap (the same occupancy-based latency hiding as the FMA kernels).
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

extern "C" int i2p_get_mlp_tensor_cores(void);

namespace i2p {
namespace tc {

constexpr int BM = 128, BK = 32, THREADS = 256;
constexpr int MAXC = 256;   // widest batch-normalised layer the dY tables hold

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_addr), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void sts128(unsigned char *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    *reinterpret_cast<uint4 *>(p) = make_uint4(a, b, c, d);
}

// split 4 values and store the hi / lo quadruples (16 bytes each) at the same offset of the two tiles
__device__ __forceinline__ void split_store4(unsigned char *hi_tile, unsigned char *lo_tile, uint32_t off, float v0,
                                             float v1, float v2, float v3) {
    uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
    umma::split_tf32(v0, h0, l0);
    umma::split_tf32(v1, h1, l1);
    umma::split_tf32(v2, h2, l2);
    umma::split_tf32(v3, h3, l3);
    sts128(hi_tile + off, h0, h1, h2, h3);
    sts128(lo_tile + off, l0, l1, l2, l3);
}

// 4 consecutive floats of a row whose start is only VEC-element aligned; elements at or beyond `valid` read 0
template <int VEC>
__device__ __forceinline__ float4 load4(const float *p, int valid) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (VEC == 4) {
        if (valid >= 4) v = __ldg(reinterpret_cast<const float4 *>(p));
    } else if (VEC == 2) {
        if (valid >= 2) { const float2 t = __ldg(reinterpret_cast<const float2 *>(p)); v.x = t.x; v.y = t.y; }
        if (valid >= 4) { const float2 t = __ldg(reinterpret_cast<const float2 *>(p + 2)); v.z = t.x; v.w = t.y; }
    } else {
        if (valid >= 1) v.x = __ldg(p);
        if (valid >= 2) v.y = __ldg(p + 1);
        if (valid >= 3) v.z = __ldg(p + 2);
        if (valid >= 4) v.w = __ldg(p + 3);
    }
    return v;
}

// ---------------------------------------------------------------------------------------------
// weight pack: W (cout, cin) -> tf32 (hi, lo) halves in the UMMA canonical K-major layout, per
// (n-tile, k-chunk) block [hi: BN x 32 | lo: BN x 32], zero-padded.  Two sections:
//   F (forward, y = A W^T):  n = output channel, k = input channel;  BN = 64 if cout <= 64 else 128
//   X (dX = dY W):           n = input channel,  k = output channel; BN = 64 if cin  <= 64 else 128
// ---------------------------------------------------------------------------------------------
struct PackGeom {
    int bn_f, nt_f, nc_f, bn_x, nt_x, nc_x;
    long long floats_f, floats_x;
};

__host__ __device__ inline PackGeom pack_geom(int cin, int cout) {
    PackGeom g;
    g.bn_f = cout <= 64 ? 64 : 128;
    g.nt_f = (cout + g.bn_f - 1) / g.bn_f;
    g.nc_f = (cin + BK - 1) / BK;
    g.floats_f = (long long)g.nt_f * g.nc_f * 2 * g.bn_f * BK;
    g.bn_x = cin <= 64 ? 64 : 128;
    g.nt_x = (cin + g.bn_x - 1) / g.bn_x;
    g.nc_x = (cout + BK - 1) / BK;
    g.floats_x = (long long)g.nt_x * g.nc_x * 2 * g.bn_x * BK;
    return g;
}

// layout 0: no-swizzle canonical order ((k / 4) * bn + n) * 4 + k % 4; layout 1: SWIZZLE_128B atoms of 8 rows x 128 B
// (row n of the tile holds its 32 k contiguously, the 16-byte chunk index XORed with n % 8)
__device__ __forceinline__ void pack_weights_body(int cin, int cout, int layout, const float *__restrict__ w, float *__restrict__ pack,
                                                  int block, int nblocks) {
    const PackGeom g = pack_geom(cin, cout);
    const long long half_f = g.floats_f / 2, half_x = g.floats_x / 2;
    for (long long e = (long long)block * 256 + threadIdx.x; e < half_f + half_x; e += (long long)nblocks * 256) {
        const bool fwd = e < half_f;
        const long long q = fwd ? e : e - half_f;
        const int bn = fwd ? g.bn_f : g.bn_x, nc = fwd ? g.nc_f : g.nc_x;
        const int per = bn * BK;                       // elements of one (tile, chunk) half-block
        const long long blk = q / per;
        const int r = (int)(q - blk * per);
        const int nt = (int)(blk / nc), c = (int)(blk - (long long)nt * nc);
        int nl, kl;   // (row, k) of the element stored at float offset r of the half-block
        if (layout == 0) {
            const int k4 = r / (bn * 4), rem = r - k4 * bn * 4;
            nl = rem >> 2;
            kl = k4 * 4 + (rem & 3);
        } else {
            nl = r >> 5;                                    // 32 floats per row
            kl = ((((r >> 2) & 7) ^ (nl & 7)) << 2) + (r & 3);
        }
        const int n = nt * bn + nl, k = c * BK + kl;
        float v = 0.f;
        if (fwd) { if (n < cout && k < cin) v = __ldg(w + (size_t)n * cin + k); }
        else     { if (n < cin && k < cout) v = __ldg(w + (size_t)k * cin + n); }
        uint32_t hi, lo;
        umma::split_tf32(v, hi, lo);
        float *base = pack + (fwd ? 0 : g.floats_f) + blk * 2 * per;
        base[r] = __uint_as_float(hi);
        base[per + r] = __uint_as_float(lo);
    }
}

__global__ void __launch_bounds__(256) pack_weights_kernel(int cin, int cout, int layout, const float *__restrict__ w,
                                                           float *__restrict__ pack) {
    pack_weights_body(cin, cout, layout, w, pack, blockIdx.x, gridDim.x);
}

// Every layer of a model in one launch: table[l] = { cin | cout << 32, weight pointer, pack pointer }, blockIdx.y = layer.
// (The weights change once per optimiser step; packing them layer by layer inside the forward pass put ~40 launches on
// the serial chain of the step.)
__global__ void __launch_bounds__(256) pack_weights_multi_kernel(int layout, const long long *__restrict__ table) {
    const long long *t = table + (size_t)blockIdx.y * 3;
    const int cin = (int)(t[0] & 0xffffffffll), cout = (int)(t[0] >> 32);
    pack_weights_body(cin, cout, layout, reinterpret_cast<const float *>(t[1]), reinterpret_cast<float *>(t[2]), blockIdx.x, gridDim.x);
}

// instruction descriptor, kind::tf32, f32 accumulate, M = 128; a_mn / b_mn: the operand is MN-major
__host__ __device__ constexpr uint32_t idesc_tf32(int n, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// One chunk of K = 32: four k-steps of 8, each as hi*hi + lo*hi + hi*lo into the same accumulator.
// `a_step` / `b_step`: byte distance between consecutive k-steps of an operand tile.
__device__ __forceinline__ void issue_chunk(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t a_lbo, uint32_t a_sbo,
                                            uint32_t a_step, uint32_t b_hi, uint32_t b_lo, uint32_t b_lbo, uint32_t b_sbo,
                                            uint32_t b_step, uint32_t idesc, bool first_chunk) {
#pragma unroll
    for (int j = 0; j < BK / 8; ++j) {
        const uint64_t ah = umma::smem_desc(a_hi + j * a_step, a_lbo, a_sbo), al = umma::smem_desc(a_lo + j * a_step, a_lbo, a_sbo);
        const uint64_t bh = umma::smem_desc(b_hi + j * b_step, b_lbo, b_sbo), bl = umma::smem_desc(b_lo + j * b_step, b_lbo, b_sbo);
        umma::mma_tf32(tmem_d, ah, bh, idesc, !(first_chunk && j == 0));
        umma::mma_tf32(tmem_d, al, bh, idesc, true);
        umma::mma_tf32(tmem_d, ah, bl, idesc, true);
    }
}

// The same for SWIZZLE_128B K-major tiles (8-row atoms of 1024 B; a k-step of 8 tf32 is 32 bytes further inside
// the atom; layout type 2, SBO = 1024, LBO = 16 as CUTLASS encodes the swizzled K-major canonical layout).
__device__ __forceinline__ void issue_chunk_sw(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                               uint32_t idesc, bool first_chunk) {
#pragma unroll
    for (int j = 0; j < BK / 8; ++j) {
        const uint64_t ah = umma::smem_desc(a_hi + j * 32, 16, 1024, 2), al = umma::smem_desc(a_lo + j * 32, 16, 1024, 2);
        const uint64_t bh = umma::smem_desc(b_hi + j * 32, 16, 1024, 2), bl = umma::smem_desc(b_lo + j * 32, 16, 1024, 2);
        umma::mma_tf32(tmem_d, ah, bh, idesc, !(first_chunk && j == 0));
        umma::mma_tf32(tmem_d, al, bh, idesc, true);
        umma::mma_tf32(tmem_d, ah, bl, idesc, true);
    }
}

// The same with TWO MMAs per k-step instead of three.  One thread needs ~120 cycles to issue a tcgen05.mma whatever its
// N (tools/micro/mma_rate*.cu), so the count of MMAs, not their size, is what a chunk costs.  The packed weights hold the
// lo tile right behind the hi tile, both as 8-row SWIZZLE_128B atoms: together they ARE one K-major operand of 2 BN rows.
// A_hi x [B_hi | B_lo] is therefore a single MMA of N = 2 BN whose columns [0, BN) receive hi*hi and [BN, 2 BN) hi*lo;
// A_lo x B_hi (N = BN) is accumulated into the second half as well, so the large terms and the 2^-11-sized correction
// terms also keep separate accumulators (the tensor core accumulates with truncation).  The epilogue adds the halves.
template <int BN>
__device__ __forceinline__ void issue_chunk_sw2(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, bool first_chunk) {
    constexpr uint32_t id_wide = idesc_tf32(2 * BN, false, false), id_half = idesc_tf32(BN, false, false);
#pragma unroll
    for (int j = 0; j < BK / 8; ++j) {
        const uint64_t ah = umma::smem_desc(a_hi + j * 32, 16, 1024, 2), al = umma::smem_desc(a_lo + j * 32, 16, 1024, 2);
        const uint64_t bb = umma::smem_desc(b_hi + j * 32, 16, 1024, 2);
        umma::mma_tf32(tmem_d, ah, bb, id_wide, !(first_chunk && j == 0));
        umma::mma_tf32(tmem_d + BN, al, bb, id_half, true);
    }
}

// TMEM allocation + mbarrier set-up shared by the three kernels
template <int COLS>
__device__ __forceinline__ uint32_t tc_prologue(uint64_t *bar, uint32_t *slot) {
    if ((threadIdx.x >> 5) == 0) umma::tmem_alloc(slot, COLS);
    if (threadIdx.x == 0) {
        umma::mbar_init(bar, 1);
        umma::mbar_fence_init();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    return *slot;
}

// accumulator (128 lanes x BN columns) -> shared tile[128][BN + 4] (+ per-column addend), conflict-free 128-bit stores
template <int BN, bool WIDE = false>
__device__ __forceinline__ void tmem_to_tile(uint32_t tmem_d, float *tile, const float *addend, int n0, int nvalid) {
    constexpr int LDT = BN + 4, CQ = BN / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = (warp & 3) * 32 + lane;      // a warp can only read its own 32 TMEM lanes
    const int cbase = (warp >> 2) * CQ;          // warps 0-3: first half of the columns, 4-7: second half
#pragma unroll
    for (int c0 = 0; c0 < CQ; c0 += 16) {
        float v[16];
        umma::tmem_ld16(tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cbase + c0), v);
        if (WIDE) {      // [hi*hi | correction terms]: two accumulators of BN columns
            float u[16];
            umma::tmem_ld16(tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(BN + cbase + c0), u);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += u[j];
        }
        if (addend != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = n0 + cbase + c0 + j;
                v[j] += n < nvalid ? __ldg(addend + n) : 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4 *>(&tile[row * LDT + cbase + c0 + j]) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
}

// tile[128][BN+4] -> out (rows, ld) at (r0, n0), coalesced; 128-bit stores when the row stride allows
template <int BN>
__device__ __forceinline__ void store_tile(const float *tile, float *out, int ld, int r0, int vr, int n0, int nvalid) {
    constexpr int LDT = BN + 4;
    if ((ld & 3) == 0) {
        for (int e = threadIdx.x; e < BM * (BN / 4); e += THREADS) {
            const int row = e / (BN / 4), c4 = e % (BN / 4);
            if (row < vr && n0 + c4 * 4 < nvalid)   // ld % 4 == 0 and nvalid == ld: a quad is all in or all out
                *reinterpret_cast<float4 *>(out + (size_t)(r0 + row) * ld + n0 + c4 * 4) =
                    *reinterpret_cast<const float4 *>(&tile[row * LDT + c4 * 4]);
        }
    } else {
        for (int e = threadIdx.x; e < BM * BN; e += THREADS) {
            const int row = e / BN, c = e % BN;
            if (row < vr && n0 + c < nvalid) out[(size_t)(r0 + row) * ld + n0 + c] = tile[row * LDT + c];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// forward:  y = act(x * in_scale + in_shift) W^T + bias, per-tile (mean, M2) of y
// ---------------------------------------------------------------------------------------------
struct FwdArgs {
    int rows, cin, cout;
    const float *x, *in_scale, *in_shift;
    float in_slope;
    const float *wpack, *bias;
    float *y, *tile_stats;
};

// Lane -> operand coordinates of the K-major activation tile (128 rows x 32 k): a warp pass covers 8 rows x
// 16 k, lanes 0-7 / 8-15 / 16-23 / 24-31 hold the four 16-byte k-groups of rows 0-7, so that each group's
// 128-bit stores of one phase are 128 contiguous bytes.  Warp w owns k-half (w & 1) and the row blocks
// (w >> 1) + 4 p, p = 0..3: one thread sees a single k-group per chunk.
struct KMajorCoords {
    int rowl, kg, rbase;
    uint32_t soff;   // byte offset of pass 0 inside a tile; pass p adds p * 4 * 128
    __device__ KMajorCoords() {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        rowl = lane & 7;
        kg = (warp & 1) * 4 + (lane >> 3);
        rbase = (warp >> 1) * 8 + rowl;          // row of pass p: rbase + 32 p
        soff = (uint32_t)kg * (BM * 16) + (uint32_t)rbase * 16;
    }
};

// SW variant: the tile is SWIZZLE_128B K-major -- row r holds its 32 k (128 B) contiguously at (r / 8) * 1024 +
// (r % 8) * 128 with the 16-byte chunk index XORed with r % 8.  Lanes 0-7 are the eight k-groups of one row (one
// full 128-byte line of global memory, stored to 128 contiguous bytes of shared memory), a warp pass covers 4 rows,
// warp w owns rows 4 w + (lane >> 3) + 32 p.  Half as many L1 wavefronts per load as the no-swizzle mapping, whose
// 8-lane phases had to span 8 rows: the load/store unit, not HBM or the tensor pipe, bounded that version.
struct SwizzledCoords {
    int kg, rbase;
    uint32_t soff;   // byte offset of pass 0 inside a tile; pass p adds p * 4096
    __device__ SwizzledCoords() {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        kg = lane & 7;
        rbase = warp * 4 + (lane >> 3);          // row of pass p: rbase + 32 p
        soff = (uint32_t)(rbase >> 3) * 1024u + (uint32_t)(rbase & 7) * 128u + ((uint32_t)(kg ^ (rbase & 7)) << 4);
    }
};

// A hi | A lo | B hi | B lo, or the epilogue tile that overlays them, whichever is larger; SW: + 1 KB so that the
// kernel can align the operand tiles to the 1024-byte swizzle atom itself
template <int BN, int SW = 0>
__host__ __device__ constexpr int fwd_smem_bytes() {
    return ((2 * BM * BK * 4 + 2 * BN * BK * 4) > (BM * (BN + 4) * 4) ? (2 * BM * BK * 4 + 2 * BN * BK * 4) : (BM * (BN + 4) * 4)) +
           (SW ? 1024 : 0);
}

// SW = 0: no-swizzle operand tiles, weights by cp.async.  SW = 1: SWIZZLE_128B tiles with line-coalesced operand
// loads (SwizzledCoords) and the weights of a chunk delivered by ONE bulk copy (TMA engine) that thread 0 issues
// as soon as the previous chunk's MMAs have released the buffer -- no LSU work for the weight operand at all.
template <int BN, int VEC, int SW>
__global__ void __launch_bounds__(THREADS, BN == 128 ? 3 : 4) fwd_kernel(const FwdArgs a) {
    constexpr uint32_t A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;
    constexpr uint32_t A_LBO = BM * 16, B_LBO = BN * 16, SBO = 128;
    extern __shared__ __align__(1024) unsigned char smem_raw[];   // A hi | A lo | B hi | B lo ; epilogue tile overlays
    __shared__ uint64_t mma_bar, b_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float part_n[THREADS / BN][BN], part_mu[THREADS / BN][BN], part_m2[THREADS / BN][BN];

    const int tid = threadIdx.x;
    const int nt = blockIdx.x, n0 = nt * BN, r0 = blockIdx.y * BM;
    if (SW && tid == 0) umma::mbar_init(&b_bar, 1);
    constexpr int TCOLS = SW == 2 ? 2 * BN : BN;      // SW == 2: [hi*hi | correction] accumulators, two MMAs per k-step
    const uint32_t tmem_d = tc_prologue<TCOLS>(&mma_bar, &tmem_slot);
    unsigned char *smem = SW ? smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u) : smem_raw;
    const uint32_t sbase = umma::smem_u32(smem);
    constexpr uint32_t idesc = idesc_tf32(BN, false, false);
    const KMajorCoords co;
    const SwizzledCoords cs;
    const int kg = SW ? cs.kg : co.kg, rbase = SW ? cs.rbase : co.rbase;
    const uint32_t soff = SW ? cs.soff : co.soff, pstep = SW ? 4096u : 512u;
    const int nchunks = (a.cin + BK - 1) / BK;
    const bool has_tf = a.in_scale != nullptr;
    const float *wblk = a.wpack + (size_t)nt * nchunks * 2 * BN * BK;

    float4 ra[4];
    auto fetch = [&](int c) {   // issue the global loads of chunk c; nothing depends on them until it is staged
        const int k = c * BK + kg * 4;
        const int kvalid = a.cin - k;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int r = r0 + rbase + 32 * p;
            ra[p] = load4<VEC>(a.x + (size_t)r * a.cin + k, r < a.rows ? kvalid : 0);
        }
    };
    auto copy_b = [&](int c) {   // weights of chunk c: the packed (hi | lo) block, byte for byte
        const float *wsrc = wblk + (size_t)c * 2 * BN * BK;
        if (SW) {
            if (tid == 0) {
                umma::mbar_expect_tx(&b_bar, 2 * B_BYTES);
                umma::bulk_g2s(sbase + 2 * A_BYTES, wsrc, B_BYTES, &b_bar);
                umma::bulk_g2s(sbase + 2 * A_BYTES + B_BYTES, wsrc + BN * BK, B_BYTES, &b_bar);
            }
        } else {
            constexpr int N16 = 2 * B_BYTES / 16;
#pragma unroll
            for (int i = 0; i < N16 / THREADS; ++i)
                cp_async16(sbase + 2 * A_BYTES + (uint32_t)(tid + i * THREADS) * 16, wsrc + (size_t)(tid + i * THREADS) * 4);
        }
    };

    fetch(0);
    for (int c = 0; c < nchunks; ++c) {
        float4 sc4 = make_float4(1.f, 1.f, 1.f, 1.f), sh4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_tf) {   // per-channel constants of the fused input transform (L1-resident)
            const int k = c * BK + kg * 4;
            sc4 = load4<1>(a.in_scale + k, a.cin - k);
            sh4 = load4<1>(a.in_shift + k, a.cin - k);
        }
        if (c >= 1) umma::mbar_wait(&mma_bar, (uint32_t)(c - 1) & 1u);   // the tensor core is done with the stage
        copy_b(c);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float4 v = ra[p];
            if (has_tf) {   // previous layer's normalise + activation; max(z, slope z) is act for 0 <= slope <= 1
                float z;
                z = __fmaf_rn(v.x, sc4.x, sh4.x); v.x = fmaxf(z, z * a.in_slope);
                z = __fmaf_rn(v.y, sc4.y, sh4.y); v.y = fmaxf(z, z * a.in_slope);
                z = __fmaf_rn(v.z, sc4.z, sh4.z); v.z = fmaxf(z, z * a.in_slope);
                z = __fmaf_rn(v.w, sc4.w, sh4.w); v.w = fmaxf(z, z * a.in_slope);
            }
            split_store4(smem, smem + A_BYTES, soff + (uint32_t)p * pstep, v.x, v.y, v.z, v.w);
        }
        if (!SW) cp_async_wait_all();
        umma::fence_smem_to_async();
        __syncthreads();
        if (tid == 0) {
            umma::fence_after_sync();
            const uint32_t sb = sbase + 2 * A_BYTES;
            if (SW) {
                umma::mbar_wait(&b_bar, (uint32_t)c & 1u);   // the bulk copy of this chunk's weights has landed
                if (SW == 2) issue_chunk_sw2<BN>(tmem_d, sbase, sbase + A_BYTES, sb, c == 0);
                else issue_chunk_sw(tmem_d, sbase, sbase + A_BYTES, sb, sb + B_BYTES, idesc, c == 0);
            } else {
                issue_chunk(tmem_d, sbase, sbase + A_BYTES, A_LBO, SBO, 2 * A_LBO, sb, sb + B_BYTES, B_LBO, SBO, 2 * B_LBO, idesc, c == 0);
            }
            umma::commit(&mma_bar);
        }
        if (c + 1 < nchunks) fetch(c + 1);   // global loads in flight while the tensor core runs
    }
    umma::mbar_wait(&mma_bar, (uint32_t)(nchunks - 1) & 1u);
    umma::fence_after_sync();
    __syncthreads();

    float *tile = reinterpret_cast<float *>(smem);
    constexpr int LDT = BN + 4;
    tmem_to_tile<BN, SW == 2>(tmem_d, tile, a.bias, n0, a.cout);
    umma::fence_before_sync();
    __syncthreads();
    if ((tid >> 5) == 0) umma::tmem_dealloc(tmem_d, TCOLS);

    const int vr = min(BM, a.rows - r0);
    store_tile<BN>(tile, a.y, a.cout, r0, vr, n0, a.cout);
    if (a.tile_stats == nullptr) return;
    // per-tile (mean, M2): THREADS / BN row parts per column (lanes = consecutive columns), merged with Chan's formula
    constexpr int PARTS = THREADS / BN, RPP = BM / PARTS;
    {
        const int col = tid % BN, part = tid / BN;
        const int rb0 = part * RPP, re0 = min(vr, rb0 + RPP);
        float sum = 0.f;
        for (int r = rb0; r < re0; ++r) sum += tile[r * LDT + col];
        const float n = (float)max(re0 - rb0, 0);
        const float mu = n > 0.f ? sum / n : 0.f;
        float m2 = 0.f;
        for (int r = rb0; r < re0; ++r) { const float d = tile[r * LDT + col] - mu; m2 += d * d; }
        part_n[part][col] = n; part_mu[part][col] = mu; part_m2[part][col] = m2;
    }
    __syncthreads();
    if (tid < BN && n0 + tid < a.cout) {
        float n = part_n[0][tid], mu = part_mu[0][tid], m2 = part_m2[0][tid];
#pragma unroll
        for (int p = 1; p < PARTS; ++p) {
            const float n2 = part_n[p][tid];
            if (n2 > 0.f) {
                const float tot = n + n2, d = part_mu[p][tid] - mu, f = n2 / tot;
                mu += d * f;
                m2 += part_m2[p][tid] + d * d * n * f;
                n = tot;
            }
        }
        float *ts = a.tile_stats + ((size_t)blockIdx.y * a.cout + n0 + tid) * 2;
        ts[0] = mu;
        ts[1] = m2;
    }
}

// ---------------------------------------------------------------------------------------------
// dY tables: the per-channel constants that turn (g, y) into dY = scale (dz - S1/n - yhat S2/n)
// ---------------------------------------------------------------------------------------------
struct BnRef {
    const float *y, *mean, *rstd, *scale, *shift;
    float slope;
};

struct DyTab {   // [6][MAXC]: scale, shift, mean, rstd, s1n, s2n
    float v[6][MAXC];
};

__device__ __forceinline__ void load_dytab(DyTab &t, const BnRef &bn, const double *s12, int c, long long rows) {
    for (int i = threadIdx.x; i < c; i += THREADS) {
        t.v[0][i] = bn.scale[i]; t.v[1][i] = bn.shift[i]; t.v[2][i] = bn.mean[i]; t.v[3][i] = bn.rstd[i];
        t.v[4][i] = (float)(s12[i] / (double)rows);
        t.v[5][i] = (float)(s12[c + i] / (double)rows);
    }
}

__device__ __forceinline__ float dy1(float g, float y, float sc, float sh, float mu, float rs, float s1n, float s2n, float slope) {
    const float dz = g * (__fmaf_rn(y, sc, sh) > 0.f ? 1.f : slope);
    return sc * (dz - s1n - ((y - mu) * rs) * s2n);
}

// 4 channels ch..ch+3 of one row
__device__ __forceinline__ float4 dy4(const DyTab &t, int ch, float4 g, float4 y, float slope) {
    const float4 sc = *reinterpret_cast<const float4 *>(&t.v[0][ch]), sh = *reinterpret_cast<const float4 *>(&t.v[1][ch]);
    const float4 mu = *reinterpret_cast<const float4 *>(&t.v[2][ch]), rs = *reinterpret_cast<const float4 *>(&t.v[3][ch]);
    const float4 s1 = *reinterpret_cast<const float4 *>(&t.v[4][ch]), s2 = *reinterpret_cast<const float4 *>(&t.v[5][ch]);
    return make_float4(dy1(g.x, y.x, sc.x, sh.x, mu.x, rs.x, s1.x, s2.x, slope), dy1(g.y, y.y, sc.y, sh.y, mu.y, rs.y, s1.y, s2.y, slope),
                       dy1(g.z, y.z, sc.z, sh.z, mu.z, rs.z, s1.z, s2.z, slope), dy1(g.w, y.w, sc.w, sh.w, mu.w, rs.w, s1.w, s2.w, slope));
}

// ---------------------------------------------------------------------------------------------
// dX = dY W  (rows x cin), dY rebuilt from (g, y) while the A tile is staged; K = cout (multiple of 32).
// Epilogue: store, and the batch-norm backward sums of the PREVIOUS layer from the tile.
// ---------------------------------------------------------------------------------------------
// Gradient w.r.t. a layer's activated output: dense (rows, cout), or -- MAXK -- the gradient dout (rows / k, cout) of a
// max-over-k output, routed to the row whose index inside its group equals arg (rows / k, cout).
struct GSrc {
    const float *g;       // dense, or dout
    const int32_t *arg;   // MAXK only
    int k, kshift;        // kshift = log2(k) for a power of two (every group size of the model), else -1
    __device__ __forceinline__ int group_of(int r) const { return kshift >= 0 ? (r >> kshift) : (r / k); }
};

template <bool MAXK>
struct GQuad {            // what a thread keeps in registers for 4 channels of one row between fetch and stage
    float4 v;
    int4 a;
};
template <>
struct GQuad<false> {
    float4 v;
};

template <bool MAXK>
__device__ __forceinline__ GQuad<MAXK> g_fetch(const GSrc &gs, long long r, long long grp, int ch, int cout, bool valid) {
    GQuad<MAXK> q;
    if constexpr (MAXK) {
        q.v = load4<4>(gs.g + (size_t)grp * cout + ch, valid ? 4 : 0);
        q.a = valid ? __ldg(reinterpret_cast<const int4 *>(gs.arg + (size_t)grp * cout + ch)) : make_int4(-1, -1, -1, -1);
    } else {
        q.v = load4<4>(gs.g + (size_t)r * cout + ch, valid ? 4 : 0);
    }
    return q;
}
template <bool MAXK>
__device__ __forceinline__ float4 g_resolve(const GQuad<MAXK> &q, int kk) {
    if constexpr (MAXK)
        return make_float4(q.a.x == kk ? q.v.x : 0.f, q.a.y == kk ? q.v.y : 0.f, q.a.z == kk ? q.v.z : 0.f, q.a.w == kk ? q.v.w : 0.f);
    else
        return q.v;
}

struct DxArgs {
    int rows, cin, cout;
    GSrc gs;             // gradient w.r.t. this layer's activated output
    BnRef bn;
    const double *s12;
    const float *wpack;  // section X of the pack
    float *dx;
    BnRef prev;          // prev.y == nullptr: the input is not a batch-normalised layer
    double *prev_s12;
};

template <int BN, int SW = 0>
__host__ __device__ constexpr int dx_smem_bytes() { return fwd_smem_bytes<BN, SW>() + (int)sizeof(DyTab); }

template <int BN, int SW, bool MAXK>
__global__ void __launch_bounds__(THREADS, MAXK ? 2 : 3) dx_kernel(const DxArgs a) {
    constexpr uint32_t A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;
    constexpr uint32_t A_LBO = BM * 16, B_LBO = BN * 16, SBO = 128;
    constexpr int OPER = fwd_smem_bytes<BN, 0>();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t mma_bar, b_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float part1[THREADS / BN][BN], part2[THREADS / BN][BN];
    unsigned char *smem = SW ? smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u) : smem_raw;
    DyTab &tab = *reinterpret_cast<DyTab *>(smem + OPER);

    const int tid = threadIdx.x;
    const int nt = blockIdx.x, n0 = nt * BN, r0 = blockIdx.y * BM;
    load_dytab(tab, a.bn, a.s12, a.cout, a.rows);
    if (SW && tid == 0) umma::mbar_init(&b_bar, 1);
    constexpr int TCOLS = SW == 2 ? 2 * BN : BN;
    const uint32_t tmem_d = tc_prologue<TCOLS>(&mma_bar, &tmem_slot);   // (its barrier also publishes the tables)
    const uint32_t sbase = umma::smem_u32(smem);
    constexpr uint32_t idesc = idesc_tf32(BN, false, false);
    const KMajorCoords co;
    const SwizzledCoords cs;
    const int kg = SW ? cs.kg : co.kg, rbase = SW ? cs.rbase : co.rbase;
    const uint32_t soff = SW ? cs.soff : co.soff, pstep = SW ? 4096u : 512u;
    const int nchunks = a.cout / BK;
    const float *wblk = a.wpack + (size_t)nt * nchunks * 2 * BN * BK;

    GQuad<MAXK> rg[4];
    float4 ry[4];
    int grp[4], kk[4];   // MAXK: group of each of the thread's rows and the row's index inside it
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int r = r0 + rbase + 32 * p;
        grp[p] = MAXK ? a.gs.group_of(r) : 0;
        kk[p] = MAXK ? r - grp[p] * a.gs.k : 0;
    }
    auto fetch = [&](int c) {
        const int k = c * BK + kg * 4;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int r = r0 + rbase + 32 * p;
            rg[p] = g_fetch<MAXK>(a.gs, r, grp[p], k, a.cout, r < a.rows);
            ry[p] = load4<4>(a.bn.y + (size_t)r * a.cout + k, r < a.rows ? 4 : 0);
        }
    };
    auto copy_b = [&](int c) {
        const float *wsrc = wblk + (size_t)c * 2 * BN * BK;
        if (SW) {
            if (tid == 0) {
                umma::mbar_expect_tx(&b_bar, 2 * B_BYTES);
                umma::bulk_g2s(sbase + 2 * A_BYTES, wsrc, B_BYTES, &b_bar);
                umma::bulk_g2s(sbase + 2 * A_BYTES + B_BYTES, wsrc + BN * BK, B_BYTES, &b_bar);
            }
        } else {
            constexpr int N16 = 2 * B_BYTES / 16;
#pragma unroll
            for (int i = 0; i < N16 / THREADS; ++i)
                cp_async16(sbase + 2 * A_BYTES + (uint32_t)(tid + i * THREADS) * 16, wsrc + (size_t)(tid + i * THREADS) * 4);
        }
    };

    fetch(0);
    for (int c = 0; c < nchunks; ++c) {
        if (c >= 1) umma::mbar_wait(&mma_bar, (uint32_t)(c - 1) & 1u);
        copy_b(c);
        const int ch = c * BK + kg * 4;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const float4 v = dy4(tab, ch, g_resolve<MAXK>(rg[p], kk[p]), ry[p], a.bn.slope);   // rows beyond `rows`: finite, never stored
            split_store4(smem, smem + A_BYTES, soff + (uint32_t)p * pstep, v.x, v.y, v.z, v.w);
        }
        if (!SW) cp_async_wait_all();
        umma::fence_smem_to_async();
        __syncthreads();
        if (tid == 0) {
            umma::fence_after_sync();
            const uint32_t sb = sbase + 2 * A_BYTES;
            if (SW) {
                umma::mbar_wait(&b_bar, (uint32_t)c & 1u);
                if (SW == 2) issue_chunk_sw2<BN>(tmem_d, sbase, sbase + A_BYTES, sb, c == 0);
                else issue_chunk_sw(tmem_d, sbase, sbase + A_BYTES, sb, sb + B_BYTES, idesc, c == 0);
            } else {
                issue_chunk(tmem_d, sbase, sbase + A_BYTES, A_LBO, SBO, 2 * A_LBO, sb, sb + B_BYTES, B_LBO, SBO, 2 * B_LBO, idesc, c == 0);
            }
            umma::commit(&mma_bar);
        }
        if (c + 1 < nchunks) fetch(c + 1);
    }
    umma::mbar_wait(&mma_bar, (uint32_t)(nchunks - 1) & 1u);
    umma::fence_after_sync();
    __syncthreads();

    float *tile = reinterpret_cast<float *>(smem);
    constexpr int LDT = BN + 4;
    tmem_to_tile<BN, SW == 2>(tmem_d, tile, nullptr, n0, a.cin);
    umma::fence_before_sync();
    __syncthreads();
    if ((tid >> 5) == 0) umma::tmem_dealloc(tmem_d, TCOLS);

    const int vr = min(BM, a.rows - r0);
    store_tile<BN>(tile, a.dx, a.cin, r0, vr, n0, a.cin);
    if (a.prev.y == nullptr) return;
    constexpr int PARTS = THREADS / BN, RPP = BM / PARTS;
    {
        const int col = tid % BN, part = tid / BN, ch = n0 + col;
        float s1 = 0.f, s2 = 0.f;
        if (ch < a.cin) {
            const float mu = a.prev.mean[ch], rs = a.prev.rstd[ch], sc = a.prev.scale[ch], sh = a.prev.shift[ch];
            const int rb0 = part * RPP, re0 = min(vr, rb0 + RPP);
            for (int r = rb0; r < re0; ++r) {
                const float yv = __ldg(a.prev.y + (size_t)(r0 + r) * a.cin + ch);
                const float dz = tile[r * LDT + col] * (__fmaf_rn(yv, sc, sh) > 0.f ? 1.f : a.prev.slope);
                s1 += dz;
                s2 += dz * ((yv - mu) * rs);
            }
        }
        part1[part][col] = s1;
        part2[part][col] = s2;
    }
    __syncthreads();
    if (tid < BN && n0 + tid < a.cin) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int p = 0; p < PARTS; ++p) { t1 += part1[p][tid]; t2 += part2[p][tid]; }
        atomicAdd(a.prev_s12 + n0 + tid, (double)t1);
        atomicAdd(a.prev_s12 + a.cin + n0 + tid, (double)t2);
    }
}

// ---------------------------------------------------------------------------------------------
// dW^T tile:  D[i, o] = sum_r A_prev[r, i] dY[r, o]   (M = 128 input channels, N = BN output channels,
// K = a chunk of rows), both operands MN-major straight from their row-major tensors; the CTA's
// partial sum is added to dw (cout, cin) with coalesced atomics.
// ---------------------------------------------------------------------------------------------
struct DwArgs {
    int rows, cin, cout, rows_per_block;
    GSrc gs;
    BnRef bn;
    const double *s12;
    const float *x;                 // (rows, cin) raw input / previous raw output
    const float *prev_scale, *prev_shift;
    float prev_slope;
    float *dw;
};

template <int BN>
__host__ __device__ constexpr int dw_smem_bytes() { return 2 * BM * BK * 4 + 2 * BN * BK * 4 + (int)sizeof(DyTab); }

// MN-major operand tile (MN channels x 32 rows of K).  For 32-bit operands the tensor core accepts an MN-major
// tile in exactly one shared-memory layout, UMMA SWIZZLE_128B_BASE32B: atoms of 4 k-rows x 32 channels (4 x 128 B,
// channels contiguous), inside which the 32-byte chunk index (channel / 8) is XORed with the k-row index; atom
// (channel block mb, row block kb) sits at mb * LBO + kb * SBO with LBO = 4096 (eight row blocks), SBO = 512.
// One MMA (K = 8) reads two row blocks: the descriptor start advances by 1024 per k-step.
// A warp pass stages one atom: lane -> (k-row r4, 16-byte half, 8-channel chunk c8); every 8-lane store phase
// then covers 4 rows x 32 B in 4 different chunk positions = all 32 banks once, and every global row is read
// as 128 contiguous bytes.
template <int BN, int VEC, bool MAXK>
__global__ void __launch_bounds__(THREADS, 2) dw_kernel(const DwArgs a) {
    constexpr uint32_t A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;
    constexpr uint32_t LBO = 4096, SBO = 512, KSTEP_BYTES = 1024;
    extern __shared__ __align__(1024) unsigned char smem[];   // A hi | A lo | B hi | B lo | dY tables
    __shared__ uint64_t mma_bar;
    __shared__ uint32_t tmem_slot;
    DyTab &tab = *reinterpret_cast<DyTab *>(smem + 2 * A_BYTES + 2 * B_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // tile index fastest: the CTAs that share a chunk of rows (every input-channel tile re-stages the same dY, every
    // output-channel tile the same x) are scheduled together, so the re-reads hit L2 instead of HBM
    const int mtiles = (a.cin + BM - 1) / BM;
    const int m0 = (blockIdx.x % mtiles) * BM, n0 = (blockIdx.x / mtiles) * BN;   // input-channel tile, output-channel tile
    const long long rb = (long long)blockIdx.y * a.rows_per_block;
    const long long re = min((long long)a.rows, rb + a.rows_per_block);
    load_dytab(tab, a.bn, a.s12, a.cout, a.rows);
    const uint32_t tmem_d = tc_prologue<BN>(&mma_bar, &tmem_slot);
    const uint32_t sbase = umma::smem_u32(smem);
    constexpr uint32_t idesc = idesc_tf32(BN, true, true);

    const int r4 = lane & 3, half = (lane >> 2) & 1, c8 = lane >> 3;
    const uint32_t in_atom = (uint32_t)r4 * 128u + ((uint32_t)(c8 ^ r4) << 5) + (uint32_t)half * 16u;
    // A operand (input channels): warp w owns channel block (w & 3) and the row blocks (w >> 2) + 2 p, p = 0..3
    const int a_kb0 = warp >> 2;
    const int a_ch = m0 + (warp & 3) * 32 + c8 * 8 + half * 4;
    const uint32_t a_off = (uint32_t)(warp & 3) * LBO + (uint32_t)a_kb0 * SBO + in_atom;   // + p * 2 * SBO
    // B operand (output channels): BN / 32 channel blocks; row blocks b_kb0 + p * B_KSTEP
    constexpr int B_MB = BN / 32, B_KSTEP = 8 / B_MB, B_PASSES = B_MB;   // BN = 128: 4, 2, 4;  BN = 64: 2, 4, 2
    const int b_kb0 = warp / B_MB;
    const int b_ch = n0 + (warp % B_MB) * 32 + c8 * 8 + half * 4;
    const uint32_t b_off = (uint32_t)(warp % B_MB) * LBO + (uint32_t)b_kb0 * SBO + in_atom;   // + p * B_KSTEP * SBO

    const bool has_tf = a.prev_scale != nullptr;
    float4 psc = make_float4(1.f, 1.f, 1.f, 1.f), psh = make_float4(0.f, 0.f, 0.f, 0.f);
    const int a_valid = a.cin - a_ch;          // > 0: some of the 4 channels exist
    if (has_tf && a_valid > 0) {
        psc = load4<1>(a.prev_scale + a_ch, a_valid);
        psh = load4<1>(a.prev_shift + a_ch, a_valid);
    }
    const bool b_in = b_ch < a.cout;           // cout % 4 == 0: a quad is all in or all out

    float4 rx[4], ry[B_PASSES];
    GQuad<MAXK> rg[B_PASSES];
    auto fetch = [&](long long r0) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const long long r = r0 + (a_kb0 + 2 * p) * 4 + r4;
            rx[p] = load4<VEC>(a.x + (size_t)r * a.cin + a_ch, r < re ? a_valid : 0);
        }
#pragma unroll
        for (int p = 0; p < B_PASSES; ++p) {
            const long long r = r0 + (b_kb0 + p * B_KSTEP) * 4 + r4;
            const bool valid = r < re && b_in;
            rg[p] = g_fetch<MAXK>(a.gs, r, MAXK ? a.gs.group_of((int)r) : 0, b_ch, a.cout, valid);
            ry[p] = load4<4>(a.bn.y + (size_t)r * a.cout + b_ch, valid ? 4 : 0);
        }
    };

    const int nchunks = (int)((re - rb + BK - 1) / BK);
    fetch(rb);
    for (int c = 0; c < nchunks; ++c) {
        const long long r0 = rb + (long long)c * BK;
        if (c >= 1) umma::mbar_wait(&mma_bar, (uint32_t)(c - 1) & 1u);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float4 v = rx[p];
            if (has_tf) {
                float z;
                z = __fmaf_rn(v.x, psc.x, psh.x); v.x = fmaxf(z, z * a.prev_slope);
                z = __fmaf_rn(v.y, psc.y, psh.y); v.y = fmaxf(z, z * a.prev_slope);
                z = __fmaf_rn(v.z, psc.z, psh.z); v.z = fmaxf(z, z * a.prev_slope);
                z = __fmaf_rn(v.w, psc.w, psh.w); v.w = fmaxf(z, z * a.prev_slope);
            }
            // rows beyond the chunk end / channels beyond cin: the dY operand is forced to zero for such rows, and
            // columns beyond cin are never written back, so finite values suffice here
            split_store4(smem, smem + A_BYTES, a_off + (uint32_t)p * 2u * SBO, v.x, v.y, v.z, v.w);
        }
#pragma unroll
        for (int p = 0; p < B_PASSES; ++p) {
            const long long r = r0 + (b_kb0 + p * B_KSTEP) * 4 + r4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < re && b_in) v = dy4(tab, b_ch, g_resolve<MAXK>(rg[p], MAXK ? (int)r - a.gs.group_of((int)r) * a.gs.k : 0), ry[p], a.bn.slope);
            split_store4(smem + 2 * A_BYTES, smem + 2 * A_BYTES + B_BYTES, b_off + (uint32_t)(p * B_KSTEP) * SBO, v.x, v.y, v.z, v.w);
        }
        umma::fence_smem_to_async();
        __syncthreads();
        if (tid == 0) {
            umma::fence_after_sync();
#pragma unroll
            for (int j = 0; j < BK / 8; ++j) {   // layout type 1 = SWIZZLE_128B_BASE32B
                const uint64_t ah = umma::smem_desc(sbase + j * KSTEP_BYTES, LBO, SBO, 1);
                const uint64_t al = umma::smem_desc(sbase + A_BYTES + j * KSTEP_BYTES, LBO, SBO, 1);
                const uint64_t bh = umma::smem_desc(sbase + 2 * A_BYTES + j * KSTEP_BYTES, LBO, SBO, 1);
                const uint64_t bl = umma::smem_desc(sbase + 2 * A_BYTES + B_BYTES + j * KSTEP_BYTES, LBO, SBO, 1);
                umma::mma_tf32(tmem_d, ah, bh, idesc, !(c == 0 && j == 0));
                umma::mma_tf32(tmem_d, al, bh, idesc, true);
                umma::mma_tf32(tmem_d, ah, bl, idesc, true);
            }
            umma::commit(&mma_bar);
        }
        if (c + 1 < nchunks) fetch(r0 + BK);
    }
    if (nchunks > 0) umma::mbar_wait(&mma_bar, (uint32_t)(nchunks - 1) & 1u);
    umma::fence_after_sync();

    // epilogue: lane = input channel, 16 output channels per tcgen05.ld; coalesced red.add into dw (cout, cin)
    if (nchunks > 0) {
        const int i = m0 + (warp & 3) * 32 + lane;
        constexpr int CQ = BN / 2;
        const int cbase = (warp >> 2) * CQ;
#pragma unroll
        for (int c0 = 0; c0 < CQ; c0 += 16) {
            float v[16];
            umma::tmem_ld16(tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cbase + c0), v);
            if (i < a.cin) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int o = n0 + cbase + c0 + j;
                    if (o < a.cout) atomicAdd(a.dw + (size_t)o * a.cin + i, v[j]);
                }
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_d, BN);
}

template <typename K>
static void allow_smem(K kernel, int bytes) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

// widest aligned access a (rows, ld) tensor at `p` allows for 4 consecutive elements of a row
static int vec_of(int ld, const void *p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    if ((ld & 3) == 0 && (a & 15) == 0) return 4;
    if ((ld & 1) == 0 && (a & 7) == 0) return 2;
    return 1;
}
static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Operand layout of the forward / dX kernels: bit 16 of the tensor-core mask (i2p_set_mlp_tensor_cores) selects the
// SWIZZLE_128B + bulk-copy variant.  Measured and dropped on the way here: a second register set for the activation
// operand (prefetch distance 2: 118.8 -> 120.9 us forward) and double-buffered cp.async weights (116.7 vs 120.8 us):
// the kernels were bound by load/store-unit wavefronts, not by the depth of either pipeline.
static int swizzled() { const int m = i2p_get_mlp_tensor_cores(); return (m & 16) ? ((m & 64) ? 2 : 1) : 0; }   // bit 64: two MMAs per k-step (issue_chunk_sw2)

}  // namespace tc
}  // namespace i2p

extern "C" {

/* which shapes the tensor-core kernels cover; kind: 0 forward, 1 dX, 2 dW */
int i2p_pw_tc_supported(int kind, int rows, int cin, int cout) {
    using namespace i2p::tc;
    if (rows < 1 || cin < 1 || cout < 64 || cout % 32 != 0) return 0;
    if (kind == 0) return cin <= 4096 ? 1 : 0;
    if (cout > MAXC) return 0;
    return cin >= 32 ? 1 : 0;   // skinny inputs (xyz encodings) stay on the FMA kernels
}

long long i2p_pw_pack_floats(int cin, int cout) {
    const i2p::tc::PackGeom g = i2p::tc::pack_geom(cin, cout);
    return g.floats_f + g.floats_x;
}

int i2p_pw_pack_weights(int cin, int cout, const float *w, float *pack, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(cin >= 1 && cout >= 1, "pw_pack_weights: bad sizes");
    const tc::PackGeom g = tc::pack_geom(cin, cout);
    const long long total = (g.floats_f + g.floats_x) / 2;
    const long long blocks = (total + 255) / 256;
    tc::pack_weights_kernel<<<(int)(blocks < 592 ? blocks : 592), 256, 0, as_stream(stream)>>>(cin, cout, tc::swizzled() ? 1 : 0, w, pack);
    return check_launch("pw_pack_weights");
}

int i2p_pw_pack_weights_multi(int n, const long long *table, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(n >= 0 && n <= 65535 && (n == 0 || table != nullptr), "pw_pack_weights_multi: bad table");
    if (n == 0) return I2P_OK;
    tc::pack_weights_multi_kernel<<<dim3(16, n), 256, 0, as_stream(stream)>>>(tc::swizzled() ? 1 : 0, table);
    return check_launch("pw_pack_weights_multi");
}

int i2p_pw_linear_fwd_tc(int rows, int cin, int cout, const float *x, const float *in_scale, const float *in_shift,
                         float in_slope, const float *wpack, const float *bias, float *y, float *tile_stats, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(i2p_pw_tc_supported(0, rows, cin, cout), "pw_linear_fwd_tc: shape not covered (cout %% 32 == 0, cout >= 64)");
    I2P_REQUIRE(in_slope >= 0.f && in_slope <= 1.f, "pw_linear_fwd_tc: activation slope must be in [0, 1]");
    I2P_REQUIRE(tc::aligned16(wpack) && tc::aligned16(y) && tc::aligned16(in_scale) && tc::aligned16(in_shift),
                "pw_linear_fwd_tc: wpack, y, in_scale, in_shift must be 16-byte aligned");
    tc::FwdArgs a{rows, cin, cout, x, in_scale, in_shift, in_slope, wpack, bias, y, tile_stats};
    const tc::PackGeom g = tc::pack_geom(cin, cout);
    cudaStream_t s = as_stream(stream);
    dim3 grid(g.nt_f, ceil_div(rows, tc::BM));
    const int vec = tc::vec_of(cin, x);
    const int bdb = tc::swizzled();
#define I2P_FWD(BN_, V_, P_)                                                                                    \
    do {                                                                                                        \
        static bool once = false;                                                                               \
        if (!once) { tc::allow_smem(tc::fwd_kernel<BN_, V_, P_>, tc::fwd_smem_bytes<BN_, P_>()); once = true; } \
        tc::fwd_kernel<BN_, V_, P_><<<grid, tc::THREADS, tc::fwd_smem_bytes<BN_, P_>(), s>>>(a);                \
    } while (0)
#define I2P_FWD_V(BN_, P_) do { if (vec == 4) I2P_FWD(BN_, 4, P_); else if (vec == 2) I2P_FWD(BN_, 2, P_); else I2P_FWD(BN_, 1, P_); } while (0)
    if (g.bn_f == 64) { if (bdb == 2) I2P_FWD_V(64, 2); else if (bdb) I2P_FWD_V(64, 1); else I2P_FWD_V(64, 0); }
    else { if (bdb == 2) I2P_FWD_V(128, 2); else if (bdb) I2P_FWD_V(128, 1); else I2P_FWD_V(128, 0); }
#undef I2P_FWD_V
#undef I2P_FWD
    return check_launch("pw_linear_fwd_tc");
}

static bool make_gsrc(i2p::tc::GSrc &gs, const float *g_dense, const float *dout, const int32_t *arg, int k) {
    gs.kshift = 0;
    if (g_dense != nullptr) { gs.g = g_dense; gs.arg = nullptr; gs.k = 1; return i2p::tc::aligned16(g_dense); }
    gs.g = dout; gs.arg = arg; gs.k = k;
    gs.kshift = -1;
    if (k >= 1 && (k & (k - 1)) == 0) { gs.kshift = 0; while ((1 << gs.kshift) < k) ++gs.kshift; }
    return dout != nullptr && arg != nullptr && k >= 1 && i2p::tc::aligned16(dout) && i2p::tc::aligned16(arg);
}

int i2p_pw_linear_bwd_dx_tc(int rows, int cin, int cout, const float *g_dense, const float *dout, const int32_t *arg, int k,
                            const float *y, const float *mean, const float *rstd, const float *scale, const float *shift,
                            float slope, const double *s12, const float *wpack, float *dx, const float *prev_y,
                            const float *prev_mean, const float *prev_rstd, const float *prev_scale, const float *prev_shift,
                            float prev_slope, double *prev_s12, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(i2p_pw_tc_supported(1, rows, cin, cout), "pw_linear_bwd_dx_tc: shape not covered");
    tc::DxArgs a;
    I2P_REQUIRE(make_gsrc(a.gs, g_dense, dout, arg, k), "pw_linear_bwd_dx_tc: no (16-byte aligned) gradient source");
    I2P_REQUIRE(tc::aligned16(y) && tc::aligned16(wpack) && tc::aligned16(dx), "pw_linear_bwd_dx_tc: y, wpack, dx must be 16-byte aligned");
    const tc::PackGeom g = tc::pack_geom(cin, cout);
    a.rows = rows; a.cin = cin; a.cout = cout;
    a.bn = tc::BnRef{y, mean, rstd, scale, shift, slope};
    a.s12 = s12; a.wpack = wpack + g.floats_f; a.dx = dx;
    a.prev = tc::BnRef{prev_y, prev_mean, prev_rstd, prev_scale, prev_shift, prev_slope};
    a.prev_s12 = prev_s12;
    cudaStream_t s = as_stream(stream);
    dim3 grid(g.nt_x, ceil_div(rows, tc::BM));
    const int sw = tc::swizzled();
    const bool maxk = g_dense == nullptr;
#define I2P_DX(BN_, P_, M_)                                                                                       \
    do {                                                                                                          \
        static bool once = false;                                                                                 \
        if (!once) { tc::allow_smem(tc::dx_kernel<BN_, P_, M_>, tc::dx_smem_bytes<BN_, P_>()); once = true; }     \
        tc::dx_kernel<BN_, P_, M_><<<grid, tc::THREADS, tc::dx_smem_bytes<BN_, P_>(), s>>>(a);                    \
    } while (0)
#define I2P_DX_M(BN_, P_) do { if (maxk) I2P_DX(BN_, P_, true); else I2P_DX(BN_, P_, false); } while (0)
    if (g.bn_x == 64) { if (sw == 2) I2P_DX_M(64, 2); else if (sw) I2P_DX_M(64, 1); else I2P_DX_M(64, 0); }
    else { if (sw == 2) I2P_DX_M(128, 2); else if (sw) I2P_DX_M(128, 1); else I2P_DX_M(128, 0); }
#undef I2P_DX_M
#undef I2P_DX
    return check_launch("pw_linear_bwd_dx_tc");
}

int i2p_pw_linear_bwd_dw_tc(int rows, int cin, int cout, const float *g_dense, const float *dout, const int32_t *arg, int k,
                            const float *y, const float *mean, const float *rstd, const float *scale, const float *shift,
                            float slope, const double *s12, const float *x, const float *prev_scale, const float *prev_shift,
                            float prev_slope, float *dw, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(i2p_pw_tc_supported(2, rows, cin, cout), "pw_linear_bwd_dw_tc: shape not covered");
    I2P_REQUIRE(prev_slope >= 0.f && prev_slope <= 1.f, "pw_linear_bwd_dw_tc: activation slope must be in [0, 1]");
    tc::DwArgs a;
    I2P_REQUIRE(make_gsrc(a.gs, g_dense, dout, arg, k), "pw_linear_bwd_dw_tc: no (16-byte aligned) gradient source");
    I2P_REQUIRE(tc::aligned16(y), "pw_linear_bwd_dw_tc: y must be 16-byte aligned");
    a.rows = rows; a.cin = cin; a.cout = cout;
    a.bn = tc::BnRef{y, mean, rstd, scale, shift, slope};
    a.s12 = s12; a.x = x; a.prev_scale = prev_scale; a.prev_shift = prev_shift; a.prev_slope = prev_slope; a.dw = dw;
    const int bn = cout <= 64 ? 64 : 128;
    const int mt = ceil_div(cin, tc::BM), ntl = ceil_div(cout, bn);
    // about two waves of CTAs (2 resident per SM), at least 8 chunks of rows per CTA
    static int ctas = -1;      // I2P_DW_CTAS: tuning override of the CTA count aimed at (default: two waves at two CTAs per SM)
    if (ctas < 0) { const char *e = getenv("I2P_DW_CTAS"); ctas = e ? atoi(e) : 2 * 2 * 148; if (ctas < 1) ctas = 1; }
    int chunks = (ctas + mt * ntl - 1) / (mt * ntl);
    int rpb = (rows + chunks - 1) / chunks;
    rpb = ((rpb + tc::BK - 1) / tc::BK) * tc::BK;
    if (rpb < 8 * tc::BK) rpb = 8 * tc::BK;
    a.rows_per_block = rpb;
    cudaStream_t s = as_stream(stream);
    dim3 grid(mt * ntl, ceil_div(rows, rpb));
    const int vec = tc::vec_of(cin, x);
    const bool maxk = g_dense == nullptr;
#define I2P_DW(BN_, V_, M_)                                                                                   \
    do {                                                                                                      \
        static bool once = false;                                                                             \
        if (!once) { tc::allow_smem(tc::dw_kernel<BN_, V_, M_>, tc::dw_smem_bytes<BN_>()); once = true; }     \
        tc::dw_kernel<BN_, V_, M_><<<grid, tc::THREADS, wgrad_smem(tc::dw_kernel<BN_, V_, M_>, tc::dw_smem_bytes<BN_>()), s>>>(a); \
    } while (0)
#define I2P_DW_V(BN_, M_) do { if (vec == 4) I2P_DW(BN_, 4, M_); else if (vec == 2) I2P_DW(BN_, 2, M_); else I2P_DW(BN_, 1, M_); } while (0)
    if (bn == 64) { if (maxk) I2P_DW_V(64, true); else I2P_DW_V(64, false); }
    else { if (maxk) I2P_DW_V(128, true); else I2P_DW_V(128, false); }
#undef I2P_DW_V
#undef I2P_DW
    return check_launch("pw_linear_bwd_dw_tc");
}
}
