// Furthest point sampling (reference kernel: furthest_point_sampling_kernel<bs>,
// pointnet2/src/sampling_gpu.cu:86-209; launcher :211-253; block-size rule cuda_utils.h:10-14).
//
// Reference design: ONE block per cloud; every one of the M iterations re-reads all N points
// and the N running distances from global memory and runs a shared-memory tree reduction with
// up to 11 block barriers.
// This design: a thread-block CLUSTER per cloud (1..16 CTAs, chosen from N and B).  Each CTA
// keeps its slice of the cloud on chip for the whole run -- running distances in registers,
// coordinates in registers (<= 16 points/thread) or shared memory (32 points/thread) -- so
// HBM is touched once (12N + 4N bytes in, 4M + 4N out).  One iteration = register update,
// two redux.sync, a 32-byte record per warp stored straight into every CTA's shared memory
// over DSMEM, ONE cluster barrier, and a redundant per-warp reduction of the records, which
// also carries the winner's coordinates so that no dependent global load sits on the
// critical path.
//
// Bit-exactness: the reference's winner among equal distances is decided by its thread
// mapping (thread t = k mod bs keeps its smallest k; the tree keeps the lower slot on ties,
// so the thread with the smallest BIT-REVERSED id wins).  tie_key() encodes exactly that
// order and the reductions here compare (distance, tie_key) explicitly.
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace i2p {

constexpr int FPS_THREADS = 512;
constexpr int FPS_WARPS = FPS_THREADS / 32;
constexpr int FPS_MAX_CLUSTER = 16;

// One record per warp and iteration: the warp winner's coordinates and index (16 bytes) and, in a separate array, its
// (distance bits, tie key) pair (8 bytes).  The reduction over the records reads only the key array, with consecutive
// lanes on consecutive 8-byte words -- conflict-free (32-byte records put lanes 8 words apart: an 8-way bank conflict
// on every key load, ~1000 shared-memory cycles per iteration at 128 records).
struct __align__(16) FpsRec {
    float x, y, z;
    int k;
};

struct FpsArgs {
    int n, m;
    int bs_mask, bs_log2, nper;  // reference block size bs = bs_mask + 1; nper = ceil(n / bs)
    const float *dataset;
    float *temp;
    int32_t *idxs;
};

// Smaller key wins among equal distances: (bit-reversed reference thread id, k / bs).
__device__ __forceinline__ unsigned tie_key(int k, const FpsArgs &a) {
    const unsigned t = (unsigned)k & (unsigned)a.bs_mask;
    const unsigned rev = a.bs_log2 == 0 ? 0u : (__brev(t) >> (32 - a.bs_log2));
    return rev * (unsigned)a.nper + ((unsigned)k >> a.bs_log2);
}

template <int PPT, bool SMEM_XYZ>
__global__ void __launch_bounds__(FPS_THREADS) fps_kernel(const FpsArgs a) {
    extern __shared__ __align__(16) unsigned char fps_smem[];
    FpsRec *recs = reinterpret_cast<FpsRec *>(fps_smem);  // [2][FPS_WARPS * csize]
    uint2 *keys = reinterpret_cast<uint2 *>(fps_smem + 2 * FPS_WARPS * FPS_MAX_CLUSTER * sizeof(FpsRec));  // same shape
    // coordinates of this CTA's slice, [p][tid]: the winner's are read from here by index (no select chain over the
    // register copies); with SMEM_XYZ (32 points per thread) the distance update reads them from here as well
    float *sx = reinterpret_cast<float *>(fps_smem + 2 * FPS_WARPS * FPS_MAX_CLUSTER * (sizeof(FpsRec) + sizeof(uint2)));
    float *sy = sx + PPT * FPS_THREADS;
    float *sz = sy + PPT * FPS_THREADS;

    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    const int cloud = blockIdx.x / csize;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n, m = a.m;

    const float *data = a.dataset + (size_t)cloud * n * 3;
    float *temp = a.temp + (size_t)cloud * n;
    int32_t *idxs = a.idxs + (size_t)cloud * m;

    // this thread's points: k = crank*PPT*T + p*T + tid
    const int kbase = crank * PPT * FPS_THREADS + tid;
    float px[SMEM_XYZ ? 1 : PPT], py[SMEM_XYZ ? 1 : PPT], pz[SMEM_XYZ ? 1 : PPT], td[PPT];
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        const int k = kbase + p * FPS_THREADS;
        // slots beyond the cloud keep a NEGATIVE running distance: min() leaves it negative for ever, so it never
        // equals the (non-negative) warp maximum and needs no index test in the loop
        float x = 0.f, y = 0.f, z = 0.f, t = -1.f;
        if (k < n) {
            x = data[(size_t)k * 3 + 0];
            y = data[(size_t)k * 3 + 1];
            z = data[(size_t)k * 3 + 2];
            t = temp[k];  // caller pre-fills 1e10 (pointnet2_utils.py:56)
        }
        sx[p * FPS_THREADS + tid] = x;
        sy[p * FPS_THREADS + tid] = y;
        sz[p * FPS_THREADS + tid] = z;
        if (!SMEM_XYZ) { px[p] = x; py[p] = y; pz[p] = z; }
        td[p] = t;
    }

    float x1 = data[0], y1 = data[1], z1 = data[2];  // old = 0 (sampling_gpu.cu:113)
    if (crank == 0 && tid == 0) idxs[0] = 0;
    const int nrec = FPS_WARPS * csize;
    if (csize > 1) cluster.sync();  // every CTA of the cluster is resident before remote stores

    for (int j = 1; j < m; ++j) {
        // --- update running distances, track the local maximum (sampling_gpu.cu:124-138)
        float best = 0.f;
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            float x2, y2, z2;
            if (SMEM_XYZ) {
                x2 = sx[p * FPS_THREADS + tid]; y2 = sy[p * FPS_THREADS + tid]; z2 = sz[p * FPS_THREADS + tid];
            } else {
                x2 = px[p]; y2 = py[p]; z2 = pz[p];
            }
            const float d = sqlen(__fsub_rn(x2, x1), __fsub_rn(y2, y1), __fsub_rn(z2, z1));
            const float d2 = fminf(d, td[p]);
            td[p] = d2;
            best = fmaxf(best, d2);
        }
        // --- warp winner under (distance desc, tie_key asc).  Only the lanes that hold the warp maximum (one, unless
        // distances tie exactly) look for the slots attaining it and evaluate their tie keys.
        const unsigned wd = __reduce_max_sync(FULL, __float_as_uint(best));
        unsigned tie = 0xffffffffu;
        int pbest = 0;
        if (__float_as_uint(best) == wd) {
            unsigned hit = 0;
#pragma unroll
            for (int p = 0; p < PPT; ++p) hit |= (__float_as_uint(td[p]) == wd) ? (1u << p) : 0u;
            while (hit) {
                const int p = __ffs(hit) - 1;
                hit &= hit - 1;
                const unsigned tk = tie_key(kbase + p * FPS_THREADS, a);
                if (tk < tie) { tie = tk; pbest = p; }
            }
        }
        const unsigned wt = __reduce_min_sync(FULL, tie);
        const int src = __ffs(__ballot_sync(FULL, tie == wt)) - 1;  // tie keys are unique per point
        const float bx = sx[pbest * FPS_THREADS + tid], by = sy[pbest * FPS_THREADS + tid], bz = sz[pbest * FPS_THREADS + tid];
        FpsRec r;
        r.x = __shfl_sync(FULL, bx, src);
        r.y = __shfl_sync(FULL, by, src);
        r.z = __shfl_sync(FULL, bz, src);
        r.k = __shfl_sync(FULL, kbase + pbest * FPS_THREADS, src);
        const uint2 key = make_uint2(wd, wt);  // tie 0xffffffff when this warp holds no real point
        FpsRec *buf = recs + (j & 1) * (FPS_WARPS * FPS_MAX_CLUSTER);
        uint2 *kbuf = keys + (j & 1) * (FPS_WARPS * FPS_MAX_CLUSTER);
        if (lane < csize) {
            FpsRec *dst = csize > 1 ? cluster.map_shared_rank(buf, lane) : buf;
            uint2 *kdst = csize > 1 ? cluster.map_shared_rank(kbuf, lane) : kbuf;
            dst[crank * FPS_WARPS + warp] = r;
            kdst[crank * FPS_WARPS + warp] = key;
        }
        if (csize > 1) cluster.sync(); else __syncthreads();

        // --- every warp reduces all records redundantly -> identical winner everywhere
        unsigned gd = 0, gt = 0xffffffffu;
        int gi = 0;
        for (int i = lane; i < nrec; i += 32) {
            const uint2 kv = kbuf[i];
            const unsigned d = kv.x, t = kv.y;
            if (t != 0xffffffffu && (d > gd || (d == gd && t < gt))) { gd = d; gt = t; gi = i; }
        }
        const unsigned fd = __reduce_max_sync(FULL, gt != 0xffffffffu ? gd : 0u);
        const unsigned ft = __reduce_min_sync(FULL, (gd == fd) ? gt : 0xffffffffu);
        const int wl = __ffs(__ballot_sync(FULL, gd == fd && gt == ft && gt != 0xffffffffu)) - 1;
        const int wi = __shfl_sync(FULL, gi, wl);
        x1 = buf[wi].x; y1 = buf[wi].y; z1 = buf[wi].z;
        if (crank == 0 && tid == 0) idxs[j] = buf[wi].k;
    }

    // the reference leaves the final running distances in `temp`
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        const int k = kbase + p * FPS_THREADS;
        if (k < n) temp[k] = td[p];
    }
    if (csize > 1) cluster.sync();  // no CTA exits while a peer may still address its shared memory
}

// cuda_utils.h:10-14, evaluated on the host exactly like the reference does
static int ref_block_size(int n) {
    const int pow_2 = (int)(std::log(static_cast<double>(n)) / std::log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

template <int PPT, bool SMEM_XYZ>
static int launch_fps(const FpsArgs &a, int b, int csize, cudaStream_t stream) {
    const size_t smem = 2 * FPS_WARPS * FPS_MAX_CLUSTER * (sizeof(FpsRec) + sizeof(uint2)) +
                        (size_t)3 * PPT * FPS_THREADS * sizeof(float);
    auto kern = fps_kernel<PPT, SMEM_XYZ>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess && csize > 8) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) {
        set_error("furthest_point_sampling: %s", cudaGetErrorString(e));
        return I2P_ERR_CUDA;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(b * csize));
    cfg.blockDim = dim3(FPS_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, a);
    if (e != cudaSuccess) {
        g_launches.fetch_add(1, std::memory_order_relaxed);
        set_error("furthest_point_sampling (cluster %d, %d pts/thread): %s", csize, PPT, cudaGetErrorString(e));
        (void)cudaGetLastError();
        return I2P_ERR_CUDA;
    }
    return check_launch("furthest_point_sampling");
}

}  // namespace i2p

extern "C" int i2p_furthest_point_sampling(int b, int n, int m, const float *dataset, float *temp,
                                           int32_t *idxs, void *stream) {
    using namespace i2p;
    I2P_REQUIRE(b >= 0 && n >= 1 && m >= 0, "furthest_point_sampling: bad sizes b=%d n=%d m=%d", b, n, m);
    if (b == 0 || m <= 0) return I2P_OK;  // sampling_gpu.cu:101
    FpsArgs a;
    a.n = n; a.m = m; a.dataset = dataset; a.temp = temp; a.idxs = idxs;
    const int bs = ref_block_size(n);
    a.bs_mask = bs - 1;
    a.bs_log2 = 0;
    while ((1 << a.bs_log2) < bs) ++a.bs_log2;
    a.nper = (n + bs - 1) / bs;

    // choose cluster size / points per thread: estimated cycles per iteration x waves
    int best_c = 0, best_ppt = 0;
    double best_cost = 1e300;
    for (int c = 1; c <= FPS_MAX_CLUSTER; c *= 2) {
        int ppt = 1;
        while ((long long)c * FPS_THREADS * ppt < n) ppt *= 2;
        if (ppt > 32) continue;
        if (c == 16 && best_c != 0) continue;  // non-portable size only when nothing else fits
        // cycles per sampling step, fitted to B200 measurements (profiles/r1_fps_cluster.txt): a per-point term and the
        // fixed chain of warp reductions + record exchange, which costs ~850 cycles more through a cluster barrier
        const double waves = (double)((b * c + 147) / 148);
        const double fixed = c == 1 ? 1280.0 : (c <= 4 ? 1950.0 : (c == 8 ? 2400.0 : 2800.0));
        const double cost = ((ppt == 32 ? 70.0 : 55.0) * ppt + fixed) * waves;
        if (cost < best_cost) { best_cost = cost; best_c = c; best_ppt = ppt; }
    }
    static int forced = -1;   // I2P_FPS_CLUSTER=c: tuning override of the cluster size (measurements in profiles/)
    if (forced < 0) { const char *e = getenv("I2P_FPS_CLUSTER"); forced = e ? atoi(e) : 0; }
    if (forced > 0 && forced <= FPS_MAX_CLUSTER && (forced & (forced - 1)) == 0) {
        int ppt = 1;
        while ((long long)forced * FPS_THREADS * ppt < n) ppt *= 2;
        if (ppt <= 32) { best_c = forced; best_ppt = ppt; }
    }
    if (best_c == 0) {
        set_error("furthest_point_sampling: n=%d exceeds the on-chip capacity of a 16-CTA cluster (262144)", n);
        return I2P_ERR_UNSUPPORTED;
    }
    cudaStream_t s = as_stream(stream);
    switch (best_ppt) {
        case 1: return launch_fps<1, false>(a, b, best_c, s);
        case 2: return launch_fps<2, false>(a, b, best_c, s);
        case 4: return launch_fps<4, false>(a, b, best_c, s);
        case 8: return launch_fps<8, false>(a, b, best_c, s);
        case 16: return launch_fps<16, false>(a, b, best_c, s);
        default: return launch_fps<32, true>(a, b, best_c, s);
    }
}
