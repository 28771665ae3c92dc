// Micro-benchmark 2: does the ~155-cycle cost per tcgen05.mma (kind::tf32, M = 128, K = 8) belong to the issuing
// thread, to the CTA, or to the SM?  W issuing warps per CTA (each its own accumulator columns), C CTAs per SM;
// and the same for kind::f16 (bf16 operands, K = 16) and for M = 64.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../i2pnet_b200/csrc/umma.cuh"

using namespace i2p;

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

struct P { int reps, warps, kind, m; };

template <int N>
__global__ void __launch_bounds__(128) k(P p, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 40 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x < 32) umma::tmem_alloc(&slot, 128);
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) umma::mbar_init(&bar[i], 1); umma::mbar_fence_init(); }
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tm = slot, sb = umma::smem_u32(smem);
    const int warp = threadIdx.x >> 5;
    // idesc: c_format f32 (1<<4); a/b format: tf32 = 2, bf16 = 1 at bits 7 and 10; N>>3 at 17; M>>4 at 24
    const uint32_t fmt = p.kind == 0 ? 2u : 1u;
    const uint32_t id = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(p.m >> 4) << 24);
    if ((threadIdx.x & 31) == 0 && warp < p.warps) {
        const uint32_t a0 = sb, b0 = sb + 32 * 1024;
        const uint64_t ad = umma::smem_desc(a0, 2048, 128), bd = umma::smem_desc(b0, N * 16, 128);
        const uint32_t acc = tm + (uint32_t)(warp * N);
        long long t0 = clock64();
        for (int r = 0; r < p.reps; ++r) {
            if (p.kind == 0) umma::mma_tf32(acc, ad, bd, id, r > 0);
            else mma_f16(acc, ad, bd, id, r > 0);
        }
        long long t1 = clock64();
        umma::commit(&bar[warp]);
        umma::mbar_wait(&bar[warp], 0);
        long long t2 = clock64();
        if (blockIdx.x == 0 && warp == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_dealloc(tm, 128);
}

template <int N>
static void run(P p, int ctas_per_sm, long long *d) {
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    k<N><<<148 * ctas_per_sm, 128, 48 * 1024>>>(p, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%s M %3d N %3d  reps %4d  issuing warps/CTA %d  CTAs/SM %d : per issuer: issue %6.1f, issue+complete %6.1f cyc/MMA -> SM aggregate %6.1f cyc/MMA  %s\n",
           p.kind == 0 ? "tf32" : "bf16", p.m, N, p.reps, p.warps, ctas_per_sm, (double)h[0] / p.reps, (double)h[1] / p.reps,
           (double)h[1] / p.reps / (p.warps * ctas_per_sm), e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long *d;
    cudaMalloc(&d, 16);
    const int reps = 256;
    for (int kind : {0, 1}) {
        for (int w : {1, 2, 4}) run<16>({reps, w, kind, 128}, 1, d);
        for (int c : {2, 4}) run<16>({reps, 1, kind, 128}, c, d);
        run<16>({reps, 2, kind, 128}, 2, d);
        run<32>({reps, 1, kind, 128}, 1, d);
        run<32>({reps, 4, kind, 128}, 1, d);
        run<16>({reps, 1, kind, 64}, 1, d);
        run<16>({reps, 4, kind, 64}, 1, d);
    }
    return 0;
}
