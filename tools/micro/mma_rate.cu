// Micro-benchmark: how long does one elected thread need per tcgen05.mma.kind::tf32 (M = 128, K = 8) as a function of
// N, of the operand layout (no-swizzle K-major with the convolution's plane pitch / SWIZZLE_128B K-major), of the
// alignment of the A start address and of how the MMAs rotate over accumulators?  One CTA per SM, operands zero.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "../../i2pnet_b200/csrc/umma.cuh"

using namespace i2p;

__host__ __device__ constexpr uint32_t idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct P { int n, reps, naccum, swz, a_shift, lbo, descs; };

template <int N>
__global__ void __launch_bounds__(128) k(P p, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x < 32) umma::tmem_alloc(&slot, 512);
    if (threadIdx.x == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tm = slot, sb = umma::smem_u32(smem);
    if (threadIdx.x == 0) {
        const uint32_t a0 = sb + p.a_shift, b0 = sb + 32 * 1024;
        long long t0 = clock64();
        for (int r = 0; r < p.reps; ++r) {
            const uint32_t acc = tm + (uint32_t)((r % p.naccum) * N);
            // descs = 1: rebuild both descriptors per MMA like the kernels do (different start addresses)
            const uint32_t ao = a0 + (p.descs ? (uint32_t)(r % 9) * 16u : 0u);
            uint64_t ad, bd;
            if (p.swz) { ad = umma::smem_desc(a0 + (r & 3) * 32, 16, 1024, 2); bd = umma::smem_desc(b0 + (r & 3) * 32, 16, 1024, 2); }
            else { ad = umma::smem_desc(ao, p.lbo, 128); bd = umma::smem_desc(b0, N * 16, 128); }
            umma::mma_tf32(acc, ad, bd, idesc(N), r >= p.naccum);
        }
        long long t1 = clock64();
        umma::commit(&bar);
        umma::mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_dealloc(tm, 512);
}

template <int N>
static void run(P p, long long *d) {
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k<N><<<148, 128, 64 * 1024>>>(p, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("N %3d  reps %4d  accumulators %d  %-9s a_shift %3d  lbo %5d : issue %6.1f cyc/MMA, issue+complete %6.1f cyc/MMA  %s\n", p.n,
           p.reps, p.naccum, p.swz ? "SW128" : "noswizzle", p.a_shift, p.lbo, (double)h[0] / p.reps, (double)h[1] / p.reps,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long *d;
    cudaMalloc(&d, 16);
    for (int reps : {54, 432}) {
        for (int na : {1, 2, 4}) {
            run<16>({16, reps, na, 0, 0, 6288, 1}, d);
            run<32>({32, reps, na, 0, 0, 6288, 1}, d);
        }
        run<16>({16, reps, 4, 0, 0, 6240, 1}, d);
        run<16>({16, reps, 4, 0, 0, 2048, 0}, d);
        run<16>({16, reps, 4, 0, 16, 2048, 0}, d);
        run<64>({64, reps, 1, 0, 0, 2048, 0}, d);
        run<128>({128, reps, 1, 0, 0, 2048, 0}, d);
        run<128>({128, reps, 2, 0, 0, 2048, 0}, d);
        run<128>({128, reps, 1, 1, 0, 0, 0}, d);
        run<16>({16, reps, 1, 1, 0, 0, 0}, d);
    }
    return 0;
}
