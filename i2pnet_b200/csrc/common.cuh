// Shared helpers for the sm_100a kernels of libi2p_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/i2p_b200.h"

namespace i2p {

// Squared length with the exact contraction nvcc -O2 gives every distance site of the
// reference (SASS: FMUL dy*dy; FFMA dx*dx+t; FFMA dz*dz+t -- SURVEY.md section 2.2).
// Spelled with intrinsics so the compiler cannot re-associate or re-contract it.
__device__ __forceinline__ float sqlen(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

constexpr unsigned FULL = 0xffffffffu;

extern std::atomic<uint64_t> g_launches;
void set_error(const char *fmt, ...);
int check_launch(const char *what);

// Dynamic shared memory a weight-gradient launch asks for at least (I2P_WGRAD_SMEM bytes, 0 = only what it needs): the
// weight-gradient kernels run on low-priority side streams beside the step's serial chain; asking for more than half an
// SM's shared memory keeps them at one CTA per SM, so that a chain kernel always finds room (runtime.cu).
int wgrad_smem_floor();
template <typename K>
inline int wgrad_smem(K kernel, int needed) {
    const int floor_ = wgrad_smem_floor();
    if (floor_ <= needed) return needed;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, floor_);
    return floor_;
}

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

#define I2P_REQUIRE(cond, ...)               \
    do {                                     \
        if (!(cond)) {                       \
            i2p::set_error(__VA_ARGS__);     \
            return I2P_ERR_INVALID_ARGUMENT; \
        }                                    \
    } while (0)

}  // namespace i2p
