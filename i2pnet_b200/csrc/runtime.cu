// Error state, launch accounting and ABI version of libi2p_b200.so.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace i2p {

std::atomic<uint64_t> g_launches{0};
static thread_local char t_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

// Called right after a kernel launch: counts it and converts a launch failure into an
// error code (the reference prints and calls exit(-1) here).
int check_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(err));
        return I2P_ERR_CUDA;
    }
    return I2P_OK;
}

int wgrad_smem_floor() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("I2P_WGRAD_SMEM"); v = e ? atoi(e) : 0; if (v < 0) v = 0; if (v > 227 * 1024) v = 227 * 1024; }
    return v;
}

}  // namespace i2p

extern "C" {
const char *i2p_last_error(void) { return i2p::t_error; }
int i2p_abi_version(void) { return 1; }
uint64_t i2p_launch_count(void) { return i2p::g_launches.load(std::memory_order_relaxed); }
}
