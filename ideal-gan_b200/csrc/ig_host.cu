// Host-side plumbing of libidealgan: version, thread-local error string, device check.
#include <stdarg.h>
#include <stdio.h>

#include "ig_common.cuh"

namespace ig {
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
    return static_cast<int>(e);
}
}  // namespace ig

extern "C" int ig_version(void) { return IG_VERSION; }
extern "C" const char *ig_last_error(void) { return ig::g_err; }

extern "C" int ig_device_ok(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}

extern "C" size_t ig_loss_scratch_bytes(int nb, int nv) {
    if (nb <= 0 || nv <= 0) return 0;
    // one float per block of the widest launch shape (one voxel per thread), plus the ticket header
    const size_t blocks_per_sample = (static_cast<size_t>(nv) + ig::kThreads - 1) / ig::kThreads;
    return ig::kScratchHeader + sizeof(float) * blocks_per_sample * static_cast<size_t>(nb);
}
