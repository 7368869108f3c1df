// Library-global bookkeeping: version, device check, launch counter, last error string.
#include <atomic>
#include <string.h>
#include "common.cuh"

static std::atomic<int64_t> g_launches{0};
static thread_local char g_err[256] = "";

extern "C" void am_set_error_(const char* msg) { strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1); }
extern "C" void am_count_launch_(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" int am_version(void) { return 100; }
extern "C" int64_t am_launch_count(void) { return g_launches.load(); }
extern "C" const char* am_last_error(void) { return g_err; }
extern "C" int am_check_device(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { am_set_error_("no CUDA device"); return AM_EARCH; }
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) { am_set_error_("amb200 kernels are compiled for sm_100a only"); return AM_EARCH; }
    return AM_OK;
}

// Precision mode of the tensor-core kernels (am_linear_tc, am_mha_tc_fwd): 0 = parity (3-term bf16 split, fp32-equivalent, the
// default and the mode every parity number is quoted in), 1 = fast (single bf16 pass: hi x hi only, ~1e-2 relative error).
static std::atomic<int> g_precision{0};
extern "C" int am_set_precision(int mode) {
    if (mode != 0 && mode != 1) { am_set_error_("am_set_precision: mode must be 0 (parity) or 1 (fast)"); return AM_EINVAL; }
    g_precision.store(mode);
    return AM_OK;
}
extern "C" int am_get_precision(void) { return g_precision.load(); }
