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
