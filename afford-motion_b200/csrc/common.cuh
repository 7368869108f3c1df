// Shared device/host helpers for the amb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../include/amb200.h"

extern "C" void am_set_error_(const char* msg);
extern "C" void am_count_launch_(int n);
extern "C" int am_get_precision(void);

#define AM_REQUIRE(cond, code, msg)            \
    do {                                       \
        if (!(cond)) {                         \
            am_set_error_(msg);                \
            return (code);                     \
        }                                      \
    } while (0)

#define AM_LAUNCH_CHECK(name)                                  \
    do {                                                       \
        am_count_launch_(1);                                   \
        cudaError_t e__ = cudaGetLastError();                  \
        if (e__ != cudaSuccess) {                              \
            am_set_error_(cudaGetErrorString(e__));            \
            return AM_ELAUNCH;                                 \
        }                                                      \
    } while (0)

static inline cudaStream_t as_stream(am_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- Programmatic dependent launch (PDL).  One denoise step is a chain of ~45 dependent kernels; measured on B200 the gap
// between two of them (grid drain + launch latency + the next kernel's prologue) is ~4.7 us per tcgen05 GEMM.  Kernels launched
// through am_launch() may START while their predecessor in the stream is still running: they execute their prologue
// (barrier init, TMEM allocation, tensor-map prefetch) and then block in pdl_wait() until the predecessor has completed and
// its memory is visible.  RULE: a kernel launched through am_launch() must call pdl_wait() before its first access to global
// memory another kernel may have written (or may still read).  AMB200_PDL=0 launches everything fully serialised.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
static inline bool am_pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("AMB200_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
    return on != 0;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t am_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (am_pdl_enabled()) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster_x > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = (unsigned)cluster_x; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = at; cfg.numAttrs = (unsigned)n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
// SM count of the current device (148 on B200), queried once per process instead of hard-coded
static inline int am_num_sms() {
    static int n = 0;
    if (n <= 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n = v;
        else return 148;
    }
    return n;
}
#define AM_NUM_SMS am_num_sms()

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float apply_act(float x, int act) {
    switch (act) {
        case AM_ACT_GELU: return gelu_erf(x);
        case AM_ACT_SILU: return silu_f(x);
        case AM_ACT_RELU: return fmaxf(x, 0.0f);
        default: return x;
    }
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based: pure function of (key, counter)
struct Philox4 { uint32_t x, y, z, w; };
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return {c0, c1, c2, c3};
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }
// 4 standard normals from one Philox block (Box-Muller)
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint32_t subseq, uint32_t sample, uint32_t blk, float out[4]) {
    Philox4 r = philox4x32_10(blk, sample, subseq, 0x414D4232u, (uint32_t)seed, (uint32_t)(seed >> 32));
    float r0 = sqrtf(-2.0f * logf(u01(r.x))), r1 = sqrtf(-2.0f * logf(u01(r.z)));
    float s0, c0, s1, c1;
    sincospif(2.0f * u01(r.y), &s0, &c0);
    sincospif(2.0f * u01(r.w), &s1, &c1);
    out[0] = r0 * c0; out[1] = r0 * s0; out[2] = r1 * c1; out[3] = r1 * s1;
}
