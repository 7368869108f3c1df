// Native replacement of the third-party `pointops_cuda` entry points the reference calls
// (models/scene_models/pointops.py:23 furthestsampling_cuda, :42 knnquery_cuda).
// Bit-exact contract shared with oracle/pointops_ref.c: d2 = (dx*dx + dy*dy) + dz*dz in fp32 with no
// FMA contraction (__fmul_rn/__fadd_rn), ties resolved lowest-index-first.
#include <math_constants.h>
#include "common.cuh"

namespace {

__device__ __forceinline__ float sqdist_rn(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    float s = __fmul_rn(dx, dx);
    s = __fadd_rn(s, __fmul_rn(dy, dy));
    s = __fadd_rn(s, __fmul_rn(dz, dz));
    return s;
}

// ---------------------------------------------------------------- farthest point sampling
// One CTA (1024 threads) per batch segment; the segment's coordinates and running min-distances stay in registers (PPT points per
// thread) for all m-1 dependent rounds.  Round 2: the m-1 rounds are a latency chain, and each one paid two block barriers plus a
// GLOBAL read of the winner's coordinates (~1 us per round, 2.15 ms for 8192 -> 2048).  Now the coordinates are also staged in
// shared memory (<= 8192 points: 96 KB), every warp redundantly reduces the 32 per-warp candidates (double-buffered by round parity)
// with redux.sync on the distance bit patterns (the round is issue-bound: 32 warps x ~180 instructions before, ~110 now), so a round
// is ONE barrier and no global access on the critical path.  Same arithmetic and tie rule: bit-exact vs the oracle.
constexpr int FPS_THREADS = 1024;

template <int PPT>
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ offset, const int32_t* __restrict__ new_offset,
           float* __restrict__ tmp, int32_t* __restrict__ idx) {
    extern __shared__ float s_xyz[];  // PPT > 0: [n][3] coordinates of the segment
    const int s = blockIdx.x;
    const int start_n = s == 0 ? 0 : offset[s - 1], end_n = offset[s];
    const int start_m = s == 0 ? 0 : new_offset[s - 1], end_m = new_offset[s];
    const int n = end_n - start_n, m = end_m - start_m;
    if (m <= 0 || n <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ unsigned s_key[2][32];
    __shared__ int s_idx[2][32];

    float px[PPT > 0 ? PPT : 1], py[PPT > 0 ? PPT : 1], pz[PPT > 0 ? PPT : 1], pt[PPT > 0 ? PPT : 1];
    if (PPT > 0) {
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            int k = tid + i * FPS_THREADS;
            if (k < n) {
                const float* p = xyz + 3 * (int64_t)(start_n + k);
                px[i] = p[0]; py[i] = p[1]; pz[i] = p[2];
                s_xyz[3 * k] = px[i]; s_xyz[3 * k + 1] = py[i]; s_xyz[3 * k + 2] = pz[i];
            } else { px[i] = py[i] = pz[i] = 0.f; }
            pt[i] = 1e10f;  // pointops.py:22
        }
    } else {
        for (int k = tid; k < n; k += FPS_THREADS) tmp[start_n + k] = 1e10f;
    }
    if (tid == 0) idx[start_m] = start_n;
    __syncthreads();
    int last = 0;  // local index of the most recently selected point (known to every thread)
    for (int j = 1; j < m; ++j) {
        float lx, ly, lz;
        if (PPT > 0) { lx = s_xyz[3 * last]; ly = s_xyz[3 * last + 1]; lz = s_xyz[3 * last + 2]; }
        else { const float* p = xyz + 3 * (int64_t)(start_n + last); lx = p[0]; ly = p[1]; lz = p[2]; }
        float best = -1.0f;
        int besti = 0x7fffffff;
        if (PPT > 0) {
#pragma unroll
            for (int i = 0; i < PPT; ++i) {
                int k = tid + i * FPS_THREADS;
                if (k < n) {
                    float d = sqdist_rn(px[i], py[i], pz[i], lx, ly, lz);
                    float t = fminf(pt[i], d);
                    pt[i] = t;
                    if (t > best) { best = t; besti = k; }
                }
            }
        } else {
            for (int k = tid; k < n; k += FPS_THREADS) {
                const float* p = xyz + 3 * (int64_t)(start_n + k);
                float d = sqdist_rn(p[0], p[1], p[2], lx, ly, lz);
                float t = fminf(tmp[start_n + k], d);
                tmp[start_n + k] = t;
                if (t > best) { best = t; besti = k; }
            }
        }
        // arg-max with ties to the lowest index as two redux.sync per level: distances are >= 0, so their bit patterns order like
        // unsigned integers; key = bits + 1 keeps 0 for "no candidate" (a thread past the end of the segment)
        unsigned key = besti == 0x7fffffff ? 0u : __float_as_uint(best) + 1u;
        unsigned kmax = __reduce_max_sync(0xffffffffu, key);
        int imin = __reduce_min_sync(0xffffffffu, key == kmax ? besti : 0x7fffffff);
        const int par = j & 1;
        if (lane == 0) { s_key[par][warp] = kmax; s_idx[par][warp] = imin; }
        __syncthreads();  // the only barrier of the round: the buffer of parity `par` is rewritten two rounds later, after the next barrier
        key = s_key[par][lane];
        besti = s_idx[par][lane];
        kmax = __reduce_max_sync(0xffffffffu, key);
        last = __reduce_min_sync(0xffffffffu, key == kmax ? besti : 0x7fffffff);
        if (tid == 0) idx[start_m + j] = start_n + last;
    }
}

// ---------------------------------------------------------------- brute-force kNN inside a segment
// One thread per query, sorted top-K in registers, candidates streamed through shared memory tiles.
constexpr int KNN_THREADS = 256;
constexpr int KNN_TILE = 1024;

template <int K>
__global__ void __launch_bounds__(KNN_THREADS)
knn_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, const int32_t* __restrict__ offset,
           const int32_t* __restrict__ new_offset, int32_t* __restrict__ idx, float* __restrict__ dist2, int k_out) {
    const int s = blockIdx.y;
    const int start_n = s == 0 ? 0 : offset[s - 1], end_n = offset[s];
    const int start_m = s == 0 ? 0 : new_offset[s - 1], end_m = new_offset[s];
    const int q = start_m + blockIdx.x * KNN_THREADS + threadIdx.x;
    if (start_m + blockIdx.x * KNN_THREADS >= end_m) return;  // whole CTA out of range
    const bool active = q < end_m;
    __shared__ float4 tile[KNN_TILE];
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) { const float* p = new_xyz + 3 * (int64_t)q; qx = p[0]; qy = p[1]; qz = p[2]; }
    float bd[K];
    int bi[K];
#pragma unroll
    for (int i = 0; i < K; ++i) { bd[i] = CUDART_INF_F; bi[i] = 0; }
    for (int t0 = start_n; t0 < end_n; t0 += KNN_TILE) {
        int cnt = min(KNN_TILE, end_n - t0);
        __syncthreads();
        const int cnt4 = (cnt + 3) & ~3;
        for (int i = threadIdx.x; i < cnt4; i += KNN_THREADS) {
            if (i < cnt) {
                const float* p = xyz + 3 * (int64_t)(t0 + i);
                tile[i] = make_float4(p[0], p[1], p[2], 0.f);
            } else {
                tile[i] = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f);  // sentinel: distance +inf, never inserted
            }
        }
        __syncthreads();
        if (active) {
            // candidates in groups of 4: one test of the group minimum against the current k-th distance replaces four compare +
            // branch pairs (after the first few hundred candidates almost every group is rejected); an accepted group replays its
            // members in index order with the same strict test, so results are identical to the one-at-a-time loop
            for (int i = 0; i < cnt4; i += 4) {
                const float4 c0 = tile[i], c1 = tile[i + 1], c2 = tile[i + 2], c3 = tile[i + 3];
                const float d0 = sqdist_rn(c0.x, c0.y, c0.z, qx, qy, qz), d1 = sqdist_rn(c1.x, c1.y, c1.z, qx, qy, qz);
                const float d2 = sqdist_rn(c2.x, c2.y, c2.z, qx, qy, qz), d3 = sqdist_rn(c3.x, c3.y, c3.z, qx, qy, qz);
                if (fminf(fminf(d0, d1), fminf(d2, d3)) < bd[K - 1]) {
                    const float dd[4] = {d0, d1, d2, d3};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float d = dd[u];
                        if (d < bd[K - 1]) {  // strict: on ties the earlier (lower) index stays
                            bd[K - 1] = d; bi[K - 1] = t0 + i + u;
#pragma unroll
                            for (int p = K - 1; p > 0; --p) {
                                if (bd[p - 1] > bd[p]) {
                                    float td = bd[p]; bd[p] = bd[p - 1]; bd[p - 1] = td;
                                    int ti = bi[p]; bi[p] = bi[p - 1]; bi[p - 1] = ti;
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    if (active) {
        int have = min(k_out, end_n - start_n);
#pragma unroll
        for (int i = 0; i < K; ++i) {
            // fewer candidates than k: slots keep the caller's zero fill (pointops.py:40-41)
            if (i < k_out) {
                idx[(int64_t)q * k_out + i] = i < have ? bi[i] : 0;
                dist2[(int64_t)q * k_out + i] = i < have ? bd[i] : 0.f;
            }
        }
    }
}

}  // namespace

extern "C" int am_furthestsampling(int b, int n_max, const float* xyz, const int32_t* offset, const int32_t* new_offset, float* tmp,
                                   int32_t* idx, am_stream_t stream) {
    AM_REQUIRE(b > 0 && n_max > 0 && xyz && offset && new_offset && idx, AM_EINVAL, "am_furthestsampling: bad args");
    cudaStream_t st = as_stream(stream);
    // PPT > 0: the segment's coordinates are staged in shared memory (12 B per point, 96 KB for 8192 points)
    const size_t sm1 = 12u * 1 * FPS_THREADS, sm2 = 12u * 2 * FPS_THREADS, sm4 = 12u * 4 * FPS_THREADS, sm8 = 12u * 8 * FPS_THREADS;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(fps_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm4) != cudaSuccess ||
            cudaFuncSetAttribute(fps_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm8) != cudaSuccess) {
            am_set_error_("am_furthestsampling: shared memory opt-in failed");
            return AM_ELAUNCH;
        }
        attr = true;
    }
    if (n_max <= 1 * FPS_THREADS) fps_kernel<1><<<b, FPS_THREADS, sm1, st>>>(xyz, offset, new_offset, tmp, idx);
    else if (n_max <= 2 * FPS_THREADS) fps_kernel<2><<<b, FPS_THREADS, sm2, st>>>(xyz, offset, new_offset, tmp, idx);
    else if (n_max <= 4 * FPS_THREADS) fps_kernel<4><<<b, FPS_THREADS, sm4, st>>>(xyz, offset, new_offset, tmp, idx);
    else if (n_max <= 8 * FPS_THREADS) fps_kernel<8><<<b, FPS_THREADS, sm8, st>>>(xyz, offset, new_offset, tmp, idx);
    else {
        AM_REQUIRE(tmp, AM_EINVAL, "am_furthestsampling: tmp workspace required for segments > 8192 points");
        fps_kernel<0><<<b, FPS_THREADS, 0, st>>>(xyz, offset, new_offset, tmp, idx);
    }
    AM_LAUNCH_CHECK("furthestsampling");
    return AM_OK;
}

extern "C" int am_knnquery(int b, int m, int nsample, const float* xyz, const float* new_xyz, const int32_t* offset,
                           const int32_t* new_offset, int32_t* idx, float* dist2, am_stream_t stream) {
    AM_REQUIRE(b > 0 && m > 0 && xyz && new_xyz && offset && new_offset && idx && dist2, AM_EINVAL, "am_knnquery: bad args");
    AM_REQUIRE(nsample >= 1 && nsample <= 16, AM_EINVAL, "am_knnquery: nsample must be in [1,16]");
    // grid.x covers the largest segment's queries; m is an upper bound (segments are not known on the host)
    dim3 grid(cdiv(m, KNN_THREADS), b);
    cudaStream_t st = as_stream(stream);
    if (nsample <= 4) knn_kernel<4><<<grid, KNN_THREADS, 0, st>>>(xyz, new_xyz, offset, new_offset, idx, dist2, nsample);
    else if (nsample <= 8) knn_kernel<8><<<grid, KNN_THREADS, 0, st>>>(xyz, new_xyz, offset, new_offset, idx, dist2, nsample);
    else knn_kernel<16><<<grid, KNN_THREADS, 0, st>>>(xyz, new_xyz, offset, new_offset, idx, dist2, nsample);
    AM_LAUNCH_CHECK("knnquery");
    return AM_OK;
}
