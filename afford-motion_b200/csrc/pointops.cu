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

__device__ __forceinline__ void argmax_merge(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

// ---------------------------------------------------------------- farthest point sampling
// One CTA (1024 threads) per batch segment; the segment's coordinates and running min-distances stay in
// registers (PPT points per thread) for all m-1 dependent rounds — no global round trip per round.
constexpr int FPS_THREADS = 1024;

template <int PPT>
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ offset, const int32_t* __restrict__ new_offset,
           float* __restrict__ tmp, int32_t* __restrict__ idx) {
    const int s = blockIdx.x;
    const int start_n = s == 0 ? 0 : offset[s - 1], end_n = offset[s];
    const int start_m = s == 0 ? 0 : new_offset[s - 1], end_m = new_offset[s];
    const int n = end_n - start_n, m = end_m - start_m;
    if (m <= 0 || n <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ float s_val[32];
    __shared__ int s_idx[32];
    __shared__ float s_last[3];
    __shared__ int s_best;

    float px[PPT > 0 ? PPT : 1], py[PPT > 0 ? PPT : 1], pz[PPT > 0 ? PPT : 1], pt[PPT > 0 ? PPT : 1];
    if (PPT > 0) {
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            int k = tid + i * FPS_THREADS;
            if (k < n) {
                const float* p = xyz + 3 * (int64_t)(start_n + k);
                px[i] = p[0]; py[i] = p[1]; pz[i] = p[2];
            } else { px[i] = py[i] = pz[i] = 0.f; }
            pt[i] = 1e10f;  // pointops.py:22
        }
    } else {
        for (int k = tid; k < n; k += FPS_THREADS) tmp[start_n + k] = 1e10f;
    }
    if (tid == 0) {
        idx[start_m] = start_n;
        const float* p = xyz + 3 * (int64_t)start_n;
        s_last[0] = p[0]; s_last[1] = p[1]; s_last[2] = p[2];
    }
    __syncthreads();
    for (int j = 1; j < m; ++j) {
        const float lx = s_last[0], ly = s_last[1], lz = s_last[2];
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
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            argmax_merge(best, besti, ov, oi);
        }
        if (lane == 0) { s_val[warp] = best; s_idx[warp] = besti; }
        __syncthreads();
        if (warp == 0) {
            best = s_val[lane]; besti = s_idx[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float ov = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, besti, o);
                argmax_merge(best, besti, ov, oi);
            }
            if (lane == 0) {
                s_best = besti;
                idx[start_m + j] = start_n + besti;
                const float* p = xyz + 3 * (int64_t)(start_n + besti);
                s_last[0] = p[0]; s_last[1] = p[1]; s_last[2] = p[2];
            }
        }
        __syncthreads();
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
        for (int i = threadIdx.x; i < cnt; i += KNN_THREADS) {
            const float* p = xyz + 3 * (int64_t)(t0 + i);
            tile[i] = make_float4(p[0], p[1], p[2], 0.f);
        }
        __syncthreads();
        if (active) {
            for (int i = 0; i < cnt; ++i) {
                float4 c = tile[i];
                float d = sqdist_rn(c.x, c.y, c.z, qx, qy, qz);
                if (d < bd[K - 1]) {  // strict: on ties the earlier (lower) index stays
                    bd[K - 1] = d; bi[K - 1] = t0 + i;
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
    if (n_max <= 1 * FPS_THREADS) fps_kernel<1><<<b, FPS_THREADS, 0, st>>>(xyz, offset, new_offset, tmp, idx);
    else if (n_max <= 2 * FPS_THREADS) fps_kernel<2><<<b, FPS_THREADS, 0, st>>>(xyz, offset, new_offset, tmp, idx);
    else if (n_max <= 4 * FPS_THREADS) fps_kernel<4><<<b, FPS_THREADS, 0, st>>>(xyz, offset, new_offset, tmp, idx);
    else if (n_max <= 8 * FPS_THREADS) fps_kernel<8><<<b, FPS_THREADS, 0, st>>>(xyz, offset, new_offset, tmp, idx);
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
