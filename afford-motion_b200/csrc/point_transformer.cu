// Fused eval-mode Point Transformer kernels (models/scene_models/pointtransformer.py):
//   pt_layer_kernel        : PointTransformerLayer.forward :26-38 (+ bn2/relu of the enclosing block :119)
//   transition_down_kernel : TransitionDown.forward :61-66 (gather -> linear -> BN -> ReLU -> max over k)
// The [n,k,c] grouped tensors of the reference (pointops.queryandgroup, pointops.py:79-100) are never
// materialised in HBM: neighbours are gathered straight into shared memory / registers.
#include <math_constants.h>
#include "common.cuh"

namespace {

constexpr int PT_WARPS = 8;

__global__ void __launch_bounds__(PT_WARPS * 32)
pt_layer_kernel(const float* __restrict__ p, const float* __restrict__ qkv, const int32_t* __restrict__ idx,
                const float* __restrict__ wp1, const float* __restrict__ bp1, const float* __restrict__ wp2,
                const float* __restrict__ bp2, const float* __restrict__ bnw_s, const float* __restrict__ bnw_t,
                const float* __restrict__ ww1, const float* __restrict__ bw1, const float* __restrict__ ww2,
                const float* __restrict__ bw2, const float* __restrict__ post_s, const float* __restrict__ post_t,
                float* __restrict__ out, int n, int c, int k, int ww1_in_smem) {
    extern __shared__ __align__(16) float sm[];
    const int c8 = c / 8, cs = c + 1, c8s = c8 + 1;
    const int nwarps = blockDim.x >> 5;
    float* s_ww1 = sm;                      // [c8][c+1]  (c <= 256; at c = 512 ww1 stays in global / L1, see am_pt_layer_fwd)
    float* s_ww2 = s_ww1 + (ww1_in_smem ? c8 * cs : 0);  // [c8][c8+1]
    float* s_warp = s_ww2 + c8 * c8s;
    const float* w1base = ww1_in_smem ? s_ww1 : ww1;
    const int w1s = ww1_in_smem ? cs : c;
    const int per_warp = k * cs + 2 * k * c8s + k * 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* W = s_warp + warp * per_warp;    // [k][c+1]
    float* t1 = W + k * cs;                 // [k][c8+1]
    float* t2 = t1 + k * c8s;               // [k][c8+1]
    float* hs = t2 + k * c8s;               // [k][4] : relu(linear_p.0/1) hidden (3) + neighbour index

    if (ww1_in_smem)
        for (int i = threadIdx.x; i < c8 * c; i += blockDim.x) s_ww1[(i / c) * cs + (i % c)] = ww1[i];
    for (int i = threadIdx.x; i < c8 * c8; i += blockDim.x) s_ww2[(i / c8) * c8s + (i % c8)] = ww2[i];
    __syncthreads();

    const int pt = blockIdx.x * nwarps + warp;
    if (pt >= n) return;
    const int ld = 3 * c;
    const float px = p[3 * (int64_t)pt], py = p[3 * (int64_t)pt + 1], pz = p[3 * (int64_t)pt + 2];
    const float* qrow = qkv + (int64_t)pt * ld;

    // phase A: relative position MLP hidden + attention pre-activation w = k_nbr - q + pr
    if (lane < k) {
        int nb = idx[(int64_t)pt * k + lane];
        float rx = p[3 * (int64_t)nb] - px, ry = p[3 * (int64_t)nb + 1] - py, rz = p[3 * (int64_t)nb + 2] - pz;
        float h0 = fmaxf(wp1[0] * rx + wp1[1] * ry + wp1[2] * rz + bp1[0], 0.f);
        float h1 = fmaxf(wp1[3] * rx + wp1[4] * ry + wp1[5] * rz + bp1[1], 0.f);
        float h2 = fmaxf(wp1[6] * rx + wp1[7] * ry + wp1[8] * rz + bp1[2], 0.f);
        hs[lane * 4 + 0] = h0; hs[lane * 4 + 1] = h1; hs[lane * 4 + 2] = h2; hs[lane * 4 + 3] = __int_as_float(nb);
    }
    __syncwarp();
    for (int ch = lane; ch < c; ch += 32) {
        const float w0 = wp2[ch * 3], w1 = wp2[ch * 3 + 1], w2 = wp2[ch * 3 + 2], b2 = bp2[ch];
        const float qv = qrow[ch], s = bnw_s[ch], t = bnw_t[ch];
        for (int j = 0; j < k; ++j) {
            int nb = __float_as_int(hs[j * 4 + 3]);
            float pr = w0 * hs[j * 4] + w1 * hs[j * 4 + 1] + w2 * hs[j * 4 + 2] + b2;
            float wv = qkv[(int64_t)nb * ld + c + ch] - qv + pr;
            W[j * cs + ch] = fmaxf(wv * s + t, 0.f);
        }
    }
    __syncwarp();
    // phase B: t1 = relu(ww1 w + bw1)   (BN folded into ww1 / bw1)
    for (int e = lane; e < k * c8; e += 32) {
        int j = e / c8, o = e % c8;
        float acc = bw1[o];
        const float* wr = w1base + o * w1s;
        const float* xr = W + j * cs;
        for (int ch = 0; ch < c; ++ch) acc = fmaf(wr[ch], xr[ch], acc);
        t1[j * c8s + o] = fmaxf(acc, 0.f);
    }
    __syncwarp();
    // phase C: t2 = ww2 t1 + bw2
    for (int e = lane; e < k * c8; e += 32) {
        int j = e / c8, o = e % c8;
        float acc = bw2[o];
        for (int i = 0; i < c8; ++i) acc = fmaf(s_ww2[o * c8s + i], t1[j * c8s + i], acc);
        t2[j * c8s + o] = acc;
    }
    __syncwarp();
    // phase D: softmax over the k neighbours, per shared-plane channel
    for (int o = lane; o < c8; o += 32) {
        float mx = -CUDART_INF_F;
        for (int j = 0; j < k; ++j) mx = fmaxf(mx, t2[j * c8s + o]);
        float sum = 0.f;
        for (int j = 0; j < k; ++j) { float e = expf(t2[j * c8s + o] - mx); t2[j * c8s + o] = e; sum += e; }
        float inv = 1.f / sum;
        for (int j = 0; j < k; ++j) t2[j * c8s + o] *= inv;
    }
    __syncwarp();
    // phase E: out = sum_j (v_nbr + pr) * w
    for (int ch = lane; ch < c; ch += 32) {
        const float w0 = wp2[ch * 3], w1 = wp2[ch * 3 + 1], w2 = wp2[ch * 3 + 2], b2 = bp2[ch];
        const int o = ch % c8;
        float acc = 0.f;
        for (int j = 0; j < k; ++j) {
            int nb = __float_as_int(hs[j * 4 + 3]);
            float pr = w0 * hs[j * 4] + w1 * hs[j * 4 + 1] + w2 * hs[j * 4 + 2] + b2;
            acc = fmaf(qkv[(int64_t)nb * ld + 2 * c + ch] + pr, t2[j * c8s + o], acc);
        }
        if (post_s) acc = fmaxf(acc * post_s[ch] + post_t[ch], 0.f);
        out[(int64_t)pt * c + ch] = acc;
    }
}

// ---------------------------------------------------------------- TransitionDown (stride != 1)
constexpr int TD_PTS = 8;     // output points per CTA
constexpr int TD_K = 16;      // neighbours per point (nsample of every strided stage)
constexpr int TD_ROWS = TD_PTS * TD_K;  // 128 grouped rows per CTA
constexpr int TD_OC = 32;     // output channels per pass

__global__ void __launch_bounds__(256)
transition_down_kernel(const float* __restrict__ p, const float* __restrict__ x, const float* __restrict__ new_p,
                       const int32_t* __restrict__ idx, const float* __restrict__ Wm, const float* __restrict__ shift,
                       float* __restrict__ out, int m, int cin, int cout) {
    extern __shared__ __align__(16) float sm[];
    const int kin = 3 + cin, gs = kin + 1;
    float* G = sm;                 // [TD_ROWS][kin+1] grouped features
    float* Wt = G + TD_ROWS * gs;  // [kin][TD_OC] transposed weight chunk
    const int pt0 = blockIdx.x * TD_PTS;
    // gather: cat(p[idx]-new_p, x[idx])
    for (int r = threadIdx.x >> 5; r < TD_ROWS; r += 8) {
        int pt = pt0 + r / TD_K;
        int lane = threadIdx.x & 31;
        if (pt < m) {
            int nb = idx[(int64_t)pt * TD_K + (r % TD_K)];
            if (lane < 3) G[r * gs + lane] = p[3 * (int64_t)nb + lane] - new_p[3 * (int64_t)pt + lane];
            for (int ch = lane; ch < cin; ch += 32) G[r * gs + 3 + ch] = x[(int64_t)nb * cin + ch];
        } else {
            for (int ch = lane; ch < kin; ch += 32) G[r * gs + ch] = 0.f;
        }
    }
    const int row = threadIdx.x >> 1, half = threadIdx.x & 1;  // 128 rows x 2 halves of 16 outputs
    for (int oc0 = 0; oc0 < cout; oc0 += TD_OC) {
        __syncthreads();
        for (int i = threadIdx.x; i < kin * TD_OC; i += blockDim.x) {
            int o = i / kin, kk = i % kin;  // coalesced read of W[oc0+o][kk]
            Wt[kk * TD_OC + o] = (oc0 + o) < cout ? Wm[(int64_t)(oc0 + o) * kin + kk] : 0.f;
        }
        __syncthreads();
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
        const float* g = G + row * gs;
        for (int kk = 0; kk < kin; ++kk) {
            float gv = g[kk];
            const float4* w4 = reinterpret_cast<const float4*>(Wt + kk * TD_OC + half * 16);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                float4 w = w4[v];
                acc[v * 4 + 0] = fmaf(gv, w.x, acc[v * 4 + 0]); acc[v * 4 + 1] = fmaf(gv, w.y, acc[v * 4 + 1]);
                acc[v * 4 + 2] = fmaf(gv, w.z, acc[v * 4 + 2]); acc[v * 4 + 3] = fmaf(gv, w.w, acc[v * 4 + 3]);
            }
        }
        // + shift, relu, max over the 16 neighbour rows (= 32 consecutive threads: rows differ in lane bits 1..4)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            int o = oc0 + half * 16 + i;
            float v = fmaxf(acc[i] + (o < cout ? shift[o] : 0.f), 0.f);
#pragma unroll
            for (int s = 2; s <= 16; s <<= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, s));
            int pt = pt0 + row / TD_K;
            if ((row % TD_K) == 0 && pt < m && o < cout) out[(int64_t)pt * cout + o] = v;
        }
    }
}

// ---------------------------------------------------------------- 3-NN inverse-distance interpolation (TransitionUp)
__global__ void interpolation_kernel(const float* __restrict__ feat, const int32_t* __restrict__ idx, const float* __restrict__ dist2,
                                     const float* base, float* out, int n, int c, int k) {
    const int64_t total = (int64_t)n * c;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(g / c), ch = (int)(g - (int64_t)j * c);
        float norm = 0.f;
        for (int i = 0; i < k; ++i) norm += 1.0f / (sqrtf(dist2[(int64_t)j * k + i]) + 1e-8f);
        float acc = 0.f;
        for (int i = 0; i < k; ++i) {
            float w = (1.0f / (sqrtf(dist2[(int64_t)j * k + i]) + 1e-8f)) / norm;
            acc += feat[(int64_t)idx[(int64_t)j * k + i] * c + ch] * w;
        }
        out[g] = (base ? base[g] : 0.f) + acc;
    }
}

// ---------------------------------------------------------------- per-segment mean (TransitionUp head form)
__global__ void segment_mean_kernel(const float* __restrict__ x, const int32_t* __restrict__ offset, float* __restrict__ out, int c) {
    const int s = blockIdx.x;
    const int start = s == 0 ? 0 : offset[s - 1], end = offset[s];
    const int ch = blockIdx.y * 32 + (threadIdx.x & 31), row0 = threadIdx.x >> 5, nrow = blockDim.x >> 5;
    __shared__ float part[8][33];
    float acc = 0.f;
    if (ch < c)
        for (int j = start + row0; j < end; j += nrow) acc += x[(int64_t)j * c + ch];
    part[row0][threadIdx.x & 31] = acc;
    __syncthreads();
    if (row0 == 0 && ch < c) {
        float t = 0.f;
        for (int r = 0; r < nrow; ++r) t += part[r][threadIdx.x];
        out[(int64_t)s * c + ch] = end > start ? t / (float)(end - start) : 0.f;
    }
}

}  // namespace

extern "C" int am_pt_layer_fwd(const float* p, const float* qkv, const int32_t* idx, const float* wp1, const float* bp1, const float* wp2,
                               const float* bp2, const float* bnw_s, const float* bnw_t, const float* ww1, const float* bw1,
                               const float* ww2, const float* bw2, const float* post_s, const float* post_t, float* out, int n, int c,
                               int k, am_stream_t stream) {
    AM_REQUIRE(p && qkv && idx && wp1 && bp1 && wp2 && bp2 && bnw_s && bnw_t && ww1 && bw1 && ww2 && bw2 && out, AM_EINVAL,
               "am_pt_layer_fwd: null pointer");
    AM_REQUIRE(n > 0 && c >= 8 && c % 8 == 0 && c <= 512 && k >= 1 && k <= 32, AM_EINVAL, "am_pt_layer_fwd: bad dims");
    AM_REQUIRE((post_s == nullptr) == (post_t == nullptr), AM_EINVAL, "am_pt_layer_fwd: post_s/post_t must come together");
    int c8 = c / 8;
    // c <= 256: 8 warps per CTA with ww1 staged in shared memory.  c = 512 (enc5 / dec5 of PointTransformerSeg, n = B*32 points):
    // ww1 (128 KB) is read through L1 instead and the CTA shrinks until the per-warp neighbour tiles fit.
    int nw = PT_WARPS, w1smem = 1;
    auto need = [&](int warps, int in_smem) {
        return sizeof(float) * ((size_t)(in_smem ? c8 * (c + 1) : 0) + (size_t)c8 * (c8 + 1) +
                                (size_t)warps * (k * (c + 1) + 2 * k * (c8 + 1) + k * 4));
    };
    size_t smem = need(nw, w1smem);
    if (smem > 227 * 1024) { w1smem = 0; smem = need(nw, 0); }
    while (smem > 227 * 1024 && nw > 1) { nw >>= 1; smem = need(nw, 0); }
    AM_REQUIRE(smem <= 227 * 1024, AM_EINVAL, "am_pt_layer_fwd: c*k too large for shared memory");
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        if (cudaFuncSetAttribute(pt_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            am_set_error_("am_pt_layer_fwd: smem opt-in failed");
            return AM_ELAUNCH;
        }
        attr = 227 * 1024;
    }
    pt_layer_kernel<<<cdiv(n, nw), nw * 32, smem, as_stream(stream)>>>(p, qkv, idx, wp1, bp1, wp2, bp2, bnw_s, bnw_t, ww1, bw1, ww2, bw2,
                                                                        post_s, post_t, out, n, c, k, w1smem);
    AM_LAUNCH_CHECK("pt_layer_fwd");
    return AM_OK;
}

extern "C" int am_transition_down_fwd(const float* p, const float* x, const float* new_p, const int32_t* idx, const float* W,
                                      const float* shift, float* out, int m, int cin, int cout, int k, am_stream_t stream) {
    AM_REQUIRE(p && x && new_p && idx && W && shift && out, AM_EINVAL, "am_transition_down_fwd: null pointer");
    AM_REQUIRE(m > 0 && cin > 0 && cout > 0, AM_EINVAL, "am_transition_down_fwd: bad dims");
    AM_REQUIRE(k == TD_K, AM_EINVAL, "am_transition_down_fwd: nsample must be 16 (every strided stage of the reference)");
    int kin = 3 + cin;
    size_t smem = sizeof(float) * ((size_t)TD_ROWS * (kin + 1) + (size_t)kin * TD_OC);
    AM_REQUIRE(smem <= 227 * 1024, AM_EINVAL, "am_transition_down_fwd: cin too large for shared memory");
    static size_t attr = 0;
    if (smem > 48 * 1024 && attr == 0) {
        if (cudaFuncSetAttribute(transition_down_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            am_set_error_("am_transition_down_fwd: smem opt-in failed");
            return AM_ELAUNCH;
        }
        attr = 1;
    }
    transition_down_kernel<<<cdiv(m, TD_PTS), 256, smem, as_stream(stream)>>>(p, x, new_p, idx, W, shift, out, m, cin, cout);
    AM_LAUNCH_CHECK("transition_down_fwd");
    return AM_OK;
}

extern "C" int am_interpolation(const float* feat, const int32_t* idx, const float* dist2, const float* base, float* out, int n, int c,
                                int k, am_stream_t stream) {
    AM_REQUIRE(feat && idx && dist2 && out && n > 0 && c > 0 && k >= 1 && k <= 16, AM_EINVAL, "am_interpolation: bad args");
    int64_t blocks = ((int64_t)n * c + 255) / 256;
    int grid = (int)(blocks < (int64_t)AM_NUM_SMS * 8 ? blocks : (int64_t)AM_NUM_SMS * 8);
    interpolation_kernel<<<grid, 256, 0, as_stream(stream)>>>(feat, idx, dist2, base, out, n, c, k);
    AM_LAUNCH_CHECK("interpolation");
    return AM_OK;
}

extern "C" int am_segment_mean(const float* x, const int32_t* offset, float* out, int b, int c, am_stream_t stream) {
    AM_REQUIRE(x && offset && out && b > 0 && c > 0, AM_EINVAL, "am_segment_mean: bad args");
    segment_mean_kernel<<<dim3(b, cdiv(c, 32)), 256, 0, as_stream(stream)>>>(x, offset, out, c);
    AM_LAUNCH_CHECK("segment_mean");
    return AM_OK;
}
