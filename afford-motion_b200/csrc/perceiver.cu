// CDM ContactPerceiver point-side kernels (models/cdm.py:155-188; Perceiver-IO blocks models/modules.py:222-661).
// With only 2 latent tokens the encoder/decoder cross-attentions are folded exactly:
//   score = (W_k^T q).LN(e) + q.b_k     and     sum_j p_j (W_v LN(e_j) + b_v) = W_v (sum_j p_j LN(e_j)) + b_v
// so K/V [B,N,512] are never formed; each point costs O(16*256) instead of O(256*1024) MACs.
// One warp per point, 8 channels per lane (C = 256), fp32 throughout, flash-style (max,sum,acc) partials.
#include <cuda_bf16.h>
#include <math_constants.h>
#include "common.cuh"

namespace {

constexpr int C = 256;     // encoder_kv_input_channels == decoder_q_input_channels (configs/model/cdm.yaml:38-46)
constexpr int CL = C / 32; // channels per lane
constexpr int R = 16;      // (head, latent) rows: 8 heads x 2 latents
constexpr int MAXCIN = 64;
constexpr int PW = 8;      // warps per CTA

__device__ __forceinline__ void layernorm8(float e[CL], const float* __restrict__ g, const float* __restrict__ b, int lane) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CL; ++i) s += e[i];
    float mean = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < CL; ++i) { float d = e[i] - mean; q += d * d; }
    float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + 1e-5f);
#pragma unroll
    for (int i = 0; i < CL; ++i) e[i] = (e[i] - mean) * rstd * g[lane * CL + i] + b[lane * CL + i];
}

// 16 dot products of the lane-distributed vector v with rows of M [R][ldm] (shared memory), result in every lane
__device__ __forceinline__ void dots16(const float* __restrict__ M, int ldm, const float v[CL], int lane, float s[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float4* m4 = reinterpret_cast<const float4*>(M + r * ldm + lane * CL);
        float4 a = m4[0], b = m4[1];
        float t = a.x * v[0] + a.y * v[1] + a.z * v[2] + a.w * v[3] + b.x * v[4] + b.y * v[5] + b.z * v[6] + b.w * v[7];
        s[r] = warp_sum(t);
    }
}

// part layout: [B][nchunk][R][C + 2]  (acc[C], max, sum)
__global__ void __launch_bounds__(PW * 32, 1)
cdm_encoder_partial_kernel(const float* __restrict__ x_t, const float* __restrict__ xyz, const float* __restrict__ w_ea,
                           const float* __restrict__ b_ea, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                           const float* __restrict__ qf, int ldq, float* __restrict__ part, int N, int cx, int nchunk) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float sm[];
    const int cin = cx + 3;
    float* Wt = sm;                 // [cin][C]
    float* Q = Wt + cin * C;        // [R][C + 4]: folded queries, column C = bias term q.b_k
    float* red = Q + R * (C + 4);   // [PW][R][C + 2] per-warp partials for the CTA combine
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < cin * C; i += blockDim.x) { int ch = i / cin, k = i % cin; Wt[k * C + ch] = w_ea[i]; }
    for (int i = tid; i < R * (C + 4); i += blockDim.x) { int r = i / (C + 4), k = i % (C + 4); Q[i] = k <= C ? qf[((int64_t)b * R + r) * ldq + k] : 0.f; }
    __syncthreads();

    const int per = (N + nchunk - 1) / nchunk;
    const int j0 = chunk * per, j1 = min(N, j0 + per);
    float m[R], l[R], acc[R][CL];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        m[r] = -CUDART_INF_F; l[r] = 0.f;
#pragma unroll
        for (int i = 0; i < CL; ++i) acc[r][i] = 0.f;
    }
    for (int j = j0 + warp; j < j1; j += PW) {
        const float* xp = x_t + ((int64_t)b * N + j) * cx;
        const float* pp = xyz + ((int64_t)b * N + j) * 3;
        float e[CL];
        {   // cat(x_t, xyz): models/cdm.py:167-171
            const float4* b4 = reinterpret_cast<const float4*>(b_ea + lane * CL);
            float4 b0 = b4[0], b1 = b4[1];
            e[0] = b0.x; e[1] = b0.y; e[2] = b0.z; e[3] = b0.w; e[4] = b1.x; e[5] = b1.y; e[6] = b1.z; e[7] = b1.w;
            for (int i = 0; i < cin; ++i) {
                float ui = i < cx ? xp[i] : pp[i - cx];
                const float4* w4 = reinterpret_cast<const float4*>(Wt + i * C + lane * CL);
                float4 w0 = w4[0], w1 = w4[1];
                e[0] = fmaf(w0.x, ui, e[0]); e[1] = fmaf(w0.y, ui, e[1]); e[2] = fmaf(w0.z, ui, e[2]); e[3] = fmaf(w0.w, ui, e[3]);
                e[4] = fmaf(w1.x, ui, e[4]); e[5] = fmaf(w1.y, ui, e[5]); e[6] = fmaf(w1.z, ui, e[6]); e[7] = fmaf(w1.w, ui, e[7]);
            }
        }
        layernorm8(e, ln_g, ln_b, lane);
        float s[R];
        dots16(Q, C + 4, e, lane, s);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float sr = s[r] + Q[r * (C + 4) + C];
            if (sr > m[r]) {  // warp-uniform: rescale the running partial
                float f = expf(m[r] - sr);
                l[r] *= f;
#pragma unroll
                for (int i = 0; i < CL; ++i) acc[r][i] *= f;
                m[r] = sr;
            }
            float pj = expf(sr - m[r]);
            l[r] += pj;
#pragma unroll
            for (int i = 0; i < CL; ++i) acc[r][i] = fmaf(pj, e[i], acc[r][i]);
        }
    }
    // CTA combine of the 8 warp partials
    float* mine = red + (int64_t)warp * R * (C + 2);
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int i = 0; i < CL; ++i) mine[r * (C + 2) + lane * CL + i] = acc[r][i];
        if (lane == 0) { mine[r * (C + 2) + C] = m[r]; mine[r * (C + 2) + C + 1] = l[r]; }
    }
    __syncthreads();
    float* outp = part + ((int64_t)b * nchunk + chunk) * R * (C + 2);
    for (int i = tid; i < R * C; i += blockDim.x) {
        int r = i / C, ch = i % C;
        float M = -CUDART_INF_F;
        for (int w = 0; w < PW; ++w) M = fmaxf(M, red[((int64_t)w * R + r) * (C + 2) + C]);
        float a = 0.f, ls = 0.f;
        for (int w = 0; w < PW; ++w) {
            float mw = red[((int64_t)w * R + r) * (C + 2) + C];
            float f = mw == -CUDART_INF_F ? 0.f : expf(mw - M);
            a += f * red[((int64_t)w * R + r) * (C + 2) + ch];
            ls += f * red[((int64_t)w * R + r) * (C + 2) + C + 1];
        }
        outp[r * (C + 2) + ch] = a;
        if (ch == 0) { outp[r * (C + 2) + C] = M; outp[r * (C + 2) + C + 1] = ls; }
    }
}

__global__ void cdm_encoder_combine_kernel(const float* __restrict__ part, float* __restrict__ z, int nchunk) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y, r = blockIdx.x, ch = threadIdx.x;
    const float* base = part + (int64_t)b * nchunk * R * (C + 2);
    float M = -CUDART_INF_F;
    for (int c = 0; c < nchunk; ++c) M = fmaxf(M, base[((int64_t)c * R + r) * (C + 2) + C]);
    float a = 0.f, ls = 0.f;
    for (int c = 0; c < nchunk; ++c) {
        float mc = base[((int64_t)c * R + r) * (C + 2) + C];
        float f = mc == -CUDART_INF_F ? 0.f : expf(mc - M);
        a += f * base[((int64_t)c * R + r) * (C + 2) + ch];
        ls += f * base[((int64_t)c * R + r) * (C + 2) + C + 1];
    }
    z[((int64_t)b * R + r) * C + ch] = a / ls;
}

// decoder per-point stage: dq (folded adapters) -> LN_q -> 16 folded scores -> per-head softmax over 2 latents
// -> h1 = dq + sum_r p_r U_r + bo ; hn = LN_m(h1)
__global__ void __launch_bounds__(PW * 32)
cdm_decoder_point_kernel(const float* __restrict__ x_t, const float* __restrict__ xyz, const float* __restrict__ wd,
                         const float* __restrict__ bd, const float* __restrict__ lnq_g, const float* __restrict__ lnq_b,
                         const float* __restrict__ kf, int ldk, const float* __restrict__ U, const float* __restrict__ bo,
                         const float* __restrict__ lnm_g, const float* __restrict__ lnm_b, float* __restrict__ h1,
                         float* __restrict__ hn, __nv_bfloat16* __restrict__ hn2, int N, int cx, int pts_per_cta) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float sm[];
    const int cin = cx + 3;
    float* Wt = sm;                  // [cin][C]
    float* Kf = Wt + cin * C;        // [R][C+4] folded keys (+ bias column)
    float* Us = Kf + R * (C + 4);    // [R][C]
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < cin * C; i += blockDim.x) { int ch = i / cin, k = i % cin; Wt[k * C + ch] = wd[i]; }
    for (int i = tid; i < R * (C + 4); i += blockDim.x) { int r = i / (C + 4), k = i % (C + 4); Kf[i] = k <= C ? kf[((int64_t)b * R + r) * ldk + k] : 0.f; }
    for (int i = tid; i < R * C; i += blockDim.x) Us[i] = U[(int64_t)b * R * C + i];
    __syncthreads();
    const int j0 = blockIdx.x * pts_per_cta, j1 = min(N, j0 + pts_per_cta);
    for (int j = j0 + warp; j < j1; j += PW) {
        const float* xp = x_t + ((int64_t)b * N + j) * cx;
        const float* pp = xyz + ((int64_t)b * N + j) * 3;
        float dq[CL], qn[CL];
        {
            const float4* b4 = reinterpret_cast<const float4*>(bd + lane * CL);
            float4 b0 = b4[0], b1 = b4[1];
            dq[0] = b0.x; dq[1] = b0.y; dq[2] = b0.z; dq[3] = b0.w; dq[4] = b1.x; dq[5] = b1.y; dq[6] = b1.z; dq[7] = b1.w;
            for (int i = 0; i < cin; ++i) {
                float ui = i < cx ? xp[i] : pp[i - cx];
                const float4* w4 = reinterpret_cast<const float4*>(Wt + i * C + lane * CL);
                float4 w0 = w4[0], w1 = w4[1];
                dq[0] = fmaf(w0.x, ui, dq[0]); dq[1] = fmaf(w0.y, ui, dq[1]); dq[2] = fmaf(w0.z, ui, dq[2]); dq[3] = fmaf(w0.w, ui, dq[3]);
                dq[4] = fmaf(w1.x, ui, dq[4]); dq[5] = fmaf(w1.y, ui, dq[5]); dq[6] = fmaf(w1.z, ui, dq[6]); dq[7] = fmaf(w1.w, ui, dq[7]);
            }
        }
#pragma unroll
        for (int i = 0; i < CL; ++i) qn[i] = dq[i];
        layernorm8(qn, lnq_g, lnq_b, lane);
        float s[R];
        dots16(Kf, C + 4, qn, lane, s);
        float o[CL];
#pragma unroll
        for (int i = 0; i < CL; ++i) o[i] = dq[i] + bo[lane * CL + i];
#pragma unroll
        for (int h = 0; h < R / 2; ++h) {  // rows r = 2h (latent 0), 2h+1 (latent 1)
            float s0 = s[2 * h] + Kf[(2 * h) * (C + 4) + C], s1 = s[2 * h + 1] + Kf[(2 * h + 1) * (C + 4) + C];
            float mx = fmaxf(s0, s1);
            float e0 = expf(s0 - mx), e1 = expf(s1 - mx);
            float inv = 1.f / (e0 + e1);
            float p0 = e0 * inv, p1 = e1 * inv;
            const float4* u0 = reinterpret_cast<const float4*>(Us + (2 * h) * C + lane * CL);
            const float4* u1 = reinterpret_cast<const float4*>(Us + (2 * h + 1) * C + lane * CL);
            float4 a0 = u0[0], a1 = u0[1], c0 = u1[0], c1 = u1[1];
            o[0] += p0 * a0.x + p1 * c0.x; o[1] += p0 * a0.y + p1 * c0.y; o[2] += p0 * a0.z + p1 * c0.z; o[3] += p0 * a0.w + p1 * c0.w;
            o[4] += p0 * a1.x + p1 * c1.x; o[5] += p0 * a1.y + p1 * c1.y; o[6] += p0 * a1.z + p1 * c1.z; o[7] += p0 * a1.w + p1 * c1.w;
        }
        float* h1p = h1 + ((int64_t)b * N + j) * C + lane * CL;
        reinterpret_cast<float4*>(h1p)[0] = make_float4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<float4*>(h1p)[1] = make_float4(o[4], o[5], o[6], o[7]);
        layernorm8(o, lnm_g, lnm_b, lane);
        if (hn) {
            float* hnp = hn + ((int64_t)b * N + j) * C + lane * CL;
            reinterpret_cast<float4*>(hnp)[0] = make_float4(o[0], o[1], o[2], o[3]);
            reinterpret_cast<float4*>(hnp)[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
        if (hn2) {  // bf16 (hi | lo) operand of the tcgen05 MLP GEMM, row stride 2*C
            uint32_t wh[4], wl[4];
#pragma unroll
            for (int i = 0; i < CL; i += 2) {
                __nv_bfloat16 h0 = __float2bfloat16_rn(o[i]), h1 = __float2bfloat16_rn(o[i + 1]);
                __nv_bfloat16 l0 = __float2bfloat16_rn(o[i] - __bfloat162float(h0)), l1 = __float2bfloat16_rn(o[i + 1] - __bfloat162float(h1));
                wh[i / 2] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                wl[i / 2] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
            __nv_bfloat16* base2 = hn2 + ((int64_t)b * N + j) * (2 * C) + lane * CL;
            *reinterpret_cast<uint4*>(base2) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
            *reinterpret_cast<uint4*>(base2 + C) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
        }
    }
}

// Y[M, N<=8] = X1 W[:, :K1]^T + X2 W[:, K1:]^T + bias : warp per row, lanes over K (coalesced), shuffle reduce
__global__ void __launch_bounds__(256)
linear_skinny_kernel(const float* __restrict__ X1, int ldx1, int K1, const float* __restrict__ X2, int ldx2, int K2,
                     const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ Y, int ldy, int M, int N) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float sw[];  // W [N][K1+K2]
    const int K = K1 + K2;
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) sw[i] = W[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < M; row += gridDim.x * 8) {
        float acc[8];
#pragma unroll
        for (int n = 0; n < 8; ++n) acc[n] = 0.f;
        const float* x1 = X1 + (int64_t)row * ldx1;
        for (int k = lane; k < K1; k += 32) {
            float xv = x1[k];
#pragma unroll
            for (int n = 0; n < 8; ++n) if (n < N) acc[n] = fmaf(xv, sw[n * K + k], acc[n]);
        }
        if (X2) {
            const float* x2 = X2 + (int64_t)row * ldx2;
            for (int k = lane; k < K2; k += 32) {
                float xv = x2[k];
#pragma unroll
                for (int n = 0; n < 8; ++n) if (n < N) acc[n] = fmaf(xv, sw[n * K + K1 + k], acc[n]);
            }
        }
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            float v = warp_sum(acc[n]);
            if (n < N && lane == n) Y[(int64_t)row * ldy + n] = v + (bias ? bias[n] : 0.f);
        }
    }
}

}  // namespace

static int opt_in_smem(const void* fn, size_t bytes, const char* who) {
    if (bytes <= 48 * 1024) return AM_OK;
    if (bytes > 227 * 1024) { am_set_error_(who); return AM_EINVAL; }
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) { am_set_error_(who); return AM_ELAUNCH; }
    return AM_OK;
}

extern "C" int am_cdm_encoder_partial(const float* x_t, const float* xyz, const float* w_ea, const float* b_ea, const float* ln_g,
                                      const float* ln_b, const float* qf, int ldq, float* part, int B, int N, int cx, int nchunk,
                                      am_stream_t stream) {
    AM_REQUIRE(x_t && xyz && w_ea && b_ea && ln_g && ln_b && qf && part, AM_EINVAL, "am_cdm_encoder_partial: null pointer");
    AM_REQUIRE(B > 0 && N > 0 && cx > 0 && cx + 3 <= MAXCIN && nchunk > 0 && ldq >= C + 1, AM_EINVAL, "am_cdm_encoder_partial: bad dims");
    size_t smem = sizeof(float) * ((size_t)(cx + 3) * C + (size_t)R * (C + 4) + (size_t)PW * R * (C + 2));
    int rc = opt_in_smem((const void*)cdm_encoder_partial_kernel, smem, "am_cdm_encoder_partial: shared memory");
    if (rc) return rc;
    am_launch(cdm_encoder_partial_kernel, dim3(dim3(nchunk, B)), dim3(PW * 32), smem, as_stream(stream), 1, x_t, xyz, w_ea, b_ea, ln_g, ln_b, qf, ldq, part, N, cx, nchunk);
    AM_LAUNCH_CHECK("cdm_encoder_partial");
    return AM_OK;
}

extern "C" int am_cdm_encoder_combine(const float* part, float* z, int B, int nchunk, am_stream_t stream) {
    AM_REQUIRE(part && z && B > 0 && nchunk > 0, AM_EINVAL, "am_cdm_encoder_combine: bad args");
    am_launch(cdm_encoder_combine_kernel, dim3(dim3(R, B)), dim3(C), 0, as_stream(stream), 1, part, z, nchunk);
    AM_LAUNCH_CHECK("cdm_encoder_combine");
    return AM_OK;
}

extern "C" int am_cdm_decoder_point(const float* x_t, const float* xyz, const float* wd, const float* bd, const float* lnq_g,
                                    const float* lnq_b, const float* kf, int ldk, const float* U, const float* bo, const float* lnm_g,
                                    const float* lnm_b, float* h1, float* hn, void* hn2, int B, int N, int cx, am_stream_t stream) {
    AM_REQUIRE(x_t && xyz && wd && bd && lnq_g && lnq_b && kf && U && bo && lnm_g && lnm_b && h1 && (hn || hn2), AM_EINVAL,
               "am_cdm_decoder_point: null pointer");
    AM_REQUIRE(B > 0 && N > 0 && cx > 0 && cx + 3 <= MAXCIN && ldk >= C + 1, AM_EINVAL, "am_cdm_decoder_point: bad dims");
    size_t smem = sizeof(float) * ((size_t)(cx + 3) * C + (size_t)R * (C + 4) + (size_t)R * C);
    int rc = opt_in_smem((const void*)cdm_decoder_point_kernel, smem, "am_cdm_decoder_point: shared memory");
    if (rc) return rc;
    // ~2 waves of CTAs over the 148 SMs (4 CTAs of 8 warps resident per SM)
    int pts = 64;
    while ((int64_t)cdiv(N, pts) * B > AM_NUM_SMS * 8 && pts < 1024) pts *= 2;
    am_launch(cdm_decoder_point_kernel, dim3(dim3(cdiv(N, pts), B)), dim3(PW * 32), smem, as_stream(stream), 1, x_t, xyz, wd, bd, lnq_g, lnq_b, kf, ldk, U, bo,
                                                                                          lnm_g, lnm_b, h1, hn, reinterpret_cast<__nv_bfloat16*>(hn2), N, cx, pts);
    AM_LAUNCH_CHECK("cdm_decoder_point");
    return AM_OK;
}

extern "C" int am_linear_skinny(const float* X1, int ldx1, int K1, const float* X2, int ldx2, int K2, const float* W, const float* bias,
                                float* Y, int ldy, int M, int N, am_stream_t stream) {
    AM_REQUIRE(X1 && W && Y && M > 0 && N > 0 && N <= 8 && K1 > 0 && K2 >= 0, AM_EINVAL, "am_linear_skinny: bad args (N <= 8)");
    AM_REQUIRE((X2 != nullptr) == (K2 > 0), AM_EINVAL, "am_linear_skinny: X2/K2 mismatch");
    size_t smem = sizeof(float) * (size_t)N * (K1 + K2);
    AM_REQUIRE(smem <= 48 * 1024, AM_EINVAL, "am_linear_skinny: weight does not fit shared memory");
    int grid = cdiv(M, 8) < AM_NUM_SMS * 8 ? cdiv(M, 8) : AM_NUM_SMS * 8;
    am_launch(linear_skinny_kernel, dim3(grid), dim3(256), smem, as_stream(stream), 1, X1, ldx1, K1, X2, ldx2, K2, W, bias, Y, ldy, M, N);
    AM_LAUNCH_CHECK("linear_skinny");
    return AM_OK;
}
