// CDM ContactPerceiver latent side (the 2 latent tokens per sample: language + time; models/cdm.py:176-185, Perceiver-IO blocks
// models/modules.py:504-648) as TWO thread-block-cluster kernels instead of ~29 tiny launches per denoise step.
//
// After the rank collapse of the point path (perceiver_tc.cu) the latent chain — 25 dependent layers on [2B, 512] activations,
// 18 MB of weights — was the whole CDM step at the config-3 shard (8 samples per GPU: ~200 us of ~250 us, launch-latency bound).
// Here one cluster of 8 CTAs owns 4 samples (8 activation rows).  Every CTA keeps a full copy of the 8 x 512 activation block in
// shared memory, computes 1/8 of each layer's output columns (weights stored K-major [K][N]: lane = column, coalesced 128-byte
// reads; K split over 8 warps, partial sums combined through shared memory) and broadcasts its slice into all 8 CTAs' buffers
// through distributed shared memory; layers are separated by one cluster barrier (~0.4 us) instead of a kernel launch.
//   cdm_latent_pre_kernel  : L0 = [language latent ; time table[t]] -> LN_q -> q_proj (scaled) -> AE = per-head fold of the query
//                            against the collapsed key matrices  (input of cdm_enc_points_kernel)
//   cdm_latent_post_kernel : encoder softmax statistic -> z -> V / o_proj + residual -> MLP -> 2 x latent self-attention layers ->
//                            decoder LN_kv -> [K | V] tokens -> AQ (decoder q fold) and UU (o_proj stack) per head
//                            (inputs of cdm_dec_prep_kernel / cdm_dec_points_tc_kernel)
// fp32 throughout (exact erf GELU, LayerNorm eps 1e-5): same arithmetic as the per-layer kernels it replaces.
#include <math_constants.h>
#include <string.h>
#include "common.cuh"

namespace {

constexpr int DL = 512;        // latent width (encoder_q_input_channels)
constexpr int C = 256;         // point channels
constexpr int H = 8;           // heads (encoder and decoder)
constexpr int HD = DL / H;     // 64
constexpr int HDD = C / H;     // 32
constexpr int R16 = 16;        // (head, latent) rows of the encoder statistic
constexpr int KU = 10, AEW = KU + 2;
constexpr int NC = 8;          // CTAs per cluster (one per head in the per-head stages)
constexpr int SPC = 4;         // samples per cluster
constexpr int R = 2 * SPC;     // activation rows per cluster
constexpr int LT = 512;        // threads per CTA: 16 warps = 2 column groups x 8 K splits
constexpr int KS = 8;          // K splits per layer
constexpr int LQKV = 3 * DL;
// shared memory (floats): X0 stream | X1 | XQ [R][1536] | X2 [R][512] | red [KS][R][64]   (Z [SPC][16][256] aliases XQ..X2)
constexpr int SX0 = 0, SX1 = SX0 + R * DL, SXQ = SX1 + R * DL, SX2 = SXQ + R * LQKV, SRED = SX2 + R * DL, SEND = SRED + KS * R * 64;
static_assert(SPC * R16 * C <= R * LQKV + R * DL, "Z must fit in the XQ|X2 alias");
constexpr int LAT_SMEM = SEND * 4;

struct LatW {  // K-major ([K][N]) weights + biases of the latent chain; all fp32 device pointers
    const float *la_unused;
    const float *eq_g, *eq_b, *eq_wt, *eq_bias;            // encoder q_norm, q_proj (scale folded in)
    const float *e_kfold;                                   // [H][AEW][HD]
    const float *ecg, *ebeta;                               // [C][KU], [C]   (z expansion)
    const float *ev_wt, *ev_b, *eo_wt, *eo_b;              // v_proj [C][DL], o_proj [DL][DL]
    const float *em_g, *em_b, *em1_wt, *em1_b, *em2_wt, *em2_b;
    const float *s_n_g[2], *s_n_b[2], *s_qkv_wt[2], *s_qkv_b[2], *s_o_wt[2], *s_o_b[2];
    const float *s_m_g[2], *s_m_b[2], *s_m1_wt[2], *s_m1_b[2], *s_m2_wt[2], *s_m2_b[2];
    const float *dkv_g, *dkv_b, *dkv_wt, *dkv_bias;        // decoder kv_norm, [k_proj ; v_proj] [DL][2C]
    const float *d_qfold;                                   // [H][AEW][HDD]
    const float *d_ostack_t;                                // [C][NS]  (K-major o_proj stack)
};

__device__ __forceinline__ uint32_t smem_u32l(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_remote(uint32_t local_addr, uint32_t cta, float v) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(cta));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}

// One layer: Y[r][n] = act( sum_k X[row(r, n)][k] WT[k][n] + bias[n] ) (+ RES[r][n]) for this CTA's N / NC columns, all R rows; the
// slice is written into the Y buffer of EVERY CTA of the cluster.  HEADROWS: X is the encoder statistic Z [SPC][16][C] and column
// n (head h = n / HD) reads row 2h + l of its sample (the per-head V projection, cdm.py:180).  Ends with a cluster barrier.
template <int ACT, bool HEADROWS>
__device__ __forceinline__ void lat_gemm(const float* __restrict__ Xs, int ldx, int K, const float* __restrict__ WT, int N,
                                         const float* __restrict__ bias, const float* RES, float* Ys, int ldy, float* red,
                                         uint32_t rank, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    const int cg = warp & 1, kq = warp >> 1;
    const int ncol = N / NC;          // 64 (N = 512) or 192 (N = 1536)
    const int kper = K / KS;
    for (int c0 = 0; c0 < ncol; c0 += 64) {
        const int n = (int)rank * ncol + c0 + cg * 32 + lane;
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        const float* xrow[R];
#pragma unroll
        for (int r = 0; r < R; ++r) xrow[r] = HEADROWS ? Xs + ((r >> 1) * R16 + 2 * (n / HD) + (r & 1)) * ldx : Xs + r * ldx;
        const float* w = WT + (int64_t)(kq * kper) * N + n;
        // the layer is latency-bound on the weight stream (each lane walks one column, 128-byte warp requests): keep 32 loads in
        // flight per lane (KB rows of W at a time) before the FMAs consume them
        constexpr int KB = 32;
        for (int k = 0; k < kper; k += KB) {
            float wv[KB];
#pragma unroll
            for (int j = 0; j < KB; ++j) wv[j] = __ldg(w + (int64_t)j * N);
            w += (int64_t)KB * N;
#pragma unroll
            for (int j = 0; j < KB; j += 4) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float4 x = *reinterpret_cast<const float4*>(xrow[r] + kq * kper + k + j);
                    acc[r] = fmaf(x.x, wv[j], fmaf(x.y, wv[j + 1], fmaf(x.z, wv[j + 2], fmaf(x.w, wv[j + 3], acc[r]))));
                }
            }
        }
        __syncthreads();  // `red` of the previous pass / layer fully consumed
#pragma unroll
        for (int r = 0; r < R; ++r) red[(kq * R + r) * 64 + cg * 32 + lane] = acc[r];
        __syncthreads();
        for (int i = tid; i < R * 64; i += LT) {
            const int r = i >> 6, c = i & 63;
            const int nn = (int)rank * ncol + c0 + c;
            float v = 0.f;
#pragma unroll
            for (int q = 0; q < KS; ++q) v += red[(q * R + r) * 64 + c];
            if (bias) v += __ldg(bias + nn);
            if (ACT == AM_ACT_GELU) v = gelu_erf(v);
            if (RES) v += RES[r * ldy + nn];
            const uint32_t a = smem_u32l(Ys + r * ldy + nn);
#pragma unroll
            for (int t = 0; t < NC; ++t) st_remote(a, (uint32_t)t, v);
        }
    }
    cluster_sync_all();
}

// LayerNorm of the R rows (local copy; every CTA normalises its own copy): one warp per row
__device__ __forceinline__ void lat_layernorm(const float* Xs, float* Ys, const float* __restrict__ g, const float* __restrict__ b, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < R; r += LT / 32) {
        float v[DL / 32];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < DL / 32; ++i) { v[i] = Xs[r * DL + lane + 32 * i]; s += v[i]; }
        const float mean = warp_sum(s) * (1.0f / DL);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < DL / 32; ++i) { const float d = v[i] - mean; q += d * d; }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / DL) + 1e-5f);
#pragma unroll
        for (int i = 0; i < DL / 32; ++i) {
            const int c = lane + 32 * i;
            Ys[r * DL + c] = (v[i] - mean) * rstd * __ldg(g + c) + __ldg(b + c);
        }
    }
    __syncthreads();
}

// L0 rows of this cluster's samples: row 2s = language latent, row 2s + 1 = time table[t]   (cdm.py:176-178)
__device__ __forceinline__ void lat_load_l0(float* X0, const float* __restrict__ text_latent, const float* __restrict__ time_table,
                                            const int32_t* __restrict__ t, int t_stride, int b0, int B, int tid) {
    for (int i = tid; i < R * DL; i += LT) {
        const int r = i / DL, c = i - r * DL, b = b0 + (r >> 1);
        float v = 0.f;
        if (b < B) v = (r & 1) ? __ldg(time_table + (int64_t)__ldg(t + b * t_stride) * DL + c) : __ldg(text_latent + (int64_t)b * DL + c);
        X0[i] = v;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(LT, 1)
cdm_latent_pre_kernel(LatW w, const float* __restrict__ text_latent, const float* __restrict__ time_table, const int32_t* __restrict__ t,
                      int t_stride, float* __restrict__ AE, int B) {
    extern __shared__ __align__(16) float sm[];
    float *X0 = sm + SX0, *X1 = sm + SX1, *X2 = sm + SX2, *red = sm + SRED;
    const int tid = threadIdx.x;
    const uint32_t rank = cluster_rank();
    const int b0 = (blockIdx.x / NC) * SPC;
    pdl_launch_dependents();
    pdl_wait();
    lat_load_l0(X0, text_latent, time_table, t, t_stride, b0, B, tid);
    lat_layernorm(X0, X1, w.eq_g, w.eq_b, tid);
    cluster_sync_all();  // every CTA's X2 is idle before remote writes start
    lat_gemm<AM_ACT_NONE, false>(X1, DL, DL, w.eq_wt, DL, w.eq_bias, nullptr, X2, DL, red, rank, tid);
    // AE[b, 2h + l, n] = sum_k q[2s + l][h*HD + k] e_kfold[h][n][k], head h = this CTA   (modules.py:335-352 folded, cdm_fold.py)
    const int h = (int)rank;
    for (int i = tid; i < R * AEW; i += LT) {
        const int r = i / AEW, n = i - r * AEW, b = b0 + (r >> 1);
        if (b >= B) continue;
        const float* q = X2 + r * DL + h * HD;
        const float* kf = w.e_kfold + ((int64_t)h * AEW + n) * HD;
        float a = 0.f;
#pragma unroll 8
        for (int k = 0; k < HD; ++k) a = fmaf(q[k], __ldg(kf + k), a);
        AE[((int64_t)b * R16 + 2 * h + (r & 1)) * AEW + n] = a;
    }
}

__global__ void __launch_bounds__(LT, 1)
cdm_latent_post_kernel(LatW w, const float* __restrict__ text_latent, const float* __restrict__ time_table, const int32_t* __restrict__ t,
                       int t_stride, const float* __restrict__ part, int nchunk, float* __restrict__ AQ, float* __restrict__ UU, int NS,
                       int B) {
    extern __shared__ __align__(16) float sm[];
    float *X0 = sm + SX0, *X1 = sm + SX1, *XQ = sm + SXQ, *X2 = sm + SX2, *red = sm + SRED;
    float* Z = XQ;  // [SPC][16][C], dead after the V projection
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_rank();
    const int b0 = (blockIdx.x / NC) * SPC;
    pdl_launch_dependents();
    pdl_wait();
    lat_load_l0(X0, text_latent, time_table, t, t_stride, b0, B, tid);
    // ---- encoder softmax statistic: combine the flash partials, w_r = sum acc / sum l  -> red[SPC*16][AEW] (first 768 floats)
    for (int i = tid; i < SPC * R16; i += LT) {
        const int s = i / R16, r = i - s * R16, b = b0 + s;
        float o[KU], ls = 0.f, M = -CUDART_INF_F;
#pragma unroll
        for (int k = 0; k < KU; ++k) o[k] = 0.f;
        if (b < B) {
            const float* base = part + ((int64_t)b * nchunk * R16 + r) * AEW;
            for (int c = 0; c < nchunk; ++c) M = fmaxf(M, base[(int64_t)c * R16 * AEW + KU]);
            for (int c = 0; c < nchunk; ++c) {
                const float* pc = base + (int64_t)c * R16 * AEW;
                const float f = pc[KU] == -CUDART_INF_F ? 0.f : expf(pc[KU] - M);
                ls = fmaf(f, pc[KU + 1], ls);
#pragma unroll
                for (int k = 0; k < KU; ++k) o[k] = fmaf(f, pc[k], o[k]);
            }
        }
        const float inv = ls > 0.f ? 1.0f / ls : 0.f;
#pragma unroll
        for (int k = 0; k < KU; ++k) red[i * AEW + k] = o[k] * inv;
    }
    __syncthreads();
    // z[s][r][c] = diag(g) Ec w_r + beta  (the softmax-weighted mean of LN_kv(enc_kv) rows)
    for (int i = tid; i < SPC * R16 * C; i += LT) {
        const int c = i % C, sr = i / C;
        const float* e = w.ecg + c * KU;
        float v = __ldg(w.ebeta + c);
#pragma unroll
        for (int k = 0; k < KU; ++k) v = fmaf(__ldg(e + k), red[sr * AEW + k], v);
        Z[i] = v;
    }
    __syncthreads();
    cluster_sync_all();
    // ---- encoder cross-attention tail: per-head V projection, o_proj + residual (un-normalised L, modules.py:230), MLP
    lat_gemm<AM_ACT_NONE, true>(Z, C, C, w.ev_wt, DL, w.ev_b, nullptr, X1, DL, red, rank, tid);
    lat_gemm<AM_ACT_NONE, false>(X1, DL, DL, w.eo_wt, DL, w.eo_b, X0, X0, DL, red, rank, tid);
    lat_layernorm(X0, X1, w.em_g, w.em_b, tid);
    lat_gemm<AM_ACT_GELU, false>(X1, DL, DL, w.em1_wt, DL, w.em1_b, nullptr, X2, DL, red, rank, tid);
    lat_gemm<AM_ACT_NONE, false>(X2, DL, DL, w.em2_wt, DL, w.em2_b, X0, X0, DL, red, rank, tid);
    // ---- 2 x latent self-attention layers (modules.py:544-648): 2 tokens per sample, 8 heads x 64
#pragma unroll 1
    for (int li = 0; li < 2; ++li) {
        lat_layernorm(X0, X1, w.s_n_g[li], w.s_n_b[li], tid);
        lat_gemm<AM_ACT_NONE, false>(X1, DL, DL, w.s_qkv_wt[li], LQKV, w.s_qkv_b[li], nullptr, XQ, LQKV, red, rank, tid);
        // attention (every CTA computes all heads of its local copy): warp per (sample, head)
        for (int p = warp; p < SPC * H; p += LT / 32) {
            const int s = p / H, h = p - s * H;
            const float* q0 = XQ + (2 * s) * LQKV + h * HD;
            const float* q1 = q0 + LQKV;
            const float *k0 = q0 + DL, *k1 = q1 + DL, *v0 = q0 + 2 * DL, *v1 = q1 + 2 * DL;
            float s00 = q0[lane] * k0[lane] + q0[lane + 32] * k0[lane + 32], s01 = q0[lane] * k1[lane] + q0[lane + 32] * k1[lane + 32];
            float s10 = q1[lane] * k0[lane] + q1[lane + 32] * k0[lane + 32], s11 = q1[lane] * k1[lane] + q1[lane + 32] * k1[lane + 32];
            const float sc = 0.125f;  // HD^-0.5
            s00 = warp_sum(s00) * sc; s01 = warp_sum(s01) * sc; s10 = warp_sum(s10) * sc; s11 = warp_sum(s11) * sc;
            const float m0 = fmaxf(s00, s01), m1 = fmaxf(s10, s11);
            const float e00 = expf(s00 - m0), e01 = expf(s01 - m0), e10 = expf(s10 - m1), e11 = expf(s11 - m1);
            const float i0 = 1.0f / (e00 + e01), i1 = 1.0f / (e10 + e11);
#pragma unroll
            for (int d = lane; d < HD; d += 32) {
                X1[(2 * s) * DL + h * HD + d] = (e00 * v0[d] + e01 * v1[d]) * i0;
                X1[(2 * s + 1) * DL + h * HD + d] = (e10 * v0[d] + e11 * v1[d]) * i1;
            }
        }
        __syncthreads();
        cluster_sync_all();  // all CTAs are done reading XQ / writing their X1 before the next layer's remote writes
        lat_gemm<AM_ACT_NONE, false>(X1, DL, DL, w.s_o_wt[li], DL, w.s_o_b[li], X0, X0, DL, red, rank, tid);
        lat_layernorm(X0, X1, w.s_m_g[li], w.s_m_b[li], tid);
        lat_gemm<AM_ACT_GELU, false>(X1, DL, DL, w.s_m1_wt[li], DL, w.s_m1_b[li], nullptr, X2, DL, red, rank, tid);
        lat_gemm<AM_ACT_NONE, false>(X2, DL, DL, w.s_m2_wt[li], DL, w.s_m2_b[li], X0, X0, DL, red, rank, tid);
    }
    // ---- decoder: LN_kv of the latents, [K | V] tokens (2 per sample, C each), then the per-head folds (head = this CTA)
    lat_layernorm(X0, X1, w.dkv_g, w.dkv_b, tid);
    lat_gemm<AM_ACT_NONE, false>(X1, DL, DL, w.dkv_wt, 2 * C, w.dkv_bias, nullptr, X2, 2 * C, red, rank, tid);
    const int h = (int)rank;
    for (int i = tid; i < R * AEW; i += LT) {  // AQ[b, 2h + l, n] = sum_k ktok[l][h*HDD + k] d_qfold[h][n][k]
        const int r = i / AEW, n = i - r * AEW, b = b0 + (r >> 1);
        if (b >= B) continue;
        const float* kt = X2 + r * (2 * C) + h * HDD;
        const float* qf = w.d_qfold + ((int64_t)h * AEW + n) * HDD;
        float a = 0.f;
#pragma unroll 8
        for (int k = 0; k < HDD; ++k) a = fmaf(kt[k], __ldg(qf + k), a);
        AQ[((int64_t)b * R16 + 2 * h + (r & 1)) * AEW + n] = a;
    }
    for (int n = tid; n < NS; n += LT) {       // UU[b, 2h + l, n] = sum_k vtok[l][h*HDD + k] ostack[n][h*HDD + k]
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        const float* wt = w.d_ostack_t + (int64_t)(h * HDD) * NS + n;
#pragma unroll 4
        for (int k = 0; k < HDD; ++k) {
            const float wv = __ldg(wt + (int64_t)k * NS);
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = fmaf(X2[r * (2 * C) + C + h * HDD + k], wv, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int b = b0 + (r >> 1);
            if (b < B) UU[((int64_t)b * R16 + 2 * h + (r & 1)) * NS + n] = acc[r];
        }
    }
}

static int set_smem(const void* fn, const char* who) {
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, LAT_SMEM) != cudaSuccess) { am_set_error_(who); return AM_ELAUNCH; }
    return AM_OK;
}

}  // namespace

// W: HOST array of device pointers in the order of struct LatW (am_cdm_latent_nweights() entries; amb200/cdm_engine.py fills it)
extern "C" int am_cdm_latent_pre(const void* const* W, int nW, const float* text_latent, const float* time_table, const int32_t* t,
                                 int t_stride, float* AE, int B, am_stream_t stream) {
    AM_REQUIRE(W && nW == (int)(sizeof(LatW) / sizeof(void*)) && text_latent && time_table && t && AE && B > 0, AM_EINVAL,
               "am_cdm_latent_pre: bad args (pointer table must match struct LatW)");
    AM_REQUIRE(t_stride == 0 || t_stride == 1, AM_EINVAL, "am_cdm_latent_pre: t_stride must be 0 or 1");
    LatW w;
    memcpy(&w, W, sizeof(LatW));
    static bool attr = false;
    if (!attr) { int rc = set_smem((const void*)cdm_latent_pre_kernel, "am_cdm_latent_pre: shared memory opt-in failed"); if (rc) return rc; attr = true; }
    const int nclusters = cdiv(B, SPC);
    if (am_launch(cdm_latent_pre_kernel, dim3(nclusters * NC), dim3(LT), LAT_SMEM, as_stream(stream), NC, w, text_latent, time_table, t, t_stride, AE, B) != cudaSuccess) {
        am_set_error_("am_cdm_latent_pre: cluster launch failed");
        return AM_ELAUNCH;
    }
    AM_LAUNCH_CHECK("cdm_latent_pre");
    return AM_OK;
}

extern "C" int am_cdm_latent_post(const void* const* W, int nW, const float* text_latent, const float* time_table, const int32_t* t,
                                  int t_stride, const float* part, int nchunk, float* AQ, float* UU, int NS, int B, am_stream_t stream) {
    AM_REQUIRE(W && nW == (int)(sizeof(LatW) / sizeof(void*)) && text_latent && time_table && t && part && AQ && UU && B > 0 && nchunk > 0, AM_EINVAL,
               "am_cdm_latent_post: bad args (pointer table must match struct LatW)");
    AM_REQUIRE(t_stride == 0 || t_stride == 1, AM_EINVAL, "am_cdm_latent_post: t_stride must be 0 or 1");
    AM_REQUIRE(NS >= 2 * C + KU + 6 && NS % 4 == 0, AM_EINVAL, "am_cdm_latent_post: bad o_proj stack width");
    LatW w;
    memcpy(&w, W, sizeof(LatW));
    static bool attr = false;
    if (!attr) { int rc = set_smem((const void*)cdm_latent_post_kernel, "am_cdm_latent_post: shared memory opt-in failed"); if (rc) return rc; attr = true; }
    const int nclusters = cdiv(B, SPC);
    if (am_launch(cdm_latent_post_kernel, dim3(nclusters * NC), dim3(LT), LAT_SMEM, as_stream(stream), NC, w, text_latent, time_table, t, t_stride, part, nchunk,
                  AQ, UU, NS, B) != cudaSuccess) {
        am_set_error_("am_cdm_latent_post: cluster launch failed");
        return AM_ELAUNCH;
    }
    AM_LAUNCH_CHECK("cdm_latent_post");
    return AM_OK;
}

extern "C" int am_cdm_latent_nweights(void) { return (int)(sizeof(LatW) / sizeof(void*)); }
