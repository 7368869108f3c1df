// fp32 SIMT GEMM with fused epilogue: Y = act(X W^T + bias) (+ residual), arbitrary M/N/K, token row maps.
// Used for the small / odd-shaped layers (latents, adapters, point-cloud encoder linears) and as the
// on-device cross-check of the tcgen05 path (gemm_tc.cu).  Register-tiled, double-buffered smem.
#include "common.cuh"

namespace {

struct GemmParams {
    const float* X; int ldx;
    const float* W; int ldw;
    float* Y; int ldy;
    int M, N, K;
    const float* bias; int act;
    const float* residual; int ldr; int res_mod;
    int xin_g, xout_g, x_off, yin_g, yout_g, y_off;
    int vecA, vecB, vecY;
    int64_t xb, wb, yb, bb;  // per-batch element strides (blockIdx.z) of X / W / Y / bias — small-M kernel only
};

__device__ __forceinline__ int64_t map_row(int m, int gin, int gout, int off) {
    return gin > 0 ? (int64_t)(m / gin) * gout + off + (m % gin) : (int64_t)m;
}

template <int BM, int BN>
__global__ void __launch_bounds__(256, 2) gemm_f32_kernel(GemmParams p) {  // 2 CTAs/SM: <= 128 registers (130 halves the occupancy: +86 % time measured)
    pdl_launch_dependents();
    pdl_wait();
    constexpr int BK = 16;
    constexpr int TM = BM / 16, TN = BN / 16;  // 8x8 (128 tile) or 4x4 (64 tile)
    constexpr int HM = TM / 2 > 4 ? 4 : (TM >= 4 ? 4 : TM), HN = TN >= 4 ? 4 : TN;
    constexpr int NHM = TM / HM, NHN = TN / HN;  // number of 4-wide fragments per thread (1 or 2)
    constexpr int PAD = 4;
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    // global->register staging: each thread loads float4 chunks along K
    constexpr int A_CHUNKS = BM * BK / 4 / 256, B_CHUNKS = BN * BK / 4 / 256;  // 2 or 1
    float4 ra[A_CHUNKS], rb[B_CHUNKS];
    int64_t a_row[A_CHUNKS]; bool a_ok[A_CHUNKS];
    int64_t b_row[B_CHUNKS]; bool b_ok[B_CHUNKS];
    const int lr = tid >> 2, lk = (tid & 3) * 4;
#pragma unroll
    for (int c = 0; c < A_CHUNKS; ++c) {
        int m = m0 + lr + c * 64;
        a_ok[c] = m < p.M;
        a_row[c] = a_ok[c] ? map_row(m, p.xin_g, p.xout_g, p.x_off) * p.ldx : 0;
    }
#pragma unroll
    for (int c = 0; c < B_CHUNKS; ++c) {
        int n = n0 + lr + c * 64;
        b_ok[c] = n < p.N;
        b_row[c] = b_ok[c] ? (int64_t)n * p.ldw : 0;
    }
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int c = 0; c < A_CHUNKS; ++c) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            int k = k0 + lk;
            if (a_ok[c]) {
                const float* src = p.X + a_row[c] + k;
                if (p.vecA && k + 3 < p.K) v = *reinterpret_cast<const float4*>(src);
                else {
                    if (k < p.K) v.x = src[0];
                    if (k + 1 < p.K) v.y = src[1];
                    if (k + 2 < p.K) v.z = src[2];
                    if (k + 3 < p.K) v.w = src[3];
                }
            }
            ra[c] = v;
        }
#pragma unroll
        for (int c = 0; c < B_CHUNKS; ++c) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            int k = k0 + lk;
            if (b_ok[c]) {
                const float* src = p.W + b_row[c] + k;
                if (p.vecB && k + 3 < p.K) v = *reinterpret_cast<const float4*>(src);
                else {
                    if (k < p.K) v.x = src[0];
                    if (k + 1 < p.K) v.y = src[1];
                    if (k + 2 < p.K) v.z = src[2];
                    if (k + 3 < p.K) v.w = src[3];
                }
            }
            rb[c] = v;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int c = 0; c < A_CHUNKS; ++c) {
            int r = lr + c * 64;
            As[buf][lk + 0][r] = ra[c].x; As[buf][lk + 1][r] = ra[c].y; As[buf][lk + 2][r] = ra[c].z; As[buf][lk + 3][r] = ra[c].w;
        }
#pragma unroll
        for (int c = 0; c < B_CHUNKS; ++c) {
            int r = lr + c * 64;
            Bs[buf][lk + 0][r] = rb[c].x; Bs[buf][lk + 1][r] = rb[c].y; Bs[buf][lk + 2][r] = rb[c].z; Bs[buf][lk + 3][r] = rb[c].w;
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nk = (p.K + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        int buf = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int h = 0; h < NHM; ++h) {
                float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * HM + h * (BM / 2)]);
                a[h * 4 + 0] = v.x; a[h * 4 + 1] = v.y; a[h * 4 + 2] = v.z; a[h * 4 + 3] = v.w;
            }
#pragma unroll
            for (int h = 0; h < NHN; ++h) {
                float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * HN + h * (BN / 2)]);
                b[h * 4 + 0] = v.x; b[h * 4 + 1] = v.y; b[h * 4 + 2] = v.z; b[h * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * HM + (i / 4) * (BM / 2) + (i % 4);
        if (TM == 4) m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
        int64_t yrow = map_row(m, p.yin_g, p.yout_g, p.y_off);
        const float* rrow = nullptr;
        if (p.residual) rrow = p.residual + (p.res_mod > 0 ? (int64_t)(m % p.res_mod) : yrow) * p.ldr;
#pragma unroll
        for (int h = 0; h < NHN; ++h) {
            int n = n0 + tx * HN + h * (BN / 2);
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float x = acc[i][h * 4 + j];
                int nn = n + j;
                if (nn < p.N) {
                    if (p.bias) x += p.bias[nn];
                    if (p.act & AM_ACT_AFTER_RES) {
                        if (rrow) x += rrow[nn];
                        x = apply_act(x, p.act & 15);
                    } else {
                        x = apply_act(x, p.act);
                        if (rrow) x += rrow[nn];
                    }
                }
                v[j] = x;
            }
            float* dst = p.Y + yrow * p.ldy + n;
            if (p.vecY && n + 3 < p.N) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            else
                for (int j = 0; j < 4; ++j)
                    if (n + j < p.N) dst[j] = v[j];
        }
    }
}

// Small-M variant (the 2-latent-token side of the CDM Perceiver: M = 2 B rows, and its per-head fold GEMMs).  The tiled kernel
// above launches cdiv(N, 64) x cdiv(M, 64) CTAs — 8 CTAs for a [16 x 512] x [512 x 512] layer, each walking K serially: 16.6 us
// per launch, 44 launches per denoise step = half of the CDM step at 8 samples per GPU.  Here parallelism comes from N: one warp
// per output column, lanes split K (coalesced 128-bit reads of the weight row), the <= 16 activation rows of the CTA sit in
// shared memory, 16 accumulators per lane, warp-shuffle reduction.  Same epilogue semantics (bias, activation, residual,
// token row maps) as gemm_f32_kernel.
constexpr int SM_ROWS = 16, SM_WARPS = 8;
__global__ void __launch_bounds__(SM_WARPS * 32) gemm_smallm_kernel(GemmParams p) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float xs[];  // [SM_ROWS][Kp], Kp = K rounded up to 4
    const int Kp = (p.K + 3) & ~3;
    const int m0 = blockIdx.y * SM_ROWS;
    const int rows = min(SM_ROWS, p.M - m0);
    {   // batch (blockIdx.z): the per-head fold GEMMs of the Perceiver differ only by pointer offsets
        const int64_t z = blockIdx.z;
        p.X += z * p.xb; p.W += z * p.wb; p.Y += z * p.yb;
        if (p.bias) p.bias += z * p.bb;
    }
    if (p.vecA && (p.K & 3) == 0 && (p.xb & 3) == 0) {
        const int k4n = Kp >> 2;
        for (int i = threadIdx.x; i < SM_ROWS * k4n; i += blockDim.x) {
            const int r = i / k4n, k4 = i - r * k4n;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < rows) v = *reinterpret_cast<const float4*>(p.X + map_row(m0 + r, p.xin_g, p.xout_g, p.x_off) * p.ldx + 4 * k4);
            reinterpret_cast<float4*>(xs)[i] = v;
        }
    } else {
        for (int i = threadIdx.x; i < SM_ROWS * Kp; i += blockDim.x) {
            const int r = i / Kp, k = i - r * Kp;
            float v = 0.f;
            if (r < rows && k < p.K) v = p.X[map_row(m0 + r, p.xin_g, p.xout_g, p.x_off) * p.ldx + k];
            xs[i] = v;
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * SM_WARPS + warp;
    if (n >= p.N) return;
    float acc[SM_ROWS];
#pragma unroll
    for (int r = 0; r < SM_ROWS; ++r) acc[r] = 0.f;
    const float* wrow = p.W + (int64_t)n * p.ldw;
    if (p.vecB && (p.wb & 3) == 0) {
        for (int k = lane * 4; k < Kp; k += 128) {
            float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k + 3 < p.K) w4 = *reinterpret_cast<const float4*>(wrow + k);
            else { if (k < p.K) w4.x = wrow[k]; if (k + 1 < p.K) w4.y = wrow[k + 1]; if (k + 2 < p.K) w4.z = wrow[k + 2]; }
#pragma unroll
            for (int r = 0; r < SM_ROWS; ++r) {
                const float4 x4 = *reinterpret_cast<const float4*>(xs + r * Kp + k);
                acc[r] = fmaf(x4.x, w4.x, fmaf(x4.y, w4.y, fmaf(x4.z, w4.z, fmaf(x4.w, w4.w, acc[r]))));
            }
        }
    } else {
        for (int k = lane; k < p.K; k += 32) {
            const float w = wrow[k];
#pragma unroll
            for (int r = 0; r < SM_ROWS; ++r) acc[r] = fmaf(xs[r * Kp + k], w, acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < SM_ROWS; ++r) acc[r] = warp_sum(acc[r]);
    // lane r finishes output row m0 + r of column n
    float x = 0.f;
#pragma unroll
    for (int r = 0; r < SM_ROWS; ++r) x = lane == r ? acc[r] : x;
    if (lane < rows) {
        const int m = m0 + lane;
        const int64_t yrow = map_row(m, p.yin_g, p.yout_g, p.y_off);
        if (p.bias) x += p.bias[n];
        float rres = 0.f;
        if (p.residual) rres = p.residual[(p.res_mod > 0 ? (int64_t)(m % p.res_mod) : yrow) * p.ldr + n];
        if (p.act & AM_ACT_AFTER_RES) x = apply_act(x + rres, p.act & 15);
        else x = apply_act(x, p.act) + rres;
        p.Y[yrow * p.ldy + n] = x;
    }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

int am_rowgemm_fwd_(const float* X, int ldx, const float* W, int ldw, int transW, float* Y, int ldy, int M, int N, int K, const float* bias,
                    int act, const float* residual, int ldr, cudaStream_t st);
static inline bool rowgemm_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("AMB200_ROWGEMM"); on = (e && e[0] == '0') ? 0 : 1; }
    return on != 0;
}

static int linear_f32_launch(const float* X, int ldx, const float* W, int ldw, float* Y, int ldy, int M, int N, int K, const float* bias,
                             int act, const float* residual, int ldr, int res_mod, int xin_g, int xout_g, int x_off, int yin_g,
                             int yout_g, int y_off, int nbatch, int64_t xb, int64_t wb, int64_t yb, int64_t bb, am_stream_t stream) {
    AM_REQUIRE(X && W && Y, AM_EINVAL, "am_linear_f32: null pointer");
    AM_REQUIRE(M > 0 && N > 0 && K > 0 && ldx >= K && ldw >= K && ldy >= N && nbatch >= 1, AM_EINVAL, "am_linear_f32: bad dims");
    AM_REQUIRE((act & 15) >= 0 && (act & 15) <= 3 && (act & ~31) == 0, AM_EINVAL, "am_linear_f32: bad activation");
    AM_REQUIRE(!residual || ldr >= N, AM_EINVAL, "am_linear_f32: bad residual stride");
    GemmParams p{X, ldx, W, ldw, Y, ldy, M, N, K, bias, act, residual, ldr, res_mod, xin_g, xout_g, x_off, yin_g, yout_g, y_off, 0, 0, 0,
                 xb, wb, yb, bb};
    p.vecA = al16(X) && (ldx % 4 == 0);
    p.vecB = al16(W) && (ldw % 4 == 0);
    p.vecY = al16(Y) && (ldy % 4 == 0);
    // tall-skinny (N <= 32, many rows: per-neighbour MLPs of the Point-Transformer encoder): one thread per row (csrc/rowgemm.cu)
    if (nbatch == 1 && xin_g == 0 && yin_g == 0 && res_mod == 0 && rowgemm_enabled() &&
        am_rowgemm_fwd_(X, ldx, W, ldw, 0, Y, ldy, M, N, K, bias, act, residual, ldr, as_stream(stream))) {
        AM_LAUNCH_CHECK("linear_f32");
        return AM_OK;
    }
    // small M (latent tokens): column-parallel kernel when the tiled one would launch only a handful of CTAs
    const size_t sm_smem = sizeof(float) * SM_ROWS * (size_t)((K + 3) & ~3);
    if (M <= 128 && (int64_t)cdiv(M, 64) * cdiv(N, 64) < 32 && sm_smem <= 48 * 1024) {
        AM_REQUIRE(nbatch == 1 || !residual, AM_EINVAL, "am_linear_f32_batched: residual is not supported with nbatch > 1");
        am_launch(gemm_smallm_kernel, dim3(cdiv(N, SM_WARPS), cdiv(M, SM_ROWS), nbatch), dim3(SM_WARPS * 32), sm_smem, as_stream(stream), 1, p);
        AM_LAUNCH_CHECK("linear_f32");
        return AM_OK;
    }
    for (int z = 0; z < nbatch; ++z) {  // large shapes: one tiled launch per batch entry
        GemmParams q = p;
        q.X += z * xb; q.W += z * wb; q.Y += z * yb;
        if (q.bias) q.bias += z * bb;
        q.vecA = al16(q.X) && (ldx % 4 == 0);
        q.vecB = al16(q.W) && (ldw % 4 == 0);
        q.vecY = al16(q.Y) && (ldy % 4 == 0);
        // tile choice: 128x128 when it still fills the 148 SMs, else 64x64
        int64_t tiles128 = (int64_t)cdiv(M, 128) * cdiv(N, 128);
        if (tiles128 >= AM_NUM_SMS) {
            dim3 grid(cdiv(N, 128), cdiv(M, 128));
            am_launch(gemm_f32_kernel<128, 128>, dim3(grid), dim3(256), 0, as_stream(stream), 1, q);
        } else {
            dim3 grid(cdiv(N, 64), cdiv(M, 64));
            am_launch(gemm_f32_kernel<64, 64>, dim3(grid), dim3(256), 0, as_stream(stream), 1, q);
        }
        AM_LAUNCH_CHECK("linear_f32");
    }
    return AM_OK;
}

extern "C" int am_linear_f32(const float* X, int ldx, const float* W, int ldw, float* Y, int ldy, int M, int N, int K, const float* bias,
                             int act, const float* residual, int ldr, int res_mod, int xin_g, int xout_g, int x_off, int yin_g,
                             int yout_g, int y_off, am_stream_t stream) {
    return linear_f32_launch(X, ldx, W, ldw, Y, ldy, M, N, K, bias, act, residual, ldr, res_mod, xin_g, xout_g, x_off, yin_g, yout_g, y_off,
                             1, 0, 0, 0, 0, stream);
}

extern "C" int am_linear_f32_batched(const float* X, int ldx, const float* W, int ldw, float* Y, int ldy, int M, int N, int K,
                                     const float* bias, int act, int xin_g, int xout_g, int x_off, int yin_g, int yout_g, int y_off,
                                     int nbatch, int64_t x_bstride, int64_t w_bstride, int64_t y_bstride, int64_t bias_bstride,
                                     am_stream_t stream) {
    return linear_f32_launch(X, ldx, W, ldw, Y, ldy, M, N, K, bias, act, nullptr, 0, 0, xin_g, xout_g, x_off, yin_g, yout_g, y_off, nbatch,
                             x_bstride, w_bstride, y_bstride, bias_bstride, stream);
}
