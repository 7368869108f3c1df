// CDM ContactPerceiver point path, rank-collapsed (models/cdm.py:155-188,511; Perceiver-IO blocks models/modules.py:222-661).
//
// Every per-point quantity of the Perceiver is a function of the 9 input floats u = cat(x_t, xyz) pushed through affine maps,
// two LayerNorms, two softmaxes and one GELU (derivation + weights-only constants: amb200/cdm_fold.py).  Per point:
//   encoder : rstd_e(u) from a 10-dim quadratic form (Cholesky factor: sum of squares), 16 scores = rstd_e * (A_e [u;1]) + c_e,
//             online softmax over the points accumulating only the 10-vector sum_j p_j rstd_j [u_j;1]   (cdm_enc_points_kernel)
//   decoder : rstd_q(u), 16 scores, per-head softmax over the 2 latents -> z = [u; 1; p] (26 entries), rstd_1 from the 26-dim
//             quadratic form, then the ONE dense layer  W1 LN_m(h1) + b1 = rstd_1 * (M z) + c1  as a [128 x 32] x [32 x 256]
//             tcgen05 GEMM per 128-point tile (3-term bf16 split, fp32 TMEM accumulation), GELU and the folded 256 -> 6 head
//             straight out of TMEM                                                                  (cdm_dec_points_tc_kernel)
// HBM traffic per point: 36 B read twice (encoder, decoder) + 24 B written; enc_kv / K / V / dq / h1 / LN(h1) / GELU activations
// ([B,N,256..512] each in the reference) never exist.  The general-cin fallback (scene features, cin = 41) is csrc/perceiver.cu.
#include <cuda_bf16.h>
#include <math_constants.h>
#include "common.cuh"

namespace {

constexpr int C = 256;              // point channels
constexpr int R = 16;               // (head, latent) rows
constexpr int CX = 6;               // contact_dim (x_t features)
constexpr int KU = CX + 3 + 1;      // [u; 1]
constexpr int KZ = KU + R;          // [u; 1; p]
constexpr int KP = 32;              // GEMM K (KZ zero padded): one 64-byte SWIZZLE_64B row of bf16
constexpr int J = 6;                // output channels
constexpr int AEW = KU + 2;         // row stride of the per-sample score matrices: coefficients[KU], constant, pad
constexpr int NCHOL = KU * (KU + 1) / 2;
constexpr int NGT = KZ * (KZ + 1) / 2;  // 351
// per-sample parameter block (floats) written by cdm_dec_prep_kernel
constexpr int PB_AQ = 0;                    // [R][AEW]
constexpr int PB_GT = R * AEW;              // packed upper triangle of G1 (row-major, off-diagonals doubled)
constexpr int PB_HP = PB_GT + 352;          // [J][28]: head coefficients of z
constexpr int HPW = 28;
constexpr int PB_STRIDE = 768;
static_assert(PB_HP + J * HPW <= PB_STRIDE, "parameter block overflow");
constexpr int BLOB_BYTES = 2 * C * KP * 2;  // per-sample B operand: hi [256 x 64 B] then lo, rows SWIZZLE_64B
constexpr int TILE = 128;                   // points per tile (UMMA M)

__device__ __forceinline__ void load_u(const float* __restrict__ x_t, const float* __restrict__ xyz, int64_t pt, bool valid, float u[KU]) {
    if (valid) {
        const float2* xp = reinterpret_cast<const float2*>(x_t + pt * CX);
        const float2 a = __ldg(xp), b = __ldg(xp + 1), c = __ldg(xp + 2);
        const float* pp = xyz + pt * 3;
        u[0] = a.x; u[1] = a.y; u[2] = b.x; u[3] = b.y; u[4] = c.x; u[5] = c.y;
        u[6] = __ldg(pp); u[7] = __ldg(pp + 1); u[8] = __ldg(pp + 2);
    } else {
#pragma unroll
        for (int i = 0; i < KU - 1; ++i) u[i] = 0.f;
    }
    u[KU - 1] = 1.f;
}

// 1 / sqrt([u;1]^T G [u;1] + eps) with G = T^T T given by its packed upper-triangular factor (shared memory, broadcast reads)
__device__ __forceinline__ float rstd_chol(const float* __restrict__ T, const float u[KU]) {
    float var = 0.f;
    int idx = 0;
#pragma unroll
    for (int i = 0; i < KU; ++i) {
        float v = 0.f;
#pragma unroll
        for (int k = i; k < KU; ++k) v = fmaf(T[idx++], u[k], v);
        var = fmaf(v, v, var);
    }
    return rsqrtf(var + 1e-5f);
}

// ------------------------------------------------------------------------------------------------ encoder
// part layout: [B][nchunk][R][AEW] = (acc[KU], max, sum)
__global__ void __launch_bounds__(256)
cdm_enc_points_kernel(const float* __restrict__ x_t, const float* __restrict__ xyz, const float* __restrict__ chol,
                      const float* __restrict__ AE, float* __restrict__ part, int N, int nchunk) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float s_chol[NCHOL + 1];
    __shared__ float s_red[8][R][AEW];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rq = tid & 3, ps = tid >> 2;  // 4 threads per point, 4 score rows each
    if (tid < NCHOL) s_chol[tid] = chol[tid];
    float a[4][KU], ce[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float* ar = AE + ((int64_t)b * R + rq * 4 + r) * AEW;
#pragma unroll
        for (int k = 0; k < KU; ++k) a[r][k] = ar[k];
        ce[r] = ar[KU];
    }
    __syncthreads();
    const int per = (N + nchunk - 1) / nchunk;
    const int j0 = chunk * per, j1 = min(N, j0 + per);
    float m[4], l[4], acc[4][KU];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        m[r] = -CUDART_INF_F; l[r] = 0.f;
#pragma unroll
        for (int k = 0; k < KU; ++k) acc[r][k] = 0.f;
    }
    for (int j = j0 + ps; j < j1; j += 64) {
        float u[KU];
        load_u(x_t, xyz, (int64_t)b * N + j, true, u);
        const float rstd = rstd_chol(s_chol, u);
        float wr[KU];
#pragma unroll
        for (int k = 0; k < KU; ++k) wr[k] = rstd * u[k];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float s = ce[r];
#pragma unroll
            for (int k = 0; k < KU; ++k) s = fmaf(a[r][k], wr[k], s);
            if (s > m[r]) {  // rescale the running partial (expf(-inf) = 0 on the first point)
                const float f = expf(m[r] - s);
                l[r] *= f;
#pragma unroll
                for (int k = 0; k < KU; ++k) acc[r][k] *= f;
                m[r] = s;
            }
            const float p = expf(s - m[r]);
            l[r] += p;
#pragma unroll
            for (int k = 0; k < KU; ++k) acc[r][k] = fmaf(p, wr[k], acc[r][k]);
        }
    }
    // lanes with equal (lane & 3) hold partials of the same 4 rows: butterfly over lane bits 2..4
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float mo = __shfl_xor_sync(0xffffffffu, m[r], off), lo = __shfl_xor_sync(0xffffffffu, l[r], off);
            const float M = fmaxf(m[r], mo);
            const float fa = m[r] == -CUDART_INF_F ? 0.f : expf(m[r] - M), fb = mo == -CUDART_INF_F ? 0.f : expf(mo - M);
            l[r] = l[r] * fa + lo * fb;
#pragma unroll
            for (int k = 0; k < KU; ++k) {
                const float ao = __shfl_xor_sync(0xffffffffu, acc[r][k], off);
                acc[r][k] = acc[r][k] * fa + ao * fb;
            }
            m[r] = M;
        }
    }
    if (lane < 4) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float* d = s_red[warp][rq * 4 + r];
#pragma unroll
            for (int k = 0; k < KU; ++k) d[k] = acc[r][k];
            d[KU] = m[r]; d[KU + 1] = l[r];
        }
    }
    __syncthreads();
    if (tid < R) {
        float M = -CUDART_INF_F;
        for (int w = 0; w < 8; ++w) M = fmaxf(M, s_red[w][tid][KU]);
        float o[KU], ls = 0.f;
#pragma unroll
        for (int k = 0; k < KU; ++k) o[k] = 0.f;
        for (int w = 0; w < 8; ++w) {
            const float mw = s_red[w][tid][KU];
            const float f = mw == -CUDART_INF_F ? 0.f : expf(mw - M);
            ls = fmaf(f, s_red[w][tid][KU + 1], ls);
#pragma unroll
            for (int k = 0; k < KU; ++k) o[k] = fmaf(f, s_red[w][tid][k], o[k]);
        }
        float* d = part + (((int64_t)b * nchunk + chunk) * R + tid) * AEW;
#pragma unroll
        for (int k = 0; k < KU; ++k) d[k] = o[k];
        d[KU] = M; d[KU + 1] = ls;
    }
}

// z[b,r,:] = diag(g) Ec w_r + beta,  w_r = sum_chunks acc / sum_chunks l   (the softmax-weighted mean of LN(enc_kv) rows)
__global__ void __launch_bounds__(C)
cdm_enc_expand_kernel(const float* __restrict__ part, const float* __restrict__ ecg, const float* __restrict__ beta,
                      float* __restrict__ z, int nchunk) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float w[R][AEW];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid < R) {
        const float* base = part + ((int64_t)b * nchunk * R + tid) * AEW;
        float M = -CUDART_INF_F;
        for (int c = 0; c < nchunk; ++c) M = fmaxf(M, base[(int64_t)c * R * AEW + KU]);
        float o[KU], ls = 0.f;
#pragma unroll
        for (int k = 0; k < KU; ++k) o[k] = 0.f;
        for (int c = 0; c < nchunk; ++c) {
            const float* pc = base + (int64_t)c * R * AEW;
            const float f = pc[KU] == -CUDART_INF_F ? 0.f : expf(pc[KU] - M);
            ls = fmaf(f, pc[KU + 1], ls);
#pragma unroll
            for (int k = 0; k < KU; ++k) o[k] = fmaf(f, pc[k], o[k]);
        }
        const float inv = 1.0f / ls;
#pragma unroll
        for (int k = 0; k < KU; ++k) w[tid][k] = o[k] * inv;
    }
    __syncthreads();
    float e[KU];
#pragma unroll
    for (int k = 0; k < KU; ++k) e[k] = ecg[tid * KU + k];
    const float bb = beta[tid];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float v = bb;
#pragma unroll
        for (int k = 0; k < KU; ++k) v = fmaf(e[k], w[r][k], v);
        z[((int64_t)b * R + r) * C + tid] = v;
    }
}

// ------------------------------------------------------------------------------------------------ decoder, per-sample prep
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// One CTA per sample.  UU [B][R][NS]: columns [0,C) centred U, [C,2C) MPt = Uc W1g^T, [2C,2C+KU) G1up, [2C+KU,2C+KU+J) HPt
// (one batched o_proj-stack launch, amb200/cdm_fold.py d_ostack).  Writes the parameter block and the swizzled bf16 (hi|lo)
// B operand [256 x 32] of the per-point GEMM.
__global__ void __launch_bounds__(256)
cdm_dec_prep_kernel(const float* __restrict__ AQ, const float* __restrict__ UU, int NS, const float* __restrict__ g1uu,
                    const float* __restrict__ mu, const float* __restrict__ hu, float* __restrict__ PB, uint8_t* __restrict__ blob) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float Uc[R][C + 1];
    __shared__ float G[KZ][KZ + 1];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* uu = UU + (int64_t)b * R * NS;
    for (int i = tid; i < R * C; i += 256) Uc[i / C][i % C] = uu[(i / C) * NS + (i % C)];
    for (int i = tid; i < KU * KU; i += 256) G[i / KU][i % KU] = g1uu[i];
    for (int i = tid; i < R * KU; i += 256) {
        const int r = i / KU, k = i % KU;
        const float v = uu[r * NS + 2 * C + k];
        G[KU + r][k] = v; G[k][KU + r] = v;
    }
    __syncthreads();
    {
        const int r = tid >> 4, s = tid & 15;
        float d = 0.f;
        for (int c = 0; c < C; ++c) d = fmaf(Uc[r][c], Uc[s][c], d);
        G[KU + r][KU + s] = d * (1.0f / C);
    }
    __syncthreads();
    float* pb = PB + (int64_t)b * PB_STRIDE;
    for (int i = tid; i < R * AEW; i += 256) pb[PB_AQ + i] = AQ[(int64_t)b * R * AEW + i];
    for (int i = tid; i < KZ * KZ; i += 256) {
        const int ii = i / KZ, jj = i % KZ;
        if (jj >= ii) pb[PB_GT + ii * KZ - (ii * (ii - 1)) / 2 + (jj - ii)] = (ii == jj ? 1.0f : 2.0f) * G[ii][jj];
    }
    for (int i = tid; i < J * HPW; i += 256) {
        const int o = i / HPW, k = i % HPW;
        pb[PB_HP + i] = k < KU ? hu[o * KU + k] : (k < KZ ? uu[(k - KU) * NS + 2 * C + KU + o] : 0.f);
    }
    {   // row n of the GEMM operand M = [Mu | MPt^T | 0]
        const int n = tid;
        float v[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) v[k] = k < KU ? mu[n * KU + k] : (k < KZ ? uu[(k - KU) * NS + C + n] : 0.f);
        uint32_t hi[KP / 2], lo[KP / 2];
#pragma unroll
        for (int k = 0; k < KP; k += 2) {
            const uint32_t h = pack_bf16(v[k], v[k + 1]);
            hi[k / 2] = h;
            lo[k / 2] = pack_bf16(v[k] - __uint_as_float(h << 16), v[k + 1] - __uint_as_float(h & 0xffff0000u));
        }
        uint8_t* bh = blob + (int64_t)b * BLOB_BYTES;
        uint8_t* bl = bh + C * KP * 2;
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // 64-byte rows, 16-byte chunk c stored at chunk c ^ ((row >> 1) & 3)   (SWIZZLE_64B)
            const int off = n * 64 + ((c ^ ((n >> 1) & 3)) << 4);
            *reinterpret_cast<uint4*>(bh + off) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
            *reinterpret_cast<uint4*>(bl + off) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ decoder, tcgen05 point kernel
// Exact-erf GELU x * Phi(x) with Phi from the Abramowitz-Stegun 7.1.26 rational form of erfc (|abs error| <= 1.5e-7 in erf, i.e.
// <= 1e-7 |x| in the GELU: three orders below the 1e-3 parity budget and below the bf16-split noise of the GEMM feeding it).
// ~17 instructions per element against ~35 for erff(): the per-point GELU is the dominant instruction stream of this kernel.
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float gelu_as(float x) {
    const float ax = fabsf(x) * 0.70710678118654752440f;
    const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));
    float y = fmaf(1.061405429f, t, -1.453152027f);
    y = fmaf(y, t, 1.421413741f);
    y = fmaf(y, t, -0.284496736f);
    y = fmaf(y, t, 0.254829592f);
    const float h = 0.5f * (y * t) * ex2_approx(ax * ax * -1.4426950408889634f);  // 0.5 erfc(|x| / sqrt 2)
    return x * (x >= 0.f ? 1.0f - h : h);
}
__device__ __forceinline__ uint32_t smem_u32p(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_p(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32p(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait_p(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32p(bar)), "r"(parity) : "memory");
        if (ok) break;
        if (++spins > 200000000u) __trap();  // protocol bug: fail the launch instead of hanging the box
    }
}
__device__ __forceinline__ void umma_ss_p(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"   // one lane of a CONVERGENT warp (a divergent `tid == 0` makes ptxas loop over the active lanes)
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_p(uint64_t* bar) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
                 ::"r"(smem_u32p(bar)) : "memory");
}
__device__ __forceinline__ void tc_before_p() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_after_p() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16_p(uint32_t taddr, uint32_t r[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major operand tile, 64-byte rows (32 bf16), SWIZZLE_64B, 8-row groups 512 B apart (same encoding as gemm_tc.cu make_desc<32>)
__device__ __forceinline__ uint64_t make_desc64_p(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
// c_format F32 | a,b BF16 | K-major A/B | N = 256 | M = 128
constexpr uint32_t IDESC_DEC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);

constexpr int DEC_THREADS = 256;
constexpr int SO_AH = 0, SO_AL = 8192, SO_BH = 16384, SO_BL = 32768, SO_PB = 49152, SO_C1 = SO_PB + PB_STRIDE * 4,
              SO_WG = SO_C1 + C * 4, SO_CHOL = SO_WG + C * 8 * 4, SO_RED = SO_CHOL + 256, SO_BAR = SO_RED + TILE * 8 * 4,
              SO_END = SO_BAR + 16;
constexpr int DEC_SMEM = 80 * 1024;  // > 227 KB / 3: at most two CTAs (2 x 256 TMEM columns) share an SM
static_assert(SO_END + 1024 <= DEC_SMEM, "decoder shared memory layout");

__global__ void __launch_bounds__(DEC_THREADS, 2)
cdm_dec_points_tc_kernel(const float* __restrict__ x_t, const float* __restrict__ xyz, const float* __restrict__ chol,
                         const float* __restrict__ c1, const float* __restrict__ wg, const float* __restrict__ PB,
                         const uint8_t* __restrict__ blob, float* __restrict__ out, int N, int tiles_per_sample, int tiles_per_cta,
                         int total_tiles) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the shared-space pointer (an integer round-trip would demote every later
    // access to generic LD/ST)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32p(smem_raw) & 1023u)) & 1023u);
    float* s_pb = reinterpret_cast<float*>(smem + SO_PB);
    float* s_c1 = reinterpret_cast<float*>(smem + SO_C1);
    float* s_wg = reinterpret_cast<float*>(smem + SO_WG);
    float* s_chol = reinterpret_cast<float*>(smem + SO_CHOL);
    float* s_red = reinterpret_cast<float*>(smem + SO_RED);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SO_BAR);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + SO_BAR + 8);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int pt = tid & (TILE - 1), half = tid >> 7;

    if (tid == 0) {
        mbar_init_p(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32p(tmem_holder)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_launch_dependents();
    pdl_wait();
    for (int i = tid; i < NCHOL; i += DEC_THREADS) s_chol[i] = chol[i];
    for (int i = tid; i < C; i += DEC_THREADS) s_c1[i] = c1[i];
    for (int i = tid; i < C * 8 / 4; i += DEC_THREADS) reinterpret_cast<float4*>(s_wg)[i] = __ldg(reinterpret_cast<const float4*>(wg) + i);
    tc_before_p();
    __syncthreads();
    tc_after_p();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);

    const int t0 = blockIdx.x * tiles_per_cta, t1 = min(total_tiles, t0 + tiles_per_cta);
    int cur_b = -1;
    uint32_t phase = 0;
    for (int tile = t0; tile < t1; ++tile) {
        const int b = tile / tiles_per_sample;
        const int j = (tile - b * tiles_per_sample) * TILE + pt;
        const bool valid = j < N;
        if (b != cur_b) {  // per-sample operands: GEMM B tile (32 KB) + parameter block (3 KB)
            __syncthreads();
            const uint4* src = reinterpret_cast<const uint4*>(blob + (int64_t)b * BLOB_BYTES);
            uint4* dst = reinterpret_cast<uint4*>(smem + SO_BH);
#pragma unroll
            for (int i = 0; i < BLOB_BYTES / 16 / DEC_THREADS; ++i) dst[tid + i * DEC_THREADS] = __ldg(src + tid + i * DEC_THREADS);
            const float4* ps = reinterpret_cast<const float4*>(PB + (int64_t)b * PB_STRIDE);
            if (tid < PB_STRIDE / 4) reinterpret_cast<float4*>(s_pb)[tid] = __ldg(ps + tid);
            __syncthreads();
            cur_b = b;
        }
        // ---------------- phase 1: everything that depends on the 9 input floats only
        float z[KZ];
        load_u(x_t, xyz, (int64_t)b * N + j, valid, z);
        const float rq = rstd_chol(s_chol, z);
        {
            float s[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float* ar = s_pb + PB_AQ + r * AEW;
                float d = 0.f;
#pragma unroll
                for (int k = 0; k < KU; ++k) d = fmaf(ar[k], z[k], d);
                s[r] = fmaf(rq, d, ar[KU]);
            }
#pragma unroll
            for (int h = 0; h < R / 2; ++h) {  // softmax over the 2 latents of head h (rows 2h, 2h+1)
                const float mx = fmaxf(s[2 * h], s[2 * h + 1]);
                const float e0 = expf(s[2 * h] - mx), e1 = expf(s[2 * h + 1] - mx);
                const float inv = rcp_approx(e0 + e1);
                z[KU + 2 * h] = e0 * inv; z[KU + 2 * h + 1] = e1 * inv;
            }
        }
        float r1;
        {
            const float* gt = s_pb + PB_GT;
            float var = 0.f;
            int idx = 0;
#pragma unroll
            for (int i = 0; i < KZ; ++i) {
                float t = 0.f;
#pragma unroll
                for (int k = i; k < KZ; ++k) t = fmaf(gt[idx++], z[k], t);
                var = fmaf(z[i], t, var);
            }
            r1 = rsqrtf(fmaxf(var, 0.f) + 1e-5f);
        }
        float oh[J];
        if (half == 0) {
#pragma unroll
            for (int o = 0; o < J; ++o) {
                const float* hp = s_pb + PB_HP + o * HPW;
                float d = 0.f;
#pragma unroll
                for (int k = 0; k < KZ; ++k) d = fmaf(hp[k], z[k], d);
                oh[o] = d;
            }
            // A operand: row pt = bf16 (hi | lo) of z, zero padded to 32
            uint32_t hi[KP / 2], lo[KP / 2];
#pragma unroll
            for (int k = 0; k < KP; k += 2) {
                const float a = k < KZ ? z[k < KZ ? k : 0] : 0.f, c = k + 1 < KZ ? z[k + 1 < KZ ? k + 1 : 0] : 0.f;
                const uint32_t h = pack_bf16(a, c);
                hi[k / 2] = h;
                lo[k / 2] = pack_bf16(a - __uint_as_float(h << 16), c - __uint_as_float(h & 0xffff0000u));
            }
            const uint32_t ah = smem_u32p(smem + SO_AH) + pt * 64, al = smem_u32p(smem + SO_AL) + pt * 64;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t off = (uint32_t)((c ^ ((pt >> 1) & 3)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ah + off), "r"(hi[4 * c]), "r"(hi[4 * c + 1]), "r"(hi[4 * c + 2]), "r"(hi[4 * c + 3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(al + off), "r"(lo[4 * c]), "r"(lo[4 * c + 1]), "r"(lo[4 * c + 2]), "r"(lo[4 * c + 3]) : "memory");
            }
        } else {
#pragma unroll
            for (int o = 0; o < J; ++o) oh[o] = 0.f;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes (A tile, B tile) -> tensor-core reads
        tc_before_p();
        __syncthreads();
        // ---------------- the dense layer: D[128 x 256] = Z[128 x 32] M^T, 3-term split, one elected thread
        if (tid < 32) {
            tc_after_p();
            const uint64_t dah = make_desc64_p(smem_u32p(smem + SO_AH)), dal = make_desc64_p(smem_u32p(smem + SO_AL));
            const uint64_t dbh = make_desc64_p(smem_u32p(smem + SO_BH)), dbl = make_desc64_p(smem_u32p(smem + SO_BL));
#pragma unroll
            for (int k = 0; k < KP / 16; ++k) {  // K = 16 per MMA: +32 B inside the 64 B swizzle row (encoded +2)
                const uint64_t ko = (uint64_t)(k * 2);
                umma_ss_p(tmem_base, dal + ko, dbh + ko, IDESC_DEC, k ? 1u : 0u);
                umma_ss_p(tmem_base, dah + ko, dbl + ko, IDESC_DEC, 1u);
                umma_ss_p(tmem_base, dah + ko, dbh + ko, IDESC_DEC, 1u);
            }
            umma_commit_p(bar);
        }
        mbar_wait_p(bar, phase);
        phase ^= 1u;
        tc_after_p();
        // ---------------- phase 2: GELU + folded head straight out of TMEM; this thread owns row pt, columns [128 half, +128)
        float o[J];
#pragma unroll
        for (int q = 0; q < J; ++q) o[q] = 0.f;
#pragma unroll 1
        for (int ch = 0; ch < 8; ++ch) {
            const int col0 = half * 128 + ch * 16;
            uint32_t v[16];
            tmem_ld16_p(tmem_row + (uint32_t)col0, v);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float g = gelu_as(fmaf(r1, __uint_as_float(v[q]), s_c1[col0 + q]));
                const float4 w0 = *reinterpret_cast<const float4*>(s_wg + (col0 + q) * 8);
                const float2 w1 = *reinterpret_cast<const float2*>(s_wg + (col0 + q) * 8 + 4);
                o[0] = fmaf(g, w0.x, o[0]); o[1] = fmaf(g, w0.y, o[1]); o[2] = fmaf(g, w0.z, o[2]);
                o[3] = fmaf(g, w0.w, o[3]); o[4] = fmaf(g, w1.x, o[4]); o[5] = fmaf(g, w1.y, o[5]);
            }
        }
        if (half == 1) {
            *reinterpret_cast<float4*>(s_red + pt * 8) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float2*>(s_red + pt * 8 + 4) = make_float2(o[4], o[5]);
        }
        tc_before_p();
        __syncthreads();
        if (half == 0 && valid) {
            const float4 r0 = *reinterpret_cast<const float4*>(s_red + pt * 8);
            const float2 r2 = *reinterpret_cast<const float2*>(s_red + pt * 8 + 4);
            float2* dst = reinterpret_cast<float2*>(out + ((int64_t)b * N + j) * J);
            dst[0] = make_float2(o[0] + r0.x + oh[0], o[1] + r0.y + oh[1]);
            dst[1] = make_float2(o[2] + r0.z + oh[2], o[3] + r0.w + oh[3]);
            dst[2] = make_float2(o[4] + r2.x + oh[4], o[5] + r2.y + oh[5]);
        }
    }
    tc_before_p();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

}  // namespace

extern "C" int am_cdm_enc_points(const float* x_t, const float* xyz, const float* chol, const float* AE, float* part, int B, int N,
                                 int nchunk, am_stream_t stream) {
    AM_REQUIRE(x_t && xyz && chol && AE && part, AM_EINVAL, "am_cdm_enc_points: null pointer");
    AM_REQUIRE(B > 0 && N > 0 && nchunk > 0 && nchunk <= 65535, AM_EINVAL, "am_cdm_enc_points: bad dims");
    AM_REQUIRE((reinterpret_cast<uintptr_t>(x_t) & 7u) == 0, AM_EALIGN, "am_cdm_enc_points: x_t must be 8-byte aligned");
    am_launch(cdm_enc_points_kernel, dim3(nchunk, B), dim3(256), 0, as_stream(stream), 1, x_t, xyz, chol, AE, part, N, nchunk);
    AM_LAUNCH_CHECK("cdm_enc_points");
    return AM_OK;
}

extern "C" int am_cdm_enc_expand(const float* part, const float* ecg, const float* beta, float* z, int B, int nchunk, am_stream_t stream) {
    AM_REQUIRE(part && ecg && beta && z && B > 0 && nchunk > 0, AM_EINVAL, "am_cdm_enc_expand: bad args");
    am_launch(cdm_enc_expand_kernel, dim3(B), dim3(C), 0, as_stream(stream), 1, part, ecg, beta, z, nchunk);
    AM_LAUNCH_CHECK("cdm_enc_expand");
    return AM_OK;
}

extern "C" int am_cdm_dec_prep(const float* AQ, const float* UU, int NS, const float* g1uu, const float* mu, const float* hu, float* PB,
                               void* blob, int B, am_stream_t stream) {
    AM_REQUIRE(AQ && UU && g1uu && mu && hu && PB && blob && B > 0, AM_EINVAL, "am_cdm_dec_prep: bad args");
    AM_REQUIRE(NS >= 2 * C + KU + J, AM_EINVAL, "am_cdm_dec_prep: o_proj stack is too narrow");
    AM_REQUIRE((reinterpret_cast<uintptr_t>(blob) & 15u) == 0 && (reinterpret_cast<uintptr_t>(PB) & 15u) == 0, AM_EALIGN,
               "am_cdm_dec_prep: PB / blob must be 16-byte aligned");
    am_launch(cdm_dec_prep_kernel, dim3(B), dim3(256), 0, as_stream(stream), 1, AQ, UU, NS, g1uu, mu, hu, PB, reinterpret_cast<uint8_t*>(blob));
    AM_LAUNCH_CHECK("cdm_dec_prep");
    return AM_OK;
}

extern "C" int am_cdm_dec_points_tc(const float* x_t, const float* xyz, const float* chol, const float* c1, const float* wg,
                                    const float* PB, const void* blob, float* out, int B, int N, am_stream_t stream) {
    AM_REQUIRE(x_t && xyz && chol && c1 && wg && PB && blob && out, AM_EINVAL, "am_cdm_dec_points_tc: null pointer");
    AM_REQUIRE(B > 0 && N > 0, AM_EINVAL, "am_cdm_dec_points_tc: bad dims");
    AM_REQUIRE((reinterpret_cast<uintptr_t>(x_t) & 7u) == 0 && (reinterpret_cast<uintptr_t>(out) & 7u) == 0 &&
                   (reinterpret_cast<uintptr_t>(wg) & 15u) == 0 && (reinterpret_cast<uintptr_t>(blob) & 15u) == 0 &&
                   (reinterpret_cast<uintptr_t>(PB) & 15u) == 0,
               AM_EALIGN, "am_cdm_dec_points_tc: alignment (x_t/out 8 B, wg/PB/blob 16 B)");
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(cdm_dec_points_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DEC_SMEM) != cudaSuccess) {
            am_set_error_("am_cdm_dec_points_tc: shared memory opt-in failed");
            return AM_ELAUNCH;
        }
        attr_set = true;
    }
    const int tps = cdiv(N, TILE);
    const int64_t total64 = (int64_t)B * tps;
    AM_REQUIRE(total64 < (1ll << 30), AM_EINVAL, "am_cdm_dec_points_tc: too many tiles");
    const int total = (int)total64;
    const int max_ctas = 2 * am_num_sms();
    const int per = cdiv(total, total < max_ctas ? total : max_ctas);
    const int grid = cdiv(total, per);
    am_launch(cdm_dec_points_tc_kernel, dim3(grid), dim3(DEC_THREADS), DEC_SMEM, as_stream(stream), 1, x_t, xyz, chol, c1, wg, PB,
              reinterpret_cast<const uint8_t*>(blob), out, N, tps, per, total);
    AM_LAUNCH_CHECK("cdm_dec_points_tc");
    return AM_OK;
}
