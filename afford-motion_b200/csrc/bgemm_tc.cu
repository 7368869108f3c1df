// Batched small GEMMs of the training-time attention (amb200/autograd_ops.py AttentionFn / MHAFn: S = Q K^T, O = P V and their four
// backward products per (batch, head); models/cmdm.py:66-77 in train mode) on the tcgen05 tensor cores with fp32 operands in HBM.
//
//   C[b] (M x N, fp32) = alpha * op(A[b]) (M x K) * op(B[b]) (K x N) (+ beta * C[b])
//
// The round-1 path ran these as fp32 SIMT GEMMs (gemm_general_kernel: 30 launches, ~6 of the 45 ms training step).  Here the operands
// are converted on the fly: every thread reads 8 consecutive fp32 values along the operand's CONTIGUOUS dimension, splits them into
// bf16 (hi | lo) and stores two 16-byte chunks into a SWIZZLE_128B tile —
//   * contiguous along K  -> K-major tile   (rows = M / N index, 64 k per 128-byte row)
//   * contiguous along M/N -> MN-major tile (rows = k, 64 m/n per 128-byte row; 64-wide atoms LBO apart)
// so no operand is ever transposed: the UMMA descriptors (a_major / b_major bits of the instruction descriptor) take both.  Per
// 64-wide K block one thread issues 4 x 3 tcgen05.mma (M128, N = BN, K16; 3-term split = fp32-equivalent), the accumulator stays in
// TMEM across K blocks, and the epilogue goes TMEM -> shared (row per lane) -> coalesced global rows.
// One output tile per CTA, no intra-CTA pipeline: 3 CTAs (3 x 128 TMEM columns, 3 x 66 KB) share an SM and overlap each other.
#include <cuda_bf16.h>
#include <string.h>
#include "common.cuh"

namespace {

constexpr int TM = 128, BK = 64, BG_THREADS = 256;

struct BGemm {
    const float* A; const float* B; float* C;
    int M, N, K;
    int64_t sAm, sAk, sBn, sBk;   // element strides of A(m,k) and B(n,k); exactly one of each pair is 1
    int ldc;
    float alpha, beta;
    int bdiv; int64_t sA1, sA2, sB1, sB2, sC1, sC2;  // batch bi -> (bi / bdiv, bi % bdiv) offsets
    int a_mn, b_mn;               // 1: operand contiguous along M / N (MN-major tile), 0: along K
};

__device__ __forceinline__ uint32_t smem_u32b(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32b(bar)), "r"(parity) : "memory");
        if (ok) break;
        if (++spins > 200000000u) __trap();  // protocol bug: fail the launch instead of hanging the box
    }
}
__device__ __forceinline__ void umma_b(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"   // issued by one lane of a CONVERGENT warp (a divergent `tid == 0` makes ptxas loop over lanes)
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld16_b(uint32_t taddr, uint32_t r[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pack2_b(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
// SWIZZLE_128B tiles, 8-row groups 1024 B apart.  K-major: rows = M/N index (LBO unused).  MN-major: rows = k, 64-wide atoms LBO apart.
__device__ __forceinline__ uint64_t desc_b(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes ? (lbo_bytes >> 4) : 1u) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// Stage one operand tile: ROWS index values (m or n) x 64 k.  8 consecutive fp32 along the contiguous dimension per item -> bf16 (hi | lo).
template <int ROWS>
__device__ __forceinline__ void stage_operand(const float* __restrict__ P, int64_t s_r, int64_t s_k, bool mn_major, int r0, int rmax, int k0,
                                              int kmax, uint32_t hi_base, uint32_t lo_base, int tid) {
    constexpr int ITEMS = ROWS * BK / 8;
#pragma unroll
    for (int it = 0; it < ITEMS / BG_THREADS; ++it) {
        const int item = tid + it * BG_THREADS;
        float v[8];
        uint32_t off;
        if (!mn_major) {   // K contiguous: item = (row, chunk of 8 k)
            const int row = item >> 3, c = item & 7;
            const int gr = r0 + row, gk = k0 + c * 8;
            const float* src = P + (int64_t)gr * s_r + gk;
            const bool rok = gr < rmax;
            if (rok && gk + 7 < kmax && ((reinterpret_cast<uintptr_t>(src) & 7u) == 0)) {
                const float2 a = __ldg(reinterpret_cast<const float2*>(src)), b = __ldg(reinterpret_cast<const float2*>(src) + 1);
                const float2 c2 = __ldg(reinterpret_cast<const float2*>(src) + 2), d = __ldg(reinterpret_cast<const float2*>(src) + 3);
                v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c2.x; v[5] = c2.y; v[6] = d.x; v[7] = d.y;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = (rok && gk + e < kmax) ? __ldg(src + e) : 0.f;
            }
            off = (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4));
        } else {           // M / N contiguous: item = (k row, chunk of 8 m/n); 64-wide atoms of 8 KB
            constexpr int CH = ROWS / 8;          // chunks per k row
            const int krow = item / CH, cc = item % CH;
            const int gk = k0 + krow, gr = r0 + cc * 8;
            const float* src = P + (int64_t)gk * s_k + gr;
            const bool kok = gk < kmax;
            if (kok && gr + 7 < rmax && ((reinterpret_cast<uintptr_t>(src) & 7u) == 0)) {
                const float2 a = __ldg(reinterpret_cast<const float2*>(src)), b = __ldg(reinterpret_cast<const float2*>(src) + 1);
                const float2 c2 = __ldg(reinterpret_cast<const float2*>(src) + 2), d = __ldg(reinterpret_cast<const float2*>(src) + 3);
                v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c2.x; v[5] = c2.y; v[6] = d.x; v[7] = d.y;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = (kok && gr + e < rmax) ? __ldg(src + e) : 0.f;
            }
            const int atom = cc >> 3, c = cc & 7;
            off = (uint32_t)(atom * 8192 + krow * 128 + ((c ^ (krow & 7)) << 4));
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            const uint32_t h = pack2_b(v[e], v[e + 1]);
            hi[e / 2] = h;
            lo[e / 2] = pack2_b(v[e] - __uint_as_float(h << 16), v[e + 1] - __uint_as_float(h & 0xffff0000u));
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hi_base + off), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lo_base + off), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
    }
}

template <int BN>
__global__ void __launch_bounds__(BG_THREADS, 2)
bgemm_tc_kernel(BGemm g) {
    constexpr int A_BYTES = TM * BK * 2, B_BYTES = BN * BK * 2;      // one bf16 tile (hi or lo)
    constexpr int OP_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    constexpr int CS_LD = BN + 1;                                     // epilogue staging row stride (floats)
    static_assert(TM * CS_LD * 4 <= OP_BYTES + 1024, "epilogue staging must fit in the operand area");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32b(smem_raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OP_BYTES + 1024);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + OP_BYTES + 1024 + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bi = blockIdx.z;
    const float* A = g.A + (bi / g.bdiv) * g.sA1 + (bi % g.bdiv) * g.sA2;
    const float* B = g.B + (bi / g.bdiv) * g.sB1 + (bi % g.bdiv) * g.sB2;
    float* C = g.C + (bi / g.bdiv) * g.sC1 + (bi % g.bdiv) * g.sC2;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * BN;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32b(bar)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32b(tmem_holder)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t a_hi = smem_u32b(smem), a_lo = a_hi + A_BYTES, b_hi = a_lo + A_BYTES, b_lo = b_hi + B_BYTES;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)g.a_mn << 15) | ((uint32_t)g.b_mn << 16) | ((uint32_t)(BN >> 3) << 17) |
                           ((uint32_t)(TM >> 4) << 24);
    uint32_t phase = 0;
    const int nkb = (g.K + BK - 1) / BK;
    for (int kb = 0; kb < nkb; ++kb) {
        const int k0 = kb * BK;
        stage_operand<TM>(A, g.sAm, g.sAk, g.a_mn != 0, m0, g.M, k0, g.K, a_hi, a_lo, tid);
        stage_operand<BN>(B, g.sBn, g.sBk, g.b_mn != 0, n0, g.N, k0, g.K, b_hi, b_lo, tid);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> tensor-core reads
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t dah = desc_b(a_hi, g.a_mn ? 8192u : 0u), dal = desc_b(a_lo, g.a_mn ? 8192u : 0u);
            const uint64_t dbh = desc_b(b_hi, g.b_mn ? 8192u : 0u), dbl = desc_b(b_lo, g.b_mn ? 8192u : 0u);
            const uint64_t ka = g.a_mn ? (2048u >> 4) : (32u >> 4), kbs = g.b_mn ? (2048u >> 4) : (32u >> 4);  // start-address step per K16
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
                umma_b(tmem_base, dal + k * ka, dbh + k * kbs, idesc, (kb | k) ? 1u : 0u);
                umma_b(tmem_base, dah + k * ka, dbl + k * kbs, idesc, 1u);
                umma_b(tmem_base, dah + k * ka, dbh + k * kbs, idesc, 1u);
            }
            asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
                         ::"r"(smem_u32b(bar)) : "memory");
        }
        mbar_wait_b(bar, phase);   // the MMAs have finished reading this K block's tiles (and, on the last block, writing the accumulator)
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // epilogue: TMEM (row per lane) -> shared staging -> coalesced rows.  8 warps: lane quadrant warp & 3, column half warp >> 2
    float* Cs = reinterpret_cast<float*>(smem);
    {
        const int q = warp & 3, half = warp >> 2;
        const int row = q * 32 + lane;
#pragma unroll
        for (int j = 0; j < BN / 32; ++j) {
            const int col = half * (BN / 2) + j * 16;
            uint32_t v[16];
            tmem_ld16_b(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col, v);
#pragma unroll
            for (int e = 0; e < 16; ++e) Cs[row * CS_LD + col + e] = __uint_as_float(v[e]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    for (int i = tid; i < TM * BN; i += BG_THREADS) {
        const int row = i / BN, col = i - row * BN;
        const int m = m0 + row, n = n0 + col;
        if (m < g.M && n < g.N) {
            float* c = C + (int64_t)m * g.ldc + n;
            float v = g.alpha * Cs[row * CS_LD + col];
            if (g.beta != 0.f) v += g.beta * *c;
            *c = v;
        }
    }
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
}

template <int BN>
int launch_bgemm(const BGemm& g, int batch, cudaStream_t st) {
    constexpr int SMEM = 2 * (TM * BK * 2) + 2 * (BN * BK * 2) + 1024 /*align*/ + 1024 /*staging overhang*/ + 64;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(bgemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return -1;
        attr = true;
    }
    dim3 grid(cdiv(g.N, BN), cdiv(g.M, TM), batch);
    bgemm_tc_kernel<BN><<<grid, BG_THREADS, SMEM, st>>>(g);
    return 0;
}

}  // namespace

// Called by am_gemm_f32 (csrc/train_kernels.cu) for batched shapes; returns 1 if it launched, 0 if the shape is not supported here.
int am_bgemm_tc_(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda, const float* B, int ldb, float beta, float* C,
                 int ldc, int batch, int bdiv, int64_t sA1, int64_t sA2, int64_t sB1, int64_t sB2, int64_t sC1, int64_t sC2, cudaStream_t st) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("AMB200_BGEMM"); on = (e && !strcmp(e, "simt")) ? 0 : 1; }
    // batched attention products, and the un-batched tall GEMMs with odd / short K (TransitionDown linears, K = 35 / 67 / 131: they miss
    // am_linear_tc's K % 4 rule); deep-K weight gradients stay on the split-K SIMT path
    const bool batched = batch >= 8 && M >= 32;
    const bool tall = batch == 1 && M >= 4096 && K <= 512;
    if (!on || !(batched || tall) || N < 32 || K < 16) return 0;
    BGemm g;
    g.A = A; g.B = B; g.C = C; g.M = M; g.N = N; g.K = K;
    // am_gemm_f32 convention: A is [M,K] (lda) or, transA, [K,M]; B is [K,N] (ldb) or, transB, [N,K]
    g.a_mn = transA ? 1 : 0; g.sAm = transA ? 1 : lda; g.sAk = transA ? lda : 1;
    g.b_mn = transB ? 0 : 1; g.sBn = transB ? ldb : 1; g.sBk = transB ? 1 : ldb;
    g.ldc = ldc; g.alpha = alpha; g.beta = beta; g.bdiv = bdiv;
    g.sA1 = sA1; g.sA2 = sA2; g.sB1 = sB1; g.sB2 = sB2; g.sC1 = sC1; g.sC2 = sC2;
    const int rc = N > 64 ? launch_bgemm<128>(g, batch, st) : launch_bgemm<64>(g, batch, st);
    return rc == 0 ? 1 : 0;
}
