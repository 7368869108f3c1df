// tcgen05 GEMM with the residual add and the LayerNorm fused into the epilogue (CMDM trunk: out_proj + norm1, linear2 + norm2 of
// torch.nn.TransformerEncoderLayer, models/cmdm.py:66-77; post-LN: x = LN(x + sublayer(x))).
//
//   Y2[M, 2*512] (bf16 hi | lo) = split( LN( A[M,K] W[512,K]^T + bias + (R_hi + R_lo) ) * gamma + beta )
//
// Why a separate kernel: LayerNorm needs the whole 512-column row, i.e. BOTH 256-wide accumulator tiles of a row block in one CTA's
// tensor memory.  A CTA pair (tcgen05.mma.cta_group::2, UMMA M = 256) owns 256 rows; each CTA keeps its 128 rows x 512 columns of
// fp32 accumulators in all 512 TMEM columns (n-tile 0 in columns [0,256), n-tile 1 in [256,512)).  The epilogue makes three passes
// over TMEM, one accumulator row per 4 threads (four warps share a row: 128 columns each):
//   pass 1  x = acc + bias + residual (bf16 pair read from global), written BACK into the accumulator columns (tcgen05.st); row sum
//           (starts on n-tile 0 while the tensor cores are still working on n-tile 1)
//   pass 2  sum of (x - mean)^2                    (two-pass variance, same arithmetic as the stand-alone LayerNorm kernel)
//   pass 3  y = (x - mean) * rstd * gamma + beta -> bf16 (hi | lo) -> global
// This removes the fp32 [M,512] hand-off tensor (written by the GEMM, read by the LayerNorm: 42 MB per layer half) and one launch per
// LayerNorm (10 per denoise step).  3-term bf16 split / fast mode as in gemm_tc.cu; BK = 64 (SWIZZLE_128B), 3-stage TMA ring.
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>
#include <type_traits>
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 256, NOUT = 512, BK = 64, NST = 3;
constexpr int EPI_WARPS = 16;  // 4 TMEM lane quadrants x 4 column quarters: the epilogue is latency-bound (TMEM round trips, one residual
                               // row segment per lane), so it gets twice the warps of the plain GEMM kernels
constexpr int THREADS = (EPI_WARPS + 4) * 32;  // warps 0-15 epilogue, 16 TMEM allocator, 18 TMA producer, 19 MMA issuer
constexpr int ALLOC_WARP = 16, PRODUCER_WARP = 18, MMA_WARP = 19;
constexpr int A_SUB = BM * BK * 2;          // 16 KB: this CTA's 128 rows of A (hi or lo)
constexpr int B_SUB = (BN / 2) * BK * 2;    // 16 KB: this CTA's half of the 256-wide W tile (hi or lo)
constexpr int STAGE_BYTES = 2 * A_SUB + 2 * B_SUB;
constexpr int SMEM_BYTES = NST * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 4 * BM * 4 /*row partials*/;
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);

struct LnParams {
    int M, Kp;
    const float* bias;
    const __nv_bfloat16* R2; int ldr;     // residual (hi | lo) [M, 2*ldr]
    const float *gamma, *beta; float eps;
    __nv_bfloat16* Y2;                    // [M, 2*NOUT]
    int fast;
};

__device__ __forceinline__ uint32_t smem_u32g(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_g(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32g(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait_g(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32g(bar)), "r"(parity) : "memory");
        if (ok) break;
        if (++spins > 200000000u) __trap();  // protocol bug: fail the launch instead of hanging the box
    }
}
__device__ __forceinline__ void tc_before_g() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_after_g() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16_g(uint32_t taddr, uint32_t r[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16_g(uint32_t taddr, const uint32_t r[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// K-major operand tile, 128-byte rows (64 bf16), SWIZZLE_128B, 8-row groups 1024 B apart (gemm_tc.cu make_desc<64>)
__device__ __forceinline__ uint64_t make_desc128_g(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t pack2_g(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_ln_2sm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, LnParams p) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32g(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NST * STAGE_BYTES);
    uint64_t* empty = full + NST;
    uint64_t* tmem_full = empty + NST;  // [2]: one per n-tile
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full + 2);
    float* xch = reinterpret_cast<float*>(smem + NST * STAGE_BYTES + 256);  // [4 column quarters][128 rows]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.Kp / BK;
    uint32_t cta_rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    const bool leader = cta_rank == 0;
    const int item = blockIdx.x >> 1;                     // one 256-row block per CTA pair
    const int m0 = (item * 2 + (int)cta_rank) * BM;       // this CTA's 128 rows

    if (warp == PRODUCER_WARP && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    }
    if (warp == MMA_WARP && lane == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init_g(&full[s], 1); mbar_init_g(&empty[s], 1); }
        mbar_init_g(&tmem_full[0], 1); mbar_init_g(&tmem_full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == ALLOC_WARP) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32g(tmem_holder)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_before_g();
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    tc_after_g();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();  // prologue above overlapped the previous kernel; from here on its results are complete and visible

    if (warp == PRODUCER_WARP) {
        // ===================== TMA producer (both CTAs): own 128 A rows + own half of each 256-wide W tile =====================
        if (lane == 0) {
            auto produce = [&](auto fc) {
                constexpr bool FAST = decltype(fc)::value;
                constexpr uint32_t stage_tx = FAST ? (uint32_t)(A_SUB + B_SUB) : (uint32_t)STAGE_BYTES;
                uint32_t kbc = 0;
                for (int nt = 0; nt < 2; ++nt) {
                    const int nh = nt * BN + (int)cta_rank * (BN / 2);
                    for (int kb = 0; kb < nkb; ++kb, ++kbc) {
                        const int s = kbc % NST;
                        mbar_wait_g(&empty[s], ((kbc / NST) & 1) ^ 1);   // released in both CTAs by the leader's commit multicast
                        const uint32_t st = smem_u32g(smem + s * STAGE_BYTES);
                        uint32_t lbar;  // the LEADER's full[s] in the cluster shared window
                        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(lbar) : "r"(smem_u32g(&full[s])), "r"(0u));
                        if (leader) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32g(&full[s])), "r"(2u * stage_tx) : "memory");
#define AM_TMA_LN(dst_, map_, c0_, c1_)                                                                                                \
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" \
                 ::"r"(dst_), "l"(reinterpret_cast<uint64_t>(map_)), "r"(lbar), "r"(c0_), "r"(c1_) : "memory")
                        AM_TMA_LN(st, &tmA, kb * BK, m0);                                   // A_hi
                        if (!FAST) AM_TMA_LN(st + A_SUB, &tmA, p.Kp + kb * BK, m0);         // A_lo
                        AM_TMA_LN(st + 2 * A_SUB, &tmB, kb * BK, nh);                       // W_hi (own half of the n-tile)
                        if (!FAST) AM_TMA_LN(st + 2 * A_SUB + B_SUB, &tmB, p.Kp + kb * BK, nh);  // W_lo
#undef AM_TMA_LN
                    }
                }
            };
            if (p.fast) produce(std::true_type{}); else produce(std::false_type{});
        }
    } else if (warp == MMA_WARP) {
        // ===================== MMA issuer: one thread of the LEADER CTA for the pair =====================
        if (lane == 0 && leader) {
            auto issue = [&](auto fc) {
                constexpr bool FAST = decltype(fc)::value;
                uint32_t kbc = 0;
                for (int nt = 0; nt < 2; ++nt) {
                    const uint32_t d = tmem_base + (uint32_t)(nt * BN);
                    for (int kb = 0; kb < nkb; ++kb, ++kbc) {
                        const int s = kbc % NST;
                        mbar_wait_g(&full[s], (kbc / NST) & 1);   // both CTAs' operands of this stage have landed
                        tc_after_g();
                        const uint32_t base = smem_u32g(smem + s * STAGE_BYTES);
                        const uint64_t a_hi = make_desc128_g(base), a_lo = make_desc128_g(base + A_SUB);
                        const uint64_t w_hi = make_desc128_g(base + 2 * A_SUB), w_lo = make_desc128_g(base + 2 * A_SUB + B_SUB);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            const uint64_t ko = (uint64_t)(k * 2);
#define AM_UMMA_LN(a_, b_, acc_)                                                                                      \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" \
                 ::"r"(d), "l"(a_), "l"(b_), "r"(IDESC), "r"((uint32_t)(acc_)) : "memory")
                            if (FAST) { AM_UMMA_LN(a_hi + ko, w_hi + ko, (kb | k) ? 1u : 0u); continue; }
                            AM_UMMA_LN(a_lo + ko, w_hi + ko, (kb | k) ? 1u : 0u);
                            AM_UMMA_LN(a_hi + ko, w_lo + ko, 1u);
                            AM_UMMA_LN(a_hi + ko, w_hi + ko, 1u);
#undef AM_UMMA_LN
                        }
                        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                     ::"r"(smem_u32g(&empty[s])), "h"((uint16_t)3) : "memory");
                    }
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                 ::"r"(smem_u32g(&tmem_full[nt])), "h"((uint16_t)3) : "memory");
                }
            };
            if (p.fast) issue(std::true_type{}); else issue(std::false_type{});
        }
    } else if (warp < EPI_WARPS) {
        // ===================== epilogue: residual + LayerNorm over the full 512-column row, straight out of TMEM =====================
        // warp (q, qt): TMEM lane quadrant q (rows 32q..32q+31), column quarter qt: columns nt*256 + qt*64 + [0,64) of both n-tiles
        const int q = warp & 3, qt = warp >> 2;
        const int r = q * 32 + lane;              // accumulator row (TMEM lane) of this thread
        const int m = m0 + r;
        const bool ok = m < p.M;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        const __nv_bfloat16* rrow = p.R2 + (int64_t)(ok ? m : 0) * (2 * (int64_t)p.ldr);
        auto chunk_col = [&](int c) { return (c >> 2) * BN + qt * 64 + (c & 3) * 16; };  // chunk c = 0..7 -> first column (TMEM == output)
        auto load_res = [&](int col, uint4 (&rb)[4]) {
            if (ok) {
                const uint4* hp = reinterpret_cast<const uint4*>(rrow + col);
                const uint4* lp = reinterpret_cast<const uint4*>(rrow + p.ldr + col);
                rb[0] = __ldg(hp); rb[1] = __ldg(hp + 1); rb[2] = __ldg(lp); rb[3] = __ldg(lp + 1);
            } else {
                rb[0] = rb[1] = rb[2] = rb[3] = make_uint4(0, 0, 0, 0);
            }
        };
        // pass 1: x = acc + bias + residual -> back into the accumulator columns; row sum.  The residual segment of chunk c + 1 is
        // in flight while chunk c is processed (and the first one while the tensor cores still work on n-tile 0).
        float sum = 0.f;
        uint4 rb[2][4];
        load_res(chunk_col(0), rb[0]);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (c == 0) { mbar_wait_g(&tmem_full[0], 0); tc_after_g(); }
            if (c == 4) { mbar_wait_g(&tmem_full[1], 0); tc_after_g(); }
            if (c + 1 < 8) load_res(chunk_col(c + 1), rb[(c + 1) & 1]);
            const int col = chunk_col(c);
            uint32_t v[16];
            tmem_ld16_g(trow + (uint32_t)col, v);
            const uint4(&cur)[4] = rb[c & 1];
            const uint32_t hw[8] = {cur[0].x, cur[0].y, cur[0].z, cur[0].w, cur[1].x, cur[1].y, cur[1].z, cur[1].w};
            const uint32_t lw[8] = {cur[2].x, cur[2].y, cur[2].z, cur[2].w, cur[3].x, cur[3].y, cur[3].z, cur[3].w};
#pragma unroll
            for (int cc = 0; cc < 16; cc += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col + cc));
                const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int i = cc + e;
                    const uint32_t hwd = hw[i >> 1], lwd = lw[i >> 1];
                    const float res = (i & 1) ? __uint_as_float(hwd & 0xffff0000u) + __uint_as_float(lwd & 0xffff0000u)
                                              : __uint_as_float(hwd << 16) + __uint_as_float(lwd << 16);
                    const float x = __uint_as_float(v[i]) + bb[e] + res;
                    sum += x;
                    v[i] = __float_as_uint(x);
                }
            }
            tmem_st16_g(trow + (uint32_t)col, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        xch[qt * BM + r] = sum;
        asm volatile("bar.sync 1, 512;" ::: "memory");   // the 16 epilogue warps
        const float mean = (xch[r] + xch[BM + r] + xch[2 * BM + r] + xch[3 * BM + r]) * (1.0f / NOUT);
        asm volatile("bar.sync 1, 512;" ::: "memory");
        // pass 2: centred sum of squares (two TMEM loads in flight per wait)
        float sq = 0.f;
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
            uint32_t v[16], u[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(trow + (uint32_t)chunk_col(c)) : "memory");
            tmem_ld16_g(trow + (uint32_t)chunk_col(c + 1), u);  // its wait::ld covers both loads
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float d0 = __uint_as_float(v[i]) - mean, d1 = __uint_as_float(u[i]) - mean;
                sq = fmaf(d0, d0, sq); sq = fmaf(d1, d1, sq);
            }
        }
        xch[qt * BM + r] = sq;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        const float rstd = rsqrtf((xch[r] + xch[BM + r] + xch[2 * BM + r] + xch[3 * BM + r]) * (1.0f / NOUT) + p.eps);
        // pass 3: normalise, affine, bf16 (hi | lo) split, store
        __nv_bfloat16* yrow = p.Y2 + (int64_t)(ok ? m : 0) * (2 * NOUT);
#pragma unroll 2
        for (int c = 0; c < 8; ++c) {
            const int col = chunk_col(c);
            uint32_t v[16];
            tmem_ld16_g(trow + (uint32_t)col, v);
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int cc = 0; cc < 16; cc += 4) {
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + col + cc));
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.beta + col + cc));
                const float y0 = (__uint_as_float(v[cc]) - mean) * rstd * g4.x + b4.x, y1 = (__uint_as_float(v[cc + 1]) - mean) * rstd * g4.y + b4.y;
                const float y2 = (__uint_as_float(v[cc + 2]) - mean) * rstd * g4.z + b4.z, y3 = (__uint_as_float(v[cc + 3]) - mean) * rstd * g4.w + b4.w;
                const uint32_t ha = pack2_g(y0, y1), hb = pack2_g(y2, y3);
                hi[cc / 2] = ha; hi[cc / 2 + 1] = hb;
                lo[cc / 2] = pack2_g(y0 - __uint_as_float(ha << 16), y1 - __uint_as_float(ha & 0xffff0000u));
                lo[cc / 2 + 1] = pack2_g(y2 - __uint_as_float(hb << 16), y3 - __uint_as_float(hb & 0xffff0000u));
            }
            if (ok) {
                uint4* hp = reinterpret_cast<uint4*>(yrow + col);
                uint4* lp = reinterpret_cast<uint4*>(yrow + NOUT + col);
                hp[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]); hp[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                lp[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]); lp[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            }
        }
    }
    tc_before_g();
    __syncthreads();
    // neither CTA may exit (or free its TMEM) while the pair's MMAs / remote arrives can still touch it
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == ALLOC_WARP) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

typedef CUresult (*EncodeTiledFnG)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool make_map_g(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, int box_rows) {
    static EncodeTiledFnG enc = nullptr;
    if (!enc) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return false;
        enc = reinterpret_cast<EncodeTiledFnG>(fp);
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

extern "C" int am_linear_ln_tc(const void* A2, const void* W2, int M, int N, int Kp, const float* bias, const void* R2, int ldr,
                               const float* gamma, const float* beta, float eps, void* Y2, am_stream_t stream) {
    AM_REQUIRE(A2 && W2 && bias && R2 && gamma && beta && Y2, AM_EINVAL, "am_linear_ln_tc: null pointer");
    AM_REQUIRE(M > 0 && N == NOUT && Kp > 0 && Kp % BK == 0 && ldr >= NOUT && ldr % 8 == 0, AM_EINVAL,
               "am_linear_ln_tc: N must be 512, Kp a multiple of 64, ldr >= 512 and a multiple of 8");
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    AM_REQUIRE(al16(A2) && al16(W2) && al16(R2) && al16(Y2) && al16(bias) && al16(gamma) && al16(beta), AM_EALIGN,
               "am_linear_ln_tc: 16-byte alignment required");
    CUtensorMap tmA, tmB;
    AM_REQUIRE(make_map_g(&tmA, A2, (uint64_t)M, (uint64_t)2 * Kp, BM), AM_ELAUNCH, "am_linear_ln_tc: cuTensorMapEncodeTiled(A) failed");
    AM_REQUIRE(make_map_g(&tmB, W2, (uint64_t)N, (uint64_t)2 * Kp, BN / 2), AM_ELAUNCH, "am_linear_ln_tc: cuTensorMapEncodeTiled(W) failed");
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(gemm_ln_2sm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) {
            am_set_error_("am_linear_ln_tc: shared memory opt-in failed");
            return AM_ELAUNCH;
        }
        attr = true;
    }
    LnParams p{M, Kp, bias, reinterpret_cast<const __nv_bfloat16*>(R2), ldr, gamma, beta, eps, reinterpret_cast<__nv_bfloat16*>(Y2), am_get_precision()};
    const int items = cdiv(M, 2 * BM);
    if (am_launch(gemm_ln_2sm_kernel, dim3(2 * items), dim3(THREADS), SMEM_BYTES, as_stream(stream), 2, tmA, tmB, p) != cudaSuccess) {
        am_set_error_("am_linear_ln_tc: CTA-pair launch failed");
        return AM_ELAUNCH;
    }
    AM_LAUNCH_CHECK("linear_ln_tc");
    return AM_OK;
}
