// tcgen05 tensor-core GEMM with fp32-equivalent accuracy (3-term bf16 split), sm_100a only.
//
//   Y[M,N] = act( A[M,K] W[N,K]^T + bias ) (+ residual)
//   A = A_hi + A_lo, W = W_hi + W_lo (bf16 pairs, 16 significant bits each)  ->
//   A W^T ~= A_lo W_hi^T + A_hi W_lo^T + A_hi W_hi^T     (fp32 accumulation in TMEM; dropped term ~2^-16 relative)
//
// Why not plain bf16 / tf32: the parity budget is 1e-3 max-abs vs the fp32 reference; one bf16 pass measures
// 1.9e-2 on CMDM and kind::tf32 truncates to 10 mantissa bits (~2.4e-3).  3 bf16 MMAs per product
// cost 3/2.25 PF = 1.33 us/TFLOP — still 10x the fp32 SIMT pipe.
//
// Kernels in this file (am_linear_tc picks one per call):
//   gemm_tc_2sm_kernel        DEFAULT.  Persistent CTA pairs, tcgen05.mma.cta_group::2 (UMMA M = 256): each SM of the pair holds
//                             its own 128 A rows and HALF of the W tile (BK = 64, SWIZZLE_128B rows; BK = 32 when Kp % 64 != 0),
//                             256- or 128-wide tiles with the last partial round cut into half-width items, accumulators
//                             double-buffered in TMEM, 8 epilogue warps.  Epilogue modes: 3 = bias(+GELU) -> bf16 (hi|lo) via TMA
//                             store, 4 = bias(+GELU)(+ fp32 or bf16-pair residual via TMA load) -> fp32 via TMA store, 5 = token row
//                             map / broadcast residual with 16-byte LSU accesses, 0 = general (any N, any map), 1 / 2 = LSU
//                             variants of 3 / 4 (AMB200_TC_EPI=lsu).
//   gemm_tc_persistent_kernel 1-SM persistent kernel (small problems: fewer than 16 tile pairs; AMB200_TC_2SM=0)
//   gemm_tc_cluster_kernel    1-SM kernel with 2-CTA W multicast (kept for A/B runs: measured no faster)
//   gemm_tc_kernel            legacy one-128x128-tile-per-CTA kernels (AMB200_TC_VARIANT=32x3|32x6|64x3), described below:
//   warp 0      : TMA producer  — per K-block of 32 loads the FOUR 128x32 sub-tiles {A_hi, A_lo, W_hi, W_lo}
//                 (SWIZZLE_64B, 8 KB each) exactly once into a 3-stage mbarrier ring (each sub-tile feeds 2 MMAs)
//   warp 1      : MMA issuer    — one thread issues 6 x tcgen05.mma (M128 N128 K16, kind::f16 bf16->f32) per stage,
//                 tcgen05.commit releases the stage / signals the epilogue
//   warp 2      : TMEM allocator (128 columns)
//   warps 4..7  : epilogue      — tcgen05.ld 32x32b.x32 (one accumulator row per thread), bias / activation /
//                 residual / token row map, fp32 store and/or bf16 (hi|lo) split store for the next GEMM
// Operand layout in HBM: A2 [M, 2*Kp] bf16 = (hi | lo), W2 [N, 2*Kp] bf16 = (hi | lo), Kp % 32 == 0, zero padded.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128;
constexpr int TMEM_COLS = 128;
constexpr int TC_THREADS = 256;
constexpr int TCP_THREADS = 384;  // persistent kernel: 8 epilogue warps + 4 control warps
constexpr int ALLOC_WARP = 8, PRODUCER_WARP = 10, MMA_WARP = 11;
constexpr int smem_bytes(int BK, int STAGES) { return STAGES * 4 * BM * BK * 2 + 1024 /*align slack*/ + 256 /*barriers*/; }

struct TcParams {
    int M, N, Kp;
    const float* bias; int act;
    const float* residual; int ldr; int res_mod;
    float* Y; int ldy; int yin_g, yout_g, y_off;
    __nv_bfloat16* Y2; int Np2;
    int vecY;
    long long* dbg;  // optional timeline buffer (tools/tc_timeline.py): CTA (0,0) records clock64() per role
    int res_split;   // MODE 4: `residual` is a bf16 (hi|lo) split tensor [M, 2*ldr] (the previous LayerNorm's only output), not fp32
    int fast;        // am_set_precision(1): single bf16 pass (A_hi W_hi^T only; the lo halves are neither loaded nor multiplied)
    int* rowflags;   // CTA-pair kernel, MODE 4: per 128-row block, incremented as the block's output columns become globally visible
                     // (a consumer kernel that overlaps this GEMM waits on it: am_layernorm_flags); NULL = off
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded spin: a protocol bug traps (reported as a launch failure) instead of hanging the GPU box
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > 200000000u) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* holder, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t r[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int ACT>
__device__ __forceinline__ float act_ct(float x) {
    if (ACT == AM_ACT_GELU) return gelu_erf(x);
    if (ACT == AM_ACT_SILU) return silu_f(x);
    if (ACT == AM_ACT_RELU) return fmaxf(x, 0.0f);
    return x;
}

// K-major operand tile in shared memory, rows of BK bf16 = one swizzle span (BK=32: SWIZZLE_64B, BK=64: SWIZZLE_128B),
// 8-row groups 8*row_bytes apart.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout [61,64))
template <int BK>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                          // LBO (ignored for swizzled K-major; CUTLASS writes 1)
    d |= (uint64_t)((8 * BK * 2) >> 4) << 32;        // SBO = 8 rows * row bytes
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)(BK == 64 ? 2 : 4) << 61;         // LayoutType::SWIZZLE_128B : SWIZZLE_64B
    return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 [4,6)=1 | a_format BF16 [7,10)=1 | b_format BF16 [10,13)=1 | K-major A/B
// (bits 15,16 = 0) | N>>3 [17,23) | M>>4 [24,29)
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ int64_t map_row_tc(int m, int gin, int gout, int off, bool& valid) {
    valid = true;
    if (gin <= 0) return m;
    int r = off + (m % gin);
    valid = r >= 0 && r < gout;
    return (int64_t)(m / gin) * gout + r;
}

// Epilogue of one accumulator tile: thread (q, lane) owns accumulator row r = 32q + lane of the 128-row tile whose first
// TMEM column is `tmem_tile`; NCOLS output columns starting at global column n0.  bias / activation / residual / token
// row map, fp32 store and/or bf16 (hi|lo) split store.
template <int ACT, int NCOLS>
__device__ __forceinline__ void epilogue_tile(const TcParams& p, uint32_t tmem_tile, int m0, int n0, int q, int lane) {
    const int m = m0 + q * 32 + lane;
    const uint32_t tmem_base = tmem_tile;
    constexpr int BN = NCOLS;
        bool row_ok = m < p.M, map_ok = true;
        int64_t yrow = row_ok ? map_row_tc(m, p.yin_g, p.yout_g, p.y_off, map_ok) : 0;
        row_ok = row_ok && map_ok;
        const float* rrow = nullptr;
        if (p.residual && row_ok) rrow = p.residual + (p.res_mod > 0 ? (int64_t)(m % p.res_mod) : yrow) * p.ldr;
        const bool after = (p.act & AM_ACT_AFTER_RES) != 0;
        const bool rvec = rrow && ((reinterpret_cast<uintptr_t>(rrow) & 15u) == 0);
        const bool bvec = p.bias && ((reinterpret_cast<uintptr_t>(p.bias) & 15u) == 0);
        // compact epilogue: rolled loop over 16-column groups (keeps the instruction footprint small — a fully unrolled
        // 128-column epilogue with a runtime activation switch thrashed the instruction cache, see profiles/)
#pragma unroll 1
        for (int c = 0; c < BN / 16; ++c) {
            uint32_t v[16];
            __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the divergent tail of the previous group
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 16), v);
            const int nb = n0 + c * 16;
            if (!row_ok || (nb >= p.N && (!p.Y2 || nb >= p.Np2))) continue;
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
            const bool full = nb + 15 < p.N;
            if (p.bias) {
                if (full && bvec) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nb + j));
                        f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (nb + j < p.N) f[j] += __ldg(p.bias + nb + j);
                }
            }
            float rr[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) rr[j] = 0.f;
            if (rrow) {
                if (full && rvec && ((nb & 3) == 0)) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float4 r4 = *reinterpret_cast<const float4*>(rrow + nb + j);
                        rr[j] = r4.x; rr[j + 1] = r4.y; rr[j + 2] = r4.z; rr[j + 3] = r4.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (nb + j < p.N) rr[j] = rrow[nb + j];
                }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float x = after ? act_ct<ACT>(f[j] + rr[j]) : act_ct<ACT>(f[j]) + rr[j];
                f[j] = (nb + j < p.N) ? x : 0.f;
            }
            if (p.Y && nb < p.N) {
                float* dst = p.Y + yrow * p.ldy + nb;
                if (p.vecY && full) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (nb + j < p.N) dst[j] = f[j];
                }
            }
            if (p.Y2 && nb < p.Np2) {  // Np2 % 32 == 0: a 16-column group is either inside the padded width or skipped
                __nv_bfloat16* hi = p.Y2 + yrow * (2 * (int64_t)p.Np2) + nb;
                __nv_bfloat16* lo = hi + p.Np2;
                uint32_t ph[8], pl[8];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    __nv_bfloat16 h0 = __float2bfloat16_rn(f[j]), h1 = __float2bfloat16_rn(f[j + 1]);
                    __nv_bfloat16 l0 = __float2bfloat16_rn(f[j] - __bfloat162float(h0)), l1 = __float2bfloat16_rn(f[j + 1] - __bfloat162float(h1));
                    ph[j / 2] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                    pl[j / 2] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                }
                *reinterpret_cast<uint4*>(hi) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                *reinterpret_cast<uint4*>(hi + 8) = make_uint4(ph[4], ph[5], ph[6], ph[7]);
                *reinterpret_cast<uint4*>(lo) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                *reinterpret_cast<uint4*>(lo + 8) = make_uint4(pl[4], pl[5], pl[6], pl[7]);
            }
        }
    
}

// Coalesced epilogue (persistent kernel).  8 epilogue warps: warp (q, half) owns accumulator rows [32q, 32q+32) and the
// column half [half*NCOLS/2, ...).  Each 32x32 sub-block is transposed through a per-warp 4 KB staging buffer (16-byte
// chunks XOR-swizzled by row, conflict-free both ways) so that one warp instruction covers TWO whole row segments:
// lanes 0-15 -> row 2i, lanes 16-31 -> row 2i+1, lane & 15 = column pair -> 128 B fp32 / 64 B bf16 contiguous per row.
// Branch-free row loop (predicated accesses only) so the compiler can software-pipeline the shared/global loads.
template <int ACT, int NCOLS>
__device__ __forceinline__ void epilogue_tile_coalesced(const TcParams& p, uint32_t tmem_tile, int m0, int n0, int q, int half, int lane,
                                                        float* stage /* this warp's 32 x 32 fp32 buffer */) {
    const unsigned FULL = 0xffffffffu;
    const int m_own = m0 + q * 32 + lane;
    bool row_ok = m_own < p.M, map_ok = true;
    int64_t yrow64 = row_ok ? map_row_tc(m_own, p.yin_g, p.yout_g, p.y_off, map_ok) : 0;
    row_ok = row_ok && map_ok;
    const unsigned okmask = __ballot_sync(FULL, row_ok);
    const int yrow_own = (int)yrow64;
    const int rres_own = p.res_mod > 0 ? (m_own % p.res_mod) : (int)yrow64;
    const bool after = (p.act & AM_ACT_AFTER_RES) != 0;
    const bool y_v2 = p.Y && ((reinterpret_cast<uintptr_t>(p.Y) & 7u) == 0) && (p.ldy % 2 == 0);
    const bool r_v2 = p.residual && ((reinterpret_cast<uintptr_t>(p.residual) & 7u) == 0) && (p.ldr % 2 == 0);
    const int rsel = lane >> 4, cp = lane & 15;
    const long long dflags = p.dbg ? p.dbg[255] : 0;  // timeline experiments only: 1 = no global stores, 2 = no tcgen05.ld, 4 = skip phase 2
#pragma unroll 1
    for (int blk = 0; blk < NCOLS / 64; ++blk) {
        const int col0 = half * (NCOLS / 2) + blk * 32;  // first column of this 32-wide sub-block inside the tile
        const int nblk0 = n0 + col0;
        const bool any_col = nblk0 < p.N || (p.Y2 && nblk0 < p.Np2);
        __syncwarp();
        // phase 1: TMEM (lane = row) -> staging, 8 chunks of 16 B per row, chunk' = chunk ^ (row & 7)
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            uint32_t v[16];
            if (!(dflags & 2)) tmem_ld16(tmem_tile + ((uint32_t)(q * 32) << 16) + (uint32_t)(col0 + g * 16), v);
            else { for (int z = 0; z < 16; ++z) v[z] = 0; }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c4 = g * 4 + k;
                *reinterpret_cast<uint4*>(stage + lane * 32 + ((c4 ^ (lane & 7)) << 2)) = make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
            }
        }
        __syncwarp();
        if (!any_col || (dflags & 4)) continue;
        // phase 2: lane = (row parity, column pair)
        const int n = nblk0 + 2 * cp;
        const bool c0ok = n < p.N, c1ok = n + 1 < p.N;
        float b0 = 0.f, b1 = 0.f;
        if (p.bias) { if (c0ok) b0 = __ldg(p.bias + n); if (c1ok) b1 = __ldg(p.bias + n + 1); }
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const int r = 2 * i + rsel;
            const int yrow = __shfl_sync(FULL, yrow_own, r);
            const int rres = __shfl_sync(FULL, rres_own, r);
            const bool ok = (okmask >> r) & 1u;
            const float2 xv = *reinterpret_cast<const float2*>(stage + r * 32 + (((cp >> 1) ^ (r & 7)) << 2) + ((cp & 1) << 1));
            float x0 = xv.x + b0, x1 = xv.y + b1;
            float r0 = 0.f, r1 = 0.f;
            if (p.residual && ok) {
                const float* rp = p.residual + (int64_t)rres * p.ldr + n;
                if (r_v2 && c1ok) { const float2 t = *reinterpret_cast<const float2*>(rp); r0 = t.x; r1 = t.y; }
                else { if (c0ok) r0 = rp[0]; if (c1ok) r1 = rp[1]; }
            }
            if (after) { x0 = act_ct<ACT>(x0 + r0); x1 = act_ct<ACT>(x1 + r1); }
            else { x0 = act_ct<ACT>(x0) + r0; x1 = act_ct<ACT>(x1) + r1; }
            x0 = c0ok ? x0 : 0.f;
            x1 = c1ok ? x1 : 0.f;
            if (p.Y && ok && c0ok && !(dflags & 1)) {
                float* dst = p.Y + (int64_t)yrow * p.ldy + n;
                if (y_v2 && c1ok) *reinterpret_cast<float2*>(dst) = make_float2(x0, x1);
                else { dst[0] = x0; if (c1ok) dst[1] = x1; }
            }
            if (p.Y2 && ok && n < p.Np2) {
                const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
                uint32_t* hi = reinterpret_cast<uint32_t*>(p.Y2 + (int64_t)yrow * (2 * (int64_t)p.Np2) + n);
                hi[0] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                hi[p.Np2 / 2] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
        }
    }
}

// Fast-path epilogue (the four big CMDM GEMMs): no row map, no broadcast residual, tile entirely inside N, even strides.
// MODE 1: bias (+act) -> bf16 (hi|lo) only (in_proj, FFN1);  MODE 2: bias (+act) (+ fp32 residual) -> fp32 only (out_proj, FFN2,
// CDM per-point MLP).
// Everything that is loop-invariant is hoisted and all 16 shared / residual loads of a 32x32 block are issued before the
// first dependent instruction, so the two epilogue warps of a scheduler keep several memory operations in flight.
template <int ACT, int NCOLS, int MODE>
__device__ __forceinline__ void epilogue_tile_fast(const TcParams& p, uint32_t tmem_tile, int m0, int n0, int q, int half, int lane,
                                                   uint32_t stage_u32) {
    const int m_base = m0 + q * 32;
    const int rows_valid = p.M - m_base;  // rows r < rows_valid exist
    // phase-2 mapping: 8 lanes per row (16-byte chunk = 4 columns each), 4 rows per warp instruction -> half the LSU
    // instructions of the 8-byte mapping (the epilogue's shared/global traffic competes with the MMA issue path)
    const int rsel = lane >> 3, ch = lane & 7;
    const long long dflags = p.dbg ? p.dbg[255] : 0;  // timeline experiments: 1 = no global stores, 8 = no ld.shared
#pragma unroll 1
    for (int blk = 0; blk < NCOLS / 64; ++blk) {
        const int col0 = half * (NCOLS / 2) + blk * 32;
        const int n = n0 + col0 + 4 * ch;
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            uint32_t v[16];
            tmem_ld16(tmem_tile + ((uint32_t)(q * 32) << 16) + (uint32_t)(col0 + g * 16), v);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c4 = g * 4 + k;
                const uint32_t a = stage_u32 + lane * 128 + ((c4 ^ (lane & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v[4 * k]), "r"(v[4 * k + 1]), "r"(v[4 * k + 2]), "r"(v[4 * k + 3]) : "memory");
            }
        }
        __syncwarp();
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + n));
        float4 xv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + rsel;
            const uint32_t a = stage_u32 + r * 128 + ((ch ^ (r & 7)) << 4);
            if (!(dflags & 8)) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(xv[i].x), "=f"(xv[i].y), "=f"(xv[i].z), "=f"(xv[i].w) : "r"(a) : "memory");
            else xv[i] = make_float4(1.f, 2.f, 3.f, 4.f);
        }
        if (MODE == 2) {
            float4 rv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = 4 * i + rsel;
                rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.residual && r < rows_valid) rv[i] = *reinterpret_cast<const float4*>(p.residual + (int64_t)(m_base + r) * p.ldr + n);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = 4 * i + rsel;
                float4 o;
                o.x = act_ct<ACT>(xv[i].x + bv.x) + rv[i].x; o.y = act_ct<ACT>(xv[i].y + bv.y) + rv[i].y;
                o.z = act_ct<ACT>(xv[i].z + bv.z) + rv[i].z; o.w = act_ct<ACT>(xv[i].w + bv.w) + rv[i].w;
                if (r < rows_valid && !(dflags & 1)) {
                    if (dflags & 16) *reinterpret_cast<float4*>(p.Y + (int64_t)(m_base + r) * p.ldy + n) = o;
                    else __stcs(reinterpret_cast<float4*>(p.Y + (int64_t)(m_base + r) * p.ldy + n), o);  // streaming store: written once, read by a later kernel
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = 4 * i + rsel;
                const float x0 = act_ct<ACT>(xv[i].x + bv.x), x1 = act_ct<ACT>(xv[i].y + bv.y);
                const float x2 = act_ct<ACT>(xv[i].z + bv.z), x3 = act_ct<ACT>(xv[i].w + bv.w);
                __nv_bfloat162 h01 = __floats2bfloat162_rn(x0, x1), h23 = __floats2bfloat162_rn(x2, x3);
                const uint32_t u01 = *reinterpret_cast<uint32_t*>(&h01), u23 = *reinterpret_cast<uint32_t*>(&h23);
                __nv_bfloat162 l01 = __floats2bfloat162_rn(x0 - __uint_as_float(u01 << 16), x1 - __uint_as_float(u01 & 0xffff0000u));
                __nv_bfloat162 l23 = __floats2bfloat162_rn(x2 - __uint_as_float(u23 << 16), x3 - __uint_as_float(u23 & 0xffff0000u));
                if (r < rows_valid && !(dflags & 1)) {
                    uint2* hi = reinterpret_cast<uint2*>(p.Y2 + (int64_t)(m_base + r) * (2 * (int64_t)p.Np2) + n);
                    if (dflags & 16) {
                        hi[0] = make_uint2(u01, u23);
                        hi[p.Np2 / 4] = make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23));
                    } else {
                        __stcs(hi, make_uint2(u01, u23));
                        __stcs(hi + p.Np2 / 4, make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23)));
                    }
                }
            }
        }
    }
}

// MODE 5: the fast LSU epilogue with a token ROW MAP and a row-broadcast residual (motion_adapter: rows (b, t) of x_t land on
// token rows b*S + 2+G + t, residual = positional encoding row t), fp32 and / or bf16 (hi|lo) outputs.  Same 16-byte lane
// mapping as MODE 2 (8 lanes per row, 4 rows per warp instruction); the general MODE 0 path moves 8 bytes per lane with a
// shuffle-dependent address per row and took 34 us for this 1.8 GFLOP GEMM.
template <int ACT, int NCOLS>
__device__ __forceinline__ void epilogue_tile_fast_mapped(const TcParams& p, uint32_t tmem_tile, int m0, int n0, int q, int half, int lane,
                                                          uint32_t stage_u32) {
    const int m_base = m0 + q * 32;
    const int rsel = lane >> 3, ch = lane & 7;
#pragma unroll 1
    for (int blk = 0; blk < NCOLS / 64; ++blk) {
        const int col0 = half * (NCOLS / 2) + blk * 32;
        const int n = n0 + col0 + 4 * ch;
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            uint32_t v[16];
            tmem_ld16(tmem_tile + ((uint32_t)(q * 32) << 16) + (uint32_t)(col0 + g * 16), v);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c4 = g * 4 + k;
                const uint32_t a = stage_u32 + lane * 128 + ((c4 ^ (lane & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v[4 * k]), "r"(v[4 * k + 1]), "r"(v[4 * k + 2]), "r"(v[4 * k + 3]) : "memory");
            }
        }
        __syncwarp();
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + n));
        float4 xv[8], rv[8];
        int64_t yrow[8];
        bool ok[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + rsel, m = m_base + r;
            const uint32_t a = stage_u32 + r * 128 + ((ch ^ (r & 7)) << 4);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(xv[i].x), "=f"(xv[i].y), "=f"(xv[i].z), "=f"(xv[i].w) : "r"(a) : "memory");
            bool map_ok = true;
            ok[i] = m < p.M;
            yrow[i] = ok[i] ? map_row_tc(m, p.yin_g, p.yout_g, p.y_off, map_ok) : 0;
            ok[i] = ok[i] && map_ok;
            rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.residual && ok[i])
                rv[i] = *reinterpret_cast<const float4*>(p.residual + (p.res_mod > 0 ? (int64_t)(m % p.res_mod) : yrow[i]) * p.ldr + n);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (!ok[i]) continue;
            const float x0 = act_ct<ACT>(xv[i].x + bv.x) + rv[i].x, x1 = act_ct<ACT>(xv[i].y + bv.y) + rv[i].y;
            const float x2 = act_ct<ACT>(xv[i].z + bv.z) + rv[i].z, x3 = act_ct<ACT>(xv[i].w + bv.w) + rv[i].w;
            if (p.Y) *reinterpret_cast<float4*>(p.Y + yrow[i] * p.ldy + n) = make_float4(x0, x1, x2, x3);
            if (p.Y2) {
                __nv_bfloat162 h01 = __floats2bfloat162_rn(x0, x1), h23 = __floats2bfloat162_rn(x2, x3);
                const uint32_t u01 = *reinterpret_cast<uint32_t*>(&h01), u23 = *reinterpret_cast<uint32_t*>(&h23);
                __nv_bfloat162 l01 = __floats2bfloat162_rn(x0 - __uint_as_float(u01 << 16), x1 - __uint_as_float(u01 & 0xffff0000u));
                __nv_bfloat162 l23 = __floats2bfloat162_rn(x2 - __uint_as_float(u23 << 16), x3 - __uint_as_float(u23 & 0xffff0000u));
                uint2* hi = reinterpret_cast<uint2*>(p.Y2 + yrow[i] * (2 * (int64_t)p.Np2) + n);
                hi[0] = make_uint2(u01, u23);
                hi[p.Np2 / 4] = make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23));
            }
        }
    }
}

// TMA-store epilogue (MODE 3, the bf16 (hi|lo)-only outputs of in_proj / FFN1).  Measured: LSU global stores of the
// epilogue throttle the main loop (12.6K -> 17K cycles per 128x256 tile) and cost 6K of the epilogue's 10K cycles,
// whatever their cache policy; handing the writes to the TMA engine removes the LSU from the output path entirely:
//   tcgen05.ld (lane = row) -> bias / activation -> pack bf16 hi / lo -> st.shared.v4 into two 32x32 bf16 tiles laid out
//   SWIZZLE_64B (conflict-free for one-row-per-lane writes) -> fence.proxy.async -> cp.async.bulk.tensor store (hi, lo).
// No transpose staging and no shared-memory reads at all.  OOB rows (m >= M) are clipped by the tensor map.
template <int ACT, int NCOLS>
__device__ __forceinline__ void epilogue_tile_tma(const TcParams& p, const CUtensorMap* tmY, uint32_t tmem_tile, int m0, int n0, int q, int half,
                                                  int lane, uint32_t out_u32 /* this warp's 4 KB: hi tile | lo tile */) {
    const int m_base = m0 + q * 32;
#pragma unroll 1
    for (int blk = 0; blk < NCOLS / 64; ++blk) {
        const int col0 = half * (NCOLS / 2) + blk * 32;
        const int n = n0 + col0;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            uint32_t v[16];
            tmem_ld16(tmem_tile + ((uint32_t)(q * 32) << 16) + (uint32_t)(col0 + g * 16), v);
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + g * 16 + j));
                const float x0 = act_ct<ACT>(__uint_as_float(v[j]) + b4.x), x1 = act_ct<ACT>(__uint_as_float(v[j + 1]) + b4.y);
                const float x2 = act_ct<ACT>(__uint_as_float(v[j + 2]) + b4.z), x3 = act_ct<ACT>(__uint_as_float(v[j + 3]) + b4.w);
                __nv_bfloat162 h01 = __floats2bfloat162_rn(x0, x1), h23 = __floats2bfloat162_rn(x2, x3);
                const uint32_t u01 = *reinterpret_cast<uint32_t*>(&h01), u23 = *reinterpret_cast<uint32_t*>(&h23);
                __nv_bfloat162 l01 = __floats2bfloat162_rn(x0 - __uint_as_float(u01 << 16), x1 - __uint_as_float(u01 & 0xffff0000u));
                __nv_bfloat162 l23 = __floats2bfloat162_rn(x2 - __uint_as_float(u23 << 16), x3 - __uint_as_float(u23 & 0xffff0000u));
                hi[g * 8 + j / 2] = u01; hi[g * 8 + j / 2 + 1] = u23;
                lo[g * 8 + j / 2] = *reinterpret_cast<uint32_t*>(&l01); lo[g * 8 + j / 2 + 1] = *reinterpret_cast<uint32_t*>(&l23);
            }
        }
        // previous block's bulk stores must have finished READING this warp's staging tiles
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        // row = lane, 64-byte rows, 16-byte chunk c stored at chunk c ^ ((row >> 1) & 3)   (SWIZZLE_64B)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t a = out_u32 + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(hi[4 * c]), "r"(hi[4 * c + 1]), "r"(hi[4 * c + 2]), "r"(hi[4 * c + 3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + 2048), "r"(lo[4 * c]), "r"(lo[4 * c + 1]), "r"(lo[4 * c + 2]), "r"(lo[4 * c + 3]) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the async (TMA) proxy
        __syncwarp();
        if (lane == 0 && m_base < p.M && n < p.N) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(tmY)), "r"(out_u32), "r"(n), "r"(m_base) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(tmY)), "r"(out_u32 + 2048), "r"(p.Np2 + n), "r"(m_base) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
}

// TMA-load + TMA-store epilogue for fp32 outputs (MODE 4: out_proj / FFN2 / CDM per-point MLP — bias (+act) (+ fp32 residual)).
// The LSU path (MODE 2) makes these GEMMs epilogue-bound: per 32x32 block every lane waits for 8 residual loads and issues 8
// global stores, and those stores also slow the main loop (see MODE 3).  Here the residual arrives through the TMA engine one
// half-block AHEAD (two 2 KB ping-pong buffers per warp; the first load of a tile is issued before its accumulator is ready)
// and the result leaves through a TMA store from the same buffer:
//   tcgen05.ld (lane = row, 16 columns) -> + bias (+act) -> + residual from smem -> st.shared in place -> TMA store.
// Half-block = 32 rows x 16 fp32 columns = 64-byte rows, SWIZZLE_64B (chunk c of row r at c ^ ((r >> 1) & 3)): conflict-free
// for one-row-per-lane accesses.  Rows >= M are zero-filled on load and clipped on store by the tensor maps.
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        if (++spins > 200000000u) __trap();
    }
}
template <int ACT, int NCOLS>
__device__ __forceinline__ void epilogue_tile_tma_f32(const TcParams& p, const CUtensorMap* tmYf, const CUtensorMap* tmR, uint32_t tmem_tile,
                                                      int m0, int n0, int q, int half, int lane, uint32_t buf_u32 /* 2 x 2 KB */,
                                                      uint32_t bar_u32 /* 2 mbarriers */, uint32_t& phases, uint64_t* tmem_full_bar,
                                                      uint32_t tmem_full_parity) {
    constexpr int NHB = NCOLS / 32;  // 16-column half-blocks owned by this warp
    const int m_base = m0 + q * 32;
    const bool rows_ok = m_base < p.M;
    const bool has_res = p.residual != nullptr;
    const int ncol0 = n0 + half * (NCOLS / 2);
    // residual half-block -> buffer: fp32 box (32 rows x 64 B, SWIZZLE_64B) or, for a split residual, the hi and lo bf16 boxes
    // (32 rows x 32 B each, unswizzled, at +0 and +1024)
    auto load_res = [&](uint32_t dst, uint32_t bar, int n) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2048u) : "memory");
        if (p.res_split) {
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmR)), "r"(bar), "r"(n), "r"(m_base) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(dst + 1024u), "l"(reinterpret_cast<uint64_t>(tmR)), "r"(bar), "r"(p.ldr + n), "r"(m_base) : "memory");
        } else {
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmR)), "r"(bar), "r"(n), "r"(m_base) : "memory");
        }
    };
    if (rows_ok && has_res && lane == 0) {
        // residual of half-block 0, issued while this tile's main loop may still be running; buffer 0 was last read by the
        // bulk store two groups back
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        load_res(buf_u32, bar_u32, ncol0);
    }
    mbar_wait(tmem_full_bar, tmem_full_parity);
    tc_fence_after();
    if (!rows_ok) return;
#pragma unroll 1
    for (int j = 0; j < NHB; ++j) {
        const int b = j & 1;
        const uint32_t buf = buf_u32 + (uint32_t)b * 2048u;
        const int n = ncol0 + j * 16;
        uint32_t v[16];
        tmem_ld16(tmem_tile + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * (NCOLS / 2) + j * 16), v);
        float f[16];
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + c));
            f[c] = act_ct<ACT>(__uint_as_float(v[c]) + b4.x); f[c + 1] = act_ct<ACT>(__uint_as_float(v[c + 1]) + b4.y);
            f[c + 2] = act_ct<ACT>(__uint_as_float(v[c + 2]) + b4.z); f[c + 3] = act_ct<ACT>(__uint_as_float(v[c + 3]) + b4.w);
        }
        if (lane == 0) {
            if (has_res && j + 1 < NHB) {
                // the other buffer was read by the most recent bulk store: wait for that read, then prefetch the next residual
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                load_res(buf_u32 + (uint32_t)(b ^ 1) * 2048u, bar_u32 + (uint32_t)(b ^ 1) * 8u, n + 16);
            } else {
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // this buffer's previous store (two groups back) is done reading
            }
        }
        __syncwarp();
        const uint32_t rowa = buf + (uint32_t)lane * 64u;
        const uint32_t sw = (uint32_t)((lane >> 1) & 3);
        if (has_res) {
            mbar_wait_s(bar_u32 + (uint32_t)b * 8u, (phases >> b) & 1u);
            phases ^= (1u << b);
            if (p.res_split) {
                // x = hi + lo (16 significant bits): this lane's 16 columns = 32 B of the hi box and 32 B of the lo box
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(h0), "=r"(h1), "=r"(h2), "=r"(h3) : "r"(buf + (uint32_t)lane * 32u + (uint32_t)c * 16u) : "memory");
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(l0), "=r"(l1), "=r"(l2), "=r"(l3) : "r"(buf + 1024u + (uint32_t)lane * 32u + (uint32_t)c * 16u) : "memory");
                    const uint32_t hw[4] = {h0, h1, h2, h3}, lw[4] = {l0, l1, l2, l3};
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        f[8 * c + 2 * w] += __uint_as_float(hw[w] << 16) + __uint_as_float(lw[w] << 16);
                        f[8 * c + 2 * w + 1] += __uint_as_float(hw[w] & 0xffff0000u) + __uint_as_float(lw[w] & 0xffff0000u);
                    }
                }
                __syncwarp();  // the fp32 rows written below overlap OTHER lanes' hi / lo rows: everyone reads first
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float r0, r1, r2, r3;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r0), "=f"(r1), "=f"(r2), "=f"(r3) : "r"(rowa + (((uint32_t)c ^ sw) << 4)) : "memory");
                    f[4 * c] += r0; f[4 * c + 1] += r1; f[4 * c + 2] += r2; f[4 * c + 3] += r3;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + (((uint32_t)c ^ sw) << 4)), "f"(f[4 * c]), "f"(f[4 * c + 1]), "f"(f[4 * c + 2]), "f"(f[4 * c + 3]) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(tmYf)), "r"(buf), "r"(n), "r"(m_base) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
}

template <int BK, int STAGES, int MINB, int ACT>
__global__ void __launch_bounds__(TC_THREADS, MINB)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcParams p) {
    constexpr int SUB_BYTES = BM * BK * 2;      // one 128 x BK bf16 sub-tile
    constexpr int STAGE_BYTES = 4 * SUB_BYTES;  // A_hi, A_lo, W_hi, W_lo
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int nkb = p.Kp / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_holder, TMEM_COLS);
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) p.dbg[0] = clock64();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) p.dbg[1] = clock64();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                int s = kb % STAGES;
                uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && kb < 64) p.dbg[8 + kb] = clock64();
                uint8_t* st = smem + s * STAGE_BYTES;
                mbar_expect_tx(&full[s], STAGE_BYTES);
                tma_load_2d(st + 0 * SUB_BYTES, &tmA, &full[s], kb * BK, m0);          // A_hi
                tma_load_2d(st + 1 * SUB_BYTES, &tmA, &full[s], p.Kp + kb * BK, m0);   // A_lo
                tma_load_2d(st + 2 * SUB_BYTES, &tmB, &full[s], kb * BK, n0);          // W_hi
                tma_load_2d(st + 3 * SUB_BYTES, &tmB, &full[s], p.Kp + kb * BK, n0);   // W_lo
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (single thread) =====================
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                int s = kb % STAGES;
                uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&full[s], ph);
                if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && kb < 64) p.dbg[72 + kb] = clock64();
                tc_fence_after();
                uint32_t base = smem_u32(smem + s * STAGE_BYTES);
                uint64_t a_hi = make_desc<BK>(base), a_lo = make_desc<BK>(base + SUB_BYTES);
                uint64_t w_hi = make_desc<BK>(base + 2 * SUB_BYTES), w_lo = make_desc<BK>(base + 3 * SUB_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {  // K=16 per MMA -> +32 B inside the 64 B swizzle row (encoded +2)
                    uint64_t ko = (uint64_t)(k * 2);
                    umma_bf16(tmem_base, a_lo + ko, w_hi + ko, IDESC, (kb | k) ? 1u : 0u);
                    umma_bf16(tmem_base, a_hi + ko, w_lo + ko, IDESC, 1u);
                    umma_bf16(tmem_base, a_hi + ko, w_hi + ko, IDESC, 1u);
                }
                umma_commit(&empty[s]);               // stage free once these MMAs retire
                if (kb == nkb - 1) umma_commit(tmem_full);
                if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && kb < 64) p.dbg[136 + kb] = clock64();
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: one accumulator row per thread =====================
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 128) p.dbg[2] = clock64();
        epilogue_tile<ACT, 128>(p, tmem_base, m0, n0, q, lane);
    }
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 128) p.dbg[3] = clock64();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) p.dbg[4] = clock64();
}


// ---------------------------------------------------------------------------------------------------------------------
// Persistent variant (default): one CTA per SM loops over output tiles (tile = m_tile * n_tiles + n_tile, static
// round-robin), BM = 128, BN_ in {128, 256}, K-block 32.  The smem ring runs continuously ACROSS tiles (NST stages of
// {A_hi, A_lo, W_hi, W_lo}; 192 KB in flight so the ~2000-cycle TMA latency is covered), and the accumulator is
// double-buffered in TMEM (2 x BN_ columns) so the epilogue of tile i overlaps the main loop of tile i+1.
template <int BN_, int NST, int ACT, int MODE>
__global__ void __launch_bounds__(TCP_THREADS, 1)
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR, TcParams p, int n_tiles,
                          int total_tiles) {
    constexpr int BK = 32;
    constexpr int A_SUB = BM * BK * 2;          // 8 KB
    constexpr int B_SUB = BN_ * BK * 2;         // 8 / 16 KB
    constexpr int STAGE_BYTES = 2 * A_SUB + 2 * B_SUB;
    constexpr uint32_t IDESC_P = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN_ >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NST * STAGE_BYTES);
    uint64_t* empty = full + NST;
    uint64_t* tmem_full = empty + NST;     // [2]
    uint64_t* tmem_empty = tmem_full + 2;  // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* stage_buf = reinterpret_cast<float*>(smem + NST * STAGE_BYTES + 1024);  // 8 warps x 4 KB (epilogue transpose / TMA-store tiles), 1024-aligned
    uint64_t* epi_bars = reinterpret_cast<uint64_t*>(smem + NST * STAGE_BYTES + 512);  // [8 warps][2]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.Kp / BK;
    const bool dbg = p.dbg && blockIdx.x == 0;

    // Warp roles: the SM sub-partition arbiter favours the HIGHEST warp id (B300_MICROARCH.md), so the two single-thread
    // control roles get the top ids and the eight epilogue warps the low ones: warps 0-7 epilogue, 8 TMEM allocator,
    // 10 TMA producer, 11 MMA issuer.  (With the control warps at ids 0/1 the main loop slowed from 12.5K to 17K cycles per
    // tile whenever the epilogue was active.)
    if (warp == PRODUCER_WARP && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    }
    if (warp == MMA_WARP && lane == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 256); }
        for (int i = 0; i < 16; ++i) mbar_init(&epi_bars[i], 1);  // MODE 4: two residual-load barriers per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == ALLOC_WARP) tmem_alloc(tmem_holder, 2 * BN_);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();  // prologue above overlapped the previous kernel; from here on its results are complete and visible
    if (dbg && threadIdx.x == 0) p.dbg[0] = clock64();

    if (warp == PRODUCER_WARP) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            // FAST (am_set_precision(1)) is resolved ONCE per role: the parity loop carries no extra instruction (measured: a
            // per-k-step `if (p.fast)` in these single-thread loops cost 4-11 % of the GEMM time)
            auto produce = [&](auto fc) {
                constexpr bool FAST = decltype(fc)::value;
                uint32_t kbc = 0;  // running K-block counter: the ring never drains between tiles
                for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                    const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN_;
                    for (int kb = 0; kb < nkb; ++kb, ++kbc) {
                        const int s = kbc % NST;
                        mbar_wait(&empty[s], ((kbc / NST) & 1) ^ 1);
                        uint8_t* st = smem + s * STAGE_BYTES;
                        mbar_expect_tx(&full[s], FAST ? (A_SUB + B_SUB) : STAGE_BYTES);
                        tma_load_2d(st, &tmA, &full[s], kb * BK, m0);                          // A_hi
                        if (!FAST) tma_load_2d(st + A_SUB, &tmA, &full[s], p.Kp + kb * BK, m0);           // A_lo
                        tma_load_2d(st + 2 * A_SUB, &tmB, &full[s], kb * BK, n0);              // W_hi
                        if (!FAST) tma_load_2d(st + 2 * A_SUB + B_SUB, &tmB, &full[s], p.Kp + kb * BK, n0);  // W_lo
                    }
                }
            };
            if (p.fast) produce(std::true_type{}); else produce(std::false_type{});
        }
    } else if (warp == MMA_WARP) {
        // ===================== MMA issuer (single thread) =====================
        if (lane == 0) {
            auto issue = [&](auto fc) {
                constexpr bool FAST = decltype(fc)::value;
                uint32_t kbc = 0;
                int it = 0;
                for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                    const int ab = it & 1;
                    mbar_wait(&tmem_empty[ab], (((it >> 1) & 1) ^ 1));  // epilogue drained this accumulator (first two: free)
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)(ab * BN_);
                    for (int kb = 0; kb < nkb; ++kb, ++kbc) {
                        const int s = kbc % NST;
                        mbar_wait(&full[s], (kbc / NST) & 1);
                        tc_fence_after();
                        const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
                        const uint64_t a_hi = make_desc<BK>(base), a_lo = make_desc<BK>(base + A_SUB);
                        const uint64_t w_hi = make_desc<BK>(base + 2 * A_SUB), w_lo = make_desc<BK>(base + 2 * A_SUB + B_SUB);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            const uint64_t ko = (uint64_t)(k * 2);
                            if (FAST) { umma_bf16(d, a_hi + ko, w_hi + ko, IDESC_P, (kb | k) ? 1u : 0u); continue; }
                            umma_bf16(d, a_lo + ko, w_hi + ko, IDESC_P, (kb | k) ? 1u : 0u);
                            umma_bf16(d, a_hi + ko, w_lo + ko, IDESC_P, 1u);
                            umma_bf16(d, a_hi + ko, w_hi + ko, IDESC_P, 1u);
                        }
                        umma_commit(&empty[s]);
                    }
                    umma_commit(&tmem_full[ab]);
                    if (dbg && it < 16) p.dbg[8 + it] = clock64();
                }
            };
            if (p.fast) issue(std::true_type{}); else issue(std::false_type{});
        }
    } else if (warp < 8) {
        // ===================== epilogue (overlaps the next tile's main loop) =====================
        const int q = warp & 3, half = warp >> 2;
        int it = 0;
        uint32_t epi_phases = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int ab = it & 1;
            const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN_;
            if (MODE == 4) {
                epilogue_tile_tma_f32<ACT, BN_>(p, &tmY, &tmR, tmem_base + (uint32_t)(ab * BN_), m0, n0, q, half, lane,
                                                smem_u32(stage_buf + warp * (32 * 32)), smem_u32(epi_bars + warp * 2), epi_phases,
                                                &tmem_full[ab], (uint32_t)((it >> 1) & 1));
                tc_fence_before();
                mbar_arrive(&tmem_empty[ab]);
                continue;
            }
            mbar_wait(&tmem_full[ab], (it >> 1) & 1);
            tc_fence_after();
            if (dbg && threadIdx.x == 0 && it < 16) p.dbg[40 + it] = clock64();
            if (MODE == 3) epilogue_tile_tma<ACT, BN_>(p, &tmY, tmem_base + (uint32_t)(ab * BN_), m0, n0, q, half, lane, smem_u32(stage_buf + warp * (32 * 32)));
            else if (MODE == 0) epilogue_tile_coalesced<ACT, BN_>(p, tmem_base + (uint32_t)(ab * BN_), m0, n0, q, half, lane, stage_buf + warp * (32 * 32));
            else if (MODE == 5) epilogue_tile_fast_mapped<ACT, BN_>(p, tmem_base + (uint32_t)(ab * BN_), m0, n0, q, half, lane, smem_u32(stage_buf + warp * (32 * 32)));
                else epilogue_tile_fast<ACT, BN_, MODE>(p, tmem_base + (uint32_t)(ab * BN_), m0, n0, q, half, lane, smem_u32(stage_buf + warp * (32 * 32)));
            tc_fence_before();
            mbar_arrive(&tmem_empty[ab]);
            if (dbg && threadIdx.x == 0 && it < 16) p.dbg[72 + it] = clock64();
        }
        if ((MODE == 3 || MODE == 4) && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all bulk stores of this warp complete before exit
    }
    tc_fence_before();
    __syncthreads();
    if (warp == ALLOC_WARP) tmem_dealloc(tmem_base, 2 * BN_);
}

// ---------------------------------------------------------------------------------------------------------------------
// 2-CTA thread-block clusters with TMA multicast.  The persistent kernel is limited by the SM<->L2 port: TMA loads (62 B/clk)
// plus the output stores exceed what one SM can move, and the main loop drops from 12.6K to 17K cycles per tile.  Here the
// two CTAs of a cluster work on the same n-tile: each fetches HALF of the W tile and multicasts it into both shared
// memories, so per-CTA load traffic per K-block falls from 48 KB to 32 KB (BN=256).  Stage release is cluster-wide
// (tcgen05.commit multicast arrive on both CTAs' `empty` barriers).
template <int BN_, int NST, int ACT, int MODE>
__global__ void __launch_bounds__(TCP_THREADS, 1)
gemm_tc_cluster_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR, TcParams p, int n_tiles,
                          int total_tiles) {
    constexpr int BK = 32;
    constexpr int A_SUB = BM * BK * 2;          // 8 KB
    constexpr int B_SUB = BN_ * BK * 2;         // 8 / 16 KB
    constexpr int STAGE_BYTES = 2 * A_SUB + 2 * B_SUB;
    constexpr uint32_t IDESC_P = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN_ >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NST * STAGE_BYTES);
    uint64_t* empty = full + NST;
    uint64_t* tmem_full = empty + NST;     // [2]
    uint64_t* tmem_empty = tmem_full + 2;  // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* stage_buf = reinterpret_cast<float*>(smem + NST * STAGE_BYTES + 1024);  // 8 warps x 4 KB (epilogue transpose / TMA-store tiles), 1024-aligned
    uint64_t* epi_bars = reinterpret_cast<uint64_t*>(smem + NST * STAGE_BYTES + 512);  // [8 warps][2]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.Kp / BK;
    uint32_t cta_rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int total_pairs = total_tiles;  // (m-tile pair, n-tile) work items: both CTAs share the n-tile (W multicast), own one m-tile each
    const bool dbg = p.dbg && blockIdx.x == 0;

    // Warp roles: the SM sub-partition arbiter favours the HIGHEST warp id (B300_MICROARCH.md), so the two single-thread
    // control roles get the top ids and the eight epilogue warps the low ones: warps 0-7 epilogue, 8 TMEM allocator,
    // 10 TMA producer, 11 MMA issuer.  (With the control warps at ids 0/1 the main loop slowed from 12.5K to 17K cycles per
    // tile whenever the epilogue was active.)
    if (warp == PRODUCER_WARP && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    }
    if (warp == MMA_WARP && lane == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2); }  // empty: both CTAs' MMA commits
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 256); }
        for (int i = 0; i < 16; ++i) mbar_init(&epi_bars[i], 1);  // MODE 4: two residual-load barriers per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == ALLOC_WARP) tmem_alloc(tmem_holder, 2 * BN_);
    tc_fence_before();
    __syncthreads();
    // barriers of both CTAs initialised before any remote arrive / multicast write
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();  // prologue above overlapped the previous kernel; from here on its results are complete and visible
    if (dbg && threadIdx.x == 0) p.dbg[0] = clock64();

    if (warp == PRODUCER_WARP) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t kbc = 0;  // running K-block counter: the ring never drains between tiles
            for (int tile = cluster_id; tile < total_pairs; tile += num_clusters) {
                const int m0 = ((tile / n_tiles) * 2 + (int)cta_rank) * BM, n0 = (tile % n_tiles) * BN_;
                for (int kb = 0; kb < nkb; ++kb, ++kbc) {
                    const int s = kbc % NST;
                    mbar_wait(&empty[s], ((kbc / NST) & 1) ^ 1);
                    uint8_t* st = smem + s * STAGE_BYTES;
                    mbar_expect_tx(&full[s], STAGE_BYTES);
                    tma_load_2d(st, &tmA, &full[s], kb * BK, m0);                          // A_hi
                    tma_load_2d(st + A_SUB, &tmA, &full[s], p.Kp + kb * BK, m0);           // A_lo
                    // W tile: this CTA fetches rows [rank*BN/2, +BN/2) and MULTICASTS them into both CTAs of the cluster
                    const int nh = n0 + (int)cta_rank * (BN_ / 2);
                    const uint32_t hoff = cta_rank * (B_SUB / 2);
                    tma_load_2d_mc(st + 2 * A_SUB + hoff, &tmB, &full[s], kb * BK, nh, (uint16_t)3);                  // W_hi half
                    tma_load_2d_mc(st + 2 * A_SUB + B_SUB + hoff, &tmB, &full[s], p.Kp + kb * BK, nh, (uint16_t)3);   // W_lo half
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // ===================== MMA issuer (single thread) =====================
        if (lane == 0) {
            uint32_t kbc = 0;
            int it = 0;
            for (int tile = cluster_id; tile < total_pairs; tile += num_clusters, ++it) {
                const int ab = it & 1;
                mbar_wait(&tmem_empty[ab], (((it >> 1) & 1) ^ 1));  // epilogue drained this accumulator (first two: free)
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(ab * BN_);
                for (int kb = 0; kb < nkb; ++kb, ++kbc) {
                    const int s = kbc % NST;
                    mbar_wait(&full[s], (kbc / NST) & 1);
                    tc_fence_after();
                    const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
                    const uint64_t a_hi = make_desc<BK>(base), a_lo = make_desc<BK>(base + A_SUB);
                    const uint64_t w_hi = make_desc<BK>(base + 2 * A_SUB), w_lo = make_desc<BK>(base + 2 * A_SUB + B_SUB);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t ko = (uint64_t)(k * 2);
                        umma_bf16(d, a_lo + ko, w_hi + ko, IDESC_P, (kb | k) ? 1u : 0u);
                        umma_bf16(d, a_hi + ko, w_lo + ko, IDESC_P, 1u);
                        umma_bf16(d, a_hi + ko, w_hi + ko, IDESC_P, 1u);
                    }
                    umma_commit_mc(&empty[s], (uint16_t)3);  // the stage is refilled by BOTH CTAs' multicasts: release it in both
                }
                umma_commit(&tmem_full[ab]);
                if (dbg && it < 16) p.dbg[8 + it] = clock64();
            }
        }
    } else if (warp < 8) {
        // ===================== epilogue (overlaps the next tile's main loop) =====================
        const int q = warp & 3, half = warp >> 2;
        int it = 0;
        uint32_t epi_phases = 0;
        for (int tile = cluster_id; tile < total_pairs; tile += num_clusters, ++it) {
            const int ab = it & 1;
            const int m0 = ((tile / n_tiles) * 2 + (int)cta_rank) * BM, n0 = (tile % n_tiles) * BN_;
            if (MODE == 4) {
                epilogue_tile_tma_f32<ACT, BN_>(p, &tmY, &tmR, tmem_base + (uint32_t)(ab * BN_), m0, n0, q, half, lane,
                                                smem_u32(stage_buf + warp * (32 * 32)), smem_u32(epi_bars + warp * 2), epi_phases,
                                                &tmem_full[ab], (uint32_t)((it >> 1) & 1));
                tc_fence_before();
                mbar_arrive(&tmem_empty[ab]);
                continue;
            }
            mbar_wait(&tmem_full[ab], (it >> 1) & 1);
            tc_fence_after();
            if (dbg && threadIdx.x == 0 && it < 16) p.dbg[40 + it] = clock64();
            if (MODE == 3) epilogue_tile_tma<ACT, BN_>(p, &tmY, tmem_base + (uint32_t)(ab * BN_), m0, n0, q, half, lane, smem_u32(stage_buf + warp * (32 * 32)));
            else if (MODE == 0) epilogue_tile_coalesced<ACT, BN_>(p, tmem_base + (uint32_t)(ab * BN_), m0, n0, q, half, lane, stage_buf + warp * (32 * 32));
            else if (MODE == 5) epilogue_tile_fast_mapped<ACT, BN_>(p, tmem_base + (uint32_t)(ab * BN_), m0, n0, q, half, lane, smem_u32(stage_buf + warp * (32 * 32)));
                else epilogue_tile_fast<ACT, BN_, MODE>(p, tmem_base + (uint32_t)(ab * BN_), m0, n0, q, half, lane, smem_u32(stage_buf + warp * (32 * 32)));
            tc_fence_before();
            mbar_arrive(&tmem_empty[ab]);
            if (dbg && threadIdx.x == 0 && it < 16) p.dbg[72 + it] = clock64();
        }
        if ((MODE == 3 || MODE == 4) && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all bulk stores of this warp complete before exit
    }
    tc_fence_before();
    __syncthreads();
    // neither CTA may exit while the peer can still multicast into its smem / arrive on its barriers
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == ALLOC_WARP) tmem_dealloc(tmem_base, 2 * BN_);
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair kernel: tcgen05.mma.cta_group::2 (UMMA M = 256 across two SMs of a TPC).
// The 1-SM kernels above are SHARED-MEMORY-BANDWIDTH bound, not tensor bound: with the 3-term split every K16 step issues three
// MMAs and each re-reads its A (4 KB) and B (8 KB at N=256) slices from smem — 96 B/clk — while the TMA engine writes 62 B/clk
// of new operands and the epilogue stages 21 B/clk, against 128 B/clk of smem bandwidth (measured: 12.6K cycles per 128x256
// tile alone, 17K with the epilogue running, 22K average; tile width, multicast and epilogue type make no difference —
// profiles/r1_tc_tile_sweep.txt).  In a CTA pair each SM keeps only HALF of the W tile: the pair computes a 256 x BN tile,
// SM r holds A rows [128r, 128r+128) and W rows [r*BN/2, (r+1)*BN/2), the tensor cores read the peer's half through the
// pair's smem path.  Per SM: MMA reads 64 B/clk, TMA writes 42 B/clk, operand traffic from L2 drops by a third.
//   * both CTAs: TMA producer (own A tile, own W half) — completion is signalled on the LEADER's `full` barrier
//     (cp.async.bulk.tensor .cta_group::2, leader expects 2 x STAGE bytes); epilogue of the own 128 x BN accumulator
//   * leader CTA (rank 0) only: one thread issues the MMAs; tcgen05.commit multicasts the stage release / accumulator-ready
//     arrives to both CTAs; the peer's epilogue threads arrive remotely on the leader's `tmem_empty`
template <int BN_, int NST, int ACT, int MODE, int BK>
__global__ void __launch_bounds__(TCP_THREADS, 1)
gemm_tc_2sm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmBq, const __grid_constant__ CUtensorMap tmBo, const __grid_constant__ CUtensorMap tmY,
                   const __grid_constant__ CUtensorMap tmR, TcParams p, int n_tiles, int wide_items, int total_items, int tail_shift) {
    // Work items: [0, wide_items) are 256 x BN_ pair tiles (m-pair = idx / n_tiles, n-tile = idx % n_tiles); items beyond are the
    // REMAINING pair tiles cut into 2^tail_shift pieces of BN_ >> tail_shift columns each.  M*N / (148 SMs x tile) is rarely an
    // integer (1.1 for the N = 512 trunk GEMMs, 2.2 for N = 1024): instead of another full round on a fraction of the SMs, the last
    // partial round runs as half- or quarter-width tiles spread over 2x / 4x as many CTA pairs (tmBq / tmBo: W boxes of BN_/4 / BN_/8
    // rows per CTA); the host picks the shift with the smaller estimated makespan.
    auto decode = [&](int idx, int& mp, int& n0, int& sh) {
        if (idx < wide_items) { mp = idx / n_tiles; n0 = (idx % n_tiles) * BN_; sh = 0; }
        else {
            const int u = idx - wide_items, w = wide_items + (u >> tail_shift);
            mp = w / n_tiles; n0 = (w % n_tiles) * BN_ + (u & ((1 << tail_shift) - 1)) * (BN_ >> tail_shift); sh = tail_shift;
        }
    };
    // BK = 32: 64-byte operand rows (SWIZZLE_64B); BK = 64: 128-byte rows (SWIZZLE_128B)
    constexpr int A_SUB = BM * BK * 2;            // this CTA's 128 rows of A (hi or lo)
    constexpr int B_SUB = (BN_ / 2) * BK * 2;     // this CTA's half of the W tile (hi or lo)
    constexpr int STAGE_BYTES = 2 * A_SUB + 2 * B_SUB;
    constexpr uint32_t IDESC_P = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN_ >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NST * STAGE_BYTES);
    uint64_t* empty = full + NST;
    uint64_t* tmem_full = empty + NST;     // [2]
    uint64_t* tmem_empty = tmem_full + 2;  // [2]  (used in the leader CTA only)
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* stage_buf = reinterpret_cast<float*>(smem + NST * STAGE_BYTES + 1024);
    uint64_t* epi_bars = reinterpret_cast<uint64_t*>(smem + NST * STAGE_BYTES + 512);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.Kp / BK;
    uint32_t cta_rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    const bool leader = cta_rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (warp == PRODUCER_WARP && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    }
    if (warp == MMA_WARP && lane == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 512); }  // both CTAs' 256 epilogue threads
        for (int i = 0; i < 16; ++i) mbar_init(&epi_bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == ALLOC_WARP) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"((uint32_t)(2 * BN_)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();  // prologue above overlapped the previous kernel; from here on its results are complete and visible
    const bool dbg = p.dbg && blockIdx.x == 0;
    if (dbg && threadIdx.x == 0) {
        p.dbg[0] = clock64();
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        p.dbg[1] = (long long)ns;
    }

    if (warp == PRODUCER_WARP) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            auto produce = [&](auto fc) {
            constexpr bool FAST = decltype(fc)::value;
            uint32_t kbc = 0;
            for (int tile = cluster_id; tile < total_items; tile += num_clusters) {
                int mp, n0, sh;
                decode(tile, mp, n0, sh);
                const int m0 = (mp * 2 + (int)cta_rank) * BM;
                const int nh = n0 + (int)cta_rank * (BN_ >> (sh + 1));
                const CUtensorMap* tmW = sh == 0 ? &tmB : (sh == 1 ? &tmBq : &tmBo);
                const uint32_t stage_tx = FAST ? (uint32_t)(A_SUB + (B_SUB >> sh)) : (uint32_t)(2 * A_SUB + 2 * (B_SUB >> sh));
                for (int kb = 0; kb < nkb; ++kb, ++kbc) {
                    const int s = kbc % NST;
                    mbar_wait(&empty[s], ((kbc / NST) & 1) ^ 1);   // released in both CTAs by the leader's commit multicast
                    const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
                    uint32_t lbar;  // the LEADER's full[s] in the cluster shared window
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(lbar) : "r"(smem_u32(&full[s])), "r"(0u));
                    if (leader) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(2u * stage_tx) : "memory");
#define AM_TMA_2SM(dst_, map_, c0_, c1_)                                                                                              \
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" \
                 ::"r"(dst_), "l"(reinterpret_cast<uint64_t>(map_)), "r"(lbar), "r"(c0_), "r"(c1_) : "memory")
                    AM_TMA_2SM(st, &tmA, kb * BK, m0);                               // A_hi (own 128 rows)
                    if (!FAST) AM_TMA_2SM(st + A_SUB, &tmA, p.Kp + kb * BK, m0);   // A_lo
                    AM_TMA_2SM(st + 2 * A_SUB, tmW, kb * BK, nh);                    // W_hi, own half of the n-tile
                    if (!FAST) AM_TMA_2SM(st + 2 * A_SUB + B_SUB, tmW, p.Kp + kb * BK, nh);     // W_lo
#undef AM_TMA_2SM
                }
            }
            };
            if (p.fast) produce(std::true_type{}); else produce(std::false_type{});
        }
    } else if (warp == MMA_WARP) {
        // ===================== MMA issuer: the LEADER CTA's warp for the pair =====================
        // The whole warp runs the loop (convergent) and `elect.sync` predicates each tcgen05 instruction: under a divergent
        // `if (lane == 0)` ptxas wraps every tcgen05.mma in an ELECT / 7 x R2UR.BROADCAST / BRA.U.ANY loop over the active lanes.
        if (leader) {
            auto issue = [&](auto fc) {
            constexpr bool FAST = decltype(fc)::value;
            uint32_t kbc = 0;
            int it = 0;
            for (int tile = cluster_id; tile < total_items; tile += num_clusters, ++it) {
                const int ab = it & 1;
                const uint32_t idesc = tile < wide_items ? IDESC_P
                    : ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((BN_ >> tail_shift) >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24));
                mbar_wait(&tmem_empty[ab], (((it >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(ab * BN_);
                for (int kb = 0; kb < nkb; ++kb, ++kbc) {
                    const int s = kbc % NST;
                    mbar_wait(&full[s], (kbc / NST) & 1);   // both CTAs' operands of this stage have landed
                    tc_fence_after();
                    const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
                    const uint64_t a_hi = make_desc<BK>(base), a_lo = make_desc<BK>(base + A_SUB);
                    const uint64_t w_hi = make_desc<BK>(base + 2 * A_SUB), w_lo = make_desc<BK>(base + 2 * A_SUB + B_SUB);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t ko = (uint64_t)(k * 2);
#define AM_UMMA_2SM(a_, b_, acc_)                                                                                     \
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" \
                 ::"r"(d), "l"(a_), "l"(b_), "r"(idesc), "r"((uint32_t)(acc_)) : "memory")
                        if (FAST) { AM_UMMA_2SM(a_hi + ko, w_hi + ko, (kb | k) ? 1u : 0u); continue; }
                        AM_UMMA_2SM(a_lo + ko, w_hi + ko, (kb | k) ? 1u : 0u);
                        AM_UMMA_2SM(a_hi + ko, w_lo + ko, 1u);
                        AM_UMMA_2SM(a_hi + ko, w_hi + ko, 1u);
#undef AM_UMMA_2SM
                    }
                    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
                                 ::"r"(smem_u32(&empty[s])), "h"((uint16_t)3) : "memory");
                }
                asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
                             ::"r"(smem_u32(&tmem_full[ab])), "h"((uint16_t)3) : "memory");
                if (dbg && lane == 0 && it < 16) p.dbg[8 + it] = clock64();
            }
            };
            if (p.fast) issue(std::true_type{}); else issue(std::false_type{});
        }
    } else if (warp < 8) {
        // ===================== epilogue (both CTAs: own 128 x BN accumulator) =====================
        const int q = warp & 3, half = warp >> 2;
        int it = 0;
        uint32_t epi_phases = 0;
        for (int tile = cluster_id; tile < total_items; tile += num_clusters, ++it) {
            const int ab = it & 1;
            int mp, n0, sh;
            decode(tile, mp, n0, sh);
            const int m0 = (mp * 2 + (int)cta_rank) * BM;
            uint32_t lempty;  // the leader's tmem_empty[ab]
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(lempty) : "r"(smem_u32(&tmem_empty[ab])), "r"(0u));
            if (MODE == 4) {
                if (sh == 2)
                    epilogue_tile_tma_f32<ACT, BN_ / 4>(p, &tmY, &tmR, tmem_base + (uint32_t)(ab * BN_), m0, n0, q, half, lane,
                                                        smem_u32(stage_buf + warp * (32 * 32)), smem_u32(epi_bars + warp * 2), epi_phases,
                                                        &tmem_full[ab], (uint32_t)((it >> 1) & 1));
                else if (sh == 1)
                    epilogue_tile_tma_f32<ACT, BN_ / 2>(p, &tmY, &tmR, tmem_base + (uint32_t)(ab * BN_), m0, n0, q, half, lane,
                                                        smem_u32(stage_buf + warp * (32 * 32)), smem_u32(epi_bars + warp * 2), epi_phases,
                                                        &tmem_full[ab], (uint32_t)((it >> 1) & 1));
                else
                    epilogue_tile_tma_f32<ACT, BN_>(p, &tmY, &tmR, tmem_base + (uint32_t)(ab * BN_), m0, n0, q, half, lane,
                                                    smem_u32(stage_buf + warp * (32 * 32)), smem_u32(epi_bars + warp * 2), epi_phases,
                                                    &tmem_full[ab], (uint32_t)((it >> 1) & 1));
            } else {
                mbar_wait(&tmem_full[ab], (it >> 1) & 1);
                tc_fence_after();
                if (MODE == 3 && sh == 2) epilogue_tile_tma<ACT, BN_ / 4>(p, &tmY, tmem_base + (uint32_t)(ab * BN_), m0, n0, q, half, lane, smem_u32(stage_buf + warp * (32 * 32)));
                else if (MODE == 3 && sh == 1) epilogue_tile_tma<ACT, BN_ / 2>(p, &tmY, tmem_base + (uint32_t)(ab * BN_), m0, n0, q, half, lane, smem_u32(stage_buf + warp * (32 * 32)));
                else if (MODE == 3) epilogue_tile_tma<ACT, BN_>(p, &tmY, tmem_base + (uint32_t)(ab * BN_), m0, n0, q, half, lane, smem_u32(stage_buf + warp * (32 * 32)));
                else if (MODE == 0) epilogue_tile_coalesced<ACT, BN_>(p, tmem_base + (uint32_t)(ab * BN_), m0, n0, q, half, lane, stage_buf + warp * (32 * 32));
                else if (MODE == 5) epilogue_tile_fast_mapped<ACT, BN_>(p, tmem_base + (uint32_t)(ab * BN_), m0, n0, q, half, lane, smem_u32(stage_buf + warp * (32 * 32)));
                else epilogue_tile_fast<ACT, BN_, MODE>(p, tmem_base + (uint32_t)(ab * BN_), m0, n0, q, half, lane, smem_u32(stage_buf + warp * (32 * 32)));
            }
            tc_fence_before();
            asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(lempty) : "memory");
            if (MODE == 4 && p.rowflags) {
                // This warp's share of the item (32 rows x half of the item's columns) is globally visible: bump the row block's
                // counter.  The block is complete at 8 warps x N / 2 = 4 N; the consumer (am_layernorm_flags) resets it.
                if (lane == 0) {
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                    asm volatile("fence.proxy.async;" ::: "memory");
                    __threadfence();
                    if (m0 < p.M) atomicAdd(p.rowflags + m0 / BM, (BN_ >> sh) / 2);
                }
            }
            if (dbg && threadIdx.x == 0 && it < 16) {
                p.dbg[72 + it] = clock64();
                unsigned long long ns;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
                p.dbg[2] = (long long)ns;  // wall clock at the end of the latest epilogue -> actual SM frequency of this launch
            }
        }
        if ((MODE == 3 || MODE == 4) && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    // neither CTA may exit (or free its TMEM) while the pair's MMAs / remote arrives can still touch it
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == ALLOC_WARP) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN_)) : "memory");
}

// fp32 [M,K] -> bf16 (hi | lo) [M, 2*Kp], zero padded to Kp
__global__ void split_bf16_kernel(const float* __restrict__ X, int ldx, __nv_bfloat16* __restrict__ X2, int Kp, int M, int K) {
    pdl_launch_dependents();
    pdl_wait();
    int64_t total = (int64_t)M * Kp;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        int m = (int)(g / Kp), k = (int)(g - (int64_t)m * Kp);
        float v = k < K ? X[(int64_t)m * ldx + k] : 0.f;
        __nv_bfloat16 h = __float2bfloat16_rn(v);
        __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        X2[(int64_t)m * 2 * Kp + k] = h;
        X2[(int64_t)m * 2 * Kp + Kp + k] = l;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D bf16 tensor [rows, cols] row-major, box = BK cols x 128 rows, swizzle span = BK*2 bytes, OOB -> zeros
bool make_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, int BK, int box_rows = BM) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace

static long long* g_tc_dbg = nullptr;
static thread_local int* g_tc_next_rowflags = nullptr;   // consumed by the next am_linear_tc call of this thread (am_linear_tc_set_rowflags)
static int g_tc_bn = 0;        // tuning hook (tools/tc_tile_sweep.py): force the tile width (128 / 256), 0 = automatic
static int g_tc_cluster = -1;  // tuning hook: force the 2-CTA multicast kernel on (1) / off (0), -1 = AMB200_TC_CLUSTER / default
static int g_tc_mixed = 1;     // tuning hook: 0 disables the half-width tail items of the CTA-pair kernel
extern "C" void am_tc_set_mixed_(int on) { g_tc_mixed = on; }
static int g_tc_bk = 0;        // tuning hook: 32 forces the 64-byte-row K block in the CTA-pair kernel
extern "C" void am_tc_set_bk_(int bk) { g_tc_bk = bk; }
static int g_tc_2sm = -1;      // tuning hook: CTA-pair (cta_group::2) kernel on (1) / off (0), -1 = AMB200_TC_2SM / default (on)
extern "C" void am_tc_set_2sm_(int on) { g_tc_2sm = on; }
static int g_tc_lsu = -1;      // tuning hook: 1 = LSU epilogues (MODE 1 / 2) instead of the TMA ones (MODE 3 / 4), -1 = AMB200_TC_EPI / default
extern "C" void am_tc_set_tile_(int bn, int cluster) { g_tc_bn = bn; g_tc_cluster = cluster; }
extern "C" void am_tc_set_epi_(int lsu) { g_tc_lsu = lsu; }
// debug hook (not part of the public header): timeline buffer of >= 200 int64 for the next am_linear_tc launches
extern "C" void am_tc_set_debug_(void* buf) { g_tc_dbg = reinterpret_cast<long long*>(buf); }

extern "C" int am_split_bf16(const float* X, int ldx, void* X2, int Kp, int M, int K, am_stream_t stream) {
    AM_REQUIRE(X && X2 && M > 0 && K > 0 && Kp >= K && Kp % 32 == 0 && ldx >= K, AM_EINVAL, "am_split_bf16: bad args (Kp % 32 == 0)");
    int64_t total = (int64_t)M * Kp;
    int64_t blocks = (total + 255) / 256;
    int grid = (int)(blocks < (int64_t)AM_NUM_SMS * 8 ? blocks : (int64_t)AM_NUM_SMS * 8);
    am_launch(split_bf16_kernel, dim3(grid), dim3(256), 0, as_stream(stream), 1, X, ldx, reinterpret_cast<__nv_bfloat16*>(X2), Kp, M, K);
    AM_LAUNCH_CHECK("split_bf16");
    return AM_OK;
}

// Tail plan of the CTA-pair kernel: `pairs` 256 x 256 tiles on `clusters` CTA pairs.  Full rounds run as whole tiles; the remaining
// `rem` tiles are cut into 2^shift column pieces each so that the last round is short.  Estimated makespan in tile-times: a half-width
// piece costs 0.54, a quarter-width piece 0.30 (A is re-read per piece and the fixed per-item cost does not shrink); shift 0 = no cut.
// AMB200_TC_TAIL=half|quarter|none forces one plan (A/B timing).
static void tc_tail_plan(int pairs, int clusters, int& wide, int& shift) {
    static int force = -2;
    if (force == -2) {
        const char* e = getenv("AMB200_TC_TAIL");
        force = !e ? -1 : (!strcmp(e, "none") ? 0 : (!strcmp(e, "half") ? 1 : (!strcmp(e, "quarter") ? 2 : -1)));
    }
    const int full = (pairs / clusters) * clusters, rem = pairs - full;
    if (rem == 0 || force == 0) { wide = pairs; shift = 1; return; }
    const double c[3] = {1.0, 0.54, 0.30};
    int best = 0;
    double best_cost = 1e30;
    for (int sh = 0; sh <= 2; ++sh) {
        if (force > 0 && sh != force) continue;
        const int items = rem << sh;
        const double cost = (double)((items + clusters - 1) / clusters) * c[sh];
        if (cost < best_cost - 1e-9) { best_cost = cost; best = sh; }
    }
    if (best == 0) { wide = pairs; shift = 1; } else { wide = full; shift = best; }
}

extern "C" int am_linear_tc(const void* A2, const void* W2, int M, int N, int Kp, const float* bias, int act, const float* residual, int ldr,
                            int res_mod, float* Y, int ldy, int yin_g, int yout_g, int y_off, void* Y2, int Np2, am_stream_t stream) {
    AM_REQUIRE(A2 && W2 && (Y || Y2), AM_EINVAL, "am_linear_tc: null pointer");
    AM_REQUIRE(M > 0 && N > 0 && Kp > 0 && Kp % 32 == 0, AM_EINVAL, "am_linear_tc: Kp must be a positive multiple of 32");
    AM_REQUIRE((act & 15) <= 3 && (act & ~31) == 0, AM_EINVAL, "am_linear_tc: bad activation");
    AM_REQUIRE(!Y || ldy >= N, AM_EINVAL, "am_linear_tc: bad ldy");
    AM_REQUIRE(!Y2 || (Np2 >= N && Np2 % 32 == 0), AM_EINVAL, "am_linear_tc: Np2 must be a multiple of 32 and >= N");
    AM_REQUIRE(!residual || ldr >= N, AM_EINVAL, "am_linear_tc: bad residual stride");
    AM_REQUIRE((reinterpret_cast<uintptr_t>(A2) & 15u) == 0 && (reinterpret_cast<uintptr_t>(W2) & 15u) == 0 &&
               (!Y2 || (reinterpret_cast<uintptr_t>(Y2) & 15u) == 0), AM_EALIGN, "am_linear_tc: operands must be 16-byte aligned");
    // kernel variant (tuning knob, read once): default "persistent"; legacy one-tile-per-CTA kernels "32x3" (BK=32, 3 stages,
    // 2 CTAs/SM), "32x6", "64x3" (1 CTA/SM) are kept for A/B measurements
    static int variant = -1;
    if (variant < 0) {
        const char* e = getenv("AMB200_TC_VARIANT");
        variant = 3;
        if (e && !strcmp(e, "32x3")) variant = 0;
        if (e && !strcmp(e, "32x6")) variant = 1;
        if (e && !strcmp(e, "64x3")) variant = 2;
    }
    // res_mod == -1: `residual` is a bf16 (hi|lo) split tensor [M, 2*ldr] (what the LayerNorm kernel writes), x = hi + lo
    const int res_split = (residual && res_mod == -1) ? 1 : 0;
    if (res_split) {
        res_mod = 0;
        AM_REQUIRE(variant == 3 && (ldr % 8) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15u) == 0, AM_EINVAL,
                   "am_linear_tc: split residual needs ldr % 8 == 0 and 16-byte alignment");
    }
    if (variant == 3) {
        // tile width: 256 when that still gives every SM >= ~2 tiles, else 128 (more, smaller tiles)
        const int mt = cdiv(M, BM);
        static int sm2_env = -1;
        if (sm2_env < 0) { const char* e = getenv("AMB200_TC_2SM"); sm2_env = (e && !strcmp(e, "0")) ? 0 : 1; }
        const bool sm2_on = (g_tc_2sm >= 0 ? g_tc_2sm : sm2_env) != 0;
        // trunk-shaped GEMMs (plain epilogue -> TMA-store modes 3 / 4) run 256-wide on the CTA-pair kernel, whose last partial
        // round is cut into half-width items; everything else: 256 only when that still gives every SM >= ~2 tiles
        const bool trunk_like = yin_g == 0 && res_mod == 0 && (N % 256) == 0 && !(act & AM_ACT_AFTER_RES) &&
                                ((Y2 && !Y && !residual && Np2 == N) || (Y && !Y2));
        const bool wide = g_tc_bn ? (g_tc_bn == 256 && N >= 256)
                                  : (N >= 256 && ((int64_t)mt * cdiv(N, 256) >= 2 * AM_NUM_SMS ||
                                                  (sm2_on && g_tc_mixed && trunk_like && cdiv(mt, 2) * (N / 256) >= 16)));
        const int bn = wide ? 256 : 128;
        const int nt = cdiv(N, bn), total = mt * nt;
        CUtensorMap tmA, tmB;
        AM_REQUIRE(make_map(&tmA, A2, (uint64_t)M, (uint64_t)2 * Kp, 32, BM), AM_ELAUNCH, "am_linear_tc: cuTensorMapEncodeTiled(A) failed");
        AM_REQUIRE(make_map(&tmB, W2, (uint64_t)N, (uint64_t)2 * Kp, 32, bn), AM_ELAUNCH, "am_linear_tc: cuTensorMapEncodeTiled(W) failed");
        TcParams p{M, N, Kp, bias, act, residual, ldr, res_mod, Y, ldy, yin_g, yout_g, y_off, reinterpret_cast<__nv_bfloat16*>(Y2), Np2, 0, g_tc_dbg, res_split, 0, nullptr};
        p.fast = am_get_precision();
        int* const want_flags = g_tc_next_rowflags;
        g_tc_next_rowflags = nullptr;
        p.vecY = Y && (reinterpret_cast<uintptr_t>(Y) & 15u) == 0 && (ldy % 4 == 0);
        const int grid = total < AM_NUM_SMS ? total : AM_NUM_SMS;
        cudaStream_t st = as_stream(stream);
        // 2-CTA cluster + W multicast variant (AMB200_TC_CLUSTER=0 disables): work items = (m-tile pair, n-tile)
        static int cl_env = -1;
        if (cl_env < 0) { const char* e = getenv("AMB200_TC_CLUSTER"); cl_env = (e && !strcmp(e, "0")) ? 0 : 1; }
        const int pairs = cdiv(mt, 2) * nt;
        const bool use_2sm = sm2_on && pairs >= 16;
        const bool use_cluster = !use_2sm && (g_tc_cluster >= 0 ? g_tc_cluster : cl_env) && pairs >= 16;
        const int grid_cl = 2 * (pairs < AM_NUM_SMS / 2 ? pairs : AM_NUM_SMS / 2);
        CUtensorMap tmBh = tmB;
        // CTA-pair kernel K block: 64 (128-byte smem rows, SWIZZLE_128B) when Kp allows, else 32 (SWIZZLE_64B)
        const int bk2 = (use_2sm && Kp % 64 == 0 && g_tc_bk != 32) ? 64 : 32;
        CUtensorMap tmA64 = tmA, tmBh64 = tmB;
        if (bk2 == 64) {
            AM_REQUIRE(make_map(&tmA64, A2, (uint64_t)M, (uint64_t)2 * Kp, 64, BM), AM_ELAUNCH, "am_linear_tc: cuTensorMapEncodeTiled(A, BK=64) failed");
            AM_REQUIRE(make_map(&tmBh64, W2, (uint64_t)N, (uint64_t)2 * Kp, 64, bn / 2), AM_ELAUNCH, "am_linear_tc: cuTensorMapEncodeTiled(W half, BK=64) failed");
        }
        // mixed-width tail (CTA-pair kernel, 256-wide tiles): W quarter-tile map for the half-width items
        CUtensorMap tmBq = tmB, tmBo = tmB;
        const bool mixed_ok = use_2sm && bn == 256 && g_tc_mixed != 0;
        if (mixed_ok) AM_REQUIRE(make_map(&tmBq, W2, (uint64_t)N, (uint64_t)2 * Kp, bk2, bn / 4), AM_ELAUNCH, "am_linear_tc: cuTensorMapEncodeTiled(W quarter) failed");
        if (mixed_ok) AM_REQUIRE(make_map(&tmBo, W2, (uint64_t)N, (uint64_t)2 * Kp, bk2, bn / 8), AM_ELAUNCH, "am_linear_tc: cuTensorMapEncodeTiled(W eighth) failed");
        if (use_cluster || use_2sm) AM_REQUIRE(make_map(&tmBh, W2, (uint64_t)N, (uint64_t)2 * Kp, 32, bn / 2), AM_ELAUNCH, "am_linear_tc: cuTensorMapEncodeTiled(W half) failed");
#define AM_TC2_LAUNCH(BN_, NST2_, ACT_, MODE_, BK_)                                                                             \
    do {                                                                                                                        \
        constexpr int smem2_ = NST2_ * (2 * BM * BK_ * 2 + BN_ * BK_ * 2) + 1024 + 1024 + 8 * 32 * 32 * 4;                       \
        static bool attr3_ = false;                                                                                             \
        if (!attr3_) {                                                                                                          \
            if (cudaFuncSetAttribute(gemm_tc_2sm_kernel<BN_, NST2_, ACT_, MODE_, BK_>,                                          \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, smem2_) != cudaSuccess) {                     \
                am_set_error_("am_linear_tc: shared memory opt-in failed");                                                     \
                return AM_ELAUNCH;                                                                                              \
            }                                                                                                                   \
            attr3_ = true;                                                                                                      \
        }                                                                                                                       \
        const bool mixed_ = mixed_ok && BN_ == 256 && (MODE_ == 3 || MODE_ == 4);                                               \
        int wide_ = pairs, shift_ = 1;                                                                                          \
        if (mixed_) tc_tail_plan(pairs, grid_cl / 2, wide_, shift_);                                                            \
        const int items_ = wide_ + ((pairs - wide_) << shift_);                                                                 \
        if (am_launch(gemm_tc_2sm_kernel<BN_, NST2_, ACT_, MODE_, BK_>, dim3(grid_cl), dim3(TCP_THREADS), smem2_, st, 2,       \
                      (bk2 == 64 ? tmA64 : tmA), (bk2 == 64 ? tmBh64 : tmBh), tmBq, tmBo, tmY, tmR, p, nt, wide_, items_, shift_) != cudaSuccess) { \
            am_set_error_("am_linear_tc: CTA-pair launch failed");                                                              \
            return AM_ELAUNCH;                                                                                                  \
        }                                                                                                                       \
    } while (0)
#define AM_TCP_LAUNCH(BN_, NST_, ACT_, MODE_)                                                                                   \
    do {                                                                                                                        \
        constexpr int smem_ = NST_ * (2 * BM * 32 * 2 + 2 * BN_ * 32 * 2) + 1024 + 1024 + 8 * 32 * 32 * 4;                       \
        static bool attr_ = false;                                                                                              \
        if (!attr_) {                                                                                                           \
            if (cudaFuncSetAttribute(gemm_tc_persistent_kernel<BN_, NST_, ACT_, MODE_>,                                         \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, smem_) != cudaSuccess) {                      \
                am_set_error_("am_linear_tc: shared memory opt-in failed");                                                     \
                return AM_ELAUNCH;                                                                                              \
            }                                                                                                                   \
            attr_ = true;                                                                                                       \
        }                                                                                                                       \
        if (use_2sm && bk2 == 64) {                                                                                             \
            AM_TC2_LAUNCH(BN_, (BN_ == 256 ? 3 : 4), ACT_, MODE_, 64);                                                          \
        } else if (use_2sm) {                                                                                                   \
            AM_TC2_LAUNCH(BN_, (BN_ == 256 ? 6 : 8), ACT_, MODE_, 32);                                                          \
        } else if (use_cluster) {                                                                                               \
            static bool attr2_ = false;                                                                                         \
            if (!attr2_) {                                                                                                      \
                if (cudaFuncSetAttribute(gemm_tc_cluster_kernel<BN_, NST_, ACT_, MODE_>,                                        \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem_) != cudaSuccess) {                  \
                    am_set_error_("am_linear_tc: shared memory opt-in failed");                                                 \
                    return AM_ELAUNCH;                                                                                          \
                }                                                                                                               \
                attr2_ = true;                                                                                                  \
            }                                                                                                                   \
            if (am_launch(gemm_tc_cluster_kernel<BN_, NST_, ACT_, MODE_>, dim3(grid_cl), dim3(TCP_THREADS), smem_, st, 2, tmA, tmBh, tmY, tmR, \
                          p, nt, pairs) != cudaSuccess) {                                                                       \
                am_set_error_("am_linear_tc: cluster launch failed");                                                           \
                return AM_ELAUNCH;                                                                                              \
            }                                                                                                                   \
        } else                                                                                                                  \
        am_launch(gemm_tc_persistent_kernel<BN_, NST_, ACT_, MODE_>, dim3(grid), dim3(TCP_THREADS), smem_, st, 1, tmA, tmB, tmY, tmR, p, nt, total); \
    } while (0)
        // epilogue mode: 1 / 2 = specialised fast paths of the big trunk GEMMs, 0 = general
        const int a15 = act & 15;
        const bool plain = yin_g == 0 && res_mod == 0 && (N % bn) == 0 && !(act & AM_ACT_AFTER_RES) &&
                           (!bias || (reinterpret_cast<uintptr_t>(bias) & 15u) == 0);
        int mode = 0;
        CUtensorMap tmY = tmA;  // placeholder unless the TMA-store epilogue is selected
        if (plain && Y2 && !Y && !residual && Np2 == N && (N % 4) == 0 && (a15 == AM_ACT_NONE || a15 == AM_ACT_GELU)) mode = 1;
        static int lsu_epi = -1;
        if (lsu_epi < 0) { const char* e = getenv("AMB200_TC_EPI"); lsu_epi = (e && !strcmp(e, "lsu")) ? 1 : 0; }
        const bool lsu = g_tc_lsu >= 0 ? g_tc_lsu != 0 : lsu_epi != 0;
        if (mode == 1 && !lsu && (N % 32) == 0) {
            // output tensor map: Y2 [M, 2*Np2] bf16, box 32 cols x 32 rows, SWIZZLE_64B
            EncodeTiledFn enc = get_encode();
            cuuint64_t gdim[2] = {(cuuint64_t)2 * Np2, (cuuint64_t)M};
            cuuint64_t gstr[1] = {(cuuint64_t)2 * Np2 * 2};
            cuuint32_t box[2] = {32, 32};
            cuuint32_t estr[2] = {1, 1};
            if (enc && enc(&tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Y2, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) mode = 3;
        }
        if (plain && Y && !Y2 && (a15 == AM_ACT_NONE || a15 == AM_ACT_GELU) && (ldy % 4) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15u) == 0 &&
            (!residual || res_split || ((ldr % 4) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15u) == 0))) mode = 2;
        CUtensorMap tmR = tmA;  // placeholder unless MODE 4 reads a residual
        if (mode == 2 && !lsu) {
            // fp32 tensor maps: Y [M, ldy] and residual [M, ldr], box 16 cols x 32 rows (64-byte rows), SWIZZLE_64B
            EncodeTiledFn enc = get_encode();
            auto f32map = [&](CUtensorMap* mp, const float* ptr, int ld) {
                cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)M};
                cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
                cuuint32_t box[2] = {16, 32};
                cuuint32_t estr[2] = {1, 1};
                return enc && enc(mp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
            };
            auto splitmap = [&](CUtensorMap* mp) {  // bf16 [M, 2*ldr], box 16 cols x 32 rows (32-byte rows), no swizzle
                cuuint64_t gdim[2] = {(cuuint64_t)2 * ldr, (cuuint64_t)M};
                cuuint64_t gstr[1] = {(cuuint64_t)2 * ldr * 2};
                cuuint32_t box[2] = {16, 32};
                cuuint32_t estr[2] = {1, 1};
                return enc && enc(mp, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<float*>(residual), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
            };
            if (f32map(&tmY, Y, ldy) && (!residual || (res_split ? splitmap(&tmR) : f32map(&tmR, residual, ldr)))) mode = 4;
        }
        AM_REQUIRE(!res_split || mode == 4, AM_EINVAL, "am_linear_tc: a split residual is only supported by the TMA fp32 epilogue (plain layout, fp32 Y)");
        AM_REQUIRE(!want_flags || (mode == 4 && use_2sm), AM_EINVAL, "am_linear_tc_set_rowflags: row-block flags need the CTA-pair kernel with the TMA fp32 epilogue");
        p.rowflags = want_flags;
        // MODE 5: row-mapped / broadcast-residual outputs with 16-byte rows (motion_adapter)
        if (mode == 0 && (yin_g > 0 || res_mod > 0) && (N % bn) == 0 && !(act & AM_ACT_AFTER_RES) && (a15 == AM_ACT_NONE || a15 == AM_ACT_GELU) &&
            (!bias || (reinterpret_cast<uintptr_t>(bias) & 15u) == 0) &&
            (!Y || ((ldy % 4) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15u) == 0)) &&
            (!residual || ((ldr % 4) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15u) == 0))) mode = 5;
#define AM_TCP_BY_MODE(BN_, NST_)                                                                  \
    if (mode == 3 && a15 == AM_ACT_GELU) AM_TCP_LAUNCH(BN_, NST_, AM_ACT_GELU, 3);                  \
    else if (mode == 3) AM_TCP_LAUNCH(BN_, NST_, AM_ACT_NONE, 3);                                   \
    else if (mode == 1 && a15 == AM_ACT_GELU) AM_TCP_LAUNCH(BN_, NST_, AM_ACT_GELU, 1);             \
    else if (mode == 1) AM_TCP_LAUNCH(BN_, NST_, AM_ACT_NONE, 1);                                   \
    else if (mode == 4 && a15 == AM_ACT_GELU) AM_TCP_LAUNCH(BN_, NST_, AM_ACT_GELU, 4);             \
    else if (mode == 4) AM_TCP_LAUNCH(BN_, NST_, AM_ACT_NONE, 4);                                   \
    else if (mode == 5 && a15 == AM_ACT_GELU) AM_TCP_LAUNCH(BN_, NST_, AM_ACT_GELU, 5);             \
    else if (mode == 5) AM_TCP_LAUNCH(BN_, NST_, AM_ACT_NONE, 5);                                   \
    else if (mode == 2 && a15 == AM_ACT_GELU) AM_TCP_LAUNCH(BN_, NST_, AM_ACT_GELU, 2);             \
    else if (mode == 2) AM_TCP_LAUNCH(BN_, NST_, AM_ACT_NONE, 2);                                   \
    else switch (a15) {                                                                            \
        case AM_ACT_GELU: AM_TCP_LAUNCH(BN_, NST_, AM_ACT_GELU, 0); break;                          \
        case AM_ACT_SILU: AM_TCP_LAUNCH(BN_, NST_, AM_ACT_SILU, 0); break;                          \
        case AM_ACT_RELU: AM_TCP_LAUNCH(BN_, NST_, AM_ACT_RELU, 0); break;                          \
        default: AM_TCP_LAUNCH(BN_, NST_, AM_ACT_NONE, 0); break;                                   \
    }
        if (wide) { AM_TCP_BY_MODE(256, 4) } else { AM_TCP_BY_MODE(128, 6) }
#undef AM_TCP_BY_MODE
#undef AM_TCP_LAUNCH
#undef AM_TC2_LAUNCH
        AM_LAUNCH_CHECK("linear_tc");
        return AM_OK;
    }
    const int BKsel = variant == 2 ? 64 : 32;
    AM_REQUIRE(Kp % BKsel == 0, AM_EINVAL, "am_linear_tc: Kp must be a multiple of the K block");
    CUtensorMap tmA, tmB;
    AM_REQUIRE(make_map(&tmA, A2, (uint64_t)M, (uint64_t)2 * Kp, BKsel), AM_ELAUNCH, "am_linear_tc: cuTensorMapEncodeTiled(A) failed");
    AM_REQUIRE(make_map(&tmB, W2, (uint64_t)N, (uint64_t)2 * Kp, BKsel), AM_ELAUNCH, "am_linear_tc: cuTensorMapEncodeTiled(W) failed");
    TcParams p{M, N, Kp, bias, act, residual, ldr, res_mod, Y, ldy, yin_g, yout_g, y_off, reinterpret_cast<__nv_bfloat16*>(Y2), Np2, 0, g_tc_dbg, 0, 0, nullptr};
    p.vecY = Y && (reinterpret_cast<uintptr_t>(Y) & 15u) == 0 && (ldy % 4 == 0);
    dim3 grid(cdiv(N, BN), cdiv(M, BM));
    cudaStream_t st = as_stream(stream);
#define AM_TC_LAUNCH(BK_, ST_, MB_, ACT_)                                                                                     \
    do {                                                                                                                      \
        static bool attr_ = false;                                                                                            \
        if (!attr_) {                                                                                                         \
            if (cudaFuncSetAttribute(gemm_tc_kernel<BK_, ST_, MB_, ACT_>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                     smem_bytes(BK_, ST_)) != cudaSuccess) {                                                  \
                am_set_error_("am_linear_tc: shared memory opt-in failed");                                                   \
                return AM_ELAUNCH;                                                                                            \
            }                                                                                                                 \
            attr_ = true;                                                                                                     \
        }                                                                                                                     \
        gemm_tc_kernel<BK_, ST_, MB_, ACT_><<<grid, TC_THREADS, smem_bytes(BK_, ST_), st>>>(tmA, tmB, p);                      \
    } while (0)
#define AM_TC_BY_ACT(BK_, ST_, MB_)                                   \
    switch (act & 15) {                                               \
        case AM_ACT_GELU: AM_TC_LAUNCH(BK_, ST_, MB_, AM_ACT_GELU); break; \
        case AM_ACT_SILU: AM_TC_LAUNCH(BK_, ST_, MB_, AM_ACT_SILU); break; \
        case AM_ACT_RELU: AM_TC_LAUNCH(BK_, ST_, MB_, AM_ACT_RELU); break; \
        default: AM_TC_LAUNCH(BK_, ST_, MB_, AM_ACT_NONE); break;      \
    }
    if (variant == 0) { AM_TC_BY_ACT(32, 3, 2) }
    else if (variant == 1) { AM_TC_BY_ACT(32, 6, 1) }
    else { AM_TC_BY_ACT(64, 3, 1) }
#undef AM_TC_BY_ACT
#undef AM_TC_LAUNCH
    AM_LAUNCH_CHECK("linear_tc");
    return AM_OK;
}

namespace {
// fp32 X [M,K] (row-major, ldx) -> bf16 (hi | lo) of X^T: XT2 [K, 2*Mp] (Mp = M padded to 32, zero filled).  32x32 smem-tiled transpose.
__global__ void transpose_split_bf16_kernel(const float* __restrict__ X, int ldx, __nv_bfloat16* __restrict__ XT2, int Mp, int M, int K) {
    __shared__ float tile[32][33];
    const int m0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        int m = m0 + ty + i, k = k0 + tx;
        tile[ty + i][tx] = (m < M && k < K) ? X[(int64_t)m * ldx + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        int k = k0 + ty + i, m = m0 + tx;
        if (k < K && m < Mp) {
            float v = tile[tx][ty + i];
            __nv_bfloat16 h = __float2bfloat16_rn(v);
            XT2[(int64_t)k * 2 * Mp + m] = h;
            XT2[(int64_t)k * 2 * Mp + Mp + m] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
    }
}
}  // namespace

extern "C" int am_transpose_split_bf16(const float* X, int ldx, void* XT2, int Mp, int M, int K, am_stream_t stream) {
    AM_REQUIRE(X && XT2 && M > 0 && K > 0 && Mp >= M && Mp % 32 == 0 && ldx >= K, AM_EINVAL, "am_transpose_split_bf16: bad args (Mp % 32 == 0)");
    dim3 grid(Mp / 32, cdiv(K, 32));
    transpose_split_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>(X, ldx, reinterpret_cast<__nv_bfloat16*>(XT2), Mp, M, K);
    AM_LAUNCH_CHECK("transpose_split_bf16");
    return AM_OK;
}

// Row-block completion flags for the NEXT am_linear_tc call of this thread (CTA-pair kernel, fp32 TMA epilogue only): flags[m / 128] is
// incremented by (columns / 2) per epilogue warp as the block's outputs become globally visible and reaches 4 * N when the 128-row
// block of Y is complete.  A consumer launched with programmatic dependent launch (am_layernorm_flags) starts on finished row blocks
// while the GEMM is still computing others, and resets the counters.  flags: int32 [ceil(M / 128)], zero before the first use.
extern "C" int am_linear_tc_set_rowflags(int* flags) {
    g_tc_next_rowflags = flags;
    return AM_OK;
}

// host-only view of the tail plan (tests/test_host_cpu.py checks it without a GPU)
extern "C" int am_tc_tail_plan_(int pairs, int clusters, int* wide, int* shift) {
    if (pairs <= 0 || clusters <= 0 || !wide || !shift) return AM_EINVAL;
    tc_tail_plan(pairs, clusters, *wide, *shift);
    return AM_OK;
}
