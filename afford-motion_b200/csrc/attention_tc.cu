// tcgen05 multi-head self-attention for the CMDM trunk (S <= 384 keys, head dim 64), fp32-equivalent accuracy.
// Replaces the SDPA / native-MHA library kernel inside torch.nn.TransformerEncoderLayer (models/cmdm.py:66-77,167).
//
// One persistent CTA per (batch, head): K and V of the head (bf16 hi|lo, 192 KB) are TMA-staged ONCE and stay in
// shared memory for all ceil(S/128) query tiles; per query tile:
//   S = Q K^T   : 3 key tiles x 3 split terms x 4 k-steps of tcgen05.mma M128 N128 K16 -> TMEM columns [0,384)
//   softmax     : 8 warps in two groups (each owns every other 64-key block of all 128 rows); pass 1 row max
//                 (tcgen05.ld) exchanged through smem, pass 2 per 64-key block
//                 p = exp(s*scale - max) written BACK INTO THE S COLUMNS as packed bf16 (hi | lo) with tcgen05.st —
//                 P never touches shared memory or HBM
//   O = P V     : A operand = P from TMEM (tcgen05.mma [d], [a_tmem], b_desc), B = V tile in its natural
//                 [key, d] layout (MN-major descriptor), 3 split terms, accumulated in TMEM columns [384,448)
//   epilogue    : O / rowsum -> bf16 (hi|lo) operand of the out_proj GEMM (and optional fp32)
// The whole key row fits in TMEM, so the softmax is exact (no online rescaling).  3-term bf16 split as in gemm_tc.cu.
// Input QKV2 [B*S, 2*3*H*64] bf16 = (hi | lo) written by the in_proj GEMM epilogue.
#include <cuda.h>
#include <cuda_bf16.h>
#include <math_constants.h>
#include <type_traits>
#include <cstdlib>
#include "common.cuh"

namespace {

constexpr int HD = 64;
constexpr int QT = 128;            // queries per tile (UMMA M)
constexpr int KT = 128;            // keys per K/V smem tile
constexpr int NKT = 3;             // key tiles -> 384 keys max
constexpr int TILE_BYTES = 128 * HD * 2;   // 16 KB: 128 rows x 64 bf16 (128-byte rows, SWIZZLE_128B)
constexpr int ATT_THREADS = 384;   // 4 control warps + 8 softmax warps (two per TMEM lane quadrant)
constexpr int S_COL = 0, O_COL = 384;
// smem: Q_hi, Q_lo, K_hi[3], K_lo[3], V_hi[3], V_lo[3]  = 14 tiles = 224 KB
constexpr int OFF_QH = 0, OFF_QL = TILE_BYTES, OFF_KH = 2 * TILE_BYTES, OFF_KL = 5 * TILE_BYTES, OFF_VH = 8 * TILE_BYTES,
              OFF_VL = 11 * TILE_BYTES, OFF_BAR = 14 * TILE_BYTES;
constexpr int ATT_SMEM = OFF_BAR + 256 + 2048 + 768;  // barriers + mask bits, row max and row sum exchange [2][128] fp32 each, alignment slack
static_assert(ATT_SMEM <= 227 * 1024, "attention shared memory exceeds the 227 KB opt-in limit");

struct AttParams {
    int B, S, H;
    float scale;
    const uint8_t* key_pad;   // [B,S] 1 = ignore key, or NULL
    float* out;               // [B*S, H*64] fp32 or NULL
    __nv_bfloat16* out2;      // [B*S, 2*H*64] bf16 (hi | lo) or NULL
    long long* dbg;           // optional clock64 timeline of CTA 0 (tools/att_timeline.py)
    int fast;                 // am_set_precision(1): single bf16 pass (Q_hi K_hi^T, P_hi V_hi); the lo tiles are not even loaded
    int q0;                   // first query row computed per sample (pipelined kernel only); outputs hold S - q0 rows per sample
};

__device__ __forceinline__ uint32_t smem_u32a(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_a(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32a(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_a(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32a(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32a(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32a(bar)), "r"(parity) : "memory");
        if (ok) break;
        if (++spins > 200000000u) __trap();  // protocol bug: fail the launch instead of hanging the box
    }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32a(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32a(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_a(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32a(bar)) : "memory");
}
__device__ __forceinline__ void fence_before_a() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_a() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32a(uint32_t taddr, uint32_t r[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16a(uint32_t taddr, const uint32_t r[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// named barrier among the 256 softmax threads (barrier 0 is __syncthreads)
__device__ __forceinline__ void softmax_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// (a, b) -> packed bf16x2 (a in the low half) with one cvt; the rounded values come back as floats by shifting
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(a, b);
    lo = pack_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}

// K-major tile (Q, K): rows of 64 bf16 = 128 B, SWIZZLE_128B, 8-row groups 1024 B apart (validated by gemm_tc 64x3)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major tile (V as the B operand of P V: N = head dim contiguous, K = keys): canonical SW128 layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units — 64 d-values contiguous (128 B), 8-key groups SBO = 1024 B apart;
// n = 1 for N = 64 so LBO is unused (cute::UMMA::make_umma_desc<Major::MN>).
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// [V_hi | V_lo] as ONE MN-major B operand with N = 128: two 64-wide SW128 atoms, the second LBO = OFF_VL - OFF_VH bytes after
// the first (same key rows of the lo tile) -> P_hi . [V_hi | V_lo] is a single MMA (the A slice is fetched from TMEM once).
__device__ __forceinline__ uint64_t desc_mnmajor_hilo(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((3 * TILE_BYTES) >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// instruction descriptors (cute::UMMA::InstrDescriptor): F32 accum, BF16 A/B
constexpr uint32_t IDESC_S = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t IDESC_PV = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) /*B MN-major*/ | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t IDESC_PV2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) /*B MN-major*/ | ((uint32_t)((2 * HD) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__global__ void __launch_bounds__(ATT_THREADS, 1)
mha_tc_kernel(const __grid_constant__ CUtensorMap tm, AttParams p) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* kv_full = bars + 0;
    uint64_t* q_full = bars + 1;
    uint64_t* q_empty = bars + 2;
    uint64_t* s_full = bars + 12;   // [3] one per 128-key tile: the row-max pass starts on key tile 0 while tiles 1, 2 are still in the tensor pipe
    uint64_t* o_full = bars + 4;
    uint64_t* acc_free = bars + 5;
    uint64_t* p_ready = bars + 6;   // [6] one per 64-key block
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 16);
    uint32_t* maskbits = tmem_holder + 1;  // [12] bit k of word w: key 32w+k is attendable
    float* xchg = reinterpret_cast<float*>(maskbits + 12);  // [2][128] row max / row sum exchange between the two softmax groups

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int S = p.S, D = p.H * HD;   // model width
    const int nq = (S + QT - 1) / QT;
    const int nblk = (S + 63) / 64;    // 64-key blocks that contain at least one real key

    // warp roles (highest warp id wins the sub-partition arbiter): 0-7 softmax/epilogue, 8 TMEM allocator, 9 key-mask bits,
    // 10 TMA producer, 11 MMA issuer
    if (warp == 10 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm)) : "memory");
    if (warp == 11 && lane == 0) {
        mbar_init_a(kv_full, 1); mbar_init_a(q_full, 1); mbar_init_a(q_empty, 1); mbar_init_a(o_full, 1);
        for (int j = 0; j < 3; ++j) mbar_init_a(&s_full[j], 1);
        mbar_init_a(acc_free, 256);
        for (int j = 0; j < 6; ++j) mbar_init_a(&p_ready[j], 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32a(tmem_holder)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 9 && lane < 12) {  // attendable-key bitmasks (key < S and not padded)
        pdl_wait();  // key_pad may have been written by the kernel just before this one
        uint32_t bits = 0;
        for (int k = 0; k < 32; ++k) {
            int key = lane * 32 + k;
            bool ok = key < S && !(p.key_pad && p.key_pad[(int64_t)b * S + key]);
            bits |= ok ? (1u << k) : 0u;
        }
        maskbits[lane] = bits;
    }
    fence_before_a();
    __syncthreads();
    fence_after_a();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();  // barrier init / TMEM allocation above overlapped the in_proj GEMM's tail
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 320) p.dbg[0] = clock64();

    if (warp == 10) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            // FAST is resolved once per role (a per-MMA `if (p.fast)` in these single-thread loops cost 12 % of the kernel)
            auto produce = [&](auto fc) {
            constexpr bool FAST = decltype(fc)::value;
            // K / V of this (batch, head): hi at column h*64 of the K / V thirds, lo 3*D columns further
            mbar_expect_tx_a(kv_full, (FAST ? 6 : 12) * TILE_BYTES);
            for (int kt = 0; kt < NKT; ++kt) {
                tma_load_3d(smem + OFF_KH + kt * TILE_BYTES, &tm, kv_full, D + h * HD, kt * KT, b);
                if (!FAST) tma_load_3d(smem + OFF_KL + kt * TILE_BYTES, &tm, kv_full, 3 * D + D + h * HD, kt * KT, b);
                tma_load_3d(smem + OFF_VH + kt * TILE_BYTES, &tm, kv_full, 2 * D + h * HD, kt * KT, b);
                if (!FAST) tma_load_3d(smem + OFF_VL + kt * TILE_BYTES, &tm, kv_full, 3 * D + 2 * D + h * HD, kt * KT, b);
            }
            for (int t = 0; t < nq; ++t) {
                if (t > 0) mbar_wait_a(q_empty, (t - 1) & 1);
                mbar_expect_tx_a(q_full, (FAST ? 1 : 2) * TILE_BYTES);
                tma_load_3d(smem + OFF_QH, &tm, q_full, h * HD, t * QT, b);
                if (!FAST) tma_load_3d(smem + OFF_QL, &tm, q_full, 3 * D + h * HD, t * QT, b);
            }
            };
            if (p.fast) produce(std::true_type{}); else produce(std::false_type{});
        }
    } else if (warp == 11) {
        // ===================== MMA issuer (single thread) =====================
        if (lane == 0) {
            auto issue = [&](auto fc) {
            constexpr bool FAST = decltype(fc)::value;
            const uint32_t sb = smem_u32a(smem);
            mbar_wait_a(kv_full, 0);
            if (p.dbg && blockIdx.x == 0) p.dbg[1] = clock64();
            for (int t = 0; t < nq; ++t) {
                const uint32_t pt = t & 1;
                mbar_wait_a(q_full, pt);
                if (p.dbg && blockIdx.x == 0) p.dbg[8 + t] = clock64();
                if (t > 0) mbar_wait_a(acc_free, (t - 1) & 1);  // S and O of the previous tile fully consumed
                fence_after_a();
                // ---- S = Q K^T (3-term split), key tile kt -> TMEM columns [128 kt, 128 kt + 128)
                for (int kt = 0; kt < NKT; ++kt) {
                    if (kt * KT >= S) break;
                    const uint64_t qh = desc_kmajor(sb + OFF_QH), ql = desc_kmajor(sb + OFF_QL);
                    const uint64_t kh = desc_kmajor(sb + OFF_KH + kt * TILE_BYTES), kl = desc_kmajor(sb + OFF_KL + kt * TILE_BYTES);
                    const uint32_t d = tmem_base + S_COL + kt * KT;
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k) {
                        const uint64_t ko = (uint64_t)(k * 2);  // +32 B per K=16 step inside the 128 B swizzle row
                        if (FAST) { umma_ss(d, qh + ko, kh + ko, IDESC_S, k ? 1u : 0u); continue; }
                        umma_ss(d, ql + ko, kh + ko, IDESC_S, k ? 1u : 0u);
                        umma_ss(d, qh + ko, kl + ko, IDESC_S, 1u);
                        umma_ss(d, qh + ko, kh + ko, IDESC_S, 1u);
                    }
                    umma_commit_a(&s_full[kt]);
                }
                umma_commit_a(q_empty);
                if (p.dbg && blockIdx.x == 0) p.dbg[16 + t] = clock64();
                // ---- O = P V per 64-key block, P (bf16 hi | lo) read from TMEM where the softmax warps stored it
                for (int j = 0; j < nblk; ++j) {
                    mbar_wait_a(&p_ready[j], pt);
                    fence_after_a();
                    const uint32_t a_hi = tmem_base + S_COL + j * 64, a_lo = a_hi + 32;
                    const uint32_t voff = (uint32_t)(j * 64) * 128u;  // 64 keys * 128 B rows
#pragma unroll
                    for (int k = 0; k < 4; ++k) {  // 16 keys per MMA: A advances 8 TMEM columns, V advances 16 rows = 2048 B
                        // O[:, 0:128] += P_hi . [V_hi | V_lo]   (one MMA, N = 128);   O[:, 0:64] += P_lo . V_hi   (N = 64)
                        // the epilogue adds the two 64-column halves.  Measured: three N = 64 MMAs per step ran at ~93 cycles each
                        // (the 4 KB A slice comes from TMEM for every MMA); two MMAs fetch it twice instead of three times.
                        const uint64_t vhl = desc_mnmajor_hilo(sb + OFF_VH + voff + k * 2048), vh = desc_mnmajor(sb + OFF_VH + voff + k * 2048);
                        if (FAST) { umma_ts(tmem_base + O_COL, a_hi + k * 8, vh, IDESC_PV, (j | k) ? 1u : 0u); continue; }
                        umma_ts(tmem_base + O_COL, a_hi + k * 8, vhl, IDESC_PV2, (j | k) ? 1u : 0u);
                        umma_ts(tmem_base + O_COL, a_lo + k * 8, vh, IDESC_PV, 1u);
                    }
                }
                umma_commit_a(o_full);
                if (p.dbg && blockIdx.x == 0) p.dbg[24 + t] = clock64();
            }
            };
            if (p.fast) issue(std::true_type{}); else issue(std::false_type{});
        }
    } else if (warp < 8) {
        // ===================== softmax + epilogue =====================
        // Two groups of 4 warps share every query row (thread (q4, lane) of both groups owns row 32 q4 + lane): group g
        // handles the 64-key blocks j = g, g+2, g+4 (row max / exp / P write-back) and the output columns [32 g, 32 g + 32).
        const int q4 = warp & 3, grp = warp >> 2;
        const int r = q4 * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
        for (int t = 0; t < nq; ++t) {
            const uint32_t pt = t & 1;
            const int qi = t * QT + r;
            // pass 1: row max of the scaled scores over attendable keys (own blocks), then exchange with the other group
            float mx = -CUDART_INF_F;
            for (int j = grp; j < nblk; j += 2) {
                mbar_wait_a(&s_full[j >> 1], pt);  // key tile j/2 of S is complete
                fence_after_a();
                if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0 && j == 0) p.dbg[32 + t] = clock64();
#pragma unroll
                for (int hlf = 0; hlf < 2; ++hlf) {
                    uint32_t v[32];
                    tmem_ld32a(lane_addr + S_COL + j * 64 + hlf * 32, v);
                    const uint32_t mb = maskbits[2 * j + hlf];
                    if (mb == 0xffffffffu) {  // warp-uniform fast path: every key of the chunk attendable
#pragma unroll
                        for (int c = 0; c < 32; ++c) mx = fmaxf(mx, __uint_as_float(v[c]));
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            if ((mb >> c) & 1u) mx = fmaxf(mx, __uint_as_float(v[c]));
                    }
                }
            }
            xchg[grp * 128 + r] = mx;  // max of the RAW scores (scale > 0 commutes with max)
            softmax_bar();
            mx = fmaxf(mx, xchg[(grp ^ 1) * 128 + r]);
            softmax_bar();  // xchg is reused for the row sums below
            const float sc2 = p.scale * 1.4426950408889634f;  // exp(s*scale - max*scale) = 2^(s*sc2 - mx2)
            const float mx2 = mx * sc2;
            if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[40 + t] = clock64();
            // pass 2: p = exp(s - max) -> packed bf16 (hi | lo) back into the block's own S columns
            float sum = 0.f;
            for (int j = grp; j < nblk; j += 2) {
                uint32_t v0[32], v1[32];
                tmem_ld32a(lane_addr + S_COL + j * 64, v0);
                tmem_ld32a(lane_addr + S_COL + j * 64 + 32, v1);
#pragma unroll
                for (int hlf = 0; hlf < 2; ++hlf) {
                    const uint32_t* v = hlf ? v1 : v0;
                    const uint32_t mb = maskbits[2 * j + hlf];
                    uint32_t ph[16], pl[16];
                    if (mb == 0xffffffffu) {
#pragma unroll
                        for (int c = 0; c < 32; c += 2) {
                            const float e0 = ex2_approx(fmaf(__uint_as_float(v[c]), sc2, -mx2));
                            const float e1 = ex2_approx(fmaf(__uint_as_float(v[c + 1]), sc2, -mx2));
                            sum += e0 + e1;
                            split_pair(e0, e1, ph[c / 2], pl[c / 2]);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; c += 2) {
                            const float e0 = ((mb >> c) & 1u) ? ex2_approx(fmaf(__uint_as_float(v[c]), sc2, -mx2)) : 0.f;
                            const float e1 = ((mb >> (c + 1)) & 1u) ? ex2_approx(fmaf(__uint_as_float(v[c + 1]), sc2, -mx2)) : 0.f;
                            sum += e0 + e1;
                            split_pair(e0, e1, ph[c / 2], pl[c / 2]);
                        }
                    }
                    // keys 64j + 32 hlf + (2w, 2w+1) -> P_hi word 16 hlf + w (columns [64j, 64j+32)), P_lo 32 columns further
                    tmem_st16a(lane_addr + S_COL + j * 64 + hlf * 16, ph);
                    tmem_st16a(lane_addr + S_COL + j * 64 + 32 + hlf * 16, pl);
                }
                tmem_wait_st();
                fence_before_a();
                mbar_arrive_a(&p_ready[j]);
            }
            if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[48 + t] = clock64();
            xchg[grp * 128 + r] = sum;
            softmax_bar();
            sum += xchg[(grp ^ 1) * 128 + r];
            // epilogue: this group's 32 output columns of O / rowsum
            mbar_wait_a(o_full, pt);
            fence_after_a();
            if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[56 + t] = clock64();
            uint32_t o0[32];
            if (p.fast) {
                tmem_ld32a(lane_addr + O_COL + grp * 32, o0);
            } else {
                uint32_t o1[32];
                tmem_ld32a(lane_addr + O_COL + grp * 32, o0);
                tmem_ld32a(lane_addr + O_COL + HD + grp * 32, o1);   // the P_hi . V_lo term
#pragma unroll
                for (int c = 0; c < 32; ++c) o0[c] = __float_as_uint(__uint_as_float(o0[c]) + __uint_as_float(o1[c]));
            }
            fence_before_a();
            mbar_arrive_a(acc_free);  // S / O columns may be overwritten by the next query tile
            softmax_bar();            // xchg free for the next tile
            if (qi < S) {
                const float inv = 1.0f / sum;
                const int64_t row = (int64_t)b * S + qi;
                if (p.out) {
                    float* dst = p.out + row * D + h * HD + grp * 32;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(o0[j]) * inv, __uint_as_float(o0[j + 1]) * inv,
                                                                          __uint_as_float(o0[j + 2]) * inv, __uint_as_float(o0[j + 3]) * inv);
                }
                if (p.out2) {
                    __nv_bfloat16* hi = p.out2 + row * (2 * (int64_t)D) + h * HD + grp * 32;
                    __nv_bfloat16* lo = hi + D;
                    uint32_t wh[16], wl[16];
#pragma unroll
                    for (int j = 0; j < 32; j += 2) split_pair(__uint_as_float(o0[j]) * inv, __uint_as_float(o0[j + 1]) * inv, wh[j / 2], wl[j / 2]);
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        *reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(hi) + j) = make_uint4(wh[j], wh[j + 1], wh[j + 2], wh[j + 3]);
                        *reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(lo) + j) = make_uint4(wl[j], wl[j + 1], wl[j + 2], wl[j + 3]);
                    }
                }
            }
        }
    }
    fence_before_a();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}


// ---------------------------------------------------------------------------------------------------------------------------------
// Round 2: the same arithmetic, software-pipelined across query tiles.  The round-1 kernel above ran each query
// tile as a strict chain  QK -> row max -> exp/split -> PV -> epilogue  (11.8K cycles per tile of which the tensor pipe was busy
// 4.6K: `tools/probes/mma_rate_probe.cu` measures M128 N128 K16 at 64 cycles and the N128 + N64 TMEM-A pair of one PV k-step at 96,
// i.e. at the tcgen05 floor, so the MMAs themselves were never the problem).  Here
//   * the MMA thread issues QK(t+1, key tile kt) right behind PV(t, last block of key tile kt): tcgen05.mma executes in issue
//     order, so the P columns of tile t are consumed before S of tile t+1 overwrites them — no second S buffer is needed;
//   * the softmax warps compute the row max of tile t+1 BEFORE they wait for O of tile t, which hides the PV tail and the QK of
//     the next tile; the (max, sum) exchanges use separate arrays -> two named barriers per tile instead of four;
//   * tcgen05.ld is issued one 32-column chunk ahead of the exp / split arithmetic; for that P(hi | lo) of a 32-key chunk is
//     written back into the chunk's OWN 32 columns (hi words [0,16), lo words [16,32)), so a store never lands in columns a
//     pending load still needs;
//   * K tiles arrive on their own barriers (QK of the first tile starts after 1/6 of the K/V bytes instead of all of them).
__device__ __forceinline__ bool mbar_test_a(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32a(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// Whole-warp (convergent) variants: every lane of the issuing warp runs the loop and `elect.sync` predicates the instruction
// itself.  Under a divergent `if (lane == 0)` ptxas lowers each tcgen05.mma to an ELECT / R2UR / BRA.U.ANY loop over the active lanes.
__device__ __forceinline__ void umma_ss_w(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts_w(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_w(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32a(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t r[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
// wait for every outstanding tcgen05.ld of this thread; the registers are operands so that no use of them is scheduled above the wait
__device__ __forceinline__ void tmem_ld32_wait(uint32_t r[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
          "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
          "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
          "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
        :: "memory");
}

__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t r[32]) {   // columns [0,16) into r[0..15]
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8a(uint32_t taddr, const uint32_t r[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
mha_tc_pipe_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmo, AttParams p) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* k_full = bars + 0;    // [3] one per 128-key K tile
    uint64_t* v_full = bars + 3;
    uint64_t* q_full = bars + 4;
    uint64_t* q_empty = bars + 5;
    uint64_t* o_full = bars + 6;
    uint64_t* o_free = bars + 7;    // 256 arrivals: both softmax groups have read O of the tile
    uint64_t* s_full = bars + 8;    // [3] one per 128-key tile of S
    uint64_t* p_ready = bars + 11;  // [12] one per 32-key chunk of P
    uint64_t* stage_full = bars + 23;   // 256 arrivals: O of the tile is staged (bf16 hi | lo) in the Q buffer
    uint64_t* q_free = bars + 24;       // the TMA store has read the staged tile: Q of the tile after next may be loaded
    uint64_t* pv_done = bars + 25;      // [3] P V of key tile kt has completed: S of the next query tile may overwrite its columns
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 28);
    float* xmax = reinterpret_cast<float*>(smem + OFF_BAR + 256);   // [2][128]
    float* xsum = xmax + 256;                                         // [2][128]
    uint32_t* maskbits = reinterpret_cast<uint32_t*>(xsum + 256);     // [12] bit k of word w: key 32w+k is attendable
    if (smem - smem_raw > 704) __trap();   // the 1024-byte round-up above may use at most the slack left in ATT_SMEM (768 - 64 for the mask)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int S = p.S, D = p.H * HD;
    const int So = S - p.q0;              // query rows computed (and written, compactly) per sample: rows [q0, S)
    const int nq = (So + QT - 1) / QT;
    const int nkt = (S + KT - 1) / KT;
    const int nch = (S + 31) >> 5;                        // 32-key chunks holding at least one key
    const bool tail16 = S - (nch - 1) * 32 <= 16;        // the last chunk is handled as 16 columns / one k-step

    if (warp == 10 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmo)) : "memory");
    }
    if (warp == 11 && lane == 0) {
        for (int j = 0; j < 3; ++j) { mbar_init_a(&k_full[j], 1); mbar_init_a(&s_full[j], 1); }
        mbar_init_a(v_full, 1); mbar_init_a(q_full, 1); mbar_init_a(q_empty, 1); mbar_init_a(o_full, 1);
        mbar_init_a(o_free, 256);
        mbar_init_a(stage_full, 256); mbar_init_a(q_free, 1);
        for (int j = 0; j < 3; ++j) mbar_init_a(&pv_done[j], 1);
        for (int j = 0; j < 12; ++j) mbar_init_a(&p_ready[j], 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32a(tmem_holder)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 9 && lane < 12) {
        pdl_wait();
        uint32_t bits = 0;
        for (int k = 0; k < 32; ++k) {
            int key = lane * 32 + k;
            bool ok = key < S && !(p.key_pad && p.key_pad[(int64_t)b * S + key]);
            bits |= ok ? (1u << k) : 0u;
        }
        maskbits[lane] = bits;
    }
    fence_before_a();
    __syncthreads();
    fence_after_a();
    // All 512 columns are allocated (one CTA per SM), so the allocation can only start at column 0.  Treating the base as the
    // CONSTANT 0 keeps every TMEM operand of the single-thread MMA loop in uniform registers; with a base loaded from shared memory
    // ptxas wraps each tcgen05.mma in an ELECT / R2UR / branch loop, and the issuing thread (not the tensor pipe) paces the kernel
    // (measured: 96 cycles per MMA issued against a 32-64 cycle floor).
    if (*tmem_holder != 0u) __trap();
    constexpr uint32_t tmem_base = 0u;
    pdl_wait();
    const bool dbg = p.dbg && blockIdx.x == 0;
    if (dbg && threadIdx.x == 320) p.dbg[0] = clock64();

    if (warp == 10) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            auto produce = [&](auto fc) {
                constexpr bool FAST = decltype(fc)::value;
                mbar_expect_tx_a(q_full, (FAST ? 1 : 2) * TILE_BYTES);
                tma_load_3d(smem + OFF_QH, &tm, q_full, h * HD, p.q0, b);
                if (!FAST) tma_load_3d(smem + OFF_QL, &tm, q_full, 3 * D + h * HD, p.q0, b);
                for (int kt = 0; kt < NKT; ++kt) {
                    mbar_expect_tx_a(&k_full[kt], (FAST ? 1 : 2) * TILE_BYTES);
                    tma_load_3d(smem + OFF_KH + kt * TILE_BYTES, &tm, &k_full[kt], D + h * HD, kt * KT, b);
                    if (!FAST) tma_load_3d(smem + OFF_KL + kt * TILE_BYTES, &tm, &k_full[kt], 3 * D + D + h * HD, kt * KT, b);
                }
                mbar_expect_tx_a(v_full, (FAST ? 3 : 6) * TILE_BYTES);
                for (int kt = 0; kt < NKT; ++kt) {
                    tma_load_3d(smem + OFF_VH + kt * TILE_BYTES, &tm, v_full, 2 * D + h * HD, kt * KT, b);
                    if (!FAST) tma_load_3d(smem + OFF_VL + kt * TILE_BYTES, &tm, v_full, 3 * D + 2 * D + h * HD, kt * KT, b);
                }
                for (int t = 1; t < nq; ++t) {
                    // the Q buffer doubles as the staging tile of the epilogue: Q(1) may land once S of tile 0 is complete; Q(t >= 2)
                    // once the TMA store of tile t-2 has read its staged O (the epilogue staged it after S of tile t-1 was complete)
                    if (t == 1) mbar_wait_a(q_empty, 0); else mbar_wait_a(q_free, (t - 2) & 1);
                    mbar_expect_tx_a(q_full, (FAST ? 1 : 2) * TILE_BYTES);
                    tma_load_3d(smem + OFF_QH, &tm, q_full, h * HD, p.q0 + t * QT, b);
                    if (!FAST) tma_load_3d(smem + OFF_QL, &tm, q_full, 3 * D + h * HD, p.q0 + t * QT, b);
                }
            };
            if (p.fast) produce(std::true_type{}); else produce(std::false_type{});
        }
    } else if (warp == 8) {
        // ===================== S = Q K^T issuer (whole warp, elect.sync per instruction) =====================
        // Two issuing warps: this one for Q K^T, warp 11 for P V.  A single issuer was the pacing resource — it shares its scheduler
        // with two softmax warps and spends ~60-110 cycles per MMA against a 32-64 cycle tensor floor, 84 MMAs per query tile.
        // Ordering between the two streams goes through mbarriers: S(t+1) of key tile kt may only overwrite the P(t) columns once
        // P V of that key tile has COMPLETED (pv_done[kt], a tcgen05.commit of the P V warp).
        auto issue = [&](auto fc) {
            constexpr bool FAST = decltype(fc)::value;
            const uint32_t sb = smem_u32a(smem);
            const uint64_t qh = desc_kmajor(sb + OFF_QH), ql = desc_kmajor(sb + OFF_QL);
            // keys of the last key tile rounded up to the MMA's N granularity (16): S = 326 -> 80 columns instead of 128.  The
            // columns beyond are never written; the key mask keeps the softmax away from them.
            const int n_last = min(KT, ((S - (nkt - 1) * KT + 15) >> 4) << 4);
            // S(t)[:, 128 kt : 128 kt + ncols] = Q K_kt^T, 3-term split
            auto qk = [&](int kt, int ncols) {
                const uint64_t kh = desc_kmajor(sb + OFF_KH + kt * TILE_BYTES), kl = desc_kmajor(sb + OFF_KL + kt * TILE_BYTES);
                const uint32_t d = tmem_base + S_COL + kt * KT;
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ncols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) {
                    const uint64_t ko = (uint64_t)(k * 2);
                    if (FAST) { umma_ss_w(d, qh + ko, kh + ko, idesc, k ? 1u : 0u); continue; }
                    umma_ss_w(d, ql + ko, kh + ko, idesc, k ? 1u : 0u);
                    umma_ss_w(d, qh + ko, kl + ko, idesc, 1u);
                    umma_ss_w(d, qh + ko, kh + ko, idesc, 1u);
                }
                umma_commit_w(&s_full[kt]);
            };
            // tile 0: each key tile as soon as its K tile has landed
            mbar_wait_a(q_full, 0);
            if (dbg && lane == 0) p.dbg[8] = clock64();
            for (int kt = 0; kt < nkt; ++kt) {
                mbar_wait_a(&k_full[kt], 0);
                if (dbg && lane == 0 && kt == 0) p.dbg[1] = clock64();
                fence_after_a();
                qk(kt, kt == nkt - 1 ? n_last : KT);
            }
            umma_commit_w(q_empty);
            if (dbg && lane == 0) p.dbg[16] = clock64();
            // tile t+1 behind P V of tile t, key tile by key tile
            for (int t = 1; t < nq; ++t) {
                mbar_wait_a(q_full, t & 1);
                if (dbg && lane == 0) p.dbg[8 + t] = clock64();
                for (int kt = 0; kt < nkt; ++kt) {
                    mbar_wait_a(&pv_done[kt], (t - 1) & 1);
                    fence_after_a();
                    qk(kt, kt == nkt - 1 ? n_last : KT);
                    if (dbg && lane == 0 && t == 2) p.dbg[88 + kt] = clock64();
                }
                if (dbg && lane == 0) p.dbg[16 + t] = clock64();
            }
        };
        if (p.fast) issue(std::true_type{}); else issue(std::false_type{});
    } else if (warp == 11) {
        // ===================== O = P V issuer (whole warp, elect.sync per instruction) =====================
        auto issue = [&](auto fc) {
            constexpr bool FAST = decltype(fc)::value;
            const uint32_t sb = smem_u32a(smem);
            // Every instruction this warp spends between two MMAs delays the tensor pipe (it shares its scheduler with two softmax
            // warps): the V descriptors and the P address advance by constants (32 keys = 4096 B of V rows = +256 in the descriptor's
            // 16-byte address field, 32 TMEM columns).
            const uint64_t vhl0 = desc_mnmajor_hilo(sb + OFF_VH), vh0 = desc_mnmajor(sb + OFF_VH);
            mbar_wait_a(v_full, 0);
            for (int t = 0; t < nq; ++t) {
                const uint32_t pt = t & 1;
                const bool more = t + 1 < nq;
                for (int kt = 0; kt < nkt; ++kt) {
                    // The two softmax groups work on the 64-key blocks 2 kt and 2 kt + 1 side by side, each chunk by chunk: consume
                    // chunks 4 kt + {0, 2, 1, 3} in the order they are produced.  (Assigning CHUNKS alternately to the groups
                    // balances a short last block better but measured slower: 45.9 vs 41.7 us per launch.)
                    const uint64_t vhl_kt = vhl0 + (uint64_t)(kt * 1024), vh_kt = vh0 + (uint64_t)(kt * 1024);
                    const uint32_t a_kt = tmem_base + S_COL + kt * 128;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int cc = ((e & 1) << 1) | (e >> 1);   // 0, 2, 1, 3
                        const int c = 4 * kt + cc;
                        if (c >= nch) continue;
                        mbar_wait_a(&p_ready[c], pt);
                        if (kt == 0 && e == 0 && t > 0) mbar_wait_a(o_free, (t - 1) & 1);   // the epilogue of tile t-1 has read O
                        fence_after_a();
                        if (dbg && lane == 0 && t == 1) p.dbg[64 + c] = clock64();
                        // 16 keys per MMA: P_hi words at column a (+8 for the second k-step), P_lo 16 columns further; V rows 16 further = +128
                        const uint32_t a = a_kt + cc * 32;
                        const uint64_t vhl = vhl_kt + (uint64_t)(cc * 256), vh = vh_kt + (uint64_t)(cc * 256);
                        const bool two = c != nch - 1 || !tail16;   // a last chunk of <= 16 keys is one k-step
                        const uint32_t acc = (kt | e) ? 1u : 0u;
                        if (FAST) {
                            umma_ts_w(tmem_base + O_COL, a, vh, IDESC_PV, acc);
                            if (two) umma_ts_w(tmem_base + O_COL, a + 8, vh + 128, IDESC_PV, 1u);
                        } else {
                            umma_ts_w(tmem_base + O_COL, a, vhl, IDESC_PV2, acc);
                            umma_ts_w(tmem_base + O_COL, a + 16, vh, IDESC_PV, 1u);
                            if (two) {
                                umma_ts_w(tmem_base + O_COL, a + 8, vhl + 128, IDESC_PV2, 1u);
                                umma_ts_w(tmem_base + O_COL, a + 24, vh + 128, IDESC_PV, 1u);
                            }
                        }
                    }
                    if (more) umma_commit_w(&pv_done[kt]);   // the P columns of key tile kt are consumed
                    if (kt == nkt - 1) {
                        umma_commit_w(o_full);
                        if (dbg && lane == 0) p.dbg[24 + t] = clock64();
                    }
                }
            }
        };
        if (p.fast) issue(std::true_type{}); else issue(std::false_type{});
    } else if (warp == 9) {
        // ===================== O store: TMA store of the staged tile =====================
        if (lane == 0) {
            for (int t = 0; t < nq; ++t) {
                mbar_wait_a(stage_full, t & 1);
                if (p.out2) {
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                                 ::"l"(reinterpret_cast<uint64_t>(&tmo)), "r"(smem_u32a(smem + OFF_QH)), "r"(h * HD), "r"(t * QT), "r"(b) : "memory");
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                                 ::"l"(reinterpret_cast<uint64_t>(&tmo)), "r"(smem_u32a(smem + OFF_QL)), "r"(D + h * HD), "r"(t * QT), "r"(b) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                mbar_arrive_a(q_free);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (warp < 8) {
        // ===================== softmax + epilogue =====================
        const int q4 = warp & 3, grp = warp >> 2;
        const int r = q4 * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
        const float sc2 = p.scale * 1.4426950408889634f;
        uint32_t va[32], vb[32];

        // ---- chunk helpers.  Chunk c = keys [32 c, 32 c + 32) = S columns [32 c, 32 c + 32); a 64-key block is two chunks.  A last
        // chunk with <= 16 keys (S = 326: 6 keys) is loaded / processed / stored as 16 columns.
        auto ld_issue = [&](uint32_t (&v)[32], int c) {
            if (c == nch - 1 && tail16) tmem_ld16_issue(lane_addr + S_COL + c * 32, v);
            else tmem_ld32_issue(lane_addr + S_COL + c * 32, v);
        };
        auto chunk_max = [&](const uint32_t (&v)[32], int c, float mx) {
            const uint32_t mb = maskbits[c];
            if (mb == 0xffffffffu) {
#pragma unroll
                for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(v[e]));
            } else if (c == nch - 1 && tail16) {
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if ((mb >> e) & 1u) mx = fmaxf(mx, __uint_as_float(v[e]));
            } else {
#pragma unroll
                for (int e = 0; e < 32; ++e)
                    if ((mb >> e) & 1u) mx = fmaxf(mx, __uint_as_float(v[e]));
            }
            return mx;
        };
        // row max of the raw scores of tile t over this group's 64-key blocks in [jlo, jhi) (block j = chunks 2 j, 2 j + 1; group g owns
        // the blocks j % 2 == g), tcgen05.ld one chunk ahead of the fmax chain
        auto pass1 = [&](int t, int jlo, int jhi, float mx) {
            const uint32_t pt = t & 1;
            int j = jlo + ((jlo ^ grp) & 1);
            if (j >= jhi) return mx;
            mbar_wait_a(&s_full[j >> 1], pt);
            fence_after_a();
            if (dbg && threadIdx.x == 0 && jlo == 0) p.dbg[32 + t] = clock64();
            ld_issue(va, 2 * j);
            for (; j < jhi; j += 2) {
                const int c0 = 2 * j, c1 = c0 + 1, jn = j + 2;
                const bool has1 = c1 < nch;
                tmem_ld32_wait(va);
                if (has1) ld_issue(vb, c1);
                mx = chunk_max(va, c0, mx);
                if (has1) tmem_ld32_wait(vb);
                bool pre = false;
                if (jn < jhi && mbar_test_a(&s_full[jn >> 1], pt)) {   // next block's key tile already complete: prefetch under the fmax chain
                    fence_after_a();
                    ld_issue(va, 2 * jn);
                    pre = true;
                }
                if (has1) mx = chunk_max(vb, c1, mx);
                if (jn < jhi && !pre) {
                    mbar_wait_a(&s_full[jn >> 1], pt);
                    fence_after_a();
                    ld_issue(va, 2 * jn);
                }
            }
            return mx;
        };
        // p = 2^(s*sc2 - mx2) of one chunk -> packed bf16 hi words [0,16) | lo words [16,32) of the chunk's own columns
        auto chunk_exp = [&](const uint32_t (&v)[32], int c, float mx2, float& sum) {
            const uint32_t mb = maskbits[c];
            if (c == nch - 1 && tail16) {
                uint32_t ph[8], pl[8];
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    const float e0 = ((mb >> e) & 1u) ? ex2_approx(fmaf(__uint_as_float(v[e]), sc2, -mx2)) : 0.f;
                    const float e1 = ((mb >> (e + 1)) & 1u) ? ex2_approx(fmaf(__uint_as_float(v[e + 1]), sc2, -mx2)) : 0.f;
                    sum += e0 + e1;
                    split_pair(e0, e1, ph[e / 2], pl[e / 2]);
                }
                tmem_st8a(lane_addr + S_COL + c * 32, ph);
                tmem_st8a(lane_addr + S_COL + c * 32 + 16, pl);
            } else {
                uint32_t ph[16], pl[16];
                if (mb == 0xffffffffu) {
#pragma unroll
                    for (int e = 0; e < 32; e += 2) {
                        const float e0 = ex2_approx(fmaf(__uint_as_float(v[e]), sc2, -mx2));
                        const float e1 = ex2_approx(fmaf(__uint_as_float(v[e + 1]), sc2, -mx2));
                        sum += e0 + e1;
                        split_pair(e0, e1, ph[e / 2], pl[e / 2]);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 32; e += 2) {
                        const float e0 = ((mb >> e) & 1u) ? ex2_approx(fmaf(__uint_as_float(v[e]), sc2, -mx2)) : 0.f;
                        const float e1 = ((mb >> (e + 1)) & 1u) ? ex2_approx(fmaf(__uint_as_float(v[e + 1]), sc2, -mx2)) : 0.f;
                        sum += e0 + e1;
                        split_pair(e0, e1, ph[e / 2], pl[e / 2]);
                    }
                }
                tmem_st16a(lane_addr + S_COL + c * 32, ph);
                tmem_st16a(lane_addr + S_COL + c * 32 + 16, pl);
            }
            tmem_wait_st();
            fence_before_a();
            mbar_arrive_a(&p_ready[c]);
        };

        const int nblk = (nch + 1) >> 1;      // 64-key blocks holding at least one key
        float mx_own = pass1(0, 0, nblk, -CUDART_INF_F);
        const int jsplit = min(nblk, 4);      // blocks of key tiles 0, 1 | key tile 2
        for (int t = 0; t < nq; ++t) {
            const uint32_t pt = t & 1;
            const int qi = t * QT + r;
            xmax[grp * 128 + r] = mx_own;
            softmax_bar();
            const float mx2 = fmaxf(mx_own, xmax[(grp ^ 1) * 128 + r]) * sc2;
            if (dbg && threadIdx.x == 0) p.dbg[40 + t] = clock64();
            // pass 2
            float sum = 0.f;
            if (grp < nblk) ld_issue(va, 2 * grp);
            for (int j = grp; j < nblk; j += 2) {
                const int c0 = 2 * j, c1 = c0 + 1;
                const bool has1 = c1 < nch;
                tmem_ld32_wait(va);
                if (has1) ld_issue(vb, c1);
                chunk_exp(va, c0, mx2, sum);
                if (has1) tmem_ld32_wait(vb);
                if (j + 2 < nblk) ld_issue(va, 2 * (j + 2));
                if (has1) chunk_exp(vb, c1, mx2, sum);
                if (dbg && t == 1 && (threadIdx.x & 127) == 0) p.dbg[96 + j] = clock64();
            }
            if (dbg && threadIdx.x == 0) p.dbg[48 + t] = clock64();
            // row max of the NEXT tile over key tiles 0, 1 (their S is complete: it was issued behind P V of chunks 0-7) while the
            // tensor pipe finishes P V of this tile and S of key tile 2
            if (t + 1 < nq) mx_own = pass1(t + 1, 0, jsplit, -CUDART_INF_F);
            xsum[grp * 128 + r] = sum;
            softmax_bar();
            sum += xsum[(grp ^ 1) * 128 + r];
            // epilogue: this group's 32 output columns of O / rowsum
            mbar_wait_a(o_full, pt);
            fence_after_a();
            if (dbg && threadIdx.x == 0) p.dbg[56 + t] = clock64();
            if (p.fast) {
                tmem_ld32_issue(lane_addr + O_COL + grp * 32, va);
                tmem_ld32_wait(va);
            } else {
                tmem_ld32_issue(lane_addr + O_COL + grp * 32, va);
                tmem_ld32_issue(lane_addr + O_COL + HD + grp * 32, vb);   // the P_hi . V_lo term
                tmem_ld32_wait(va);
                tmem_ld32_wait(vb);
#pragma unroll
                for (int c = 0; c < 32; ++c) va[c] = __float_as_uint(__uint_as_float(va[c]) + __uint_as_float(vb[c]));
            }
            fence_before_a();
            mbar_arrive_a(o_free);
            if (dbg && (threadIdx.x & 127) == 0) p.dbg[104 + 4 * t + grp] = clock64();
            if (qi < So) {
                const float inv = 1.0f / sum;
                const int64_t row = (int64_t)b * So + qi;
                if (p.out) {
                    float* dst = p.out + row * D + h * HD + grp * 32;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(va[j]) * inv, __uint_as_float(va[j + 1]) * inv,
                                                                          __uint_as_float(va[j + 2]) * inv, __uint_as_float(va[j + 3]) * inv);
                }
            }
            {
                // bf16 (hi | lo) of O / rowsum staged in the Q buffer (SWIZZLE_128B rows of 64 bf16: 16-byte chunk c of row r sits at
                // chunk c ^ (r % 8)) and written by TMA: with one row per lane a direct store touches 32 cache lines per instruction
                // and the 8 x 16-byte stores per thread cost 2-3.4K cycles per tile (measured); rows >= S are clipped by the tensor map.
                if (t + 1 < nq) mbar_wait_a(&s_full[nkt - 1], (t + 1) & 1);   // S of the next tile is complete: Q is no longer being read
                if (t > 0) mbar_wait_a(q_free, (t - 1) & 1);                   // the previous tile's store has read its staging
                if (p.out2) {
                    const float inv = 1.0f / sum;
                    uint32_t wh[16], wl[16];
#pragma unroll
                    for (int j = 0; j < 32; j += 2) split_pair(__uint_as_float(va[j]) * inv, __uint_as_float(va[j + 1]) * inv, wh[j / 2], wl[j / 2]);
                    uint8_t* sh = smem + OFF_QH + r * 128;
                    uint8_t* sl = smem + OFF_QL + r * 128;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int off = ((grp * 4 + c) ^ (r & 7)) << 4;
                        *reinterpret_cast<uint4*>(sh + off) = make_uint4(wh[4 * c], wh[4 * c + 1], wh[4 * c + 2], wh[4 * c + 3]);
                        *reinterpret_cast<uint4*>(sl + off) = make_uint4(wl[4 * c], wl[4 * c + 1], wl[4 * c + 2], wl[4 * c + 3]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                }
                mbar_arrive_a(stage_full);
            }
            if (dbg && (threadIdx.x & 127) == 0) p.dbg[104 + 4 * t + 2 + grp] = clock64();
            if (t + 1 < nq) mx_own = pass1(t + 1, jsplit, nblk, mx_own);   // key tile 2 of the next tile
        }
    }
    fence_before_a();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

typedef CUresult (*EncodeTiledFnA)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

static long long* g_att_dbg = nullptr;
extern "C" void am_att_set_debug_(void* buf) { g_att_dbg = reinterpret_cast<long long*>(buf); }  // debug hook, not in the public header

extern "C" int am_mha_tc_fwd_rows(const void* qkv2, float* out, void* out2, const uint8_t* key_pad, int B, int S, int H, int hd, float scale,
                                  int q_row0, am_stream_t stream) {
    AM_REQUIRE(qkv2 && (out || out2) && B > 0 && S > 0 && H > 0, AM_EINVAL, "am_mha_tc_fwd: bad args");
    AM_REQUIRE(q_row0 >= 0 && q_row0 < S, AM_EINVAL, "am_mha_tc_fwd_rows: q_row0 must be in [0, S)");
    AM_REQUIRE(hd == HD, AM_EINVAL, "am_mha_tc_fwd: head dim must be 64");
    AM_REQUIRE(S <= NKT * KT, AM_EINVAL, "am_mha_tc_fwd: S must be <= 384 (whole key row lives in TMEM)");
    AM_REQUIRE((reinterpret_cast<uintptr_t>(qkv2) & 15u) == 0 && (!out || (reinterpret_cast<uintptr_t>(out) & 15u) == 0) &&
               (!out2 || (reinterpret_cast<uintptr_t>(out2) & 15u) == 0), AM_EALIGN, "am_mha_tc_fwd: 16-byte alignment required");
    static EncodeTiledFnA enc = nullptr;
    if (!enc) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            am_set_error_("am_mha_tc_fwd: cuTensorMapEncodeTiled unavailable");
            return AM_ELAUNCH;
        }
        enc = reinterpret_cast<EncodeTiledFnA>(fp);
    }
    const uint64_t cols = (uint64_t)2 * 3 * H * HD;  // (hi | lo) x (q | k | v)
    CUtensorMap tm;
    cuuint64_t gdim[3] = {cols, (cuuint64_t)S, (cuuint64_t)B};
    cuuint64_t gstr[2] = {cols * 2, cols * 2 * (cuuint64_t)S};
    cuuint32_t box[3] = {HD, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv2), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        am_set_error_("am_mha_tc_fwd: cuTensorMapEncodeTiled failed");
        return AM_ELAUNCH;
    }
    static bool attr = false;
    static bool pipe = true;   // AMB200_ATTN_PIPE=0: the round-1 kernel (one strict chain per query tile), kept for A/B timing
    if (!attr) {
        const char* e = getenv("AMB200_ATTN_PIPE");
        pipe = !(e && e[0] == '0');
        if (cudaFuncSetAttribute(mha_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(mha_tc_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM) != cudaSuccess) {
            am_set_error_("am_mha_tc_fwd: shared memory opt-in failed");
            return AM_ELAUNCH;
        }
        attr = true;
    }
    AM_REQUIRE(pipe || q_row0 == 0, AM_EINVAL, "am_mha_tc_fwd_rows: q_row0 > 0 needs the pipelined kernel (AMB200_ATTN_PIPE=0 is set)");
    AttParams p{B, S, H, scale, key_pad, out, reinterpret_cast<__nv_bfloat16*>(out2), g_att_dbg, am_get_precision(), q_row0};
    CUtensorMap tmo = tm;   // store map of the bf16 (hi | lo) output [B, S, 2 H 64]: 128-row x 64-column boxes, rows >= S clipped
    if (pipe && out2) {
        const uint64_t ocols = (uint64_t)2 * H * HD;
        cuuint64_t odim[3] = {ocols, (cuuint64_t)(S - q_row0), (cuuint64_t)B};
        cuuint64_t ostr[2] = {ocols * 2, ocols * 2 * (cuuint64_t)(S - q_row0)};
        if (enc(&tmo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out2, odim, ostr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            am_set_error_("am_mha_tc_fwd: cuTensorMapEncodeTiled (output) failed");
            return AM_ELAUNCH;
        }
    }
    if (pipe) am_launch(mha_tc_pipe_kernel, dim3(B * H), dim3(ATT_THREADS), ATT_SMEM, as_stream(stream), 1, tm, tmo, p);
    else am_launch(mha_tc_kernel, dim3(B * H), dim3(ATT_THREADS), ATT_SMEM, as_stream(stream), 1, tm, p);
    AM_LAUNCH_CHECK("mha_tc_fwd");
    return AM_OK;
}

extern "C" int am_mha_tc_fwd(const void* qkv2, float* out, void* out2, const uint8_t* key_pad, int B, int S, int H, int hd, float scale,
                             am_stream_t stream) {
    return am_mha_tc_fwd_rows(qkv2, out, out2, key_pad, B, S, H, hd, scale, 0, stream);
}
