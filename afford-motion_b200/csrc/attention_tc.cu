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
constexpr int ATT_SMEM = OFF_BAR + 256 + 1024 + 1024;  // barriers + mask bits, row max/sum exchange [2][128] fp32, alignment slack

struct AttParams {
    int B, S, H;
    float scale;
    const uint8_t* key_pad;   // [B,S] 1 = ignore key, or NULL
    float* out;               // [B*S, H*64] fp32 or NULL
    __nv_bfloat16* out2;      // [B*S, 2*H*64] bf16 (hi | lo) or NULL
    long long* dbg;           // optional clock64 timeline of CTA 0 (tools/att_timeline.py)
    int fast;                 // am_set_precision(1): single bf16 pass (Q_hi K_hi^T, P_hi V_hi); the lo tiles are not even loaded
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

typedef CUresult (*EncodeTiledFnA)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

static long long* g_att_dbg = nullptr;
extern "C" void am_att_set_debug_(void* buf) { g_att_dbg = reinterpret_cast<long long*>(buf); }  // debug hook, not in the public header

extern "C" int am_mha_tc_fwd(const void* qkv2, float* out, void* out2, const uint8_t* key_pad, int B, int S, int H, int hd, float scale,
                             am_stream_t stream) {
    AM_REQUIRE(qkv2 && (out || out2) && B > 0 && S > 0 && H > 0, AM_EINVAL, "am_mha_tc_fwd: bad args");
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
    if (!attr) {
        if (cudaFuncSetAttribute(mha_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM) != cudaSuccess) {
            am_set_error_("am_mha_tc_fwd: shared memory opt-in failed");
            return AM_ELAUNCH;
        }
        attr = true;
    }
    AttParams p{B, S, H, scale, key_pad, out, reinterpret_cast<__nv_bfloat16*>(out2), g_att_dbg, am_get_precision()};
    am_launch(mha_tc_kernel, dim3(B * H), dim3(ATT_THREADS), ATT_SMEM, as_stream(stream), 1, tm, p);
    AM_LAUNCH_CHECK("mha_tc_fwd");
    return AM_OK;
}
