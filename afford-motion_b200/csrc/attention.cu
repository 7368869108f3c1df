// Multi-head self-attention core (fp32 SIMT, exact softmax over the whole key row): replaces the
// SDPA / native-MHA library call inside torch.nn.TransformerEncoderLayer (models/cmdm.py:66-77,167).
// One CTA per (batch, head, 128-query tile); K^T and V of the head are staged once in shared memory
// (S <= 352 keeps K^T + V + P inside the 227 KB carve-out); each warp processes 4 queries at a time:
// scores with lanes over keys, softmax in registers, P through smem, PV with lanes over head dims.
#include <math_constants.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace {

constexpr int HD = 64;
constexpr int ATT_WARPS = 8;
constexpr int QW = 4;        // queries per warp pass
constexpr int QTILE = 128;   // queries per CTA
constexpr int MAXJ = 11;     // ceil(352/32)

__global__ void __launch_bounds__(ATT_WARPS * 32, 1)
mha_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out, const uint8_t* __restrict__ key_pad, int S, int H, float scale,
               __nv_bfloat16* __restrict__ out2) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float smem[];
    const int SP = ((S + 31) / 32) * 32 + 1;
    float* Kt = smem;                         // [HD][SP]
    float* Vs = Kt + HD * SP;                 // [S][HD]
    float* Ps = Vs + ((S * HD + 3) / 4) * 4;  // [ATT_WARPS][SP-1][QW]
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QTILE;
    const int D3 = 3 * H * HD;
    const float* base = qkv + (int64_t)b * S * D3;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // stage K^T and V (float4 along the head dim)
    for (int i = tid; i < S * (HD / 4); i += blockDim.x) {
        int key = i / (HD / 4), d4 = (i % (HD / 4)) * 4;
        float4 kv = *reinterpret_cast<const float4*>(base + (int64_t)key * D3 + H * HD + h * HD + d4);
        Kt[(d4 + 0) * SP + key] = kv.x; Kt[(d4 + 1) * SP + key] = kv.y; Kt[(d4 + 2) * SP + key] = kv.z; Kt[(d4 + 3) * SP + key] = kv.w;
        float4 vv = *reinterpret_cast<const float4*>(base + (int64_t)key * D3 + 2 * H * HD + h * HD + d4);
        *reinterpret_cast<float4*>(Vs + key * HD + d4) = vv;
    }
    __syncthreads();

    const int nj = (S + 31) / 32;
    float* Pw = Ps + (int64_t)warp * (SP - 1) * QW;
    const uint8_t* pad = key_pad ? key_pad + (int64_t)b * S : nullptr;

    for (int qb = q0 + warp * QW; qb < min(q0 + QTILE, S); qb += ATT_WARPS * QW) {
        // each lane keeps q[qi][d = lane], q[qi][d = lane+32] (pre-scaled); broadcast by shuffle
        float qlo[QW], qhi[QW];
#pragma unroll
        for (int qi = 0; qi < QW; ++qi) {
            int q = min(qb + qi, S - 1);
            const float* qp = base + (int64_t)q * D3 + h * HD;
            qlo[qi] = qp[lane] * scale; qhi[qi] = qp[lane + 32] * scale;
        }
        float acc[QW][MAXJ];
#pragma unroll
        for (int qi = 0; qi < QW; ++qi)
#pragma unroll
            for (int j = 0; j < MAXJ; ++j) acc[qi][j] = 0.f;
#pragma unroll 4
        for (int d = 0; d < HD; ++d) {
            float qd[QW];
#pragma unroll
            for (int qi = 0; qi < QW; ++qi) qd[qi] = __shfl_sync(0xffffffffu, d < 32 ? qlo[qi] : qhi[qi], d & 31);
            const float* kr = Kt + d * SP + lane;
#pragma unroll
            for (int j = 0; j < MAXJ; ++j) {
                if (j < nj) {
                    float kv = kr[j * 32];  // columns >= S hold stale smem, masked below
#pragma unroll
                    for (int qi = 0; qi < QW; ++qi) acc[qi][j] = fmaf(qd[qi], kv, acc[qi][j]);
                }
            }
        }
        // mask + softmax (row max / sum across the warp)
#pragma unroll
        for (int qi = 0; qi < QW; ++qi) {
            float mx = -CUDART_INF_F;
#pragma unroll
            for (int j = 0; j < MAXJ; ++j) {
                int key = lane + j * 32;
                bool ok = j < nj && key < S && !(pad && pad[key]);
                acc[qi][j] = ok ? acc[qi][j] : -CUDART_INF_F;
                mx = fmaxf(mx, acc[qi][j]);
            }
            mx = warp_max(mx);
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < MAXJ; ++j) {
                float e = acc[qi][j] == -CUDART_INF_F ? 0.f : expf(acc[qi][j] - mx);
                acc[qi][j] = e; sum += e;
            }
            sum = warp_sum(sum);
            float inv = 1.0f / sum;
#pragma unroll
            for (int j = 0; j < MAXJ; ++j) acc[qi][j] *= inv;
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
            int key = lane + j * 32;
            if (j < nj && key < S) *reinterpret_cast<float4*>(Pw + key * QW) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
        }
        __syncwarp();
        // PV: lane owns head dims 2*lane, 2*lane+1
        float o[QW][2];
#pragma unroll
        for (int qi = 0; qi < QW; ++qi) o[qi][0] = o[qi][1] = 0.f;
#pragma unroll 4
        for (int key = 0; key < S; ++key) {
            float4 p4 = *reinterpret_cast<const float4*>(Pw + key * QW);
            float2 v2 = *reinterpret_cast<const float2*>(Vs + key * HD + 2 * lane);
            o[0][0] = fmaf(p4.x, v2.x, o[0][0]); o[0][1] = fmaf(p4.x, v2.y, o[0][1]);
            o[1][0] = fmaf(p4.y, v2.x, o[1][0]); o[1][1] = fmaf(p4.y, v2.y, o[1][1]);
            o[2][0] = fmaf(p4.z, v2.x, o[2][0]); o[2][1] = fmaf(p4.z, v2.y, o[2][1]);
            o[3][0] = fmaf(p4.w, v2.x, o[3][0]); o[3][1] = fmaf(p4.w, v2.y, o[3][1]);
        }
#pragma unroll
        for (int qi = 0; qi < QW; ++qi) {
            int q = qb + qi;
            if (q < S) {
                int64_t row = (int64_t)b * S + q;
                if (out) *reinterpret_cast<float2*>(out + row * (H * HD) + h * HD + 2 * lane) = make_float2(o[qi][0], o[qi][1]);
                if (out2) {  // bf16 (hi | lo) operand of the out_proj tcgen05 GEMM; row stride 2*H*HD
                    __nv_bfloat16 h0 = __float2bfloat16_rn(o[qi][0]), h1 = __float2bfloat16_rn(o[qi][1]);
                    __nv_bfloat162 hi2; hi2.x = h0; hi2.y = h1;
                    __nv_bfloat162 lo2; lo2.x = __float2bfloat16_rn(o[qi][0] - __bfloat162float(h0)); lo2.y = __float2bfloat16_rn(o[qi][1] - __bfloat162float(h1));
                    __nv_bfloat16* base2 = out2 + row * (2 * H * HD) + h * HD + 2 * lane;
                    *reinterpret_cast<__nv_bfloat162*>(base2) = hi2;
                    *reinterpret_cast<__nv_bfloat162*>(base2 + H * HD) = lo2;
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace

extern "C" int am_mha_fwd(const float* qkv, float* out, const uint8_t* key_pad, int B, int S, int H, int hd, float scale, void* out2,
                          am_stream_t stream) {
    AM_REQUIRE(qkv && (out || out2) && B > 0 && S > 0 && H > 0, AM_EINVAL, "am_mha_fwd: bad args");
    AM_REQUIRE(hd == HD, AM_EINVAL, "am_mha_fwd: head dim must be 64");
    AM_REQUIRE(S <= 32 * MAXJ, AM_EINVAL, "am_mha_fwd: S must be <= 352");
    AM_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 7u) == 0, AM_EALIGN, "am_mha_fwd: alignment");
    int SP = ((S + 31) / 32) * 32 + 1;
    size_t smem = sizeof(float) * ((size_t)HD * SP + ((size_t)S * HD + 3) / 4 * 4 + (size_t)ATT_WARPS * (SP - 1) * QW);
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(mha_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            am_set_error_("am_mha_fwd: cannot opt in to 227 KB shared memory");
            return AM_ELAUNCH;
        }
        attr_set = true;
    }
    AM_REQUIRE(smem <= 227 * 1024, AM_EINVAL, "am_mha_fwd: sequence too long for the shared-memory staging");
    dim3 grid(cdiv(S, QTILE), H, B);
    am_launch(mha_fwd_kernel, dim3(grid), dim3(ATT_WARPS * 32), smem, as_stream(stream), 1, qkv, out, key_pad, S, H, scale, reinterpret_cast<__nv_bfloat16*>(out2));
    AM_LAUNCH_CHECK("mha_fwd");
    return AM_OK;
}
