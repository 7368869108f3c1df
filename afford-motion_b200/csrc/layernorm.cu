// Y = LayerNorm(X (+R)) * gamma + beta — one warp per row, row cached in registers (D <= 1024),
// two-pass mean / variance in fp32 (matches torch.nn.LayerNorm, eps inside the sqrt).
#include <cuda_bf16.h>
#include "common.cuh"

namespace {

template <int MAXV>  // MAXV = ceil(D/32) upper bound
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ R, int ldr,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float* __restrict__ Y, int ldy, int M, int D, float eps,
                                                        __nv_bfloat16* __restrict__ Y2, int Np2) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    int lane = threadIdx.x & 31;
    const float* x = X + (int64_t)row * ldx;
    const float* r = R ? R + (int64_t)row * ldr : nullptr;
    float v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        int d = lane + i * 32;
        float t = 0.f;
        if (d < D) { t = x[d]; if (r) t += r[d]; }
        v[i] = t; s += t;
    }
    float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        int d = lane + i * 32;
        float t = d < D ? v[i] - mean : 0.f;
        q += t * t;
    }
    float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
    float* y = Y ? Y + (int64_t)row * ldy : nullptr;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        int d = lane + i * 32;
        if (d < D) {
            float o = (v[i] - mean) * rstd * gamma[d] + beta[d];
            if (Y) y[d] = o;
            if (Y2) {  // bf16 (hi | lo) operand of the next tcgen05 GEMM
                __nv_bfloat16 h = __float2bfloat16_rn(o);
                Y2[(int64_t)row * 2 * Np2 + d] = h;
                Y2[(int64_t)row * 2 * Np2 + Np2 + d] = __float2bfloat16_rn(o - __bfloat162float(h));
            }
        } else if (Y2 && d < Np2) {
            Y2[(int64_t)row * 2 * Np2 + d] = __float2bfloat16_rn(0.f);
            Y2[(int64_t)row * 2 * Np2 + Np2 + d] = __float2bfloat16_rn(0.f);
        }
    }
}

}  // namespace

extern "C" int am_layernorm(const float* X, int ldx, const float* R, int ldr, const float* gamma, const float* beta, float* Y, int ldy,
                            int M, int D, float eps, void* Y2, int Np2, am_stream_t stream) {
    AM_REQUIRE(X && gamma && beta && (Y || Y2) && M > 0 && D > 0 && D <= 1024, AM_EINVAL, "am_layernorm: bad args (D <= 1024)");
    AM_REQUIRE(ldx >= D && (!Y || ldy >= D) && (!R || ldr >= D), AM_EINVAL, "am_layernorm: bad strides");
    AM_REQUIRE(!Y2 || (Np2 >= D && Np2 % 32 == 0 && Np2 <= 1024), AM_EINVAL, "am_layernorm: bad Np2");
    __nv_bfloat16* y2 = reinterpret_cast<__nv_bfloat16*>(Y2);
    int grid = cdiv(M, 8);
    if (D <= 256) layernorm_kernel<8><<<grid, 256, 0, as_stream(stream)>>>(X, ldx, R, ldr, gamma, beta, Y, ldy, M, D, eps, y2, Np2);
    else if (D <= 512) layernorm_kernel<16><<<grid, 256, 0, as_stream(stream)>>>(X, ldx, R, ldr, gamma, beta, Y, ldy, M, D, eps, y2, Np2);
    else layernorm_kernel<32><<<grid, 256, 0, as_stream(stream)>>>(X, ldx, R, ldr, gamma, beta, Y, ldy, M, D, eps, y2, Np2);
    AM_LAUNCH_CHECK("layernorm");
    return AM_OK;
}
