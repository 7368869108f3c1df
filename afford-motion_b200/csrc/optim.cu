// Fused flat-buffer AdamW: one launch updates EVERY parameter of the model.
// Replaces the ~100 foreach kernels of torch.optim.AdamW that utils/training.py:48-50,154 drives once per step.
// Parameters, gradients and both moments live in four flat fp32 buffers (amb200.optim.FusedAdamW lays them out);
// HBM-bound: 16 B read + 12 B written per element, 128-bit accesses, grid = 148 SMs x 8 CTAs, grid-stride.
#include "common.cuh"

namespace {

// torch.optim.AdamW, single-tensor form (decoupled weight decay, no amsgrad, not maximize):
//   p *= 1 - lr*wd;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;
//   p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)         bc1 = 1-b1^t, bc2 = 1-b2^t
__global__ void adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                  int64_t n, float lr, float b1, float b2, float omb1, float omb2, float eps, float wd, float step_size, float inv_sqrt_bc2,
                                  float grad_scale, int zero_grad, float* __restrict__ g_rw) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const float decay = 1.0f - lr * wd;  // lr*wd <= 1e-5 in every reference config: fp32 is exact enough here
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P = reinterpret_cast<float4*>(p)[i], G = reinterpret_cast<const float4*>(g)[i];
        float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
        float pe[4] = {P.x, P.y, P.z, P.w}, ge[4] = {G.x, G.y, G.z, G.w}, me[4] = {M.x, M.y, M.z, M.w}, ve[4] = {V.x, V.y, V.z, V.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gj = ge[j] * grad_scale;
            pe[j] *= decay;
            me[j] = b1 * me[j] + omb1 * gj;
            ve[j] = b2 * ve[j] + omb2 * gj * gj;
            pe[j] -= step_size * (me[j] / (sqrtf(ve[j]) * inv_sqrt_bc2 + eps));
        }
        reinterpret_cast<float4*>(p)[i] = make_float4(pe[0], pe[1], pe[2], pe[3]);
        reinterpret_cast<float4*>(m)[i] = make_float4(me[0], me[1], me[2], me[3]);
        reinterpret_cast<float4*>(v)[i] = make_float4(ve[0], ve[1], ve[2], ve[3]);
        if (zero_grad) reinterpret_cast<float4*>(g_rw)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // tail (n % 4 elements)
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gj = g[i] * grad_scale;
        float pj = p[i] * decay;
        const float mj = b1 * m[i] + omb1 * gj;
        const float vj = b2 * v[i] + omb2 * gj * gj;
        pj -= step_size * (mj / (sqrtf(vj) * inv_sqrt_bc2 + eps));
        p[i] = pj; m[i] = mj; v[i] = vj;
        if (zero_grad) g_rw[i] = 0.f;
    }
}

}  // namespace

extern "C" int am_adamw_flat(float* p, float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2, double eps,
                             double weight_decay, int64_t step, float grad_scale, int zero_grad, am_stream_t stream) {
    AM_REQUIRE(p && g && m && v && n > 0 && step >= 1, AM_EINVAL, "am_adamw_flat: bad args");
    AM_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v)) & 15u) == 0, AM_EALIGN, "am_adamw_flat: buffers must be 16-byte aligned");
    // bias corrections in double on the host, like torch (step is a host integer here)
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    // 1-beta in double then rounded once, like torch's scalar arguments (1.0f - 0.999f is off by 4.7e-5 relative)
    const float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
    const float step_size = (float)(lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    int64_t blocks = ((n >> 2) + 255) / 256;
    int grid = (int)(blocks < (int64_t)AM_NUM_SMS * 8 ? (blocks > 0 ? blocks : 1) : (int64_t)AM_NUM_SMS * 8);
    adamw_flat_kernel<<<grid, 256, 0, as_stream(stream)>>>(p, g, m, v, n, (float)lr, (float)beta1, (float)beta2, omb1, omb2, (float)eps, (float)weight_decay, step_size, inv_sqrt_bc2,
                                                           grad_scale, zero_grad, g);
    AM_LAUNCH_CHECK("adamw_flat");
    return AM_OK;
}
