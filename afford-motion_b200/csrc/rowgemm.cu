// Tall-skinny fp32 GEMMs of the Point-Transformer contact encoder in TRAINING (models/scene_models/pointtransformer.py:9-38
// per-neighbour MLPs: linear_p 3->3->c, linear_w c->c/8->c/8 on [n*k, .] grouped tensors with n*k up to 2.1 M rows).
// The tiled SIMT kernels (gemm_f32_kernel / gemm_general_kernel, 64- or 128-wide tiles) waste 75-97 % of every tile when one
// GEMM dimension is 3..32; measured: these GEMMs ran at ~2 TFLOP/s and were 54 % of the CMDM training step.
//   rowgemm_fwd_kernel : Y[M,N] = act(X[M,K] Weff^T + b) (+res), N <= 32 — one thread per row, Weff ([N][K] or its transpose
//                        for dX = dY W) in shared memory as [k][NP] so four outputs cost one 128-bit broadcast read
//   rowgemm_dw_kernel  : C[P,Q] = A[R,P]^T B[R,Q] over R >= 8192 rows with min(P,Q) <= 32 (weight gradients dW = dY^T X):
//                        the wide side is spread over threadIdx.x (coalesced row reads), the narrow side lives in registers,
//                        rows are split over threadIdx.y and CTAs, partials meet in shared memory and then in fp32 atomics
// Both are HBM-bound streaming kernels (each input row is read once).
#include "common.cuh"

namespace {

template <int NP>
__global__ void __launch_bounds__(256)
rowgemm_fwd_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw, int transW, float* __restrict__ Y, int ldy,
                   int M, int N, int K, const float* __restrict__ bias, int act, const float* __restrict__ residual, int ldr, int vecX) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float ws[];  // [nchunk][K][NP]: ws[(c*K + k)*NP + n] = Weff[c*NP + n][k]
    const int nchunk = (N + NP - 1) / NP;         // > 1 only for the 3 -> c position MLP (K <= 8): one thread still owns a row
    for (int i = threadIdx.x; i < nchunk * K * NP; i += blockDim.x) {
        const int c = i / (K * NP), k = (i / NP) % K, n = c * NP + (i % NP);
        float v = 0.f;
        if (n < N) v = transW ? W[(int64_t)k * ldw + n] : W[(int64_t)n * ldw + k];
        ws[i] = v;
    }
    __syncthreads();
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float* xr = X + (int64_t)m * ldx;
    const bool after = (act & AM_ACT_AFTER_RES) != 0;
    float* yr = Y + (int64_t)m * ldy;
    const float* rr = residual ? residual + (int64_t)m * ldr : nullptr;
  for (int c = 0; c < nchunk; ++c) {
    const float* wc = ws + (int64_t)c * K * NP;
    const int nb = c * NP;
    float acc[NP];
#pragma unroll
    for (int n = 0; n < NP; ++n) acc[n] = 0.f;
    auto fma_k = [&](float xv, int k) {
        const float4* w4 = reinterpret_cast<const float4*>(wc + k * NP);
#pragma unroll
        for (int q = 0; q < NP / 4; ++q) {
            const float4 w = w4[q];
            acc[4 * q] = fmaf(xv, w.x, acc[4 * q]); acc[4 * q + 1] = fmaf(xv, w.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(xv, w.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(xv, w.w, acc[4 * q + 3]);
        }
    };
    int k = 0;
    if (vecX) {
        for (; k + 3 < K; k += 4) {
            const float4 x4 = *reinterpret_cast<const float4*>(xr + k);
            fma_k(x4.x, k); fma_k(x4.y, k + 1); fma_k(x4.z, k + 2); fma_k(x4.w, k + 3);
        }
    }
    for (; k < K; ++k) fma_k(xr[k], k);
#pragma unroll
    for (int n = 0; n < NP; ++n) {
        if (nb + n < N) {
            float v = acc[n] + (bias ? bias[nb + n] : 0.f);
            const float r = rr ? rr[nb + n] : 0.f;
            v = after ? apply_act(v + r, act & 15) : apply_act(v, act & 15) + r;
            yr[nb + n] = v;
        }
    }
  }
}

// C[P,Q] += A[R,P]^T B[R,Q].  narrowA != 0: P <= 32 in registers, thread column t of B (Q <= 256); else Q <= 32 in registers,
// thread column t of A.  blockDim = (WT, NY): WT = wide side rounded up to 32, NY row lanes.
template <int NP>
__global__ void __launch_bounds__(256)
rowgemm_dw_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc, int R, int P,
                  int Q, int narrowA, int rows_per_cta) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float red[];  // [NY][NP][WT]
    const int WT = blockDim.x, NY = blockDim.y, t = threadIdx.x, ty = threadIdx.y;
    const int wide = narrowA ? Q : P, narrow = narrowA ? P : Q;
    const float* Wd = narrowA ? B : A;   // wide operand
    const float* Nr = narrowA ? A : B;   // narrow operand
    const int ldw = narrowA ? ldb : lda, ldn = narrowA ? lda : ldb;
    float acc[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) acc[i] = 0.f;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(R, r0 + rows_per_cta);
    if (t < wide) {
        // four rows in flight per thread: the loop is pure load latency otherwise (one coalesced + NP broadcast loads per row)
        int r = r0 + ty;
        for (; r + 3 * NY < r1; r += 4 * NY) {
            float wv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) wv[u] = Wd[(int64_t)(r + u * NY) * ldw + t];
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                if (i < narrow) {
                    float a = acc[i];
#pragma unroll
                    for (int u = 0; u < 4; ++u) a = fmaf(__ldg(Nr + (int64_t)(r + u * NY) * ldn + i), wv[u], a);
                    acc[i] = a;
                }
            }
        }
        for (; r < r1; r += NY) {
            const float wv = Wd[(int64_t)r * ldw + t];
            const float* nr = Nr + (int64_t)r * ldn;
#pragma unroll
            for (int i = 0; i < NP; ++i)
                if (i < narrow) acc[i] = fmaf(__ldg(nr + i), wv, acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) red[(ty * NP + i) * WT + t] = acc[i];
    __syncthreads();
    for (int e = ty * WT + t; e < NP * WT; e += NY * WT) {
        const int i = e / WT, tt = e - i * WT;
        if (i >= narrow || tt >= wide) continue;
        float s = 0.f;
        for (int y = 0; y < NY; ++y) s += red[(y * NP + i) * WT + tt];
        float* dst = narrowA ? C + (int64_t)i * ldc + tt : C + (int64_t)tt * ldc + i;
        atomicAdd(dst, s);
    }
}

}  // namespace

// internal entry points (called from am_linear_f32 / am_gemm_f32 when the shape qualifies); 1 = handled, 0 = not applicable
int am_rowgemm_fwd_(const float* X, int ldx, const float* W, int ldw, int transW, float* Y, int ldy, int M, int N, int K, const float* bias,
                    int act, const float* residual, int ldr, cudaStream_t st) {
    if (M < 8192 || K > 512 || (N > 32 && (K > 8 || N > 512))) return 0;
    const int NP = N <= 4 ? 4 : (N <= 8 ? 8 : (N <= 16 ? 16 : 32));
    const size_t smem = sizeof(float) * (size_t)K * NP * ((N + NP - 1) / NP);
    if (smem > 48 * 1024) return 0;
    const int vecX = ((reinterpret_cast<uintptr_t>(X) & 15u) == 0) && (ldx % 4 == 0);
    dim3 grid(cdiv(M, 256)), block(256);
    switch (NP) {
        case 4: am_launch(rowgemm_fwd_kernel<4>, grid, block, smem, st, 1, X, ldx, W, ldw, transW, Y, ldy, M, N, K, bias, act, residual, ldr, vecX); break;
        case 8: am_launch(rowgemm_fwd_kernel<8>, grid, block, smem, st, 1, X, ldx, W, ldw, transW, Y, ldy, M, N, K, bias, act, residual, ldr, vecX); break;
        case 16: am_launch(rowgemm_fwd_kernel<16>, grid, block, smem, st, 1, X, ldx, W, ldw, transW, Y, ldy, M, N, K, bias, act, residual, ldr, vecX); break;
        default: am_launch(rowgemm_fwd_kernel<32>, grid, block, smem, st, 1, X, ldx, W, ldw, transW, Y, ldy, M, N, K, bias, act, residual, ldr, vecX); break;
    }
    return 1;
}

int am_rowgemm_dw_(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int R, int P, int Q, cudaStream_t st) {
    const int narrow = P < Q ? P : Q, wide = P < Q ? Q : P;
    if (narrow > 32 || wide > 256 || R < 8192) return 0;
    const int narrowA = P <= Q ? 1 : 0;
    const int NP = narrow <= 4 ? 4 : (narrow <= 8 ? 8 : (narrow <= 16 ? 16 : 32));
    const int WT = ((wide + 31) / 32) * 32, NY = 256 / WT > 0 ? 256 / WT : 1;
    const size_t smem = sizeof(float) * (size_t)NY * NP * WT;
    if (smem > 48 * 1024) return 0;
    int ctas = cdiv(R, 256);  // >= 256 rows per CTA; up to 4 CTAs per SM so the row loop's load latency overlaps across CTAs
    if (ctas > 4 * AM_NUM_SMS) ctas = 4 * AM_NUM_SMS;
    const int rows_per_cta = cdiv(R, ctas);
    cudaMemsetAsync(C, 0, sizeof(float) * ((size_t)(P - 1) * ldc + Q), st);
    dim3 grid(cdiv(R, rows_per_cta)), block(WT, NY);
    switch (NP) {
        case 4: am_launch(rowgemm_dw_kernel<4>, grid, block, smem, st, 1, A, lda, B, ldb, C, ldc, R, P, Q, narrowA, rows_per_cta); break;
        case 8: am_launch(rowgemm_dw_kernel<8>, grid, block, smem, st, 1, A, lda, B, ldb, C, ldc, R, P, Q, narrowA, rows_per_cta); break;
        case 16: am_launch(rowgemm_dw_kernel<16>, grid, block, smem, st, 1, A, lda, B, ldb, C, ldc, R, P, Q, narrowA, rows_per_cta); break;
        default: am_launch(rowgemm_dw_kernel<32>, grid, block, smem, st, 1, A, lda, B, ldb, C, ldc, R, P, Q, narrowA, rows_per_cta); break;
    }
    return 1;
}
