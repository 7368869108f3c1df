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
    pdl_launch_dependents();
    pdl_wait();
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


// Vectorised variant for D % 128 == 0 and 16-byte aligned rows: every lane owns NV float4 chunks (chunk c = lane + 32 i),
// 128-bit loads / stores, bf16 (hi | lo) written as 8-byte pairs.
template <int NV>
__global__ void __launch_bounds__(256) layernorm_vec_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ R, int ldr,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ Y, int ldy, int M, int D, float eps,
                                                            __nv_bfloat16* __restrict__ Y2, int Np2,
                                                            __nv_bfloat16* __restrict__ Y2w, int seg, int seg_q0,
                                                            const __nv_bfloat16* __restrict__ R2) {
    pdl_launch_dependents();
    pdl_wait();
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    int lane = threadIdx.x & 31;
    // optional second, COMPACT copy of the rows [seg_q0, seg) of every `seg`-row segment (am_layernorm_win)
    __nv_bfloat16* w2 = nullptr;
    if (Y2w) {
        const int sg = row / seg, i = row - sg * seg;
        if (i >= seg_q0) w2 = Y2w + ((int64_t)sg * (seg - seg_q0) + (i - seg_q0)) * 2 * Np2;
    }
    const float4* x4 = reinterpret_cast<const float4*>(X + (int64_t)row * ldx);
    const float4* r4 = R ? reinterpret_cast<const float4*>(R + (int64_t)row * ldr) : nullptr;
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float4 t = x4[lane + 32 * i];
        if (r4) { float4 u = r4[lane + 32 * i]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
        if (R2) {  // residual as a bf16 (hi | lo) pair tensor [M, 2*Np2]: t += hi + lo, the sum the GEMM epilogue used to form
            const uint2* rh = reinterpret_cast<const uint2*>(R2 + (int64_t)row * 2 * Np2) + lane + 32 * i;
            const uint2 h = rh[0], l = rh[Np2 / 4];
            t.x += __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
            t.y += __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u);
            t.z += __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
            t.w += __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u);
        }
        v[i] = t; s += (t.x + t.y) + (t.z + t.w);
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        const float4 g = g4[c], b = b4[c];
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + b.x; o.y = (v[i].y - mean) * rstd * g.y + b.y;
        o.z = (v[i].z - mean) * rstd * g.z + b.z; o.w = (v[i].w - mean) * rstd * g.w + b.w;
        if (Y) reinterpret_cast<float4*>(Y + (int64_t)row * ldy)[c] = o;
        if (Y2) {
            __nv_bfloat162 h01 = __floats2bfloat162_rn(o.x, o.y), h23 = __floats2bfloat162_rn(o.z, o.w);
            uint32_t u01 = *reinterpret_cast<uint32_t*>(&h01), u23 = *reinterpret_cast<uint32_t*>(&h23);
            __nv_bfloat162 l01 = __floats2bfloat162_rn(o.x - __uint_as_float(u01 << 16), o.y - __uint_as_float(u01 & 0xffff0000u));
            __nv_bfloat162 l23 = __floats2bfloat162_rn(o.z - __uint_as_float(u23 << 16), o.w - __uint_as_float(u23 & 0xffff0000u));
            uint2* hi = reinterpret_cast<uint2*>(Y2 + (int64_t)row * 2 * Np2) + c;
            hi[0] = make_uint2(u01, u23);
            hi[Np2 / 4] = make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23));
            if (w2) {
                uint2* wh = reinterpret_cast<uint2*>(w2) + c;
                wh[0] = make_uint2(u01, u23);
                wh[Np2 / 4] = make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23));
            }
        }
    }
}

// LayerNorm that OVERLAPS its producer GEMM (am_linear_tc_set_rowflags): LNF_SUB CTAs per 128-row block, each 16 rows (8 warps x 2
// rows in flight).  Thread 0 waits until the block's completion counter (low 20 bits) has reached `expect`; when a CTA is done it adds
// 1 << 20, and the last of the LNF_SUB consumers resets the counter for the next GEMM.  No griddepcontrol.wait: the kernel is launched as
// a programmatic dependent of the GEMM and its CTAs become resident next to the (persistent, one-per-SM) GEMM CTAs once those have all
// started — row blocks finished in the GEMM's first round are normalised while its tail round is still running.  X is read with
// ld.global.cg (L2): the lines were written by another SM's TMA store.  (A first version with ONE CTA per row block, 16 rows per warp
// in sequence, had too few rows in flight: 878 vs 979 denoise-steps/s.)
constexpr int LNF_SUB = 8;
template <int NV>
__global__ void __launch_bounds__(256) layernorm_flag_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, int M, int D, float eps,
                                                             __nv_bfloat16* __restrict__ Y2, int Np2, __nv_bfloat16* __restrict__ Y2w, int seg,
                                                             int seg_q0, int* __restrict__ flags, int expect) {
    pdl_launch_dependents();
    const int mt = blockIdx.x / LNF_SUB, sub = blockIdx.x % LNF_SUB;
    if (threadIdx.x == 0) {
        uint32_t spins = 0;
        while (true) {
            int v;
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + mt) : "memory");
            if ((v & 0xFFFFF) >= expect) break;
            __nanosleep(64);
            if (++spins > 40000000u) __trap();   // the producer never signalled: fail the launch instead of hanging the box
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
    const int row0 = mt * 128 + sub * 16 + warp * 2;
    float4 v[2][NV];
    float s[2] = {0.f, 0.f};
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        if (row0 + rr < M) {
            const float4* x4 = reinterpret_cast<const float4*>(X + (int64_t)(row0 + rr) * ldx);
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                v[rr][i] = __ldcg(x4 + lane + 32 * i);
                s[rr] += (v[rr][i].x + v[rr][i].y) + (v[rr][i].z + v[rr][i].w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < NV; ++i) v[rr][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        const int row = row0 + rr;
        const float mean = warp_sum(s[rr]) / (float)D;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float a = v[rr][i].x - mean, b = v[rr][i].y - mean, c = v[rr][i].z - mean, d = v[rr][i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
        if (row >= M) continue;
        __nv_bfloat16* w2 = nullptr;
        if (Y2w) {
            const int sg = row / seg, i = row - sg * seg;
            if (i >= seg_q0) w2 = Y2w + ((int64_t)sg * (seg - seg_q0) + (i - seg_q0)) * 2 * Np2;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            const float4 g = g4[c], b = b4[c];
            float4 o;
            o.x = (v[rr][i].x - mean) * rstd * g.x + b.x; o.y = (v[rr][i].y - mean) * rstd * g.y + b.y;
            o.z = (v[rr][i].z - mean) * rstd * g.z + b.z; o.w = (v[rr][i].w - mean) * rstd * g.w + b.w;
            __nv_bfloat162 h01 = __floats2bfloat162_rn(o.x, o.y), h23 = __floats2bfloat162_rn(o.z, o.w);
            uint32_t u01 = *reinterpret_cast<uint32_t*>(&h01), u23 = *reinterpret_cast<uint32_t*>(&h23);
            __nv_bfloat162 l01 = __floats2bfloat162_rn(o.x - __uint_as_float(u01 << 16), o.y - __uint_as_float(u01 & 0xffff0000u));
            __nv_bfloat162 l23 = __floats2bfloat162_rn(o.z - __uint_as_float(u23 << 16), o.w - __uint_as_float(u23 & 0xffff0000u));
            const uint2 hv = make_uint2(u01, u23), lv = make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23));
            uint2* hi = reinterpret_cast<uint2*>(Y2 + (int64_t)row * 2 * Np2) + c;
            hi[0] = hv;
            hi[Np2 / 4] = lv;
            if (w2) {
                uint2* wh = reinterpret_cast<uint2*>(w2) + c;
                wh[0] = hv;
                wh[Np2 / 4] = lv;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int old = atomicAdd(flags + mt, 1 << 20);
        if ((old >> 20) == LNF_SUB - 1) atomicExch(flags + mt, 0);   // last consumer of the block: ready for the next GEMM
    }
}

}  // namespace

extern "C" int am_layernorm_win(const float* X, int ldx, const float* R, int ldr, const float* gamma, const float* beta, float* Y, int ldy,
                                int M, int D, float eps, void* Y2, int Np2, void* Y2w, int seg, int seg_q0, const void* Rsplit,
                                am_stream_t stream) {
    AM_REQUIRE(!Rsplit || (Np2 >= D && Np2 % 32 == 0 && (reinterpret_cast<uintptr_t>(Rsplit) & 15u) == 0), AM_EINVAL,
               "am_layernorm_win: a split residual needs Np2 (its row layout) and 16-byte alignment");
    const __nv_bfloat16* r2 = reinterpret_cast<const __nv_bfloat16*>(Rsplit);
    AM_REQUIRE(!Y2w || (Y2 && seg > 0 && seg_q0 >= 0 && seg_q0 < seg && M % seg == 0 && (reinterpret_cast<uintptr_t>(Y2w) & 15u) == 0), AM_EINVAL,
               "am_layernorm_win: the window copy needs Y2, 0 <= seg_q0 < seg, M % seg == 0 and a 16-byte aligned Y2w");
    __nv_bfloat16* y2w = reinterpret_cast<__nv_bfloat16*>(Y2w);
    AM_REQUIRE(X && gamma && beta && (Y || Y2) && M > 0 && D > 0 && D <= 1024, AM_EINVAL, "am_layernorm: bad args (D <= 1024)");
    AM_REQUIRE(ldx >= D && (!Y || ldy >= D) && (!R || ldr >= D), AM_EINVAL, "am_layernorm: bad strides");
    AM_REQUIRE(!Y2 || (Np2 >= D && Np2 % 32 == 0 && Np2 <= 1024), AM_EINVAL, "am_layernorm: bad Np2");
    __nv_bfloat16* y2 = reinterpret_cast<__nv_bfloat16*>(Y2);
    int grid = cdiv(M, 8);
    auto a16 = [](const void* p_) { return (reinterpret_cast<uintptr_t>(p_) & 15u) == 0; };
    const bool vec = (D % 128 == 0) && D <= 1024 && a16(X) && (ldx % 4 == 0) && (!R || (a16(R) && ldr % 4 == 0)) && a16(gamma) && a16(beta) &&
                     (!Y || (a16(Y) && ldy % 4 == 0)) && (!Y2 || (a16(Y2) && Np2 == D)) && (!Rsplit || Np2 == D);
    if (vec) {
        switch (D / 128) {
            case 1: am_launch(layernorm_vec_kernel<1>, dim3(grid), dim3(256), 0, as_stream(stream), 1, X, ldx, R, ldr, gamma, beta, Y, ldy, M, D, eps, y2, Np2, y2w, seg, seg_q0, r2); break;
            case 2: am_launch(layernorm_vec_kernel<2>, dim3(grid), dim3(256), 0, as_stream(stream), 1, X, ldx, R, ldr, gamma, beta, Y, ldy, M, D, eps, y2, Np2, y2w, seg, seg_q0, r2); break;
            case 4: am_launch(layernorm_vec_kernel<4>, dim3(grid), dim3(256), 0, as_stream(stream), 1, X, ldx, R, ldr, gamma, beta, Y, ldy, M, D, eps, y2, Np2, y2w, seg, seg_q0, r2); break;
            case 8: am_launch(layernorm_vec_kernel<8>, dim3(grid), dim3(256), 0, as_stream(stream), 1, X, ldx, R, ldr, gamma, beta, Y, ldy, M, D, eps, y2, Np2, y2w, seg, seg_q0, r2); break;
            default: goto scalar_path;
        }
        AM_LAUNCH_CHECK("layernorm");
        return AM_OK;
    }
scalar_path:
    AM_REQUIRE(!Y2w && !Rsplit, AM_EINVAL, "am_layernorm_win: window copy / split residual are implemented for the vectorised path only (D % 128 == 0, aligned)");
    if (D <= 256) am_launch(layernorm_kernel<8>, dim3(grid), dim3(256), 0, as_stream(stream), 1, X, ldx, R, ldr, gamma, beta, Y, ldy, M, D, eps, y2, Np2);
    else if (D <= 512) am_launch(layernorm_kernel<16>, dim3(grid), dim3(256), 0, as_stream(stream), 1, X, ldx, R, ldr, gamma, beta, Y, ldy, M, D, eps, y2, Np2);
    else am_launch(layernorm_kernel<32>, dim3(grid), dim3(256), 0, as_stream(stream), 1, X, ldx, R, ldr, gamma, beta, Y, ldy, M, D, eps, y2, Np2);
    AM_LAUNCH_CHECK("layernorm");
    return AM_OK;
}

extern "C" int am_layernorm(const float* X, int ldx, const float* R, int ldr, const float* gamma, const float* beta, float* Y, int ldy,
                            int M, int D, float eps, void* Y2, int Np2, am_stream_t stream) {
    return am_layernorm_win(X, ldx, R, ldr, gamma, beta, Y, ldy, M, D, eps, Y2, Np2, nullptr, 0, 0, nullptr, stream);
}

// LayerNorm overlapped with its producer GEMM (see layernorm_flag_kernel): X fp32 [M, D] (written by an am_linear_tc launch armed with
// am_linear_tc_set_rowflags(flags)), bf16 (hi|lo) output only, D % 128 == 0; `expect` = 4 * N of that GEMM.  Same arithmetic as
// am_layernorm / am_layernorm_win (bit-identical outputs).
extern "C" int am_layernorm_flags(const float* X, int ldx, const float* gamma, const float* beta, int M, int D, float eps, void* Y2, int Np2,
                                  void* Y2w, int seg, int seg_q0, int* flags, int expect, am_stream_t stream) {
    AM_REQUIRE(X && gamma && beta && Y2 && flags && expect > 0 && M > 0, AM_EINVAL, "am_layernorm_flags: bad args");
    auto a16 = [](const void* p_) { return (reinterpret_cast<uintptr_t>(p_) & 15u) == 0; };
    AM_REQUIRE((D % 128 == 0) && D <= 1024 && Np2 == D && ldx >= D && (ldx % 4 == 0) && a16(X) && a16(gamma) && a16(beta) && a16(Y2), AM_EINVAL,
               "am_layernorm_flags: D % 128 == 0, Np2 == D and 16-byte alignment required");
    AM_REQUIRE(!Y2w || (seg > 0 && seg_q0 >= 0 && seg_q0 < seg && M % seg == 0 && a16(Y2w)), AM_EINVAL, "am_layernorm_flags: bad window");
    __nv_bfloat16* y2 = reinterpret_cast<__nv_bfloat16*>(Y2);
    __nv_bfloat16* y2w = reinterpret_cast<__nv_bfloat16*>(Y2w);
    const int grid = cdiv(M, 128) * LNF_SUB;
    switch (D / 128) {
        case 1: am_launch(layernorm_flag_kernel<1>, dim3(grid), dim3(256), 0, as_stream(stream), 1, X, ldx, gamma, beta, M, D, eps, y2, Np2, y2w, seg, seg_q0, flags, expect); break;
        case 2: am_launch(layernorm_flag_kernel<2>, dim3(grid), dim3(256), 0, as_stream(stream), 1, X, ldx, gamma, beta, M, D, eps, y2, Np2, y2w, seg, seg_q0, flags, expect); break;
        case 4: am_launch(layernorm_flag_kernel<4>, dim3(grid), dim3(256), 0, as_stream(stream), 1, X, ldx, gamma, beta, M, D, eps, y2, Np2, y2w, seg, seg_q0, flags, expect); break;
        case 8: am_launch(layernorm_flag_kernel<8>, dim3(grid), dim3(256), 0, as_stream(stream), 1, X, ldx, gamma, beta, M, D, eps, y2, Np2, y2w, seg, seg_q0, flags, expect); break;
        default: am_set_error_("am_layernorm_flags: D / 128 must be 1, 2, 4 or 8"); return AM_EINVAL;
    }
    AM_LAUNCH_CHECK("layernorm_flags");
    return AM_OK;
}
