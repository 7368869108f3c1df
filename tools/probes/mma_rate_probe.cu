// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (bf16, M128, K16) for the operand sources / N sizes the attention kernel
// uses — A from shared memory (SS) vs A from TMEM (TS), N = 64 / 128 / 256 — optionally with 8 warps hammering tcgen05.ld at the
// same time (the softmax warps).  One CTA per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate_probe mma_rate_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_k(uint32_t a) {
    return (uint64_t)((a & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t a, uint32_t lbo) {
    return (uint64_t)((a & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t idesc(int N, bool bmn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (bmn ? (1u << 16) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(id) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t ph) {
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(bar)), "r"(ph) : "memory");
        if (++spins > 100000000u) __trap();
    }
}

// mode: 0 SS N128 (K-major B) | 1 SS N64 | 2 TS N64 (MN-major B) | 3 TS N128 | 4 TS N128 + TS N64 alternating (current PV) | 5 TS N256
//       6 SS N256 | 7 TS N64 x3 (old PV)
template <int MODE>
__global__ void __launch_bounds__(384, 1) probe(int iters, int ldwarps, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        stop = 0;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&holder)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = holder;
    if (warp == 11 && lane == 0) {
        const uint32_t sb = s32(smem);
        const uint64_t a = desc_k(sb), bk = desc_k(sb + 32768), bmn = desc_mn(sb + 32768, 16384);
        long long t0 = clock64();
        // descriptors of the 4 k-steps precomputed: the loop body is nothing but MMAs (a single thread issues ~1 instruction / 5 cycles)
        uint64_t ak[4], bkk[4], bm[4], bm2[4];
        uint32_t ta[4], tl[4];
        for (int k = 0; k < 4; ++k) { ak[k] = a + k * 2; bkk[k] = bk + k * 2; bm[k] = bmn + k * 128; bm2[k] = bmn + 1024 + k * 128; ta[k] = tb + k * 8; tl[k] = tb + 32 + k * 8; }
        constexpr uint32_t I64 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t I128 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t I256 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t MN = 1u << 16;
        for (int it = 0; it < iters; it += 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (MODE == 0) mma_ss(tb + 256, ak[k], bkk[k], I128);
                if (MODE == 1) mma_ss(tb + 256, ak[k], bkk[k], I64);
                if (MODE == 2) mma_ts(tb + 256, ta[k], bm[k], I64 | MN);
                if (MODE == 3) mma_ts(tb + 256, ta[k], bm[k], I128 | MN);
                if (MODE == 4) { mma_ts(tb + 256, ta[k], bm[k], I128 | MN); mma_ts(tb + 256, tl[k], bm[k], I64 | MN); }
                if (MODE == 5) mma_ts(tb + 256, ta[k], bm[k], I256 | MN);
                if (MODE == 6) mma_ss(tb + 256, ak[k], bkk[k], I256);
                if (MODE == 7) { mma_ts(tb + 256, ta[k], bm[k], I64 | MN); mma_ts(tb + 256, ta[k], bm2[k], I64 | MN); mma_ts(tb + 256, tl[k], bm[k], I64 | MN); }
                if (MODE == 8) { mma_ts(tb + 256, ta[k], bm[k], I128 | MN); mma_ts(tb + 384, tl[k], bm[k], I64 | MN); }   // two accumulators
                if (MODE == 9) { mma_ss(tb + 256, ak[k], bkk[k], I128); mma_ss(tb + 384, ak[k], bkk[k], I128); }          // two accumulators
            }
        }
        long long t1 = clock64();
        commit(&bar);
        wait(&bar, 0);
        long long t2 = clock64();
        stop = 1;
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    } else if (warp < ldwarps) {
        // background tcgen05.ld traffic from the "softmax" warps: 32x32b.x32 loads of the first 256 columns of the own lane quadrant
        const uint32_t la = tb + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t acc = 0, n = 0;
        while (!stop) {
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(la + (n & 7) * 32) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 32; ++c) acc ^= r[c];
            ++n;
        }
        if (acc == 0x12345678u) out[8] = acc;
        if (blockIdx.x == 0 && lane == 0) out[16 + warp] = n;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

template <int M> void go(int smem, int iters, int ldw, long long* d) {
    cudaFuncSetAttribute(probe<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<M><<<148, 384, smem>>>(iters, ldw, d);
}
void launch(int mode, int smem, int iters, int ldw, long long* d) {
    switch (mode) {
        case 0: go<0>(smem, iters, ldw, d); break; case 1: go<1>(smem, iters, ldw, d); break; case 2: go<2>(smem, iters, ldw, d); break;
        case 3: go<3>(smem, iters, ldw, d); break; case 4: go<4>(smem, iters, ldw, d); break; case 5: go<5>(smem, iters, ldw, d); break;
        case 6: go<6>(smem, iters, ldw, d); break; case 7: go<7>(smem, iters, ldw, d); break; case 8: go<8>(smem, iters, ldw, d); break;
        case 9: go<9>(smem, iters, ldw, d); break;
    }
}
int main() {
    long long* d;
    cudaMalloc(&d, 64 * sizeof(long long));
    const int smem = 130 * 1024;

    const char* names[] = {"SS N128", "SS N64", "TS N64", "TS N128", "TS N128 + TS N64 (2 MMAs)", "TS N256", "SS N256", "TS N64 x3 (3 MMAs)", "TS N128 + TS N64, 2 accumulators", "SS N128 x2, 2 accumulators"};
    const int iters = 2048;
    for (int ldw = 0; ldw <= 8; ldw += 8)
        for (int mode = 0; mode < 10; ++mode) {
            cudaMemset(d, 0, 64 * sizeof(long long));
            launch(mode, smem, iters, ldw, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
            long long h[64];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            printf("ld warps %d  %-28s issue %.1f cyc/iter   complete %.1f cyc/iter   (background tcgen05.ld x32 per warp: %lld)\n", ldw, names[mode],
                   (double)h[0] / iters, (double)h[1] / iters, h[16]);
        }
    return 0;
}
