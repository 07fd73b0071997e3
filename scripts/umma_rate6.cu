// umma_rate.cu -- tcgen05.mma issue/execute rate for small-N tiles: no-swizzle vs 128B-swizzle K-major operands (timing only,
// operands are whatever is in shared memory).  Development microbenchmark.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// mode 0: no swizzle, LBO = rows*16, SBO = 128, k-step advance 2*rows*16
// mode 1: 128B swizzle, rows of 128 B (64 k), SBO = 1024, k-step advance 32 B, 4 k-steps per 64-k block then + rows*128
__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int M, int N, int rowsA, int nmma, long long *out, const uint8_t *gsrc, int stream, int commit_every)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, sbar[2], cbar;
    __shared__ uint32_t tmem_s;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 4; i += 128) ((uint32_t *)smem)[i] = 0x3c003c00u;
    if (tid == 0) { stop = 0; asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&sbar[0]))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&sbar[1]))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&cbar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    if (warp == 0) {
        uint32_t pred;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t a_base = s32(smem), b_base = s32(smem) + 128 * 1024;
        uint32_t hi, a_lo0, b_lo0, a_inc, b_inc;
        if (mode == 0) {
            hi = (128u >> 4) | (1u << 14);
            a_lo0 = ((a_base & 0x3FFFF) >> 4) | ((uint32_t)(rowsA * 16 >> 4) << 16);
            b_lo0 = ((b_base & 0x3FFFF) >> 4) | ((uint32_t)(N * 16 >> 4) << 16);
            a_inc = rowsA * 2; b_inc = N * 2;
        } else {
            hi = (1024u >> 4) | (1u << 14) | (2u << 29);           // SBO 1024, version 1, layout SWIZZLE_128B (bits 61-63 = 2)
            a_lo0 = ((a_base & 0x3FFFF) >> 4) | (1u << 16);
            b_lo0 = ((b_base & 0x3FFFF) >> 4) | (1u << 16);
            a_inc = 2; b_inc = 2;                                   // 32 B per k-step inside the 128 B row
        }
        long long t0 = clock64();
        uint32_t a_lo = a_lo0, b_lo = b_lo0;
        for (int i = 0; i < nmma; ++i) {
            if (pred)
                asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(i > 0 ? 1u : 0u));
            if (commit_every && (i & (commit_every - 1)) == commit_every - 1) {
                if (pred) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&cbar)) : "memory");
                if (stream & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (stream & 4) { while (!try_wait(s32(&sbar[1]), 1)) {} }
                if (stream & 8) __syncwarp();
            }
            if ((i & 3) == 3) {
                if (mode == 0) { a_lo = a_lo0 + ((i >> 2) & 3) * 4 * a_inc; b_lo = b_lo0; }
                else { a_lo = a_lo0 + (((i >> 2) & 3) * rowsA * 128 >> 4); b_lo = b_lo0; }
            } else { a_lo += a_inc; b_lo += b_inc; }
        }
        if (pred) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
        __syncwarp();
        long long t1 = clock64();
        while (!try_wait(s32(&bar), 0)) {}
        long long t2 = clock64();
        if (pred) { out[0] = t1 - t0; out[1] = t2 - t0; }
        stop = 1;
    }
    if ((warp == 1 || warp == 2) && (tid & 31) == 0 && stream) {
        const int pw = warp - 1;
        const uint32_t dst = s32(smem) + 136 * 1024 + pw * 32768, sb = s32(&sbar[pw]);
        long long bytes = 0;
        for (unsigned it = 0; !stop; ++it) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(32768) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(gsrc + (size_t)((it * 2 + pw) % 200) * 32768), "r"(32768), "r"(sb) : "memory");
            while (!try_wait(sb, it & 1)) {}
            bytes += 32768;
        }
        out[2 + pw] = bytes;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
int main()
{
    long long *d, h[4];
    uint8_t *g;
    CK(cudaMalloc(&d, 32));
    CK(cudaMalloc(&g, 200 * 32768));
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int nmma = 4000;
    const char *names[] = {"commit only", "commit + tcgen05.fence::after_thread_sync", "commit + try_wait on a completed mbarrier", "commit + fence + try_wait", "commit + __syncwarp", "all"};
    const int flags[] = {0, 2, 4, 6, 8, 14};
    for (int v = 0; v < 6; ++v)
        for (int ce : {4, 8, 16}) {
            CK(cudaMemset(d, 0, 32));
            rate_kernel<<<1, 128, 200 * 1024>>>(0, 128, 32, 128, nmma, d, g, flags[v], ce);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost));
            printf("every %2d MMAs: %-45s %.1f clk/mma\n", ce, names[v], (double)h[1] / nmma);
        }
    return 0;
}
