// Per-SM ingest rate of a 32 KiB all-CTAs-read-the-same-buffer gather for the load flavours usable for polling
// (development aid).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bw gather_bw.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
template <int F> __device__ __forceinline__ uint4 ld16(const void *p)
{
    uint4 v;
    if (F == 0) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    if (F == 1) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    if (F == 2) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    if (F == 3) asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    if (F == 4) asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    if (F == 5) asm volatile("ld.relaxed.cta.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void tma_bulk(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// F < 6: every thread loads 4 x 16 B (the CTA reads `bytes` = 32 KiB); F == 6: one cp.async.bulk of 32 KiB per CTA
template <int F> __global__ void __launch_bounds__(512, 1) gather(const unsigned char *buf, int iters, int same, long long *out, unsigned *sink)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ unsigned long long bar;
    const unsigned char *src = buf + (same ? 0 : (size_t)blockIdx.x * 32768);
    unsigned acc = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (F < 6) {
            uint4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = ld16<F>(src + 16 * (threadIdx.x + 512 * j));
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4 *>(sm + 16 * (threadIdx.x + 512 * j)) = v[j];
        } else {
            if (threadIdx.x == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(32768u) : "memory");
                tma_bulk(sm, src, 32768u, &bar);
            }
            unsigned ok = 0;
            while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(&bar)), "r"((unsigned)(it & 1)) : "memory");
        }
        __syncthreads();
        acc += sm[threadIdx.x * 4];
        __syncthreads();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters;
    if (acc == 0x12345678u) sink[0] = acc;
}
int main()
{
    unsigned char *buf; long long *out; unsigned *sink;
    CK(cudaMalloc(&buf, 128 * 32768)); CK(cudaMemset(buf, 1, 128 * 32768)); CK(cudaMalloc(&out, 128 * 8)); CK(cudaMalloc(&sink, 4));
    const char *names[] = {"ld.volatile.v4", "ld.relaxed.gpu.v4", "ld.global.cg.v4", "ld.global.cv.v4", "ld.L1::no_allocate.v4", "ld.relaxed.cta.v4", "cp.async.bulk 32 KiB"};
    long long h[128];
#define RUN(F) { CK(cudaFuncSetAttribute(gather<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536)); \
    for (int same = 1; same >= 0; --same) for (int ncta : {128, 16, 1}) { \
        gather<F><<<ncta, 512, 65536>>>(buf, 500, same, out, sink); CK(cudaDeviceSynchronize()); \
        gather<F><<<ncta, 512, 65536>>>(buf, 500, same, out, sink); CK(cudaDeviceSynchronize()); \
        CK(cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost)); long long mx = 0; for (int i = 0; i < ncta; ++i) mx = h[i] > mx ? h[i] : mx; \
        printf("%-24s %3d CTAs, %s buffer: %6lld cycles per 32 KiB gather (%.1f B/clk/SM)\n", names[F], ncta, same ? "same    " : "distinct", mx, 32768.0 / mx); } }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
    return 0;
}
