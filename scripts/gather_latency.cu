// Round-2 microbenchmark (development aid): how long does the wide kernel's warp-local quad gather take when the data is
// ALREADY in L2 (no producer running)?  128 CTAs x 512 threads, thread pattern of wavernn_wide.cuh::gather_rows
// (warp w reads the nq quads of units 32w..32w+31), for load flavours {volatile (STRONG.SYS), relaxed.gpu, ld.cg} and with
// the warp -> slice assignment rotated per CTA (so the 128 SMs do not walk the same L2 lines in the same order).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_latency gather_latency.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
constexpr int NT = 512, NQ = 7, UROW = 28;

template <int FLAVOUR>
__device__ __forceinline__ uint4 ldq(const unsigned *p)
{
    uint4 v;
    if (FLAVOUR == 0) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    else if (FLAVOUR == 1) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    else asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// mode bit 0: rotate the slice by CTA; bit 1: one warp only (latency of a lone warp)
template <int FLAVOUR>
__global__ void __launch_bounds__(NT, 1) gather(int nq, int mode, int iters, const unsigned *buf, long long *cyc, float *sink)
{
    extern __shared__ float sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cta = blockIdx.x;
    const int slice = (mode & 1) ? (warp + cta) & 15 : warp;
    const unsigned rcp = (65536u + nq - 1) / nq;
    float acc = 0.f;
    long long total = 0;
    for (int it = 0; it < iters; ++it) {
        const unsigned *vec = buf + (size_t)(it & 3) * 512 * UROW;
        __syncthreads();
        const long long t0 = clock64();
        if (!(mode & 2) || warp == 0) {
            uint4 v[NQ];
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                v[j] = make_uint4(0, 0, 0, 0);
                if (j < nq) {
                    const int g = j * 32 + lane, unit = (int)(((unsigned)g * rcp) >> 16);
                    v[j] = ldq<FLAVOUR>(buf ? vec + (32 * slice + unit) * UROW + (g - unit * nq) * 4 : nullptr);
                }
            }
#pragma unroll
            for (int j = 0; j < NQ; ++j)
                if (j < nq) {
                    const int g = j * 32 + lane, unit = (int)(((unsigned)g * rcp) >> 16);
                    *reinterpret_cast<uint4 *>(sm + (32 * slice + unit) * UROW + (g - unit * nq) * 4) = v[j];
                }
            __syncwarp();
        }
        if (tid == 0) total += clock64() - t0;
        __syncthreads();
        acc += sm[(tid * 29) % (512 * UROW)];
    }
    if (tid == 0) cyc[cta] = total;
    if (acc == 123.f) *sink = acc;
}

template <int FLAVOUR>
static void run(const char *name, int nq, int mode, int ncta, int iters, unsigned *buf, long long *cyc, float *sink)
{
    const size_t smem = 200 * 1024;
    CK(cudaFuncSetAttribute(gather<FLAVOUR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gather<FLAVOUR><<<ncta, NT, smem>>>(nq, mode, iters, buf, cyc, sink);
    CK(cudaDeviceSynchronize());
    long long h[256];
    CK(cudaMemcpy(h, cyc, ncta * sizeof(long long), cudaMemcpyDeviceToHost));
    double mean = 0, mx = 0;
    for (int i = 0; i < ncta; ++i) { mean += (double)h[i] / iters; if ((double)h[i] / iters > mx) mx = (double)h[i] / iters; }
    printf("ctas %3d  %-12s nq %d  %-8s %-9s  warp-0 gather: mean %6.0f clk  max %6.0f clk\n", ncta, name, nq, (mode & 1) ? "rotated" : "same", (mode & 2) ? "one warp" : "16 warps", mean / ncta, mx);
}

int main()
{
    unsigned *buf;
    long long *cyc;
    float *sink;
    CK(cudaMalloc(&buf, (size_t)4 * 512 * UROW * 4 + 64));
    CK(cudaMemset(buf, 1, (size_t)4 * 512 * UROW * 4 + 64));
    CK(cudaMalloc(&cyc, 256 * sizeof(long long)));
    CK(cudaMalloc(&sink, 4));
    const int iters = 2000;
    for (int ncta : {1, 16, 128})
        for (int nq : {1, 7})
            for (int mode : {0, 1, 2}) {
                run<0>("volatile", nq, mode, ncta, iters, buf, cyc, sink);
                run<1>("relaxed.gpu", nq, mode, ncta, iters, buf, cyc, sink);
                run<2>("ld.cg", nq, mode, ncta, iters, buf, cyc, sink);
            }
    return 0;
}
