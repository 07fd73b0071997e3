// umma_probe.cu -- bring-up probes for the dense (tcgen05) WaveRNN step kernel.  Development tool, not product.
//   A: tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM), no-swizzle K-major operand images, checked against the host
//   B: cluster-of-8 all-to-all over distributed shared memory with cp.async.bulk shared::cta -> shared::cluster
//   C: L2 -> shared-memory streaming rate per SM with cp.async.bulk (unicast), whole chip and one cluster
//   D: the same stream with .multicast::cluster (each CTA fetches 1/8 of a chunk for all 8)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, unsigned parity)
{
    for (long i = 0; i < 20000000; ++i) if (mbar_try_wait(bar, parity)) return true;
    return false;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- A: one 128 x N x K GEMM on tcgen05
__host__ __device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;     // descriptor version (Blackwell)
    return d;                   // layout type 0 = no swizzle
}
__host__ __device__ inline uint32_t make_idesc(int M, int N)
{
    uint32_t d = 0;
    d |= 1u << 4;               // D format F32
    d |= 1u << 7;               // A format BF16
    d |= 1u << 10;              // B format BF16
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;                   // A, B K-major
}

template <int N, int K>
__global__ void __launch_bounds__(128, 1) mma_probe(const __nv_bfloat16 *a_img, const __nv_bfloat16 *b_img, float *d_out, int swap_offsets, int *status)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __nv_bfloat16 *sa = (__nv_bfloat16 *)smem;                       // 128 x K
    __nv_bfloat16 *sb = (__nv_bfloat16 *)(smem + 128 * K * 2);       // N x K
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * K / 8; i += 128) ((uint4 *)sa)[i] = ((const uint4 *)a_img)[i];
    for (int i = tid; i < N * K / 8; i += 128) ((uint4 *)sb)[i] = ((const uint4 *)b_img)[i];
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, N);
        for (int j = 0; j < K / 16; ++j) {
            uint32_t a_lbo = 128 * 16, a_sbo = 128, b_lbo = N * 16, b_sbo = 128;
            if (swap_offsets) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
            const uint64_t da = make_desc(smem_u32(sa) + j * 2 * 128 * 16, a_lbo, a_sbo);
            const uint64_t db = make_desc(smem_u32(sb) + j * 2 * N * 16, b_lbo, b_sbo);
            const uint32_t acc = j > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    if (!mbar_wait(&bar, 0)) { if (tid == 0) *status = 1; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                       "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) d_out[(size_t)tid * N + c0 + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

static void pack_kmajor(const std::vector<float> &src, int rows, int K, std::vector<__nv_bfloat16> &img)   // [K/8][rows/8][8][8]
{
    img.resize((size_t)rows * K);
    for (int m = 0; m < rows; ++m)
        for (int k = 0; k < K; ++k) img[((size_t)(k / 8) * (rows / 8) + m / 8) * 64 + (m % 8) * 8 + k % 8] = __float2bfloat16(src[(size_t)m * K + k]);
}

template <int N, int K>
static void run_mma_probe()
{
    std::vector<float> A(128 * K), B(N * K);
    for (auto &v : A) v = (float)((rand() % 17) - 8) / 8.0f;
    for (auto &v : B) v = (float)((rand() % 13) - 6) / 4.0f;
    std::vector<__nv_bfloat16> ai, bi;
    pack_kmajor(A, 128, K, ai);
    pack_kmajor(B, N, K, bi);
    __nv_bfloat16 *da, *db; float *dd; int *ds;
    CK(cudaMalloc(&da, ai.size() * 2)); CK(cudaMalloc(&db, bi.size() * 2)); CK(cudaMalloc(&dd, 128 * N * 4)); CK(cudaMalloc(&ds, 4));
    CK(cudaMemcpy(da, ai.data(), ai.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, bi.data(), bi.size() * 2, cudaMemcpyHostToDevice));
    const int smem = (128 + N) * K * 2;
    CK(cudaFuncSetAttribute(mma_probe<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int swap = 0; swap < 1; ++swap) {
        CK(cudaMemset(dd, 0, 128 * N * 4)); CK(cudaMemset(ds, 0, 4));
        mma_probe<N, K><<<1, 128, smem>>>(da, db, dd, swap, ds);
        CK(cudaDeviceSynchronize());
        std::vector<float> D(128 * N); int st;
        CK(cudaMemcpy(D.data(), dd, 128 * N * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int k = 0; k < K; ++k) ref += (double)__bfloat162float(__float2bfloat16(A[m * K + k])) * (double)__bfloat162float(__float2bfloat16(B[n * K + k]));
                maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
            }
        printf("A: mma 128x%dx%d  %s  status %d  max|err| %.3g  %s\n", N, K, swap ? "LBO=row-group SBO=k-chunk" : "LBO=k-chunk SBO=row-group", st, maxerr, maxerr < 1e-3 ? "MATCH" : "mismatch");
    }
    cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(ds);
}

// ---------------------------------------------------------------- B: cluster all-to-all over DSMEM bulk copies
template <int CL>
__global__ void __launch_bounds__(128, 1) dsmem_probe(int rounds, int slice_bytes, long long *cycles, int *status)
{
    extern __shared__ __align__(1024) uint8_t smem[];     // [2 buffers][CL slices][slice_bytes]
    __shared__ uint64_t bar[2];
    const int tid = threadIdx.x;
    const uint32_t rank = cluster_rank();
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = tid; i < 2 * CL * slice_bytes / 4; i += 128) ((uint32_t *)smem)[i] = 0;
    __syncthreads();
    cluster_sync();
    long long t0 = clock64();
    int bad = 0;
    for (int r = 0; r < rounds; ++r) {
        const int b = r & 1;
        uint8_t *buf = smem + (size_t)b * CL * slice_bytes;
        uint32_t *mine = (uint32_t *)(buf + (size_t)rank * slice_bytes);
        for (int i = tid; i < slice_bytes / 4; i += 128) mine[i] = (uint32_t)(r * 1000003 + rank * 4099 + i);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&bar[b], (CL - 1) * slice_bytes);
            for (int p = 1; p < CL; ++p) {
                const uint32_t peer = (rank + p) % CL;
                const uint32_t dst = mapa(smem_u32(mine), peer), rbar = mapa(smem_u32(&bar[b]), peer);
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "r"(smem_u32(mine)), "r"(slice_bytes), "r"(rbar) : "memory");
            }
        }
        if (!mbar_wait(&bar[b], (r >> 1) & 1)) { bad = 1; break; }
        if (r == rounds - 1)
            for (int p = 0; p < CL; ++p) {
                const uint32_t *s = (const uint32_t *)(buf + (size_t)p * slice_bytes);
                for (int i = tid; i < slice_bytes / 4; i += 128) if (s[i] != (uint32_t)(r * 1000003 + p * 4099 + i)) bad = 2;
            }
    }
    long long t1 = clock64();
    if (bad) atomicExch(status, bad);
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    cluster_sync();
}

template <int CL>
static void run_dsmem_probe(int slice_bytes)
{
    long long *dc; int *ds;
    CK(cudaMalloc(&dc, 64 * 8)); CK(cudaMalloc(&ds, 4)); CK(cudaMemset(ds, 0, 4));
    const int smem = 2 * CL * slice_bytes, rounds = 200;
    CK(cudaFuncSetAttribute(dsmem_probe<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (CL > 8) CK(cudaFuncSetAttribute(dsmem_probe<CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, dsmem_probe<CL>, rounds, slice_bytes, dc, ds));
    CK(cudaDeviceSynchronize());
    long long c[16]; int st;
    CK(cudaMemcpy(c, dc, CL * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
    printf("B: cluster %d, slice %d B to each of %d peers: %.0f cycles per all-to-all round (in %.1f B/clk per CTA), status %d\n", CL, slice_bytes, CL - 1, (double)c[0] / rounds,
           (double)(CL - 1) * slice_bytes / ((double)c[0] / rounds), st);
    cudaFree(dc); cudaFree(ds);
}

// ---------------------------------------------------------------- C/D: L2 -> smem streaming
template <int CL, bool MC>
__global__ void __launch_bounds__(128, 1) stream_probe(const uint8_t *src, size_t total_bytes, int chunk_bytes, int stages, int passes, long long *cycles, int *status)
{
    extern __shared__ __align__(1024) uint8_t smem[];     // [stages][chunk_bytes]
    __shared__ uint64_t full[8], empty[8];
    const int tid = threadIdx.x;
    const uint32_t rank = CL > 1 ? cluster_rank() : 0;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], MC ? CL : 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (CL > 1) cluster_sync();
    const int nchunks = (int)(total_bytes / chunk_bytes) * passes, per_pass = (int)(total_bytes / chunk_bytes);
    long long t0 = clock64();
    int bad = 0;
    if (tid == 0) {            // producer
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % stages, use = c / stages;
            if (use > 0 && !mbar_wait(&empty[s], (use - 1) & 1)) { bad = 1; break; }
            mbar_expect_tx(&full[s], chunk_bytes);
            const uint8_t *g = src + (size_t)(c % per_pass) * chunk_bytes;
            if (MC) {
                const int part = chunk_bytes / CL;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem + (size_t)s * chunk_bytes + rank * part)),
                             "l"(g + rank * part), "r"(part), "r"(smem_u32(&full[s])), "h"((uint16_t)((1u << CL) - 1))
                             : "memory");
            } else {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + (size_t)s * chunk_bytes)), "l"(g), "r"(chunk_bytes), "r"(smem_u32(&full[s])) : "memory");
            }
        }
    } else if (tid == 32) {    // consumer: waits, touches nothing, releases
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % stages, use = c / stages;
            if (!mbar_wait(&full[s], use & 1)) { bad = 2; break; }
            if (MC) {
                for (int p = 0; p < CL; ++p) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa(smem_u32(&empty[s]), p)) : "memory");
            } else mbar_arrive(&empty[s]);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (bad) atomicExch(status, bad);
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    if (CL > 1) cluster_sync();
}

template <int CL, bool MC>
static void run_stream_probe(const uint8_t *src, size_t total, int grid, int chunk, int stages)
{
    long long *dc; int *ds;
    CK(cudaMalloc(&dc, 256 * 8)); CK(cudaMalloc(&ds, 4)); CK(cudaMemset(ds, 0, 4));
    const int smem = chunk * stages, passes = 6;
    CK(cudaFuncSetAttribute(stream_probe<CL, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    CK(cudaLaunchKernelEx(&cfg, stream_probe<CL, MC>, src, total, chunk, stages, 1, dc, ds));      // warm L2
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    CK(cudaLaunchKernelEx(&cfg, stream_probe<CL, MC>, src, total, chunk, stages, passes, dc, ds));
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> c(grid); int st;
    CK(cudaMemcpy(c.data(), dc, grid * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
    long long mx = 0; for (auto v : c) mx = v > mx ? v : mx;
    const double bytes = (double)total * passes;
    printf("%s: grid %3d cluster %d chunk %5d x %d stages: %.1f B/clk per CTA into smem, %.2f TB/s aggregate smem fill (%.3f ms), status %d\n", MC ? "D multicast" : "C unicast  ", grid, CL, chunk, stages,
           bytes / (double)mx, bytes * grid / (ms * 1e-3) / 1e12, ms, st);
    cudaFree(dc); cudaFree(ds);
}

// ---------------------------------------------------------------- E: streaming with several producer threads
__global__ void __launch_bounds__(256, 1) stream_probe2(const uint8_t *src, size_t total_bytes, int chunk_bytes, int stages, int nprod, int passes, long long *cycles, int *status)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[16], empty[16];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int per_pass = (int)(total_bytes / chunk_bytes), nchunks = per_pass * passes;
    long long t0 = clock64();
    int bad = 0;
    if (warp < nprod && lane == 0) {
        for (int c = warp; c < nchunks; c += nprod) {
            const int s = c % stages, use = c / stages;
            if (use > 0 && !mbar_wait(&empty[s], (use - 1) & 1)) { bad = 1; break; }
            mbar_expect_tx(&full[s], chunk_bytes);
            const uint8_t *g = src + (size_t)(c % per_pass) * chunk_bytes;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + (size_t)s * chunk_bytes)), "l"(g), "r"(chunk_bytes), "r"(smem_u32(&full[s])) : "memory");
        }
    } else if (warp == 7 && lane == 0) {
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % stages, use = c / stages;
            if (!mbar_wait(&full[s], use & 1)) { bad = 2; break; }
            mbar_arrive(&empty[s]);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (bad) atomicExch(status, bad);
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

static void run_stream_probe2(const uint8_t *src, size_t total, int grid, int chunk, int stages, int nprod)
{
    long long *dc; int *ds;
    CK(cudaMalloc(&dc, 256 * 8)); CK(cudaMalloc(&ds, 4)); CK(cudaMemset(ds, 0, 4));
    const int smem = chunk * stages, passes = 4;
    CK(cudaFuncSetAttribute(stream_probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    stream_probe2<<<grid, 256, smem>>>(src, total, chunk, stages, nprod, 1, dc, ds);
    CK(cudaDeviceSynchronize());
    stream_probe2<<<grid, 256, smem>>>(src, total, chunk, stages, nprod, passes, dc, ds);
    CK(cudaDeviceSynchronize());
    std::vector<long long> c(grid); int st;
    CK(cudaMemcpy(c.data(), dc, grid * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
    long long mx = 0; for (auto v : c) mx = v > mx ? v : mx;
    printf("E: grid %3d chunk %6d x %2d stages, %d producers: %6.1f B/clk per CTA (%.0f clk per chunk), status %d\n", grid, chunk, stages, nprod, (double)total * passes / (double)mx,
           (double)mx / ((double)total * passes / chunk), st);
    cudaFree(dc); cudaFree(ds);
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs, %d MHz\n", prop.name, prop.multiProcessorCount, prop.clockRate / 1000);
    run_mma_probe<64, 64>();
    const size_t total = 8u << 20;
    uint8_t *src; CK(cudaMalloc(&src, total)); CK(cudaMemset(src, 1, total));
    for (int grid : {1, 148})
        for (int chunk : {4096, 8192, 16384, 32768, 65536})
            for (int stages : {2, 4, 8})
                for (int nprod : {1, 2, 4}) {
                    if (chunk * stages > 200 * 1024 || nprod > stages) continue;
                    run_stream_probe2(src, total, grid, chunk, stages, nprod);
                }
    printf("done\n");
    return 0;
}
