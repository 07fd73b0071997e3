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

template <int M, int N, int K>
__global__ void __launch_bounds__(128, 1) mma_probe(const __nv_bfloat16 *a_img, const __nv_bfloat16 *b_img, float *d_out, int swap_offsets, int *status)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __nv_bfloat16 *sa = (__nv_bfloat16 *)smem;                       // 128 x K
    __nv_bfloat16 *sb = (__nv_bfloat16 *)(smem + M * K * 2);       // N x K
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * K / 8; i += 128) ((uint4 *)sa)[i] = ((const uint4 *)a_img)[i];
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
        const uint32_t idesc = make_idesc(M, N);
        for (int j = 0; j < K / 16; ++j) {
            uint32_t a_lbo = M * 16, a_sbo = 128, b_lbo = N * 16, b_sbo = 128;
            if (swap_offsets) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
            const uint64_t da = make_desc(smem_u32(sa) + j * 2 * M * 16, a_lbo, a_sbo);
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

template <int M, int N, int K>
static void run_mma_probe()
{
    std::vector<float> A(M * K), B(N * K);
    for (auto &v : A) v = (float)((rand() % 17) - 8) / 8.0f;
    for (auto &v : B) v = (float)((rand() % 13) - 6) / 4.0f;
    std::vector<__nv_bfloat16> ai, bi;
    pack_kmajor(A, M, K, ai);
    pack_kmajor(B, N, K, bi);
    __nv_bfloat16 *da, *db; float *dd; int *ds;
    CK(cudaMalloc(&da, ai.size() * 2)); CK(cudaMalloc(&db, bi.size() * 2)); CK(cudaMalloc(&dd, 128 * N * 4)); CK(cudaMalloc(&ds, 4));
    CK(cudaMemcpy(da, ai.data(), ai.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, bi.data(), bi.size() * 2, cudaMemcpyHostToDevice));
    const int smem = (M + N) * K * 2;
    CK(cudaFuncSetAttribute(mma_probe<M, N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int swap = 0; swap < 1; ++swap) {
        CK(cudaMemset(dd, 0, 128 * N * 4)); CK(cudaMemset(ds, 0, 4));
        mma_probe<M, N, K><<<1, 128, smem>>>(da, db, dd, swap, ds);
        CK(cudaDeviceSynchronize());
        std::vector<float> D(128 * N); int st;
        CK(cudaMemcpy(D.data(), dd, 128 * N * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
        printf("M=%d N=%d: tensor-memory lane -> D row: ", M, N);
        for (int lane = 0; lane < 128; ++lane) {
            int found = -1;
            for (int m = 0; m < M && found < 0; ++m) {
                bool okr = true;
                for (int n = 0; n < N && okr; ++n) {
                    double ref = 0;
                    for (int k = 0; k < K; ++k) ref += (double)__bfloat162float(__float2bfloat16(A[m * K + k])) * (double)__bfloat162float(__float2bfloat16(B[n * K + k]));
                    okr = fabs(ref - D[lane * N + n]) < 1e-3;
                }
                if (okr) found = m;
            }
            if (lane % 16 == 0) printf("| %d:", lane);
            printf("%d ", found);
        }
        printf("\n");
    }
    cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(ds);
}

int main()
{
    run_mma_probe<128, 32, 64>();
    run_mma_probe<64, 32, 64>();
    return 0;
}
