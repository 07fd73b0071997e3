// umma_issue.cu -- the dense kernel's MMA issue loop in isolation (development microbenchmark): bundles of NK products issued inside
// an elect.sync region, with the per-bundle mbarrier wait / tcgen05 fence / commits of the real kernel switched on one by one.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate));
}
struct Seg { uint16_t off16, rows, nk, bsrc16, dcol, first; };
template <int NK>
__device__ __forceinline__ void issue_run(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t a_inc, uint32_t hi, uint32_t idesc, uint32_t acc0)
{
    tc_mma(d, a_lo, b_lo, hi, idesc, acc0);
#pragma unroll
    for (int k = 1; k < NK; ++k) tc_mma(d, a_lo + k * a_inc, b_lo + k * 64, hi, idesc, 1u);
}
// flags: 1 = wait on the (self-completing) ring barrier per bundle, 2 = tcgen05.fence::after_thread_sync per bundle, 4 = commit to an "empty" barrier
// per bundle (which a producer-like thread turns into the next "full"), 8 = segment record read from shared memory (else registers)
__global__ void __launch_bounds__(384, 1) issue_kernel(int flags, int nk, int nbundles, long long *out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[2], empty[2], done;
    __shared__ uint32_t tmem_s;
    __shared__ Seg segs[4];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 4; i += 384) ((uint32_t *)smem)[i] = 0x3c003c00u;
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s]))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty[s]))); }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&done)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        segs[0] = Seg{0, 128, (uint16_t)nk, 0, 0, 1};
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s, sb = s32(smem);
    const uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (4u << 17) | (8u << 24), DESC_HI = (128u >> 4) | (1u << 14);
    if (warp == 1 && (tid & 31) == 0 && (flags & 4)) {
        // stands in for the producer: slot s becomes full again as soon as its products have completed
        for (int b = 0; b < nbundles; ++b) {
            const int s = b & 1, use = b >> 1;
            if (use > 0) while (!try_wait(s32(&empty[s]), (use - 1) & 1)) {}
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&full[s])) : "memory");
        }
    } else if (warp == 8) {
        unsigned ph = 0;
        const long long t0 = clock64();
        if (flags & 256) {
            if (elect_one()) {
                uint32_t na = ((sb & 0x3FFFFu) >> 4) | ((2048u >> 4) << 16), nb_ = (((sb + 65536u) & 0x3FFFFu) >> 4) | ((512u >> 4) << 16);
                for (int b = 0; b < nbundles; ++b) {
                    const uint32_t a_lo0 = na, b_lo0 = nb_;
                    // descriptors of the NEXT bundle, computed before this bundle's products so that the set-up hides behind them
                    const int nslot = (b + 1) & 1;
                    Seg sg2;
                    if (flags & 8) sg2 = segs[0]; else sg2 = Seg{0, 128, (uint16_t)nk, 0, 0, 1};
                    na = (((sb + nslot * 32768 + sg2.off16 * 16u) & 0x3FFFFu) >> 4) | ((sg2.rows * 16u >> 4) << 16);
                    nb_ = (((sb + 65536u + sg2.bsrc16 * 16u) & 0x3FFFFu) >> 4) | ((512u >> 4) << 16);
                    const int slot = b & 1;
                    if (flags & 4) { while (!try_wait(s32(&full[slot]), (ph >> slot) & 1u)) {} ph ^= 1u << slot; }
                    if (flags & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    uint32_t a_lo = a_lo0, b_lo = b_lo0;
                    tc_mma(tmem, a_lo, b_lo, DESC_HI, IDESC, 0u);
#pragma unroll 4
                    for (int k = 1; k < nk; ++k) {
                        a_lo += 256;
                        b_lo += 64;
                        tc_mma(tmem, a_lo, b_lo, DESC_HI, IDESC, 1u);
                    }
                    if (flags & 4) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&empty[slot])) : "memory");
                }
            }
            __syncwarp();
        } else         if (flags & 16) {
            if (elect_one()) {
                for (int b = 0; b < nbundles; ++b) {
                    const int slot = b & 1;
                    if (flags & 4) { while (!try_wait(s32(&full[slot]), (ph >> slot) & 1u)) {} ph ^= 1u << slot; }
                    if (flags & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t slot_base = sb + slot * 32768;
                    Seg sg;
                    if (flags & 8) sg = segs[0]; else sg = Seg{0, 128, (uint16_t)nk, 0, (uint16_t)((flags & 128) ? 64 * (b & 1) : 0), (uint16_t)((flags & 64) ? 0 : 1)};
                    const uint32_t rows = sg.rows, d = tmem + sg.dcol;
                    uint32_t a_lo = (((slot_base + sg.off16 * 16u) & 0x3FFFFu) >> 4) | ((rows * 16u >> 4) << 16);
                    uint32_t b_lo = (((sb + 65536u + sg.bsrc16 * 16u) & 0x3FFFFu) >> 4) | ((512u >> 4) << 16);
                    const uint32_t a_inc = rows * 2;
                    const int n = sg.nk;
                    if (flags & 32) {
                        if (n == 8) issue_run<8>(d, a_lo, b_lo, a_inc, DESC_HI, IDESC, sg.first ? 0u : 1u);
                        else if (n == 16) issue_run<16>(d, a_lo, b_lo, a_inc, DESC_HI, IDESC, sg.first ? 0u : 1u);
                        else issue_run<1>(d, a_lo, b_lo, a_inc, DESC_HI, IDESC, sg.first ? 0u : 1u);
                    } else {
                        tc_mma(d, a_lo, b_lo, DESC_HI, IDESC, sg.first ? 0u : 1u);
#pragma unroll 4
                        for (int k = 1; k < n; ++k) {
                            a_lo += a_inc;
                            b_lo += 64;
                            tc_mma(d, a_lo, b_lo, DESC_HI, IDESC, 1u);
                        }
                    }
                    if (flags & 4) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&empty[slot])) : "memory");
                }
            }
            __syncwarp();
        } else
        for (int b = 0; b < nbundles; ++b) {
            const int slot = b & 1;
            if (flags & 4) { while (!try_wait(s32(&full[slot]), (ph >> slot) & 1u)) {} ph ^= 1u << slot; }
            else if (flags & 1) { while (!try_wait(s32(&done), 1)) {} }
            if (flags & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t slot_base = sb + slot * 32768;
            if (elect_one()) {
                Seg sg;
                if (flags & 8) sg = segs[0]; else sg = Seg{0, 128, (uint16_t)nk, 0, (uint16_t)((flags & 128) ? 64 * (b & 1) : 0), (uint16_t)((flags & 64) ? 0 : 1)};
                const uint32_t rows = sg.rows, d = tmem + sg.dcol;
                uint32_t a_lo = (((slot_base + sg.off16 * 16u) & 0x3FFFFu) >> 4) | ((rows * 16u >> 4) << 16);
                uint32_t b_lo = (((sb + 65536u + sg.bsrc16 * 16u) & 0x3FFFFu) >> 4) | ((512u >> 4) << 16);
                const uint32_t a_inc = rows * 2;
                tc_mma(d, a_lo, b_lo, DESC_HI, IDESC, sg.first ? 0u : 1u);
                const int n = sg.nk;
#pragma unroll 4
                for (int k = 1; k < n; ++k) {
                    a_lo += a_inc;
                    b_lo += 64;
                    tc_mma(d, a_lo, b_lo, DESC_HI, IDESC, 1u);
                }
                if (flags & 4) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&empty[slot])) : "memory");
            }
            __syncwarp();
        }
        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&done)) : "memory");
        __syncwarp();
        const long long t1 = clock64();
        while (!try_wait(s32(&done), 0)) {}
        const long long t2 = clock64();
        if ((tid & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
int main()
{
    long long *d, h[2];
    CK(cudaMalloc(&d, 16));
    CK(cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int nb = 500;
    const int fl[] = {0, 16, 30, 256, 256 + 8, 256 + 14};
    const char *nm[] = {"bare, elect region per bundle", "bare, ONE elect region", "the kernel, ONE elect region", "software-pipelined set-up, bare", "software-pipelined, record from shared memory", "software-pipelined, the kernel (ring + fence + record)"};
    for (int v = 0; v < 6; ++v)
        for (int nk : {8, 16}) {
            issue_kernel<<<1, 384, 200 * 1024>>>(fl[v], nk, nb, d);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
            printf("%-58s bundles of %2d: issue %.1f, complete %.1f clk/mma\n", nm[v], nk, (double)h[0] / (nb * nk), (double)h[1] / (nb * nk));
        }
    return 0;
}
