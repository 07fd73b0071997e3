// L2 round-trip latency of the load flavours usable for polling, and a 2-CTA ping-pong (development aid).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_latency l2_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int F> __device__ __forceinline__ unsigned ld(const unsigned *p)
{
    unsigned v;
    if (F == 0) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (F == 1) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (F == 2) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (F == 3) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (F == 4) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (F == 5) asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
template <int F> __global__ void chase(const unsigned *buf, int n, long long *out)
{
    // buf[i] = (i + stride) % size : dependent loads
    unsigned idx = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) idx = ld<F>(buf + idx);
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (t1 - t0) / n; out[1] = idx; }
}
// ping-pong between CTA 0 and CTA `peer`: one-way latency = total / (2 n)
template <int ST> __device__ __forceinline__ void st(unsigned *p, unsigned v)
{
    if (ST == 0) asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (ST == 1) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (ST == 2) asm volatile("st.global.cg.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (ST == 3) asm volatile("st.global.wt.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (ST == 4) { unsigned o; asm volatile("atom.global.exch.b32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory"); }
    if (ST == 5) asm volatile("red.global.max.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (ST == 6) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <int ST> __global__ void pingpong2(unsigned *flags, int n, int peer, long long *out)
{
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    if (me != 0 && me != peer) return;
    unsigned *mine = flags + (me == 0 ? 0 : 32), *other = flags + (me == 0 ? 32 : 0);
    long long t0 = clock64();
    for (int i = 1; i <= n; ++i) {
        if (me == 0) {
            st<ST>(other, (unsigned)i);
            while (ld<0>(mine) != (unsigned)i) { }
        } else {
            while (ld<0>(mine) != (unsigned)i) { }
            st<ST>(other, (unsigned)i);
        }
    }
    long long t1 = clock64();
    if (me == 0) out[0] = (t1 - t0) / n;
}
template <int F> __global__ void pingpong(unsigned *flags, int n, int peer, long long *out)
{
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    if (me != 0 && me != peer) return;
    unsigned *mine = flags + (me == 0 ? 0 : 32), *other = flags + (me == 0 ? 32 : 0);
    long long t0 = clock64();
    for (int i = 1; i <= n; ++i) {
        if (me == 0) {
            asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(other), "r"((unsigned)i) : "memory");
            while (ld<F>(mine) != (unsigned)i) { }
        } else {
            while (ld<F>(mine) != (unsigned)i) { }
            asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(other), "r"((unsigned)i) : "memory");
        }
    }
    long long t1 = clock64();
    if (me == 0) out[0] = (t1 - t0) / n;
}
int main()
{
    const int size = 1 << 20;   // 4 MiB: stays in L2
    unsigned *h = new unsigned[size];
    for (int i = 0; i < size; ++i) h[i] = (unsigned)((i + 4099 * 32) % size);
    unsigned *buf, *flags; long long *out;
    CK(cudaMalloc(&buf, size * 4)); CK(cudaMalloc(&flags, 256)); CK(cudaMalloc(&out, 64));
    CK(cudaMemcpy(buf, h, size * 4, cudaMemcpyHostToDevice));
    const char *names[] = {"ld.volatile", "ld.relaxed.gpu", "ld.global.cg", "ld.global.cv", "ld.acquire.gpu", "ld.global.ca"};
    long long r[2];
#define RUN(F) chase<F><<<1, 32>>>(buf, 2000, out); CK(cudaDeviceSynchronize()); chase<F><<<1, 32>>>(buf, 2000, out); CK(cudaDeviceSynchronize()); \
    CK(cudaMemcpy(r, out, 16, cudaMemcpyDeviceToHost)); printf("dependent %-16s %lld cycles per load\n", names[F], r[0]);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
#define PP(F, peer) CK(cudaMemset(flags, 0, 256)); pingpong<F><<<148, 32>>>(flags, 2000, peer, out); CK(cudaDeviceSynchronize()); \
    CK(cudaMemcpy(r, out, 8, cudaMemcpyDeviceToHost)); printf("ping-pong %-16s CTA0<->CTA%-3d %lld cycles per round trip (2 stores + 2 successful polls)\n", names[F], peer, r[0]);
    PP(0, 1) PP(0, 74) PP(0, 147) PP(1, 1) PP(1, 74) PP(1, 147) PP(2, 74) PP(4, 74)
    const char *sn[] = {"st.volatile", "st.relaxed.gpu", "st.global.cg", "st.global.wt", "atom.exch", "red.max", "st.release.gpu"};
#define PP2(ST) CK(cudaMemset(flags, 0, 256)); pingpong2<ST><<<148, 32>>>(flags, 2000, 74, out); CK(cudaDeviceSynchronize()); \
    CK(cudaMemcpy(r, out, 8, cudaMemcpyDeviceToHost)); printf("ping-pong store %-16s + ld.volatile: %lld cycles per round trip\n", sn[ST], r[0]);
    PP2(0) PP2(1) PP2(2) PP2(3) PP2(4) PP2(5) PP2(6)
    return 0;
}
