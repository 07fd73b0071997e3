// Round-2 microbenchmark (development aid, not product): grid-level all-gather of a [512 units][F folds] fp32
// vector in two wire formats,
//   mode 0  "sector":  32-byte sectors {7 values, epoch}, ONE 256-bit store per sector (st.global.v8.b32, sm_100),
//                      readers poll with ONE 256-bit load per sector; NSEC sectors per unit (7 * NSEC folds)
//   mode 1  "LL pair": 8-byte {value, epoch} pairs (the round-1 format), 8 * NSEC pairs per unit,
// and an integrity check of the sector format: every value carries (iteration, unit, slot), so a torn sector
// (epoch of iteration i next to a value of iteration i-2) is detected and counted.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sector_exchange sector_exchange.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int NT = 512, NPROD = 128, UNITS = 4;

struct Sec { unsigned v[8]; };
__device__ __forceinline__ Sec ld_sector(const unsigned *p)
{
    Sec s;
    asm volatile("ld.relaxed.gpu.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(s.v[0]), "=r"(s.v[1]), "=r"(s.v[2]), "=r"(s.v[3]), "=r"(s.v[4]), "=r"(s.v[5]), "=r"(s.v[6]), "=r"(s.v[7]) : "l"(p) : "memory");
    return s;
}
__device__ __forceinline__ void st_sector(unsigned *p, const Sec &s)
{
    asm volatile("st.relaxed.gpu.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "r"(s.v[0]), "r"(s.v[1]), "r"(s.v[2]), "r"(s.v[3]), "r"(s.v[4]), "r"(s.v[5]), "r"(s.v[6]), "r"(s.v[7]) : "memory");
}
__device__ __forceinline__ uint4 ld_pairs2(const void *p) { uint4 v; asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_pair(void *p, unsigned v, unsigned e) { asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v), "r"(e) : "memory"); }

__device__ __forceinline__ unsigned tag(int it, int unit, int slot) { return ((unsigned)it << 16) ^ ((unsigned)unit << 5) ^ (unsigned)slot; }

// buffers: 2 alternating vectors.  sector mode: [2][512][NSEC][8] words; LL mode: [2][512][8*NSEC][2] words
template <int NSEC>
__global__ void __launch_bounds__(NT, 1) exch(int mode, int iters, unsigned *buf, int *errors, float *sink, long long *cyc)
{
    extern __shared__ float sm[];
    const int tid = threadIdx.x, cta = blockIdx.x;
    int err = 0, cap = 1 << 20;
    float acc = 0.f;
    long long t0 = clock64();
    if (mode == 0) {
        for (int it = 0; it < iters; ++it) {
            const unsigned epoch = (unsigned)it + 1u;
            unsigned *vec = buf + (size_t)(it & 1) * 512 * NSEC * 8;
            if (cta < NPROD && tid < UNITS * NSEC) {
                const int unit = cta * UNITS + tid / NSEC, sec = tid % NSEC;
                Sec s;
#pragma unroll
                for (int j = 0; j < 7; ++j) s.v[j] = tag(it, unit, sec * 7 + j);
                s.v[7] = epoch;
                st_sector(vec + ((size_t)unit * NSEC + sec) * 8, s);
            }
            Sec s[NSEC];
            const unsigned *src = vec + (size_t)tid * NSEC * 8;
#pragma unroll
            for (int j = 0; j < NSEC; ++j) s[j] = ld_sector(src + j * 8);
            for (int spin = 0;; ++spin) {
                bool bad = false;
#pragma unroll
                for (int j = 0; j < NSEC; ++j) {
                    const bool b = s[j].v[7] != epoch;
                    if (b) s[j] = ld_sector(src + j * 8);
                    bad |= b;
                }
                if (!bad) break;
                if (spin > cap) { err += 1000000; cap = 0; break; }
            }
#pragma unroll
            for (int j = 0; j < NSEC; ++j)
#pragma unroll
                for (int q = 0; q < 7; ++q) {
                    if (s[j].v[q] != tag(it, tid, j * 7 + q)) ++err;
                    sm[(tid * NSEC + j) * 8 + q] = __uint_as_float(s[j].v[q]);
                }
            __syncthreads();
            acc += sm[(tid * 37) % (512 * NSEC * 8)] * 1e-30f;
            __syncthreads();
        }
    } else if (mode == 2) {
        // same wire format as mode 0, but a warp's 32 loads cover 32 CONSECUTIVE sectors (1 KB): thread tid polls
        // sectors tid, tid + 512, ... of the vector instead of the NSEC sectors of unit tid
        for (int it = 0; it < iters; ++it) {
            const unsigned epoch = (unsigned)it + 1u;
            unsigned *vec = buf + (size_t)(it & 1) * 512 * NSEC * 8;
            if (cta < NPROD && tid < UNITS * NSEC) {
                const int unit = cta * UNITS + tid / NSEC, sec = tid % NSEC;
                Sec s;
#pragma unroll
                for (int j = 0; j < 7; ++j) s.v[j] = tag(it, unit, sec * 7 + j);
                s.v[7] = epoch;
                st_sector(vec + ((size_t)unit * NSEC + sec) * 8, s);
            }
            Sec s[NSEC];
            const unsigned *src = vec + (size_t)tid * 8;
#pragma unroll
            for (int j = 0; j < NSEC; ++j) s[j] = ld_sector(src + (size_t)j * NT * 8);
            for (int spin = 0;; ++spin) {
                bool bad = false;
#pragma unroll
                for (int j = 0; j < NSEC; ++j) {
                    const bool b = s[j].v[7] != epoch;
                    if (b) s[j] = ld_sector(src + (size_t)j * NT * 8);
                    bad |= b;
                }
                if (!bad) break;
                if (spin > cap) { err += 1000000; cap = 0; break; }
            }
#pragma unroll
            for (int j = 0; j < NSEC; ++j) {
                const int sidx = tid + j * NT, unit = sidx / NSEC, sec = sidx % NSEC;
#pragma unroll
                for (int q = 0; q < 7; ++q) {
                    if (s[j].v[q] != tag(it, unit, sec * 7 + q)) ++err;
                    sm[sidx * 8 + q] = __uint_as_float(s[j].v[q]);
                }
            }
            __syncthreads();
            acc += sm[(tid * 37) % (512 * NSEC * 8)] * 1e-30f;
            __syncthreads();
        }
    } else if (mode == 3) {
        // 16-byte quads {3 values, epoch}: 7 * NSEC folds = ceil(7 NSEC / 3) quads per unit, one st.v4 / ld.v4 each
        constexpr int NQ = (7 * NSEC + 2) / 3;
        constexpr int TOT = 512 * NQ, PER = (TOT + NT - 1) / NT;
        for (int it = 0; it < iters; ++it) {
            const unsigned epoch = (unsigned)it + 1u;
            unsigned *vec = buf + (size_t)(it & 1) * 512 * NSEC * 8;
            if (cta < NPROD && tid < UNITS * NQ) {
                const int unit = cta * UNITS + tid / NQ, q = tid % NQ;
                asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(vec + ((size_t)unit * NQ + q) * 4), "r"(tag(it, unit, q * 3)),
                             "r"(tag(it, unit, q * 3 + 1)), "r"(tag(it, unit, q * 3 + 2)), "r"(epoch) : "memory");
            }
            uint4 v[PER];
#pragma unroll
            for (int j = 0; j < PER; ++j) if (tid + j * NT < TOT) v[j] = ld_pairs2(vec + 4 * (size_t)(tid + j * NT));
            for (int spin = 0;; ++spin) {
                bool bad = false;
#pragma unroll
                for (int j = 0; j < PER; ++j) if (tid + j * NT < TOT) {
                    const bool b = v[j].w != epoch;
                    if (b) v[j] = ld_pairs2(vec + 4 * (size_t)(tid + j * NT));
                    bad |= b;
                }
                if (!bad) break;
                if (spin > cap) { err += 1000000; cap = 0; break; }
            }
#pragma unroll
            for (int j = 0; j < PER; ++j) if (tid + j * NT < TOT) {
                const int qi = tid + j * NT, unit = qi / NQ, q = qi % NQ;
                if (v[j].x != tag(it, unit, q * 3) || v[j].y != tag(it, unit, q * 3 + 1) || v[j].z != tag(it, unit, q * 3 + 2)) ++err;
                *reinterpret_cast<float4 *>(sm + 4 * qi) = make_float4(__uint_as_float(v[j].x), __uint_as_float(v[j].y), __uint_as_float(v[j].z), 0.f);
            }
            __syncthreads();
            acc += sm[(tid * 37) % (TOT * 4)] * 1e-30f;
            __syncthreads();
        }
    } else {
        constexpr int NPAIR = 8 * NSEC;            // pairs per unit
        for (int it = 0; it < iters; ++it) {
            const unsigned epoch = (unsigned)it + 1u;
            unsigned *vec = buf + (size_t)(it & 1) * 512 * NPAIR * 2;
            if (cta < NPROD && tid < UNITS * NPAIR) {
                const int unit = cta * UNITS + tid / NPAIR, slot = tid % NPAIR;
                st_pair(vec + ((size_t)unit * NPAIR + slot) * 2, tag(it, unit, slot), epoch);
            }
            // 512 * NPAIR pairs = 256 * NPAIR 16-byte chunks; 4 in flight per thread
            constexpr int NCH = 256 * NPAIR;
            for (int base = tid; base < NCH; base += 4 * NT) {
                uint4 v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = ld_pairs2(vec + 4 * (size_t)(base + j * NT));
                for (int spin = 0;; ++spin) {
                    bool bad = false;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool b = (v[j].y != epoch) | (v[j].w != epoch);
                        if (b) v[j] = ld_pairs2(vec + 4 * (size_t)(base + j * NT));
                        bad |= b;
                    }
                    if (!bad) break;
                    if (spin > cap) { err += 1000000; cap = 0; break; }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ch = base + j * NT, unit = ch / (NPAIR / 2), slot = (ch % (NPAIR / 2)) * 2;
                    if (v[j].x != tag(it, unit, slot) || v[j].z != tag(it, unit, slot + 1)) ++err;
                    *reinterpret_cast<float2 *>(sm + 2 * ch) = make_float2(__uint_as_float(v[j].x), __uint_as_float(v[j].z));
                }
            }
            __syncthreads();
            acc += sm[(tid * 37) % (512 * NPAIR)] * 1e-30f;
            __syncthreads();
        }
    }
    if (tid == 0) cyc[cta] = clock64() - t0;
    if (err) atomicAdd(errors, err);
    if (acc == 123.f) *sink = acc;
}

template <int NSEC>
static void run(int mode, int ncta, int iters)
{
    unsigned *buf;
    int *errors;
    float *sink;
    long long *cyc;
    const size_t bytes = (size_t)2 * 512 * NSEC * 8 * 8;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMalloc(&errors, 4));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMalloc(&cyc, 8 * 256));
    const size_t smem = 160 * 1024;
    CK(cudaFuncSetAttribute(exch<NSEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    int errs = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(buf, 0, bytes));
        CK(cudaMemset(errors, 0, 4));
        void *args[] = {&mode, &iters, &buf, &errors, &sink, &cyc};
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((const void *)exch<NSEC>, dim3(ncta), dim3(NT), args, smem, 0));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
        int e;
        CK(cudaMemcpy(&e, errors, 4, cudaMemcpyDeviceToHost));
        errs += e;
    }
    printf("ctas %3d  %-8s  %2d folds/unit (%5.1f KB gathered per CTA)  %.3f us/exchange   integrity errors %d\n", ncta,
           mode == 0 ? "sector" : mode == 2 ? "sector-c" : mode == 3 ? "quad" : "LL pair", mode == 1 ? 8 * NSEC : 7 * NSEC,
           (mode == 1 ? 512.0 * NSEC * 64 : mode == 3 ? 512.0 * ((7 * NSEC + 2) / 3) * 16 : 512.0 * NSEC * 32) / 1024.0, best * 1000.f / iters, errs);
    cudaFree(buf); cudaFree(errors); cudaFree(sink); cudaFree(cyc);
}

int main(int argc, char **argv)
{
    const int iters = argc > 1 ? atoi(argv[1]) : 20000;
    for (int ncta : {128}) {
        for (int mode = 0; mode < 4; ++mode) {
            run<1>(mode, ncta, iters);
            run<2>(mode, ncta, iters);
            run<3>(mode, ncta, iters);
            run<4>(mode, ncta, iters);
        }
    }
    return 0;
}
