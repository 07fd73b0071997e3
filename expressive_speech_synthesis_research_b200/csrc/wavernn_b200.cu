// C-ABI implementation (include/wavernn_b200.h): handle management, weight repack, launches.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include "../../include/wavernn_b200.h"
#include "wavernn_kernel.cuh"
#include "wavernn_dense.cuh"
#include "wavernn_wide.cuh"
#include "wavernn_cond.cuh"

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <functional>
#include <string>
#include <vector>

using namespace wrnn;

static thread_local std::string g_err;
static std::atomic<long long> g_epilogue_launches{0};

static int32_t fail(int32_t code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CUDA_TRY(expr)                                                                             \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) return fail(WRNN_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

struct wrnn_handle {
    wrnn_config cfg;
    int device = 0, sm_count = 0;
    int rows5 = 4, nprod5 = 128, cpad = 512, n_u = 1;
    int smem_bytes = 0, smem_limit = 0;   // dynamic shared memory of the last launch / opt-in limit of the device
    int last_teams = 1;
    int bf16w = 0;                        // precision bf16: item images hold bf16 weights
    bool loaded = false;
    float *wimg = nullptr;
    unsigned long long *xb = nullptr;   // LL exchange buffers
    int *status = nullptr;
    int *status_host = nullptr;      // pinned copy of `status`, filled by the D2H that closes every call
    bool pending = false;            // a generate call has been enqueued and not yet waited for
    long long *fold_dev = nullptr;   // [2][fold_cap]
    int fold_cap = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0, epilogue_launches = 0;
    int last_status = 0;
    float last_ms = 0.f;
    long long *prof_dev = nullptr;   // [NCTA][PROF_SLOTS], allocated by wrnn_set_profiling
    bool profiling = false;
    // wide kernel (csrc/wavernn_wide.cuh): all folds of a launch through one exchange per stage; fp32 RAW-512 / MOL
    int wide = 0, wide_nsamp = 0;
    int last_kernel = 0;                  // 0 grouped (round-1) kernel, 1 wide kernel, 2 dense kernel
    int kernel_choice = -1;               // wrnn_set_kernel: -1 by fold count, 0 grouped, 1 wide
    float *wide_img = nullptr;            // [NWORK][IMG_FLOATS]
    unsigned *wide_xb = nullptr;          // XW_TOTAL words
    int *progress_host = nullptr, *progress_dev = nullptr;   // mapped step counter (wrnn_progress)
    // precision bf16-dense (csrc/wavernn_dense.cuh): streamed operand tiles, bundle table, per-row fp32 vectors
    int dense = 0, dense_nb = 0, dense_clusters = 0;
    unsigned dense_stream_bytes = 0;
    unsigned char *dense_stream = nullptr;
    wrnn_dense::Bundle *dense_table = nullptr;
    float *dense_sv = nullptr;
    long long *dense_prof = nullptr;       // [CTAs of the launch][PROF_N], development profiling
    size_t dense_prof_slots = 0;
};

// Every entry point that touches a device restores the caller's current device on return.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

extern "C" int32_t wrnn_abi_version(void) { return WRNN_ABI_VERSION; }
extern "C" const char *wrnn_last_error(void) { return g_err.c_str(); }

// fold_with_overlap index arithmetic -- fatchord_version.py:298-309
extern "C" int32_t wrnn_fold_index(int64_t total_len, int64_t target, int64_t overlap,
                                   int64_t *num_folds, int64_t *padded_len)
{
    if (!num_folds || !padded_len) return fail(WRNN_ERR_INVALID, "null output pointer");
    if (target < 0 || overlap < 0 || target + overlap <= 0 || total_len < 0)
        return fail(WRNN_ERR_INVALID, "fold_index: need target >= 0, overlap >= 0, target+overlap > 0, total_len >= 0");
    const int64_t hop = target + overlap;
    int64_t num = total_len - overlap, n = num / hop;
    if (num % hop != 0 && num < 0) --n;                        // python floor division
    const int64_t remaining = total_len - (n * hop + overlap);
    int64_t plen = total_len;
    if (remaining != 0) {
        n += 1;
        plen = total_len + (target + 2 * overlap - remaining);
    }
    *num_folds = n;
    *padded_len = plen;
    return WRNN_OK;
}

static int32_t derive_layout(const wrnn_config &c, int &rows5, int &nprod5, int &n_u)
{
    if (c.rnn_dims != HID || c.fc_dims != HID)
        return fail(WRNN_ERR_INVALID, "this build keeps rnn_dims == fc_dims == %d resident (got %d / %d)", HID, c.rnn_dims, c.fc_dims);
    if (c.feat_dims != 80 || c.aux_dims != 32)
        return fail(WRNN_ERR_INVALID, "conditioning layout is fixed to feat_dims 80 + 4 x aux_dims 32 (got %d, %d)", c.feat_dims, c.aux_dims);
    if (c.precision != WRNN_PREC_FP32 && c.precision != WRNN_PREC_BF16 && c.precision != WRNN_PREC_BF16_DENSE)
        return fail(WRNN_ERR_INVALID, "unknown precision %d", c.precision);
    if (c.precision == WRNN_PREC_BF16_DENSE && !((c.mode == WRNN_MODE_RAW && c.n_classes == wrnn_dense::NCLASS) || (c.mode == WRNN_MODE_MOL && c.n_classes == wrnn_dense::MOL_C)))
        return fail(WRNN_ERR_INVALID, "precision bf16-dense (tcgen05 path) is built for RAW with %d classes and MOL with %d outputs", wrnn_dense::NCLASS, wrnn_dense::MOL_C);
    if (c.mode == WRNN_MODE_RAW) {
        const int C = c.n_classes;
        if (C != 64 && C != 128 && C != 256 && C != 512 && C != 1024)
            return fail(WRNN_ERR_INVALID, "RAW n_classes must be 2**bits with bits in 6..10 (got %d)", C);
        rows5 = C > 512 ? C / NCTA : 4;
        nprod5 = C / rows5;
        n_u = 1;
    } else if (c.mode == WRNN_MODE_MOL) {
        if (c.n_classes != 30) return fail(WRNN_ERR_INVALID, "MOL n_classes must be 30 (got %d)", c.n_classes);
        rows5 = 4;
        nprod5 = 8;
        n_u = 11;
    } else
        return fail(WRNN_ERR_INVALID, "unknown mode %d", c.mode);
    return WRNN_OK;
}

// one kernel per (teams per CTA, model geometry, weight image precision); profiling exists for the generic fp32 kernels
#define K3(stem) {(const void *)stem, (const void *)stem##_t2, (const void *)stem##_t3}
static const void *persistent_kernel(int T, int model, int bf16w, int prof)
{
    static const void *plain[3][2][MAXT] = {{K3(wavernn_persistent_kernel_any), K3(wavernn_persistent_kernel_any_bf16w)},
                                            {K3(wavernn_persistent_kernel), K3(wavernn_persistent_kernel_bf16w)},
                                            {K3(wavernn_persistent_kernel_mol), K3(wavernn_persistent_kernel_mol_bf16w)}};
    static const void *profk[MAXT] = K3(wavernn_persistent_kernel_prof);
    if (prof && !bf16w) return profk[T - 1];
    return plain[model][bf16w ? 1 : 0][T - 1];
}
#undef K3
static bool wide_supported(const wrnn_config &c)
{
    return c.precision == WRNN_PREC_FP32 && ((c.mode == WRNN_MODE_RAW && c.n_classes == 512) || (c.mode == WRNN_MODE_MOL && c.n_classes == 30));
}

static int model_of(const wrnn_config &c) { return c.mode == WRNN_MODE_MOL ? 2 : c.n_classes == 512 ? 1 : 0; }

extern "C" int32_t wrnn_create(const wrnn_config *cfg, int32_t device, wrnn_handle **out)
{
    if (!cfg || !out) return fail(WRNN_ERR_INVALID, "null argument");
    int rows5, nprod5, n_u;
    int32_t rc = derive_layout(*cfg, rows5, nprod5, n_u);
    if (rc) return rc;
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(WRNN_ERR_CUDA, "device %d not present (%d CUDA devices)", device, ndev);
    DeviceGuard guard_(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(WRNN_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    if (prop.multiProcessorCount < NCTA) return fail(WRNN_ERR_CUDA, "need >= %d SMs, device has %d", NCTA, prop.multiProcessorCount);
    int coop = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    if (!coop) return fail(WRNN_ERR_CUDA, "device does not support cooperative launch");

    wrnn_handle *h = new wrnn_handle();
    h->cfg = *cfg;
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->rows5 = rows5;
    h->nprod5 = nprod5;
    h->n_u = n_u;
    h->cpad = rows5 * nprod5;
    h->smem_limit = (int)prop.sharedMemPerBlockOptin;
    h->bf16w = cfg->precision == WRNN_PREC_BF16 ? 1 : 0;
    h->smem_bytes = smem_map(rows5, cfg->mode, cfg->n_classes, 1, 1, h->bf16w).total * (int)sizeof(float);   // smallest configuration
    if (h->smem_bytes > h->smem_limit) {
        const int need = h->smem_bytes;
        delete h;
        return fail(WRNN_ERR_CUDA, "kernel needs %d B shared memory, device allows %zu", need, prop.sharedMemPerBlockOptin);
    }
    cudaError_t e;
#define H_TRY(expr)                                                                  \
    if ((e = (expr)) != cudaSuccess) {                                               \
        wrnn_destroy(h);                                                             \
        return fail(WRNN_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e));          \
    }
    for (int pr = 0; pr < 2; ++pr)
        for (int T = 1; T <= MAXT; ++T)
            H_TRY(cudaFuncSetAttribute(persistent_kernel(T, model_of(*cfg), h->bf16w, pr), cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_limit));
    H_TRY(cudaFuncSetAttribute(wavernn_exchange_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_limit));
    int occ = 0;
    H_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, persistent_kernel(1, model_of(*cfg), h->bf16w, 0), NTHREADS, h->smem_bytes));
    if (occ < 1) {
        wrnn_destroy(h);
        return fail(WRNN_ERR_CUDA, "persistent kernel does not fit on an SM");
    }
    H_TRY(cudaMalloc(&h->wimg, (size_t)NCTA * w_image_floats(rows5, h->bf16w) * sizeof(float)));
    H_TRY(cudaMalloc(&h->xb, (size_t)MAXG * xb_group(h->cpad) * sizeof(unsigned long long) ));
    H_TRY(cudaMalloc(&h->status, 4 * sizeof(int)));
    H_TRY(cudaHostAlloc(&h->status_host, 4 * sizeof(int), cudaHostAllocDefault));
    memset(h->status_host, 0, 4 * sizeof(int));
    H_TRY(cudaHostAlloc(&h->progress_host, sizeof(int), cudaHostAllocMapped));
    *h->progress_host = 0;
    H_TRY(cudaHostGetDevicePointer((void **)&h->progress_dev, h->progress_host, 0));
    if (wide_supported(*cfg) && prop.multiProcessorCount > wrnn_wide::NWORK && wrnn_wide::SM_BYTES <= h->smem_limit) {
        // the wide kernel: 128 worker CTAs + up to 20 sampler CTAs, one per SM (cooperative launch)
        const void *wk[] = {(const void *)wrnn_wide::wavernn_wide_kernel, (const void *)wrnn_wide::wavernn_wide_kernel_mol,
                            (const void *)wrnn_wide::wavernn_wide_kernel_prof, (const void *)wrnn_wide::wavernn_wide_kernel_mol_prof,
                            (const void *)wrnn_wide::wavernn_wide_probe_kernel};
        for (const void *k : wk) H_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, wrnn_wide::SM_BYTES));
        int wocc = 0;
        H_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wocc, wk[0], wrnn_wide::WTHREADS, wrnn_wide::SM_BYTES));
        if (wocc >= 1) {
            h->wide = 1;
            h->wide_nsamp = prop.multiProcessorCount - wrnn_wide::NWORK;
            if (h->wide_nsamp > wrnn_wide::MAXSAMP) h->wide_nsamp = wrnn_wide::MAXSAMP;
            H_TRY(cudaMalloc(&h->wide_img, (size_t)wrnn_wide::NWORK * wrnn_wide::IMG_FLOATS * sizeof(float)));
            H_TRY(cudaMalloc(&h->wide_xb, (size_t)wrnn_wide::XW_TOTAL * sizeof(unsigned)));
        }
    }
    H_TRY(cudaEventCreate(&h->ev0));
    H_TRY(cudaEventCreate(&h->ev1));
    if (cfg->precision == WRNN_PREC_BF16_DENSE) {
        h->dense = 1;
        if (wrnn_dense::SM_TOTAL > h->smem_limit) {
            wrnn_destroy(h);
            return fail(WRNN_ERR_CUDA, "dense kernel needs %d B shared memory, device allows %d", wrnn_dense::SM_TOTAL, h->smem_limit);
        }
        H_TRY(cudaFuncSetAttribute(wrnn_dense::wavernn_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wrnn_dense::SM_TOTAL));
        H_TRY(cudaFuncSetAttribute(wrnn_dense::wavernn_dense_kernel_prof, cudaFuncAttributeMaxDynamicSharedMemorySize, wrnn_dense::SM_TOTAL));
        H_TRY(cudaFuncSetAttribute(wrnn_dense::wavernn_dense_kernel_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, wrnn_dense::SM_TOTAL));
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(wrnn_dense::CL * 64);
        lc.blockDim = dim3(wrnn_dense::DTHREADS);
        lc.dynamicSmemBytes = wrnn_dense::SM_TOTAL;
        int ncl = 0;
        H_TRY(cudaOccupancyMaxActiveClusters(&ncl, wrnn_dense::wavernn_dense_kernel, &lc));
        if (ncl < 1) {
            wrnn_destroy(h);
            return fail(WRNN_ERR_CUDA, "no cluster of %d CTAs of the dense kernel fits on this device", wrnn_dense::CL);
        }
        h->dense_clusters = ncl;
    }
#undef H_TRY
    *out = h;
    return WRNN_OK;
}

extern "C" void wrnn_destroy(wrnn_handle *h)
{
    if (!h) return;
    DeviceGuard guard_(h->device);
    if (h->pending && h->ev1) cudaEventSynchronize(h->ev1);
    cudaFree(h->wimg);
    cudaFree(h->xb);
    cudaFree(h->status);
    if (h->status_host) cudaFreeHost(h->status_host);
    if (h->progress_host) cudaFreeHost(h->progress_host);
    cudaFree(h->wide_img);
    cudaFree(h->wide_xb);
    cudaFree(h->fold_dev);
    cudaFree(h->prof_dev);
    cudaFree(h->dense_stream);
    cudaFree(h->dense_table);
    cudaFree(h->dense_sv);
    cudaFree(h->dense_prof);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    delete h;
}

// ---- weight repack ---------------------------------------------------------------------------
// out[M][N] = A[M][K(lda)] (fp32) * Bm[K][N] (fp64), accumulated in fp64
static void gemm_f64(const float *A, int lda, int M, int K, const std::vector<double> &Bm, int N, std::vector<double> &out)
{
    out.assign((size_t)M * N, 0.0);
    for (int i = 0; i < M; ++i) {
        double *o = &out[(size_t)i * N];
        const float *a = A + (size_t)i * lda;
        for (int k = 0; k < K; ++k) {
            const double av = a[k];
            const double *b = &Bm[(size_t)k * N];
            for (int j = 0; j < N; ++j) o[j] += av * b[j];
        }
    }
}

// write one item image: rows r0..r0+3 (callback per row), columns kbase..kbase+127.
// Slot r of lane l holds row r ^ (l >> 3): the permutation that makes the kernel's butterfly
// reduce-scatter select-free (wavernn_kernel.cuh, reduce_scatter32).
template <class F>
static void put_item(float *dst, int kbase, F rowval)
{
    for (int r = 0; r < 4; ++r)
        for (int l = 0; l < 32; ++l)
            for (int i = 0; i < 4; ++i) dst[(r * 32 + l) * 4 + i] = (float)rowval(r ^ (l >> 3), kbase + l + 32 * i);
}

// Build the per-CTA shared-memory images (host, fp64 folding).  img: [NCTA][w_total(rows5)].
static void pack_images(int C, int rows5, const wrnn_weights *w, std::vector<float> &img)
{
    const int R = HID, F = 80, A = 32, KI = 1 + F + A, KC = F + A;   // I: [512, 113]

    // folded products in fp64:  Wc = I.weight[:, 1:], w0 = I.weight[:, 0]
    std::vector<double> Wc((size_t)R * KC), w0(R), bI(R);
    for (int j = 0; j < R; ++j) {
        w0[j] = w->I_w[(size_t)j * KI];
        bI[j] = w->I_b[j];
        for (int k = 0; k < KC; ++k) Wc[(size_t)j * KC + k] = w->I_w[(size_t)j * KI + 1 + k];
    }
    std::vector<double> G1, G2, G3, u1, u2, u3, c1, c2, c3;
    gemm_f64(w->r1_wih, R, 3 * R, R, Wc, KC, G1);
    gemm_f64(w->r2_wih, R + A, 3 * R, R, Wc, KC, G2);
    gemm_f64(w->fc1_w, R + A, R, R, Wc, KC, G3);
    gemm_f64(w->r1_wih, R, 3 * R, R, w0, 1, u1);
    gemm_f64(w->r2_wih, R + A, 3 * R, R, w0, 1, u2);
    gemm_f64(w->fc1_w, R + A, R, R, w0, 1, u3);
    gemm_f64(w->r1_wih, R, 3 * R, R, bI, 1, c1);
    gemm_f64(w->r2_wih, R + A, 3 * R, R, bI, 1, c2);
    gemm_f64(w->fc1_w, R + A, R, R, bI, 1, c3);

    const size_t per = (size_t)w_total(rows5);
    img.assign((size_t)NCTA * per, 0.f);
    for (int c = 0; c < NCTA; ++c) {
        float *base = &img[(size_t)c * per];
        const int j0 = UNITS * c;
        // M2: rg 0..2 = Wih2[:, :512] gate rows, rg 3..5 = Whh1 gate rows
        for (int rg = 0; rg < 6; ++rg)
            for (int kc = 0; kc < 4; ++kc)
                put_item(base + W_M2 + (rg * 4 + kc) * ITEM, kc * 128, [&](int r, int k) -> double {
                    const int row = j0 + r + R * (rg % 3);
                    return rg < 3 ? w->r2_wih[(size_t)row * (R + A) + k] : w->r1_whh[(size_t)row * R + k];
                });
        // M3: rg 0 = fc1[:, :512] rows, rg 1..3 = Whh2 gate rows
        for (int rg = 0; rg < 4; ++rg)
            for (int kc = 0; kc < 4; ++kc)
                put_item(base + W_M3 + (rg * 4 + kc) * ITEM, kc * 128, [&](int r, int k) -> double {
                    if (rg == 0) return w->fc1_w[(size_t)(j0 + r) * (R + A) + k];
                    return w->r2_whh[(size_t)(j0 + r + R * (rg - 1)) * R + k];
                });
        for (int kc = 0; kc < 4; ++kc)
            put_item(base + W_M4 + kc * ITEM, kc * 128, [&](int r, int k) -> double { return w->fc2_w[(size_t)(j0 + r) * (R + A) + k]; });
        for (int rg = 0; rg < rows5 / 4; ++rg)
            for (int kc = 0; kc < 4; ++kc)
                put_item(base + W_M5 + (rg * 4 + kc) * ITEM, kc * 128, [&](int r, int k) -> double {
                    const int cls = rows5 * c + rg * 4 + r;
                    return cls < C ? w->fc3_w[(size_t)cls * R + k] : 0.0;
                });
        // conditioning items; cond k: [0,80) mel | [80,112) a1 | [112,144) a2 | [144,176) a3 | [176,208) a4
        float *mc = base + w_mc(rows5);
        for (int rg = 0; rg < 3; ++rg)                                        // P1 (chunk A)
            put_item(mc + rg * ITEM, 0, [&](int r, int k) -> double { return k < KC ? G1[(size_t)(j0 + r + R * rg) * KC + k] : 0.0; });
        for (int rg = 0; rg < 3; ++rg)                                        // P2 (chunks A, B)
            for (int ch = 0; ch < 2; ++ch)
                put_item(mc + (3 + rg * 2 + ch) * ITEM, ch * 128, [&](int r, int k) -> double {
                    const int row = j0 + r + R * rg;
                    if (k < KC) return G2[(size_t)row * KC + k];
                    if (k < KC + A) return w->r2_wih[(size_t)row * (R + A) + R + (k - KC)];
                    return 0.0;
                });
        for (int ch = 0; ch < 2; ++ch)                                        // P3 (chunks A, B)
            put_item(mc + (9 + ch) * ITEM, ch * 128, [&](int r, int k) -> double {
                const int row = j0 + r;
                if (k < KC) return G3[(size_t)row * KC + k];
                if (k >= KC + A && k < KC + 2 * A) return w->fc1_w[(size_t)row * (R + A) + R + (k - KC - A)];
                return 0.0;
            });
        put_item(mc + 11 * ITEM, 128, [&](int r, int k) -> double {           // P4 (chunk B)
            const int row = j0 + r;
            if (k >= KC + 2 * A && k < KC + 3 * A) return w->fc2_w[(size_t)row * (R + A) + R + (k - KC - 2 * A)];
            return 0.0;
        });
        float *sv = base + w_small(rows5);
        for (int q = 0; q < 3; ++q)
            for (int u = 0; u < UNITS; ++u) {
                const int row = j0 + u + R * q;
                sv[SV_U1 + q * 4 + u] = (float)u1[row];
                sv[SV_B1 + q * 4 + u] = (float)(c1[row] + (double)w->r1_bih[row]);
                sv[SV_BHH1 + q * 4 + u] = w->r1_bhh[row];
                sv[SV_U2 + q * 4 + u] = (float)u2[row];
                sv[SV_B2 + q * 4 + u] = (float)(c2[row] + (double)w->r2_bih[row]);
                sv[SV_BHH2 + q * 4 + u] = w->r2_bhh[row];
            }
        for (int u = 0; u < UNITS; ++u) {
            sv[SV_U3 + u] = (float)u3[j0 + u];
            sv[SV_B3 + u] = (float)(c3[j0 + u] + (double)w->fc1_b[j0 + u]);
            sv[SV_B4 + u] = w->fc2_b[j0 + u];
        }
        for (int r = 0; r < rows5; ++r) {
            const int cls = rows5 * c + r;
            sv[SV_B5 + r] = cls < C ? w->fc3_b[cls] : 0.f;
        }
    }
}

// ---- wide kernel images (csrc/wavernn_wide.cuh) --------------------------------------------------
// Same algebra as pack_images (input layer folded in fp64, rounded once to fp32); different layouts:
//   gate matrices  [warp 16][ig 4][ks 2][unit 4][ii 4][gate 3]   k = 32 warp + 2 (4 ig + ii) + ks   (Wih2x, Whh1, Whh2)
//   fc matrices    [k 512][unit 4]                                                                   (fc1, fc2, fc3)
//   conditioning   [k' 176][row block 8][4]: row block = which * 4 + unit, rows {P1 r, z, n, P3} (which 0) or
//                  {P2 r, z, n, P4} (which 1); k' < 112: mel + a1 columns; [112, 144): a3 (which 0: the P3 row only) or
//                  a2 (which 1: the P2 rows only); [144, 176): a4 (which 1: the P4 row only)
static void pack_wide(int C, const wrnn_weights *w, std::vector<float> &img)
{
    using namespace wrnn_wide;
    const int R = HID, F = 80, A = 32, KI = 1 + F + A, KC = F + A;
    std::vector<double> Wc((size_t)R * KC), w0(R), bI(R);
    for (int j = 0; j < R; ++j) {
        w0[j] = w->I_w[(size_t)j * KI];
        bI[j] = w->I_b[j];
        for (int k = 0; k < KC; ++k) Wc[(size_t)j * KC + k] = w->I_w[(size_t)j * KI + 1 + k];
    }
    std::vector<double> G1, G2, G3, u1, u2, u3, c1, c2, c3;
    gemm_f64(w->r1_wih, R, 3 * R, R, Wc, KC, G1);
    gemm_f64(w->r2_wih, R + A, 3 * R, R, Wc, KC, G2);
    gemm_f64(w->fc1_w, R + A, R, R, Wc, KC, G3);
    gemm_f64(w->r1_wih, R, 3 * R, R, w0, 1, u1);
    gemm_f64(w->r2_wih, R + A, 3 * R, R, w0, 1, u2);
    gemm_f64(w->fc1_w, R + A, R, R, w0, 1, u3);
    gemm_f64(w->r1_wih, R, 3 * R, R, bI, 1, c1);
    gemm_f64(w->r2_wih, R + A, 3 * R, R, bI, 1, c2);
    gemm_f64(w->fc1_w, R + A, R, R, bI, 1, c3);

    img.assign((size_t)NWORK * IMG_FLOATS, 0.f);
    for (int c = 0; c < NWORK; ++c) {
        float *base = &img[(size_t)c * IMG_FLOATS];
        const int j0 = UNITS * c;
        auto gate_matrix = [&](int off, const float *M, int ld) {
            for (int wp = 0; wp < 16; ++wp)
                for (int ig = 0; ig < 4; ++ig)
                    for (int ks = 0; ks < 2; ++ks)
                        for (int u = 0; u < UNITS; ++u)
                            for (int ii = 0; ii < 4; ++ii)
                                for (int g = 0; g < 3; ++g) {
                                    const int k = 32 * wp + 2 * (4 * ig + ii) + ks;
                                    base[off + ((((wp * 4 + ig) * 2 + ks) * 4 + u) * 12) + ii * 3 + g] = M[(size_t)(j0 + u + R * g) * ld + k];
                                }
        };
        gate_matrix(OFF_IH2, w->r2_wih, R + A);
        gate_matrix(OFF_HH1, w->r1_whh, R);
        gate_matrix(OFF_HH2, w->r2_whh, R);
        auto fc_matrix = [&](int off, const float *M, int ld, int rows_valid) {
            for (int k = 0; k < R; ++k)
                for (int u = 0; u < UNITS; ++u) base[off + k * 4 + u] = j0 + u < rows_valid ? M[(size_t)(j0 + u) * ld + k] : 0.f;
        };
        fc_matrix(OFF_FC1, w->fc1_w, R + A, R);
        fc_matrix(OFF_FC2, w->fc2_w, R + A, R);
        fc_matrix(OFF_FC3, w->fc3_w, R, C);              // class = 4 * cta + unit (MOL: 30 rows, CTAs 0..7)
        float *wc = base + OFF_WC;
        for (int kp = 0; kp < KC2; ++kp)
            for (int which = 0; which < 2; ++which)
                for (int u = 0; u < UNITS; ++u) {
                    float *dst = wc + (kp * 8 + which * 4 + u) * 4;
                    const int row = j0 + u;
                    for (int g = 0; g < 3; ++g) {
                        double v = 0.0;
                        if (kp < KC) v = which == 0 ? G1[(size_t)(row + R * g) * KC + kp] : G2[(size_t)(row + R * g) * KC + kp];
                        else if (which == 1 && kp < KC + A) v = w->r2_wih[(size_t)(row + R * g) * (R + A) + R + (kp - KC)];   // a2
                        dst[g] = (float)v;
                    }
                    double v = 0.0;
                    if (which == 0) {                    // P3: G3 | a3
                        if (kp < KC) v = G3[(size_t)row * KC + kp];
                        else if (kp < KC + A) v = w->fc1_w[(size_t)row * (R + A) + R + (kp - KC)];
                    } else if (kp >= KC + A) v = w->fc2_w[(size_t)row * (R + A) + R + (kp - KC - A)];       // P4: a4
                    dst[3] = (float)v;
                }
        float *sv = base + OFF_SV;
        for (int q = 0; q < 3; ++q)
            for (int u = 0; u < UNITS; ++u) {
                const int row = j0 + u + R * q;
                sv[SV_U1 + q * 4 + u] = (float)u1[row];
                sv[SV_B1 + q * 4 + u] = (float)(c1[row] + (double)w->r1_bih[row]);
                sv[SV_BHH1 + q * 4 + u] = w->r1_bhh[row];
                sv[SV_U2 + q * 4 + u] = (float)u2[row];
                sv[SV_B2 + q * 4 + u] = (float)(c2[row] + (double)w->r2_bih[row]);
                sv[SV_BHH2 + q * 4 + u] = w->r2_bhh[row];
            }
        for (int u = 0; u < UNITS; ++u) {
            sv[SV_U3 + u] = (float)u3[j0 + u];
            sv[SV_B3 + u] = (float)(c3[j0 + u] + (double)w->fc1_b[j0 + u]);
            sv[SV_B4 + u] = w->fc2_b[j0 + u];
            sv[SV_B5 + u] = j0 + u < C ? w->fc3_b[j0 + u] : 0.f;
        }
    }
}

// fp32 -> bf16, round to nearest even (finite inputs)
static unsigned short bf16_rne(float f)
{
    unsigned u;
    memcpy(&u, &f, 4);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (unsigned short)(u >> 16);
}

// Per-CTA images in the layout the kernel keeps resident.  bf16 precision: every item image is rounded to bf16
// (after the fp64 folding) and packed two per float slot; the small vectors stay fp32.
static void finish_images(int rows5, int bf16w, const std::vector<float> &img32, std::vector<float> &out)
{
    if (!bf16w) {
        out = img32;
        return;
    }
    const size_t per32 = (size_t)w_total(rows5), per = (size_t)w_image_floats(rows5, 1), items = (size_t)w_small(rows5);
    out.assign((size_t)NCTA * per, 0.f);
    for (int c = 0; c < NCTA; ++c) {
        const float *src = &img32[(size_t)c * per32];
        unsigned short *dst = reinterpret_cast<unsigned short *>(&out[(size_t)c * per]);
        for (size_t i = 0; i < items; ++i) dst[i] = bf16_rne(src[i]);
        memcpy(&out[(size_t)c * per + items / 2], src + items, SV_SIZE * sizeof(float));
    }
}

// ---- precision bf16-dense: operand stream, bundle table and per-row vectors of csrc/wavernn_dense.cuh ------------
// One step of the tensor-core program, in issue order (wavernn_dense.cuh has the dependency argument):
//   (a) after h1(t):  Wih2x.h1 -> g2 [commit G2] | Wfc1x.h1 -> f1 | P1.c(t+1) -> g1 (first touch)
//   (b) after h2(t):  Wfc1x.h2 -> f1 [commit F1] | P2.c(t+1) -> g2 (first touch) | Whh2.h2 -> g2 (1 of 6 bundles)
//   (c) after y1(t):  Wfc2x.y1 -> f2 [commit F2] | P3.c(t+1) -> f1 (first touch) | Whh2.h2 (2 of 6)
//   (d) after y2(t):  Wfc3.y2 -> f3 (first touch) [commit F3] | P4 -> f2 (first touch) | Whh1.h1 -> g1 [commit G1] | Whh2.h2 (3 of 6) [commit H2RD]
// The part after each "|" is work for step t+1 placed where the tensor pipe would otherwise wait for an epilogue + exchange.
// Tiles: T0 = [r | z] rows of the CTA's 64 units (128 rows), T1 = the n rows (64), F = 64 fc rows / classes.
namespace {
using wrnn_dense::Bundle;
using wrnn_dense::Seg;
typedef std::function<double(int rank, int m, int k)> DenseVal;
struct DenseSeg {
    int rows, nk, bsrc, dcol, first;
    DenseVal val;
    DenseSeg(int rows_, int nk_, int bsrc_, int dcol_, int first_, DenseVal val_) : rows(rows_), nk(nk_), bsrc(bsrc_), dcol(dcol_), first(first_), val(val_) {}
};
struct DenseBundle { int wait, commit; std::vector<DenseSeg> segs; };

struct DensePack {
    std::vector<Bundle> table;
    std::vector<unsigned char> stream;     // [CL][stream_bytes]
    std::vector<float> sv;                 // [CL][NSV][UPC]
    unsigned stream_bytes = 0;
};

static int32_t pack_dense(int C, const wrnn_weights *w, DensePack &out)
{
    using namespace wrnn_dense;
    const int R = DHID, F = 80, A = 32, KI = 1 + F + A, KC = F + A, RA = R + A;
    std::vector<double> Wc((size_t)R * KC), w0(R), bI(R);
    for (int j = 0; j < R; ++j) {
        w0[j] = w->I_w[(size_t)j * KI];
        bI[j] = w->I_b[j];
        for (int k = 0; k < KC; ++k) Wc[(size_t)j * KC + k] = w->I_w[(size_t)j * KI + 1 + k];
    }
    std::vector<double> G1, G2, G3, u1, u2, u3, c1, c2, c3;
    gemm_f64(w->r1_wih, R, 3 * R, R, Wc, KC, G1);
    gemm_f64(w->r2_wih, RA, 3 * R, R, Wc, KC, G2);
    gemm_f64(w->fc1_w, RA, R, R, Wc, KC, G3);
    gemm_f64(w->r1_wih, R, 3 * R, R, w0, 1, u1);
    gemm_f64(w->r2_wih, RA, 3 * R, R, w0, 1, u2);
    gemm_f64(w->fc1_w, RA, R, R, w0, 1, u3);
    gemm_f64(w->r1_wih, R, 3 * R, R, bI, 1, c1);
    gemm_f64(w->r2_wih, RA, 3 * R, R, bI, 1, c2);
    gemm_f64(w->fc1_w, RA, R, R, bI, 1, c3);

    // tile 0: accumulator lane m = 32 q + j holds r (j < 16) or z (j >= 16) of unit 16 q + j % 16, so that a unit's r shares its lane
    // with the unit's rows of the 64-row tiles (M = 64: row i -> lane 32 (i / 16) + i % 16) and its z sits 16 lanes above
    auto rowT0 = [](int rank, int m) { return ((m & 31) < 16 ? 0 : R) + UPC * rank + 16 * (m >> 5) + (m & 15); };
    auto rowN = [](int rank, int m) { return 2 * R + UPC * rank + m; };
    auto unit = [](int rank, int m) { return UPC * rank + m; };
    auto img = [](int i, int chunk) { return i * IMG_B + chunk * CHUNK_B; };
    auto cnd = [](int chunk) { return 4 * IMG_B + chunk * CHUNK_B; };

    std::vector<DenseBundle> prog;
    // K = 512 products: T0 in four 128-wide bundles, 64-row tiles in two 256-wide bundles, interleaved so both advance
    auto hidden = [&](int wait, int commit, int bimg, const float *W, int ld, int dT0, int dT1, int firstT0, int firstT1) {
        for (int half = 0; half < 2; ++half) {
            for (int kb = 2 * half; kb < 2 * half + 2; ++kb) {
                DenseBundle b;
                b.wait = kb == 0 ? wait : W_NONE;
                b.commit = C_NONE;
                b.segs.push_back(DenseSeg(128, 8, img(bimg, kb * 16), dT0, firstT0 && kb == 0, [=](int rank, int m, int k) { return (double)W[(size_t)rowT0(rank, m) * ld + kb * 128 + k]; }));
                prog.push_back(b);
            }
            DenseBundle b;
            b.wait = W_NONE;
            b.commit = half == 1 ? commit : C_NONE;
            b.segs.push_back(DenseSeg(UPC, 16, img(bimg, half * 32), dT1, firstT1 && half == 0, [=](int rank, int m, int k) { return (double)W[(size_t)rowN(rank, m) * ld + half * 256 + k]; }));
            prog.push_back(b);
        }
    };
    // single bundles of a K = 512 product, for work that is spread over several gaps
    auto t0_block = [&](int commit, int bimg, const float *W, int ld, int dT0, int kb, int first) {
        DenseBundle b;
        b.wait = W_NONE;
        b.commit = commit;
        b.segs.push_back(DenseSeg(128, 8, img(bimg, kb * 16), dT0, first, [=](int rank, int m, int k) { return (double)W[(size_t)rowT0(rank, m) * ld + kb * 128 + k]; }));
        prog.push_back(b);
    };
    auto t1_half = [&](int commit, int bimg, const float *W, int ld, int dT1, int half, int first) {
        DenseBundle b;
        b.wait = W_NONE;
        b.commit = commit;
        b.segs.push_back(DenseSeg(UPC, 16, img(bimg, half * 32), dT1, first, [=](int rank, int m, int k) { return (double)W[(size_t)rowN(rank, m) * ld + half * 256 + k]; }));
        prog.push_back(b);
    };
    auto fc = [&](int wait, int commit, int bimg, const float *W, int ld, int dcol, int first, bool classes) {
        for (int half = 0; half < 2; ++half) {
            DenseBundle b;
            b.wait = half == 0 ? wait : W_NONE;
            b.commit = half == 1 ? commit : C_NONE;
            // fc3: output row 64 rank + m exists only below C (MOL: the 30 outputs are rows 0-29 of rank 0, everything else is zero)
            b.segs.push_back(DenseSeg(UPC, 16, img(bimg, half * 32), dcol, first && half == 0, [=](int rank, int m, int k) {
                return classes && unit(rank, m) >= C ? 0.0 : (double)W[(size_t)unit(rank, m) * ld + half * 256 + k]; }));
            prog.push_back(b);
        }
    };
    const double *g1 = G1.data(), *g2 = G2.data(), *g3 = G3.data();
    const float *r2 = w->r2_wih, *f1 = w->fc1_w, *f2 = w->fc2_w;
    auto p2 = [=](int row, int k) { return k < KC ? g2[(size_t)row * KC + k] : (double)r2[(size_t)row * RA + R + (k - KC)]; };
    auto one = [&](int wait, int commit, std::vector<DenseSeg> segs) {
        DenseBundle b;
        b.wait = wait;
        b.commit = commit;
        b.segs = segs;
        prog.push_back(b);
    };
    // conditioning inputs k: [0,80) mel | [80,112) a1 | [112,144) a2 | [144,176) a3 | [176,208) a4 (8 per image chunk)
    // (a) critical: Wih2x.h1(t) -> g2.  Then, inside E2 + the h2 exchange: Wfc1x.h1 -> f1, P1.c(t+1) -> g1 (first touch)
    hidden(W_H1, C_G2, IMG_H1, w->r2_wih, RA, D_G2_T0, D_G2_1I, 0, 0);
    fc(W_NONE, C_NONE, IMG_H1, w->fc1_w, RA, D_F1, 0, false);
    one(W_COND, C_NONE, {DenseSeg(128, 7, cnd(0), D_G1_T0, 1, [=](int rank, int m, int k) { return g1[(size_t)rowT0(rank, m) * KC + k]; })});
    one(W_NONE, C_NONE, {DenseSeg(UPC, 7, cnd(0), D_G1_1I, 1, [=](int rank, int m, int k) { return g1[(size_t)rowN(rank, m) * KC + k]; })});
    t1_half(C_NONE, IMG_H1, w->r1_whh, R, D_G1_1H, 0, 1);       // one sixth of Whh1.h1(t) here, the rest in the tail of (d)
    // (b) critical: Wfc1x.h2(t) -> f1.  Inside E3 + the y1 exchange: P2.c(t+1) -> g2 (first touch; E2(t) has drained g2)
    fc(W_H2, C_F1, IMG_H2, w->fc1_w, RA, D_F1, 0, false);
    one(W_NONE, C_NONE, {DenseSeg(128, 8, cnd(0), D_G2_T0, 1, [=](int rank, int m, int k) { return p2(rowT0(rank, m), k); })});
    one(W_NONE, C_NONE, {DenseSeg(128, 1, cnd(16), D_G2_T0, 0, [=](int rank, int m, int k) { return p2(rowT0(rank, m), 128 + k); }),
                         DenseSeg(UPC, 9, cnd(0), D_G2_1I, 1, [=](int rank, int m, int k) { return p2(rowN(rank, m), k); })});
    // Whh2.h2(t) (step t+1's hidden-side gates) is spread over the gaps of (b), (c) and the tail of (d): the n rows' first half here
    t1_half(C_NONE, IMG_H2, w->r2_whh, R, D_G2_1H, 0, 1);
    // (c) critical: Wfc2x.y1 -> f2.  Inside E4 + the y2 exchange: P3.c(t+1) -> f1 (first touch; E3(t) has drained f1), Whh2 [r | z] K 0..255
    fc(W_Y1, C_F2, IMG_Y1, w->fc2_w, RA, D_F2, 0, false);
    one(W_NONE, C_NONE, {DenseSeg(UPC, 7, cnd(0), D_F1, 1, [=](int rank, int m, int k) { return g3[(size_t)unit(rank, m) * KC + k]; }),
                         DenseSeg(UPC, 2, cnd(18), D_F1, 0, [=](int rank, int m, int k) { return (double)f1[(size_t)unit(rank, m) * RA + R + k]; })});
    t0_block(C_NONE, IMG_H2, w->r2_whh, R, D_G2_T0, 0, 0);
    t0_block(C_NONE, IMG_H2, w->r2_whh, R, D_G2_T0, 1, 0);
    // (d) critical: Wfc3.y2 -> f3.  Inside E5, the logits exchange, sampling and E1(t+1): P4 -> f2 (first touch), Whh1.h1(t) -> g1
    // [commit G1: E1(t+1) may start], then Whh2.h2(t) -> g2 [commit H2RD: this CTA no longer reads the h2(t) image]
    fc(W_Y2, C_F3, IMG_Y2, w->fc3_w, R, D_F3, 1, true);
    one(W_NONE, C_NONE, {DenseSeg(UPC, 2, cnd(22), D_F2, 1, [=](int rank, int m, int k) { return (double)f2[(size_t)unit(rank, m) * RA + R + k]; })});
    for (int kb = 0; kb < 4; ++kb) t0_block(C_NONE, IMG_H1, w->r1_whh, R, D_G1_T0, kb, 0);
    t1_half(C_G1, IMG_H1, w->r1_whh, R, D_G1_1H, 1, 0);
    t0_block(C_NONE, IMG_H2, w->r2_whh, R, D_G2_T0, 2, 0);
    t0_block(C_NONE, IMG_H2, w->r2_whh, R, D_G2_T0, 3, 0);
    t1_half(C_H2RD, IMG_H2, w->r2_whh, R, D_G2_1H, 1, 0);
    if ((int)prog.size() > MAXBUNDLE) return fail(WRNN_ERR_INVALID, "dense program has %zu bundles (max %d)", prog.size(), MAXBUNDLE);

    // serialise: table (identical for every rank) and the per-rank streams
    out.table.assign(prog.size(), Bundle{});
    unsigned off = 0;
    for (size_t b = 0; b < prog.size(); ++b) {
        Bundle &t = out.table[b];
        t.src_off = off;
        t.nseg = (uint16_t)prog[b].segs.size();
        t.wait = (uint16_t)prog[b].wait;
        t.commit = (uint16_t)prog[b].commit;
        if (t.nseg > MAXSEG) return fail(WRNN_ERR_INVALID, "bundle %zu has %d segments", b, (int)t.nseg);
        unsigned bytes = 0;
        for (int s = 0; s < t.nseg; ++s) {
            const DenseSeg &d = prog[b].segs[s];
            t.seg[s] = Seg{(uint16_t)(bytes / 16), (uint16_t)d.rows, (uint16_t)d.nk, (uint16_t)(d.bsrc / 16), (uint16_t)d.dcol, (uint16_t)d.first};
            bytes += (unsigned)d.rows * d.nk * 32;
        }
        if (bytes > (unsigned)SLOT) return fail(WRNN_ERR_INVALID, "bundle %zu is %u bytes (slot %d)", b, bytes, SLOT);
        t.bytes = bytes;
        off += bytes;
    }
    out.stream_bytes = off;
    out.stream.assign((size_t)CL * off, 0);
    for (int rank = 0; rank < CL; ++rank) {
        unsigned short *dst = reinterpret_cast<unsigned short *>(out.stream.data() + (size_t)rank * off);
        size_t e = 0;
        for (size_t b = 0; b < prog.size(); ++b)
            for (const DenseSeg &d : prog[b].segs)
                for (int kc = 0; kc < 2 * d.nk; ++kc)             // image: [k chunk][row][8 k]
                    for (int m = 0; m < d.rows; ++m)
                        for (int i = 0; i < 8; ++i) dst[e++] = bf16_rne((float)d.val(rank, m, kc * 8 + i));
    }
    out.sv.assign((size_t)CL * NSV * UPC, 0.f);
    for (int rank = 0; rank < CL; ++rank) {
        float *sv = &out.sv[(size_t)rank * NSV * UPC];
        for (int m = 0; m < UPC; ++m) {
            const int ur = UPC * rank + m, rr = ur, rz = R + ur, rn = 2 * R + ur;
            sv[DV_B1R * UPC + m] = (float)(c1[rr] + (double)w->r1_bih[rr] + (double)w->r1_bhh[rr]);
            sv[DV_U1R * UPC + m] = (float)u1[rr];
            sv[DV_B1Z * UPC + m] = (float)(c1[rz] + (double)w->r1_bih[rz] + (double)w->r1_bhh[rz]);
            sv[DV_U1Z * UPC + m] = (float)u1[rz];
            sv[DV_B1NI * UPC + m] = (float)(c1[rn] + (double)w->r1_bih[rn]);
            sv[DV_U1N * UPC + m] = (float)u1[rn];
            sv[DV_B1NH * UPC + m] = w->r1_bhh[rn];
            sv[DV_B2R * UPC + m] = (float)(c2[rr] + (double)w->r2_bih[rr] + (double)w->r2_bhh[rr]);
            sv[DV_U2R * UPC + m] = (float)u2[rr];
            sv[DV_B2Z * UPC + m] = (float)(c2[rz] + (double)w->r2_bih[rz] + (double)w->r2_bhh[rz]);
            sv[DV_U2Z * UPC + m] = (float)u2[rz];
            sv[DV_B2NI * UPC + m] = (float)(c2[rn] + (double)w->r2_bih[rn]);
            sv[DV_U2N * UPC + m] = (float)u2[rn];
            sv[DV_B2NH * UPC + m] = w->r2_bhh[rn];
            sv[DV_B3 * UPC + m] = (float)(c3[ur] + (double)w->fc1_b[ur]);
            sv[DV_U3 * UPC + m] = (float)u3[ur];
            sv[DV_B4 * UPC + m] = w->fc2_b[ur];
            sv[DV_B5 * UPC + m] = ur < C ? w->fc3_b[ur] : 0.f;
        }
    }
    return WRNN_OK;
}
}   // namespace

extern "C" int32_t wrnn_dense_layout(const wrnn_config *cfg, int64_t *layout)
{
    if (!cfg || !layout) return fail(WRNN_ERR_INVALID, "null argument");
    int rows5, nprod5, n_u;
    wrnn_config c = *cfg;
    c.precision = WRNN_PREC_BF16_DENSE;
    int32_t rc = derive_layout(c, rows5, nprod5, n_u);
    if (rc) return rc;
    // the program does not depend on the weight values: pack zeros to measure it
    std::vector<float> z((size_t)1536 * 544, 0.f);
    wrnn_weights w = {z.data(), z.data(), z.data(), z.data(), z.data(), z.data(), z.data(), z.data(), z.data(), z.data(), z.data(), z.data(), z.data(), z.data(), z.data(), z.data()};
    DensePack dp;
    rc = pack_dense(c.n_classes, &w, dp);
    if (rc) return rc;
    const int64_t v[8] = {(int64_t)dp.table.size(), dp.stream_bytes, (int64_t)sizeof(Bundle), wrnn_dense::CL, wrnn_dense::UPC, wrnn_dense::BC, wrnn_dense::NSV, 0};
    memcpy(layout, v, sizeof v);
    return WRNN_OK;
}

extern "C" int32_t wrnn_dense_pack_host(const wrnn_config *cfg, const wrnn_weights *w, uint8_t *stream, uint8_t *table, float *sv)
{
    if (!cfg || !w || !stream || !table || !sv) return fail(WRNN_ERR_INVALID, "null argument");
    int rows5, nprod5, n_u;
    wrnn_config c = *cfg;
    c.precision = WRNN_PREC_BF16_DENSE;
    int32_t rc = derive_layout(c, rows5, nprod5, n_u);
    if (rc) return rc;
    DensePack dp;
    rc = pack_dense(c.n_classes, w, dp);
    if (rc) return rc;
    memcpy(stream, dp.stream.data(), dp.stream.size());
    memcpy(table, dp.table.data(), dp.table.size() * sizeof(Bundle));
    memcpy(sv, dp.sv.data(), dp.sv.size() * sizeof(float));
    return WRNN_OK;
}

extern "C" int64_t wrnn_packed_floats(const wrnn_config *cfg)
{
    int rows5, nprod5, n_u;
    if (!cfg || derive_layout(*cfg, rows5, nprod5, n_u)) return -1;
    return (int64_t)NCTA * w_image_floats(rows5, cfg->precision == WRNN_PREC_BF16 ? 1 : 0);
}

extern "C" int32_t wrnn_pack_weights_host(const wrnn_config *cfg, const wrnn_weights *w, float *out, int64_t out_floats)
{
    if (!cfg || !w || !out) return fail(WRNN_ERR_INVALID, "null argument");
    int rows5, nprod5, n_u;
    int32_t rc = derive_layout(*cfg, rows5, nprod5, n_u);
    if (rc) return rc;
    const int bf16w = cfg->precision == WRNN_PREC_BF16 ? 1 : 0;
    const long long want = (long long)NCTA * w_image_floats(rows5, bf16w);
    if (out_floats != want) return fail(WRNN_ERR_INVALID, "out_floats must be %lld", want);
    std::vector<float> img32, img;
    pack_images(cfg->n_classes, rows5, w, img32);
    finish_images(rows5, bf16w, img32, img);
    memcpy(out, img.data(), img.size() * sizeof(float));
    return WRNN_OK;
}

extern "C" int64_t wrnn_wide_packed_floats(const wrnn_config *cfg, int64_t *layout)
{
    int rows5, nprod5, n_u;
    if (!cfg || derive_layout(*cfg, rows5, nprod5, n_u) || !wide_supported(*cfg)) return -1;
    using namespace wrnn_wide;
    if (layout) {
        const int64_t v[8] = {IMG_FLOATS, OFF_IH2, OFF_HH1, OFF_HH2, OFF_FC1, OFF_FC2, OFF_FC3, OFF_WC};
        memcpy(layout, v, sizeof v);
    }
    return (int64_t)NWORK * IMG_FLOATS;
}

extern "C" int32_t wrnn_wide_pack_host(const wrnn_config *cfg, const wrnn_weights *w, float *out, int64_t out_floats)
{
    if (!cfg || !w || !out) return fail(WRNN_ERR_INVALID, "null argument");
    const int64_t want = wrnn_wide_packed_floats(cfg, nullptr);
    if (want < 0) return fail(WRNN_ERR_INVALID, "the wide kernel serves fp32 RAW-512 and MOL models only");
    if (out_floats != want) return fail(WRNN_ERR_INVALID, "out_floats must be %lld", (long long)want);
    std::vector<float> img;
    pack_wide(cfg->n_classes, w, img);
    memcpy(out, img.data(), img.size() * sizeof(float));
    return WRNN_OK;
}

extern "C" int32_t wrnn_load_weights(wrnn_handle *h, const wrnn_weights *w)
{
    if (!h || !w) return fail(WRNN_ERR_INVALID, "null argument");
    const float *const *pp = reinterpret_cast<const float *const *>(w);
    for (int i = 0; i < 16; ++i)
        if (!pp[i]) return fail(WRNN_ERR_INVALID, "weight pointer %d is null", i);
    DeviceGuard guard_(h->device);
    if (h->dense) {
        DensePack dp;
        int32_t rc = pack_dense(h->cfg.n_classes, w, dp);
        if (rc) return rc;
        cudaFree(h->dense_stream);
        cudaFree(h->dense_table);
        cudaFree(h->dense_sv);
        h->dense_stream = nullptr;
        h->dense_table = nullptr;
        h->dense_sv = nullptr;
        CUDA_TRY(cudaMalloc(&h->dense_stream, dp.stream.size()));
        CUDA_TRY(cudaMalloc(&h->dense_table, dp.table.size() * sizeof(wrnn_dense::Bundle)));
        CUDA_TRY(cudaMalloc(&h->dense_sv, dp.sv.size() * sizeof(float)));
        CUDA_TRY(cudaMemcpy(h->dense_stream, dp.stream.data(), dp.stream.size(), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(h->dense_table, dp.table.data(), dp.table.size() * sizeof(wrnn_dense::Bundle), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(h->dense_sv, dp.sv.data(), dp.sv.size() * sizeof(float), cudaMemcpyHostToDevice));
        h->dense_nb = (int)dp.table.size();
        h->dense_stream_bytes = dp.stream_bytes;
        h->loaded = true;
        return WRNN_OK;
    }
    std::vector<float> img32, img;
    pack_images(h->cfg.n_classes, h->rows5, w, img32);
    finish_images(h->rows5, h->bf16w, img32, img);
    CUDA_TRY(cudaMemcpy(h->wimg, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice));
    if (h->wide) {
        std::vector<float> wimg;
        pack_wide(h->cfg.n_classes, w, wimg);
        CUDA_TRY(cudaMemcpy(h->wide_img, wimg.data(), wimg.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    h->loaded = true;
    return WRNN_OK;
}

// ---- step loop launch --------------------------------------------------------------------------
// A generate call is ENQUEUED: begin_call / end_call bracket its launches with two events and a D2H of the kernels' status
// word into pinned memory; finish_pending (wrnn_synchronize, wrnn_get_info, the next generate call, wrnn_destroy) waits for
// it and turns a fired watchdog into WRNN_ERR_TIMEOUT.
static int32_t finish_pending(wrnn_handle *h)
{
    if (!h->pending) return WRNN_OK;
    h->pending = false;
    CUDA_TRY(cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    h->last_status = h->status_host[0];
    if (h->last_status != 0)
        return fail(WRNN_ERR_TIMEOUT, "step-loop kernel watchdog fired (status %d): an exchange or pipeline barrier never completed", h->last_status);
    return WRNN_OK;
}
static int32_t begin_call(wrnn_handle *h, cudaStream_t st)
{
    int32_t rc = finish_pending(h);          // a handle is not re-entrant: one call in flight
    if (rc) return rc;
    *h->progress_host = 0;
    CUDA_TRY(cudaMemsetAsync(h->status, 0, 4 * sizeof(int), st));
    CUDA_TRY(cudaEventRecord(h->ev0, st));
    return WRNN_OK;
}
static int32_t end_call(wrnn_handle *h, cudaStream_t st)
{
    CUDA_TRY(cudaMemcpyAsync(h->status_host, h->status, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(h->ev1, st));
    h->pending = true;
    return WRNN_OK;
}

// Teams per CTA and conditioning buffers per team: as many teams as there are groups (at most MAXT)
// and double-buffered staging when it fits the opt-in shared memory, else fewer.
static void choose_teams(const wrnn_handle *h, int G, int &T, int &nbuf)
{
    for (T = G < MAXT ? G : MAXT; T >= 1; --T)
        for (nbuf = 2; nbuf >= 1; --nbuf)
            if (smem_map(h->rows5, h->cfg.mode, h->cfg.n_classes, T, nbuf, h->bf16w).total * (int)sizeof(float) <= h->smem_limit) return;
    T = 1;
    nbuf = 1;
}

static int32_t launch_chunk(wrnn_handle *h, KParams &p, cudaStream_t st, bool probe)
{
    void *args[] = {&p};
    if (probe) {
        p.T = 1;
        p.nbuf = 1;
    } else {
        const char *force = getenv("WRNN_FORCE_TEAMS");          // development knob: cap the number of teams
        choose_teams(h, p.G, p.T, p.nbuf);
        if (force && atoi(force) >= 1 && atoi(force) < p.T) {
            p.T = atoi(force);
            p.nbuf = 2;
            if (smem_map(h->rows5, h->cfg.mode, h->cfg.n_classes, p.T, 2, h->bf16w).total * (int)sizeof(float) > h->smem_limit) p.nbuf = 1;
        }
    }
    h->smem_bytes = smem_map(h->rows5, h->cfg.mode, h->cfg.n_classes, p.T, p.nbuf, h->bf16w).total * (int)sizeof(float);
    h->last_teams = p.T;
    h->last_kernel = 0;
    // epochs restart at 1 every launch: clear stale {value, epoch} pairs of the previous one
    CUDA_TRY(cudaMemsetAsync(h->xb, 0, (size_t)MAXG * xb_group(h->cpad) * sizeof(unsigned long long), st));
    const void *fn = probe ? (const void *)wavernn_exchange_probe_kernel
                           : persistent_kernel(p.T, model_of(h->cfg), h->bf16w, p.prof != nullptr);
    CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(NCTA), dim3(NTHREADS), args, (size_t)h->smem_bytes, st));
    h->launches += 1;
    return WRNN_OK;
}

static void fill_common(wrnn_handle *h, KParams &p)
{
    memset(&p, 0, sizeof p);
    p.wimg = h->wimg;
    p.xb = h->xb;
    p.status = h->status;
    p.C = h->cfg.n_classes;
    p.mode = h->cfg.mode;
    p.rows5 = h->rows5;
    p.nprod5 = h->nprod5;
    p.n_u = h->n_u;
    p.bf16w = h->bf16w;
    p.feat = h->cfg.feat_dims;
    p.auxw = 4 * h->cfg.aux_dims;
}

// wide kernel (csrc/wavernn_wide.cuh): one cooperative launch of 128 workers + sampler CTAs per <= 21 folds
static int32_t launch_wide(wrnn_handle *h, wrnn_wide::WParams &p, cudaStream_t st, bool probe)
{
    using namespace wrnn_wide;
    p.wimg = h->wide_img;
    p.xb = h->wide_xb;
    p.status = h->status;
    p.C = h->cfg.n_classes;
    p.mode = h->cfg.mode;
    p.n_u = h->n_u;
    p.feat = h->cfg.feat_dims;
    p.auxw = 4 * h->cfg.aux_dims;
    p.nq = (p.F + 2) / 3;
    p.nsamp = h->wide_nsamp;
    CUDA_TRY(cudaMemsetAsync(h->wide_xb, 0, (size_t)XW_TOTAL * sizeof(unsigned), st));
    const bool mol = h->cfg.mode == WRNN_MODE_MOL;
    const void *fn = probe ? (const void *)wavernn_wide_probe_kernel
                   : p.prof ? (mol ? (const void *)wavernn_wide_kernel_mol_prof : (const void *)wavernn_wide_kernel_prof)
                            : (mol ? (const void *)wavernn_wide_kernel_mol : (const void *)wavernn_wide_kernel);
    void *args[] = {&p};
    CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(NWORK + h->wide_nsamp), dim3(probe ? NTHREADS : WTHREADS), args, (size_t)SM_BYTES, st));
    h->smem_bytes = SM_BYTES;
    h->last_kernel = 1;
    h->launches += 1;
    return WRNN_OK;
}
// Which fp32 kernel serves a call: the wide kernel whenever the model is one it is built for (fp32 RAW-512 / MOL).  Since its
// publishes are whole 32-byte sectors it is at least as fast as the grouped kernel at every fold count (1 fold: 10.7 vs 11.2 us per
// step, 8 folds: 11.07 vs 11.06, 20 folds: 12.0 vs 24.0; profiles/r02_summary.md).  The grouped kernel keeps bf16 resident weights and other
// class counts.  wrnn_set_kernel / WRNN_KERNEL=grouped | wide force one (the two add partial sums in different orders).
static bool use_wide(const wrnn_handle *h, int num_folds)
{
    if (!h->wide) return false;
    if (h->kernel_choice == 0) return false;
    if (h->kernel_choice == 1) return true;
    const char *k = getenv("WRNN_KERNEL");
    if (k && strcmp(k, "grouped") == 0) return false;
    (void)num_folds;
    return true;
}

// dense step loop (csrc/wavernn_dense.cuh): one launch holds every fold; clusters are independent
static int32_t launch_dense(wrnn_handle *h, wrnn_dense::DParams dp, bool frames, int32_t num_folds, int32_t steps, cudaStream_t st)
{
    using namespace wrnn_dense;
    // clusters are independent (no grid-level synchronisation): one launch holds every fold, the hardware runs the
    // clusters in waves of h->dense_clusters
    const int cap = 2048 * BC;
    for (int b0 = 0; b0 < num_folds; b0 += cap) {
        const int nb = num_folds - b0 < cap ? num_folds - b0 : cap;
        // whole waves of co-resident clusters, folds dealt evenly: a cluster's step time hardly depends on its fold count
        int ncl = (nb + BC - 1) / BC;
        if (ncl > h->dense_clusters) ncl = (ncl + h->dense_clusters - 1) / h->dense_clusters * h->dense_clusters;
        if (ncl > nb) ncl = nb;
        dp.wstream = h->dense_stream;
        dp.table = h->dense_table;
        dp.sv = h->dense_sv;
        dp.status = h->status;
        dp.stream_bytes = h->dense_stream_bytes;
        dp.nb = h->dense_nb;
        dp.B = num_folds;
        dp.S = steps;
        dp.fold0 = b0;
        dp.nfolds = nb;
        dp.per = (nb + ncl - 1) / ncl;
        dp.prof = nullptr;
        dp.mol = h->cfg.mode == WRNN_MODE_MOL ? 1 : 0;
        if (h->profiling && !frames) {
            // the counters are per CTA of the launch (the grid can exceed one wave of clusters)
            const size_t need = (size_t)ncl * CL * PROF_N;
            if (need > h->dense_prof_slots) {
                cudaFree(h->dense_prof);
                h->dense_prof = nullptr;
                h->dense_prof_slots = 0;
                CUDA_TRY(cudaMalloc(&h->dense_prof, need * sizeof(long long)));
                h->dense_prof_slots = need;
            }
            CUDA_TRY(cudaMemsetAsync(h->dense_prof, 0, h->dense_prof_slots * sizeof(long long), st));
            dp.prof = h->dense_prof;
        }
        if (frames) wavernn_dense_kernel_frames<<<dim3(ncl * CL), dim3(DTHREADS), SM_TOTAL, st>>>(dp);
        else if (dp.prof) wavernn_dense_kernel_prof<<<dim3(ncl * CL), dim3(DTHREADS), SM_TOTAL, st>>>(dp);
        else wavernn_dense_kernel<<<dim3(ncl * CL), dim3(DTHREADS), SM_TOTAL, st>>>(dp);
        CUDA_TRY(cudaGetLastError());
        h->launches += 1;
        h->smem_bytes = SM_TOTAL;
        h->last_kernel = 2;
    }
    return WRNN_OK;
}

extern "C" int32_t wrnn_generate_folds(wrnn_handle *h, const float *mels, const float *aux, int64_t cond_rows,
                                       const int64_t *fold_start, const int64_t *fold_limit,
                                       int32_t num_folds, int32_t steps,
                                       const float *uniforms, uint64_t seed,
                                       const float *forced_x, float *logits_out,
                                       float *samples_out, int32_t *labels_out, void *stream)
{
    if (!h) return fail(WRNN_ERR_INVALID, "null handle");
    if (!h->loaded) return fail(WRNN_ERR_STATE, "wrnn_load_weights has not been called");
    if (!mels || !aux || !fold_start || !fold_limit || !samples_out) return fail(WRNN_ERR_INVALID, "null pointer argument");
    if (num_folds <= 0 || steps <= 0) return fail(WRNN_ERR_INVALID, "num_folds (%d) and steps (%d) must be positive", num_folds, steps);
    if (((uintptr_t)mels | (uintptr_t)aux) & 15) return fail(WRNN_ERR_INVALID, "conditioning pointers must be 16-byte aligned");
    for (int b = 0; b < num_folds; ++b) {
        if (fold_start[b] < 0 || fold_limit[b] > cond_rows || fold_limit[b] < 0)
            return fail(WRNN_ERR_INVALID, "fold %d: start %lld / limit %lld outside [0, %lld]", b, (long long)fold_start[b], (long long)fold_limit[b], (long long)cond_rows);
    }
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard guard_(h->device);
    int32_t rc = begin_call(h, st);
    if (rc) return rc;
    if (num_folds > h->fold_cap) {
        cudaFree(h->fold_dev);
        h->fold_dev = nullptr;
        h->fold_cap = 0;
        CUDA_TRY(cudaMalloc(&h->fold_dev, (size_t)2 * num_folds * sizeof(long long)));
        h->fold_cap = num_folds;
    }
    CUDA_TRY(cudaMemcpyAsync(h->fold_dev, fold_start, (size_t)num_folds * sizeof(long long), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->fold_dev + h->fold_cap, fold_limit, (size_t)num_folds * sizeof(long long), cudaMemcpyHostToDevice, st));

    if (h->dense) {
        if (cond_rows > 0x7fffffffll) return fail(WRNN_ERR_INVALID, "dense path indexes conditioning rows with 32 bits (got %lld rows)", (long long)cond_rows);
        wrnn_dense::DParams dp;
        memset(&dp, 0, sizeof dp);
        dp.mels = mels;
        dp.aux = aux;
        dp.fold_start = h->fold_dev;
        dp.fold_limit = h->fold_dev + h->fold_cap;
        dp.uniforms = uniforms;
        dp.forced_x = forced_x;
        dp.logits_out = logits_out;
        dp.samples_out = samples_out;
        dp.labels_out = labels_out;
        dp.seed = seed;
        rc = launch_dense(h, dp, false, num_folds, steps, st);
        if (rc) return rc;
        return end_call(h, st);
    }
    if (use_wide(h, num_folds)) {
        // balanced launches of at most FMAX folds; a fold's arithmetic does not depend on its launch mates
        const int nl = (num_folds + wrnn_wide::FMAX - 1) / wrnn_wide::FMAX;
        int next = 0;
        for (int l = 0; l < nl; ++l) {
            const int nf = num_folds / nl + (l < num_folds % nl ? 1 : 0);
            wrnn_wide::WParams p;
            memset(&p, 0, sizeof p);
            p.mels = mels;
            p.aux = aux;
            p.fold_start = h->fold_dev;
            p.fold_limit = h->fold_dev + h->fold_cap;
            p.uniforms = uniforms;
            p.forced_x = forced_x;
            p.logits_out = logits_out;
            p.samples_out = samples_out;
            p.labels_out = labels_out;
            p.seed = seed;
            p.prof = h->profiling ? h->prof_dev : nullptr;
            p.progress = nl == 1 ? h->progress_dev : nullptr;
            p.B = num_folds;
            p.S = steps;
            p.F = nf;
            p.fold0 = next;
            next += nf;
            rc = launch_wide(h, p, st, false);
            if (rc) return rc;
        }
        return end_call(h, st);
    }
    const int max_chunk = MAXG * BT;
    for (int b0 = 0; b0 < num_folds; b0 += max_chunk) {
        const int nb = num_folds - b0 < max_chunk ? num_folds - b0 : max_chunk;
        KParams p;
        fill_common(h, p);
        p.mels = mels;
        p.aux = aux;
        p.fold_start = h->fold_dev;
        p.fold_limit = h->fold_dev + h->fold_cap;
        p.uniforms = uniforms;
        p.forced_x = forced_x;
        p.logits_out = logits_out;
        p.samples_out = samples_out;
        p.labels_out = labels_out;
        p.seed = seed;
        p.prof = h->profiling ? h->prof_dev : nullptr;
        p.B = num_folds;
        p.S = steps;
        p.G = (nb + BT - 1) / BT;
        int next = b0;
        for (int g = 0; g < p.G; ++g) {           // balanced groups: sizes differ by at most one
            const int nf = nb / p.G + (g < nb % p.G ? 1 : 0);
            p.group_fold0[g] = next;
            p.group_nf[g] = nf;
            next += nf;
        }
        rc = launch_chunk(h, p, st, false);
        if (rc) return rc;
    }
    return end_call(h, st);
}

extern "C" int32_t wrnn_generate_folds_frames(wrnn_handle *h, const float *mel_frames, int64_t frame_rows, const float *aux_frames, int64_t aux_rows,
                                              const float *interp, int32_t hop, int32_t pad, const int32_t *fold_geo, int32_t num_folds, int32_t steps,
                                              const float *uniforms, uint64_t seed, const float *forced_x, float *logits_out,
                                              float *samples_out, int32_t *labels_out, void *stream)
{
    if (!h) return fail(WRNN_ERR_INVALID, "null handle");
    if (!h->dense) return fail(WRNN_ERR_INVALID, "in-kernel conditioning expansion is implemented by the dense kernel (precision WRNN_PREC_BF16_DENSE)");
    if (!h->loaded) return fail(WRNN_ERR_STATE, "wrnn_load_weights has not been called");
    if (!mel_frames || !aux_frames || !interp || !fold_geo || !samples_out) return fail(WRNN_ERR_INVALID, "null pointer argument");
    if (num_folds <= 0 || steps <= 0 || hop <= 0 || pad < 2) return fail(WRNN_ERR_INVALID, "num_folds (%d), steps (%d), hop (%d) must be positive and pad (%d) >= 2", num_folds, steps, hop, pad);
    if (((uintptr_t)mel_frames | (uintptr_t)aux_frames) & 15) return fail(WRNN_ERR_INVALID, "frame pointers must be 16-byte aligned");
    if (frame_rows > 0x7fffffffll / 128 || aux_rows > 0x7fffffffll / 128) return fail(WRNN_ERR_INVALID, "too many frame rows");
    for (int b = 0; b < num_folds; ++b) {
        const int32_t *g = fold_geo + 4 * b;
        // frames touched by the last valid sample: padded mel frame (len - 1 + pad*hop) / hop + 2, aux frame (len - 1) / hop
        if (g[0] < 0 || g[1] <= 0 || g[2] < 0 || g[3] < 0 || g[2] + (g[1] - 1 + pad * hop) / hop + 2 >= frame_rows || g[3] + (g[1] - 1) / hop >= aux_rows)
            return fail(WRNN_ERR_INVALID, "fold %d: geometry {%d, %d, %d, %d} outside the frame tensors (%lld, %lld rows)", b, g[0], g[1], g[2], g[3],
                        (long long)frame_rows, (long long)aux_rows);
    }
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard guard_(h->device);
    int32_t rc = begin_call(h, st);
    if (rc) return rc;
    if (num_folds > h->fold_cap) {
        cudaFree(h->fold_dev);
        h->fold_dev = nullptr;
        h->fold_cap = 0;
        CUDA_TRY(cudaMalloc(&h->fold_dev, (size_t)2 * num_folds * sizeof(long long)));
        h->fold_cap = num_folds;
    }
    CUDA_TRY(cudaMemcpyAsync(h->fold_dev, fold_geo, (size_t)num_folds * 4 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    wrnn_dense::DParams dp;
    memset(&dp, 0, sizeof dp);
    dp.mel_frames = mel_frames;
    dp.aux_frames = aux_frames;
    dp.interp = interp;
    dp.fold_geo = reinterpret_cast<const int *>(h->fold_dev);
    dp.hop = hop;
    dp.indent = pad * hop;
    dp.uniforms = uniforms;
    dp.forced_x = forced_x;
    dp.logits_out = logits_out;
    dp.samples_out = samples_out;
    dp.labels_out = labels_out;
    dp.seed = seed;
    rc = launch_dense(h, dp, true, num_folds, steps, st);
    if (rc) return rc;
    return end_call(h, st);
}

extern "C" int32_t wrnn_synchronize(wrnn_handle *h)
{
    if (!h) return fail(WRNN_ERR_INVALID, "null handle");
    DeviceGuard guard_(h->device);
    return finish_pending(h);
}

extern "C" int32_t wrnn_query(wrnn_handle *h, int32_t *done, int32_t *steps_done)
{
    if (!h || !done) return fail(WRNN_ERR_INVALID, "null argument");
    DeviceGuard guard_(h->device);
    *done = 1;
    if (h->pending) {
        const cudaError_t e = cudaEventQuery(h->ev1);
        if (e == cudaErrorNotReady) *done = 0;
        else if (e != cudaSuccess) return fail(WRNN_ERR_CUDA, "cudaEventQuery: %s", cudaGetErrorString(e));
    }
    if (steps_done) *steps_done = *reinterpret_cast<volatile int *>(h->progress_host);
    return WRNN_OK;
}

extern "C" int32_t wrnn_measure_exchange(wrnn_handle *h, int32_t iters, float *usec_per_exchange)
{
    if (!h || !usec_per_exchange || iters <= 0) return fail(WRNN_ERR_INVALID, "bad argument");
    DeviceGuard guard_(h->device);
    for (int rep = 0; rep < 2; ++rep) {          // warm-up, then the measured launch
        int32_t rc = begin_call(h, nullptr);
        if (rc) return rc;
        if (use_wide(h, wrnn_wide::FMAX)) {
            wrnn_wide::WParams p;
            memset(&p, 0, sizeof p);
            p.F = wrnn_wide::FMAX;
            p.probe_iters = iters;
            rc = launch_wide(h, p, nullptr, true);
        } else {
            KParams p;
            fill_common(h, p);
            p.G = 1;
            p.probe_iters = iters;
            rc = launch_chunk(h, p, nullptr, true);
        }
        if (rc) return rc;
        rc = end_call(h, nullptr);
        if (rc) return rc;
        rc = finish_pending(h);
        if (rc) return rc;
    }
    *usec_per_exchange = h->last_ms * 1000.f / (float)iters;
    return WRNN_OK;
}

extern "C" int32_t wrnn_set_profiling(wrnn_handle *h, int32_t enable)
{
    if (!h) return fail(WRNN_ERR_INVALID, "null handle");
    DeviceGuard guard_(h->device);
    if (h->dense) {
        h->profiling = enable != 0;
        return WRNN_OK;
    }
    if (enable && !h->prof_dev) CUDA_TRY(cudaMalloc(&h->prof_dev, (size_t)NCTA * PROF_SLOTS * sizeof(long long)));
    h->profiling = enable != 0;
    return WRNN_OK;
}

extern "C" int32_t wrnn_get_stage_cycles(wrnn_handle *h, int64_t *out, int32_t n)
{
    if (!h || !out) return fail(WRNN_ERR_INVALID, "null argument");
    DeviceGuard guard_(h->device);
    int32_t rc = finish_pending(h);
    if (rc) return rc;
    if (h->dense) {     // dense kernel: [CTAs of the last launch][32] counters, slot map in csrc/wavernn_dense.cuh
        if (!h->dense_prof) return fail(WRNN_ERR_STATE, "profiling was never enabled");
        const size_t have = h->dense_prof_slots;
        CUDA_TRY(cudaMemcpy(out, h->dense_prof, ((size_t)n < have ? (size_t)n : have) * sizeof(long long), cudaMemcpyDeviceToHost));
        return WRNN_OK;
    }
    if (!h->prof_dev) return fail(WRNN_ERR_STATE, "profiling was never enabled");
    if (n != NCTA * PROF_SLOTS) return fail(WRNN_ERR_INVALID, "n must be %d", NCTA * PROF_SLOTS);
    CUDA_TRY(cudaMemcpy(out, h->prof_dev, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost));
    return WRNN_OK;
}

extern "C" int32_t wrnn_get_info(wrnn_handle *h, wrnn_info *out)
{
    if (!h || !out) return fail(WRNN_ERR_INVALID, "null argument");
    DeviceGuard guard_(h->device);
    finish_pending(h);                           // status and time of the last call (a fired watchdog shows in last_kernel_status)
    const bool wide = !h->dense && use_wide(h, wrnn_wide::FMAX);
    out->ctas = h->dense ? h->dense_clusters * wrnn_dense::CL : wide ? wrnn_wide::NWORK + h->wide_nsamp : NCTA;
    out->threads = h->dense ? wrnn_dense::DTHREADS : wide ? wrnn_wide::WTHREADS : NTHREADS;
    out->smem_bytes = h->dense ? wrnn_dense::SM_TOTAL : wide ? wrnn_wide::SM_BYTES : h->smem_bytes;
    out->folds_per_group = h->dense ? wrnn_dense::BC : wide ? wrnn_wide::FMAX : BT;
    out->max_folds_per_launch = h->dense ? h->dense_clusters * wrnn_dense::BC : wide ? wrnn_wide::FMAX : MAXG * BT;
    out->exchanges_per_step = h->dense ? 6 : wide ? 6 : NEXCH;
    out->sm_count = h->sm_count;
    out->launches = h->launches;
    out->epilogue_launches = g_epilogue_launches;
    out->last_kernel_status = h->last_status;
    out->last_kernel_ms = h->last_ms;
    out->kernel_kind = h->last_kernel;
    return WRNN_OK;
}

// ---- epilogue ----------------------------------------------------------------------------------
// numpy.linspace(start, stop, num): y_i = i*step + start with two roundings, last element = stop
static void np_linspace(double start, double stop, int64_t num, double *y)
{
    if (num <= 0) return;
    const int64_t div = num - 1;
    const double delta = stop - start;
    if (div > 0) {
        const double step = delta / (double)div;
        for (int64_t i = 0; i < num; ++i) {
            volatile double m = (double)i * step;
            y[i] = m + start;
        }
        y[num - 1] = stop;
    } else {
        volatile double m = 0.0 * delta;
        y[0] = m + start;
    }
}

// One thread per output sample.  fatchord_version.py:222-237 + utility/dsp.py:100-105.
// Rows of `samples` are the global folds [fold0, fold0 + B); out[i] is global position seg_start + i.
// A fold outside the given rows contributes nothing (the multi-GPU caller passes the neighbour's
// overlap samples as an extra row, expressive_speech_synthesis_research_b200/distributed.py).
__global__ void xfade_unfold_kernel(const float *__restrict__ samples, int B, int S, int batched, int overlap,
                                    const double *__restrict__ fade_in, const double *__restrict__ fade_out,
                                    const double *__restrict__ tail, int mu_classes, long long wave_len, int tail_len,
                                    long long fold0, long long seg_start, long long seg_len, double *__restrict__ out)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= seg_len) return;
    const long long p = seg_start + idx;
    double v;
    if (batched) {
        const long long hop = (long long)S - overlap;            // target + overlap
        const long long hi = p / hop;
        v = 0.0;                                                  // np.zeros(total_len), :375
        for (long long gi = hi - 1; gi <= hi; ++gi) {             // ascending fold order, :378-381
            const long long i = gi - fold0;
            if (i < 0 || i >= B) continue;
            const long long s = p - gi * hop;
            if (s < 0 || s >= S) continue;
            double y = (double)samples[i * S + s];                // .astype(np.float64), :224
            if (s < overlap) y = __dmul_rn(y, fade_in[s]);        // :372
            if (s >= S - overlap) y = __dmul_rn(y, fade_out[s - (S - overlap)]);   // :373
            v = __dadd_rn(v, y);
        }
    } else
        v = (double)samples[p];                                   // output[0], :229
    if (mu_classes) {                                             // decode_mu_law, dsp.py:103-104
        const double mu = (double)(mu_classes - 1);
        const double sgn = v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : 0.0);
        const double q = __ddiv_rn(sgn, mu);
        const double pw = pow(1.0 + mu, fabs(v));
        v = __dmul_rn(q, __dsub_rn(pw, 1.0));
    }
    if (p >= wave_len - tail_len) v = __dmul_rn(v, tail[p - (wave_len - tail_len)]);   // :235-237
    out[idx] = v;
}

static thread_local double *t_tables_dev = nullptr;
static thread_local int t_overlap = -1, t_tail = -1, t_device = -1;

static int32_t xfade_impl(const float *samples, int32_t num_folds, int32_t steps, int32_t batched, int32_t overlap,
                          int32_t mu_law_classes, int64_t wave_len, int32_t tail_fade, int64_t fold0, int64_t total_folds,
                          int64_t seg_start, int64_t seg_len, double *out, void *stream)
{
    if (!samples || !out) return fail(WRNN_ERR_INVALID, "null pointer argument");
    if (num_folds <= 0 || steps <= 0 || wave_len <= 0) return fail(WRNN_ERR_INVALID, "num_folds, steps and wave_len must be positive");
    if (tail_fade < 0 || tail_fade > wave_len)
        return fail(WRNN_ERR_INVALID, "tail fade of %d samples does not fit wave_len %lld (reference: broadcast ValueError, needs T >= 21 frames)", tail_fade, (long long)wave_len);
    int64_t total;
    if (batched) {
        if (overlap <= 0) return fail(WRNN_ERR_INVALID, "batched crossfade needs overlap > 0 (reference: y[:, -0:] broadcast error)");
        if (steps < 2 * overlap) return fail(WRNN_ERR_INVALID, "steps (%d) < 2 * overlap (%d)", steps, overlap);
        total = total_folds * (steps - overlap) + overlap;
    } else {
        if (num_folds != 1) return fail(WRNN_ERR_INVALID, "unbatched epilogue takes exactly one fold");
        total = steps;
        overlap = 0;
    }
    if (wave_len > total) return fail(WRNN_ERR_INVALID, "wave_len %lld exceeds the unfolded length %lld", (long long)wave_len, (long long)total);
    if (seg_start < 0 || seg_len <= 0 || seg_start + seg_len > wave_len)
        return fail(WRNN_ERR_INVALID, "segment [%lld, +%lld) outside [0, wave_len %lld)", (long long)seg_start, (long long)seg_len, (long long)wave_len);
    if (fold0 < 0 || fold0 + num_folds > total_folds) return fail(WRNN_ERR_INVALID, "fold rows [%lld, +%d) outside [0, %lld)", (long long)fold0, num_folds, (long long)total_folds);
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    cudaStream_t st = (cudaStream_t)stream;
    if (t_overlap != overlap || t_tail != tail_fade || t_device != dev) {
        // fade tables, fatchord_version.py:357-369 and :235
        const int sil = overlap / 2, fl = overlap - sil;
        std::vector<double> tab((size_t)2 * overlap + tail_fade + 1, 0.0), t((size_t)fl + 1);
        np_linspace(-1.0, 1.0, fl, t.data());
        for (int i = 0; i < fl; ++i) {
            volatile double a = 1.0 + t[i], b = 0.5 * a, c = 1.0 - t[i], d = 0.5 * c;
            tab[sil + i] = sqrt(b);               // fade_in  = [0]*sil ++ sqrt(.5(1+t))
            tab[overlap + i] = sqrt(d);           // fade_out = sqrt(.5(1-t)) ++ [0]*sil
        }
        np_linspace(1.0, 0.0, tail_fade, tab.data() + 2 * overlap);
        CUDA_TRY(cudaStreamSynchronize(st));      // previous users of the cached table are done
        if (t_tables_dev) cudaFree(t_tables_dev);
        t_tables_dev = nullptr;
        CUDA_TRY(cudaMalloc(&t_tables_dev, tab.size() * sizeof(double)));
        // on the caller's stream (a non-blocking stream is not ordered after the legacy stream), host buffer alive until done
        CUDA_TRY(cudaMemcpyAsync(t_tables_dev, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        t_overlap = overlap;
        t_tail = tail_fade;
        t_device = dev;
    }
    const int threads = 256;
    const long long blocks = (seg_len + threads - 1) / threads;
    xfade_unfold_kernel<<<(unsigned)blocks, threads, 0, st>>>(samples, num_folds, steps, batched, overlap, t_tables_dev,
                                                              t_tables_dev + overlap, t_tables_dev + 2 * overlap,
                                                              mu_law_classes, wave_len, tail_fade, fold0, seg_start, seg_len, out);
    CUDA_TRY(cudaGetLastError());
    g_epilogue_launches += 1;
    return WRNN_OK;
}

extern "C" int32_t wrnn_xfade_unfold(const float *samples, int32_t num_folds, int32_t steps,
                                     int32_t batched, int32_t overlap, int32_t mu_law_classes,
                                     int64_t wave_len, int32_t tail_fade, double *out, void *stream)
{
    return xfade_impl(samples, num_folds, steps, batched, overlap, mu_law_classes, wave_len, tail_fade, 0, num_folds, 0, wave_len, out, stream);
}

extern "C" int32_t wrnn_xfade_unfold_segment(const float *samples, int32_t num_rows, int32_t steps, int32_t overlap,
                                             int32_t mu_law_classes, int64_t wave_len, int32_t tail_fade,
                                             int64_t first_fold, int64_t total_folds, int64_t seg_start, int64_t seg_len,
                                             double *out, void *stream)
{
    return xfade_impl(samples, num_rows, steps, 1, overlap, mu_law_classes, wave_len, tail_fade, first_fold, total_folds, seg_start, seg_len, out, stream);
}

// ---------------------------------------------------------------------------------------------
// MelResNet at frame rate (csrc/wavernn_cond.cuh): one object per (device, model), independent of the step-loop engines
// ---------------------------------------------------------------------------------------------
struct wrnn_cond {
    int device = 0, res_blocks = 0;
    float *blob = nullptr;
    wrnn_mel::Tile *tiles_dev = nullptr, *tiles_host = nullptr;    // device table | pinned staging
    int tiles_cap = 0;
    cudaEvent_t copied = nullptr;                                   // the staging buffer has been read by the last call's copy
    int64_t launches = 0;
};

extern "C" int32_t wrnn_set_kernel(wrnn_handle *h, int32_t choice)
{
    if (!h || choice < -1 || choice > 1) return fail(WRNN_ERR_INVALID, "wrnn_set_kernel: choice must be -1 (by fold count), 0 (grouped) or 1 (wide)");
    h->kernel_choice = choice;
    return WRNN_OK;
}

extern "C" int64_t wrnn_cond_blob_floats(int32_t res_blocks) { return res_blocks < 0 ? -1 : (int64_t)wrnn_mel::blob_floats(res_blocks); }

extern "C" int32_t wrnn_cond_create(int32_t device, const float *blob_host, int64_t n_floats, int32_t res_blocks, wrnn_cond **out)
{
    if (!blob_host || !out || res_blocks < 0) return fail(WRNN_ERR_INVALID, "wrnn_cond_create: null argument");
    if (n_floats != (int64_t)wrnn_mel::blob_floats(res_blocks))
        return fail(WRNN_ERR_INVALID, "wrnn_cond_create: blob has %lld floats, %lld expected for %d residual blocks (80 x 5 -> 128 -> 128 only)",
                    (long long)n_floats, (long long)wrnn_mel::blob_floats(res_blocks), res_blocks);
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if ((int)prop.sharedMemPerBlockOptin < wrnn_mel::SM_BYTES) return fail(WRNN_ERR_INVALID, "wrnn_cond_create: device offers %d bytes of shared memory, %d needed", (int)prop.sharedMemPerBlockOptin, wrnn_mel::SM_BYTES);
    wrnn_cond *c = new wrnn_cond;
    c->device = device;
    c->res_blocks = res_blocks;
    cudaError_t e = cudaSuccess;
    if ((e = cudaMalloc(&c->blob, (size_t)n_floats * sizeof(float))) != cudaSuccess ||
        (e = cudaMemcpy(c->blob, blob_host, (size_t)n_floats * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c->copied, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(wrnn_mel::wavernn_melresnet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wrnn_mel::SM_BYTES)) != cudaSuccess) {
        cudaFree(c->blob);
        if (c->copied) cudaEventDestroy(c->copied);
        delete c;
        return fail(WRNN_ERR_CUDA, "wrnn_cond_create: %s", cudaGetErrorString(e));
    }
    *out = c;
    return WRNN_OK;
}

extern "C" void wrnn_cond_destroy(wrnn_cond *c)
{
    if (!c) return;
    DeviceGuard guard(c->device);
    cudaFree(c->blob);
    cudaFree(c->tiles_dev);
    cudaFreeHost(c->tiles_host);
    if (c->copied) cudaEventDestroy(c->copied);
    delete c;
}

extern "C" int64_t wrnn_cond_launches(const wrnn_cond *c) { return c ? c->launches : -1; }

// segments [nseg][3] (host): {first row of the segment's zero-padded mel frames in `mel_frames`, output frames T (the segment has
// T + 4 input rows), first output row in `aux_out`}.  Enqueues on `stream` and returns.
extern "C" int32_t wrnn_cond_frames(wrnn_cond *c, const float *mel_frames, const int32_t *segments, int32_t nseg, float *aux_out, void *stream)
{
    using namespace wrnn_mel;
    if (!c || !mel_frames || !segments || !aux_out || nseg < 0) return fail(WRNN_ERR_INVALID, "wrnn_cond_frames: null argument");
    DeviceGuard guard(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    int ntiles = 0;
    for (int s = 0; s < nseg; ++s) {
        if (segments[3 * s + 1] < 0 || segments[3 * s] < 0 || segments[3 * s + 2] < 0) return fail(WRNN_ERR_INVALID, "wrnn_cond_frames: negative segment field");
        ntiles += (segments[3 * s + 1] + TF - 1) / TF;
    }
    if (ntiles == 0) return WRNN_OK;
    if (ntiles > c->tiles_cap) {
        CUDA_TRY(cudaEventSynchronize(c->copied));
        cudaFree(c->tiles_dev);
        cudaFreeHost(c->tiles_host);
        c->tiles_dev = nullptr;
        c->tiles_host = nullptr;
        c->tiles_cap = 0;
        const int cap = ntiles + ntiles / 2 + 64;
        CUDA_TRY(cudaMalloc(&c->tiles_dev, (size_t)cap * sizeof(Tile)));
        CUDA_TRY(cudaMallocHost(&c->tiles_host, (size_t)cap * sizeof(Tile)));
        c->tiles_cap = cap;
    }
    CUDA_TRY(cudaEventSynchronize(c->copied));             // the previous call's copy has read the staging buffer (normally long ago)
    int k = 0;
    for (int s = 0; s < nseg; ++s) {
        const int row0 = segments[3 * s], T = segments[3 * s + 1], out0 = segments[3 * s + 2];
        for (int f = 0; f < T; f += TF, ++k) {
            Tile &t = c->tiles_host[k];
            t.mel_row0 = row0 + f;
            t.mel_rows = T + KS - 1 - f;                   // rows of the segment from this tile's first input row on
            t.out_row0 = out0 + f;
            t.nvalid = T - f < TF ? T - f : TF;
        }
    }
    CUDA_TRY(cudaMemcpyAsync(c->tiles_dev, c->tiles_host, (size_t)ntiles * sizeof(Tile), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaEventRecord(c->copied, st));
    MParams p;
    p.blob = c->blob;
    p.mel = mel_frames;
    p.aux = aux_out;
    p.tiles = c->tiles_dev;
    p.res_blocks = c->res_blocks;
    wavernn_melresnet_kernel<<<ntiles, NT, SM_BYTES, st>>>(p);
    CUDA_TRY(cudaGetLastError());
    c->launches += 1;
    return WRNN_OK;
}
