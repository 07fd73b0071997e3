// wavernn_dense.cuh -- dense-regime step loop of WaveRNN.generate (fatchord_version.py:171-222) on the Blackwell
// tensor cores: tcgen05.mma (bf16 x bf16 -> fp32 in tensor memory) with the weights STREAMED from L2 by bulk-TMA and
// the folds of a cluster advanced together.  precision = WRNN_PREC_BF16_DENSE.
//
// Work split.  A thread-block cluster of CL = 8 CTAs owns up to BC = 32 folds for all their steps; there is no
// grid-level synchronisation (clusters are independent, so a launch may hold more clusters than fit at once).
// CTA `rank` of a cluster owns hidden units [64 rank, 64 rank + 64) of rnn1, rnn2, fc1, fc2 and classes
// [64 rank, +64) of fc3 (MOL: the 30 outputs are rows 0-29 of rank 0).  Per step it multiplies its weight rows (MMA operand A,
// M = 128 for the [r | z] tiles, M = 64 for the n-gate and fc tiles, K-major, no swizzle) with the activations of ALL 512 units
// of the cluster's folds (operand B, N = 32 folds, K-major), which every CTA keeps as bf16 images in shared memory:
// image byte (k, f) = (k / 8) * 512 + f * 16 + (k % 8) * 2.
// A CTA's freshly computed 64 units are a contiguous 4 KB block of such an image; it is written locally and
// broadcast to the 7 peers with cp.async.bulk shared::cta -> shared::cluster, completing on the peers' mbarriers.
//
// Algebra (same folding as the fp32 kernel, csrc/wavernn_b200.cu::pack_images): the input layer I is folded into
// rnn1 / rnn2 / fc1, the sample x enters only through fp32 rank-1 terms in the epilogues, and everything that does
// not depend on the value produced by the current stage is issued early ("deferred" segments), so the per-step
// dependent chain is  E1 -> Wih2x.h1 -> E2 -> Wfc1x.h2 -> E3 -> Wfc2x.y1 -> E4 -> Wfc3.y2 -> E5 -> sample.
//
// Warp roles (384 threads): warps 0-7 epilogue (tensor memory -> registers -> gates / relu / sampling), warp 8 issues every
// tcgen05.mma (one elected lane), warps 9 and 10 lane 0 stream the weight bundles (one ring slot each), warp 11 lane 0
// relays "h2 image no longer read" between the CTAs of the cluster.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace wrnn_dense {

constexpr int CL = 8;                         // CTAs per cluster
constexpr int DHID = 512;
constexpr int UPC = DHID / CL;                 // hidden units (and RAW classes) per CTA
constexpr int BC = 32;                        // folds per cluster = MMA N
constexpr int NEPI = 256;                     // epilogue threads
constexpr int MMA_WARP = 8, PROD_WARP0 = 9, RELAY_WARP = 11;
constexpr int DTHREADS = NEPI + 128;
constexpr int SLOT = 32768, NSLOT = 2;        // weight ring
constexpr int CHUNK_B = BC * 16;              // one k-chunk (8 k) of an activation image
constexpr int IMG_B = (DHID / 8) * CHUNK_B;    // 32 KB: one activation image
constexpr int COND_CHUNKS = 26;               // 208 conditioning inputs: mel 80 | a1 | a2 | a3 | a4
constexpr int COND_B = COND_CHUNKS * CHUNK_B;
constexpr int SLICE_B = (UPC / 8) * CHUNK_B;  // a CTA's 64 units of an image
constexpr int NCLASS = 512;                   // RAW 9 bit
constexpr int MOL_C = 30, MOL_NR = 10, MOL_NU = 11;   // MOL: 10 mixtures x (logit, mean, log scale); 10 + 1 uniforms per draw
constexpr int FPC = BC / CL;                  // folds sampled by each CTA
constexpr int MAXSEG = 4, MAXBUNDLE = 40;
constexpr int NSV = 18;                       // per-row fp32 vectors (biases, x coefficients), [NSV][UPC] per CTA

enum { IMG_H1 = 0, IMG_H2 = 1, IMG_Y1 = 2, IMG_Y2 = 3 };
enum { W_NONE = 0, W_H1 = 1, W_H2 = 2, W_Y1 = 3, W_Y2 = 4, W_COND = 5 };       // wait before a bundle
enum { C_NONE = 0, C_G2 = 1, C_F1 = 2, C_F2 = 3, C_F3 = 4, C_G1 = 5, C_H2RD = 6 };   // commit after a bundle
// tensor-memory columns of the fp32 accumulators (BC columns each)
enum { D_G1_T0 = 0, D_G1_1H = 32, D_G1_1I = 64, D_G2_T0 = 96, D_G2_1H = 128, D_G2_1I = 160, D_F1 = 192, D_F2 = 224, D_F3 = 256 };
enum { DV_B1R = 0, DV_U1R, DV_B1Z, DV_U1Z, DV_B1NI, DV_U1N, DV_B1NH, DV_B2R, DV_U2R, DV_B2Z, DV_U2Z, DV_B2NI, DV_U2N, DV_B2NH, DV_B3, DV_U3, DV_B4, DV_B5 };

struct Seg {
    uint16_t off16;    // byte offset of the A tile in the ring slot / 16
    uint16_t rows;     // rows of the A tile: 128 -> M = 128 product, 64 -> M = 64 product
    uint16_t nk;       // k-steps of 16
    uint16_t bsrc16;   // B operand: byte offset from the first activation image / 16
    uint16_t dcol;     // accumulator column
    uint16_t first;    // 1: the first k-step overwrites the accumulator
};
struct Bundle {
    uint32_t bytes, src_off;     // payload size and offset in the per-rank stream
    uint16_t nseg, wait, commit, pad;
    Seg seg[MAXSEG];
};

// shared-memory map (bytes)
constexpr int SM_RING = 0;
constexpr int SM_IMG = SM_RING + NSLOT * SLOT;                 // 4 activation images, then the conditioning image
constexpr int SM_COND = SM_IMG + 4 * IMG_B;
constexpr int SM_SCRATCH = SM_COND + COND_B;                   // [BC][UPC] fp32: logits staging (E5)
constexpr int SM_SAMP = SM_SCRATCH + BC * UPC * 4;             // [CL src][FPC][UPC] fp32 logits of this CTA's folds
constexpr int SM_X = SM_SAMP + CL * FPC * UPC * 4;             // [BC] fp32 fed-back sample
constexpr int SM_FOLD = SM_X + BC * 4;                         // [4][BC] int32 fold geometry (rows mode uses the first two)
constexpr int SM_TAB = SM_FOLD + 4 * BC * 4;
constexpr int SM_BAR = SM_TAB + MAXBUNDLE * (int)sizeof(Bundle);
enum { B_FULL = 0, B_EMPTY = 2, B_ACT = 4, B_LG = 8, B_X = 9, B_COND = 10, B_ACC = 11, B_DONE = 16, B_H2RD = 17, B_H2OK = 18, NBAR = 20 };   // B_H2OK: two barriers, alternating by step
constexpr int SM_TMEM = SM_BAR + NBAR * 8;
constexpr int SM_TOTAL = SM_TMEM + 16;

struct DParams {
    const uint8_t *wstream;        // [CL][stream_bytes] bf16 operand tiles in bundle order
    const Bundle *table;           // [nb]
    const float *sv;               // [CL][NSV][UPC]
    const float *mels, *aux;       // fp32 [rows][80], [rows][128]  (UpsampleNetwork output, unfolded)
    const long long *fold_start, *fold_limit;
    // frames mode (wrnn_generate_folds_frames): the conditioning is expanded in the kernel from FRAME-rate tensors
    const float *mel_frames;       // [frames][80] mel frames of every utterance, zero-padded by `pad` frames on both sides
    const float *aux_frames;       // [frames][128] MelResNet output at frame rate
    const float *interp;           // [hop][5]: four composite interpolation weights of the three (repeat, FIR) stages + first-frame offset
    const int *fold_geo;           // [B][4]: first sample of the fold in its utterance, samples of the utterance, row of the
                                   // utterance's first (padded) mel frame, row of its first aux frame
    int hop, indent;               // samples per frame, pad * hop
    int mol;                       // 0: RAW, 512 classes (softmax + inverse CDF); 1: MOL, 30 outputs (discretized mix of logistics)
    const float *uniforms, *forced_x;
    float *logits_out, *samples_out;
    int *labels_out;
    unsigned long long seed;
    int *status;
    uint32_t stream_bytes;
    int nb;                        // bundles per step
    int B, S;                      // folds of the whole call (row stride of uniforms / logits), steps
    int fold0, nfolds, per;        // this launch: folds [fold0, fold0 + nfolds), `per` per cluster
    long long *prof;               // optional [CTAs][PROF_N] cycle counters (development), else nullptr
};
constexpr int PROF_N = 192;       // 0-31 counters, 32.. per-bundle trace of step 10: {wait done, ring slot full, MMAs issued} cycles since the step began
// profile slots: MMA thread 0-4 wait H1/H2/Y1/Y2/COND, 5 wait ring full, 6 issue, 7 total | epilogue (thread 0) 8-12 wait
// G1/G2/F1/F2/F3, 13 wait x, 14 wait logits, 15-19 stages E1..E5, 20 sampling, 21 conditioning, 22 total | producer 0: 24 wait empty, 25 total

// ---- PTX helpers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void mbar_init(uint32_t bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) { asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory"); }
// CTA-scope acquire is enough for everything that completes through the async proxy (TMA fill, DSMEM bulk copies, tcgen05.commit)
// or through local arrivals.  Only the sample broadcast (a remote generic store followed by a remote arrive) needs cluster scope,
// and ptxas follows every cluster-scope acquire with CCTL.IVALL, which throws the L1 contents away.
template <bool CLUSTER>
__device__ __forceinline__ bool mbar_try(uint32_t bar, unsigned parity)
{
    unsigned ok;
    if (CLUSTER)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// capped wait (~1 s): a stuck pipeline sets *status and lets every role run out instead of hanging the GPU
template <bool CLUSTER>
__device__ __noinline__ bool mbar_wait_slow(uint32_t bar, unsigned parity, int *status, int code)
{
    const long long t0 = clock64();
    for (;;) {
        for (int i = 0; i < 256; ++i)
            if (mbar_try<CLUSTER>(bar, parity)) return true;
        if (clock64() - t0 > 2000000000ll) {
            atomicCAS(status, 0, code);
            return false;
        }
    }
}
template <bool CLUSTER = false>
__device__ __forceinline__ bool mbar_wait(uint32_t bar, unsigned parity, int *status, int code)
{
    if (mbar_try<CLUSTER>(bar, parity)) return true;
    return mbar_wait_slow<CLUSTER>(bar, parity, status, code);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, unsigned bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2peer(uint32_t dst_cluster, uint32_t src_cta, unsigned bytes, uint32_t bar_cluster)
{
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}
// remote 4-byte store that completes 4 bytes of the destination CTA's mbarrier transaction count: no release fence on the sender
// (mbarrier.arrive.release.cluster costs a MEMBAR.ALL.GPU), no cluster-scope acquire on the receiver
__device__ __forceinline__ void st_async_remote(uint32_t dst_cluster, float v, uint32_t bar_cluster)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(dst_cluster), "r"(__float_as_uint(v)), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc]; descriptors are passed as their 32-bit halves so that stepping along K is one
// 32-bit add on the address field (the issuing thread's instruction latency is what paces short MMAs)
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 B stored contiguously (128 B),
// LBO = byte distance between the two core matrices of a k-step along K, SBO = between 8-row groups (128 B here)
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t addr, uint32_t lbo_bytes) { return ((addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16); }
constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);       // SBO = 128 B | descriptor version 1 (Blackwell)
// instruction descriptor: D fp32, A/B bf16, both K-major, N = BC; M = 128 for the [r | z] tiles, M = 64 for the 64-row tiles
// (an M = 64 product reads half the A bytes: 24 instead of 40 cycles, scripts/umma_rate.cu).  Accumulator lanes:
// M = 128: row i -> lane i;  M = 64: row i -> lane 32 (i / 16) + i % 16  (scripts/umma_m64.cu)
constexpr uint32_t IDESC_BASE = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BC >> 3) << 17);
constexpr uint32_t IDESC128 = IDESC_BASE | ((uint32_t)(128 >> 4) << 24), IDESC64 = IDESC_BASE | ((uint32_t)(64 >> 4) << 24);

// one MUFU per gate (the GRU epilogues are MUFU-bound): tanh.approx has 2^-11 relative error, well inside the bf16 rounding
// (2^-9) that every activation of this path goes through anyway
__device__ __forceinline__ float tanh_(float v) { float r; asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float sigmoid_(float v) { return fmaf(0.5f, tanh_(0.5f * v), 0.5f); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    const __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&b);
}

// Philox4x32-10, same stream as the fp32 kernel: uniform j of (step, fold) = word j & 3 of Philox({step, fold, j >> 2, 0}, seed)
__device__ __forceinline__ float philox_u01(unsigned long long seed, int step, int fold, int j = 0)
{
    uint4 c = make_uint4((unsigned)step, (unsigned)fold, (unsigned)(j >> 2), 0u);
    uint2 k = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
#pragma unroll 1
    for (int i = 0; i < 10; ++i) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    const unsigned w = (j & 3) == 0 ? c.x : (j & 3) == 1 ? c.y : (j & 3) == 2 ? c.z : c.w;
    return (float)(w >> 8) * (1.0f / 16777216.0f);
}

// ---- conditioning: item (kc, f) = 8 consecutive conditioning inputs of fold f, fp32 in global, bf16 in the image ----
struct CondRegs { uint4 v[4]; };                                // step t+2's items of this thread, already bf16
__device__ __forceinline__ uint4 cond_pack(const float4 &a, const float4 &b) { return make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w)); }
// rows mode: the UpsampleNetwork output is materialised ([rows][80], [rows][128]); fold f reads row fold_row0[f] + step
__device__ __forceinline__ void cond_load(const DParams &p, const int *geo, int nf, int tid, int step, CondRegs &cr)
{
    const int *fold_row0 = geo, *fold_lim = geo + BC;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = tid + NEPI * j;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (i < COND_CHUNKS * BC) {
            const int f = i & (BC - 1), kc = i >> 5;
            const long long row = (long long)fold_row0[f] + step;
            if (f < nf && step < p.S && row < (long long)fold_lim[f]) {
                const float *src = kc < 10 ? p.mels + row * 80 + kc * 8 : p.aux + row * 128 + (kc - 10) * 8;
                a = __ldg(reinterpret_cast<const float4 *>(src));
                b = __ldg(reinterpret_cast<const float4 *>(src) + 1);
            }
        }
        cr.v[j] = cond_pack(a, b);
    }
}
// frames mode: UpsampleNetwork.forward (fatchord_version.py:79-86) and fold_with_overlap's gather (:311-319) fused into the load.
//   aux[p]    = resnet_out[p / hop]                                  (Stretch2d of the MelResNet output, :80-82)
//   mel_up[p] = sum_j w[r][j] * mel_pad[q + s(r) + j],  q = (p + indent) / hop, r = (p + indent) % hop
// where w / s are the composite response of the three (Stretch2d, Conv2d box filter) stages (:83-85), measured once per model
// by pushing an impulse through those layers (host, wavernn.py); the crop by `indent` (:85) keeps the zero padding of the
// intermediate stages out of every retained sample.  Each item keeps (sample, frame, phase) and advances by one sample per call.
struct CondState { int pos[4], q[4], r[4]; };
__device__ __forceinline__ void cond_state_init(const DParams &p, const int *geo, int tid, CondState &st)
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = tid + NEPI * j;
        const int f = i & (BC - 1), kc = i >> 5;
        const int pos = geo[f];                                 // first sample of the fold
        const int shifted = kc < 10 ? pos + p.indent : pos;
        st.pos[j] = pos;
        st.q[j] = shifted / p.hop;
        st.r[j] = shifted - st.q[j] * p.hop;
    }
}
__device__ __forceinline__ void cond_load_frames(const DParams &p, const int *geo, int nf, int tid, int step, CondState &st, CondRegs &cr)
{
    const int *len = geo + BC, *mel_base = geo + 2 * BC, *aux_base = geo + 3 * BC;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = tid + NEPI * j;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (i < COND_CHUNKS * BC) {
            const int f = i & (BC - 1), kc = i >> 5;
            if (f < nf && step < p.S && st.pos[j] < len[f]) {
                if (kc < 10) {
                    const float *w = p.interp + st.r[j] * 5;
                    const float w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
                    const int first = (int)__ldg(w + 4);
                    const float4 *src = reinterpret_cast<const float4 *>(p.mel_frames + (size_t)(mel_base[f] + st.q[j] + first) * 80 + kc * 8);
                    const float4 a0 = __ldg(src), b0 = __ldg(src + 1), a1 = __ldg(src + 20), b1 = __ldg(src + 21);
                    const float4 a2 = __ldg(src + 40), b2 = __ldg(src + 41), a3 = __ldg(src + 60), b3 = __ldg(src + 61);
#define WRNN_MIX(c) fmaf(w3, a3.c, fmaf(w2, a2.c, fmaf(w1, a1.c, w0 * a0.c)))
#define WRNN_MIXB(c) fmaf(w3, b3.c, fmaf(w2, b2.c, fmaf(w1, b1.c, w0 * b0.c)))
                    a = make_float4(WRNN_MIX(x), WRNN_MIX(y), WRNN_MIX(z), WRNN_MIX(w));
                    b = make_float4(WRNN_MIXB(x), WRNN_MIXB(y), WRNN_MIXB(z), WRNN_MIXB(w));
#undef WRNN_MIX
#undef WRNN_MIXB
                } else {
                    const float4 *src = reinterpret_cast<const float4 *>(p.aux_frames + (size_t)(aux_base[f] + st.q[j]) * 128 + (kc - 10) * 8);
                    a = __ldg(src);
                    b = __ldg(src + 1);
                }
            }
        }
        cr.v[j] = cond_pack(a, b);
        st.pos[j] += 1;
        if (++st.r[j] == p.hop) {
            st.r[j] = 0;
            st.q[j] += 1;
        }
    }
}
__device__ __forceinline__ void cond_store(uint8_t *cond_img, int tid, const CondRegs &cr)
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = tid + NEPI * j;
        if (i < COND_CHUNKS * BC) {
            const int f = i & (BC - 1), kc = i >> 5;
            *reinterpret_cast<uint4 *>(cond_img + kc * CHUNK_B + f * 16) = cr.v[j];
        }
    }
}

// broadcast this CTA's 4 KB slice of activation image `img` to the 7 peers: called by warp 0, lane q > 0 serves peer
// rank + q (issuing a bulk copy stalls the issuing thread, so the seven copies are issued by seven lanes), lane 0 arms
// the CTA's own barrier with the bytes it expects from its peers
__device__ __forceinline__ void send_slice(uint32_t smem_base, int img, uint32_t rank, int lane)
{
    const uint32_t src = smem_base + SM_IMG + img * IMG_B + rank * SLICE_B;
    const uint32_t bar = smem_base + SM_BAR + (B_ACT + img) * 8;
    if (lane == 0) mbar_expect_tx(bar, (CL - 1) * SLICE_B);
    else if (lane < CL) {
        const uint32_t peer = (rank + (uint32_t)lane) & (CL - 1);
        bulk_s2peer(mapa(src, peer), src, SLICE_B, mapa(bar, peer));
    }
    __syncwarp();
}

// GRU epilogue of one layer (torch.nn.GRUCell, fatchord_version.py:190,194).  Accumulator lanes of warp quad q:
//   lanes 0-15  ("r lanes"): unit 16 q + l:      r (tile 0, M = 128), W_hn.h and W_in.x (64-row tiles, M = 64)
//   lanes 16-31 ("z lanes"): unit 16 q + l - 16: z (tile 0); they own the fp32 state of the unit
// Both kinds run the same instruction stream: gate = sigmoid(acc0 + ...) is r or z by lane, the candidate n = tanh(i_n + r h_n)
// is meaningful in r lanes and handed to the unit's z lane with one shuffle, which finishes h' = n + z (h - n) and writes bf16.
__device__ __forceinline__ void gru_epilogue(uint32_t tmem, int col_t0, int col_1h, int col_1i, uint8_t *smem, int img, uint32_t rank, int warp, int lane,
                                             float b_g, float u_g, float b_ni, float u_n, float b_nh, float (&hprev)[16])
{
    const int q = warp & 3, hf = warp >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float *xs = reinterpret_cast<const float *>(smem + SM_X);
    float g0[16], nh[16], ni[16];
    tc_ld16(tmem + lane_base + col_t0 + 16 * hf, g0);
    tc_ld16(tmem + lane_base + col_1h + 16 * hf, nh);
    tc_ld16(tmem + lane_base + col_1i + 16 * hf, ni);
    tc_ld_wait();
    const bool zlane = lane >= 16;
    const int u = 16 * q + (lane & 15);
    uint8_t *dst = smem + SM_IMG + img * IMG_B + (rank * (UPC / 8) + (u >> 3)) * CHUNK_B + (u & 7) * 2;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int f = 16 * hf + j;
        const float x = xs[f];
        const float gate = sigmoid_(g0[j] + fmaf(x, u_g, b_g));
        float nn = tanh_(ni[j] + fmaf(x, u_n, b_ni) + gate * (nh[j] + b_nh));
        nn = __shfl_sync(0xffffffffu, nn, lane & 15);
        if (zlane) {
            const float h = fmaf(gate, hprev[j] - nn, nn);
            hprev[j] = h;
            *reinterpret_cast<__nv_bfloat16 *>(dst + f * 16) = __float2bfloat16_rn(h);
        }
    }
    fence_async_smem();
    tc_fence_before();
    epi_sync();
}

// fc epilogue: 64-row tile (M = 64): unit 16 q + l in lanes 0-15 of every quad; relu(acc + b + x u) -> bf16 image
__device__ __forceinline__ void fc_epilogue(uint32_t tmem, int col, uint8_t *smem, int img, uint32_t rank, int warp, int lane, float b, float ux)
{
    const int q = warp & 3, hf = warp >> 2;
    const int u = 16 * q + (lane & 15);
    const float *xs = reinterpret_cast<const float *>(smem + SM_X);
    float a[16];
    tc_ld16(tmem + ((uint32_t)(q * 32) << 16) + col + 16 * hf, a);
    tc_ld_wait();
    if (lane < 16) {
        uint8_t *dst = smem + SM_IMG + img * IMG_B + (rank * (UPC / 8) + (u >> 3)) * CHUNK_B + (u & 7) * 2;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int f = 16 * hf + j;
            const float y = fmaxf(a[j] + fmaf(xs[f], ux, b), 0.0f);
            *reinterpret_cast<__nv_bfloat16 *>(dst + f * 16) = __float2bfloat16_rn(y);
        }
    }
    fence_async_smem();
    tc_fence_before();
    epi_sync();
}

template <bool PROF, bool FRAMES>
__device__ __forceinline__ void dense_body(const DParams &p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int ci = blockIdx.x / CL;
    const int cfold0 = p.fold0 + ci * p.per;
    int nf = p.fold0 + p.nfolds - cfold0;
    nf = nf < p.per ? nf : p.per;
    if (nf <= 0) return;                                       // whole cluster: nothing to do
    const uint32_t sb = s32(smem);
    const uint32_t bar0 = sb + SM_BAR;
    int *geo = reinterpret_cast<int *>(smem + SM_FOLD);
    Bundle *tab = reinterpret_cast<Bundle *>(smem + SM_TAB);

    // ---- set-up: zero the images / conditioning / x, barriers, bundle table, tensor memory
    for (int i = tid; i < (SM_FOLD - SM_IMG) / 16; i += DTHREADS) reinterpret_cast<uint4 *>(smem + SM_IMG)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < p.nb * (int)(sizeof(Bundle) / 4); i += DTHREADS) reinterpret_cast<uint32_t *>(tab)[i] = reinterpret_cast<const uint32_t *>(p.table)[i];
    if (tid < BC) {
        const int f = tid;
        if (FRAMES) {
            for (int c = 0; c < 4; ++c) geo[c * BC + f] = f < nf ? p.fold_geo[(size_t)(cfold0 + f) * 4 + c] : 0;
        } else {
            geo[f] = f < nf ? (int)p.fold_start[cfold0 + f] : 0;
            geo[BC + f] = f < nf ? (int)p.fold_limit[cfold0 + f] : 0;
        }
    }
    if (tid == 0) {
        for (int s = 0; s < NSLOT; ++s) {
            mbar_init(bar0 + (B_FULL + s) * 8, 1);
            mbar_init(bar0 + (B_EMPTY + s) * 8, 1);
        }
        for (int i = 0; i < 4; ++i) mbar_init(bar0 + (B_ACT + i) * 8, 1);
        mbar_init(bar0 + B_LG * 8, 1);
        mbar_init(bar0 + B_X * 8, 1);
        mbar_init(bar0 + B_COND * 8, NEPI / 32);
        for (int i = 0; i < 5; ++i) mbar_init(bar0 + (B_ACC + i) * 8, 1);
        mbar_init(bar0 + B_DONE * 8, 1);
        mbar_init(bar0 + B_H2RD * 8, 1);
        mbar_init(bar0 + B_H2OK * 8, CL);
        mbar_init(bar0 + (B_H2OK + 1) * 8, CL);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sb + SM_TMEM), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<const uint32_t *>(smem + SM_TMEM);
    const int S = p.S;

    if (warp == RELAY_WARP) {
        // ===== relay: once this CTA's Whh2.h2(t) products have completed (they read the h2(t) image long after the step's
        // other users), tell every CTA of the cluster; a CTA broadcasts its slice of h2(t+1) only when all eight have
        if (lane == 0) {
            for (int t = -1; t < S; ++t) {
                if (!mbar_wait(bar0 + B_H2RD * 8, (unsigned)((t + 1) & 1), p.status, 30)) break;
#pragma unroll 1
                for (uint32_t d = 0; d < CL; ++d) mbar_arrive_remote(mapa(bar0 + (B_H2OK + ((t + 1) & 1)) * 8, d));   // read by step t+1
            }
        }
    } else if (warp >= PROD_WARP0) {
        // ===== weight stream: producer pw owns ring slot pw; bundle gb of the launch goes to slot gb & 1 =====
        if (lane == 0) {
            const int pw = warp - PROD_WARP0;
            const uint8_t *src = p.wstream + (size_t)rank * p.stream_bytes;
            const uint32_t full = bar0 + (B_FULL + pw) * 8, empty = bar0 + (B_EMPTY + pw) * 8, dst = sb + SM_RING + pw * SLOT;
            const long long total = (long long)(S + 1) * p.nb;
            int b = pw;                                         // bundle index within the step
            long long t_empty = 0;
            const long long t_begin = clock64();
            for (long long gb = pw, use = 0; gb < total; gb += NSLOT, ++use) {
                const long long t0 = PROF ? clock64() : 0;
                if (use > 0 && !mbar_wait(empty, (unsigned)((use - 1) & 1), p.status, 1)) break;
                if (PROF) t_empty += clock64() - t0;
                const uint32_t bytes = tab[b].bytes;
                mbar_expect_tx(full, bytes);
                bulk_g2s(dst, src + tab[b].src_off, bytes, full);
                b += NSLOT;
                if (b >= p.nb) b -= p.nb;
            }
            if (PROF && p.prof && pw == 0) {
                p.prof[(size_t)blockIdx.x * PROF_N + 24] = t_empty;
                p.prof[(size_t)blockIdx.x * PROF_N + 25] = clock64() - t_begin;
            }
        }
    } else if (warp == MMA_WARP) {
        // ===== MMA issue.  One elected lane runs the whole role: inside an elect.sync region ptxas keeps descriptors and loop state
        // on the uniform datapath (a plain `lane == 0` test costs a register->uniform waterfall per product).  The tensor pipe holds
        // next to nothing in flight, so whatever the issuing thread does between two products idles the pipe (scripts/umma_issue.cu:
        // 40 cycles per product back to back, +285 cycles per bundle when every bundle re-enters the elect region, +170 like this).
        // t = -1 runs the same table on zero images and primes the accumulators of step 0.
        if (elect_one()) {
            unsigned ph_full = 0, ph_wait = 0;                  // phase bits: ring slots / wait events
            long long gb = 0;
            bool ok = true;
            long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const long long t_begin = clock64();
            for (int t = -1; t < S && ok; ++t) {
                const bool pre = t < 0;
                for (int b = 0; b < p.nb && ok; ++b, ++gb) {
                    const Bundle &bd = tab[b];
                    const int w = bd.wait;
                    if (w != W_NONE && (w == W_COND || !pre)) {
                        const uint32_t wb = w == W_COND ? bar0 + B_COND * 8 : bar0 + (B_ACT + (w - W_H1)) * 8;
                        const long long t0 = PROF ? clock64() : 0;
                        ok = mbar_wait(wb, (ph_wait >> w) & 1u, p.status, 10 + w);
                        if (PROF) pf[w - 1] += clock64() - t0;
                        ph_wait ^= 1u << w;
                        if (!ok) break;
                    }
                    const int slot = (int)(gb & 1);
                    const long long t1 = PROF ? clock64() : 0;
                    ok = mbar_wait(bar0 + (B_FULL + slot) * 8, (ph_full >> slot) & 1u, p.status, 2);
                    const long long t2 = PROF ? clock64() : 0;
                    if (PROF) pf[5] += t2 - t1;
                    ph_full ^= 1u << slot;
                    if (!ok) break;
                    tc_fence_after();
                    const uint32_t slot_base = sb + SM_RING + slot * SLOT;
                    const int nseg = bd.nseg;
                    for (int s = 0; s < nseg; ++s) {
                        const Seg sg = bd.seg[s];
                        const uint32_t rows = sg.rows, d = tmem + sg.dcol;
                        const uint32_t IDESC = rows == 64 ? IDESC64 : IDESC128;
                        uint32_t a_lo = smem_desc_lo(slot_base + (uint32_t)sg.off16 * 16, rows * 16);
                        uint32_t b_lo = smem_desc_lo(sb + SM_IMG + (uint32_t)sg.bsrc16 * 16, CHUNK_B);
                        const uint32_t a_inc = rows * 2;                  // one k-step = 2 chunks of rows x 16 B, in 16-byte units
                        constexpr uint32_t b_inc = 2 * CHUNK_B / 16;
                        tc_mma(d, a_lo, b_lo, DESC_HI, IDESC, sg.first ? 0u : 1u);
                        const int nk = sg.nk;
#pragma unroll 4
                        for (int k = 1; k < nk; ++k) {
                            a_lo += a_inc;
                            b_lo += b_inc;
                            tc_mma(d, a_lo, b_lo, DESC_HI, IDESC, 1u);
                        }
                    }
                    const int c = bd.commit;
                    tc_commit(bar0 + (B_EMPTY + slot) * 8);
                    if (c == C_H2RD) tc_commit(bar0 + B_H2RD * 8);
                    else if (c != C_NONE && (c == C_G1 || !pre)) tc_commit(bar0 + (B_ACC + (c == C_G1 ? 0 : c)) * 8);
                    if (PROF) pf[6] += clock64() - t2;
                }
            }
            tc_commit(bar0 + B_DONE * 8);
            mbar_wait(bar0 + B_DONE * 8, 0, p.status, 3);
            pf[7] = clock64() - t_begin;
            if (PROF && p.prof)
                for (int i = 0; i < 8; ++i) p.prof[(size_t)blockIdx.x * PROF_N + i] = pf[i];
        }
        __syncwarp();
    } else {
        // ===== epilogue warps =====
        const int q = warp & 3;
        const float *sv = p.sv + (size_t)rank * NSV * UPC;
        const int u = 16 * q + (lane & 15);                     // unit of this thread's accumulator lane
        const bool zl = lane >= 16;                             // z lane (owns the state) or r lane
        const float b1g = sv[(zl ? DV_B1Z : DV_B1R) * UPC + u], u1g = sv[(zl ? DV_U1Z : DV_U1R) * UPC + u];
        const float b1ni = sv[DV_B1NI * UPC + u], u1n = sv[DV_U1N * UPC + u], b1nh = sv[DV_B1NH * UPC + u];
        const float b2g = sv[(zl ? DV_B2Z : DV_B2R) * UPC + u], u2g = sv[(zl ? DV_U2Z : DV_U2R) * UPC + u];
        const float b2ni = sv[DV_B2NI * UPC + u], u2n = sv[DV_U2N * UPC + u], b2nh = sv[DV_B2NH * UPC + u];
        const float b3 = sv[DV_B3 * UPC + u], u3 = sv[DV_U3 * UPC + u], b4 = sv[DV_B4 * UPC + u], b5 = sv[DV_B5 * UPC + u];
        float h1[16], h2[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) h1[j] = h2[j] = 0.f;
        uint8_t *cond_img = smem + SM_COND;
        CondRegs cr;
        // prologue: conditioning of step 0 -> image, conditioning of step 1 -> registers
        CondState cst;
        if (FRAMES) {
            cond_state_init(p, geo, tid, cst);
            cond_load_frames(p, geo, nf, tid, 0, cst, cr);
        } else
            cond_load(p, geo, nf, tid, 0, cr);
        cond_store(cond_img, tid, cr);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + B_COND * 8);
        if (FRAMES) cond_load_frames(p, geo, nf, tid, 1, cst, cr);
        else cond_load(p, geo, nf, tid, 1, cr);

        if (tid == 0) mbar_expect_tx(bar0 + B_X * 8, BC * 4);   // the 32 samples of step 0 (one st.async of 4 bytes per fold)
        unsigned ph = 0;                                        // phase bits of B_ACC + {0..4}, B_X (bit 5), B_LG (bit 6)
        bool ok = true;
        long long pe[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tk = clock64();
        const long long te_begin = tk;
#define TICK(slot) do { if (PROF) { const long long now_ = clock64(); pe[slot] += now_ - tk; tk = now_; } } while (0)
        for (int t = 0; t < S && ok; ++t) {
            // ---- E1: h1(t) = GRU1(I(x(t-1), c(t)), h1(t-1)) -- fatchord_version.py:188-190
            ok = mbar_wait(bar0 + (B_ACC + 0) * 8, ph & 1u, p.status, 20);
            ph ^= 1u;
            if (!ok) break;
            TICK(0);
            if (t > 0) {
                ok = mbar_wait(bar0 + B_X * 8, (ph >> 5) & 1u, p.status, 25);
                ph ^= 1u << 5;
                if (!ok) break;
                if (tid == 0) mbar_expect_tx(bar0 + B_X * 8, BC * 4);   // arm the next step's phase
            }
            TICK(5);
            tc_fence_after();
            gru_epilogue(tmem, D_G1_T0, D_G1_1H, D_G1_1I, smem, IMG_H1, rank, warp, lane, b1g, u1g, b1ni, u1n, b1nh, h1);
            if (warp == 0) send_slice(sb, IMG_H1, rank, lane);
            TICK(7);
            // conditioning of step t+1 -> image (all MMAs that read the image of step t completed before the G1 commit)
            cond_store(cond_img, tid, cr);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + B_COND * 8);
            TICK(13);
            // ---- E2: h2(t) = GRU2([x + h1, a2], h2(t-1)) -- :192-194
            ok = mbar_wait(bar0 + (B_ACC + C_G2) * 8, (ph >> C_G2) & 1u, p.status, 21);
            ph ^= 1u << C_G2;
            if (!ok) break;
            TICK(1);
            tc_fence_after();
            gru_epilogue(tmem, D_G2_T0, D_G2_1H, D_G2_1I, smem, IMG_H2, rank, warp, lane, b2g, u2g, b2ni, u2n, b2nh, h2);
            if (warp == 0) {                                    // every CTA has finished reading the h2(t-1) image (relay warp)
                ok = mbar_wait(bar0 + (B_H2OK + (t & 1)) * 8, (unsigned)((t >> 1) & 1), p.status, 27);
                send_slice(sb, IMG_H2, rank, lane);
            }
            TICK(8);
            // ---- E3: y1 = relu(fc1([x + h1 + h2, a3])) -- :196-198
            ok = mbar_wait(bar0 + (B_ACC + C_F1) * 8, (ph >> C_F1) & 1u, p.status, 22);
            ph ^= 1u << C_F1;
            if (!ok) break;
            TICK(2);
            tc_fence_after();
            fc_epilogue(tmem, D_F1, smem, IMG_Y1, rank, warp, lane, b3, u3);
            if (warp == 0) send_slice(sb, IMG_Y1, rank, lane);
            TICK(9);
            // conditioning of step t+2 -> registers (bf16): the global-load latency it waits for hides in the wait for fc2's products
            if (FRAMES) cond_load_frames(p, geo, nf, tid, t + 2, cst, cr);
            else cond_load(p, geo, nf, tid, t + 2, cr);
            TICK(13);
            // ---- E4: y2 = relu(fc2([y1, a4])) -- :200-201
            ok = mbar_wait(bar0 + (B_ACC + C_F2) * 8, (ph >> C_F2) & 1u, p.status, 23);
            ph ^= 1u << C_F2;
            if (!ok) break;
            TICK(3);
            tc_fence_after();
            fc_epilogue(tmem, D_F2, smem, IMG_Y2, rank, warp, lane, b4, 0.0f);
            if (warp == 0) send_slice(sb, IMG_Y2, rank, lane);
            TICK(10);
            // ---- E5: logits = fc3(y2) (:202): this CTA's 64 classes of every fold -> the CTA that samples the fold
            ok = mbar_wait(bar0 + (B_ACC + C_F3) * 8, (ph >> C_F3) & 1u, p.status, 24);
            ph ^= 1u << C_F3;
            if (!ok) break;
            TICK(4);
            tc_fence_after();
            {
                const int hf = warp >> 2;
                float a[16];
                tc_ld16(tmem + ((uint32_t)(q * 32) << 16) + D_F3 + 16 * hf, a);
                tc_ld_wait();
                float *stage = reinterpret_cast<float *>(smem + SM_SCRATCH);
                if (lane < 16) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) stage[(16 * hf + j) * UPC + u] = a[j] + b5;
                }
                fence_async_smem();
            }
            tc_fence_before();
            epi_sync();
            if (warp == 0) {                                    // lane d < CL: folds FPC d .. FPC d + FPC - 1 are sampled by CTA d
                const uint32_t lgbar = bar0 + B_LG * 8;
                if (lane == 0) mbar_expect_tx(lgbar, CL * FPC * UPC * 4);
                __syncwarp();
                if (lane < CL)
                    bulk_s2peer(mapa(sb + SM_SAMP + rank * (FPC * UPC * 4), (uint32_t)lane), sb + SM_SCRATCH + lane * (FPC * UPC * 4), FPC * UPC * 4, mapa(lgbar, (uint32_t)lane));
                __syncwarp();
            }
            TICK(11);
            // ---- sampling (:210-216): warp w < FPC takes fold FPC rank + w; softmax + inverse CDF as in the fp32 kernel
            if (warp < FPC) {
                const int fl = FPC * (int)rank + warp;          // fold within the cluster
                const int bglob = cfold0 + fl;
                float uu = 0.f, fx = 0.f;                       // the draws and the forced value are fetched while the logits travel
                if (fl < nf) {
                    if (p.mol) {                                // lane j < 11: uniform j of this (step, fold)
                        if (lane < MOL_NU) uu = p.uniforms ? __ldg(p.uniforms + ((size_t)t * p.B + bglob) * MOL_NU + lane) : philox_u01(p.seed, t, bglob, lane);
                    } else if (lane == 0)
                        uu = p.uniforms ? __ldg(p.uniforms + (size_t)t * p.B + bglob) : philox_u01(p.seed, t, bglob);
                    if (lane == 0 && p.forced_x) fx = __ldg(p.forced_x + (size_t)t * p.B + bglob);
                }
                ok = mbar_wait(bar0 + B_LG * 8, (ph >> 6) & 1u, p.status, 26);
                ph ^= 1u << 6;
                if (!ok) break;
                TICK(6);
                if (p.mol) {
                    // sample_from_discretized_mix_logistic (utility/distribution.py:87-123) as in the fp32 kernel: the 30 outputs are
                    // rows 0-29 of CTA 0's tile; lane i < 10 scores mixture i (Gumbel-max), lane 0 draws from the chosen logistic
                    const float *l30 = reinterpret_cast<const float *>(smem + SM_SAMP) + warp * UPC;     // source CTA 0, fold `warp`
                    const float my_logit = lane < MOL_C ? l30[lane] : 0.f;
                    float best = -INFINITY;
                    int arg = 1 << 20;
                    if (lane < MOL_NR) {
                        const float u = (float)(1e-5 + ((1.0 - 1e-5) - 1e-5) * (double)uu);
                        best = l30[lane] - logf(-logf(u));
                        arg = lane;
                    }
#pragma unroll
                    for (int off = 1; off <= 8; off <<= 1) {
                        const float b2 = __shfl_xor_sync(0xffffffffu, best, off);
                        const int a2 = __shfl_xor_sync(0xffffffffu, arg, off);
                        if (b2 > best || (b2 == best && a2 < arg)) {
                            best = b2;
                            arg = a2;
                        }
                    }
                    arg = __shfl_sync(0xffffffffu, arg, 0);
                    const float u2r = __shfl_sync(0xffffffffu, uu, MOL_NR);
                    float x = 0.f;
                    if (lane == 0 && fl < nf) {
                        const float mean = l30[MOL_NR + arg];
                        const float ls = fmaxf(l30[2 * MOL_NR + arg], -32.23619130191664f);
                        const float u2 = (float)(1e-5 + ((1.0 - 1e-5) - 1e-5) * (double)u2r);
                        x = mean + expf(ls) * (logf(u2) - logf(1.0f - u2));
                        x = fminf(fmaxf(x, -1.0f), 1.0f);
                    }
                    x = __shfl_sync(0xffffffffu, x, 0);
                    fx = __shfl_sync(0xffffffffu, fx, 0);
                    const float xnext = fl < nf ? (p.forced_x ? fx : x) : 0.f;
                    if (lane < CL) {
                        st_async_remote(mapa(sb + SM_X + fl * 4, (uint32_t)lane), xnext, mapa(bar0 + B_X * 8, (uint32_t)lane));
                    }
                    if (lane == 0 && fl < nf) {
                        p.samples_out[(size_t)bglob * S + t] = x;
                        if (p.labels_out) p.labels_out[(size_t)bglob * S + t] = arg;
                    }
                    if (p.logits_out && fl < nf && lane < MOL_C) p.logits_out[((size_t)t * p.B + bglob) * MOL_C + lane] = my_logit;
                    TICK(12);
                    continue;
                }
                const float *lg = reinterpret_cast<const float *>(smem + SM_SAMP) + ((lane >> 2) * FPC + warp) * UPC + (lane & 3) * 16;
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 qv = *reinterpret_cast<const float4 *>(lg + j);
                    v[j] = qv.x; v[j + 1] = qv.y; v[j + 2] = qv.z; v[j + 3] = qv.w;
                }
                float lgv[16];                                  // the raw logits, written out (if asked for) after the feedback is on its way
#pragma unroll
                for (int j = 0; j < 16; ++j) lgv[j] = v[j];
                float m = v[0];
#pragma unroll
                for (int j = 1; j < 16; ++j) m = fmaxf(m, v[j]);
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
                float run = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    run += __expf(v[j] - m);
                    v[j] = run;
                }
                float incl = run;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const float tt = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += tt;
                }
                const float excl = incl - run;
                const float total = __shfl_sync(0xffffffffu, incl, 31);
                uu = __shfl_sync(0xffffffffu, uu, 0);
                fx = __shfl_sync(0xffffffffu, fx, 0);
                const float thr = uu * total;
                int cnt = 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) cnt += (excl + v[j] <= thr) ? 1 : 0;
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
                const int kk = cnt > NCLASS - 1 ? NCLASS - 1 : cnt;
                const float sample = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, (float)kk), (float)NCLASS - 1.0f), 1.0f);   // :214
                const float xnext = fl < nf ? (p.forced_x ? fx : sample) : 0.f;
                // the feedback first: the release that follows the remote store is a full memory barrier for the thread, so the global
                // stores of the outputs are issued after it, not before
                if (lane < CL) {                                // lane d delivers x to CTA d
                    st_async_remote(mapa(sb + SM_X + fl * 4, (uint32_t)lane), xnext, mapa(bar0 + B_X * 8, (uint32_t)lane));
                }
                if (lane == 0 && fl < nf) {
                    p.samples_out[(size_t)bglob * S + t] = sample;
                    if (p.labels_out) p.labels_out[(size_t)bglob * S + t] = kk;
                }
                if (p.logits_out && fl < nf) {
                    float *dst = p.logits_out + ((size_t)t * p.B + bglob) * NCLASS + lane * 16;
#pragma unroll
                    for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4 *>(dst + j) = make_float4(lgv[j], lgv[j + 1], lgv[j + 2], lgv[j + 3]);
                }
                TICK(12);
            }
        }
#undef TICK
        if (PROF && p.prof && tid == 0) {
            // waits: 8-12 acc G1/G2/F1/F2/F3, 13 x, 14 logits | work: 15-19 E1..E5, 20 sampling, 21 conditioning
            static const int map[14] = {8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21};
            for (int i = 0; i < 14; ++i) p.prof[(size_t)blockIdx.x * PROF_N + map[i]] = pe[i];
            p.prof[(size_t)blockIdx.x * PROF_N + 22] = clock64() - te_begin;
        }
    }
    // ---- teardown: every role is done (or timed out); peers may still be writing into this CTA until the cluster barrier
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(DTHREADS, 1) wavernn_dense_kernel(const DParams p) { dense_body<false, false>(p); }
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(DTHREADS, 1) wavernn_dense_kernel_prof(const DParams p) { dense_body<true, false>(p); }
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(DTHREADS, 1) wavernn_dense_kernel_frames(const DParams p) { dense_body<false, true>(p); }

}   // namespace wrnn_dense
