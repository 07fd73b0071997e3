// Persistent sm_100a kernel for the WaveRNN step loop (reference: WaveRNN/models/
// fatchord_version.py:171-222).  See DESIGN.md for the derivation; summary:
//
//  * 128 CTAs (one per SM, cooperative launch), each owns 4 of the 512 hidden units of every
//    layer; its rows of every weight matrix stay resident in shared memory for all steps.
//  * Folds advance in groups of 8.  A group goes through 5 grid-level exchanges per step
//    (h1 | h2 | y1 | y2 | logits).  The exchange is flag-free ("LL" protocol): every value
//    travels as an 8-byte {fp32 value, step epoch} pair written with one store; a consumer
//    polls the pairs it needs until their epoch matches -- one L2 round trip, no fences, no
//    contended flag lines (profiles/r01_exchange_microbench.md).
//  * The 16 warps of a CTA are split into up to 3 TEAMS; team j advances groups j, j+T, ...
//    on its own (named barriers, private staging buffers), so the exchange latency, sampling
//    and pointwise phases of one group overlap the mat-vecs of the others.
//  * The input layer I and every conditioning term are folded algebraically into the
//    downstream layers at load time (host, fp64), so the recurrence only multiplies
//    h1, h2, s=h1+h2, y1, y2; conditioning projections for step t+1 are computed during
//    step t from TMA-staged rows of the UNFOLDED conditioning (fold gather fused).
//  * Mat-vec work items use packed FFMA2 (fma.rn.f32x2, two folds per instruction, each lane
//    IEEE-rounded exactly like fmaf) and a select-free butterfly reduce-scatter: weight rows
//    and fold slots are pre-permuted per lane so every level is shuffle + add only.
//  * Sampling (softmax inverse-CDF / mixture-of-logistics) is done redundantly by every
//    CTA straight from the polled logits, so no broadcast exchange is needed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wrnn {

constexpr int HID = 512;          // rnn_dims == fc_dims
constexpr int NCTA = 128;         // CTAs; UNITS hidden units each
constexpr int UNITS = 4;
constexpr int BT = 8;             // folds per group (one 32-byte row of an exchanged vector)
constexpr int NTHREADS = 512;
constexpr int NWARPS = NTHREADS / 32;
constexpr int MAXG = 8;           // groups per launch -> 64 folds
constexpr int MAXT = 3;           // teams per CTA
constexpr int VEC = HID * BT;     // values of one exchanged vector (16 KiB in smem, 32 KiB of LL pairs in L2)
constexpr int ITEM = 4 * 32 * 4;  // floats of one weight item image: 4 rows x 128 k
constexpr int NEXCH = 5;
constexpr int CROW = 208;         // conditioning floats per fold-step: 80 mel + 4*32 aux
constexpr int COND_ITEMS = 12;
constexpr int PART_FLOATS = 24 * 32;   // per-team partial sums: 16 critical item slots + 8 deferred unit slots, x (4 rows x 8 folds)

// ---- per-CTA weight image (floats) --------------------------------------------------------
// An item image is [4 slots][32 lanes] float4: slot r of lane l holds W[row r ^ (l >> 3)][kbase + l + 32 i], i = 0..3
// (the per-lane row permutation makes the butterfly reduce select-free, see reduce_scatter32).
constexpr int W_M2 = 0;                       // 24 rows: Wih2x gates (12) | Whh1 gates (12), x H1
constexpr int W_M3 = W_M2 + 24 * HID;         // 16 rows: Wfc1x (4) x H1 (S2) and x H2 (S3) | Whh2 gates (12) x H2
constexpr int W_M4 = W_M3 + 16 * HID;         // 4 rows: Wfc2x x Y1
constexpr int W_M5 = W_M4 + 4 * HID;          // rows5 rows: Wfc3 x Y2
__host__ __device__ constexpr int w_mc(int rows5) { return W_M5 + rows5 * HID; }           // 12 cond items
__host__ __device__ constexpr int w_small(int rows5) { return w_mc(rows5) + COND_ITEMS * ITEM; }
// small vectors (offsets inside the small block)
constexpr int SV_U1 = 0, SV_B1 = 12, SV_BHH1 = 24, SV_U2 = 36, SV_B2 = 48, SV_BHH2 = 60,
              SV_U3 = 72, SV_B3 = 76, SV_B4 = 80, SV_B5 = 84, SV_SIZE = 128;
__host__ __device__ constexpr int w_total(int rows5) { return w_small(rows5) + SV_SIZE; }
// bf16-weight images (WRNN_PREC_BF16): every item image holds bf16 instead of fp32, i.e. all item offsets halve
// (they are multiples of ITEM); the small vectors stay fp32.  The arithmetic stays fp32 FFMA2.
__host__ __device__ constexpr int w_image_floats(int rows5, int bf16w) { return (bf16w ? w_small(rows5) / 2 : w_small(rows5)) + SV_SIZE; }

// ---- per-group private state (floats) -----------------------------------------------------
constexpr int PG_GH1 = 0, PG_GH2 = 96, PG_P1 = 192, PG_P2 = 288, PG_P3 = 384, PG_P4 = 416,
              PG_H1 = 448, PG_H2 = 480, PG_X = 512, PG_U = 520, PG_FX = 608, PG_F1 = 616, PG_SIZE = 656;

// ---- shared memory map (floats) -----------------------------------------------------------
// [weights | per-group state | fold ranges | profiling | team 0 | team 1 | team 2]
// team block: [stage | part | cond staging (nbuf x 8 folds x 208) | 16 control words]
struct SmemMap {
    int w, priv, fs, prof, team0, team_stride, stage_floats, nbuf, total;
    int t_stage, t_part, t_cst, t_ctl;      // offsets inside a team block
};
__host__ __device__ inline int stage_floats_for(int mode, int C) { return (mode == 0 && BT * C > VEC) ? BT * C : VEC; }
__host__ __device__ inline SmemMap smem_map(int rows5, int mode, int C, int T, int nbuf, int bf16w = 0)
{
    SmemMap m;
    m.w = 0;
    m.priv = m.w + w_image_floats(rows5, bf16w);
    m.fs = m.priv + MAXG * PG_SIZE;                 // [MAXG*8] fold starts | [MAXG*8] fold limits (long long)
    m.prof = m.fs + 4 * MAXG * BT;
    m.team0 = m.prof + 2 * 32;
    m.stage_floats = stage_floats_for(mode, C);
    m.nbuf = nbuf;
    m.t_stage = 0;
    m.t_part = m.t_stage + m.stage_floats;
    m.t_cst = m.t_part + PART_FLOATS;
    m.t_ctl = m.t_cst + nbuf * BT * CROW;           // [0..3] two mbarriers, [8] abort flag
    m.team_stride = m.t_ctl + 16;
    m.total = m.team0 + T * m.team_stride;
    return m;
}
__host__ __device__ constexpr int team_warps(int T) { return T == 1 ? 16 : T == 2 ? 8 : 5; }

struct KParams {
    const float *wimg;                 // [NCTA][w_total]
    const float *mels, *aux;           // unfolded conditioning [rows, 80] / [rows, 128]
    const long long *fold_start, *fold_limit;
    const float *uniforms, *forced_x;
    float *logits_out, *samples_out;
    int *labels_out;
    unsigned long long *xb;            // [G][xb_group] exchange buffers of {value, epoch} pairs
    int *status;
    unsigned long long seed;
    int B, S, G, C, mode, rows5, nprod5, n_u;   // n_u: uniforms per fold-step (1 RAW, 11 MOL)
    int feat, auxw;                    // 80, 128
    int T, nbuf;                       // teams per CTA, conditioning staging buffers per team
    int bf16w;                         // 1: item images hold bf16 weights (precision bf16), else fp32
    int group_fold0[MAXG], group_nf[MAXG];
    int probe_iters;
    long long *prof;                   // optional [NCTA][PROF_SLOTS] per-stage cycle counters (clock64, team 0 thread 0)
};
constexpr int PROF_SLOTS = 32;
// exchange buffer of one group, in pairs: H1 | H2 | Y1 | Y2 as [k][8 folds], logits as [8 folds][cpad]
constexpr int XB_H1 = 0, XB_H2 = VEC, XB_Y1 = 2 * VEC, XB_Y2 = 3 * VEC, XB_LG = 4 * VEC;
__host__ __device__ constexpr int xb_group(int cpad) { return 4 * VEC + cpad * BT; }

// ============================================================================================
// device helpers
// ============================================================================================
__device__ __forceinline__ uint4 ld_pairs2(const unsigned long long *p)      // two {value, epoch} pairs, L2 (never L1)
{
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint2 ld_pair(const unsigned long long *p)
{
    uint2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_pair(unsigned long long *p, float v, unsigned epoch)   // one 8-byte store
{
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(epoch) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_team(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// Shared-memory position of (k, fold f) of an exchanged vector: 32-byte rows [k][8]; fold f sits in slot
// f ^ (k & 3).  A lane working on k = lane + 32 i loads the two 16-byte halves of its row in the order
// (bit 2 of lane) first -- conflict-free -- and thereby finds fold (s ^ (lane & 7)) in register slot s.
__device__ __forceinline__ int xidx(int k, int f) { return k * BT + (f ^ (k & 3)); }

// mbarrier + 1-D bulk TMA (cp.async.bulk) -- conditioning rows are staged with these
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Philox4x32-10 (Salmon et al. 2011): counter-based RNG for the in-kernel uniforms
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll 1
    for (int i = 0; i < 10; ++i) {
        unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float u01(unsigned x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
// uniform j of (step, fold b): word j & 3 of Philox(counter = {step, b, j >> 2, 0}, key = seed).  Out of line: one
// warp per visit needs it and the step loop has to stay small.
__device__ __noinline__ float philox_uniform(unsigned long long seed, int step, int b, int j)
{
    const uint4 r = philox4x32_10(make_uint4((unsigned)step, (unsigned)b, (unsigned)(j >> 2), 0u), make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    return u01(((j & 3) == 0) ? r.x : ((j & 3) == 1) ? r.y : ((j & 3) == 2) ? r.z : r.w);
}

constexpr int POLL_CAP = 1 << 22;    // watchdog: ~1 s of polling

// ---- packed fp32 math (Blackwell FFMA2): d.{x,y} = a.{x,y} * b.{x,y} + d.{x,y}, each half rounded like fmaf ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void fma2(f32x2 &d, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }

// The four k-consecutive weights (k = kbase + lane + 32 i) of row slot r of an item image: one LDS.128 of fp32, or one
// LDS.64 of bf16 widened to fp32 (exact) when the image holds bf16 weights.
template <bool BF16W>
__device__ __forceinline__ float4 ld_w4(const float *wimg, int r, int lane)
{
    if (!BF16W) return *reinterpret_cast<const float4 *>(wimg + (r * 32 + lane) * 4);
    const uint2 u = *reinterpret_cast<const uint2 *>(reinterpret_cast<const unsigned short *>(wimg) + (r * 32 + lane) * 4);
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
}

// One work item: acc[slot r][fold slots] += W[4][128 k] * X[128 k][8 folds] for this lane's four k.
// wimg: item image (see above).  xs: exchanged vector in shared memory at row kbase (a multiple of 128).
// acc[r][j] holds fold slots (2j, 2j+1) of row slot r.
template <bool BF16W>
__device__ __forceinline__ void item_fma(const float *wimg, const float *xs, int lane, f32x2 (&acc)[4][4])
{
    float4 w[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) w[r] = ld_w4<BF16W>(wimg, r, lane);
    const int sw = ((lane >> 2) & 1) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float *xp = xs + (lane + 32 * i) * BT;
        const float4 lo = *reinterpret_cast<const float4 *>(xp + sw);        // fold slots 0..3
        const float4 hi = *reinterpret_cast<const float4 *>(xp + (4 - sw));  // fold slots 4..7
        const f32x2 x0 = pack2(lo.x, lo.y), x1 = pack2(lo.z, lo.w), x2 = pack2(hi.x, hi.y), x3 = pack2(hi.z, hi.w);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float wv = (i == 0) ? w[r].x : (i == 1) ? w[r].y : (i == 2) ? w[r].z : w[r].w;
            const f32x2 ww = pack2(wv, wv);
            fma2(acc[r][0], ww, x0);
            fma2(acc[r][1], ww, x1);
            fma2(acc[r][2], ww, x2);
            fma2(acc[r][3], ww, x3);
        }
    }
}
// The same item against the fold-major conditioning staging buffer cst[fold][208] (k = kbase + lane + 32 i;
// k >= 208 is the zero padding of the 256-wide conditioning K space).
template <bool BF16W>
__device__ __forceinline__ void item_fma_cond(const float *wimg, const float *cst, int kbase, int lane, f32x2 (&acc)[4][4])
{
    float4 w[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) w[r] = ld_w4<BF16W>(wimg, r, lane);
    // register slot s holds fold s ^ (lane & 7); its four k are kbase + lane + 32 i at fixed offsets from one pointer
    const float *row[BT];
#pragma unroll
    for (int s = 0; s < BT; ++s) row[s] = cst + (s ^ (lane & 7)) * CROW + kbase + lane;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool live = kbase + lane + 32 * i < CROW;            // only the tail of chunk B is padding
        float x[BT];
#pragma unroll
        for (int s = 0; s < BT; ++s) x[s] = live ? row[s][32 * i] : 0.f;
        const f32x2 x0 = pack2(x[0], x[1]), x1 = pack2(x[2], x[3]), x2 = pack2(x[4], x[5]), x3 = pack2(x[6], x[7]);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float wv = (i == 0) ? w[r].x : (i == 1) ? w[r].y : (i == 2) ? w[r].z : w[r].w;
            const f32x2 ww = pack2(wv, wv);
            fma2(acc[r][0], ww, x0);
            fma2(acc[r][1], ww, x1);
            fma2(acc[r][2], ww, x2);
            fma2(acc[r][3], ww, x3);
        }
    }
}

// Cross-lane reduce-scatter of the 32 accumulator slots.  Slot (r, f) of lane l holds the partial sum of
// (row r ^ (l >> 3), fold f ^ (l & 7)), so at every level both partners want to KEEP the lower half of their
// slots and SEND the upper half: 31 shuffles + 31 adds, no selects.  Lane l returns the warp-wide sum of
// (row l >> 3, fold l & 7).
__device__ __forceinline__ float reduce_scatter32(f32x2 (&acc)[4][4])
{
    float v[32];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) unpack2(acc[r][j], v[r * 8 + 2 * j], v[r * 8 + 2 * j + 1]);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
        for (int j = 0; j < off; ++j) v[j] = v[j] + __shfl_xor_sync(0xffffffffu, v[j + off], off);
    }
    return v[0];
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

// Publish this CTA's 4 x 8 values of one exchanged vector: 32 lanes x 8-byte pairs = 256 contiguous
// bytes.  One warp; lane = unit*8 + fold.  No fence, no flag: the epoch rides with the value.
__device__ __forceinline__ void publish_line(unsigned long long *vec, int cta, int lane, float val, unsigned epoch)
{
    st_pair(vec + (UNITS * cta) * BT + lane, val, epoch);
}

// Shared-memory position of class k in a fold's logits row (RAW sampler): lane k / NPL owns NPL consecutive
// classes; its float4 groups are XOR-swizzled so that the 16-byte reads of a quarter-warp hit distinct banks.
template <int NPL>
__device__ __forceinline__ int lg_swz(int owner)
{
    return NPL == 32 ? (owner & 7) : NPL == 16 ? ((owner >> 1) & 3) : NPL == 8 ? ((owner >> 2) & 1) : 0;
}
template <int NPL>
__device__ __forceinline__ int lg_pos(int k)
{
    if (NPL < 4) return k;
    const int owner = k / NPL, within = k % NPL;
    return owner * NPL + ((((within >> 2) ^ lg_swz<NPL>(owner)) << 2) | (within & 3));
}
__device__ __forceinline__ int lg_pos_dyn(int npl, int k)
{
    switch (npl) {
        case 32: return lg_pos<32>(k);
        case 16: return lg_pos<16>(k);
        case 8: return lg_pos<8>(k);
        case 4: return lg_pos<4>(k);
        default: return k;
    }
}

// ============================================================================================
// the persistent kernel
// ============================================================================================
struct Ctx {
    const KParams *p;
    float *sm;
    SmemMap m;
    int tid, lane, warp, cta;
    int C, rows5, nprod5, mode;       // model geometry: compile-time constants in the specialised kernels (MODEL 1, 2)
    int team, nw, nt, tw, ttid, ng;   // team index, warps / threads per team, warp / thread inside the team, groups of this team
    float *stage, *part, *cst;        // this team's staging buffers
    uint64_t *mbar;                   // this team's two conditioning mbarriers
    int *abort_flag;
    int cv_buf, cv_par;  // conditioning visit being consumed: staging buffer index, mbarrier phase parity
    int is_g, is_step, is_buf;   // next conditioning visit to issue (maintained by the issuing warp only)
    long long tprev;     // profiling: last timestamp (team 0, thread 0)
    float du[3], dfx;    // draws fetched by draws_issue, waiting for draws_commit (last warp of the team)
};
// profiling tick: charge the cycles since the previous tick to `slot` (thread 0 of team 0 only).
// Compiled out of the production kernel (PROF = false).
template <bool PROF>
__device__ __forceinline__ void tick(Ctx &c, int slot)
{
    if (PROF && c.tid == 0) {
        const long long now = clock64();
        reinterpret_cast<long long *>(c.sm + c.m.prof)[slot] += now - c.tprev;
        c.tprev = now;
    }
}

__device__ __forceinline__ float *priv(const Ctx &c, int g) { return c.sm + c.m.priv + g * PG_SIZE; }
__device__ __forceinline__ const float *small(const Ctx &c) { return c.sm + c.m.w + (w_image_floats(c.rows5, c.p->bf16w) - SV_SIZE); }
__device__ __forceinline__ unsigned long long *xb_base(const Ctx &c, int g)
{
    return c.p->xb + (size_t)g * xb_group(c.rows5 * c.nprod5);
}
__device__ __forceinline__ void team_sync(const Ctx &c) { bar_team(1 + c.team, c.nt); }

__device__ __forceinline__ void poll_timeout(Ctx &c)
{
    *c.abort_flag = 1;
    atomicExch(c.p->status, -4);
}

// LL gather of one exchanged vector (VEC {value, epoch} pairs, [k][8] in L2) into the team's staging
// buffer.  Every thread polls 16-byte chunks (two pairs) in batches of four until both epochs match.
// Team-uniform result: false = the watchdog fired somewhere in this team (the kernel then exits and the
// host reports WRNN_ERR_TIMEOUT).
template <bool PROF, int BATCH>
__device__ __forceinline__ bool gather_vec(Ctx &c, const unsigned long long *src, unsigned epoch, int prof_slot = 19)
{
    float *dst = c.stage;
    constexpr int NCH = VEC / 2;
    tick<PROF>(c, prof_slot + 4);                // everything between the previous phase and the first poll
#pragma unroll 1
    for (int base = c.ttid; base < NCH; base += BATCH * c.nt) {
        uint4 v[BATCH];
        bool need[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
            need[j] = base + j * c.nt < NCH;
            if (need[j]) v[j] = ld_pairs2(src + 2 * (base + j * c.nt));
        }
        for (int spin = 0;; ++spin) {
            bool bad = false;
#pragma unroll
            for (int j = 0; j < BATCH; ++j) {
                const bool b = need[j] && ((v[j].y != epoch) | (v[j].w != epoch));
                if (b) v[j] = ld_pairs2(src + 2 * (base + j * c.nt));
                bad |= b;
            }
            if (!bad) break;
            if (spin > POLL_CAP) {
                poll_timeout(c);
                break;
            }
        }
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
            if (!need[j]) continue;
            const int i = base + j * c.nt, k = i >> 2, f0 = (i & 3) * 2;
            // folds f0, f0+1 live in slots (f0 ^ (k&3)), (f0+1) ^ (k&3): an aligned pair, swapped when k is odd
            const float a = __uint_as_float(v[j].x), b = __uint_as_float(v[j].z);
            *reinterpret_cast<float2 *>(dst + k * BT + (f0 ^ (k & 2))) = (k & 1) ? make_float2(b, a) : make_float2(a, b);
        }
    }
    tick<PROF>(c, prof_slot);                    // own chunks validated + scattered
    team_sync(c);
    return *c.abort_flag == 0;
}

// Cold path of cond_issue: zero the staging rows of folds that have run past their conditioning (fold padding,
// fatchord_version.py:306-309) or do not exist.  Kept out of line so the step loop stays small.
__device__ __noinline__ void cond_zero_rows(float *buf, unsigned mask, int lane)
{
    for (int ff = 0; ff < BT; ++ff)
        if (!((mask >> ff) & 1))
            for (int i = lane; i < CROW; i += 32) buf[ff * CROW + i] = 0.f;
}

// The conditioning visits of a team run over (step, its groups) in that order; visit number v uses staging
// buffer v % nbuf.  Issue the TMA row copies of the next visit (is_g, is_step) and advance.  Called by ALL lanes
// of the team's last warp: lane f < 8 copies fold f's two rows (mel 320 B + aux 512 B).
__device__ __forceinline__ void cond_issue_next(Ctx &c)
{
    const KParams &p = *c.p;
    const int g = c.is_g, step = c.is_step, b = c.is_buf;
    c.is_g += p.T;
    if (c.is_g >= p.G) {
        c.is_g = c.team;
        c.is_step += 1;
    }
    c.is_buf = (b + 1 == p.nbuf) ? 0 : b + 1;
    if (step >= p.S) return;
    float *buf = c.cst + b * (BT * CROW);
    uint64_t *bar = c.mbar + b;
    const int f = c.lane;
    bool valid = false;
    long long row = 0;
    if (f < p.group_nf[g]) {
        const long long *fs = reinterpret_cast<const long long *>(c.sm + c.m.fs);   // [MAXG*8] starts | [MAXG*8] limits
        row = fs[g * BT + f] + step;
        valid = row < fs[MAXG * BT + g * BT + f];
    }
    const unsigned m = __ballot_sync(0xffffffffu, valid) & 0xffu;
    // folds that exist but have run past their conditioning read zeros (rows of folds that do not exist were zeroed
    // once in the prologue and are never written)
    const unsigned exist = (1u << p.group_nf[g]) - 1u;
    if ((m & exist) != exist) cond_zero_rows(buf, m | ~exist, c.lane);
    // the staging buffer was last read through the generic proxy (work items, before the team barrier)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (c.lane == 0) mbar_expect_tx(bar, (unsigned)(__popc(m) * (p.feat + p.auxw) * 4));
    __syncwarp();
    if (valid) {
        tma_bulk_g2s(buf + f * CROW, p.mels + row * p.feat, (unsigned)(p.feat * 4), bar);
        tma_bulk_g2s(buf + f * CROW + p.feat, p.aux + row * p.auxw, (unsigned)(p.auxw * 4), bar);
    }
}

// Wait until the conditioning rows of the visit being consumed have landed.  All threads of the team.
__device__ __forceinline__ void cond_wait(Ctx &c)
{
    uint64_t *bar = c.mbar + c.cv_buf;
    for (int spin = 0; !mbar_try_wait(bar, (unsigned)c.cv_par); ++spin) {
        if (spin > POLL_CAP) {                   // a conditioning copy never completed (bad pointer?): give up loudly
            poll_timeout(c);
            break;
        }
    }
}

// Fetch the uniforms / forced value that step `step` of group g will need at sampling time.
// Two-phase so the global-load latency (injected uniforms) overlaps the gather + items of the
// visit: issue early, commit later.  Same warp for both.
__device__ __forceinline__ void draws_issue(Ctx &c, int g, int step)
{
    const KParams &p = *c.p;
    const int nu = p.n_u;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int i = c.lane + 32 * q;
        float u = 0.f;
        if (i < BT * nu) {
            const int f = i / nu, j = i - f * nu;
            if (f < p.group_nf[g]) {
                const int b = p.group_fold0[g] + f;
                u = p.uniforms ? p.uniforms[((size_t)step * p.B + b) * nu + j] : philox_uniform(p.seed, step, b, j);
            }
        }
        c.du[q] = u;
    }
    c.dfx = 0.f;
    if (p.forced_x && c.lane < p.group_nf[g]) c.dfx = p.forced_x[(size_t)step * p.B + p.group_fold0[g] + c.lane];
}
__device__ __forceinline__ void draws_commit(Ctx &c, int g)
{
    float *pg = priv(c, g);
    const int nu = c.p->n_u;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int i = c.lane + 32 * q;
        if (i < BT * nu) {
            const int f = i / nu, j = i - f * nu;
            pg[PG_U + f * 11 + j] = c.du[q];
        }
    }
    if (c.lane < BT) pg[PG_FX + c.lane] = c.dfx;
}

// ---- sampling: logits of step s are polled straight from L2 ([fold][class] LL pairs); writes the fed-back x ----
// RAW: softmax (fatchord_version.py:211) + inverse CDF with one uniform per fold (oracle/ref_shim.py:
// k = #{c : cdf_c <= u}, clamped) + label -> float (:214).  One warp per fold (a warp takes folds tw, tw + nw, ...),
// NPL = C/32 consecutive classes per lane, warp shuffles only (no block barrier before the sample exists).
template <int NPL>
__device__ __forceinline__ void sample_raw(Ctx &c, int g, int s, unsigned epoch)
{
    const KParams &p = *c.p;
    float *pg = priv(c, g);
    const int lane = c.lane;
    constexpr int C = NPL * 32;
    constexpr int NCH = (NPL / 2) > 0 ? NPL / 2 : 1;        // 16-byte chunks (2 classes) per lane
    constexpr int BATCH = NPL == 16 ? 8 : (NCH < 4 ? NCH : 4);   // 512 classes: all chunks of a fold in flight (one round trip)
    for (int f = c.tw; f < BT; f += c.nw) {
        const unsigned long long *src = xb_base(c, g) + XB_LG + (size_t)f * C;
        float *row = c.stage + f * C;
#pragma unroll 1
        for (int j0 = 0; j0 < NCH; j0 += BATCH) {
            uint4 v[BATCH];
#pragma unroll
            for (int j = 0; j < BATCH; ++j) v[j] = ld_pairs2(src + 2 * ((j0 + j) * 32 + lane));
            for (int spin = 0;; ++spin) {
                bool bad = false;
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    const bool b = (v[j].y != epoch) | (v[j].w != epoch);
                    if (b) v[j] = ld_pairs2(src + 2 * ((j0 + j) * 32 + lane));
                    bad |= b;
                }
                if (!bad) break;
                if (spin > POLL_CAP) {
                    poll_timeout(c);
                    break;
                }
            }
#pragma unroll
            for (int j = 0; j < BATCH; ++j)
                *reinterpret_cast<float2 *>(row + lg_pos<NPL>(2 * ((j0 + j) * 32 + lane))) =
                    make_float2(__uint_as_float(v[j].x), __uint_as_float(v[j].z));
        }
        __syncwarp();
        float v[NPL];
        if (NPL >= 4) {
#pragma unroll
            for (int j = 0; j < NPL; j += 4) {
                const float4 q = *reinterpret_cast<const float4 *>(row + lane * NPL + (((j >> 2) ^ lg_swz<NPL>(lane)) << 2));
                v[j] = q.x; v[j + 1] = q.y; v[j + 2] = q.z; v[j + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < NPL; ++j) v[j] = row[lane * NPL + j];
        }
        float m = v[0];
#pragma unroll
        for (int j = 1; j < NPL; ++j) m = fmaxf(m, v[j]);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
        float run = 0.f;                              // inclusive prefix inside the lane
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            run += expf(v[j] - m);
            v[j] = run;
        }
        float incl = run;                             // inclusive scan of the lane totals across the warp
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const float t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        const float excl = incl - run;
        const float total = __shfl_sync(0xffffffffu, incl, 31);
        const float thr = pg[PG_U + f * 11] * total;
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < NPL; ++j) cnt += (excl + v[j] <= thr) ? 1 : 0;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        if (lane == 0) {
            const int k = cnt > C - 1 ? C - 1 : cnt;
            // 2 * k.float() / (C - 1.) - 1.  (fatchord_version.py:214), three separately rounded fp32 ops
            const float sample = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, (float)k), (float)C - 1.0f), 1.0f);
            if (f < p.group_nf[g] && c.cta == (s * p.G + g) % NCTA) {
                const int b = p.group_fold0[g] + f;
                p.samples_out[(size_t)b * p.S + s] = sample;
                if (p.labels_out) p.labels_out[(size_t)b * p.S + s] = k;
            }
            pg[PG_X + f] = p.forced_x ? pg[PG_FX + f] : sample;
        }
    }
}

// sample_from_discretized_mix_logistic, utility/distribution.py:87-123.  Warp 0 of the team; the 30 logits of
// the 8 folds are polled into stage[i * 8 + f].
__device__ __forceinline__ void sample_mol(Ctx &c, int g, int s, unsigned epoch)
{
    const KParams &p = *c.p;
    if (c.tw != 0) return;
    float *lg = c.stage;
    float *pg = priv(c, g);
    const int lane = c.lane, f = lane & 7, cs = lane >> 3;
    const int nr = c.C / 3, cpad = c.rows5 * c.nprod5;
    const unsigned long long *src = xb_base(c, g) + XB_LG + (size_t)f * cpad;
    for (int i = cs; i < c.C; i += 4) {
        uint2 v = ld_pair(src + i);
        for (int spin = 0; v.y != epoch; ++spin) {
            if (spin > POLL_CAP) {
                poll_timeout(c);
                break;
            }
            v = ld_pair(src + i);
        }
        lg[i * BT + f] = __uint_as_float(v.x);
    }
    __syncwarp();
    float best = -INFINITY;
    int arg = cs;                   // a valid index even when every comparison fails (NaN logits)
    for (int i = cs; i < nr; i += 4) {
        const float u = (float)(1e-5 + ((1.0 - 1e-5) - 1e-5) * (double)pg[PG_U + f * 11 + i]);
        const float t = lg[i * BT + f] - logf(-logf(u));
        if (t > best) {
            best = t;
            arg = i;
        }
    }
#pragma unroll
    for (int off = 8; off <= 16; off <<= 1) {
        const float b2 = __shfl_xor_sync(0xffffffffu, best, off);
        const int a2 = __shfl_xor_sync(0xffffffffu, arg, off);
        if (b2 > best || (b2 == best && a2 < arg)) {
            best = b2;
            arg = a2;
        }
    }
    if (cs == 0) {
        const float mean = lg[(nr + arg) * BT + f];
        const float ls = fmaxf(lg[(2 * nr + arg) * BT + f], -32.23619130191664f);
        const float u2 = (float)(1e-5 + ((1.0 - 1e-5) - 1e-5) * (double)pg[PG_U + f * 11 + nr]);
        float x = mean + expf(ls) * (logf(u2) - logf(1.0f - u2));
        x = fminf(fmaxf(x, -1.0f), 1.0f);
        if (f < p.group_nf[g] && c.cta == (s * p.G + g) % NCTA) {
            const int b = p.group_fold0[g] + f;
            p.samples_out[(size_t)b * p.S + s] = x;
            if (p.labels_out) p.labels_out[(size_t)b * p.S + s] = arg;
        }
        pg[PG_X + f] = p.forced_x ? pg[PG_FX + f] : x;
    }
}

// logits_out[s][b][c] from the staged logits (teacher-forced parity runs only; after the team barrier)
__device__ __forceinline__ void dump_logits(Ctx &c, int g, int s)
{
    const KParams &p = *c.p;
    if (c.cta != (s * p.G + g) % NCTA) return;
    const float *lg = c.stage;
    const int npl = c.mode == 0 ? c.C >> 5 : 0;
    for (int f = 0; f < p.group_nf[g]; ++f) {
        float *dst = p.logits_out + ((size_t)s * p.B + p.group_fold0[g] + f) * c.C;
        for (int k = c.ttid; k < c.C; k += c.nt) dst[k] = npl ? lg[f * c.C + lg_pos_dyn(npl, k)] : lg[k * BT + f];
    }
}

// The mat-vec work of a stage runs in two phases.
//  CRITICAL: only what the value published by this stage needs, cut into single ITEMS (4 rows x 8 folds x 128 k,
//    one reduce-scatter each) so that as many warps as possible shorten the dependent chain:
//      S2 12 items Wih2x{r,z,n} x 4 K chunks | S3 4 items Wfc1x | S4 4 items Wfc2x | S5 rows5 items Wfc3
//    Item i leaves its partial sums in part[i * 32 + lane] (lane = row * 8 + fold); the finalize adds the four
//    K chunks in a fixed order.
//  DEFERRED: everything whose result is only needed later, run AFTER the stage has published (i.e. inside the
//    exchange latency), cut into UNITS of two consecutive K chunks with ONE reduce-scatter (the shuffles share
//    the shared-memory pipe with the operand loads, so they are worth saving):
//      S2 8 units {Whh1 r,z,n | Wfc1x} x 2 K halves (next step's gh1, this step's fc1 input)
//      S3 6 units Whh2{r,z,n} x 2 (next step's gh2)
//      S4 8 conditioning units: P1 r,z,n (item 0,1,2; chunk A) | P2 r,z,n (items 3+2q, 4+2q; A,B) | P3 (9,10; A,B) | P4 (11; B)
//    Unit u leaves its sums in part[(DEF0 + u) * 32 + lane].
// Work is dealt round-robin to the warps of the team and combined in a fixed order, so the arithmetic does not
// depend on the team size.  One copy of the mat-vec code serves every stage (instruction cache).
constexpr int DEF0 = 16;    // first deferred slot of the partial-sum buffer

__device__ __forceinline__ void zero_acc(f32x2 (&acc)[4][4])
{
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[r][j] = 0ull;
}

template <bool BF16W>
__device__ __forceinline__ void run_critical(Ctx &c, int stage)
{
    constexpr int SH = BF16W ? 1 : 0;            // item offsets halve when the images hold bf16
    const int rows5 = c.rows5;
    const int n = stage == 1 ? 12 : stage == 4 ? rows5 : 4;
    const int woff = stage == 1 ? W_M2 : stage == 2 ? W_M3 : stage == 3 ? W_M4 : W_M5;
    const float *W = c.sm + c.m.w + (woff >> SH);
#pragma unroll 1
    for (int i = c.tw; i < n; i += c.nw) {
        f32x2 acc[4][4];
        zero_acc(acc);
        item_fma<BF16W>(W + ((i * ITEM) >> SH), c.stage + (i & 3) * 128 * BT, c.lane, acc);
        c.part[i * 32 + c.lane] = reduce_scatter32(acc);
    }
}

template <bool BF16W>
__device__ __forceinline__ void run_deferred(Ctx &c, int stage)
{
    constexpr int SH = BF16W ? 1 : 0;
    const int rows5 = c.rows5;
    const int n = stage == 1 ? 8 : stage == 2 ? 6 : 8;
    const float *W = c.sm + c.m.w;
    // deal from warp 1 on: warp 0 is busy with the pointwise math + publish of this stage
    int first = c.tw - 1;
    if (first < 0) first += c.nw;
#pragma unroll 1
    for (int u = first; u < n; u += c.nw) {
        f32x2 acc[4][4];
        zero_acc(acc);
        if (stage == 3) {
            const int it0 = u < 3 ? u : u < 6 ? 3 + 2 * (u - 3) : u == 6 ? 9 : 11;
            const int nit = (u < 3 || u == 7) ? 1 : 2;
            const int chunk0 = u == 7 ? 1 : 0;
            const float *cst = c.cst + c.cv_buf * (BT * CROW);
#pragma unroll 1
            for (int q = 0; q < nit; ++q) item_fma_cond<BF16W>(W + ((w_mc(rows5) + (it0 + q) * ITEM) >> SH), cst, (chunk0 + q) * 128, c.lane, acc);
        } else {
            const int rg = u >> 1, h = u & 1;
            // S2: row groups 3..5 of M2 (Whh1), then Wfc1x = row group 0 of M3;  S3: row groups 1..3 of M3 (Whh2)
            const int woff = stage == 1 ? (rg < 3 ? W_M2 + ((3 + rg) * 4 + 2 * h) * ITEM : W_M3 + 2 * h * ITEM)
                                        : W_M3 + ((1 + rg) * 4 + 2 * h) * ITEM;
#pragma unroll
            for (int q = 0; q < 2; ++q) item_fma<BF16W>(W + ((woff + q * ITEM) >> SH), c.stage + (2 * h + q) * 128 * BT, c.lane, acc);
        }
        c.part[(DEF0 + u) * 32 + c.lane] = reduce_scatter32(acc);
    }
}
__device__ __forceinline__ float sum4(const float *part, int rg, int lane)
{
    const float *q = part + rg * 128 + lane;
    return (q[0] + q[32]) + (q[64] + q[96]);
}
__device__ __forceinline__ float sum2d(const float *part, int pair, int lane)      // deferred units 2*pair, 2*pair+1
{
    return part[(DEF0 + 2 * pair) * 32 + lane] + part[(DEF0 + 2 * pair + 1) * 32 + lane];
}

// GRU cell of this CTA's 4 units x 8 folds (torch gate order r, z, n) + publish of the new state.  One warp.
// which = 0: rnn1 (input side folded into P1; gate pre-activations need no mat-vec), 1: rnn2 (input side =
// Wih2x . h1 from the critical items + P2).
template <bool PROF>
__device__ __forceinline__ void gru_publish(Ctx &c, float *pg, unsigned long long *xb, const float *sv, const float *part, int which, unsigned epoch)
{
    const int lane = c.lane, u = lane >> 3, f = lane & 7;
    const float x = pg[PG_X + f];
    const int P = which == 0 ? PG_P1 : PG_P2, GH = which == 0 ? PG_GH1 : PG_GH2, H = which == 0 ? PG_H1 : PG_H2;
    const int SU = which == 0 ? SV_U1 : SV_U2, SB = which == 0 ? SV_B1 : SV_B2;
    float gi[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        gi[q] = pg[P + q * 32 + lane] + x * sv[SU + q * 4 + u] + sv[SB + q * 4 + u];
        if (which == 1) gi[q] += sum4(part, q, lane);
    }
    const float r = sigmoidf_(gi[0] + pg[GH + lane]);
    const float z = sigmoidf_(gi[1] + pg[GH + 32 + lane]);
    const float n = tanhf(gi[2] + r * pg[GH + 64 + lane]);
    const float h = (1.0f - z) * n + z * pg[H + lane];
    pg[H + lane] = h;
    publish_line(xb + (which == 0 ? XB_H1 : XB_H2), c.cta, lane, h, epoch);
    tick<PROF>(c, which == 0 ? 2 : 5);
}

// One visit: stage `stage` of step t for group g.  Returns false when the watchdog fired (team-uniform).
template <bool PROF, bool BF16W, int BATCH>
__device__ __forceinline__ bool visit(Ctx &c, int t, int stage, int g)
{
    const KParams &p = *c.p;
    const int S = p.S, lane = c.lane, tw = c.tw, nw = c.nw;
    const unsigned epoch = (unsigned)t + 1u;
    float *pg = priv(c, g);
    unsigned long long *xb = xb_base(c, g);
    const float *sv = small(c);
    const float *part = c.part;
    const bool warm = t < 0;
    // The refill of a conditioning staging buffer costs its issuing warp ~1400 cycles; unless this team's next visits
    // need the buffer before S2 comes round again (more groups than buffers) it is postponed to S2's deferred phase.
    const bool late_issue = c.ng <= p.nbuf;

    if (warm) team_sync(c);                              // no gather barrier in the warm-up pass: part is reused
    if (stage == 0) {
        if (t > 0) {
            // sample step t-1 from its logits (epoch t).  The barrier orders the draws committed at S4 (CTAs that
            // do not produce logits come here straight from S4's finalize).
            team_sync(c);
            if (c.mode != 0) sample_mol(c, g, t - 1, (unsigned)t);
            else switch (c.C) {
                case 1024: sample_raw<32>(c, g, t - 1, (unsigned)t); break;
                case 512: sample_raw<16>(c, g, t - 1, (unsigned)t); break;
                case 256: sample_raw<8>(c, g, t - 1, (unsigned)t); break;
                case 128: sample_raw<4>(c, g, t - 1, (unsigned)t); break;
                default: sample_raw<2>(c, g, t - 1, (unsigned)t); break;
            }
            team_sync(c);
            if (*c.abort_flag) return false;
            tick<PROF>(c, 1);
            if (p.logits_out) {
                dump_logits(c, g, t - 1);
                team_sync(c);                            // the staging buffer is reused by the next gather
            }
        }
        if (t == S) return true;
    }

    // stages 1..4: gather + critical items, then pointwise math + publish, then the deferred units inside the
    // exchange latency of the value just published (SA has no mat-vec work: GRU1's input side is all precomputed).
    const bool cond_live = t + 1 < S;
    const int rows5 = c.rows5;
    if (!warm && stage > 0) {
        if (!gather_vec<PROF, BATCH>(c, xb + (stage - 1) * VEC, epoch, 19 + stage)) return false;   // H1 | H2 | Y1 | Y2
        tick<PROF>(c, 3 * stage + (stage >= 3 ? 1 : 0));
        run_critical<BF16W>(c, stage);
        team_sync(c);
        tick<PROF>(c, 3 * stage + 1 + (stage >= 3 ? 1 : 0));
    }
    if (!warm) {
        if (stage <= 1) {
            if (tw == 0) gru_publish<PROF>(c, pg, xb, sv, part, stage, epoch);
            if (stage == 0) return true;
        } else if (stage == 2) {
            if (tw == 0) {       // fc1: Wfc1x . (h1 + h2) with the h1 half saved by S2's deferred phase
                const int u = lane >> 3, f = lane & 7;
                float y = (sum4(part, 0, lane) + pg[PG_F1 + lane]) + pg[PG_P3 + lane] + pg[PG_X + f] * sv[SV_U3 + u] + sv[SV_B3 + u];
                y = fmaxf(y, 0.f);
                publish_line(xb + XB_Y1, c.cta, lane, y, epoch);
                tick<PROF>(c, 8);
            }
        } else if (stage == 3) {
            if (tw == 0) {       // fc2
                float y = sum4(part, 0, lane) + pg[PG_P4 + lane] + sv[SV_B4 + (lane >> 3)];
                y = fmaxf(y, 0.f);
                publish_line(xb + XB_Y2, c.cta, lane, y, epoch);
                tick<PROF>(c, 12);
            }
        } else {
            // S5: logits of this CTA's rows5 classes, published fold-major
            for (int role = tw; role < (rows5 >> 2); role += nw) {
                const float v = sum4(part, role, lane) + sv[SV_B5 + role * 4 + (lane >> 3)];
                const int k = rows5 * c.cta + role * 4 + (lane >> 3);
                st_pair(xb + XB_LG + (size_t)(lane & 7) * (rows5 * c.nprod5) + k, v, epoch);
            }
            tick<PROF>(c, 15);
            return true;
        }
    }
    if (stage == 1 && late_issue && tw == nw - 1) cond_issue_next(c);
    if (stage == 3) {
        if (cond_live) cond_wait(c);                     // conditioning rows of step t+1 of this group
        if (tw == nw - 1 && !warm) draws_issue(c, g, t); // draws consumed by the sample of step t (at SA of t+1)
    }
    if (stage != 3 || cond_live) run_deferred<BF16W>(c, stage);
    team_sync(c);
    if (*c.abort_flag) return false;
    tick<PROF>(c, 16 + (stage - 1));
    if (stage == 3) {
        // the staging buffer of this conditioning visit is free again: refill it for visit v + nbuf
        if (tw == nw - 1 && !late_issue) cond_issue_next(c);
        if (++c.cv_buf == p.nbuf) {
            c.cv_buf = 0;
            c.cv_par ^= 1;
        }
    }
    const int nroles = stage == 1 ? 4 : stage == 2 ? 3 : 8;
#pragma unroll 1
    for (int role = tw; role < nroles; role += nw) {
        if (stage == 1) {
            if (role < 3)        // gh1 of the NEXT step: Whh1 . h1_t + b_hh1
                pg[PG_GH1 + role * 32 + lane] = sum2d(part, role, lane) + sv[SV_BHH1 + role * 4 + (lane >> 3)];
            else                 // Wfc1x . h1_t, consumed by S3 of this step
                pg[PG_F1 + lane] = sum2d(part, 3, lane);
        } else if (stage == 2) {
            pg[PG_GH2 + role * 32 + lane] = sum2d(part, role, lane) + sv[SV_BHH2 + role * 4 + (lane >> 3)];
        } else if (cond_live) {
            // conditioning projections of step t+1: units 0-2 P1 | 3-5 P2 | 6 P3 | 7 P4
            const int dst = role < 3 ? PG_P1 + role * 32 : role < 6 ? PG_P2 + (role - 3) * 32 : role == 6 ? PG_P3 : PG_P4;
            pg[dst + lane] = part[(DEF0 + role) * 32 + lane];
        }
    }
    if (stage == 3 && tw == nw - 1 && !warm) draws_commit(c, g);
    return true;
}

// TEAMS is a compile-time constant (one kernel per team count): warps / threads per team fold into immediates
// and the gather can keep the right number of polls in flight.
template <bool PROF, bool BF16W, int TEAMS, int MODEL>
__device__ __forceinline__ void persistent_body(const KParams &prm)
{
    extern __shared__ __align__(128) float sm[];
    const KParams &p = prm;
    Ctx c;
    c.p = &prm;
    c.sm = sm;
    // MODEL 1 = RAW with 512 classes, 2 = MOL (30 logits): geometry folds into immediates; 0 = any supported model
    c.C = MODEL == 1 ? 512 : MODEL == 2 ? 30 : p.C;
    c.rows5 = MODEL ? 4 : p.rows5;
    c.nprod5 = MODEL == 1 ? 128 : MODEL == 2 ? 8 : p.nprod5;
    c.mode = MODEL == 1 ? 0 : MODEL == 2 ? 1 : p.mode;
    c.m = smem_map(c.rows5, c.mode, c.C, TEAMS, p.nbuf, BF16W ? 1 : 0);
    c.tid = threadIdx.x;
    c.lane = c.tid & 31;
    c.warp = c.tid >> 5;
    c.cta = blockIdx.x;
    c.nw = team_warps(TEAMS);
    c.nt = c.nw * 32;
    constexpr int BATCH = TEAMS == 1 ? 4 : TEAMS == 2 ? 8 : 7;
    c.team = c.warp / c.nw;
    c.tw = c.warp - c.team * c.nw;
    c.ttid = c.tid - c.team * c.nt;
    c.cv_buf = c.cv_par = 0;
    c.is_step = c.is_buf = 0;
    c.tprev = 0;
    const bool member = c.team < TEAMS;
    c.ng = member ? (p.G - c.team + TEAMS - 1) / TEAMS : 0;
    c.is_g = c.team;
    {
        float *tb = sm + c.m.team0 + (member ? c.team : 0) * c.m.team_stride;
        c.stage = tb + c.m.t_stage;
        c.part = tb + c.m.t_part;
        c.cst = tb + c.m.t_cst;
        c.mbar = reinterpret_cast<uint64_t *>(tb + c.m.t_ctl);
        c.abort_flag = reinterpret_cast<int *>(tb + c.m.t_ctl + 8);
    }
    const int G = p.G, S = p.S;

    // ---- prologue (whole CTA): resident weights, zero state ------------------------------------
    {
        const int wfloats = w_image_floats(c.rows5, BF16W ? 1 : 0);
        const float4 *src = reinterpret_cast<const float4 *>(p.wimg + (size_t)c.cta * wfloats);
        float4 *dst = reinterpret_cast<float4 *>(sm + c.m.w);
        for (int i = c.tid; i < wfloats / 4; i += NTHREADS) dst[i] = src[i];
        for (int i = c.tid; i < MAXG * PG_SIZE; i += NTHREADS) sm[c.m.priv + i] = 0.f;
        if (c.tid < MAXG * BT) {                 // fold row ranges cached in shared memory (cond_issue reads them every visit)
            const int g = c.tid / BT, f = c.tid % BT;
            long long *fs = reinterpret_cast<long long *>(sm + c.m.fs);
            const bool live = g < p.G && f < p.group_nf[g];
            fs[c.tid] = live ? p.fold_start[p.group_fold0[g] + f] : 0;
            fs[MAXG * BT + c.tid] = live ? p.fold_limit[p.group_fold0[g] + f] : 0;
        }
        if (c.tid < TEAMS) {
            float *tb = sm + c.m.team0 + c.tid * c.m.team_stride;
            uint64_t *bar = reinterpret_cast<uint64_t *>(tb + c.m.t_ctl);
            mbar_init(bar, 1);
            mbar_init(bar + 1, 1);
            *reinterpret_cast<int *>(tb + c.m.t_ctl + 8) = 0;
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (c.tid < 2 * PROF_SLOTS) sm[c.m.prof + c.tid] = 0.f;
        for (int tm = 0; tm < TEAMS; ++tm)         // conditioning staging starts as zeros (rows of absent folds stay so)
            for (int i = c.tid; i < p.nbuf * BT * CROW; i += NTHREADS) sm[c.m.team0 + tm * c.m.team_stride + c.m.t_cst + i] = 0.f;
        __syncthreads();
        const float *sv = small(c);
        for (int g = 0; g < G; ++g) {           // h = 0  =>  gh = b_hh  (fatchord_version.py:173-174)
            float *pg = priv(c, g);
            if (c.tid < 96) {
                pg[PG_GH1 + c.tid] = sv[SV_BHH1 + (c.tid >> 3)];
                pg[PG_GH2 + c.tid] = sv[SV_BHH2 + (c.tid >> 3)];
            }
        }
        __syncthreads();
    }
    if (!member) return;                         // spare warp (3 teams x 5 warps)
    if (c.tw == c.nw - 1)
        for (int v = 0; v < p.nbuf; ++v) cond_issue_next(c);

    if (PROF && c.tid == 0) c.tprev = clock64();
    // t = -1 is the warm-up pass: only the conditioning half of stage 3 runs (projections of step 0).
    // Visit order inside a team: stage-major, its groups inside each stage, so that with several groups
    // per team one group's exchange still overlaps the others' mat-vecs.
    for (int t = -1; t <= S; ++t) {
        for (int stage = (t < 0 ? 3 : 0); stage < (t < 0 ? 4 : 5); ++stage) {
            if (stage == 4 && c.cta >= c.nprod5) break;            // only the logits producers run S5
            if (stage > 0 && t == S) break;
            for (int g = c.team; g < G; g += TEAMS)
                if (!visit<PROF, BF16W, BATCH>(c, t, stage, g)) return;
        }
    }
    if (PROF && p.prof && c.tid == 0)
        for (int i = 0; i < PROF_SLOTS; ++i) p.prof[(size_t)c.cta * PROF_SLOTS + i] = reinterpret_cast<long long *>(sm + c.m.prof)[i];
}

#ifndef WRNN_HELPERS_ONLY   /* development: a translation unit that only wants the device helpers above */
#define WRNN_KERNEL(name, PROF, BF16W, TEAMS, MODEL) \
    extern "C" __global__ void __launch_bounds__(NTHREADS, 1) name(const KParams prm) { persistent_body<PROF, BF16W, TEAMS, MODEL>(prm); }
#define WRNN_KERNELS3(stem, PROF, BF16W, MODEL) \
    WRNN_KERNEL(stem, PROF, BF16W, 1, MODEL) WRNN_KERNEL(stem##_t2, PROF, BF16W, 2, MODEL) WRNN_KERNEL(stem##_t3, PROF, BF16W, 3, MODEL)
// One kernel per (teams per CTA) x (model geometry: RAW-512 | MOL | any) x (fp32 | bf16 weight images); the generic
// fp32 kernels also exist with the per-stage clock64 accounting compiled in (wrnn_set_profiling).
WRNN_KERNELS3(wavernn_persistent_kernel, false, false, 1)
WRNN_KERNELS3(wavernn_persistent_kernel_mol, false, false, 2)
WRNN_KERNELS3(wavernn_persistent_kernel_any, false, false, 0)
WRNN_KERNELS3(wavernn_persistent_kernel_bf16w, false, true, 1)
WRNN_KERNELS3(wavernn_persistent_kernel_mol_bf16w, false, true, 2)
WRNN_KERNELS3(wavernn_persistent_kernel_any_bf16w, false, true, 0)
WRNN_KERNELS3(wavernn_persistent_kernel_prof, true, false, 0)
#undef WRNN_KERNELS3
#undef WRNN_KERNEL

// Exchange microbenchmark: the same publish / LL-gather sequence on an otherwise empty kernel (one team).
extern "C" __global__ void __launch_bounds__(NTHREADS, 1) wavernn_exchange_probe_kernel(const KParams prm)
{
    extern __shared__ __align__(128) float sm[];
    Ctx c;
    c.p = &prm;
    c.sm = sm;
    c.C = prm.C;
    c.rows5 = prm.rows5;
    c.nprod5 = prm.nprod5;
    c.mode = prm.mode;
    c.m = smem_map(prm.rows5, prm.mode, prm.C, 1, 1, prm.bf16w);
    c.tid = threadIdx.x;
    c.lane = c.tid & 31;
    c.warp = c.tid >> 5;
    c.cta = blockIdx.x;
    c.team = 0;
    c.nw = NWARPS;
    c.nt = NTHREADS;
    c.tw = c.warp;
    c.ttid = c.tid;
    c.tprev = 0;
    float *tb = sm + c.m.team0;
    c.stage = tb + c.m.t_stage;
    c.abort_flag = reinterpret_cast<int *>(tb + c.m.t_ctl + 8);
    if (c.tid == 0) *c.abort_flag = 0;
    __syncthreads();
    unsigned long long *xb = xb_base(c, 0);
    float acc = 0.f;
    for (int it = 0; it < prm.probe_iters; ++it) {
        unsigned long long *vec = xb + (it & 3) * VEC;
        const unsigned epoch = (unsigned)it + 1u;
        if (c.warp == 0) publish_line(vec, c.cta, c.lane, acc + (float)it, epoch);
        if (!gather_vec<false, 4>(c, vec, epoch)) return;
        acc += c.stage[c.tid] * 1e-30f;
        __syncthreads();
    }
    if (acc == 123.456f) prm.status[1] = 1;      // keep the loads alive
}
#endif  // WRNN_HELPERS_ONLY

}  // namespace wrnn
