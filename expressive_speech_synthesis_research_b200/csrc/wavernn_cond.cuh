// MelResNet at frame rate (reference: WaveRNN/models/fatchord_version.py:28-45, ResBlock :10-25), the `aux` half of
// UpsampleNetwork.forward (:79-86) before the repeat.  One launch serves every utterance of a pooled call: the grid runs over
// tiles of 64 output frames, a table gives each tile its rows.
//
//   x0  = relu(bn0(conv_in(mel)))                       conv_in: Conv1d(80 -> 128, k = 5, no bias, no padding: T + 4 rows -> T)
//   x   = x + bn2(conv2(relu(bn1(conv1(x)))))            x res_blocks, 1x1 convolutions without bias
//   aux = conv_out(x) + b                                1x1
//
// Every output (frame, channel) is one sequential fp32 dot product over k in a fixed order, so a frame's result does not depend on
// the tile or the launch it is computed in (pooled and single-utterance calls are bit-identical, tests/test_gpu_dense.py), which
// per-utterance cuDNN calls do not promise across shapes; and a pooled call costs one launch instead of ~65 per utterance (the
// host launches of 32 utterances were 58 ms of a 290 ms call on 8 GPUs).  Eval-mode batch norm is folded to a scale and a shift
// per channel on the host in float64 (WaveRNN.pack_melresnet); the convolution sums themselves are untouched.
//
// Layout of the weight blob (floats): W0 [5 taps][80 in][128 out] | s0 [128] | b0 [128] | res_blocks x { W1 [128 k][128 out] | s1 | b1 |
// W2 [128 k][128 out] | s2 | b2 } | Wout [128 k][128 out] | bout [128].  k-major so that the 16 lanes (8 outputs each) of a warp read
// 512 contiguous bytes of shared memory per k.
#pragma once
#include <cuda_runtime.h>

namespace wrnn_mel {

constexpr int CD = 128;                 // compute_dims = res_out_dims (hparams.py:37-38; both geometries)
constexpr int FEAT = 80, KS = 5;
constexpr int TF = 64;                  // output frames per CTA
constexpr int NT = 256;                 // thread = (og = tid % 16: outputs 8 og .. 8 og + 7, fg = tid / 16: frames 4 fg .. 4 fg + 3)
constexpr int LAYER_FLOATS = CD * CD + 2 * CD;
constexpr int W0_FLOATS = KS * FEAT * CD + 2 * CD;
__host__ __device__ constexpr long long blob_floats(int res_blocks) { return W0_FLOATS + (long long)(2 * res_blocks) * LAYER_FLOATS + CD * CD + CD; }

// shared memory (floats)
constexpr int SM_WL = 0;                          // weights of the layer in flight [128 k][128 out] (conv_in: one tap [80][128])
constexpr int SM_X = SM_WL + CD * CD;             // state [128 ch][64 frames]
constexpr int SM_Y = SM_X + CD * TF;              // hidden [128][64]; conv_in: the tile's input rows [68][80]
constexpr int SM_FLOATS = SM_Y + CD * TF;
constexpr int SM_BYTES = SM_FLOATS * 4;           // 128 KB
static_assert((TF + KS - 1) * FEAT <= CD * TF, "the input rows of a tile fit the hidden buffer");

struct Tile { int mel_row0, mel_rows, out_row0, nvalid; };    // input rows [mel_row0, +mel_rows) exist (rest reads zero); outputs [out_row0, +nvalid)

struct MParams {
    const float *blob;
    const float *mel;                  // [rows][80] zero-padded mel frames of all segments
    float *aux;                        // [rows][128]
    const Tile *tiles;
    int res_blocks;
};

__device__ __forceinline__ void load_weights(float *dst, const float *src, int n4, int tid)
{
    const float4 *s = reinterpret_cast<const float4 *>(src);
    float4 *d = reinterpret_cast<float4 *>(dst);
    for (int i = tid; i < n4; i += NT) d[i] = __ldg(s + i);
}
// acc[o][f] += sum_k W[k][8 og + o] * X[k][4 fg + f]
__device__ __forceinline__ void layer_products(const float *W, const float *X, int og, int fg, float (&acc)[8][4])
{
    const float4 *w4 = reinterpret_cast<const float4 *>(W) + 2 * og;
    const float4 *x4 = reinterpret_cast<const float4 *>(X) + fg;
#pragma unroll 4
    for (int k = 0; k < CD; ++k) {
        const float4 wa = w4[k * (CD / 4)], wb = w4[k * (CD / 4) + 1], x = x4[k * (TF / 4)];
        const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
        const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int o = 0; o < 8; ++o)
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[o][f] = fmaf(w[o], xv[f], acc[o][f]);
    }
}

extern "C" __global__ void __launch_bounds__(NT, 1) wavernn_melresnet_kernel(const MParams p)
{
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, og = tid & 15, fg = tid >> 4;
    const Tile tile = p.tiles[blockIdx.x];
    float *Wl = sm + SM_WL, *X = sm + SM_X, *Y = sm + SM_Y;
    const float *blob = p.blob;

    // ---- conv_in: the tile's 68 input rows [row][80] (zeros past the segment), five taps of [80][128] weights ----
    for (int i = tid; i < (TF + KS - 1) * FEAT; i += NT) {
        const int r = i / FEAT;
        Y[i] = r < tile.mel_rows ? p.mel[(size_t)tile.mel_row0 * FEAT + i] : 0.f;
    }
    float acc[8][4];
#pragma unroll
    for (int o = 0; o < 8; ++o)
#pragma unroll
        for (int f = 0; f < 4; ++f) acc[o][f] = 0.f;
    for (int c = 0; c < KS; ++c) {
        __syncthreads();                                   // the previous tap's weights are no longer read (and Y is complete)
        load_weights(Wl, blob + c * FEAT * CD, FEAT * CD / 4, tid);
        __syncthreads();
        const float4 *w4 = reinterpret_cast<const float4 *>(Wl) + 2 * og;
        const float *xin = Y + (4 * fg + c) * FEAT;
#pragma unroll 4
        for (int i = 0; i < FEAT; ++i) {
            const float4 wa = w4[i * (CD / 4)], wb = w4[i * (CD / 4) + 1];
            const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
            const float xv[4] = {xin[i], xin[FEAT + i], xin[2 * FEAT + i], xin[3 * FEAT + i]};
#pragma unroll
            for (int o = 0; o < 8; ++o)
#pragma unroll
                for (int f = 0; f < 4; ++f) acc[o][f] = fmaf(w[o], xv[f], acc[o][f]);
        }
    }
    {
        const float *s = blob + KS * FEAT * CD, *b = s + CD;
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const int ch = 8 * og + o;
            const float sc = __ldg(s + ch), sh = __ldg(b + ch);
            float4 v;
            v.x = fmaxf(fmaf(acc[o][0], sc, sh), 0.f);
            v.y = fmaxf(fmaf(acc[o][1], sc, sh), 0.f);
            v.z = fmaxf(fmaf(acc[o][2], sc, sh), 0.f);
            v.w = fmaxf(fmaf(acc[o][3], sc, sh), 0.f);
            reinterpret_cast<float4 *>(X + ch * TF)[fg] = v;
        }
    }
    // ---- residual blocks ----
    const float *lw = blob + W0_FLOATS;
    for (int blk = 0; blk < p.res_blocks; ++blk) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half, lw += LAYER_FLOATS) {
            __syncthreads();                               // X / Y of the previous layer are written, its weights no longer read
            load_weights(Wl, lw, CD * CD / 4, tid);
            __syncthreads();
#pragma unroll
            for (int o = 0; o < 8; ++o)
#pragma unroll
                for (int f = 0; f < 4; ++f) acc[o][f] = 0.f;
            layer_products(Wl, half == 0 ? X : Y, og, fg, acc);
            const float *s = lw + CD * CD, *b = s + CD;
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                const int ch = 8 * og + o;
                const float sc = __ldg(s + ch), sh = __ldg(b + ch);
                float4 v = make_float4(fmaf(acc[o][0], sc, sh), fmaf(acc[o][1], sc, sh), fmaf(acc[o][2], sc, sh), fmaf(acc[o][3], sc, sh));
                if (half == 0) {
                    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                    reinterpret_cast<float4 *>(Y + ch * TF)[fg] = v;
                } else {                                   // x + residual: this thread is the only one touching these four words of X
                    float4 *xp = reinterpret_cast<float4 *>(X + ch * TF) + fg;
                    const float4 r = *xp;
                    *xp = make_float4(v.x + r.x, v.y + r.y, v.z + r.z, v.w + r.w);
                }
            }
        }
    }
    // ---- conv_out ----
    __syncthreads();
    load_weights(Wl, lw, CD * CD / 4, tid);
    __syncthreads();
#pragma unroll
    for (int o = 0; o < 8; ++o)
#pragma unroll
        for (int f = 0; f < 4; ++f) acc[o][f] = 0.f;
    layer_products(Wl, X, og, fg, acc);
    const float *bo = lw + CD * CD;
    float bias[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) bias[o] = __ldg(bo + 8 * og + o);
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        const int fr = 4 * fg + f;
        if (fr < tile.nvalid) {
            float4 *dst = reinterpret_cast<float4 *>(p.aux + (size_t)(tile.out_row0 + fr) * CD + 8 * og);
            dst[0] = make_float4(acc[0][f] + bias[0], acc[1][f] + bias[1], acc[2][f] + bias[2], acc[3][f] + bias[3]);
            dst[1] = make_float4(acc[4][f] + bias[4], acc[5][f] + bias[5], acc[6][f] + bias[6], acc[7][f] + bias[7]);
        }
    }
}

}  // namespace wrnn_mel
