/*
 * wavernn_b200.h -- C ABI of the B200-native WaveRNN batched-generation hot path.
 *
 * The reference (sankar-mukherjee/Expressive-Speech-Synthesis-Research) has no FFI for this
 * path: the boundary is the Python method WaveRNN.generate (WaveRNN/models/fatchord_version.py:150).
 * Each entry point below replaces the part of that method named beside it; the Python host
 * mirror (expressive_speech_synthesis_research_b200/wavernn.py) binds them with ctypes and
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions: plain C types, raw pointers and sizes only (no torch types).  Every function
 * returns 0 on success or a negative wrnn_status; wrnn_last_error() gives the thread-local
 * message.  Pointers marked [dev] are CUDA device pointers borrowed for the duration of the
 * call; [host] are host pointers.  `stream` is a cudaStream_t passed as void* (NULL = default
 * stream).  The generate and epilogue entry points ENQUEUE their work on `stream` and return; a
 * fired in-kernel watchdog is reported by wrnn_synchronize / the next generate call on the
 * handle (WRNN_ERR_TIMEOUT).  wrnn_load_weights and wrnn_measure_exchange are synchronous.
 * One handle per device; a handle is NOT re-entrant: one generate call in flight (the
 * persistent kernels occupy every SM), a second one first waits for the first.
 * Entry points leave the caller's current CUDA device unchanged.
 * There is no CPU fallback: every compute entry point fails with WRNN_ERR_CUDA when no
 * sm_100 device is usable.
 */
#ifndef WAVERNN_B200_H
#define WAVERNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WRNN_ABI_VERSION 3

typedef enum {
    WRNN_OK = 0,
    WRNN_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
    WRNN_ERR_CUDA = -2,        /* CUDA runtime error (message has the cudaError string) */
    WRNN_ERR_STATE = -3,       /* e.g. generate before load_weights */
    WRNN_ERR_TIMEOUT = -4      /* in-kernel watchdog: a grid-level exchange never completed */
} wrnn_status;

enum { WRNN_MODE_RAW = 0, WRNN_MODE_MOL = 1 };
enum { WRNN_PREC_FP32 = 0, WRNN_PREC_BF16 = 1, WRNN_PREC_BF16_DENSE = 2 };

/* Model geometry: WaveRNN.__init__, fatchord_version.py:90-114. */
typedef struct {
    int32_t rnn_dims;    /* 512 */
    int32_t fc_dims;     /* 512 */
    int32_t feat_dims;   /* 80  */
    int32_t aux_dims;    /* res_out_dims / 4 = 32 (fatchord_version.py:104) */
    int32_t n_classes;   /* 2**bits (RAW) or 30 (MOL) (fatchord_version.py:96-99) */
    int32_t mode;        /* WRNN_MODE_* */
    int32_t precision;   /* WRNN_PREC_FP32 / WRNN_PREC_BF16: storage of the RESIDENT WEIGHTS of the persistent FFMA kernel
                            (bf16 = rounded after the fp64 folding of the input layer); activations, accumulation and
                            the exchange stay fp32 in both.
                            WRNN_PREC_BF16_DENSE: the dense-regime kernel (north star: "tcgen05 tensor-core tiles when
                            the fold batch makes the per-step matmul dense"): bf16 weights streamed from L2 by TMA,
                            bf16 activations, fp32 accumulation in tensor memory, fp32 recurrent state and sampling;
                            clusters of 8 CTAs advance 32 folds each, 480 folds in flight per GPU.  RAW with 512 classes, or MOL. */
} wrnn_config;

/* The sixteen state_dict tensors on the step path, torch [out, in] row-major fp32, HOST memory
 * (keys I.*, rnn1.*_l0, rnn2.*_l0, fc1.*, fc2.*, fc3.*). */
typedef struct {
    const float *I_w, *I_b;
    const float *r1_wih, *r1_whh, *r1_bih, *r1_bhh;
    const float *r2_wih, *r2_whh, *r2_bih, *r2_bhh;
    const float *fc1_w, *fc1_b, *fc2_w, *fc2_b, *fc3_w, *fc3_b;
} wrnn_weights;

typedef struct wrnn_handle wrnn_handle;

/* ABI version of the loaded library (== WRNN_ABI_VERSION of the header it was built from). */
int32_t wrnn_abi_version(void);

/* Thread-local message of the last failing call on this thread ("" if none). */
const char *wrnn_last_error(void);

/* Create / destroy the per-device engine (allocates exchange buffers, flags, workspace).
 * Replaces: module construction + .cuda() placement, synthesizer_wavernn.py:17-28. */
int32_t wrnn_create(const wrnn_config *cfg, int32_t device, wrnn_handle **out);
void wrnn_destroy(wrnn_handle *h);

/* Repack the reference's weights into per-SM shared-memory images and upload them
 * (synchronous).  Replaces: restore()/load(), fatchord_version.py:396-405, and get_gru_cell,
 * :252-258.  May be called again to swap checkpoints. */
int32_t wrnn_load_weights(wrnn_handle *h, const wrnn_weights *w /* [host] */);

/* Host-only view of the repack (no GPU needed): the per-CTA shared-memory weight images that
 * wrnn_load_weights uploads, [128][wrnn_packed_floats/128] floats.  Used by the CPU tests to
 * check the algebraic folding of I / conditioning terms against the oracle. */
int64_t wrnn_packed_floats(const wrnn_config *cfg);
int32_t wrnn_pack_weights_host(const wrnn_config *cfg, const wrnn_weights *w /* [host] */,
                               float *out /* [host] */, int64_t out_floats);

/* The same for the WIDE kernel (csrc/wavernn_wide.cuh: every fold of a launch, up to 21, through one exchange per stage;
 * fp32, RAW with 512 classes or MOL): [128 worker CTAs][wrnn_wide_packed_floats / 128] floats; layout[8] = {floats per
 * CTA, offsets of Wih2x, Whh1, Whh2 (gate layout), fc1, fc2, fc3 (fc layout), the conditioning block}; the small vectors
 * follow the conditioning block.  Returns -1 / WRNN_ERR_INVALID for configurations the wide kernel does not serve. */
int64_t wrnn_wide_packed_floats(const wrnn_config *cfg, int64_t *layout /* [host] int64[8] or NULL */);
int32_t wrnn_wide_pack_host(const wrnn_config *cfg, const wrnn_weights *w /* [host] */,
                            float *out /* [host] */, int64_t out_floats);

/* Host-only view of the WRNN_PREC_BF16_DENSE repack (no GPU needed), used by the CPU tests to replay the kernel's
 * tensor-core program in numpy.  layout[8] = {bundles per step, stream bytes per CTA rank, sizeof(bundle record),
 * CTAs per cluster, units per CTA, folds per cluster, per-row vectors, 0}.  stream: [cluster CTAs][stream bytes] bf16
 * operand tiles in issue order; table: the bundle records (csrc/wavernn_dense.cuh::Bundle); sv: fp32
 * [cluster CTAs][vectors][units per CTA] biases and sample coefficients. */
int32_t wrnn_dense_layout(const wrnn_config *cfg, int64_t *layout /* [host] int64[8] */);
int32_t wrnn_dense_pack_host(const wrnn_config *cfg, const wrnn_weights *w /* [host] */,
                             uint8_t *stream /* [host] */, uint8_t *table /* [host] */, float *sv /* [host] */);

/* fold_with_overlap index arithmetic, fatchord_version.py:298-309 (pure host integer code).
 * num_folds may be 0 (total_len <= overlap).  padded_len is the reference's padded length. */
int32_t wrnn_fold_index(int64_t total_len, int64_t target, int64_t overlap,
                        int64_t *num_folds, int64_t *padded_len);

/*
 * The hot loop: fatchord_version.py:171-222 for `num_folds` independent folds of `steps`
 * samples each, reading the UNFOLDED conditioning by index (fold_with_overlap's gather,
 * :311-319, is fused: fold b reads rows fold_start[b] + s; rows >= fold_limit[b] are the
 * zero padding of :306-309).  Folds of several utterances may be pooled in one call.
 *
 *   mels  [dev] float32 [cond_rows, feat_dims]      upsampled mel  (UpsampleNetwork, :79-86)
 *   aux   [dev] float32 [cond_rows, 4*aux_dims]     upsampled aux
 *   fold_start, fold_limit [host] int64 [num_folds] first row / one-past-last valid row
 *   uniforms [dev] float32: RAW [steps, num_folds], MOL [steps, num_folds, n_classes/3+1];
 *            NULL => drawn in-kernel from Philox4x32-10(seed)
 *   forced_x [dev] float32 [steps, num_folds] or NULL: teacher forcing, value fed back after
 *            step s instead of the sample (WaveRNN.forward semantics, :119-148)
 *   logits_out  [dev] float32 [steps, num_folds, n_classes] or NULL
 *   samples_out [dev] float32 [num_folds, steps]    (the tensor of :222, before .cpu())
 *   labels_out  [dev] int32   [num_folds, steps] or NULL   RAW: class index; MOL: mixture index
 */
int32_t wrnn_generate_folds(wrnn_handle *h,
                            const float *mels, const float *aux, int64_t cond_rows,
                            const int64_t *fold_start, const int64_t *fold_limit,
                            int32_t num_folds, int32_t steps,
                            const float *uniforms, uint64_t seed,
                            const float *forced_x, float *logits_out,
                            float *samples_out, int32_t *labels_out, void *stream);

/*
 * The same step loop with the conditioning EXPANDED IN THE KERNEL from frame-rate tensors: the three
 * (Stretch2d, box-filter Conv2d) stages of UpsampleNetwork.forward, the repeat of the MelResNet output (fatchord_version.py:
 * 79-86) and fold_with_overlap's gather (:311-319) are fused into the kernel's conditioning load, so the [samples, 208] fp32
 * tensors (832 B per sample) are never materialised.  WRNN_PREC_BF16_DENSE handles only; others fail with WRNN_ERR_INVALID.
 *
 *   mel_frames [dev] float32 [frame_rows, feat_dims]   mel frames of every utterance, zero-padded by `pad` frames on both sides
 *   aux_frames [dev] float32 [aux_rows, 4*aux_dims]    MelResNet output (one row per unpadded frame)
 *   interp     [dev] float32 [hop, 5]   per phase r = (sample + pad*hop) % hop: four weights of the composite response of the
 *                                       (repeat, FIR) stages and the offset (-2 or -1, stored as float) of the first frame they apply to
 *   fold_geo   [host] int32 [num_folds, 4]  first sample of the fold within its utterance, samples of the utterance (rows beyond
 *                                       are the zero padding of :306-309), row of the utterance's first padded mel frame, row of its
 *                                       first aux frame
 * Remaining arguments as in wrnn_generate_folds.
 */
int32_t wrnn_generate_folds_frames(wrnn_handle *h,
                                   const float *mel_frames, int64_t frame_rows, const float *aux_frames, int64_t aux_rows,
                                   const float *interp, int32_t hop, int32_t pad,
                                   const int32_t *fold_geo, int32_t num_folds, int32_t steps,
                                   const float *uniforms, uint64_t seed,
                                   const float *forced_x, float *logits_out,
                                   float *samples_out, int32_t *labels_out, void *stream);

/*
 * generate() epilogue on the device, fatchord_version.py:222-237:
 * widen to float64, xfade_and_unfold (:321-383, bit-exact: mul and add kept separate),
 * optional decode_mu_law (utility/dsp.py:100-105; pow within 2 ulp), trim to wave_len,
 * linear tail fade over the last tail_fade samples (:235-237).
 *   samples [dev] float32 [num_folds, steps];  batched=0: num_folds must be 1, no crossfade
 *   mu_law_classes: 0 = off, else n_classes
 *   out [dev] float64 [wave_len]
 * Fails with WRNN_ERR_INVALID where the reference raises (overlap <= 0 when batched,
 * tail_fade > wave_len, wave_len beyond the unfolded length).
 */
int32_t wrnn_xfade_unfold(const float *samples, int32_t num_folds, int32_t steps,
                          int32_t batched, int32_t overlap, int32_t mu_law_classes,
                          int64_t wave_len, int32_t tail_fade, double *out, void *stream);

/*
 * The same epilogue for ONE SEGMENT of the waveform, used when the folds of an utterance are
 * sharded over several GPUs (north star: "only the overlap samples are gathered"): `samples`
 * holds rows for the global folds [first_fold, first_fold + num_rows) of a `total_folds`-fold
 * utterance (a neighbour's overlap samples travel as one extra, otherwise zero, row), and
 * out[i] receives global sample seg_start + i for i < seg_len.  Arithmetic per sample is the
 * one of wrnn_xfade_unfold, so concatenated segments are bit-identical to the single-GPU result.
 */
int32_t wrnn_xfade_unfold_segment(const float *samples, int32_t num_rows, int32_t steps, int32_t overlap,
                                  int32_t mu_law_classes, int64_t wave_len, int32_t tail_fade,
                                  int64_t first_fold, int64_t total_folds, int64_t seg_start, int64_t seg_len,
                                  double *out, void *stream);

/* Wait for the generate call enqueued last on this handle (if any) and report its outcome: WRNN_ERR_TIMEOUT when an
 * in-kernel watchdog fired.  wrnn_get_info also waits (it reports the call's device time and status). */
int32_t wrnn_synchronize(wrnn_handle *h);

/* Non-blocking progress of the call in flight: *done = 1 when nothing is pending; *steps_done = sample steps completed so far
 * (written by the step loop every 128 steps into mapped host memory; single-launch calls of the wide kernel only, else 0).
 * Replaces the reference's gen_display progress line, fatchord_version.py:220,246-250. */
int32_t wrnn_query(wrnn_handle *h, int32_t *done, int32_t *steps_done /* may be NULL */);

/* Introspection used by bench.py / tests. */
typedef struct {
    int32_t ctas;                /* CTAs of the persistent kernel (one per SM) */
    int32_t threads;             /* threads per CTA */
    int32_t smem_bytes;          /* dynamic shared memory per CTA */
    int32_t folds_per_group;     /* folds advanced together through one exchange */
    int32_t max_folds_per_launch;
    int32_t exchanges_per_step;  /* grid-level exchanges on the critical path of one step */
    int32_t sm_count;
    int64_t launches;            /* persistent-kernel launches since create */
    int64_t epilogue_launches;
    int32_t last_kernel_status;  /* 0 ok, else WRNN_ERR_TIMEOUT */
    float   last_kernel_ms;      /* device time of the last generate_folds (CUDA events) */
    int32_t kernel_kind;         /* step-loop kernel of the last launch: 0 grouped FFMA (round 1), 1 wide FFMA, 2 dense tcgen05 */
} wrnn_info;
int32_t wrnn_get_info(wrnn_handle *h, wrnn_info *out);

/* Optional in-kernel stage timing (development / profiles/): when enabled, thread 0 of every
 * CTA accumulates clock64() cycles per (stage, phase) of the step loop; the counters of the
 * last wrnn_generate_folds launch are returned as int64 [128 CTAs][32 slots]
 * (slot map in csrc/wavernn_kernel.cuh).  Adds ~20 clock reads per step. */
int32_t wrnn_set_profiling(wrnn_handle *h, int32_t enable);
int32_t wrnn_get_stage_cycles(wrnn_handle *h, int64_t *out /* [host] */, int32_t n);

/* Microbenchmark of the grid-level exchange used by the step loop (wide kernel: each worker CTA publishes its 4 units
 * x 7 quads {3 values, epoch}, every warp then polls the 224 quads of its 32 units from L2 -- 56 KB per CTA; grouped
 * kernel: 32 {value, epoch} pairs published, 4096 gathered), `iters` times on an otherwise empty persistent kernel.
 * Writes the mean device time per exchange in microseconds.  Synchronous. */
int32_t wrnn_measure_exchange(wrnn_handle *h, int32_t iters, float *usec_per_exchange);

/* Which fp32 step-loop kernel serves the next calls on this handle: -1 (default) the wide kernel whenever the model is one it is
 * built for (fp32, RAW with 512 classes or MOL), else the grouped kernel; 0 the grouped kernel; 1 the wide kernel.  The two kernels add
 * their partial sums in different orders, so their samples differ in the last bits: a caller comparing runs pins one.  No effect on
 * the dense precision. */
int32_t wrnn_set_kernel(wrnn_handle *h, int32_t choice);

/* ---- frame-rate conditioning network (csrc/wavernn_cond.cuh) -------------------------------------------------------
 * Replaces: MelResNet.forward, WaveRNN/models/fatchord_version.py:28-45 (ResBlock :10-25): the `aux` half of
 * UpsampleNetwork.forward (:79-86) before Stretch2d repeats it.  Dimensions are the reference's (hparams.py:35-39):
 * 80 mel channels, kernel 5, 128 compute / output channels; `res_blocks` is free.  Eval-mode batch norm arrives folded:
 * blob (host floats, wrnn_cond_blob_floats(res_blocks) of them) =
 *   W0 [5 taps][80 in][128 out] | scale0 [128] | shift0 [128] |
 *   res_blocks x { W1 [128 in][128 out] | scale1 | shift1 | W2 [128 in][128 out] | scale2 | shift2 } | Wout [128 in][128 out] | bias [128]
 * (WaveRNN.pack_melresnet builds it from the state_dict in float64).  One object per device and model; not tied to an engine. */
typedef struct wrnn_cond wrnn_cond;
int64_t wrnn_cond_blob_floats(int32_t res_blocks);
int32_t wrnn_cond_create(int32_t device, const float *blob_host, int64_t n_floats, int32_t res_blocks, wrnn_cond **out);
void wrnn_cond_destroy(wrnn_cond *c);
/* kernels launched by this object so far (bench.py's gpu_launches) */
int64_t wrnn_cond_launches(const wrnn_cond *c);
/* aux = MelResNet(mel) for `nseg` utterances in ONE launch.  mel_frames (device) [rows][80]: every segment's mel frames, already
 * zero-padded by `pad` = 2 frames on both sides (fatchord_version.py:164); segments (HOST) [nseg][3] int32 =
 * {first row of the segment in mel_frames, T = frames out (the segment has T + 4 rows), first row of its output in aux_out};
 * aux_out (device) [rows][128].  A frame's result does not depend on the segment or launch it shares.  Enqueues on `stream`
 * (cudaStream_t) and returns. */
int32_t wrnn_cond_frames(wrnn_cond *c, const float *mel_frames, const int32_t *segments, int32_t nseg, float *aux_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* WAVERNN_B200_H */
