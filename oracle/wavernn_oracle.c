/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the WaveRNN batched-generation hot path.
 *
 * A plain-C restatement of the reference's algorithm, used as the checker by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  Nothing in
 * the product package may link, import or call this file.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/WaveRNN).  Third-party arithmetic the reference delegates to
 * (torch ATen nn.Linear / nn.GRUCell / softmax / Categorical, numpy linspace /
 * sqrt / power) is restated from its documented semantics; the restatement is
 * pinned against outputs of the reference itself (tests/golden/ *.npz, minted by
 * oracle/make_golden.py from the live reference in the build container).
 *
 * Summation order: plain ascending-k accumulation.  REAL selects the working
 * precision: float mirrors the reference's fp32 model, double gives the
 * "truth" twin used to bound fp32 re-association noise.
 *
 * Build: see oracle/Makefile (gcc -O3 -mavx2 -ffp-contract=off -pthread).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>

typedef struct {
    int rnn_dims;   /* 512 (hparams.py:36) */
    int fc_dims;    /* 512 (hparams.py:37) */
    int feat_dims;  /* 80  (hparams.py:18) */
    int aux_dims;   /* res_out_dims // 4 = 32 (fatchord_version.py:104) */
    int n_classes;  /* 2**bits RAW | 30 MOL (fatchord_version.py:96-99) */
    int mode;       /* 0 RAW, 1 MOL */
} orc_dims;

/* torch state_dict tensors, [out, in] row-major fp32 (fatchord_version.py:109-114) */
typedef struct {
    const float *I_w, *I_b;
    const float *r1_wih, *r1_whh, *r1_bih, *r1_bhh;
    const float *r2_wih, *r2_whh, *r2_bih, *r2_bhh;
    const float *fc1_w, *fc1_b, *fc2_w, *fc2_b, *fc3_w, *fc3_b;
} orc_weights;

/* ------------------------------------------------------------------------- */
/* fold_with_overlap index arithmetic -- fatchord_version.py:298-309          */
/* ------------------------------------------------------------------------- */
void orc_fold_index(int64_t total_len, int64_t target, int64_t overlap,
                    int64_t *num_folds, int64_t *padded_len)
{
    int64_t n = (total_len - overlap) / (target + overlap);       /* :301 (python // on the values generate() can pass) */
    if ((total_len - overlap) < 0 && (total_len - overlap) % (target + overlap) != 0) n -= 1; /* floor division */
    int64_t extended = n * (overlap + target) + overlap;           /* :302 */
    int64_t remaining = total_len - extended;                      /* :303 */
    int64_t plen = total_len;
    if (remaining != 0) {                                          /* :306-309 */
        n += 1;
        plen = total_len + (target + 2 * overlap - remaining);
    }
    *num_folds = n;
    *padded_len = plen;
}

/* fold_with_overlap gather -- fatchord_version.py:311-319.  x: [L, F]; folded: [B, S, F] */
void orc_fold(const float *x, int64_t total_len, int64_t feat, int64_t target, int64_t overlap,
              float *folded)
{
    int64_t n, plen;
    orc_fold_index(total_len, target, overlap, &n, &plen);
    int64_t S = target + 2 * overlap;
    for (int64_t i = 0; i < n; ++i) {
        int64_t start = i * (target + overlap);                    /* :315 */
        for (int64_t s = 0; s < S; ++s) {
            int64_t p = start + s;
            float *dst = folded + (i * S + s) * feat;
            if (p < total_len) memcpy(dst, x + p * feat, (size_t)feat * sizeof(float));
            else memset(dst, 0, (size_t)feat * sizeof(float));     /* pad_tensor(..., side='after') :309 */
        }
    }
}

/* ------------------------------------------------------------------------- */
/* numpy.linspace(start, stop, num) float64 -- used by fatchord_version.py:235,363 */
/* y_i = i*step + start (two roundings), last element forced to stop.          */
/* ------------------------------------------------------------------------- */
static void np_linspace(double start, double stop, int64_t num, double *y)
{
    if (num <= 0) return;
    int64_t div = num - 1;
    double delta = stop - start;
    if (div > 0) {
        double step = delta / (double)div;
        for (int64_t i = 0; i < num; ++i) {
            volatile double m = (double)i * step;                  /* y *= step  */
            y[i] = m + start;                                      /* y += start */
        }
        y[num - 1] = stop;
    } else {
        volatile double m = 0.0 * delta;
        y[0] = m + start;
    }
}

/* fade tables of xfade_and_unfold -- fatchord_version.py:357-369 */
void orc_fade_tables(int64_t overlap, double *fade_in, double *fade_out)
{
    int64_t silence_len = overlap / 2;                             /* :358 */
    int64_t fade_len = overlap - silence_len;                      /* :359 */
    double *t = (double *)malloc(sizeof(double) * (size_t)(fade_len > 0 ? fade_len : 1));
    np_linspace(-1.0, 1.0, fade_len, t);                           /* :363 */
    for (int64_t i = 0; i < silence_len; ++i) fade_in[i] = 0.0;    /* :368 */
    for (int64_t i = 0; i < fade_len; ++i) {
        volatile double a = 1.0 + t[i];
        volatile double b = 0.5 * a;
        fade_in[silence_len + i] = sqrt(b);                        /* :364 */
        volatile double c = 1.0 - t[i];
        volatile double d = 0.5 * c;
        fade_out[i] = sqrt(d);                                     /* :365 */
    }
    for (int64_t i = 0; i < silence_len; ++i) fade_out[fade_len + i] = 0.0;   /* :369 */
    free(t);
}

/* xfade_and_unfold -- fatchord_version.py:353-383.  y: [B, S] (modified in place, as the
 * reference does); out: [B*(target+overlap)+overlap] */
void orc_xfade_unfold(double *y, int64_t num_folds, int64_t length, int64_t overlap, double *out)
{
    int64_t target = length - 2 * overlap;                         /* :354 */
    int64_t total_len = num_folds * (target + overlap) + overlap;  /* :355 */
    double *fi = (double *)malloc(sizeof(double) * (size_t)(overlap > 0 ? overlap : 1));
    double *fo = (double *)malloc(sizeof(double) * (size_t)(overlap > 0 ? overlap : 1));
    orc_fade_tables(overlap, fi, fo);
    for (int64_t b = 0; b < num_folds; ++b) {
        double *row = y + b * length;
        for (int64_t j = 0; j < overlap; ++j) row[j] *= fi[j];                    /* :372 */
        for (int64_t j = 0; j < overlap; ++j) row[length - overlap + j] *= fo[j]; /* :373 */
    }
    for (int64_t p = 0; p < total_len; ++p) out[p] = 0.0;          /* :375 */
    for (int64_t b = 0; b < num_folds; ++b) {                      /* :378-381 */
        int64_t start = b * (target + overlap);
        for (int64_t s = 0; s < length; ++s) out[start + s] += y[b * length + s];
    }
    free(fi);
    free(fo);
}

/* decode_mu_law(y, mu=n_classes, from_labels=False) -- utility/dsp.py:100-105 */
void orc_decode_mu_law(double *y, int64_t n, int64_t n_classes)
{
    double mu = (double)(n_classes - 1);                           /* :103 */
    for (int64_t i = 0; i < n; ++i) {
        double v = y[i];
        double sgn = (v > 0.0) ? 1.0 : ((v < 0.0) ? -1.0 : 0.0);   /* np.sign */
        volatile double q = sgn / mu;
        volatile double p = pow(1.0 + mu, fabs(v));
        volatile double pm1 = p - 1.0;
        y[i] = q * pm1;                                            /* :104 */
    }
}

/* generate() tail -- fatchord_version.py:235-237: out[-n_fade:] *= linspace(1, 0, n_fade) */
void orc_tail_fade(double *out, int64_t wave_len, int64_t n_fade)
{
    double *f = (double *)malloc(sizeof(double) * (size_t)(n_fade > 0 ? n_fade : 1));
    np_linspace(1.0, 0.0, n_fade, f);
    for (int64_t i = 0; i < n_fade; ++i) out[wave_len - n_fade + i] *= f[i];
    free(f);
}

/* label -> float of the RAW branch -- fatchord_version.py:214 (fp32 ops in this order) */
float orc_label_to_float(int k, int n_classes)
{
    volatile float a = 2.0f * (float)k;
    volatile float b = a / ((float)n_classes - 1.0f);
    return b - 1.0f;
}

/* ------------------------------------------------------------------------- */
/* The step loop -- fatchord_version.py:180-216, instantiated for float and    */
/* double working precision.                                                   */
/* ------------------------------------------------------------------------- */
#define REAL float
#define RNAME(x) x##_f32
#define R_EXP expf
#define R_LOG logf
#define R_TANH tanhf
#include "wavernn_oracle_steps.inc"
#undef REAL
#undef RNAME
#undef R_EXP
#undef R_LOG
#undef R_TANH

#define REAL double
#define RNAME(x) x##_f64
#define R_EXP exp
#define R_LOG log
#define R_TANH tanh
#include "wavernn_oracle_steps.inc"
#undef REAL
#undef RNAME
#undef R_EXP
#undef R_LOG
#undef R_TANH

/*
 * Run the autoregressive loop for B folds of S steps.
 *   mels [B,S,feat], aux [B,S,4*aux_dims]  folded conditioning (fatchord_version.py:167-169)
 *   uniforms  RAW: [S,B]   MOL: [S,B,11]   pre-drawn U[0,1) (see oracle/ref_shim.py)
 *   forced_x  [S,B] or NULL: teacher forcing -- the value fed back as x after step s
 *             (fatchord_version.py:119-148 semantics: x[:,0]=0, x[:,s+1]=forced[s])
 *   logits_out [S,B,C] or NULL;  samples_out [B,S] fp32;  labels_out [B,S] (RAW) or NULL
 *   mix_out [B,S] (MOL mixture index) or NULL
 *   precision 0 = fp32 working precision, 1 = fp64
 * Folds are independent (zero initial state, :173-175) so they are run on a thread pool.
 */
typedef struct {
    const orc_dims *d; const orc_weights *w; const float *mels, *aux; int64_t B, S;
    const float *uniforms, *forced_x; float *logits_out, *samples_out;
    int32_t *labels_out, *mix_out; int precision;
    int64_t next; int err; pthread_mutex_t mu;
} orc_job;

static void *orc_worker(void *arg)
{
    orc_job *j = (orc_job *)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int64_t b = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (b >= j->B) break;
        int r = j->precision
            ? run_fold_f64(j->d, j->w, j->mels, j->aux, j->B, j->S, b, j->uniforms, j->forced_x,
                           j->logits_out, j->samples_out, j->labels_out, j->mix_out)
            : run_fold_f32(j->d, j->w, j->mels, j->aux, j->B, j->S, b, j->uniforms, j->forced_x,
                           j->logits_out, j->samples_out, j->labels_out, j->mix_out);
        if (r) { pthread_mutex_lock(&j->mu); j->err = r; pthread_mutex_unlock(&j->mu); }
    }
    return NULL;
}

int orc_generate_folds(const orc_dims *d, const orc_weights *w,
                       const float *mels, const float *aux, int64_t B, int64_t S,
                       const float *uniforms, const float *forced_x,
                       float *logits_out, float *samples_out, int32_t *labels_out,
                       int32_t *mix_out, int precision, int num_threads)
{
    if (d->mode != 0 && d->mode != 1) return 1;
    if (d->mode == 1 && d->n_classes % 3 != 0) return 2;           /* distribution.py:98 */
    orc_job job = { d, w, mels, aux, B, S, uniforms, forced_x, logits_out, samples_out,
                    labels_out, mix_out, precision, 0, 0, PTHREAD_MUTEX_INITIALIZER };
    int nt = num_threads > 0 ? num_threads : 1;
    if (nt > B) nt = (int)B;
    if (nt > 256) nt = 256;
    pthread_t tid[256];
    for (int t = 1; t < nt; ++t) pthread_create(&tid[t], NULL, orc_worker, &job);
    orc_worker(&job);
    for (int t = 1; t < nt; ++t) pthread_join(tid[t], NULL);
    return job.err;
}
