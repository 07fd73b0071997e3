"""Shared checkers for the parity tests (test infrastructure; may use oracle/)."""
import os

import numpy as np
import torch

from oracle import c_oracle, synth, torch_port

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Stated tolerances (DESIGN.md "Parity contract")
TOL_LOGITS_FP32 = 2e-5      # teacher-forced logits, fp32 weights, max abs
TOL_LOGITS_BF16 = 6e-2      # teacher-forced logits, bf16 weights, max abs
TOL_CDF = 2e-6              # half-width of the "justified flip" window on the inverse-CDF test
TOL_MOL_X = 2e-5            # MOL sample value, fp32
TOL_MULAW_ABS = 1e-15       # numpy SVML pow vs glibc / CUDA pow (1 ulp of the power, before the -1)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def digest_of(g, key):
    return bytes(g[key]).decode()


def state_for(mode, geometry, expect_digest=None):
    sd = synth.make_state(mode, geometry, 0)
    if expect_digest is not None:
        assert synth.state_digest(sd) == expect_digest, "synthetic weights differ from the ones the goldens were minted with"
    return sd


def folded_conditioning(sd, mel, geometry, batched, target, overlap):
    """Conditioning via the torch-port oracle -> (mels [B,S,80], aux [B,S,128]) numpy."""
    g = synth.GEOMETRY[geometry]
    with torch.no_grad():
        m, a = torch_port.conditioning(sd, mel, g["upsample_factors"], 2)
    if batched:
        return c_oracle.fold(m[0].numpy(), target, overlap), c_oracle.fold(a[0].numpy(), target, overlap)
    return m.numpy().copy(), a.numpy().copy()


def labels_to_float(labels, C):
    k = torch.as_tensor(np.asarray(labels).astype(np.int64))
    return (2 * k.float() / (C - 1.) - 1.).numpy()


def check_raw_labels_consistent(sd, mels_f, aux_f, U, labels, eps=TOL_CDF, precision="fp64"):
    """'Justified flip' criterion: teacher-force the fp64 oracle on the candidate's
    OWN label history and require, at every (step, fold),  cdf[k-1]-eps <= u < cdf[k]+eps.
    Returns (n_bad, worst_margin)."""
    labels = np.asarray(labels)
    B, S = labels.shape
    C = sd["fc3.weight"].shape[0]
    forced = labels_to_float(labels, C).T.copy()                    # [S,B]
    U = np.asarray(U, dtype=np.float32)
    n_bad, worst = 0, 0.0
    for b0 in range(0, B, 4):                                       # bound memory: S*4*C logits
        b1 = min(B, b0 + 4)
        r = c_oracle.generate_folds(sd, "RAW", mels_f[b0:b1], aux_f[b0:b1], U[:, b0:b1].copy(),
                                    forced_x=forced[:, b0:b1].copy(), want_logits=True, precision=precision)
        lg = r["logits"].astype(np.float64)                          # [S,b,C]
        p = np.exp(lg - lg.max(-1, keepdims=True))
        p /= p.sum(-1, keepdims=True)
        cdf = np.cumsum(p, -1)
        k = labels[b0:b1].T.astype(np.int64)                         # [S,b]
        hi = np.take_along_axis(cdf, k[..., None], -1)[..., 0]
        lo = np.where(k > 0, np.take_along_axis(cdf, np.maximum(k - 1, 0)[..., None], -1)[..., 0], 0.0)
        hi = np.where(k == C - 1, np.inf, hi)
        u = U[:, b0:b1].astype(np.float64)
        margin = np.maximum(lo - u, u - hi)                          # <= 0 strictly inside; u == hi is outside
        bad = (margin > eps)
        n_bad += int(bad.sum())
        worst = max(worst, float(margin.max()))
    return n_bad, worst


def check_mol_samples_consistent(sd, mels_f, aux_f, U, samples, tol=TOL_MOL_X):
    """Teacher-force the fp64 oracle on the candidate's own sample history; at every step the
    oracle's draw (same uniforms) must equal the candidate's within tol.  Returns max abs diff."""
    samples = np.asarray(samples, dtype=np.float32)
    forced = samples.T.copy()
    r = c_oracle.generate_folds(sd, "MOL", mels_f, aux_f, np.asarray(U, np.float32), forced_x=forced,
                                precision="fp64")
    return float(np.abs(r["samples"].astype(np.float64) - samples.astype(np.float64)).max())
