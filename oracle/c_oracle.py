"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the C oracle (oracle/wavernn_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may
import this module; the product package never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_wavernn.so")
_LIB = None

WEIGHT_KEYS = (
    "I.weight", "I.bias",
    "rnn1.weight_ih_l0", "rnn1.weight_hh_l0", "rnn1.bias_ih_l0", "rnn1.bias_hh_l0",
    "rnn2.weight_ih_l0", "rnn2.weight_hh_l0", "rnn2.bias_ih_l0", "rnn2.bias_hh_l0",
    "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc3.weight", "fc3.bias",
)


class _Dims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in
                ("rnn_dims", "fc_dims", "feat_dims", "aux_dims", "n_classes", "mode")]


class _Weights(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("I_w", "I_b", "r1_wih", "r1_whh", "r1_bih", "r1_bhh",
                 "r2_wih", "r2_whh", "r2_bih", "r2_bhh",
                 "fc1_w", "fc1_b", "fc2_w", "fc2_b", "fc3_w", "fc3_b")]


def build(force=False):
    """Compile the oracle with oracle/Makefile (gcc only)."""
    src = [os.path.join(_HERE, f) for f in ("wavernn_oracle.c", "wavernn_oracle_steps.inc", "Makefile")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src)):
        return _SO
    subprocess.run(["make", "-s", "-C", _HERE, "-f", os.path.join(_HERE, "Makefile")], check=True)
    return _SO


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        i64, p = ctypes.c_int64, ctypes.c_void_p
        L.orc_fold_index.argtypes = [i64, i64, i64, ctypes.POINTER(i64), ctypes.POINTER(i64)]
        L.orc_fold_index.restype = None
        L.orc_fold.argtypes = [p, i64, i64, i64, i64, p]
        L.orc_fold.restype = None
        L.orc_fade_tables.argtypes = [i64, p, p]
        L.orc_fade_tables.restype = None
        L.orc_xfade_unfold.argtypes = [p, i64, i64, i64, p]
        L.orc_xfade_unfold.restype = None
        L.orc_decode_mu_law.argtypes = [p, i64, i64]
        L.orc_decode_mu_law.restype = None
        L.orc_tail_fade.argtypes = [p, i64, i64]
        L.orc_tail_fade.restype = None
        L.orc_label_to_float.argtypes = [ctypes.c_int, ctypes.c_int]
        L.orc_label_to_float.restype = ctypes.c_float
        L.orc_generate_folds.argtypes = [ctypes.POINTER(_Dims), ctypes.POINTER(_Weights), p, p, i64, i64,
                                         p, p, p, p, p, p, ctypes.c_int, ctypes.c_int]
        L.orc_generate_folds.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def fold_index(total_len, target, overlap):
    """-> (num_folds, padded_len); fatchord_version.py:298-309"""
    n, pl = ctypes.c_int64(), ctypes.c_int64()
    lib().orc_fold_index(total_len, target, overlap, ctypes.byref(n), ctypes.byref(pl))
    return int(n.value), int(pl.value)


def fold(x, target, overlap):
    """x: [L, F] float32 -> [B, S, F]; fatchord_version.py:272-319"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    L, F = x.shape
    B, _ = fold_index(L, target, overlap)
    out = np.zeros((max(B, 0), target + 2 * overlap, F), dtype=np.float32)
    if B > 0:
        lib().orc_fold(_ptr(x), L, F, target, overlap, _ptr(out))
    return out


def fade_tables(overlap):
    fi = np.zeros(max(overlap, 1), dtype=np.float64)
    fo = np.zeros(max(overlap, 1), dtype=np.float64)
    lib().orc_fade_tables(overlap, _ptr(fi), _ptr(fo))
    return fi[:overlap], fo[:overlap]


def xfade_unfold(y, overlap):
    """y: [B, S] float64 (NOT modified; a copy is faded) -> unfolded 1-D; fatchord_version.py:321-383"""
    y = np.array(y, dtype=np.float64, order="C", copy=True)
    B, S = y.shape
    target = S - 2 * overlap
    out = np.zeros(B * (target + overlap) + overlap, dtype=np.float64)
    lib().orc_xfade_unfold(_ptr(y), B, S, overlap, _ptr(out))
    return out


def decode_mu_law(y, n_classes):
    """utility/dsp.py:100-105 with from_labels=False"""
    y = np.array(y, dtype=np.float64, order="C", copy=True)
    lib().orc_decode_mu_law(_ptr(y), y.size, n_classes)
    return y


def tail_fade(out, n_fade):
    """fatchord_version.py:235-237 (caller has already trimmed to wave_len)"""
    out = np.array(out, dtype=np.float64, order="C", copy=True)
    if n_fade > out.size:
        raise ValueError("operands could not be broadcast together (wave_len < 20*hop)")
    lib().orc_tail_fade(_ptr(out), out.size, n_fade)
    return out


def label_to_float(k, n_classes):
    return np.float32(lib().orc_label_to_float(int(k), int(n_classes)))


def pack_weights(state):
    """state: mapping of torch state_dict keys -> array-like; returns (struct, keepalive, dims-less)"""
    keep = []
    w = _Weights()
    for field, key in zip([f[0] for f in _Weights._fields_], WEIGHT_KEYS):
        t = state[key]
        a = np.ascontiguousarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t, dtype=np.float32)
        keep.append(a)
        setattr(w, field, a.ctypes.data)
    return w, keep


def dims_from_state(state, mode):
    I_w = state["I.weight"]
    rnn = I_w.shape[0]
    fc = state["fc1.weight"].shape[0]
    aux = state["fc1.weight"].shape[1] - rnn
    feat = I_w.shape[1] - 1 - aux
    C = state["fc3.weight"].shape[0]
    return _Dims(rnn, fc, feat, aux, C, 0 if mode == "RAW" else 1)


def generate_folds(state, mode, mels, aux, uniforms, forced_x=None, want_logits=False,
                   precision="fp32", threads=None):
    """Step loop over folded conditioning (fatchord_version.py:171-222).

    mels [B,S,feat], aux [B,S,4*aux_dims] float32; uniforms [S,B] (RAW) / [S,B,11] (MOL).
    Returns dict(samples [B,S] f32, labels [B,S] i32 | None, mix [B,S] | None, logits [S,B,C] | None).
    """
    mels = np.ascontiguousarray(mels, dtype=np.float32)
    aux = np.ascontiguousarray(aux, dtype=np.float32)
    B, S, _ = mels.shape
    d = dims_from_state(state, mode)
    w, keep = pack_weights(state)
    uniforms = np.ascontiguousarray(uniforms, dtype=np.float32)
    need = (S, B) if mode == "RAW" else (S, B, d.n_classes // 3 + 1)
    if uniforms.shape[:2] != need[:2] or (mode == "MOL" and uniforms.shape != need):
        raise ValueError("uniforms shape %r, need %r" % (uniforms.shape, need))
    fx = None if forced_x is None else np.ascontiguousarray(forced_x, dtype=np.float32)
    logits = np.zeros((S, B, d.n_classes), dtype=np.float32) if want_logits else None
    samples = np.zeros((B, S), dtype=np.float32)
    labels = np.zeros((B, S), dtype=np.int32) if mode == "RAW" else None
    mix = np.zeros((B, S), dtype=np.int32) if mode == "MOL" else None
    nt = threads if threads else (os.cpu_count() or 1)
    rc = lib().orc_generate_folds(ctypes.byref(d), ctypes.byref(w), _ptr(mels), _ptr(aux), B, S,
                                  _ptr(uniforms), _ptr(fx), _ptr(logits), _ptr(samples), _ptr(labels),
                                  _ptr(mix), 1 if precision == "fp64" else 0, nt)
    if rc:
        raise RuntimeError("oracle orc_generate_folds failed rc=%d" % rc)
    del keep
    return dict(samples=samples, labels=labels, mix=mix, logits=logits)


def assemble(samples, batched, target, overlap, mu_law_classes, wave_len, hop_length):
    """generate() epilogue, fatchord_version.py:222-237: widen, xfade/unfold, mu-law, trim, tail fade."""
    out = np.asarray(samples, dtype=np.float32).astype(np.float64)          # :222-224
    out = xfade_unfold(out, overlap) if batched else out[0].copy()           # :226-229
    if mu_law_classes:
        out = decode_mu_law(out, mu_law_classes)                             # :231-232
    out = out[:wave_len]                                                     # :236
    return tail_fade(out, 20 * hop_length)                                   # :235,237


def generate_from_conditioning(state, mode, mels_up, aux_up, batched, target, overlap, mu_law,
                               hop_length, wave_len, uniforms, precision="fp32", threads=None):
    """Everything after the conditioning network: fold -> step loop -> epilogue.

    mels_up [L, feat], aux_up [L, 4*aux] are the upsample network's outputs for batch item 0.
    """
    if batched:
        m = fold(mels_up, target, overlap)
        a = fold(aux_up, target, overlap)
    else:
        m, a = np.asarray(mels_up, np.float32)[None], np.asarray(aux_up, np.float32)[None]
    r = generate_folds(state, mode, m, a, uniforms, precision=precision, threads=threads)
    C = dims_from_state(state, mode).n_classes
    mu = C if (mu_law and mode == "RAW") else 0                              # :152
    return assemble(r["samples"], batched, target, overlap, mu, wave_len, hop_length), r
