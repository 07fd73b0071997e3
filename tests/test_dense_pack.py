"""CPU tests of precision "bf16-dense" (the tcgen05 path): the packed operand stream + bundle table, replayed in numpy
(tests/dense_replay.py), must reproduce the oracle's teacher-forced logits within the stated bf16 tolerances."""
import ctypes

import numpy as np
import pytest

from expressive_speech_synthesis_research_b200 import _lib
from oracle import c_oracle, synth
from tests.dense_replay import DenseReplay, BUNDLE

# Stated tolerances (teacher-forced logits, max abs, random-init weights, logits of O(0.1-1)):
TOL_DENSE_WEIGHTS_ONLY = 3e-3      # bf16 weights (rounded after the fp64 folding), exact activations
TOL_DENSE = 5e-3                   # bf16 weights and bf16 activations / conditioning (what the kernel computes)


def _case(seed=5, B=5, S=6, mode="RAW"):
    sd = synth.make_state(mode, "ref", 3)
    rng = np.random.default_rng(seed)
    mels = rng.uniform(0, 1, (B, S, 80)).astype(np.float32)
    aux = rng.normal(0, 1, (B, S, 128)).astype(np.float32)
    forced = rng.uniform(-1, 1, (S, B)).astype(np.float32)
    U = np.zeros((S, B) if mode == "RAW" else (S, B, 11), np.float32) + 0.5
    want = c_oracle.generate_folds(sd, mode, mels, aux, U, forced_x=forced, want_logits=True, precision="fp64")["logits"]
    return sd, mels, aux, forced, want


def test_dense_program_shape():
    sd, *_ = _case()
    r = DenseReplay(sd)
    assert (r.CL, r.UPC, r.BC) == (8, 64, 32)
    assert r.table.dtype == BUNDLE and len(r.table) == r.nb
    off = 0
    for bd in r.table:
        assert int(bd["src_off"]) == off and 0 < int(bd["bytes"]) <= 32768 and int(bd["bytes"]) % 16 == 0
        seg_bytes = sum(int(s["rows"]) * int(s["nk"]) * 32 for s in bd["seg"][:int(bd["nseg"])])
        assert seg_bytes == int(bd["bytes"])
        off += int(bd["bytes"])
    assert off == r.stream_bytes
    # every accumulator the epilogues read is overwritten (first touch) exactly once per step
    firsts = sorted(int(s["dcol"]) for bd in r.table for s in bd["seg"][:int(bd["nseg"])] if int(s["first"]))
    assert firsts == [0, 32, 64, 96, 128, 160, 192, 224, 256]
    commits = [int(bd["commit"]) for bd in r.table if int(bd["commit"])]
    assert commits == [1, 2, 3, 4, 5, 6]


def test_dense_stream_reproduces_oracle_logits_with_exact_activations():
    sd, mels, aux, forced, want = _case()
    got = DenseReplay(sd).run(mels.astype(np.float64), aux.astype(np.float64), forced, round_act=False)
    assert np.isfinite(got).all()
    err = np.abs(got - want).max()
    assert 1e-7 < err < TOL_DENSE_WEIGHTS_ONLY, err


@pytest.mark.parametrize("mode", ["RAW", "MOL"])
def test_dense_stream_with_bf16_activations_within_tolerance(mode):
    sd, mels, aux, forced, want = _case(mode=mode)
    got = DenseReplay(sd, mode).run(mels.astype(np.float64), aux.astype(np.float64), forced, round_act=True)
    err = np.abs(got - want).max()
    assert err < TOL_DENSE, err
    if mode == "MOL":
        assert got.shape[-1] == 30
        return
    # the approximation must not change what is sampled in any material way: softmax distributions stay close
    p = lambda lg: np.exp(lg - lg.max(-1, keepdims=True)) / np.exp(lg - lg.max(-1, keepdims=True)).sum(-1, keepdims=True)
    tv = 0.5 * np.abs(p(got) - p(want)).sum(-1).max()
    assert tv < 2e-2, tv


def test_dense_rejects_unsupported_configs():
    L = _lib.lib()
    lay = (ctypes.c_int64 * 8)()
    for cfg in (_lib.Config(512, 512, 80, 32, 30, 0, 2), _lib.Config(512, 512, 80, 32, 256, 0, 2), _lib.Config(512, 512, 80, 32, 512, 1, 2)):
        assert L.wrnn_dense_layout(ctypes.byref(cfg), lay) != 0
