"""CPU suite for the C-ABI boundary: the library loads, exports every symbol the header
declares, and its host-only entry points agree with the oracle.  No compute calls (no GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from expressive_speech_synthesis_research_b200 import WaveRNN, _lib, build
from oracle import c_oracle, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    header = open(os.path.join(ROOT, "include", "wavernn_b200.h")).read()
    declared = set(re.findall(r"\b(wrnn_[a-z_0-9]+)\s*\(", header))
    declared -= {"wrnn_status"}
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().wrnn_abi_version() == _lib.ABI_VERSION


def test_fold_index_matches_oracle_grid():
    rng = np.random.default_rng(0)
    cases = [(10, 2, 1), (160200, 11000, 550), (220825, 11000, 550), (100, 1000, 100), (0, 5, 5), (5, 5, 5), (6, 5, 5)]
    cases += [(int(rng.integers(0, 50000)), int(rng.integers(0, 3000)), int(rng.integers(1, 400))) for _ in range(300)]
    for L, t, o in cases:
        assert _lib.fold_index(L, t, o) == c_oracle.fold_index(L, t, o), (L, t, o)
    with pytest.raises(RuntimeError):
        _lib.fold_index(10, 0, 0)
    with pytest.raises(RuntimeError):
        _lib.fold_index(10, -1, 4)


def test_error_reporting_without_gpu():
    L = _lib.lib()
    h = _lib.vp()
    bad = _lib.Config(100, 512, 80, 32, 512, 0, 0)
    assert L.wrnn_create(ctypes.byref(bad), 0, ctypes.byref(h)) == -1
    assert b"rnn_dims" in L.wrnn_last_error()
    if not torch.cuda.is_available():
        ok = _lib.Config(512, 512, 80, 32, 512, 0, 0)
        assert L.wrnn_create(ctypes.byref(ok), 0, ctypes.byref(h)) == -2       # WRNN_ERR_CUDA, no fallback
        out = np.zeros(4)
        assert L.wrnn_xfade_unfold(None, 1, 1, 0, 0, 0, 1, 0, out.ctypes.data, None) == -1


def test_model_mirror_state_dict_and_signature():
    for mode in ("RAW", "MOL"):
        m = WaveRNN(**synth.model_kwargs(mode, "ref"))
        sd = synth.make_state(mode, "ref", 0)
        res = m.load_state_dict(sd, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        assert [k for k, _ in synth.state_shapes(mode, "ref")] == list(m.state_dict().keys())
    with pytest.raises(RuntimeError):
        WaveRNN(**synth.model_kwargs("RAW", "ref") | {"mode": "XYZ"})
    m = WaveRNN(**synth.model_kwargs("RAW", "ref"))
    mel = synth.make_mel(30)
    with pytest.raises(TypeError):
        m.generate(mel, True, 11000)                                   # missing overlap, mu_law
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU"):               # product path fails loudly without CUDA
            m.generate(mel, True, 11000, 550, True)
        assert m.training                                               # :241 -- train() restored even on failure


def test_conditioning_matches_golden():
    from tests import helpers as H
    g = H.load_golden("conditioning.npz")
    for geometry in ("ref", "fatchord"):
        sd = H.state_for("RAW", geometry, H.digest_of(g, geometry + "_digest"))
        m = WaveRNN(**synth.model_kwargs("RAW", geometry)).eval()
        m.load_state_dict(sd)
        with torch.no_grad():
            mu, aux = m.conditioning(synth.make_mel(9, seed=5))
        rows = g[geometry + "_rows"]
        np.testing.assert_allclose(mu.numpy()[rows], g[geometry + "_mels"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(aux.numpy()[rows], g[geometry + "_aux"], rtol=0, atol=1e-5)


def test_wav_roundtrip(tmp_path):
    from expressive_speech_synthesis_research_b200.wavio import load_wav, save_wav
    x = np.linspace(-1, 1, 1000)
    p = tmp_path / "a.wav"
    save_wav(x, p, 16000)
    y, sr = load_wav(p)
    assert sr == 16000 and np.array_equal(y, x.astype(np.float32))
