"""CPU suite: pin the oracle restatements (C + torch port) against fixtures minted from the
LIVE reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, synth, torch_port
from tests import helpers as H


@pytest.fixture(scope="module")
def ge():
    return H.load_golden("index_epilogue.npz")


def _ramp(L):
    return ((np.arange(L * 3, dtype=np.float32).reshape(L, 3)) % 8191) + 1.0


def test_fold_docstring_example(ge):
    # fatchord_version.py:291-295
    x = np.arange(1, 11, dtype=np.float32).reshape(10, 1)
    f = c_oracle.fold(x, 2, 1)
    assert np.array_equal(f, ge["fold_docstring"])
    assert np.array_equal(f[..., 0], [[1, 2, 3, 4], [4, 5, 6, 7], [7, 8, 9, 10]])
    assert np.array_equal(c_oracle.xfade_unfold(np.array([[1., 2, 3, 4], [4, 5, 6, 7], [7, 8, 9, 10]]), 1),
                          ge["xfade_docstring"])


def test_fold_cases_bit_exact(ge):
    for L, t, o, n in ge["fold_cases"]:
        L, t, o, n = int(L), int(t), int(o), int(n)
        B, plen = c_oracle.fold_index(L, t, o)
        if n < 0:
            continue
        assert B == n, (L, t, o)
        if B == 0:
            continue
        f = c_oracle.fold(_ramp(L), t, o)
        ft = torch_port.fold_with_overlap(torch.from_numpy(_ramp(L))[None], t, o).numpy()
        assert np.array_equal(f, ft)
        key = "fold_%d_%d_%d" % (L, t, o)
        if key in ge:
            assert np.array_equal(f, ge[key]), key
        else:
            assert np.array_equal(f[:, :3], ge[key + "_head"])
            assert np.array_equal(f[:, -3:], ge[key + "_tail"])
            assert np.array_equal(f.astype(np.float64).sum((1, 2)), ge[key + "_sum"])


def test_fold_edge_cases():
    # L <= overlap -> zero folds (reference then crashes downstream); overlap < L < target+2*overlap -> 1 fold
    assert c_oracle.fold_index(100, 1000, 100)[0] == 0
    assert c_oracle.fold_index(150, 1000, 100) == (1, 1300)   # reference pads t+2o-rem, more than the fold needs
    assert c_oracle.fold_index(1200, 1000, 100) == (1, 1200)
    assert c_oracle.fold_index(1201, 1000, 100) == (2, 2400)
    # BASELINE configs: ref 10 s -> 14 folds, fatchord 10 s -> 20, 10 min -> 832 / 1146
    assert c_oracle.fold_index(801 * 200, 11000, 550)[0] == 14
    assert c_oracle.fold_index(803 * 275, 11000, 550)[0] == 20
    assert c_oracle.fold_index(48001 * 200, 11000, 550)[0] == 832
    assert c_oracle.fold_index(48110 * 275, 11000, 550)[0] == 1146


def test_xfade_unfold_bit_exact(ge):
    for i, (B, t, o) in enumerate(ge["xfade_cases"]):
        y = ge["xfade_in_%d" % i]
        want = ge["xfade_out_%d" % i]
        got = c_oracle.xfade_unfold(y, int(o))
        assert got.dtype == np.float64 and np.array_equal(got, want), (B, t, o)
        assert np.array_equal(torch_port.xfade_and_unfold(y, int(o)), want)


def test_label_to_float_bit_exact(ge):
    for C in (512, 1024):
        want = ge["label_to_float_%d" % C]
        got = np.array([c_oracle.label_to_float(k, C) for k in range(C)], dtype=np.float32)
        assert np.array_equal(got, want)


def test_mu_law_decode(ge):
    for k in ("levels", "rand"):
        got = c_oracle.decode_mu_law(ge["mulaw_%s_in" % k], 512)
        want = ge["mulaw_%s_out" % k]
        assert np.abs(got - want).max() <= H.TOL_MULAW_ABS
    assert c_oracle.decode_mu_law(np.array([0.0, 1.0, -1.0]), 512).tolist() == [0.0, 1.0, -1.0]


def test_tail_fade_bit_exact(ge):
    for hop in (200, 275):
        got = c_oracle.tail_fade(np.ones(20 * hop + 7), 20 * hop)
        assert np.array_equal(got[7:], ge["tail_fade_%d" % hop]) and np.all(got[:7] == 1.0)
    with pytest.raises(ValueError):                       # wave_len < 20*hop (T < 21): reference raises
        c_oracle.tail_fade(np.ones(100), 4000)


def test_conditioning_torch_port():
    g = H.load_golden("conditioning.npz")
    for geometry in ("ref", "fatchord"):
        sd = H.state_for("RAW", geometry, H.digest_of(g, geometry + "_digest"))
        mel = synth.make_mel(9, seed=5)
        m, a = torch_port.conditioning(sd, mel, synth.GEOMETRY[geometry]["upsample_factors"], 2)
        rows = g[geometry + "_rows"]
        np.testing.assert_allclose(m[0].numpy()[rows], g[geometry + "_mels"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(a[0].numpy()[rows], g[geometry + "_aux"], rtol=0, atol=1e-5)
        np.testing.assert_allclose(m[0].double().sum(0).numpy(), g[geometry + "_mels_colsum"], rtol=1e-6)


@pytest.mark.parametrize("mode", ["RAW", "MOL"])
def test_teacher_forced_logits(mode):
    g = H.load_golden("teacher_forced.npz")
    sd = H.state_for(mode, "ref", H.digest_of(g, mode + "_digest"))
    x, mel, steps, want = g[mode + "_x"], torch.from_numpy(g[mode + "_mel"]), g[mode + "_steps"], g[mode + "_logits"]
    B, seq = x.shape
    with torch.no_grad():
        mu, aux = torch_port.upsample(sd, mel, (5, 5, 8), 2)          # forward(): no extra padding (:124)
    assert mu.shape[1] == seq
    forced = np.concatenate([x[:, 1:], np.zeros((B, 1), np.float32)], 1).T.copy()   # value fed after step s
    U = synth.make_uniforms(seq, B, mode).numpy()
    for precision, tol in (("fp32", H.TOL_LOGITS_FP32), ("fp64", 5e-6)):
        r = c_oracle.generate_folds(sd, mode, mu.numpy(), aux.numpy(), U, forced_x=forced, want_logits=True,
                                    precision=precision)
        got = r["logits"].transpose(1, 0, 2)[:, steps, :]
        assert np.abs(got - want).max() <= tol, (precision, np.abs(got - want).max())
    _, lg = torch_port.step_loop(sd, mode, mu, aux, uniforms=torch.from_numpy(U), forced_x=torch.from_numpy(forced),
                                 want_logits=True)
    got = lg.numpy().transpose(1, 0, 2)[:, steps, :]
    assert np.abs(got - want).max() <= 1e-5


def _cases():
    g = H.load_golden("free_running.npz")
    return g, [tuple(c) for c in g["cases"]]


@pytest.mark.parametrize("idx", range(7))
def test_free_running_generate(idx):
    g, cases = _cases()
    name, mode, geometry, T, batched, target, overlap, mu_law, B, S = cases[idx]
    T, batched, target, overlap, mu_law, B, S = int(T), bool(int(batched)), int(target), int(overlap), bool(int(mu_law)), int(B), int(S)
    if S > 6000:
        pytest.skip("long single-fold case is exercised by the GPU suite")
    sd = H.state_for(mode, geometry, H.digest_of(g, name + "_digest"))
    geo = synth.GEOMETRY[geometry]
    mel = synth.make_mel(T, seed=21)
    U = synth.make_uniforms(S, B, mode, seed=123).numpy()
    mf, af = H.folded_conditioning(sd, mel, geometry, batched, target, overlap)
    assert mf.shape[:2] == (B, S)
    wave_len = (T - 1) * geo["hop_length"]
    C = sd["fc3.weight"].shape[0]
    for precision in ("fp32", "fp64"):
        r = c_oracle.generate_folds(sd, mode, mf, af, U, precision=precision)
        if mode == "RAW":
            want = g[name + "_labels"].astype(np.int32)
            nbad, worst = H.check_raw_labels_consistent(sd, mf, af, U, r["labels"])
            assert nbad == 0, (precision, nbad, worst)
            nbad, worst = H.check_raw_labels_consistent(sd, mf, af, U, want)      # the reference's own labels
            assert nbad == 0, ("reference labels", nbad, worst)
            mism = int((r["labels"] != want).sum())
            assert mism <= max(2, B * S // 2000), (precision, mism)               # flips are rare events
            if mism == 0:
                wav = c_oracle.assemble(r["samples"], batched, target, overlap, C if mu_law else 0, wave_len,
                                        geo["hop_length"])
                assert wav.shape == g[name + "_wav"].shape
                assert np.abs(wav - g[name + "_wav"]).max() <= H.TOL_MULAW_ABS
        else:
            want = g[name + "_samples"]
            assert np.abs(r["samples"] - want).max() <= H.TOL_MOL_X, precision
            wav = c_oracle.assemble(r["samples"], batched, target, overlap, 0, wave_len, geo["hop_length"])
            assert np.abs(wav - g[name + "_wav"]).max() <= 2 * H.TOL_MOL_X
    # torch port end to end (same ATen ops as the reference)
    wav = torch_port.generate(sd, mel, batched, target, overlap, mu_law, mode=mode,
                              upsample_factors=geo["upsample_factors"], pad=2, hop_length=geo["hop_length"],
                              uniforms=torch.from_numpy(U))
    assert wav.shape == g[name + "_wav"].shape
    frac_equal = float((wav == g[name + "_wav"]).mean())
    assert frac_equal > 0.99 or np.abs(wav - g[name + "_wav"]).max() <= 1e-4
