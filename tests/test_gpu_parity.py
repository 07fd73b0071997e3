"""GPU parity suite (-m gpu): the CUDA path, called through the C ABI / the reference-facing
generate(), against the oracle and the fixtures minted from the live reference."""
import ctypes

import numpy as np
import pytest
import torch

from expressive_speech_synthesis_research_b200 import WaveRNN, _lib
from oracle import c_oracle, synth
from tests import helpers as H

pytestmark = pytest.mark.gpu

_MODELS = {}


def model(mode, geometry="ref"):
    key = (mode, geometry)
    if key not in _MODELS:
        m = WaveRNN(**synth.model_kwargs(mode, geometry))
        m.load_state_dict(synth.make_state(mode, geometry, 0))
        _MODELS[key] = m.cuda()
    return _MODELS[key]


def run_folds(m, mels_f, aux_f, U, forced=None, logits=False):
    """Feed already-folded conditioning [B,S,*] through wrnn_generate_folds (rows = b*S + s)."""
    dev = torch.device("cuda", 0)
    B, S, _ = mels_f.shape
    mu = torch.as_tensor(mels_f).reshape(B * S, -1).contiguous().to(dev)
    au = torch.as_tensor(aux_f).reshape(B * S, -1).contiguous().to(dev)
    starts = np.arange(B, dtype=np.int64) * S
    limits = starts + S
    eng = m._engine(dev)
    r = m._run_folds(eng, dev, mu, au, starts, limits, S, U, 0, forced, logits)
    torch.cuda.synchronize()
    return {k: (v.cpu().numpy() if v is not None else None) for k, v in r.items()}


# ---------------------------------------------------------------------------------------------
# epilogue: bit-exact fp64 crossfade / unfold, label map, tail fade; mu-law within tolerance
# ---------------------------------------------------------------------------------------------
def xfade_gpu(y32, batched, overlap, mu, wave_len, tail):
    L = _lib.lib()
    y = torch.as_tensor(np.ascontiguousarray(y32, dtype=np.float32)).cuda()
    out = torch.empty(wave_len, dtype=torch.float64, device="cuda")
    _lib.check(L.wrnn_xfade_unfold(y.data_ptr(), y.shape[0], y.shape[1], int(batched), overlap, mu, wave_len, tail,
                                   out.data_ptr(), None))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_xfade_unfold_bit_exact_vs_golden_and_oracle():
    ge = H.load_golden("index_epilogue.npz")
    rng = np.random.default_rng(3)
    for i, (B, t, o) in enumerate(ge["xfade_cases"]):
        B, t, o = int(B), int(t), int(o)
        y64 = ge["xfade_in_%d" % i]
        y32 = y64.astype(np.float32)                       # the device path consumes the fp32 samples
        total = B * (t + o) + o
        want = c_oracle.xfade_unfold(y32.astype(np.float64), o)
        got = xfade_gpu(y32, True, o, 0, total, 0)
        assert np.array_equal(got, want), (B, t, o)
        if np.array_equal(y32.astype(np.float64), y64):    # inputs exactly representable -> golden applies directly
            assert np.array_equal(got, ge["xfade_out_%d" % i])
    # exactly-representable inputs against the reference's own output (docstring example + label-grid values)
    doc = np.array([[1., 2, 3, 4], [4, 5, 6, 7], [7, 8, 9, 10]])
    assert np.array_equal(xfade_gpu(doc, True, 1, 0, 10, 0), ge["xfade_docstring"])
    lab = rng.integers(0, 512, size=(5, 1200))
    y = H.labels_to_float(lab, 512)
    want = c_oracle.assemble(y, True, 1000, 100, 0, 5000, 200)
    assert np.array_equal(xfade_gpu(y, True, 100, 0, 5000, 4000), want)
    # trimmed + tail-faded + mu-law
    want = c_oracle.assemble(y, True, 1000, 100, 512, 5000, 200)
    got = xfade_gpu(y, True, 100, 512, 5000, 4000)
    assert np.abs(got - want).max() <= H.TOL_MULAW_ABS
    # unbatched
    want = c_oracle.assemble(y[:1], False, 0, 0, 512, 1100, 50)
    got = xfade_gpu(y[:1], False, 0, 512, 1100, 1000)
    assert np.abs(got - want).max() <= H.TOL_MULAW_ABS
    want = c_oracle.assemble(y[:1], False, 0, 0, 0, 1100, 50)
    assert np.array_equal(xfade_gpu(y[:1], False, 0, 0, 1100, 1000), want)


def test_mu_law_levels_and_errors():
    ge = H.load_golden("index_epilogue.npz")
    lv = ge["label_to_float_512"]                           # fp32 label grid
    got = xfade_gpu(lv[None, :], False, 0, 512, 512, 0)
    assert np.abs(got - ge["mulaw_levels_out"]).max() <= H.TOL_MULAW_ABS
    L = _lib.lib()
    y = torch.zeros(2, 100, device="cuda")
    out = torch.empty(100, dtype=torch.float64, device="cuda")
    assert L.wrnn_xfade_unfold(y.data_ptr(), 2, 100, 1, 0, 0, 100, 0, out.data_ptr(), None) == -1    # overlap 0
    assert L.wrnn_xfade_unfold(y.data_ptr(), 2, 100, 1, 10, 0, 100, 200, out.data_ptr(), None) == -1  # tail > wave_len
    assert L.wrnn_xfade_unfold(y.data_ptr(), 2, 100, 1, 10, 0, 500, 0, out.data_ptr(), None) == -1    # beyond unfolded


# ---------------------------------------------------------------------------------------------
# teacher-forced logits vs WaveRNN.forward of the reference (golden) and vs the oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["RAW", "MOL"])
def test_teacher_forced_logits_vs_reference_forward(mode):
    g = H.load_golden("teacher_forced.npz")
    H.state_for(mode, "ref", H.digest_of(g, mode + "_digest"))
    m = model(mode)
    x, mel, steps, want = g[mode + "_x"], torch.from_numpy(g[mode + "_mel"]), g[mode + "_steps"], g[mode + "_logits"]
    B, seq = x.shape
    with torch.no_grad():
        m.eval()
        mu, aux = m.upsample_fp32(mel.cuda())              # forward(): no extra padding (fatchord_version.py:124)
        m.train()
    forced = np.concatenate([x[:, 1:], np.zeros((B, 1), np.float32)], 1).T.copy()
    U = synth.make_uniforms(seq, B, mode).numpy()
    r = run_folds(m, mu.cpu().numpy(), aux.cpu().numpy(), U, forced=forced, logits=True)
    got = r["logits"].transpose(1, 0, 2)[:, steps, :]
    err = np.abs(got - want).max()
    assert err <= H.TOL_LOGITS_FP32, err
    # and against the fp64 oracle on every step
    o = c_oracle.generate_folds(synth.make_state(mode, "ref", 0), mode, mu.cpu().numpy(), aux.cpu().numpy(), U,
                                forced_x=forced, want_logits=True, precision="fp64")
    assert np.abs(r["logits"] - o["logits"]).max() <= H.TOL_LOGITS_FP32


# ---------------------------------------------------------------------------------------------
# free-running generate() vs the reference's own generate() (golden) with injected uniforms
# ---------------------------------------------------------------------------------------------
def _cases():
    g = H.load_golden("free_running.npz")
    return g, [tuple(c) for c in g["cases"]]


@pytest.mark.parametrize("idx", range(7))
def test_generate_matches_reference(idx):
    g, cases = _cases()
    name, mode, geometry, T, batched, target, overlap, mu_law, B, S = cases[idx]
    T, batched, target, overlap, mu_law, B, S = int(T), bool(int(batched)), int(target), int(overlap), bool(int(mu_law)), int(B), int(S)
    sd = H.state_for(mode, geometry, H.digest_of(g, name + "_digest"))
    m = model(mode, geometry)
    mel = synth.make_mel(T, seed=21)
    U = synth.make_uniforms(S, B, mode, seed=123)
    wav, ex = m.generate(mel, batched, target, overlap, mu_law, uniforms=U, return_samples=True)
    assert m.training                                       # generate() leaves the module in train() (:241)
    want_wav = g[name + "_wav"]
    assert wav.dtype == np.float64 and wav.shape == want_wav.shape
    labels = ex["labels"].cpu().numpy()
    samples = ex["samples"].cpu().numpy()
    if mode == "RAW":
        want = g[name + "_labels"].astype(np.int32)
        assert np.array_equal(samples, H.labels_to_float(labels, 512))          # label -> float bit-exact
        mism = int((labels != want).sum())
        print("%s: %d / %d labels differ from the reference" % (name, mism, want.size))
        if mism:
            mf, af = H.folded_conditioning(sd, mel, geometry, batched, target, overlap)
            nbad, worst = H.check_raw_labels_consistent(sd, mf, af, U.numpy(), labels)
            assert nbad == 0, (nbad, worst)                                     # every flip is a justified flip
        assert mism <= max(2, want.size // 2000)
        if mism == 0:
            assert np.abs(wav - want_wav).max() <= H.TOL_MULAW_ABS
            if not mu_law:
                assert np.array_equal(wav, want_wav)                             # crossfade/unfold/trim/fade bit-exact
    else:
        assert np.abs(samples - g[name + "_samples"]).max() <= H.TOL_MOL_X
        assert np.abs(wav - want_wav).max() <= 2 * H.TOL_MOL_X


def test_generate_both_arities_and_save_path(tmp_path):
    from expressive_speech_synthesis_research_b200.wavio import load_wav
    m = model("RAW")
    mel = synth.make_mel(24, seed=2)
    U = synth.make_uniforms(700 + 120, 7, "RAW", seed=9)
    a = m.generate(mel, True, 700, 60, True, uniforms=U)
    p = tmp_path / "out.wav"
    b = m.generate(mel, str(p), True, 700, 60, True, uniforms=U)                # upstream 6-arg form (gen_wavernn.py:34)
    c = m.generate(mel, None, batched=True, target=700, overlap=60, mu_law=True, uniforms=U)
    assert np.array_equal(a, b) and np.array_equal(a, c)
    y, sr = load_wav(p)
    assert sr == 16000 and np.array_equal(y, a.astype(np.float32))
    with pytest.raises(ValueError):                                             # T < 21 frames (reference: broadcast error)
        m.generate(synth.make_mel(10), True, 700, 60, True)
    with pytest.raises(RuntimeError):                                           # L <= overlap -> no folds
        m.generate(synth.make_mel(25), True, 700, 6000, True)


# ---------------------------------------------------------------------------------------------
# full-size properties (BASELINE config 2 geometry) + pooling / chunking equivalences
# ---------------------------------------------------------------------------------------------
def test_full_size_config2_properties():
    m = model("RAW")
    T = 801                                                 # 10 s @ 16 kHz / hop 200 -> 14 folds of 12100
    mel = synth.make_mel(T, seed=0)
    w1, e1 = m.generate(mel, True, 11000, 550, True, seed=77, return_samples=True)
    w2, e2 = m.generate(mel, True, 11000, 550, True, seed=77, return_samples=True)
    w3 = m.generate(mel, True, 11000, 550, True, seed=78)
    assert w1.shape == ((T - 1) * 200,) and np.isfinite(w1).all() and np.abs(w1).max() <= 1.0
    assert np.array_equal(w1, w2)                           # deterministic for a fixed seed
    assert not np.array_equal(w1, w3)
    lab = e1["labels"].cpu().numpy()
    assert lab.shape == (14, 12100) and lab.min() >= 0 and lab.max() <= 511
    assert len(np.unique(lab)) > 100                        # the sampler actually spreads over classes
    assert np.array_equal(w1[-1:], [0.0])                   # tail fade ends at exactly 0 (linspace endpoint)
    # the oracle agrees on the first steps of two folds given the same uniforms (folds are independent)
    S0 = 300
    U = synth.make_uniforms(12100, 14, "RAW", seed=5)
    _, e = m.generate(mel, True, 11000, 550, True, uniforms=U, return_samples=True)
    lab = e["labels"].cpu().numpy()
    sd = synth.make_state("RAW", "ref", 0)
    mf, af = H.folded_conditioning(sd, mel, "ref", True, 11000, 550)
    sel = [0, 13]
    nbad, worst = H.check_raw_labels_consistent(sd, mf[sel, :S0], af[sel, :S0], U.numpy()[:S0][:, sel], lab[sel, :S0])
    assert nbad == 0, (nbad, worst)


def test_pooled_and_chunked_folds_match_single_calls():
    m = model("RAW")
    t, o = 500, 50
    S = t + 2 * o
    mels = [synth.make_mel(T, seed=40 + i) for i, T in enumerate((24, 31, 27))]
    folds = [_lib.fold_index(mm.shape[-1] * 200, t, o)[0] for mm in mels]
    U = synth.make_uniforms(S, sum(folds), "RAW", seed=4)
    pooled = m.generate_many(mels, t, o, True, uniforms=U)
    b0 = 0
    for mm, nb, wav in zip(mels, folds, pooled):
        single = m.generate(mm, True, t, o, True, uniforms=U[:, b0:b0 + nb])
        assert np.array_equal(single, wav)
        b0 += nb
    # > 64 folds: the library splits the pool into several launches; results must not change
    mel = synth.make_mel(200, seed=3)                       # L = 40000 -> 73 folds of 600
    nb = _lib.fold_index(200 * 200, t, o)[0]
    assert nb > 64
    U = synth.make_uniforms(S, nb, "RAW", seed=6)
    wav, ex = m.generate(mel, True, t, o, True, uniforms=U, return_samples=True)
    lab = ex["labels"].cpu().numpy()
    sd = synth.make_state("RAW", "ref", 0)
    mf, af = H.folded_conditioning(sd, mel, "ref", True, t, o)
    sel = [0, 63, 64, nb - 1]
    nbad, worst = H.check_raw_labels_consistent(sd, mf[sel], af[sel], U.numpy()[:, sel], lab[sel])
    assert nbad == 0, (nbad, worst)


def test_mol_in_kernel_rng_and_bits10():
    m = model("MOL")
    mel = synth.make_mel(40, seed=8)
    a = m.generate(mel, True, 1000, 100, True, seed=5)
    b = m.generate(mel, True, 1000, 100, True, seed=5)
    assert np.array_equal(a, b) and np.isfinite(a).all() and np.abs(a).max() <= 1.0
    # a 10-bit RAW model uses 8 output rows per CTA
    kw = synth.model_kwargs("RAW", "ref", bits=10)
    m10 = WaveRNN(**kw)
    sd = synth.make_state("RAW", "ref", 1, bits=10)
    m10.load_state_dict(sd)
    m10.cuda()
    mel = synth.make_mel(24, seed=2)
    U = synth.make_uniforms(820, 7, "RAW", seed=9)
    _, ex = m10.generate(mel, True, 700, 60, True, uniforms=U, return_samples=True)
    lab = ex["labels"].cpu().numpy()
    assert lab.max() > 512
    mf, af = H.folded_conditioning(sd, mel, "ref", True, 700, 60)
    nbad, worst = H.check_raw_labels_consistent(sd, mf[:2], af[:2], U.numpy()[:, :2], lab[:2])
    assert nbad == 0, (nbad, worst)


# ---------------------------------------------------------------------------------------------
# precision="bf16": resident weights rounded to bf16, fp32 activations / accumulation (config 3 of BASELINE.json)
# ---------------------------------------------------------------------------------------------
TOL_LOGITS_BF16 = 3e-2          # bf16-weight mode, teacher-forced logits, absolute


@pytest.mark.parametrize("mode", ["RAW", "MOL"])
def test_bf16_weights_teacher_forced_logits_and_free_running(mode):
    sd = synth.make_state(mode, "ref", 0)
    m = WaveRNN(**synth.model_kwargs(mode, "ref"))
    m.load_state_dict(sd)
    m.cuda()
    m.precision = "bf16"
    rng = np.random.default_rng(11)
    B, S = 11, 48
    mels = rng.uniform(0, 1, (B, S, 80)).astype(np.float32)
    aux = rng.normal(0, 1, (B, S, 128)).astype(np.float32)
    forced = rng.uniform(-1, 1, (S, B)).astype(np.float32)
    U = synth.make_uniforms(S, B, mode, seed=4).numpy()
    want = c_oracle.generate_folds(sd, mode, mels, aux, U, forced_x=forced, want_logits=True, precision="fp64")["logits"]
    got = run_folds(m, mels, aux, U, forced=forced, logits=True)["logits"]
    err = np.abs(got - want).max()
    cos = float((got * want).sum() / np.sqrt((got * got).sum() * (want * want).sum()))
    assert 1e-6 < err <= TOL_LOGITS_BF16, err             # > 1e-6: the bf16 images are really in use
    assert cos > 0.9999, cos
    # the same model in fp32 on the same inputs stays at fp32 accuracy (separate engine per precision)
    m.precision = "fp32"
    got32 = run_folds(m, mels, aux, U, forced=forced, logits=True)["logits"]
    assert np.abs(got32 - want).max() <= H.TOL_LOGITS_FP32
    # free-running generate() in bf16: finite, in range, right length
    m.precision = "bf16"
    mel = synth.make_mel(40, seed=2)
    wav = m.generate(mel, True, 1500, 150, True, seed=3)
    assert wav.shape == ((40 - 1) * 200,) and np.isfinite(wav).all() and np.abs(wav).max() <= 1.0 + 1e-9


def test_results_do_not_depend_on_the_team_layout(monkeypatch):
    """The host picks how many teams share a CTA (1-3, development cap WRNN_FORCE_TEAMS); work is dealt to warps in
    fixed items / units and partial sums are combined in a fixed order, so samples must be bit-identical."""
    m = model("RAW")
    rng = np.random.default_rng(3)
    B, S = 20, 40
    mels = rng.uniform(0, 1, (B, S, 80)).astype(np.float32)
    aux = rng.normal(0, 1, (B, S, 128)).astype(np.float32)
    U = synth.make_uniforms(S, B, "RAW", seed=8).numpy()
    monkeypatch.delenv("WRNN_FORCE_TEAMS", raising=False)
    ref = run_folds(m, mels, aux, U, logits=True)
    for cap in ("1", "2"):
        monkeypatch.setenv("WRNN_FORCE_TEAMS", cap)
        got = run_folds(m, mels, aux, U, logits=True)
        assert np.array_equal(got["labels"], ref["labels"]) and np.array_equal(got["logits"], ref["logits"]), cap


@pytest.mark.parametrize("mode", ["RAW", "MOL"])
def test_wide_kernel_every_fold_count_and_ragged_limits(monkeypatch, mode):
    """The wide kernel (csrc/wavernn_wide.cuh) at EVERY fold count 1..21: nq = 1..7
    quads per unit, sampler warps with and without a fold, finalize threads without a quad.  A fold's samples and logits must not
    depend on how many folds share its launch (the partial sums are added in a fixed order and handed between the pass warps and
    the finalize warps through named barriers + a consumption counter: a lost hand-off shows up here), must match the grouped
    round-1 kernel's logits to fp32 round-off when both are teacher-forced on the same history, and folds that run past their
    conditioning must read zeros from then on (fatchord_version.py:306-309)."""
    m = model(mode)
    rng = np.random.default_rng(11)
    B, S = 21, 48
    mels = rng.uniform(0, 1, (B, S, 80)).astype(np.float32)
    aux = rng.normal(0, 1, (B, S, 128)).astype(np.float32)
    U = synth.make_uniforms(S, B, mode, seed=9).numpy()
    monkeypatch.setenv("WRNN_KERNEL", "wide")
    ref = run_folds(m, mels, aux, U, logits=True)
    assert m._engine(torch.device("cuda", 0)).info().kernel_kind == 1
    for nb in (1, 2, 3, 4, 7, 8, 9, 13, 20):
        got = run_folds(m, mels[:nb], aux[:nb], np.ascontiguousarray(U[:, :nb]), logits=True)
        assert np.array_equal(got["samples"], ref["samples"][:nb]), nb
        assert np.array_equal(got["logits"], ref["logits"][:, :nb]), nb
    # same history through the grouped kernel: logits agree to round-off (different summation trees)
    forced = ref["samples"].T.copy()                                    # [S, B] the wide kernel's own history
    monkeypatch.setenv("WRNN_KERNEL", "grouped")
    grp = run_folds(m, mels, aux, U, forced=forced, logits=True)
    assert m._engine(torch.device("cuda", 0)).info().kernel_kind == 0
    monkeypatch.setenv("WRNN_KERNEL", "wide")
    wid = run_folds(m, mels, aux, U, forced=forced, logits=True)
    assert np.abs(wid["logits"] - grp["logits"]).max() <= 2e-5
    # very short runs (the conditioning prefetch, the end-of-step pass and the barrier phases have special cases at S = 1, 2, 3)
    for S2 in (1, 2, 3):
        monkeypatch.setenv("WRNN_KERNEL", "wide")
        w = run_folds(m, mels[:, :S2], aux[:, :S2], np.ascontiguousarray(U[:S2]), forced=np.ascontiguousarray(forced[:S2]), logits=True)
        assert np.array_equal(w["logits"], wid["logits"][:S2]), S2        # a prefix of the long run, bit for bit
        assert np.array_equal(w["samples"], wid["samples"][:, :S2]), S2
    # ragged limits: fold f has only 5 + 2 f rows of conditioning, the rest of its steps read zeros
    dev = torch.device("cuda", 0)
    eng = m._engine(dev)
    mu = torch.as_tensor(mels).reshape(B * S, -1).contiguous().to(dev)
    au = torch.as_tensor(aux).reshape(B * S, -1).contiguous().to(dev)
    starts = np.arange(B, dtype=np.int64) * S
    limits = starts + np.minimum(S, 5 + 2 * np.arange(B))
    r = m._run_folds(eng, dev, mu, au, starts, limits, S, U, 0, None, True)
    mz, az = mels.copy(), aux.copy()
    for f in range(B):
        mz[f, limits[f] - starts[f]:] = 0
        az[f, limits[f] - starts[f]:] = 0
    full = run_folds(m, mz, az, U, logits=True)
    assert np.array_equal(r["samples"].cpu().numpy(), full["samples"])
    assert np.array_equal(r["logits"].cpu().numpy(), full["logits"])
