"""GPU parity of precision "bf16-dense" (tcgen05 / tensor-memory step loop, csrc/wavernn_dense.cuh), called through the
C ABI and through generate(), against the fp64 oracle.  Tolerances are bf16 tolerances and are written here:
  teacher-forced logits   max |dense - oracle fp64| <= TOL_DENSE_LOGITS   (logits of the synthetic model are O(0.3))
  sampling                every label is the inverse-CDF outcome of the kernel's OWN logits (window TOL_DENSE_CDF)
  free running            on the dense path's own history the fp64 oracle draws the same class up to +-DENSE_LABEL_SLACK
Fold-level work is bit-reproducible: pooling / chunking / cluster placement must not change a fold's labels."""
import numpy as np
import pytest
import torch

from expressive_speech_synthesis_research_b200 import WaveRNN
from oracle import c_oracle, synth
from tests import helpers as H
from tests.test_gpu_parity import run_folds

pytestmark = pytest.mark.gpu

TOL_DENSE_LOGITS = 5e-3
TOL_DENSE_CDF = 1e-5
DENSE_LABEL_SLACK = 2

_M = {}


def dense_model():
    if "m" not in _M:
        m = WaveRNN(**synth.model_kwargs("RAW", "ref"))
        m.load_state_dict(synth.make_state("RAW", "ref", 0))
        m = m.cuda()
        m.precision = "bf16-dense"
        _M["m"] = m
    return _M["m"]


def _inputs(B, S, seed):
    rng = np.random.default_rng(seed)
    mels = rng.uniform(0, 1, (B, S, 80)).astype(np.float32)
    aux = rng.normal(0, 1, (B, S, 128)).astype(np.float32)
    U = rng.uniform(0, 1, (S, B)).astype(np.float32)
    return mels, aux, U


def _own_cdf_margin(logits, U, labels):
    lg = logits.astype(np.float64)
    p = np.exp(lg - lg.max(-1, keepdims=True))
    cdf = np.cumsum(p, -1) / p.sum(-1, keepdims=True)
    k = labels.T.astype(np.int64)
    hi = np.take_along_axis(cdf, k[..., None], -1)[..., 0]
    lo = np.where(k > 0, np.take_along_axis(cdf, np.maximum(k - 1, 0)[..., None], -1)[..., 0], 0.0)
    hi = np.where(k == lg.shape[-1] - 1, np.inf, hi)
    u = U.astype(np.float64)
    return np.maximum(lo - u, u - hi)


def test_dense_teacher_forced_logits_vs_oracle():
    m = dense_model()
    sd = synth.make_state("RAW", "ref", 0)
    B, S = 37, 48                                            # two clusters, ragged fold counts
    mels, aux, U = _inputs(B, S, 11)
    forced = np.random.default_rng(12).uniform(-1, 1, (S, B)).astype(np.float32)
    r = run_folds(m, mels, aux, U, forced=forced, logits=True)
    want = c_oracle.generate_folds(sd, "RAW", mels, aux, U, forced_x=forced, want_logits=True, precision="fp64")["logits"]
    err = np.abs(r["logits"] - want).max()
    assert err <= TOL_DENSE_LOGITS, err
    assert _own_cdf_margin(r["logits"], U, r["labels"]).max() <= TOL_DENSE_CDF
    assert np.array_equal(r["samples"], H.labels_to_float(r["labels"], 512))          # label -> float map is exact


def test_dense_free_running_is_consistent_with_the_oracle():
    m = dense_model()
    sd = synth.make_state("RAW", "ref", 0)
    B, S = 9, 300
    mels, aux, U = _inputs(B, S, 21)
    r = run_folds(m, mels, aux, U, logits=True)
    assert _own_cdf_margin(r["logits"], U, r["labels"]).max() <= TOL_DENSE_CDF
    # fp64 oracle teacher-forced on the dense path's own sample history: same logits within the bf16 tolerance at EVERY
    # step of a free run (errors do not build up), and the classes it would draw are the dense ones up to a few levels
    forced = r["samples"].T.copy()
    o = c_oracle.generate_folds(sd, "RAW", mels, aux, U, forced_x=forced, want_logits=True, precision="fp64")
    assert np.abs(r["logits"] - o["logits"]).max() <= TOL_DENSE_LOGITS
    d = np.abs(o["labels"].astype(np.int64) - r["labels"].astype(np.int64))
    assert d.max() <= DENSE_LABEL_SLACK, d.max()
    assert (d == 0).mean() > 0.98, (d == 0).mean()                 # measured 99.6 % equal, the rest off by one level


def test_dense_pooling_chunking_and_placement_do_not_change_a_fold():
    m = dense_model()
    B, S = 70, 40
    mels, aux, U = _inputs(B, S, 31)
    whole = run_folds(m, mels, aux, U)["labels"]
    a = run_folds(m, mels[:17], aux[:17], U[:, :17].copy())["labels"]
    b = run_folds(m, mels[17:], aux[17:], U[:, 17:].copy())["labels"]
    assert np.array_equal(whole, np.concatenate([a, b]))


def test_dense_philox_is_reproducible_and_seed_dependent():
    m = dense_model()
    dev = torch.device("cuda", 0)
    B, S = 5, 64
    mels, aux, _ = _inputs(B, S, 41)
    mu = torch.as_tensor(mels).reshape(B * S, -1).contiguous().to(dev)
    au = torch.as_tensor(aux).reshape(B * S, -1).contiguous().to(dev)
    starts = np.arange(B, dtype=np.int64) * S
    eng = m._engine(dev)
    runs = [m._run_folds(eng, dev, mu, au, starts, starts + S, S, None, seed, None, False)["labels"].cpu().numpy() for seed in (7, 7, 8)]
    assert np.array_equal(runs[0], runs[1]) and not np.array_equal(runs[0], runs[2])


def test_dense_generate_end_to_end_matches_assembly_of_its_own_samples():
    """generate() on the dense path: same fold / crossfade / mu-law epilogue as the fp32 path (bit-exact vs the oracle's
    assembly of the SAME per-fold samples), waveform shape and range as the reference's."""
    m = dense_model()
    g = synth.GEOMETRY["ref"]
    T = 120
    mel = synth.make_mel(T, 3)
    target, overlap = 2000, 100
    wav, ex = m.generate(mel, True, target, overlap, True, seed=5, return_samples=True)
    hop = g["hop_length"]
    assert wav.dtype == np.float64 and wav.shape == ((T - 1) * hop,)
    assert np.isfinite(wav).all() and np.abs(wav).max() <= 1.0
    samples = ex["samples"].cpu().numpy()
    want = c_oracle.assemble(samples, True, target, overlap, 512, (T - 1) * hop, hop)
    assert np.abs(wav - want).max() <= H.TOL_MULAW_ABS


def test_dense_generate_many_chunks_match_single_utterance_calls():
    """A sentence set larger than one chunk of generate_many (> 960 pooled folds): every waveform must equal the one
    generate() returns for that utterance alone with the matching slice of the draws (pooling, chunking and cluster
    placement do not change a fold)."""
    from expressive_speech_synthesis_research_b200 import _lib
    m = dense_model()
    t, o = 300, 20
    S = t + 2 * o
    mels = [synth.make_mel(50 + (7 * i) % 23, seed=60 + i) for i in range(32)]
    folds = [_lib.fold_index(mm.shape[-1] * 200, t, o)[0] for mm in mels]
    assert sum(folds) > 960
    U = synth.make_uniforms(S, sum(folds), "RAW", seed=8)
    pooled = m.generate_many(mels, t, o, True, uniforms=U)
    assert m.last_stats["chunks"] >= 2
    b0 = 0
    for k, (mm, nb, wav) in enumerate(zip(mels, folds, pooled)):
        if k % 5 == 0 or k == len(mels) - 1:
            single = m.generate(mm, True, t, o, True, uniforms=U[:, b0:b0 + nb])
            assert np.array_equal(single, wav), k
        b0 += nb


def test_dense_in_kernel_conditioning_expansion_matches_the_materialised_path():
    """generate() on the dense kernel expands the conditioning inside the kernel from frame-rate tensors.  On the
    same utterance and draws the teacher-forced logits must agree with the path that materialises UpsampleNetwork's output
    (both round the conditioning to bf16: a last-bit difference of the fp32 value can move one bf16 ulp) and with the fp64 oracle."""
    m = dense_model()
    sd = synth.make_state("RAW", "ref", 0)
    T, target, overlap = 64, 900, 60
    mel = synth.make_mel(T, seed=9)
    S = target + 2 * overlap
    B = c_oracle.fold_index(T * 200, target, overlap)[0]
    U = synth.make_uniforms(S, B, "RAW", seed=10)
    forced = np.random.default_rng(13).uniform(-1, 1, (S, B)).astype(np.float32)
    out = {}
    for mode in (True, False):
        m.expand_in_kernel = mode
        try:
            _, ex = m.generate(mel, True, target, overlap, True, uniforms=U, forced_x=forced, return_logits=True)
        finally:
            m.expand_in_kernel = True
        out[mode] = ex["logits"].cpu().numpy()
    assert np.abs(out[True] - out[False]).max() <= 1e-3
    mf, af = H.folded_conditioning(sd, mel, "ref", True, target, overlap)
    want = c_oracle.generate_folds(sd, "RAW", mf, af, U.numpy(), forced_x=forced, want_logits=True, precision="fp64")["logits"]
    assert np.abs(out[True] - want).max() <= TOL_DENSE_LOGITS
    # unbatched: one fold over the whole utterance
    m.generate(synth.make_mel(22, seed=11), False, target, overlap, True, seed=1)


def _mol_from_logits(logits, U):
    """sample_from_discretized_mix_logistic (utility/distribution.py:87-123) in float32 from given logits [S,B,30], U [S,B,11]."""
    lg = logits.astype(np.float32)
    u1 = (1e-5 + ((1.0 - 1e-5) - 1e-5) * U[..., :10].astype(np.float64)).astype(np.float32)
    score = lg[..., :10] - np.log(-np.log(u1))
    arg = score.argmax(-1)
    mean = np.take_along_axis(lg[..., 10:20], arg[..., None], -1)[..., 0]
    ls = np.maximum(np.take_along_axis(lg[..., 20:30], arg[..., None], -1)[..., 0], np.float32(-32.23619130191664))
    u2 = (1e-5 + ((1.0 - 1e-5) - 1e-5) * U[..., 10].astype(np.float64)).astype(np.float32)
    x = mean + np.exp(ls) * (np.log(u2) - np.log(np.float32(1.0) - u2))
    return np.clip(x, -1, 1).astype(np.float32), arg


def test_dense_mol_logits_and_sampling():
    """MOL on the dense kernel: the 30 outputs are rows 0-29 of CTA 0's fc3 tile.  Teacher-forced logits vs the fp64 oracle within the
    bf16 tolerance; every sample is the mixture-of-logistics draw of the kernel's OWN logits (same mixture, value within 1e-5)."""
    m = WaveRNN(**synth.model_kwargs("MOL", "ref"))
    sd = synth.make_state("MOL", "ref", 0)
    m.load_state_dict(sd)
    m = m.cuda()
    m.precision = "bf16-dense"
    B, S = 37, 40
    rng = np.random.default_rng(51)
    mels = rng.uniform(0, 1, (B, S, 80)).astype(np.float32)
    aux = rng.normal(0, 1, (B, S, 128)).astype(np.float32)
    U = rng.uniform(0, 1, (S, B, 11)).astype(np.float32)
    forced = rng.uniform(-1, 1, (S, B)).astype(np.float32)
    r = run_folds(m, mels, aux, U, forced=forced, logits=True)
    want = c_oracle.generate_folds(sd, "MOL", mels, aux, U, forced_x=forced, want_logits=True, precision="fp64")["logits"]
    assert r["logits"].shape == want.shape == (S, B, 30)
    assert np.abs(r["logits"] - want).max() <= TOL_DENSE_LOGITS
    r = run_folds(m, mels, aux, U, logits=True)                       # free running
    x, arg = _mol_from_logits(r["logits"], U)
    assert np.array_equal(arg.T, r["labels"])
    assert np.abs(x.T - r["samples"]).max() <= 1e-5
    mel = synth.make_mel(40, seed=4)                                  # generate(): frames mode, Philox draws
    wav = m.generate(mel, True, 1000, 100, True, seed=3)
    assert wav.shape == (39 * 200,) and np.isfinite(wav).all() and np.abs(wav).max() <= 1.0


def test_dense_edge_cases_fold_counts_and_short_runs():
    """Fold counts around the cluster (32) and wave (480) boundaries, runs of 1-3 steps (the conditioning prefetch reaches past the
    end), the shortest legal utterance through generate(): no watchdog, finite logits, labels consistent with the kernel's logits."""
    from scripts import dense_edge_sweep
    assert dense_edge_sweep.main() == 0


def test_native_melresnet_matches_torch_and_pooling_is_bit_identical():
    """a2 (MelResNet.forward, fatchord_version.py:28-45) as csrc/wavernn_cond.cuh: within fp32 round-off of the PyTorch network, and a
    frame's result independent of the utterances it is pooled with (one launch, tiles of 64 frames; per-utterance cuDNN calls do not
    promise that across shapes)."""
    from tests.test_melresnet_pack import replay
    m = WaveRNN(**synth.model_kwargs("RAW", "fatchord"))
    m.load_state_dict(synth.make_state("RAW", "fatchord", 0))
    m = m.cuda()
    m.eval()
    dev = torch.device("cuda", 0)
    assert m.melresnet_native()
    mels = [synth.make_mel(T, seed=70 + T) for T in (21, 64, 65, 130, 803)]
    mf, af, mrows, arows = m.conditioning_frames_many(mels, dev)
    torch.cuda.synchronize()
    blob, nb = m.pack_melresnet(), len(m.upsample.resnet.layers)
    for i, mel in enumerate(mels):
        T = mel.shape[-1]
        single_m, single_a = m.conditioning_frames(mel.to(dev))
        assert torch.equal(single_m, mf[mrows[i]:mrows[i + 1]])
        assert torch.equal(single_a, af[arows[i]:arows[i + 1]]), T          # bit-identical whatever the pooling
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            want = m.upsample.resnet(torch.nn.functional.pad(mel.to(dev), (2, 2)))[0].t()
        err = float((single_a - want).abs().max())
        assert err <= 2e-5 * max(1.0, float(want.abs().max())), (T, err)
        rep = replay(blob, single_m.cpu().numpy(), nb)
        assert np.abs(rep - single_a.cpu().numpy()).max() <= 2e-5 * max(1.0, float(want.abs().max()))
