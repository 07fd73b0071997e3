"""TEST INFRASTRUCTURE ONLY -- mint tests/golden/*.npz from the LIVE reference.

Run in the build container (needs /root/reference):  python oracle/make_golden.py
Every array is produced by the reference's own code (fatchord_version.py, utility/dsp.py,
utility/distribution.py) imported through oracle/ref_shim.py; nothing here comes from the
oracle restatements or the product.  The fixtures are small and committed; the script is
committed with them so they can be re-minted.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE in sys.path:
    sys.path.remove(HERE)
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim, synth  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def build(mode, geometry, seed=0):
    fv = ref_shim.install_reference()
    import contextlib
    with contextlib.redirect_stdout(open(os.devnull, "w")):
        m = fv.WaveRNN(**synth.model_kwargs(mode, geometry))
    sd = synth.make_state(mode, geometry, seed)
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    m.eval()
    return m, sd


def golden_index_and_epilogue():
    """fold / xfade / label / mu-law / tail-fade known answers (pure index + fp64 work)."""
    m, _ = build("RAW", "ref")
    g = {}
    # --- fold_with_overlap on an integer-valued ramp so the gather is checkable exactly
    cases = [(10, 2, 1), (11, 2, 1), (9, 2, 1), (3, 2, 1), (4, 2, 1), (2, 2, 1), (1, 2, 1),
             (6000, 1000, 100), (6100, 1000, 100), (5999, 1000, 100), (1201, 1000, 100),
             (1200, 1000, 100), (150, 1000, 100), (100, 1000, 100), (32200, 11000, 550),
             (160200, 11000, 550), (220825, 11000, 550), (12100, 11000, 550), (12101, 11000, 550),
             (23650, 11000, 550), (23651, 11000, 550), (777, 50, 7), (64, 5, 0)]
    meta = []
    for (L, t, o) in cases:
        x = (torch.arange(L * 3, dtype=torch.float32).reshape(1, L, 3) % 8191) + 1.0
        try:
            f = m.fold_with_overlap(x, t, o).numpy()
            meta.append((L, t, o, f.shape[0]))
            if f.size <= 20000:
                g["fold_%d_%d_%d" % (L, t, o)] = f
            else:                                    # keep first/last rows + a checksum
                g["fold_%d_%d_%d_head" % (L, t, o)] = f[:, :3, :].copy()
                g["fold_%d_%d_%d_tail" % (L, t, o)] = f[:, -3:, :].copy()
                g["fold_%d_%d_%d_sum" % (L, t, o)] = f.astype(np.float64).sum(axis=(1, 2))
        except Exception as e:                       # L <= overlap -> reference fails downstream
            meta.append((L, t, o, -1))
    g["fold_cases"] = np.array(meta, dtype=np.int64)
    # docstring example fatchord_version.py:291-295
    x = torch.arange(1, 11, dtype=torch.float32).reshape(1, 10, 1)
    g["fold_docstring"] = m.fold_with_overlap(x, 2, 1).numpy()

    # --- xfade_and_unfold on random float64
    rng = np.random.default_rng(7)
    xmeta = []
    for i, (B, t, o) in enumerate([(3, 2, 1), (1, 10, 4), (4, 37, 5), (6, 1000, 100), (2, 300, 1),
                                   (2, 11000, 550), (3, 16, 8), (2, 0, 3)]):
        y = rng.uniform(-1, 1, size=(B, t + 2 * o))
        out = m.xfade_and_unfold(y.copy(), t, o)
        xmeta.append((B, t, o))
        g["xfade_in_%d" % i] = y
        g["xfade_out_%d" % i] = out
    g["xfade_cases"] = np.array(xmeta, dtype=np.int64)
    g["xfade_docstring"] = m.xfade_and_unfold(
        np.array([[1., 2, 3, 4], [4, 5, 6, 7], [7, 8, 9, 10]]), 2, 1)

    # --- RAW label -> float (fatchord_version.py:214) for every label of the 9-bit model
    k = torch.arange(512)
    g["label_to_float_512"] = (2 * k.float() / (512 - 1.) - 1.).numpy()
    k = torch.arange(1024)
    g["label_to_float_1024"] = (2 * k.float() / (1024 - 1.) - 1.).numpy()

    # --- decode_mu_law(y, 512, False) (utility/dsp.py:100-105) on the levels and on random f64
    fv = ref_shim.install_reference()
    lv = g["label_to_float_512"].astype(np.float64)
    g["mulaw_levels_in"] = lv
    g["mulaw_levels_out"] = fv.decode_mu_law(lv.copy(), 512, False)
    r = rng.uniform(-1, 1, size=4096)
    r[:4] = [0.0, 1.0, -1.0, 1e-300]
    g["mulaw_rand_in"] = r
    g["mulaw_rand_out"] = fv.decode_mu_law(r.copy(), 512, False)

    # --- tail fade tables (fatchord_version.py:235) for both hops
    g["tail_fade_200"] = np.linspace(1, 0, 20 * 200)
    g["tail_fade_275"] = np.linspace(1, 0, 20 * 275)
    np.savez_compressed(os.path.join(OUT, "index_epilogue.npz"), **g)
    print("index_epilogue.npz", len(g), "arrays")


def golden_conditioning():
    """UpsampleNetwork output (fatchord_version.py:79-86) through generate()'s prologue."""
    g = {}
    for geometry in ("ref", "fatchord"):
        m, sd = build("RAW", geometry)
        hop = m.hop_length
        T = 9
        mel = synth.make_mel(T, seed=5)
        with torch.no_grad():
            x = m.pad_tensor(mel.transpose(1, 2), pad=m.pad, side="both")
            mu, aux = m.upsample(x.transpose(1, 2))
        assert mu.shape == (1, T * hop, 80) and aux.shape == (1, T * hop, 128)
        rows = np.unique(np.concatenate([np.arange(0, T * hop, 37), np.arange(0, 12),
                                         np.arange(T * hop - 12, T * hop)]))
        g[geometry + "_rows"] = rows
        g[geometry + "_mels"] = mu[0].numpy()[rows]
        g[geometry + "_aux"] = aux[0].numpy()[rows]
        g[geometry + "_mels_colsum"] = mu[0].double().sum(0).numpy()
        g[geometry + "_aux_colsum"] = aux[0].double().sum(0).numpy()
        g[geometry + "_digest"] = np.frombuffer(synth.state_digest(sd).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "conditioning.npz"), **g)
    print("conditioning.npz")


def golden_teacher_forced():
    """Teacher-forced logits from WaveRNN.forward (fatchord_version.py:119-148)."""
    g = {}
    for mode in ("RAW", "MOL"):
        m, sd = build(mode, "ref")
        hop, pad = m.hop_length, m.pad
        B, frames = 3, 2
        seq = frames * hop
        mel = torch.rand(B, 80, frames + 2 * pad, generator=torch.Generator().manual_seed(11))
        x = torch.rand(B, seq, generator=torch.Generator().manual_seed(12)) * 2 - 1
        if mode == "RAW":                           # feed values on the label grid, as generate() would
            x = 2 * torch.floor((x + 1) / 2 * 511 + 0.5) / 511. - 1.
        x[:, 0] = 0.0                               # generate() starts from x = 0 (:175)
        with torch.no_grad():
            logits = m.forward(x, mel)              # (B, seq, C)
            mu, aux = m.upsample(mel)               # conditioning actually consumed
        steps = np.unique(np.concatenate([np.arange(0, 8), np.arange(8, seq, 13), [seq - 1]]))
        g[mode + "_x"] = x.numpy()
        g[mode + "_mel"] = mel.numpy()
        g[mode + "_steps"] = steps
        g[mode + "_logits"] = logits.numpy()[:, steps, :]
        g[mode + "_cond_mels_sum"] = mu.double().sum((1, 2)).numpy()
        g[mode + "_cond_aux_sum"] = aux.double().sum((1, 2)).numpy()
        g[mode + "_digest"] = np.frombuffer(synth.state_digest(sd).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "teacher_forced.npz"), **g)
    print("teacher_forced.npz")


def _capture_generate(m, mel, batched, target, overlap, mu_law, U):
    """Run the reference's generate() with injected uniforms, also recording the raw
    per-fold samples (the tensor built at fatchord_version.py:222)."""
    rec = {}
    orig_stack = torch.stack

    def stack(tensors, *a, **k):
        out = orig_stack(tensors, *a, **k)
        rec["stack"] = out
        return out

    torch.stack = stack
    try:
        wav = ref_shim.reference_generate(m, mel, batched, target, overlap, mu_law, uniforms=U)
    finally:
        torch.stack = orig_stack
    samples = rec["stack"].transpose(0, 1).numpy().copy()      # (B, S)
    return wav, samples


def golden_losses():
    """Training losses (train_wavernn.py:28-44): the reference's discretized_mix_logistic_loss
    (utility/distribution.py:16-84) on random mixture parameters, reduce=True and False, including targets
    in the edge bins and a narrow component (the cdf_delta <= 1e-5 branch)."""
    ref_shim.install_reference()
    from utility.distribution import discretized_mix_logistic_loss
    g = {}
    gen = torch.Generator().manual_seed(21)
    B, T, K = 3, 50, 10
    y_hat = torch.randn(B, T, 3 * K, generator=gen)
    y_hat[..., 2 * K:] = y_hat[..., 2 * K:] * 2 - 4                  # log-scales around -4
    y_hat[0, :5, 2 * K:] = -16.0                                     # very narrow components
    y = torch.rand(B, T, 1, generator=gen) * 2 - 1
    y[1, :4, 0] = torch.tensor([-1.0, -0.9995, 0.9995, 1.0])         # edge bins
    g["mol_y_hat"], g["mol_y"] = y_hat.numpy(), y.numpy()
    g["mol_loss"] = np.array(discretized_mix_logistic_loss(y_hat, y).item())
    g["mol_loss_unreduced"] = discretized_mix_logistic_loss(y_hat, y, reduce=False).numpy()
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **g)
    print("losses.npz")


def golden_free_running():
    """End-to-end generate() (fatchord_version.py:150-243) with pre-drawn uniforms."""
    g = {}
    cases = [
        # name, mode, geometry, T, batched, target, overlap, mu_law
        ("raw_batched", "RAW", "ref", 30, True, 1000, 100, True),
        ("raw_batched_nomu", "RAW", "ref", 24, True, 700, 60, False),
        ("raw_unbatched", "RAW", "ref", 22, False, 11000, 550, True),
        ("mol_batched", "MOL", "ref", 30, True, 1000, 100, True),
        ("mol_unbatched", "MOL", "ref", 22, False, 11000, 550, True),
        ("raw_fatchord", "RAW", "fatchord", 23, True, 900, 80, True),
        ("raw_default_fold", "RAW", "ref", 62, True, 11000, 550, True),   # one fold of 12100 steps
    ]
    meta = []
    for name, mode, geometry, T, batched, target, overlap, mu_law in cases:
        m, sd = build(mode, geometry)
        hop = m.hop_length
        mel = synth.make_mel(T, seed=21)
        L = T * hop
        if batched:
            n = (L - overlap) // (target + overlap)
            if L - (n * (target + overlap) + overlap) != 0:
                n += 1
            B, S = n, target + 2 * overlap
        else:
            B, S = 1, L
        U = synth.make_uniforms(S, B, mode, seed=123)
        wav, samples = _capture_generate(m, mel, batched, target, overlap, mu_law, U)
        assert samples.shape == (B, S), (samples.shape, B, S)
        g[name + "_wav"] = wav
        if mode == "RAW":
            lab = np.rint((samples.astype(np.float64) + 1) * 511 / 2).astype(np.int16)
            assert np.array_equal((2 * torch.tensor(lab).float() / 511. - 1.).numpy(), samples)
            g[name + "_labels"] = lab
        else:
            g[name + "_samples"] = samples
        g[name + "_digest"] = np.frombuffer(synth.state_digest(sd).encode(), dtype=np.uint8)
        meta.append((name, mode, geometry, T, int(batched), target, overlap, int(mu_law), B, S))
        print(" ", name, "B=%d S=%d" % (B, S), "wav", wav.shape)
    g["cases"] = np.array(meta, dtype=np.str_)
    np.savez_compressed(os.path.join(OUT, "free_running.npz"), **g)
    print("free_running.npz")


if __name__ == "__main__":
    torch.set_num_threads(1)            # fixed MKL reduction order for the minted fixtures
    os.makedirs(OUT, exist_ok=True)
    golden_index_and_epilogue()
    golden_conditioning()
    golden_teacher_forced()
    golden_free_running()
    golden_losses()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
