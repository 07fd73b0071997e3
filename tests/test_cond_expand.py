"""CPU test of the in-kernel conditioning expansion: the [hop, 5] interpolation table that WaveRNN.interp_table
measures from the (Stretch2d, box-filter) stages, applied to frame-rate tensors exactly as csrc/wavernn_dense.cuh::cond_load_frames
does, must reproduce UpsampleNetwork.forward (fatchord_version.py:79-86).  Tolerance: the cascade rounds after every stage, the
fused form once; mel values are in [0, 1)."""
import numpy as np
import pytest
import torch

from expressive_speech_synthesis_research_b200 import WaveRNN
from oracle import synth

TOL_EXPAND = 2e-6


def expand(model, mel):
    """numpy restatement of cond_load_frames for every sample of one utterance -> (mel_up [L, 80], aux [L, 128])."""
    hop, pad = model.hop_length, model.pad
    with torch.no_grad():
        mf, af = model.conditioning_frames(mel)
    table = model.interp_table().numpy().astype(np.float32)
    mf, af = mf.numpy(), af.numpy()
    T = mel.shape[-1]
    L = T * hop
    p = np.arange(L)
    q, r = (p + pad * hop) // hop, (p + pad * hop) % hop
    first = table[r, 4].astype(np.int64)
    out = np.zeros((L, mf.shape[1]), np.float32)
    for j in range(4):                                      # same order as the kernel: w0*a0, then fma w1, w2, w3
        term = table[r, j][:, None] * mf[q + first + j]
        out = term if j == 0 else (out + term).astype(np.float32)
    return out, af[p // hop]


@pytest.mark.parametrize("geometry", ["ref", "fatchord"])
@pytest.mark.parametrize("trained", [False, True])
def test_interp_table_reproduces_the_upsample_network(geometry, trained):
    kw = synth.model_kwargs("RAW", "ref")
    if geometry == "fatchord":
        kw.update(upsample_factors=(5, 5, 11), hop_length=275, sample_rate=22050)
    torch.manual_seed(0)
    m = WaveRNN(**kw)
    m.eval()
    if trained:                                             # the box filters are trainable (fatchord_version.py:75): perturb them
        g = torch.Generator().manual_seed(1)
        for i in (1, 3, 5):
            w = m.upsample.up_layers[i].weight
            w.data.mul_(1.0 + 0.5 * (torch.rand(w.shape, generator=g) - 0.5))
    mel = synth.make_mel(37, seed=2)
    with torch.no_grad():
        want_m, want_a = m.conditioning(mel)
    got_m, got_a = expand(m, mel)
    assert got_m.shape == tuple(want_m.shape) and got_a.shape == tuple(want_a.shape)
    assert np.array_equal(got_a, want_a.numpy())            # the repeat is exact
    err = np.abs(got_m - want_m.numpy()).max()
    assert err <= TOL_EXPAND, err
    t = m.interp_table().numpy()
    assert t.shape == (kw["hop_length"], 5) and set(np.unique(t[:, 4])) <= {-2.0, -1.0}
    if not trained:                                         # box filters: every phase's weights sum to 1
        assert np.abs(t[:, :4].sum(1) - 1).max() < 1e-6


def test_interp_table_follows_the_weights():
    m = WaveRNN(**synth.model_kwargs("RAW", "ref"))
    a = m.interp_table().clone()
    assert m.interp_table() is m._interp_cache[1]           # cached
    m.upsample.up_layers[5].weight.data.mul_(2.0)          # even a .data edit (no version bump) must be seen
    b = m.interp_table()
    assert torch.allclose(b[:, :4], 2 * a[:, :4])
