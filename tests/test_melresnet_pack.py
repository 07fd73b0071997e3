"""CPU: the weight blob of the frame-rate conditioning kernel (csrc/wavernn_cond.cuh; WaveRNN.pack_melresnet) replayed in numpy with
the kernel's own arithmetic order against the PyTorch MelResNet (reference: fatchord_version.py:28-45, ResBlock :10-25), and the C ABI's
size contract.  The GPU half (kernel vs this replay vs torch, pooled vs single bit-identity) is tests/test_gpu_dense.py."""
import numpy as np
import torch

from expressive_speech_synthesis_research_b200 import WaveRNN, _lib
from oracle import synth


def replay(blob, mel_pad_frames, res_blocks):
    """aux [T, 128] from zero-padded mel frames [T + 4, 80]: same layout walk as the kernel (float32 accumulation via numpy matmul)."""
    CD, FEAT, KS = 128, 80, 5
    o = 0

    def take(n, shape):
        nonlocal o
        v = blob[o:o + n].reshape(shape)
        o += n
        return v

    W0 = take(KS * FEAT * CD, (KS, FEAT, CD))
    s0, b0 = take(CD, (CD,)), take(CD, (CD,))
    T = mel_pad_frames.shape[0] - (KS - 1)
    x = np.zeros((T, CD), dtype=np.float32)
    for c in range(KS):
        x += mel_pad_frames[c:c + T] @ W0[c]
    x = np.maximum(x * s0 + b0, 0)
    for _ in range(res_blocks):
        W1 = take(CD * CD, (CD, CD)); s1, b1 = take(CD, (CD,)), take(CD, (CD,))
        W2 = take(CD * CD, (CD, CD)); s2, b2 = take(CD, (CD,)), take(CD, (CD,))
        y = np.maximum((x @ W1) * s1 + b1, 0)
        x = ((y @ W2) * s2 + b2) + x
    Wo = take(CD * CD, (CD, CD)); bo = take(CD, (CD,))
    assert o == blob.size
    return x @ Wo + bo


def test_blob_replay_matches_the_torch_melresnet():
    m = WaveRNN(**synth.model_kwargs("RAW", "fatchord"))
    sd = synth.make_state("RAW", "fatchord", 0)
    g = torch.Generator().manual_seed(5)
    # non-trivial batch-norm statistics (make_state leaves the running stats at 0 / 1)
    for k in list(sd):
        if k.endswith("running_mean"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
        elif k.endswith("running_var"):
            sd[k] = torch.rand(sd[k].shape, generator=g) + 0.5
        elif "batch_norm" in k and k.endswith(".weight"):
            sd[k] = torch.rand(sd[k].shape, generator=g) + 0.5
        elif "batch_norm" in k and k.endswith(".bias"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
    m.load_state_dict(sd)
    m.eval()
    assert m.melresnet_native()
    blob = m.pack_melresnet()
    nb = len(m.upsample.resnet.layers)
    assert blob.dtype == np.float32 and blob.size == _lib.lib().wrnn_cond_blob_floats(nb)
    mel = synth.make_mel(37, seed=2)
    mp = torch.nn.functional.pad(mel, (2, 2))
    with torch.no_grad():
        want = m.upsample.resnet(mp)[0].t().numpy()
    got = replay(blob, mp[0].t().contiguous().numpy(), nb)
    assert got.shape == want.shape == (37, 128)
    assert np.abs(got - want).max() <= 2e-5 * max(1.0, np.abs(want).max())


def test_other_dimensions_fall_back_to_torch_on_the_gpu_not_to_the_kernel():
    kw = synth.model_kwargs("RAW", "ref")
    kw["compute_dims"] = 64
    assert not WaveRNN(**kw).melresnet_native()
