"""TEST INFRASTRUCTURE ONLY -- import harness for the *real* reference WaveRNN.

Imports `/root/reference/WaveRNN/models/fatchord_version.py` read-only, on CPU, in
the build container (the GPU box has no /root/reference; nothing here may be
called from `-m gpu` tests, smoke() or bench.py).  It is used by
`oracle/make_golden.py` to mint `tests/golden/*.npz` and by the CPU tests that
pin the C / torch restatements against the live reference when it is present.

Shims: stub matplotlib/librosa modules
(utility/display.py:2-6, utility/dsp.py:3), np.cumproduct (fatchord_version.py:65),
Tensor.cuda/Module.cuda -> identity (fatchord_version.py:122-123,162,173-175,208,
265,311; distribution.py:129-130), silence gen_display (fatchord_version.py:250).

Randomness injection so that the reference and the CUDA path consume the SAME
pre-drawn uniforms:
  RAW  Categorical.sample (fatchord_version.py:212-214) -> inverse CDF over the
       already re-normalised `probs`, one uniform per (step, fold).
  MOL  Tensor.uniform_ (distribution.py:106,118) -> a + (b-a)*u with u taken
       from U[step, fold, 0:10] (mixture Gumbel draws) then U[step, fold, 10].
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("WAVERNN_REFERENCE_ROOT", "/root/reference/WaveRNN")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "fatchord_version.py"))


_FV = None


def install_reference():
    """Import the reference model module with the four shims; returns the module."""
    global _FV
    if _FV is not None:
        return _FV
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot", "librosa", "librosa.filters", "librosa.output"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(np, "cumproduct"):
        np.cumproduct = np.cumprod
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    import models.fatchord_version as fv  # noqa: E402  (the reference itself)

    fv.stream = lambda msg: None
    _FV = fv
    return fv


# Reference hparams (WaveRNN/hparams.py:15-54) and the upstream-fatchord geometry
# that BASELINE.json's configs name (22.05 kHz, hop 275).
GEOMETRY = {
    "ref": dict(sample_rate=16000, hop_length=200, upsample_factors=(5, 5, 8)),
    "fatchord": dict(sample_rate=22050, hop_length=275, upsample_factors=(5, 5, 11)),
}


def model_kwargs(mode="RAW", geometry="ref", bits=9):
    g = GEOMETRY[geometry]
    return dict(rnn_dims=512, fc_dims=512, bits=bits, pad=2,
                upsample_factors=g["upsample_factors"], feat_dims=80, compute_dims=128,
                res_out_dims=128, res_blocks=10, hop_length=g["hop_length"],
                sample_rate=g["sample_rate"], mode=mode)


def build_reference_model(mode="RAW", geometry="ref", seed=0, bits=9):
    fv = install_reference()
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(open(os.devnull, "w")):
        model = fv.WaveRNN(**model_kwargs(mode, geometry, bits))
    model.eval()
    return model


class UniformFeed:
    """Pre-drawn uniforms U[S, B] (RAW) or U[S, B, 11] (MOL) consumed step by step."""

    def __init__(self, u):
        self.u = torch.as_tensor(u)
        self.i = 0


@contextlib.contextmanager
def injected_uniforms(feed: UniformFeed, mode: str):
    """Patch the reference's random draws so they read `feed` instead of the RNG."""
    if mode == "RAW":
        orig = torch.distributions.Categorical.sample

        def sample(self, sample_shape=torch.Size()):
            p = self.probs                              # p / p.sum(-1), Categorical.__init__
            c = torch.cumsum(p, dim=-1)
            u = feed.u[feed.i, : p.shape[0]].to(p.dtype).unsqueeze(-1)
            feed.i += 1
            return (c <= u).sum(-1).clamp_(max=p.shape[-1] - 1)

        torch.distributions.Categorical.sample = sample
        try:
            yield
        finally:
            torch.distributions.Categorical.sample = orig
    else:
        orig = torch.Tensor.uniform_
        state = {"phase": 0}

        def uniform_(self, a=0.0, b=1.0, generator=None):
            if not (a == 1e-5 and b == 1.0 - 1e-5):     # e.g. nn.GRUCell init inside get_gru_cell (:253)
                return orig(self, a, b, generator=generator)
            if state["phase"] == 0:                     # distribution.py:106  shape (1, B, 10)
                u = feed.u[feed.i, : self.shape[1], :10].reshape(self.shape)
                state["phase"] = 1
            else:                                       # distribution.py:118  shape (1, B)
                u = feed.u[feed.i, : self.shape[1], 10].reshape(self.shape)
                state["phase"] = 0
                feed.i += 1
            return self.copy_((a + (b - a) * u.to(torch.float64)).to(self.dtype))

        torch.Tensor.uniform_ = uniform_
        try:
            yield
        finally:
            torch.Tensor.uniform_ = orig


def reference_generate(model, mel, batched, target, overlap, mu_law, uniforms=None):
    """Run the reference's own generate() (fatchord_version.py:150-243) on CPU."""
    mode = model.mode
    if uniforms is None:
        return model.generate(mel, batched, target, overlap, mu_law)
    feed = UniformFeed(uniforms)
    with injected_uniforms(feed, mode):
        out = model.generate(mel, batched, target, overlap, mu_law)
    return out
