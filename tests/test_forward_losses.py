"""CPU tests of the training-side twin: WaveRNN.forward (fatchord_version.py:119-148) against the
teacher-forced goldens minted from the live reference, and discretized_mix_logistic_loss (utility/distribution.py:16-84)
against tests/golden/losses.npz.  PyTorch only: these do not touch the CUDA library."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from expressive_speech_synthesis_research_b200 import WaveRNN
from expressive_speech_synthesis_research_b200.distribution import discretized_mix_logistic_loss, loss_for_mode
from oracle import synth
from tests import helpers as H


@pytest.mark.parametrize("mode", ["RAW", "MOL"])
def test_forward_matches_the_reference_forward_goldens(mode):
    g = H.load_golden("teacher_forced.npz")
    sd = H.state_for(mode, "ref", H.digest_of(g, mode + "_digest"))
    m = WaveRNN(**synth.model_kwargs(mode, "ref"))
    m.load_state_dict(sd)
    m.eval()
    step0 = m.get_step()
    with torch.no_grad():
        logits = m(torch.from_numpy(g[mode + "_x"]), torch.from_numpy(g[mode + "_mel"]))
    assert m.get_step() == step0 + 1                                 # forward counts training steps (:120)
    assert logits.shape == (3, 400, 512 if mode == "RAW" else 30)
    got = logits.numpy()[:, g[mode + "_steps"], :]
    assert np.abs(got - g[mode + "_logits"]).max() <= 1e-5           # same ATen ops, same container: ~1e-7 in practice


def test_forward_is_trainable_like_train_wavernn():
    """train_wavernn.py:33-47 for both modes: loss.backward() reaches every step-path parameter."""
    for mode in ("RAW", "MOL"):
        m = WaveRNN(**synth.model_kwargs(mode, "ref"))
        m.train()
        x = torch.rand(2, 2 * 200) * 2 - 1
        mel = torch.rand(2, 80, 2 + 4)
        y_hat = m(x, mel)
        if mode == "RAW":
            y = torch.randint(0, 512, (2, 400))
            loss = loss_for_mode(mode)(y_hat.transpose(1, 2).unsqueeze(-1), y.unsqueeze(-1))
        else:
            loss = loss_for_mode(mode)(y_hat, (torch.rand(2, 400) * 2 - 1).unsqueeze(-1))
        loss.backward()
        for name in ("I.weight", "rnn1.weight_hh_l0", "rnn2.weight_ih_l0", "fc1.weight", "fc2.weight", "fc3.bias",
                     "upsample.resnet.conv_in.weight"):
            grad = dict(m.named_parameters())[name].grad
            assert grad is not None and torch.isfinite(grad).all() and grad.abs().sum() > 0, (mode, name)


def test_mixture_loss_matches_the_reference():
    g = H.load_golden("losses.npz")
    y_hat, y = torch.from_numpy(g["mol_y_hat"]), torch.from_numpy(g["mol_y"])
    assert abs(discretized_mix_logistic_loss(y_hat, y).item() - float(g["mol_loss"])) <= 1e-6
    got = discretized_mix_logistic_loss(y_hat, y, reduce=False).numpy()
    assert got.shape == g["mol_loss_unreduced"].shape
    assert np.abs(got - g["mol_loss_unreduced"]).max() <= 1e-5
    assert loss_for_mode("RAW") is F.cross_entropy
    with pytest.raises(ValueError):
        loss_for_mode("XYZ")
