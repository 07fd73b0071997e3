"""Training-side losses of the reference's vocoder, PyTorch.

`discretized_mix_logistic_loss` follows WaveRNN/utility/distribution.py:16-84 (itself adapted from r9y9's wavenet_vocoder
mixture.py): y_hat (B, T, 3*K) mixture parameters from WaveRNN.forward in MOL mode, y (B, T, 1) targets in [-1, 1].
train_wavernn.py:28-30 selects it for MOL and F.cross_entropy for RAW (`loss_for_mode`).  The generation-side sampler
(distribution.py:87-123) lives in the CUDA kernels (csrc/wavernn_wide.cuh, sampler_body)."""
import math

import torch
import torch.nn.functional as F


def log_sum_exp(x):
    """log(sum(exp(x))) over the last axis, shifted by the maximum (distribution.py:6-12)."""
    m = x.max(dim=-1, keepdim=True).values
    return m.squeeze(-1) + torch.log(torch.exp(x - m).sum(dim=-1))


def discretized_mix_logistic_loss(y_hat, y, num_classes=65536, log_scale_min=None, reduce=True):
    if log_scale_min is None:
        log_scale_min = float(math.log(1e-14))
    if y_hat.dim() != 3 or y_hat.size(-1) % 3 != 0:
        raise ValueError("y_hat must be (B, T, 3 * nr_mix), got %s" % (tuple(y_hat.shape),))
    k = y_hat.size(-1) // 3
    # the reference permutes to (B, C, T) and straight back (:21-27); parameters are read from the last axis
    logit_probs, means = y_hat[..., :k], y_hat[..., k:2 * k]
    log_scales = torch.clamp(y_hat[..., 2 * k:3 * k], min=log_scale_min)
    y = y.expand_as(means)

    half_bin = 1. / (num_classes - 1)
    centered = y - means
    inv_std = torch.exp(-log_scales)
    plus_in = inv_std * (centered + half_bin)
    min_in = inv_std * (centered - half_bin)
    cdf_delta = torch.sigmoid(plus_in) - torch.sigmoid(min_in)          # mass of the bin around y
    log_cdf_plus = plus_in - F.softplus(plus_in)                        # left edge bin (y < -0.999)
    log_one_minus_cdf_min = -F.softplus(min_in)                         # right edge bin (y > 0.999)
    mid_in = inv_std * centered
    log_pdf_mid = mid_in - log_scales - 2. * F.softplus(mid_in)         # density at the bin centre, for vanishing bins

    wide = (cdf_delta > 1e-5).float()
    inner = wide * torch.log(torch.clamp(cdf_delta, min=1e-12)) + (1. - wide) * (log_pdf_mid - math.log((num_classes - 1) / 2))
    right = (y > 0.999).float()
    inner = right * log_one_minus_cdf_min + (1. - right) * inner
    left = (y < -0.999).float()
    log_probs = left * log_cdf_plus + (1. - left) * inner
    log_probs = log_probs + F.log_softmax(logit_probs, -1)
    if reduce:
        return -torch.mean(log_sum_exp(log_probs))
    return -log_sum_exp(log_probs).unsqueeze(-1)


def loss_for_mode(mode):
    """train_wavernn.py:28-30."""
    if mode == 'RAW':
        return F.cross_entropy
    if mode == 'MOL':
        return discretized_mix_logistic_loss
    raise ValueError("unknown mode %r" % (mode,))
