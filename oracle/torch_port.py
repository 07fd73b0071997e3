"""TEST INFRASTRUCTURE ONLY -- torch-CPU restatement of the reference's generate().

A functional re-statement (state_dict in, waveform out) that issues the same ATen
ops in the same order as /root/reference/WaveRNN/models/fatchord_version.py, so on
the same host/threads it reproduces the reference's fp32 results and its CPU cost
profile (8 addmm + 2 gru_cell + softmax/sampling per step).  It is what
`bench.py --impl reference` and the `cpu_baseline` leg time on the GPU box, where
/root/reference does not exist; it also provides the conditioning-network oracle
(`upsample`) for the product's PyTorch UpsampleNetwork.

Pinned against the goldens minted from the live reference (oracle/make_golden.py ->
tests/golden/*.npz) by tests/test_oracle_golden.py.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _bn(x, sd, prefix, eps=1e-5):
    # nn.BatchNorm1d in eval mode -- fatchord_version.py:15-16,33 (generate() calls self.eval(), :154)
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], False, 0.0, eps)


def mel_resnet(sd, m, res_blocks):
    """MelResNet.forward -- fatchord_version.py:39-45.  m: (1, feat, T+2*pad)"""
    x = F.conv1d(m, sd["upsample.resnet.conv_in.weight"])
    x = F.relu(_bn(x, sd, "upsample.resnet.batch_norm"))
    for i in range(res_blocks):
        p = "upsample.resnet.layers.%d" % i
        r = x
        x = F.conv1d(x, sd[p + ".conv1.weight"])
        x = F.relu(_bn(x, sd, p + ".batch_norm1"))
        x = F.conv1d(x, sd[p + ".conv2.weight"])
        x = _bn(x, sd, p + ".batch_norm2") + r
    return F.conv1d(x, sd["upsample.resnet.conv_out.weight"], sd["upsample.resnet.conv_out.bias"])


def upsample(sd, m, upsample_factors, pad, res_blocks=None):
    """UpsampleNetwork.forward -- fatchord_version.py:79-86.  m: (1, feat, T+2*pad) ->
    (mels (1, L, feat), aux (1, L, res_out)), L = T*hop."""
    if res_blocks is None:
        res_blocks = sum(1 for k in sd if k.startswith("upsample.resnet.layers.") and k.endswith(".conv1.weight"))
    total = int(np.prod(upsample_factors))
    aux = mel_resnet(sd, m, res_blocks)
    aux = aux.repeat_interleave(total, dim=-1)                       # Stretch2d(total, 1) :48-58,80-82
    x = m.unsqueeze(1)
    for i, s in enumerate(upsample_factors):
        x = x.repeat_interleave(s, dim=-1)                          # Stretch2d(s, 1)
        x = F.conv2d(x, sd["upsample.up_layers.%d.weight" % (2 * i + 1)], padding=(0, s))   # :74
    indent = pad * total
    x = x.squeeze(1)[:, :, indent:-indent]                           # :85
    return x.transpose(1, 2), aux.transpose(1, 2)


def conditioning(sd, mels, upsample_factors, pad):
    """generate() prologue -- fatchord_version.py:162-165: zero-pad `pad` frames each side, upsample."""
    m = F.pad(mels, (pad, pad))                                      # pad_tensor(..., side='both') :260-270
    return upsample(sd, m, upsample_factors, pad)


def fold_with_overlap(x, target, overlap):
    """fatchord_version.py:272-319 (index arithmetic restated; values copied)."""
    _, total_len, feats = x.shape
    n = (total_len - overlap) // (target + overlap)
    remaining = total_len - (n * (overlap + target) + overlap)
    if remaining != 0:
        n += 1
        x = F.pad(x, (0, 0, 0, target + 2 * overlap - remaining))
    hop = target + overlap
    out = x.new_zeros(n, target + 2 * overlap, feats)
    for i in range(n):
        out[i] = x[0, i * hop:i * hop + target + 2 * overlap]
    return out


def xfade_and_unfold(y, overlap):
    """fatchord_version.py:321-383 (numpy float64)."""
    y = np.array(y, dtype=np.float64)
    n, length = y.shape
    target = length - 2 * overlap
    sil = overlap // 2
    t = np.linspace(-1, 1, overlap - sil, dtype=np.float64)
    fade_in = np.concatenate([np.zeros(sil), np.sqrt(0.5 * (1 + t))])
    fade_out = np.concatenate([np.sqrt(0.5 * (1 - t)), np.zeros(sil)])
    y[:, :overlap] *= fade_in
    y[:, -overlap:] *= fade_out
    out = np.zeros(n * (target + overlap) + overlap, dtype=np.float64)
    for i in range(n):
        s = i * (target + overlap)
        out[s:s + length] += y[i]
    return out


def decode_mu_law(y, n_classes):
    """utility/dsp.py:100-105, from_labels=False."""
    mu = n_classes - 1
    return np.sign(y) / mu * ((1 + mu) ** np.abs(y) - 1)


def _gru_cell(x, h, sd, name):
    return torch.gru_cell(x, h, sd[name + ".weight_ih_l0"], sd[name + ".weight_hh_l0"],
                          sd[name + ".bias_ih_l0"], sd[name + ".bias_hh_l0"])


def step_loop(sd, mode, mels, aux, uniforms=None, forced_x=None, want_logits=False, generator=None):
    """The hot loop -- fatchord_version.py:171-222.  mels (B,S,feat), aux (B,S,4d) on CPU.

    uniforms: injected draws ([S,B] RAW inverse-CDF / [S,B,11] MOL), else torch's own RNG
    (Categorical.sample / uniform_) exactly like the reference.
    """
    B, S, _ = mels.shape
    rnn = sd["rnn1.weight_hh_l0"].shape[1]
    d = aux.shape[2] // 4
    C = sd["fc3.weight"].shape[0]
    h1 = mels.new_zeros(B, rnn)
    h2 = mels.new_zeros(B, rnn)
    x = mels.new_zeros(B, 1)
    a = [aux[:, :, d * i:d * (i + 1)] for i in range(4)]
    out, logit_log = [], []
    with torch.no_grad():
        for i in range(S):
            x = torch.cat([x, mels[:, i, :], a[0][:, i, :]], dim=1)
            x = F.linear(x, sd["I.weight"], sd["I.bias"])
            h1 = _gru_cell(x, h1, sd, "rnn1")
            x = x + h1
            h2 = _gru_cell(torch.cat([x, a[1][:, i, :]], dim=1), h2, sd, "rnn2")
            x = x + h2
            x = F.relu(F.linear(torch.cat([x, a[2][:, i, :]], dim=1), sd["fc1.weight"], sd["fc1.bias"]))
            x = F.relu(F.linear(torch.cat([x, a[3][:, i, :]], dim=1), sd["fc2.weight"], sd["fc2.bias"]))
            logits = F.linear(x, sd["fc3.weight"], sd["fc3.bias"])
            if want_logits:
                logit_log.append(logits)
            if mode == "MOL":
                nr = C // 3
                lo, hi = 1e-5, 1.0 - 1e-5
                if uniforms is None:
                    u1 = logits.new_empty(B, nr).uniform_(lo, hi, generator=generator)
                    u2 = logits.new_empty(B).uniform_(lo, hi, generator=generator)
                else:
                    u = uniforms[i].to(torch.float64)
                    u1 = (lo + (hi - lo) * u[:, :nr]).to(logits.dtype)
                    u2 = (lo + (hi - lo) * u[:, nr]).to(logits.dtype)
                arg = (logits[:, :nr] - torch.log(-torch.log(u1))).argmax(dim=-1, keepdim=True)
                mean = logits[:, nr:2 * nr].gather(1, arg).squeeze(1)
                ls = torch.clamp(logits[:, 2 * nr:3 * nr].gather(1, arg).squeeze(1), min=float(np.log(1e-14)))
                sample = mean + torch.exp(ls) * (torch.log(u2) - torch.log(1. - u2))
                sample = torch.clamp(torch.clamp(sample, min=-1.), max=1.)
            else:
                post = F.softmax(logits, dim=1)
                if uniforms is None:
                    k = torch.distributions.Categorical(post).sample()
                else:
                    p = post / post.sum(-1, keepdim=True)
                    k = (torch.cumsum(p, dim=-1) <= uniforms[i].to(p.dtype).unsqueeze(-1)).sum(-1).clamp_(max=C - 1)
                sample = 2 * k.float() / (C - 1.) - 1.
            out.append(sample)
            x = (sample if forced_x is None else forced_x[i].to(sample.dtype)).unsqueeze(-1)
    samples = torch.stack(out).transpose(0, 1)
    return samples, (torch.stack(logit_log) if want_logits else None)


def generate(sd, mels, batched, target, overlap, mu_law, *, mode, upsample_factors, pad, hop_length,
             uniforms=None):
    """WaveRNN.generate -- fatchord_version.py:150-243, functional form."""
    mu_law = mu_law if mode == "RAW" else False
    with torch.no_grad():
        wave_len = (mels.size(-1) - 1) * hop_length
        m, aux = conditioning(sd, mels, upsample_factors, pad)
        if batched:
            m = fold_with_overlap(m, target, overlap)
            aux = fold_with_overlap(aux, target, overlap)
        samples, _ = step_loop(sd, mode, m, aux, uniforms=uniforms)
    out = samples.cpu().numpy().astype(np.float64)
    out = xfade_and_unfold(out, overlap) if batched else out[0]
    if mu_law:
        out = decode_mu_law(out, sd["fc3.weight"].shape[0])
    fade = np.linspace(1, 0, 20 * hop_length)
    out = out[:wave_len]
    out[-20 * hop_length:] *= fade
    return out
