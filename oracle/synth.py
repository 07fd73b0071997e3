"""TEST INFRASTRUCTURE ONLY -- deterministic synthetic weights / mels / uniforms.

Weights are drawn key by key from an explicit torch.Generator (CPU Philox/MT stream is
stable for a given torch version), NOT via module construction order, so the reference
model, the oracle and the product all load the very same state_dict.  Shapes follow the
reference's state_dict (fatchord_version.py:10-117).
"""
import hashlib
import math

import numpy as np
import torch

GEOMETRY = {
    "ref": dict(sample_rate=16000, hop_length=200, upsample_factors=(5, 5, 8)),        # hparams.py:15,19,34
    "fatchord": dict(sample_rate=22050, hop_length=275, upsample_factors=(5, 5, 11)),  # BASELINE.json configs
}


def model_kwargs(mode="RAW", geometry="ref", bits=9):
    g = GEOMETRY[geometry]
    return dict(rnn_dims=512, fc_dims=512, bits=bits, pad=2, upsample_factors=g["upsample_factors"],
                feat_dims=80, compute_dims=128, res_out_dims=128, res_blocks=10,
                hop_length=g["hop_length"], sample_rate=g["sample_rate"], mode=mode)


def state_shapes(mode="RAW", geometry="ref", bits=9):
    kw = model_kwargs(mode, geometry, bits)
    R, FCD, F, CD, RO = kw["rnn_dims"], kw["fc_dims"], kw["feat_dims"], kw["compute_dims"], kw["res_out_dims"]
    A = RO // 4
    C = 2 ** bits if mode == "RAW" else 30
    k = 2 * kw["pad"] + 1
    shapes = [("step", (1,)), ("upsample.resnet.conv_in.weight", (CD, F, k))]
    shapes += _bn_shapes("upsample.resnet.batch_norm", CD)
    for i in range(kw["res_blocks"]):
        p = "upsample.resnet.layers.%d" % i
        shapes += [(p + ".conv1.weight", (CD, CD, 1)), (p + ".conv2.weight", (CD, CD, 1))]
        shapes += _bn_shapes(p + ".batch_norm1", CD) + _bn_shapes(p + ".batch_norm2", CD)
    shapes += [("upsample.resnet.conv_out.weight", (RO, CD, 1)), ("upsample.resnet.conv_out.bias", (RO,))]
    for i, s in enumerate(kw["upsample_factors"]):
        shapes.append(("upsample.up_layers.%d.weight" % (2 * i + 1), (1, 1, 1, 2 * s + 1)))
    shapes += [("I.weight", (R, F + A + 1)), ("I.bias", (R,))]
    for name, inp in (("rnn1", R), ("rnn2", R + A)):
        shapes += [(name + ".weight_ih_l0", (3 * R, inp)), (name + ".weight_hh_l0", (3 * R, R)),
                   (name + ".bias_ih_l0", (3 * R,)), (name + ".bias_hh_l0", (3 * R,))]
    shapes += [("fc1.weight", (FCD, R + A)), ("fc1.bias", (FCD,)),
               ("fc2.weight", (FCD, FCD + A)), ("fc2.bias", (FCD,)),
               ("fc3.weight", (C, FCD)), ("fc3.bias", (C,))]
    return shapes


def _bn_shapes(prefix, n):
    return [(prefix + ".weight", (n,)), (prefix + ".bias", (n,)), (prefix + ".running_mean", (n,)),
            (prefix + ".running_var", (n,)), (prefix + ".num_batches_tracked", ())]


def make_state(mode="RAW", geometry="ref", seed=0, bits=9, gain=1.0):
    """A full reference-compatible state_dict with non-trivial BN statistics.

    Linear/GRU/conv tensors ~ U(+-gain/sqrt(fan_in)) (torch's default family); BN affine
    weight ~ U(0.5,1.5), bias/mean ~ U(-0.1,0.1), var ~ U(0.5,1.5); box-filter up-convs
    = 1/(2s+1) * U(0.9,1.1) (the reference fills them with 1/(2s+1) at init, :75, but they
    are trainable).
    """
    g = torch.Generator().manual_seed(1000003 * seed + (17 if mode == "MOL" else 0))
    sd = {}
    for key, shape in state_shapes(mode, geometry, bits):
        if key == "step":
            sd[key] = torch.zeros(1, dtype=torch.long)
        elif key.endswith("num_batches_tracked"):
            sd[key] = torch.tensor(0, dtype=torch.long)
        elif key.endswith("running_var") or (".batch_norm" in key and key.endswith(".weight")):
            sd[key] = 0.5 + torch.rand(shape, generator=g)
        elif ".batch_norm" in key:
            sd[key] = (torch.rand(shape, generator=g) - 0.5) * 0.2
        elif ".up_layers." in key:
            sd[key] = (0.9 + 0.2 * torch.rand(shape, generator=g)) / shape[-1]
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else None
            if key.startswith("rnn") or fan_in is None:
                fan = 512 if key.startswith("rnn") else {"I.bias": 113, "fc1.bias": 544, "fc2.bias": 544,
                                                         "fc3.bias": 512}.get(key, 128)
                bound = gain / math.sqrt(fan)
            else:
                bound = gain / math.sqrt(fan_in)
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


def state_digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def make_mel(T, seed=0, feat=80):
    """(1, feat, T) float32 in [0,1) -- the range of (mel+4)/8, synthesize_sentences.py:54."""
    return torch.rand(1, feat, T, generator=torch.Generator().manual_seed(seed))


def make_uniforms(S, B, mode, seed=123):
    shape = (S, B) if mode == "RAW" else (S, B, 11)
    return torch.rand(*shape, generator=torch.Generator().manual_seed(seed))


def frames_for_seconds(seconds, geometry):
    g = GEOMETRY[geometry]
    return int(round(seconds * g["sample_rate"] / g["hop_length"])) + 1
