"""Bring-up of precision "bf16-dense" on a B200 (development tool): teacher-forced logits against the fp64 oracle and
the numpy replay of the packed stream, sampling consistency, then step timing at several fold counts."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from expressive_speech_synthesis_research_b200 import WaveRNN  # noqa: E402
from oracle import c_oracle, synth  # noqa: E402
from tests.dense_replay import DenseReplay  # noqa: E402


def run_folds(m, mels_f, aux_f, U, forced=None, logits=False, seed=0):
    dev = torch.device("cuda", 0)
    B, S, _ = mels_f.shape
    mu = torch.as_tensor(mels_f).reshape(B * S, -1).contiguous().to(dev)
    au = torch.as_tensor(aux_f).reshape(B * S, -1).contiguous().to(dev)
    starts = np.arange(B, dtype=np.int64) * S
    eng = m._engine(dev)
    r = m._run_folds(eng, dev, mu, au, starts, starts + S, S, U, seed, forced, logits)
    torch.cuda.synchronize()
    out = {k: (v.cpu().numpy() if v is not None else None) for k, v in r.items()}
    out["ms"] = eng.info().last_kernel_ms
    return out


def main():
    quick = "--quick" in sys.argv
    sd = synth.make_state("RAW", "ref", 0)
    m = WaveRNN(**synth.model_kwargs("RAW", "ref"))
    m.load_state_dict(sd)
    m = m.cuda()
    m.precision = "bf16-dense"
    rng = np.random.default_rng(7)
    res = {}
    for B, S in ((5, 12), (37, 40)):
        mels = rng.uniform(0, 1, (B, S, 80)).astype(np.float32)
        aux = rng.normal(0, 1, (B, S, 128)).astype(np.float32)
        forced = rng.uniform(-1, 1, (S, B)).astype(np.float32)
        U = rng.uniform(0, 1, (S, B)).astype(np.float32)
        r = run_folds(m, mels, aux, U, forced=forced, logits=True)
        want = c_oracle.generate_folds(sd, "RAW", mels, aux, U, forced_x=forced, want_logits=True, precision="fp64")["logits"]
        err = float(np.abs(r["logits"] - want).max())
        rep = DenseReplay(sd).run(mels.astype(np.float64), aux.astype(np.float64), forced, round_act=True) if B <= 8 else None
        err_rep = float(np.abs(r["logits"] - rep).max()) if rep is not None else None
        # labels must be the inverse-CDF outcome of the kernel's own logits
        lg = r["logits"].astype(np.float64)
        p = np.exp(lg - lg.max(-1, keepdims=True))
        cdf = np.cumsum(p, -1)
        k = np.minimum((cdf <= (U[..., None].astype(np.float64) * cdf[..., -1:])).sum(-1), 511)
        mism = int((k.T != r["labels"]).sum())
        print("teacher-forced B=%d S=%d: max|logits - oracle| %.3e, vs replay %s, label mismatches vs own logits %d / %d, kernel %.3f ms"
              % (B, S, err, err_rep, mism, k.size, r["ms"]), flush=True)
        res["tf_%d_%d" % (B, S)] = dict(err_oracle=err, err_replay=err_rep, label_mismatch=mism)
    if quick:
        return
    # free running + timing
    for B, S in ((20, 2000), (32, 2000), (64, 2000), (512, 2000), (1024, 1000)):
        mels = rng.uniform(0, 1, (B, S, 80)).astype(np.float32)
        aux = rng.normal(0, 1, (B, S, 128)).astype(np.float32)
        r = run_folds(m, mels, aux, None, seed=3)
        r = run_folds(m, mels, aux, None, seed=3)
        lab = r["labels"]
        print("free-running B=%d S=%d: kernel %.2f ms = %.2f us/step, %.2f fold-steps/us; label range %d..%d, distinct %d"
              % (B, S, r["ms"], 1e3 * r["ms"] / S, B * S / (1e3 * r["ms"]), lab.min(), lab.max(), len(np.unique(lab))), flush=True)
        res["free_%d" % B] = dict(ms=r["ms"], us_per_step=1e3 * r["ms"] / S, fold_steps_per_us=B * S / (1e3 * r["ms"]))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
