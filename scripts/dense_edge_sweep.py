"""Edge-case sweep of the dense kernel (development aid): fold counts around cluster / wave boundaries, very short runs, both
conditioning modes; every label must be the inverse-CDF outcome of the kernel's own logits and no watchdog may fire."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from expressive_speech_synthesis_research_b200 import WaveRNN  # noqa: E402
from oracle import synth  # noqa: E402
from scripts.dense_bringup import run_folds  # noqa: E402


def main():
    m = WaveRNN(**synth.model_kwargs("RAW", "ref"))
    m.load_state_dict(synth.make_state("RAW", "ref", 0))
    m = m.cuda()
    m.precision = "bf16-dense"
    rng = np.random.default_rng(1)
    bad = 0
    for B, S in ((1, 1), (1, 2), (1, 3), (2, 5), (7, 17), (31, 9), (32, 9), (33, 9), (64, 4), (65, 4), (100, 6), (479, 3), (480, 3), (481, 3), (961, 2), (1500, 2)):
        mels = rng.uniform(0, 1, (B, S, 80)).astype(np.float32)
        aux = rng.normal(0, 1, (B, S, 128)).astype(np.float32)
        U = rng.uniform(0, 1, (S, B)).astype(np.float32)
        r = run_folds(m, mels, aux, U, logits=True)
        lg = r["logits"].astype(np.float64)
        p = np.exp(lg - lg.max(-1, keepdims=True))
        cdf = np.cumsum(p, -1)
        k = np.minimum((cdf <= (U[..., None].astype(np.float64) * cdf[..., -1:])).sum(-1), 511)
        mism = int((k.T != r["labels"]).sum())
        fin = bool(np.isfinite(r["logits"]).all())
        bad += (mism > 0) + (not fin)
        print("B=%4d S=%2d: label mismatches %d / %d, logits finite %s, kernel %.3f ms" % (B, S, mism, k.size, fin, r["ms"]), flush=True)
    # generate(): shortest legal utterance (21 frames), unbatched and batched, frames mode
    for T, batched in ((21, False), (21, True), (23, True)):
        wav = m.generate(synth.make_mel(T, seed=T), batched, 500, 50, True, seed=1)
        ok = wav.shape == ((T - 1) * 200,) and np.isfinite(wav).all()
        bad += not ok
        print("generate T=%d batched=%s: %s" % (T, batched, "ok" if ok else "BAD"), flush=True)
    print("edge sweep", "OK" if bad == 0 else "FAILED (%d)" % bad)
    return bad


if __name__ == "__main__":
    sys.exit(main())
