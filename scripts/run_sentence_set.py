"""BASELINE.json configs[3] and [4] end to end on ONE GPU with the dense (tcgen05) kernel (development aid; feeds profiles/):
  configs[3]  sentence set: 256 synthetic utterances of 2-12 s, folds pooled across utterances (generate_many)
  configs[4]  one 10-minute utterance (1 146 folds at 22.05 kHz), batched generate
Wall clock of the public call: host mels in, conditioning network, step loop, crossfade / mu-law epilogue, float64 waveforms out."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from expressive_speech_synthesis_research_b200 import WaveRNN  # noqa: E402
from bench import GEOMETRY, model_kwargs  # noqa: E402

TARGET, OVERLAP = 11000, 550


def main():
    n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    precision = sys.argv[2] if len(sys.argv) > 2 else "bf16-dense"
    geometry = "fatchord"
    sr, hop, _ = GEOMETRY[geometry]
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = WaveRNN(**model_kwargs("RAW", geometry)).to(dev)
    m.precision = precision
    warm = torch.rand(1, 80, 200, generator=torch.Generator().manual_seed(0))
    m.generate(warm, True, TARGET, OVERLAP, True, seed=1)
    durs = np.random.default_rng(0).uniform(2, 12, 256)[:n_utt]
    mels = [torch.rand(1, 80, int(round(d * sr / hop)) + 1, generator=torch.Generator().manual_seed(10 + i)) for i, d in enumerate(durs)]
    for attempt in ("first sight of every utterance length (cuDNN plans the conditioning network per shape)", "second pass, same lengths"):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        wavs = m.generate_many(mels, TARGET, OVERLAP, True, seed=1)
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        st = dict(m.last_stats)
        total = sum(w.size for w in wavs)
        print(json.dumps({"config": "configs[3] sentence set, %d utterances pooled" % n_utt, "pass": attempt, "precision": precision, "geometry": geometry, "folds": st["folds"],
                          "samples": int(total), "audio_seconds": total / sr, "wall_s": t, "samples_per_s": total / t, "rtf": t / (total / sr),
                          "step_loop_ms": st["kernel_ms"], "fold_steps_per_us": st["folds"] * st["steps"] / (st["kernel_ms"] * 1e3)}), flush=True)
    del wavs, mels
    torch.cuda.empty_cache()
    if n_utt == 256:
        T = int(round(600.0 * sr / hop)) + 1
        mel = torch.rand(1, 80, T, generator=torch.Generator().manual_seed(3))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        wav = m.generate(mel, True, TARGET, OVERLAP, True, seed=2)
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        st = dict(m.last_stats)
        print(json.dumps({"config": "configs[4] 10-minute utterance on one GPU", "precision": precision, "geometry": geometry, "folds": st["folds"],
                          "samples": int(wav.size), "audio_seconds": wav.size / sr, "wall_s": t, "samples_per_s": wav.size / t, "rtf": t / (wav.size / sr),
                          "step_loop_ms": st["kernel_ms"], "fold_steps_per_us": st["folds"] * st["steps"] / (st["kernel_ms"] * 1e3)}), flush=True)


if __name__ == "__main__":
    main()
