"""Measure the BASELINE.json configurations other than the bench.py headline on one GPU (development aid;
feeds profiles/).  Prints one JSON line per configuration: useful samples/s, RTF, kernel us/step.

  configs[0]  RAW 9-bit unbatched, 2 s utterance (pure step-latency floor, B = 1)
  configs[1]  RAW 9-bit batched 10 s (bench.py headline), both geometries
  configs[2]  MOL batched 10 s
  configs[3]  sentence set: N utterances of 2-12 s, folds pooled across utterances (generate_many)
"""
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


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


def main():
    n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device("cuda", 0)
    for geometry in ("fatchord", "ref"):
        sr, hop, _ = GEOMETRY[geometry]
        for mode in ("RAW", "MOL"):
            torch.manual_seed(0)
            m = WaveRNN(**model_kwargs(mode, geometry)).to(dev)
            T10 = int(round(10.0 * sr / hop)) + 1
            mel10 = torch.rand(1, 80, T10, generator=torch.Generator().manual_seed(0))
            t, wav = timed(lambda: m.generate(mel10, True, TARGET, OVERLAP, True, seed=1))
            st = dict(m.last_stats)
            print(json.dumps({"config": "batched 10 s", "mode": mode, "geometry": geometry, "folds": st["folds"], "wave_len": int(wav.size),
                              "samples_per_s": wav.size / t, "rtf": t / (wav.size / sr), "kernel_us_per_step": st["kernel_ms"] * 1e3 / st["steps"]}), flush=True)
            if mode == "RAW":
                T2 = int(round(2.0 * sr / hop)) + 1
                mel2 = torch.rand(1, 80, T2, generator=torch.Generator().manual_seed(1))
                t, wav = timed(lambda: m.generate(mel2, False, TARGET, OVERLAP, True, seed=1), reps=1)
                st = dict(m.last_stats)
                print(json.dumps({"config": "unbatched 2 s", "mode": mode, "geometry": geometry, "folds": 1, "wave_len": int(wav.size),
                                  "samples_per_s": wav.size / t, "rtf": t / (wav.size / sr), "kernel_us_per_step": st["kernel_ms"] * 1e3 / st["steps"]}), flush=True)
                durs = np.random.default_rng(0).uniform(2, 12, 256)[:n_utt]
                mels = [torch.rand(1, 80, int(round(d * sr / hop)) + 1, generator=torch.Generator().manual_seed(10 + i)) for i, d in enumerate(durs)]
                t, wavs = timed(lambda: m.generate_many(mels, TARGET, OVERLAP, True, seed=1), reps=1)
                st = dict(m.last_stats)
                total = sum(w.size for w in wavs)
                print(json.dumps({"config": "sentence set, %d utterances pooled" % n_utt, "mode": mode, "geometry": geometry, "folds": st["folds"],
                                  "wave_len": int(total), "samples_per_s": total / t, "rtf": t / (total / sr),
                                  "kernel_ms_last_launch_group": st["kernel_ms"]}), flush=True)


if __name__ == "__main__":
    main()
