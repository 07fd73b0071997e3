"""BASELINE.json configs[3] and [4] across the GPUs of one box (run under torchrun, NCCL), dense kernel (development aid; feeds profiles/):
  configs[3]  256 synthetic utterances of 2-12 s: whole utterances go to ranks by longest-processing-time on their fold counts
              (distributed.plan_utterances, no communication); every rank pools its utterances with generate_many
  configs[4]  one 10-minute utterance: contiguous fold ranges per rank, one all_gather of the overlap edges (distributed.generate_sharded)
Time = max over ranks of the wall clock of the public call, between barriers; rank 0 prints one JSON line per configuration."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from expressive_speech_synthesis_research_b200 import WaveRNN, _lib, distributed as D  # noqa: E402
from bench import GEOMETRY, model_kwargs  # noqa: E402

TARGET, OVERLAP = 11000, 550


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    precision = sys.argv[1] if len(sys.argv) > 1 else "bf16-dense"
    sr, hop, _ = GEOMETRY["fatchord"]
    torch.manual_seed(0)
    m = WaveRNN(**model_kwargs("RAW", "fatchord")).to(dev)
    m.precision = precision
    m.generate(torch.rand(1, 80, 200, generator=torch.Generator().manual_seed(0)), True, TARGET, OVERLAP, True, seed=1)

    def timed(fn):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out

    durs = np.random.default_rng(0).uniform(2, 12, 256)
    mels = [torch.rand(1, 80, int(round(d * sr / hop)) + 1, generator=torch.Generator().manual_seed(10 + i)) for i, d in enumerate(durs)]
    folds = [_lib.fold_index(x.size(-1) * hop, TARGET, OVERLAP)[0] for x in mels]
    mine = D.plan_utterances(folds, world)[rank]
    for attempt in ("first call", "second call"):
        t, wavs = timed(lambda: m.generate_many([mels[i] for i in mine], TARGET, OVERLAP, True, seed=1 + rank))
        n = torch.tensor([sum(w.size for w in wavs), sum(folds[i] for i in mine)], device=dev, dtype=torch.float64)
        per_rank = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(per_rank, n)
        if rank == 0:
            total = sum(float(p[0]) for p in per_rank)
            print(json.dumps({"config": "configs[3] sentence set, 256 utterances over %d GPUs (whole utterances per rank, LPT)" % world, "pass": attempt,
                              "precision": precision, "folds_per_rank": [int(p[1]) for p in per_rank], "samples": int(total), "wall_s": t,
                              "samples_per_s": total / t, "rtf": t / (total / sr)}), flush=True)
        del wavs
    T = int(round(600.0 * sr / hop)) + 1
    mel = torch.rand(1, 80, T, generator=torch.Generator().manual_seed(3))
    for attempt in ("first call", "second call"):
        t, wav = timed(lambda: D.generate_sharded(m, mel, TARGET, OVERLAP, True, seed=2, gather_to=0))
        if rank == 0:
            print(json.dumps({"config": "configs[4] 10-minute utterance, folds sharded over %d GPUs" % world, "pass": attempt, "precision": precision,
                              "folds": int(_lib.fold_index(T * hop, TARGET, OVERLAP)[0]), "samples": int(wav.size), "wall_s": t,
                              "samples_per_s": wav.size / t, "rtf": t / (wav.size / sr)}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
