"""Multi-GPU check (run under torchrun, NCCL): folds of one utterance sharded over the ranks must reproduce the
single-GPU waveform bit for bit; prints one line per case on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from expressive_speech_synthesis_research_b200 import WaveRNN, distributed as D  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for mode, T, target, overlap, precision in (("RAW", 60, 1000, 100, "fp32"), ("MOL", 45, 700, 50, "fp32"), ("RAW", 30, 2000, 200, "fp32"),
                                                ("RAW", 400, 1000, 100, "bf16-dense")):
        sd = synth.make_state(mode, "ref", 0)
        m = WaveRNN(**synth.model_kwargs(mode, "ref"))
        m.load_state_dict(sd)
        m.cuda()
        m.precision = precision                       # bf16-dense: 73 folds, two or three clusters per rank (tcgen05 kernel)
        mel = synth.make_mel(T, seed=3)
        L = T * 200
        B = (L - overlap) // (target + overlap)
        if L - (B * (target + overlap) + overlap) != 0:
            B += 1
        S = target + 2 * overlap
        U = synth.make_uniforms(S, B, mode, seed=5)
        wav = D.generate_sharded(m, mel, target, overlap, True, uniforms=U, gather_to=0)
        if rank == 0:
            want = m.generate(mel, True, target, overlap, True, uniforms=U)
            same = wav.shape == want.shape and np.array_equal(wav, want)
            ok &= bool(same)
            print("sharded %s %s: %d folds over %d ranks, wave_len %d, bit-identical to single GPU: %s" % (mode, precision, B, world, want.size, same), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
