"""Staged GPU bring-up checks (development aid; prints rather than asserts so one call shows everything)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from expressive_speech_synthesis_research_b200 import WaveRNN  # noqa: E402
from oracle import c_oracle, synth  # noqa: E402
from tests import helpers as H  # noqa: E402


def stage(name):
    print("\n=== %s ===" % name, flush=True)


def main():
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).multi_processor_count, "SMs")
    dev = torch.device("cuda", 0)
    for mode in ("RAW", "MOL"):
        sd = synth.make_state(mode, "ref", 0)
        m = WaveRNN(**synth.model_kwargs(mode, "ref"))
        m.load_state_dict(sd)
        m.cuda()
        t0 = time.time()
        eng = m._engine(dev)
        print(mode, "engine + weight repack %.2fs" % (time.time() - t0), "smem", eng.info().smem_bytes)
        if mode == "RAW":
            stage("exchange probe")
            for it in (200, 2000):
                print("iters", it, "usec/exchange", eng.measure_exchange(it), flush=True)
        for (B, S) in ((1, 3), (3, 40), (8, 64), (11, 64), (20, 200)):
            stage("%s teacher-forced B=%d S=%d" % (mode, B, S))
            rng = np.random.default_rng(B * 100 + S)
            mels = rng.uniform(0, 1, (B, S, 80)).astype(np.float32)
            aux = rng.normal(0, 1, (B, S, 128)).astype(np.float32)
            forced = rng.uniform(-1, 1, (S, B)).astype(np.float32)
            U = synth.make_uniforms(S, B, mode, seed=1).numpy()
            want = c_oracle.generate_folds(sd, mode, mels, aux, U, forced_x=forced, want_logits=True, precision="fp64")
            mu = torch.as_tensor(mels).reshape(B * S, 80).cuda()
            au = torch.as_tensor(aux).reshape(B * S, 128).cuda()
            starts = np.arange(B, dtype=np.int64) * S
            try:
                t0 = time.time()
                r = m._run_folds(eng, dev, mu, au, starts, starts + S, S, U, 0, forced, True)
                torch.cuda.synchronize()
                lg = r["logits"].cpu().numpy()
                err = np.abs(lg - want["logits"]).max(axis=(1, 2))
                print("kernel ms %.3f (%.2f us/step); logits max err per step (first 6):" % (eng.info().last_kernel_ms, eng.info().last_kernel_ms * 1e3 / S),
                      np.array2string(err[:6], precision=2), "overall %.3g" % err.max())
                if mode == "RAW":
                    lab = r["labels"].cpu().numpy()
                    print("labels equal to oracle: %d / %d" % ((lab == want["labels"]).sum(), lab.size))
                else:
                    print("samples max diff %.3g; mix idx equal %d / %d" % (
                        np.abs(r["samples"].cpu().numpy() - want["samples"]).max(),
                        (r["labels"].cpu().numpy() == want["mix"]).sum(), want["mix"].size))
            except Exception as e:
                print("FAILED:", repr(e), flush=True)
                return 1
    stage("free-running smoke")
    import __graft_entry__ as ge
    ge.smoke()
    return 0


if __name__ == "__main__":
    sys.exit(main())
