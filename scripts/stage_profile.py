"""Per-stage cycle breakdown of the persistent kernel (development aid; feeds profiles/)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from expressive_speech_synthesis_research_b200 import WaveRNN  # noqa: E402
from oracle import synth  # noqa: E402

SLOTS = ["before first poll: roles + loop + visit entry (4 gathers)", "SA poll+sample", "SA gru1+publish", "S2 gather(h1)", "S2 items", "S2 finalize",
         "S3 gather(h2)", "S3 items", "S3 finalize", "gather: poll rounds (count, 4 gathers)", "S4 gather(y1)", "S4 items", "S4 finalize",
         "S5 gather(y2)", "S5 items", "S5 finalize", "S2 deferred+barrier", "S3 deferred+barrier",
         "S4 deferred(cond)+barrier", "(unused)", "S2 wait own chunks (H1)", "S3 wait own chunks (H2)", "S4 wait own chunks (Y1)", "S5 wait own chunks (Y2)",
         "S2 before first poll", "S3 before first poll", "S4 before first poll", "S5 before first poll"]


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "RAW"
    dev = torch.device("cuda", 0)
    m = WaveRNN(**synth.model_kwargs(mode, "ref"))
    m.load_state_dict(synth.make_state(mode, "ref", 0))
    m.cuda()
    eng = m._engine(dev)
    S = 3000
    cases = [(8, None), (14, None), (20, None), (20, "1"), (64, None)]      # (folds, WRNN_FORCE_TEAMS cap or None)
    for B, force in cases:
        if force is None:
            os.environ.pop("WRNN_FORCE_TEAMS", None)
        else:
            os.environ["WRNN_FORCE_TEAMS"] = force
        L = S + 64
        mu = torch.rand(B * L, 80, device=dev)
        au = torch.randn(B * L, 128, device=dev)
        starts = np.arange(B, dtype=np.int64) * L
        for prof in (False, True):
            eng.stage_cycles(prof)
            for _ in range(2):
                m._run_folds(eng, dev, mu, au, starts, starts + L, S, None, 1, None, False)
            ms = eng.info().last_kernel_ms
            print("B=%d G=%d teams<=%s profiling=%s: %.3f ms, %.2f us/step" % (B, (B + 7) // 8, force or "3", prof, ms, ms * 1e3 / S), flush=True)
        cyc = eng.stage_cycles().astype(np.float64) / S
        G = (B + 7) // 8
        tot = cyc[:, :28].sum(1)
        print("  cycles/step (all groups): cta0 total %.0f  mean %.0f  max %.0f  => %.2f GHz effective" % (
            tot[0], tot.mean(), tot.max(), tot.mean() / (ms * 1e3 / S) / 1e3))
        for i, name in enumerate(SLOTS):
            print("  %-20s cta0 %7.0f  mean %7.0f  min %7.0f  max %7.0f   per group %6.0f" % (
                name, cyc[0, i], cyc[:, i].mean(), cyc[:, i].min(), cyc[:, i].max(), cyc[:, i].mean() / G))


if __name__ == "__main__":
    main()
