"""Per-stage cycle breakdown of the WIDE kernel (csrc/wavernn_wide.cuh), thread 0 of every worker CTA (development aid; feeds profiles/)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from expressive_speech_synthesis_research_b200 import WaveRNN  # noqa: E402
from oracle import synth  # noqa: E402

SLOTS = ["SA wait x (sampler hop)", "SA gru1 + publish H1", "S2 gather H1", "S2 pass Wih2x + A + store + B", "S2 sum + C + gru2 + publish H2",
         "S2 deferred (Whh1, Wfc1x) + D + finalize", "S3 deferred: warp 0 pass (Whh2)", "S3 pass Wfc1x + A + store + B", "S3 sum + C + fc1 + publish Y1",
         "S3 deferred: D + finalize", "S3 deferred: fold halves + store", "S4 pass Wfc2 + A + store + B", "S4 sum + fc2 + publish Y2",
         "cond: D + finalize + E", "S3 deferred: gather finish (Y1)", "S5 pass Wfc3 + A + store + B", "S5 sum + publish logits", "cond: wait for the TMA rows", "cond: pass", "S5 finalize: sum of 16 partials", "S5 gather Y2", "-", "-", "S5 finalize: bar96"]


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "RAW"
    dev = torch.device("cuda", 0)
    m = WaveRNN(**synth.model_kwargs(mode, "ref"))
    m.load_state_dict(synth.make_state(mode, "ref", 0))
    m.cuda()
    eng = m._engine(dev)
    S = 3000
    print("exchange probe: %.3f us" % eng.measure_exchange(2000))
    for B in (1, 8, 14, 20, 21):
        L = S + 64
        mu = torch.rand(B * L, 80, device=dev)
        au = torch.randn(B * L, 128, device=dev)
        starts = np.arange(B, dtype=np.int64) * L
        for prof in (False, True):
            eng.stage_cycles(prof)
            for _ in range(2):
                m._run_folds(eng, dev, mu, au, starts, starts + L, S, None, 1, None, False)
            info = eng.info()
            ms = info.last_kernel_ms
            print("B=%d kernel_kind=%d profiling=%s: %.3f ms, %.2f us/step" % (B, info.kernel_kind, prof, ms, ms * 1e3 / S), flush=True)
        cyc = eng.stage_cycles().astype(np.float64) / S
        tot = cyc[:, :21].sum(1) + cyc[:, 23]
        print("  cycles/step: cta0 total %.0f  mean %.0f  max %.0f" % (tot[0], tot.mean(), tot.max()))
        print("  gathers (4 per step, thread 0): issue -> first answers %.0f clk each, stale poll rounds %.2f each, whole gather %.0f clk each"
              % (cyc[:, 20].mean() / 4, cyc[:, 21].mean() / 4, cyc[:, 22].mean() / 4))
        print("  sampler of fold 0 per step: poll wait %.0f clk, logits -> x published %.0f clk, loop top (draws) %.0f clk" % (cyc[0, 24], cyc[0, 25], cyc[0, 26]))
        for i, name in enumerate(SLOTS):
            print("  %-44s cta0 %7.0f  mean %7.0f  min %7.0f  max %7.0f" % (name, cyc[0, i], cyc[:, i].mean(), cyc[:, i].min(), cyc[:, i].max()))


if __name__ == "__main__":
    main()
