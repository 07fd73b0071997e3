"""Per-stage cycle breakdown of the WIDE kernel (csrc/wavernn_wide.cuh), thread 0 of every worker CTA (development aid; feeds profiles/)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from expressive_speech_synthesis_research_b200 import WaveRNN  # noqa: E402
from oracle import synth  # noqa: E402

# pass warp 0 (thread 0): slots 0..11, 16, 17; finalize warp 0 (thread 512): 12..14; 21 = gather polls that found stale data
SLOTS = ["P  wait: H1 published (BAR_GO)", "P  S2 gather H1 -> A", "P  S2 pass Wih2x (tensor-memory weights) + store", "P  S2 filler Whh1 + Wfc1x (tensor memory) + issue H2 + store",
         "P  S2 gather finish H2 -> B", "P  S3 pass Wfc1x . h2 + store", "P  cond: TMA wait + part 1", "P  S3 gather Y1 -> A",
         "P  S4 pass Wfc2 + store", "P  cond: part 2", "P  S5 gather Y2 -> A", "P  S5 pass Wfc3 + store",
         "F  wait for x (sampler round trip)", "F  GRU1 + publish H1", "F  waiting for partial sums (all stages)", "-",
         "P  cond part 3 + store, Whh2 . h2 + store", "P  pass-warp barrier + TMA issue", "-", "-", "-", "(count) stale gather polls", "-", "-",
         "F  sum Whh2 . h2 -> gates of rnn2", "F  sum + GRU2 + publish H2", "F  sum Whh1 . h1 -> gates of rnn1, fc1 part", "F  sum + fc1 + publish Y1",
         "F  sum + fc2 + publish Y2", "F  sum + logits sector", "F  sum conditioning projections (2 x 16 float4)", "-"]


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "RAW"
    dev = torch.device("cuda", 0)
    m = WaveRNN(**synth.model_kwargs(mode, "ref"))
    m.load_state_dict(synth.make_state(mode, "ref", 0))
    m.cuda()
    eng = m._engine(dev)
    S = 3000
    print("exchange probe: %.3f us" % eng.measure_exchange(2000))
    for B in [int(a) for a in sys.argv[2:]] or (1, 8, 14, 20, 21):
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
        tot = cyc[:, :12].sum(1) + cyc[:, 16] + cyc[:, 17]
        print("  cycles/step: cta0 total %.0f  mean %.0f  max %.0f" % (tot[0], tot.mean(), tot.max()))
        print("  sampler of fold 0 per step: poll wait %.0f clk, logits -> x published %.0f clk, loop top (draws) %.0f clk" % (cyc[0, 24], cyc[0, 25], cyc[0, 26]))
        for i, name in enumerate(SLOTS):
            print("  %-62s cta1 %7.0f  mean %7.0f  min %7.0f  max %7.0f" % (name, cyc[1, i], cyc[1:, i].mean(), cyc[1:, i].min(), cyc[1:, i].max()))


if __name__ == "__main__":
    main()
