"""Development aid: step time of the wide kernel with random conditioning (what scripts/ab_time.py measures) against the same kernel
inside generate() on BASELINE.json configs[1] and its ref-geometry twin.  Found the 1.6 us/step that the per-step zeroing of a padding
fold's conditioning row cost (profiles/r02_summary.md section 4)."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from expressive_speech_synthesis_research_b200 import WaveRNN
from oracle import synth
dev = torch.device("cuda", 0)
for geo in ("ref", "fatchord"):
    m = WaveRNN(**synth.model_kwargs("RAW", geo)); m.load_state_dict(synth.make_state("RAW", geo, 0)); m.cuda()
    eng = m._engine(dev)
    for S in (3000, 12100):
        B = 20; L = S + 64
        mu = torch.rand(B * L, 80, device=dev); au = torch.randn(B * L, 128, device=dev)
        starts = np.arange(B, dtype=np.int64) * L
        ms = []
        for _ in range(3):
            m._run_folds(eng, dev, mu, au, starts, starts + L, S, None, 1, None, False)
            ms.append(eng.info().last_kernel_ms)
        print(geo, "random cond S=%d: %.2f us/step" % (S, min(ms[1:]) * 1e3 / S), flush=True)
    mel = synth.make_mel(803, seed=0)
    for _ in range(3):
        m.generate(mel, True, 11000, 550, True, seed=1)
        print(geo, "generate(): kernel %.2f us/step" % (m.last_stats["kernel_ms"] * 1e3 / 12100), {k: v for k, v in m.last_stats.items() if k in ("folds", "kernel_kind")}, flush=True)
