"""A/B timing of two builds of the library on ONE box (development aid): WRNN_LIB=<path> python scripts/ab_time.py [folds ...]
prints the step-loop time of the wide kernel for each fold count, best of three launches."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from expressive_speech_synthesis_research_b200 import WaveRNN
from oracle import synth
dev = torch.device("cuda", 0)
m = WaveRNN(**synth.model_kwargs("RAW", "ref")); m.load_state_dict(synth.make_state("RAW", "ref", 0)); m.cuda()
eng = m._engine(dev)
S = int(os.environ.get("AB_STEPS", "3000"))      # AB_STEPS=40 for a compute-sanitizer pass
out = []
for B in [int(a) for a in sys.argv[1:]] or [1, 14, 20]:
    L = S + 64
    mu = torch.rand(B * L, 80, device=dev); au = torch.randn(B * L, 128, device=dev)
    starts = np.arange(B, dtype=np.int64) * L
    ms = []
    for _ in range(4):
        m._run_folds(eng, dev, mu, au, starts, starts + L, S, None, 1, None, False)
        ms.append(eng.info().last_kernel_ms)
    out.append("B=%d %.2f us/step" % (B, min(ms[1:]) * 1e3 / S))
print(os.environ.get("WRNN_LIB", "in-tree"), "|", ", ".join(out), flush=True)
