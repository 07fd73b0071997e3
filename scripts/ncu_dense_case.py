"""One dense-kernel launch for ncu (development aid): B folds x S steps, Philox draws, precision bf16-dense."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from expressive_speech_synthesis_research_b200 import WaveRNN
from oracle import synth
B, S = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda", 0)
m = WaveRNN(**synth.model_kwargs("RAW", "ref")); m.load_state_dict(synth.make_state("RAW", "ref", 0)); m.cuda()
m.precision = "bf16-dense"
eng = m._engine(dev)
mu = torch.rand(B * S, 80, device=dev); au = torch.randn(B * S, 128, device=dev)
starts = np.arange(B, dtype=np.int64) * S
for _ in range(2):
    m._run_folds(eng, dev, mu, au, starts, starts + S, S, None, 1, None, False)
print("kernel ms", eng.info().last_kernel_ms, "us/step", eng.info().last_kernel_ms * 1e3 / S)
