"""Conditioning network with and without cuDNN: time on first sight of 256 utterance lengths, and numerical difference (development aid)."""
import os, sys, time
import numpy as np, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from expressive_speech_synthesis_research_b200 import WaveRNN
from bench import GEOMETRY, model_kwargs
sr, hop, _ = GEOMETRY["fatchord"]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = WaveRNN(**model_kwargs("RAW", "fatchord")).to(dev)
m.eval()
durs = np.random.default_rng(0).uniform(2, 12, 256)
mels = [torch.rand(1, 80, int(round(d * sr / hop)) + 1, generator=torch.Generator().manual_seed(10 + i)).to(dev) for i, d in enumerate(durs)]
def T():
    torch.cuda.synchronize(); return time.perf_counter()
def cond_nocudnn(x):
    with torch.backends.cudnn.flags(enabled=False):
        mm = F.pad(x, (m.pad, m.pad))
        a, b = m.upsample(mm)
        return a[0].contiguous(), b[0].contiguous()
with torch.no_grad():
    m.conditioning(mels[0]); cond_nocudnn(mels[0])
    t0 = T()
    for x in mels[1:129]:
        m.conditioning(x)
    t1 = T()
    for x in mels[129:]:
        cond_nocudnn(x)
    t2 = T()
    for x in mels[129:]:
        cond_nocudnn(x)
    t3 = T()
    a0, b0 = m.conditioning(mels[5]); a1, b1 = cond_nocudnn(mels[5])
    print("cuDNN, 128 new lengths: %.3f s | no cuDNN, 127 new lengths: %.3f s | no cuDNN, same lengths again: %.3f s" % (t1 - t0, t2 - t1, t3 - t2))
    print("max |mel_up diff| %.3e, max |aux diff| %.3e (aux magnitude %.2f)" % ((a0 - a1).abs().max().item(), (b0 - b1).abs().max().item(), b0.abs().max().item()))
