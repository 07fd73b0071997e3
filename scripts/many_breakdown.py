"""Where generate_many's wall time goes on configs[3] (development aid)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from expressive_speech_synthesis_research_b200 import WaveRNN
from bench import GEOMETRY, model_kwargs
sr, hop, _ = GEOMETRY["fatchord"]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = WaveRNN(**model_kwargs("RAW", "fatchord")).to(dev)
m.precision = "bf16-dense"
m.eval()
durs = np.random.default_rng(0).uniform(2, 12, 256)
mels = [torch.rand(1, 80, int(round(d * sr / hop)) + 1, generator=torch.Generator().manual_seed(10 + i)) for i, d in enumerate(durs)]
m.generate(mels[0], True, 11000, 550, True, seed=1)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
with torch.no_grad():
    t0 = T()
    md = [x.to(dev) for x in mels]
    t1 = T()
    conds = [m.conditioning(x) for x in md]
    t2 = T()
    m_all = torch.cat([c[0] for c in conds]).contiguous(); a_all = torch.cat([c[1] for c in conds]).contiguous()
    t3 = T()
    wavs = [torch.empty((x.size(-1) - 1) * hop, dtype=torch.float64, device=dev) for x in md]
    t4 = T()
    host = [w.cpu().numpy() for w in wavs]
    t5 = T()
print("H2D mels %.3f s | conditioning x256 %.3f s | cat %.3f s | alloc %.3f s | D2H float64 %.3f s (%.0f MB)" % (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, sum(h.nbytes for h in host) / 1e6))
with torch.no_grad():
    del conds, m_all, a_all, wavs, host
    t0 = T()
    conds = [m.conditioning(x) for x in md]
    t1 = T()
    del conds
    order = sorted(range(256), key=lambda i: -md[i].size(-1))
    torch.cuda.empty_cache()
    t2 = T()
    conds = [m.conditioning(md[i]) for i in order]
    t3 = T()
    del conds
    t4 = T()
    for i in order[:32]:
        m.conditioning(md[i])
    t5 = T()
    import torch.nn.functional as F
    x = md[order[40]]
    t6 = T()
    for _ in range(20):
        m.conditioning(x)
    t7 = T()
print("second pass same order %.3f s | after empty_cache, longest first %.3f s | 32 longest again %.3f s | one shape x20 %.4f s each" % (t1 - t0, t3 - t2, t5 - t4, (t7 - t6) / 20))
