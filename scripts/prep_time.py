"""Development aid: host and device time of WaveRNN.conditioning_frames_many (one copy + one launch of the MelResNet kernel,
csrc/wavernn_cond.cuh) for one rank's share of configs[3] on 8 GPUs (32 utterances), plus a cProfile of the host side."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, argparse, numpy as np
from expressive_speech_synthesis_research_b200 import WaveRNN
args = argparse.Namespace(geometry="fatchord", mode="RAW", seconds=10.0, utterances=256)
ss = bench.sentence_set(args)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = WaveRNN(**bench.model_kwargs("RAW", "fatchord")).to(dev)
mels = bench.make_mels(ss["T"], 0, True)[::8][:32]
print("pinned inputs:", mels[0].is_pinned(), mels[0].shape, mels[0].dtype)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    mf, af, mr, ar = m.conditioning_frames_many(mels, dev)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("conditioning_frames_many: host %.2f ms, + device %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for it in range(5):
    m.conditioning_frames_many(mels, dev)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
