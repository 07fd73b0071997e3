import os, sys, time
sys.path.insert(0, "/root/repo")
os.environ["WRNN_TRACE"] = "1"
import torch, bench, argparse
from expressive_speech_synthesis_research_b200 import WaveRNN
args = argparse.Namespace(geometry="fatchord", mode="RAW", seconds=10.0, utterances=256)
ss = bench.sentence_set(args)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = WaveRNN(**bench.model_kwargs("RAW", "fatchord")).to(dev)
m.precision = "auto"
mels = bench.make_mels(ss["T"], 0, True)
for i in range(3):
    t0 = time.perf_counter()
    outs = m.generate_many(mels, 11000, 550, True, seed=i)
    print("call %d: %.3f s, kernel %.3f s" % (i, time.perf_counter() - t0, m.last_stats["kernel_ms"] / 1e3), flush=True)
