"""Host timeline of generate_many (development aid): WRNN_TRACE=1 prints the pipeline marks of every call.
usage: python scripts/many_trace.py [utterances of the configs[3] set to take, default all 256; 32 = one rank's share at 8 GPUs]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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
take = int(sys.argv[1]) if len(sys.argv) > 1 else len(mels)
mels = mels[::len(mels) // take][:take]
for i in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    outs = m.generate_many(mels, 11000, 550, True, seed=i)
    print("call %d: %d utterances, %d folds: %.3f s, kernel %.3f s" % (i, len(mels), m.last_stats.get("folds", -1), time.perf_counter() - t0, m.last_stats["kernel_ms"] / 1e3), flush=True)
