"""Host timeline of generate_sharded on the 10-minute utterance of BASELINE.json configs[4] (development aid; run under torchrun):
WRNN_TRACE=1 makes rank 0 print where the call's time goes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["WRNN_TRACE"] = "1"
import torch, torch.distributed as dist
import bench
from expressive_speech_synthesis_research_b200 import WaveRNN, distributed as D
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(0)
m = WaveRNN(**bench.model_kwargs("RAW", "fatchord")).cuda()
m.precision = "auto"
T = 48110
mel = torch.rand(1, 80, T, generator=torch.Generator().manual_seed(0)).pin_memory()
for i in range(4):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    w = D.generate_sharded(m, mel, 11000, 550, True, seed=i, gather_to=0)
    dist.barrier(); torch.cuda.synchronize()
    if rank == 0:
        print("call %d: %.4f s" % (i, time.perf_counter() - t0), flush=True)
dist.destroy_process_group()
