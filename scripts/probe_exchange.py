import sys, torch
sys.path.insert(0, '/root/repo')
from expressive_speech_synthesis_research_b200 import WaveRNN
from oracle import synth
m = WaveRNN(**synth.model_kwargs("RAW", "ref")); m.load_state_dict(synth.make_state("RAW", "ref", 0)); m.cuda()
eng = m._engine(torch.device("cuda", 0))
for it in (2000, 4000): print("probe usec/exchange", eng.measure_exchange(it))
