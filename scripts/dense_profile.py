"""Per-step cycle budget of the dense kernel from its in-kernel counters (development tool)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from expressive_speech_synthesis_research_b200 import WaveRNN, _lib  # noqa: E402
from oracle import synth  # noqa: E402
from scripts.dense_bringup import run_folds  # noqa: E402

NAMES = {0: "mma wait H1", 1: "mma wait H2", 2: "mma wait Y1", 3: "mma wait Y2", 4: "mma wait COND", 5: "mma wait ring full", 6: "mma issue+commit",
         7: "mma total", 8: "epi wait G1", 9: "epi wait G2", 10: "epi wait F1", 11: "epi wait F2", 12: "epi wait F3", 13: "epi wait x", 14: "epi wait logits",
         15: "E1", 16: "E2", 17: "E3", 18: "E4", 19: "E5", 20: "sampling", 21: "conditioning", 22: "epi total", 24: "producer0 wait empty", 25: "producer0 total"}


def main():
    global out
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    sd = synth.make_state("RAW", "ref", 0)
    m = WaveRNN(**synth.model_kwargs("RAW", "ref"))
    m.load_state_dict(sd)
    m = m.cuda()
    m.precision = "bf16-dense"
    rng = np.random.default_rng(7)
    mels = rng.uniform(0, 1, (B, S, 80)).astype(np.float32)
    aux = rng.normal(0, 1, (B, S, 128)).astype(np.float32)
    run_folds(m, mels, aux, None, seed=3)
    eng = m._engine(torch.device("cuda", 0))
    _lib.check(eng.lib.wrnn_set_profiling(eng.handle, 1))
    r = run_folds(m, mels, aux, None, seed=3)
    n = eng.info().ctas * 192
    out = np.zeros(n, dtype=np.int64)
    _lib.check(eng.lib.wrnn_get_stage_cycles(eng.handle, out.ctypes.data, n))
    out = out.reshape(-1, 192)
    used = min(((B + 31) // 32) * 8, out.shape[0])
    print("co-resident clusters:", eng.info().ctas // 8, "launches", eng.info().launches)
    print("B=%d S=%d kernel %.2f ms = %.2f us/step; cycles per step, CTA 0 | mean over %d CTAs" % (B, S, r["ms"], 1e3 * r["ms"] / S, used))
    for i in sorted(NAMES):
        print("  %-22s %9.0f | %9.0f" % (NAMES[i], out[0, i] / S, out[:used, i].mean() / S))


    print("bundle trace of step 10, CTA 0 (cycles since the first bundle's wait began): wait-done, slot-full, issued")
    tr = out[0, 32:32 + 4 * 40].reshape(40, 4)
    for b in range(40):
        if tr[b, 2]:
            print("  bundle %2d: %6d %6d %6d" % (b, tr[b, 0], tr[b, 1], tr[b, 2]))


if __name__ == "__main__":
    main()
