"""CPU tests (gloo, world_size 2 and 3) of the multi-GPU sharding logic in
expressive_speech_synthesis_research_b200/distributed.py: fold ranges, the overlap-edge
all_gather (the only data-path collective) and the per-rank segment assembly, which must
reproduce the single-process waveform BIT FOR BIT.  The CUDA epilogue is
replaced by the oracle (oracle/c_oracle.py) through the `unfold` hook -- this is the checker
standing in for the device call in a test, never a product path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from expressive_speech_synthesis_research_b200 import distributed as D
from oracle import c_oracle


def oracle_segment(rows, first_fold, num_folds, steps, overlap, mu_classes, wave_len, tail_fade, start, length):
    """Oracle stand-in for wrnn_xfade_unfold_segment: place the given rows at their global fold index inside an
    otherwise zero [num_folds, S] array, run the reference-order epilogue and cut the segment."""
    full = np.zeros((num_folds, steps), dtype=np.float32)
    r = rows.cpu().numpy()
    full[first_fold:first_fold + r.shape[0]] = r
    target = steps - 2 * overlap
    hop = tail_fade // 20
    wav = c_oracle.assemble(full, True, target, overlap, mu_classes, wave_len, hop)
    return torch.from_numpy(np.ascontiguousarray(wav[start:start + length]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, case, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, target, overlap, mu, wave_len, hop = case
        S = target + 2 * overlap
        rng = np.random.default_rng(7)
        samples = rng.uniform(-1, 1, (B, S)).astype(np.float32)        # what the step loop would have produced
        lo, hi = D.fold_ranges(B, world)[rank]
        local = torch.from_numpy(samples[lo:hi].copy())
        wav = D.finish_sharded(local, lo, hi, B, target, overlap, mu, wave_len, 20 * hop, gather_to=0, unfold=oracle_segment)
        if rank == 0:
            want = c_oracle.assemble(samples, True, target, overlap, mu, wave_len, hop)
            ret["equal"] = bool(np.array_equal(wav.numpy(), want))
            ret["len"] = int(wav.numel())
        else:
            assert wav is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", [
    (5, 100, 10, 0, 5 * 110 + 10 - 37, 5),        # MOL-style (no mu-law), ragged tail
    (4, 64, 8, 512, 4 * 72 + 8 - 1, 3),           # RAW mu-law decode on the crossfaded samples
    (2, 50, 6, 512, 2 * 56 + 6 - 20, 2),          # fewer folds than ranks at world 3: an empty rank
])
def test_sharded_assembly_is_bit_identical(world, case):
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, case, ret), nprocs=world, join=True)
        assert ret["equal"], "sharded waveform differs from the single-process one"
        assert ret["len"] == case[4]


def test_fold_ranges_and_utterance_plan():
    for B in range(0, 40):
        for G in (1, 2, 3, 4, 8):
            r = D.fold_ranges(B, G)
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
    counts = [11, 3, 7, 20, 5, 5, 14, 2, 9]
    for G in (1, 2, 4, 8):
        plan = D.plan_utterances(counts, G)
        assert sorted(i for p in plan for i in p) == list(range(len(counts)))
        loads = [sum(counts[i] for i in p) for p in plan]
        assert max(loads) <= sum(counts) / G + max(counts)          # LPT bound
    assert D.plan_utterances(counts, 2) == D.plan_utterances(counts, 2)   # deterministic


def test_segment_bounds_partition_the_waveform():
    B, target, overlap = 7, 100, 10
    wave_len = B * 110 + 10 - 13
    for G in (1, 2, 3, 8):
        pos = 0
        for lo, hi in D.fold_ranges(B, G):
            a, b = D.segment_bounds(lo, hi, B, target, overlap, wave_len)
            if hi > lo:
                assert a == pos
                pos = b
        assert pos == wave_len
