"""Multi-GPU sharding of the generation path: one process per GPU (torch.distributed, NCCL over
NVLink on the GPU box, gloo in the CPU tests).

Every fold is independent until the crossfade (zero initial state, fatchord_version.py:173-175),
so the path shards without any collective inside the step loop:

* sentence sets -- whole utterances go to ranks by longest-processing-time on their fold counts
  (`plan_utterances`); no communication at all.
* one long utterance -- contiguous fold ranges per rank (`fold_ranges`); after generation each
  rank needs only the LAST `overlap` samples of its left neighbour's last fold to finish the
  crossfade of its own span, so the single exchange is an all_gather of one [overlap] fp32 edge
  per rank (2.2 KB at overlap=550).  Each rank then runs the segment epilogue
  (wrnn_xfade_unfold_segment) on its span; concatenated spans are bit-identical to the
  single-GPU waveform because every output sample is still (0 + a*fade_out) + b*fade_in in fp64.
"""
import ctypes
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .wavio import decode_mu_law_host, parallel_copy


def plan_utterances(fold_counts, world_size):
    """Longest-processing-time assignment of whole utterances to ranks.
    Returns a list (per rank) of utterance indices; deterministic for equal counts."""
    order = sorted(range(len(fold_counts)), key=lambda i: (-fold_counts[i], i))
    load = [0] * world_size
    plan = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        plan[r].append(i)
        load[r] += fold_counts[i]
    return plan


def fold_ranges(num_folds, world_size):
    """Contiguous fold ranges [lo, hi) per rank: rank g gets floor(g*B/G) .. floor((g+1)*B/G)."""
    return [(g * num_folds // world_size, (g + 1) * num_folds // world_size) for g in range(world_size)]


def exchange_edges(tail, group=None):
    """all_gather of each rank's trailing-overlap edge ([overlap] fp32, zeros if the rank has no
    folds).  Returns a [world, overlap] tensor on tail's device.  This is the only collective of
    the sharded path."""
    world = dist.get_world_size(group)
    out = [torch.empty_like(tail) for _ in range(world)]
    dist.all_gather(out, tail.contiguous(), group=group)
    return torch.stack(out)


def segment_rows(local_samples, left_tail, lo):
    """Rows handed to the segment epilogue: the left neighbour's overlap samples as one extra
    (otherwise zero) row in front of this rank's folds.  Returns (rows [n(+1), S], first_fold)."""
    if lo == 0:
        return local_samples, 0
    ghost = torch.zeros_like(local_samples[:1])
    ghost[0, -left_tail.numel():] = left_tail
    return torch.cat([ghost, local_samples]), lo - 1


def segment_bounds(lo, hi, num_folds, target, overlap, wave_len):
    """Global sample range [start, stop) owned by the rank holding folds [lo, hi)."""
    hop = target + overlap
    start = lo * hop
    stop = hi * hop + (overlap if hi == num_folds else 0)
    return min(start, wave_len), min(stop, wave_len)


def assemble_segment(rows, first_fold, num_folds, steps, overlap, mu_law_classes, wave_len, tail_fade, seg, unfold=None):
    """Run the segment epilogue.  `unfold` lets the CPU tests substitute the oracle for the CUDA call."""
    start, stop = seg
    if stop <= start:
        return torch.empty(0, dtype=torch.float64, device=rows.device)
    if unfold is not None:
        return unfold(rows, first_fold, num_folds, steps, overlap, mu_law_classes, wave_len, tail_fade, start, stop - start)
    out = torch.empty(stop - start, dtype=torch.float64, device=rows.device)
    stream = torch.cuda.current_stream(rows.device).cuda_stream
    _lib.check(_lib.lib().wrnn_xfade_unfold_segment(rows.data_ptr(), rows.shape[0], steps, overlap, mu_law_classes, wave_len,
                                                    tail_fade, first_fold, num_folds, start, stop - start, out.data_ptr(),
                                                    ctypes.c_void_p(stream)))
    return out


def finish_sharded(local_samples, lo, hi, num_folds, target, overlap, mu_law_classes, wave_len, tail_fade,
                   group=None, gather_to=0, unfold=None):
    """Edge exchange + local crossfade/unfold + (optional) gather of the waveform spans.

    local_samples: [hi-lo, S] fp32 samples of this rank's folds (may be empty).
    Returns the full float64 waveform on rank `gather_to` (None elsewhere), or this rank's span
    when gather_to is None."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    S = target + 2 * overlap
    dev = local_samples.device
    tail = local_samples[-1, S - overlap:].clone() if hi > lo else torch.zeros(overlap, dtype=torch.float32, device=dev)
    edges = exchange_edges(tail, group)                       # the ONLY collective on the data path
    ranges = fold_ranges(num_folds, world)
    left = next((r for r in range(rank - 1, -1, -1) if ranges[r][1] > ranges[r][0]), None)
    seg = segment_bounds(lo, hi, num_folds, target, overlap, wave_len)
    if hi > lo:
        rows, first = segment_rows(local_samples, edges[left] if left is not None else None, lo)
        span = assemble_segment(rows, first, num_folds, S, overlap, mu_law_classes, wave_len, tail_fade, seg, unfold)
    else:
        span = torch.empty(0, dtype=torch.float64, device=dev)
    if gather_to is None:
        return span
    sizes = [max(0, segment_bounds(a, b, num_folds, target, overlap, wave_len)[1]
                 - segment_bounds(a, b, num_folds, target, overlap, wave_len)[0]) if b > a else 0 for a, b in ranges]
    # result assembly (not part of the step path): every rank's span travels ONCE, to the root only -- the round-1 all_gather
    # moved world x waveform bytes to every rank (106 MB x 8 for a 10-minute utterance)
    width = max(sizes + [1])
    padded = torch.zeros(width, dtype=torch.float64, device=dev)
    padded[:span.numel()] = span
    parts = [torch.empty_like(padded) for _ in range(world)] if rank == gather_to else None
    dist.gather(padded, parts, dst=gather_to if group is None else dist.get_global_rank(group, gather_to), group=group)
    if rank != gather_to:
        return None
    return torch.cat([p[:n] for p, n in zip(parts, sizes)])


def generate_sharded(model, mels, target, overlap, mu_law, uniforms=None, seed=0, group=None, gather_to=0):
    """One utterance, folds sharded over the ranks of `group`; equals
    model.generate(mels, True, target, overlap, mu_law, uniforms=...) bit for bit.
    `uniforms` is the full [S, B(,11)] tensor (each rank consumes its columns)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mu_law = mu_law if model.mode == 'RAW' else False
    model.eval()
    trace = [] if os.environ.get("WRNN_TRACE") else None      # development: host timeline of the call (device synchronised at every mark)

    def mark(what):
        if trace is not None:
            torch.cuda.synchronize()
            trace.append((what, time.perf_counter()))
    try:
        with torch.no_grad():
            device = model._device()
            with torch.cuda.device(device):
                mark("start")
                eng = model._engine(device)
                nfolds_total, _ = _lib.fold_index(mels.size(-1) * model.hop_length, target, overlap)
                eng = model._pick_engine(eng, device, -(-nfolds_total // world))      # regime by the folds of ONE rank
                mels = mels.to(device=device, dtype=torch.float32)
                if model.upsample.resnet.conv_in.weight.device != device:
                    model.to(device)
                wave_len = (mels.size(-1) - 1) * model.hop_length
                frames = model._frames_mode(eng)              # dense kernel: conditioning expanded in the kernel, as generate() does
                if frames:
                    L = mels.size(-1) * model.hop_length
                    if model.melresnet_native():
                        # the conditioning network is sharded with the folds.  A rank's folds read the samples
                        # [lo (t + o), hi (t + o) + o), i.e. aux frames [p0 / hop, (p1 - 1) / hop]; the MelResNet kernel computes exactly
                        # those (its k = 5 input window is the halo), and a frame's result does not depend on what is computed with it
                        # (csrc/wavernn_cond.cuh), so the shards are bit-identical to the unsplit run.  The zero-padded mel frames
                        # themselves (a transpose) are kept whole: the kernel's box-filter taps index them directly.
                        import torch.nn.functional as F
                        T = int(mels.size(-1))
                        mel_fr = F.pad(mels, (model.pad, model.pad))[0].t().contiguous()
                        aux_fr = torch.zeros(T, 4 * model.aux_dims, dtype=torch.float32, device=device)
                        Bt, _ = _lib.fold_index(L, target, overlap)
                        flo, fhi = fold_ranges(Bt, world)[rank]
                        if fhi > flo:
                            p0, p1 = flo * (target + overlap), min(L, fhi * (target + overlap) + overlap)
                            a0, a1 = p0 // model.hop_length, min(T - 1, (p1 - 1) // model.hop_length)
                            model._cond(device).frames(mel_fr, np.array([[a0, a1 - a0 + 1, a0]], dtype=np.int32), aux_fr)
                    else:
                        mel_fr, aux_fr = model.conditioning_frames(mels)
                else:
                    m_up, aux = model.conditioning(mels)
                    L = m_up.size(0)
                mark("conditioning")
                B, _ = _lib.fold_index(L, target, overlap)
                S = target + 2 * overlap
                lo, hi = fold_ranges(B, world)[rank]
                # the fp32 kernels differ in summation order: every rank must run the one the unsplit call would run; that is the
                # library's default for every fold count (the wide kernel where the model allows it), so nothing is pinned here
                eng.set_kernel(-1)
                if hi > lo:
                    starts = np.arange(lo, hi, dtype=np.int64) * (target + overlap)
                    u = None if uniforms is None else torch.as_tensor(uniforms)[:, lo:hi]
                    if frames:
                        geo = np.stack([starts, np.full(hi - lo, L), np.zeros(hi - lo), np.zeros(hi - lo)], axis=1)
                        res = model._run_folds_frames(eng, device, mel_fr, aux_fr, geo, S, u, seed + rank, None, False, wait=False)
                    else:
                        limits = np.full(hi - lo, L, dtype=np.int64)
                        res = model._run_folds(eng, device, m_up, aux, starts, limits, S, u, seed + rank, None, False, wait=False)
                    local = res["samples"]
                else:
                    local = torch.empty(0, S, dtype=torch.float32, device=device)
                # The step loop is running (the calls above only enqueue).  The root uses the wait to get the caller's array ready: a
                # fresh 106 MB array (10-minute utterance) costs ~20 ms of page faults when it is first written.
                out = None
                if gather_to is not None and rank == gather_to:
                    out = np.empty(wave_len, dtype=np.float64)
                    out[::512] = 0.0                                              # touch every 4 KB page
                    model._pinned("sharded", wave_len)
                if hi > lo:
                    eng.synchronize()                                             # watchdog status of the step loop
                mark("step loop (+ output array prepared)")
                # same decode policy as WaveRNN.generate: bit-exact numpy mu-law + tail fade on the host below 2 M samples
                host_mu = bool(mu_law) and gather_to is not None and (
                    model.mu_law_decode == "host" or (model.mu_law_decode == "auto" and wave_len < 2_000_000))
                wav = finish_sharded(local, lo, hi, B, target, overlap, model.n_classes if (mu_law and not host_mu) else 0, wave_len,
                                     0 if host_mu else 20 * model.hop_length, group=group, gather_to=gather_to)
                mark("edges, epilogue, gather to root")
                if wav is None:
                    return None
                host = model._pinned("sharded", wav.numel())            # pinned landing buffer: 106 MB of a 10-minute utterance at PCIe speed
                host.copy_(wav, non_blocking=True)
                torch.cuda.current_stream(device).synchronize()
                if out is None or out.size != wav.numel():
                    out = np.empty(wav.numel(), dtype=np.float64)
                mark("D2H")
                parallel_copy(out, host.numpy())                        # a few threads, into the pre-faulted array
                wav = out
                mark("copy to the caller's array")
                if trace is not None:
                    print("generate_sharded timeline (ms): " + ", ".join("%s +%.1f" % (w, (t - trace[i][1]) * 1e3) for i, (w, t) in enumerate(trace[1:])), flush=True)
                if host_mu:
                    mu = model.n_classes - 1
                    wav = decode_mu_law_host(wav, mu)             # decode_mu_law, dsp.py:100-105
                    wav[-20 * model.hop_length:] *= np.linspace(1, 0, 20 * model.hop_length)   # fatchord_version.py:235-237
                return wav
    finally:
        for e in model._engines.values():
            e.set_kernel(-1)                      # back to the per-call choice
        model.train()
