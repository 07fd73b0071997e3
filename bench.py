#!/usr/bin/env python
"""bench.py -- WaveRNN batched generation throughput (BASELINE.json metric) on N B200s.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (torch port of it), rank 0 only

N = 1: a "step" is one `generate(mel, batched=True, target=11000, overlap=550, mu_law=True)` of a 10 s utterance
(BASELINE.json configs[1]: RAW 9-bit, random-init weights, 22.05 kHz / hop 275 -> 20 folds of 12100 steps, fp32).
N > 1 (torchrun, one rank per GPU): a step is BASELINE.json configs[3], the whole 256-utterance sentence set (2-12 s each),
whole utterances dealt to the ranks by longest-processing-time on their fold counts (no data-path collective), each rank
running generate_many on its share (precision 'auto' -> the dense tcgen05 kernel); STRONG scaling: the set is fixed,
value = samples of the set / max-over-ranks time.  The N > 1 line also carries `long_form`: configs[4], one 10-minute
utterance with its folds sharded over the ranks and the overlap edges all-gathered (distributed.generate_sharded); the N = 1
line carries `sentence_set`, the same configs[3] on one GPU, which is what the multi-GPU values scale against.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GEOMETRY = {   # name -> (sample_rate, hop, upsample factors)
    "fatchord": (22050, 275, (5, 5, 11)),     # the geometry BASELINE.json's configs name
    "ref": (16000, 200, (5, 5, 8)),           # this repo's hparams.py
}
TARGET, OVERLAP = 11000, 550
FLOP_PER_FOLD_STEP = {"RAW": 8143872, "MOL": 7650304}
HBM_BYTES_PER_FOLD_STEP = {"RAW": 840, "MOL": 880}
SM_CLOCK_MHZ = 1965.0
# shared-memory operand wavefronts of one step of the wide kernel per CTA (the pipe delivers 32 lane-words
# per clock): gate passes 16 warps x 16 k x (6 | 7 | 9) (Wih2x and Whh1 come from tensor memory, Whh2 from shared memory),
# fc passes 3 x 16 x 80, conditioning 16 x 11 x 10, partial sums ~800
WIDE_SMEM_WAVEFRONTS = 16 * 16 * (6 + 7 + 9) + 3 * 16 * 80 + 16 * 11 * 10 + 800
SMALL_HOP_US = 0.46      # store -> successful poll of one value between two CTAs: 910 cycles (profiles/r01_l2_latency_raw.log)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--geometry", default="fatchord", choices=list(GEOMETRY))
    ap.add_argument("--mode", default="RAW", choices=["RAW", "MOL"])
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16", "bf16-dense"],
                    help="N = 1 only.  fp32 (headline) | bf16 resident weights of the grouped FFMA kernel (BASELINE.json configs[2]) | "
                         "bf16-dense: the tcgen05 / tensor-memory kernel for large fold batches")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-regime (pooled folds, tcgen05) measurement")
    ap.add_argument("--dense-folds", type=int, default=480)
    ap.add_argument("--dense-steps", type=int, default=2000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sentence-set", action="store_true", help="N = 1: skip the configs[3] extra key")
    ap.add_argument("--no-long-form", action="store_true", help="N > 1: skip the configs[4] extra key")
    ap.add_argument("--utterances", type=int, default=256, help="size of the configs[3] sentence set")
    return ap.parse_args()


def model_kwargs(mode, geometry):
    sr, hop, ups = GEOMETRY[geometry]
    return dict(rnn_dims=512, fc_dims=512, bits=9, pad=2, upsample_factors=ups, feat_dims=80, compute_dims=128,
                res_out_dims=128, res_blocks=10, hop_length=hop, sample_rate=sr, mode=mode)


def fold_count(L):
    n = (L - OVERLAP) // (TARGET + OVERLAP)
    if L - (n * (TARGET + OVERLAP) + OVERLAP) != 0:
        n += 1
    return n


def workload(args):
    sr, hop, _ = GEOMETRY[args.geometry]
    T = int(round(args.seconds * sr / hop)) + 1
    return dict(T=T, hop=hop, sr=sr, L=T * hop, folds=fold_count(T * hop), S=TARGET + 2 * OVERLAP, wave_len=(T - 1) * hop)


def sentence_set(args):
    """BASELINE.json configs[3]: durations numpy default_rng(0).uniform(2, 12, 256) s."""
    import numpy as np
    sr, hop, _ = GEOMETRY[args.geometry]
    dur = np.random.default_rng(0).uniform(2, 12, args.utterances)
    T = [int(round(d * sr / hop)) + 1 for d in dur]
    return dict(T=T, hop=hop, sr=sr, folds=[fold_count(t * hop) for t in T], wave_len=[(t - 1) * hop for t in T], S=TARGET + 2 * OVERLAP)


def config_dict(args, n_gpus):
    if n_gpus == 1:
        wl = workload(args)
        return {"workload": "configs[1]: WaveRNN %s 9-bit batched generate, target=11000 overlap=550, one %.0f s synthetic utterance"
                            % (args.mode, args.seconds),
                "geometry": "%s (%d Hz, hop %d)" % (args.geometry, wl["sr"], wl["hop"]),
                "mel_frames": wl["T"], "folds": wl["folds"], "steps_per_fold": wl["S"], "wave_len": wl["wave_len"],
                "parallelism": "one GPU", "weights": "random-init (torch.manual_seed(0))",
                "l2": "per-step inputs exceed L2: %.0f MB of upsampled conditioning rewritten every step" % (wl["L"] * 208 * 4 / 1e6)}
    ss = sentence_set(args)
    return {"workload": "configs[3]: sentence-set vocoding, %d synthetic utterances of 2-12 s, WaveRNN %s 9-bit batched generate "
                        "target=11000 overlap=550, whole utterances dealt to the GPUs by LPT on fold counts" % (args.utterances, args.mode),
            "geometry": "%s (%d Hz, hop %d)" % (args.geometry, ss["sr"], ss["hop"]),
            "utterances": args.utterances, "folds": int(sum(ss["folds"])), "steps_per_fold": ss["S"], "wave_len": int(sum(ss["wave_len"])),
            "parallelism": "utterance-sharded x%d, no data-path collective" % n_gpus, "weights": "random-init (torch.manual_seed(0))",
            "l2": "per-step working set exceeds L2: %.0f MB of frame-rate conditioning and samples per step" % (sum(ss["wave_len"]) * 12 / 1e6)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# the reference's CPU implementation (torch port of fatchord_version.py::generate)
# ------------------------------------------------------------------------------------------------
def cpu_generate(args, state, T, threads, steps=None, seed=0):
    """One generate() of a T-frame utterance by the torch port of the reference (oracle/torch_port.py: the ATen ops of
    fatchord_version.py:150-243 in the same order) on `threads` host threads.  steps=None: ALL S steps of every fold (nothing
    extrapolated); steps=k: only the first k steps of the step loop (used for the 1-thread figure).  Returns seconds."""
    import torch
    from oracle import torch_port
    _, hop, ups = GEOMETRY[args.geometry]
    torch.set_num_threads(threads)
    sd = {k: v.detach().cpu() for k, v in state.items()}
    mel = torch.rand(1, 80, T, generator=torch.Generator().manual_seed(seed))
    t0 = time.perf_counter()
    with torch.no_grad():
        if steps is None:
            torch_port.generate(sd, mel, True, TARGET, OVERLAP, True, mode=args.mode, upsample_factors=ups, pad=2, hop_length=hop)
        else:
            m, aux = torch_port.conditioning(sd, mel, ups, 2)
            m = torch_port.fold_with_overlap(m, TARGET, OVERLAP)[:, :steps].contiguous()
            aux = torch_port.fold_with_overlap(aux, TARGET, OVERLAP)[:, :steps].contiguous()
            torch_port.step_loop(sd, args.mode, m, aux)
    return time.perf_counter() - t0


def run_reference(args):
    """Reference arm: the reference's CPU generate() on this box's host cores, rank 0 only.  N = 1: every step is one FULL
    generate() of the configs[1] utterance (all 12 100 steps of all 20 folds).  N > 1: the arm's workload is configs[3]
    (41 M samples: half an hour of CPU time), so every step vocodes a bounded sample of it, the set's first utterance, in full."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from expressive_speech_synthesis_research_b200 import WaveRNN
    torch.manual_seed(0)
    state = WaveRNN(**model_kwargs(args.mode, args.geometry)).state_dict()
    cores = os.cpu_count() or 1
    if args.gpus == 1:
        wl = workload(args)
        T, wave_len, sr = wl["T"], wl["wave_len"], wl["sr"]
        sample = "the whole workload: one full generate() per step (%d folds x %d steps, nothing extrapolated), torch %s, %d threads" % (
            wl["folds"], wl["S"], torch.__version__, cores)
    else:
        ss = sentence_set(args)
        T, wave_len, sr = ss["T"][0], ss["wave_len"][0], ss["sr"]
        sample = "bounded sample of configs[3]: utterance 0 of the set (%d frames, %d folds) vocoded in full every step, torch %s, %d threads" % (
            T, ss["folds"][0], torch.__version__, cores)
    times = [cpu_generate(args, state, T, cores, seed=it) for it in range(args.warmup + args.steps)][args.warmup:]
    t = sum(times) / len(times)
    value = wave_len / t
    one = None
    try:        # the 1-thread figure of BASELINE.md section 3, on a tenth of the steps (a rate, labelled as a sample)
        k = max(1, (TARGET + 2 * OVERLAP) // 10)
        t1 = cpu_generate(args, state, T, 1, steps=k)
        one = {"value": wave_len * k / (TARGET + 2 * OVERLAP) / t1, "unit": "samples/s", "cores": 1, "kind": "port",
               "sample": "first %d of %d steps of the step loop of every fold + conditioning, 1 thread" % (k, TARGET + 2 * OVERLAP)}
    except Exception as e:  # pragma: no cover
        one = {"value": None, "sample": "failed: %r" % (e,)}
    line = {"impl": "reference", "metric": "generated_samples_per_sec", "value": value, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "rtf": t / (wave_len / sr), "config": config_dict(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "cpu_baseline_1thread": one,
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def dense_regime(args, dev, state, pk):
    """BASELINE.json configs[3]/[4] regime: hundreds of pooled folds per GPU (sentence sets, long-form).  Times the
    step loop alone (wrnn_generate_folds, conditioning resident in HBM, in-kernel Philox draws) for the tcgen05 kernel
    (precision bf16-dense) and for the fp32 wide FFMA kernel (serial launches of 21 folds) on the same folds."""
    import numpy as np
    import torch
    from expressive_speech_synthesis_research_b200 import WaveRNN
    B, S = args.dense_folds, args.dense_steps
    g = torch.Generator(device="cpu").manual_seed(5)
    mels = torch.rand(B * S, 80, generator=g).to(dev)
    aux = torch.randn(B * S, 128, generator=g).to(dev)
    starts = np.arange(B, dtype=np.int64) * S
    out = {"workload": "configs[3]/[4] regime: %d pooled folds x %d steps, step loop only, conditioning resident in HBM (%.0f MB > L2)"
                       % (B, S, B * S * 208 * 4 / 1e6)}
    for prec in ("bf16-dense", "fp32"):
        m = WaveRNN(**model_kwargs("RAW", args.geometry)).to(dev)
        m.load_state_dict(state)
        m.precision = prec
        eng = m._engine(dev)
        nb = B if prec == "bf16-dense" else min(B, 63)           # fp32: three launches of the wide kernel (21 folds each)
        ms = []
        for it in range(3):
            m._run_folds(eng, dev, mels, aux, starts[:nb], starts[:nb] + S, S, None, 11 + it, None, False)
            torch.cuda.synchronize()
            ms.append(eng.info().last_kernel_ms)
        t = min(ms[1:]) * 1e-3
        out[prec] = {"folds": nb, "us_per_step": t / S * 1e6, "fold_steps_per_us": nb * S / t / 1e6,
                     "samples_per_sec_raw": nb * S / t}
        del m, eng
    d = out["bf16-dense"]
    tf = d["fold_steps_per_us"] * 1e6 * FLOP_PER_FOLD_STEP["RAW"] / 1e12
    out["speedup_vs_fp32_kernel"] = d["fold_steps_per_us"] / out["fp32"]["fold_steps_per_us"]
    out["roofline"] = {"bound": "tensor", "kernel": "wavernn_dense_kernel", "achieved": tf, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                       "frac": tf / pk["bf16_tflops"], "traffic": None,
                       "traffic_ncu_capture": "dram read 127.8 MB + write 4.1 MB for 480 folds x 300 steps (profiles/r01_dense.md section 3; "
                                              "algorithmic 832 B per fold-step = 119.8 MB); not re-captured for this run",
                       "note": "algorithmic FLOPs (8.14 MFLOP per fold-step); every tcgen05.mma is M=128 x N=32 x K=16 and is paced by "
                               "its shared-memory operand reads (40 clk measured, scripts/umma_rate.cu), the step by the shared-memory "
                               "port (weights cross it twice: TMA fill + MMA read) and the cluster exchange (DESIGN.md 9)"}
    return out


def make_mels(T_list, seed0, pinned):
    import torch
    out = []
    for i, T in enumerate(T_list):
        m = torch.rand(1, 80, T, generator=torch.Generator().manual_seed(seed0 + i))
        out.append(m.pin_memory() if pinned else m)
    return out


def time_sentence_set(model, mels_host, dev, steps, warmup, barrier):
    """generate_many over `mels_host` (pinned host tensors): returns (seconds per step with the mels resident in HBM, seconds per
    step end to end from host buffers, launches per step)."""
    import torch
    mels_dev = [m.to(dev) for m in mels_host]
    eng_launches = lambda: (sum(int(e.info().launches) for e in model._engines.values())
                            + sum(c.launches() for c in model._conds.values()))      # step-loop kernels + the frame-rate conditioning kernel
    for i in range(warmup):
        model.generate_many(mels_dev, TARGET, OVERLAP, True, seed=1000 + i)
        model.generate_many(mels_host, TARGET, OVERLAP, True, seed=1000 + i)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng_launches()
    barrier()
    ev0.record()
    for i in range(steps):
        model.generate_many(mels_dev, TARGET, OVERLAP, True, seed=2000 + i)
    ev1.record()
    barrier()
    t_dev = ev0.elapsed_time(ev1) / 1e3 / steps
    launches = (eng_launches() - l0) / steps
    barrier()
    w0 = time.perf_counter()
    ev0.record()
    for i in range(steps):
        outs = model.generate_many(mels_host, TARGET, OVERLAP, True, seed=3000 + i)
    ev1.record()
    barrier()
    t_e2e = max(ev0.elapsed_time(ev1) / 1e3, time.perf_counter() - w0) / steps
    del outs
    return t_dev, t_e2e, launches


# ------------------------------------------------------------------------------------------------
# this repo's CUDA path
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    # stdout carries exactly one JSON line: anything a library prints there meanwhile (NCCL announces its version on stdout when
    # the first communicator is created) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    line = run_single(args, dev, local, barrier) if world == 1 else run_multi(args, dev, local, rank, world, barrier)
    if rank == 0:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_single(args, dev, local, barrier):
    import torch
    from expressive_speech_synthesis_research_b200 import WaveRNN
    wl = workload(args)
    torch.manual_seed(0)
    model = WaveRNN(**model_kwargs(args.mode, args.geometry)).to(dev)
    model.precision = args.precision
    mel_host = torch.rand(1, 80, wl["T"], generator=torch.Generator().manual_seed(0)).pin_memory()
    mel_dev = mel_host.to(dev)
    eng = model._engine(dev)

    def device_step(seed):
        # inputs already resident in HBM; conditioning net + step loop + epilogue, no host copies
        model.eval()                      # generate() leaves the module in train() like the reference (:241)
        with torch.no_grad():
            return model._generate_on_device(eng, dev, mel_dev, True, TARGET, OVERLAP, True, None, seed, None, False)[0]

    def e2e_step(seed):
        return model.generate(mel_host, True, TARGET, OVERLAP, True, seed=seed)     # host mel in, numpy out

    model.eval()
    for i in range(args.warmup):
        device_step(i)
        e2e_step(i)
    info0 = eng.info()
    cond0 = sum(c.launches() for c in model._conds.values())       # the frame-rate conditioning kernel (csrc/wavernn_cond.cuh)
    sampler = ClockSampler(local)
    sampler.start()
    # ---- kernel-resident leg -------------------------------------------------------------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    barrier()
    ev0.record()
    for i in range(args.steps):
        device_step(100 + i)
        kernel_ms.append(eng.info().last_kernel_ms)
    ev1.record()
    barrier()
    t_dev = ev0.elapsed_time(ev1) / 1e3
    info1 = eng.info()
    cond1 = sum(c.launches() for c in model._conds.values())
    # ---- end-to-end leg (host buffers, H2D + D2H inside the timed region) ----------------------
    barrier()
    ev0.record()
    w0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(200 + i)
    ev1.record()
    barrier()
    t_e2e = max(ev0.elapsed_time(ev1) / 1e3, time.perf_counter() - w0)
    clocks = sampler.stop()
    model.train()

    total_samples = wl["wave_len"] * args.steps
    pk, pk_src = peaks()
    fold_steps = wl["folds"] * wl["S"]
    k_ms = sum(kernel_ms) / len(kernel_ms)
    us_per_step = k_ms * 1e3 / wl["S"]
    kind = int(info1.kernel_kind)                 # 0 grouped FFMA (round 1), 1 wide FFMA, 2 dense tcgen05
    kernel = {0: "wavernn_persistent_kernel", 1: "wavernn_wide_kernel", 2: "wavernn_dense_kernel"}[kind]
    try:
        t_exch = eng.measure_exchange(4000)
    except Exception:  # pragma: no cover
        t_exch = None
    clk = (clocks["sm_mhz"] or SM_CLOCK_MHZ) * 1e6
    fp32_peak = 148 * 128 * 2 * clk / 1e12
    fma_floor = wl["folds"] * FLOP_PER_FOLD_STEP[args.mode] / (fp32_peak * 1e12) * 1e6
    hbm_gbs = fold_steps * HBM_BYTES_PER_FOLD_STEP[args.mode] / (k_ms * 1e-3) / 1e9
    line = {
        "metric": "generated_samples_per_sec", "value": total_samples / t_dev, "unit": "samples/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16": "f32 math, bf16 resident weights", "bf16-dense": "bf16 tensor-core products, f32 accumulation / state / sampling"}[args.precision],
        "data": "synthetic",
        "rtf": (t_dev / args.steps) / (wl["wave_len"] / wl["sr"]),
        "config": config_dict(args, 1),
        "e2e": {"value": total_samples / t_e2e, "unit": "samples/s", "h2d_bytes_per_step": int(mel_host.numel() * 4),
                "d2h_bytes_per_step": int(wl["wave_len"] * 8), "ms_per_step": t_e2e / args.steps * 1e3,
                "rtf": (t_e2e / args.steps) / (wl["wave_len"] / wl["sr"])},
        "gpu_launches": int((info1.launches - info0.launches) + (info1.epilogue_launches - info0.epilogue_launches) + (cond1 - cond0)),
        "clocks": clocks,
    }
    if kind == 1:
        # the model that binds: per step, four grid-level exchanges of a [512 units x folds]
        # vector (measured: publish + 56 KB warp-local gather on an otherwise empty grid) and the two small hops of the sampler
        # round trip, against the shared-memory operand traffic of the mat-vec passes (one wavefront = 32 lane-words per clock)
        lat_floor = (4 * t_exch + 2 * SMALL_HOP_US) if t_exch else None
        smem_floor = WIDE_SMEM_WAVEFRONTS / clk * 1e6
        floor = max(lat_floor or 0.0, smem_floor, fma_floor)
        line["roofline"] = {"bound": "latency", "kernel": kernel, "achieved": us_per_step, "peak": floor, "unit": "us/step",
                            "frac": floor / us_per_step, "traffic": None, "kernel_ms_per_launch": k_ms,
                            "terms_us": {"exchange_chain": lat_floor, "t_exchange_measured": t_exch, "small_hop": SMALL_HOP_US,
                                         "shared_memory_operands": smem_floor, "fp32_ffma": fma_floor},
                            "note": "lower is better: achieved and peak are microseconds per sample step of all %d folds; frac = floor / achieved. "
                                    "The floor is the largest of the three terms, not their sum (the kernel overlaps them only partly)" % wl["folds"],
                            "hbm": {"achieved_gbs": hbm_gbs, "peak_gbs": pk["hbm_gbs"], "frac": hbm_gbs / pk["hbm_gbs"], "peak_source": pk_src,
                                    "algorithmic_bytes_per_launch": fold_steps * HBM_BYTES_PER_FOLD_STEP[args.mode],
                                    "note": "HBM does not bind this kernel: 840 B per fold-step"},
                            "achieved_tflops_fp32": fold_steps * FLOP_PER_FOLD_STEP[args.mode] / (k_ms * 1e-3) / 1e12,
                            "fp32_peak_tflops_at_clock": fp32_peak}
    else:
        tf = fold_steps * FLOP_PER_FOLD_STEP[args.mode] / (k_ms * 1e-3) / 1e12
        line["roofline"] = {"bound": "tensor" if kind == 2 else "latency", "kernel": kernel, "achieved": tf if kind == 2 else us_per_step,
                            "peak": pk["bf16_tflops"] if kind == 2 else (5 * t_exch if t_exch else None),
                            "unit": "TFLOP/s" if kind == 2 else "us/step",
                            "frac": (tf / pk["bf16_tflops"]) if kind == 2 else ((5 * t_exch / us_per_step) if t_exch else None),
                            "traffic": None, "kernel_ms_per_launch": k_ms,
                            "hbm": {"achieved_gbs": hbm_gbs, "peak_gbs": pk["hbm_gbs"], "frac": hbm_gbs / pk["hbm_gbs"], "peak_source": pk_src}}
    line["step_latency_model"] = {"us_per_step": us_per_step, "kernel": kernel, "folds": wl["folds"], "exchanges_per_step": int(info1.exchanges_per_step),
                                  "t_exchange_us_measured": t_exch}
    if not args.no_dense and args.precision == "fp32" and args.mode == "RAW":
        try:
            line["dense_regime"] = dense_regime(args, dev, model.state_dict(), pk)
        except Exception as e:  # pragma: no cover
            line["dense_regime"] = {"failed": repr(e)}
    if not args.no_sentence_set and args.precision == "fp32" and args.mode == "RAW":
        # configs[3] on this one GPU: what the multi-GPU lines (bench.py --gpus N, same sentence set) scale against
        try:
            ss = sentence_set(args)
            m2 = WaveRNN(**model_kwargs(args.mode, args.geometry)).to(dev)
            m2.load_state_dict(model.state_dict())
            m2.precision = "auto"
            t_d, t_e, _ = time_sentence_set(m2, make_mels(ss["T"], 0, True), dev, 1, 1, barrier)
            line["sentence_set"] = {"workload": config_dict(args, 2)["workload"].replace("dealt to the GPUs by LPT on fold counts", "one GPU"),
                                    "folds": int(sum(ss["folds"])), "wave_len": int(sum(ss["wave_len"])), "seconds_per_step": t_d,
                                    "value": sum(ss["wave_len"]) / t_d, "e2e_value": sum(ss["wave_len"]) / t_e, "unit": "samples/s",
                                    "kernel_ms": m2.last_stats.get("kernel_ms"), "chunks": m2.last_stats.get("chunks"),
                                    "non_step_loop_seconds": t_e - (m2.last_stats.get("kernel_ms") or 0.0) * 1e-3,
                                    "precision": "auto -> bf16-dense (tcgen05)", "rtf": t_e / (sum(ss["wave_len"]) / ss["sr"])}
            del m2
        except Exception as e:  # pragma: no cover
            line["sentence_set"] = {"failed": repr(e)}
    if not args.no_cpu_baseline:
        try:
            cores = os.cpu_count() or 1
            t = cpu_generate(args, model.state_dict(), wl["T"], cores)
            line["cpu_baseline"] = {"value": wl["wave_len"] / t, "unit": "samples/s", "cores": cores, "kind": "port",
                                    "sample": "one full generate() of the workload (%d folds x %d steps, nothing extrapolated), torch %s CPU"
                                              % (wl["folds"], wl["S"], torch.__version__)}
        except Exception as e:  # pragma: no cover
            line["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
    return line


def run_multi(args, dev, local, rank, world, barrier):
    import numpy as np
    import torch
    import torch.distributed as dist
    from expressive_speech_synthesis_research_b200 import WaveRNN, distributed as D
    ss = sentence_set(args)
    torch.manual_seed(0)
    model = WaveRNN(**model_kwargs(args.mode, args.geometry)).to(dev)
    model.precision = "auto"
    plan = D.plan_utterances(ss["folds"], world)                 # whole utterances per rank, LPT on fold counts: no communication
    mine = plan[rank]
    mels_host = [torch.rand(1, 80, ss["T"][i], generator=torch.Generator().manual_seed(i)).pin_memory() for i in mine]
    sampler = ClockSampler(local)
    sampler.start()
    t_dev, t_e2e, launches = time_sentence_set(model, mels_host, dev, args.steps, args.warmup, barrier)
    clocks = sampler.stop()
    tt = torch.tensor([t_dev, t_e2e], device=dev, dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)                    # timing only: max over ranks
    t_dev, t_e2e = tt.tolist()
    ll = torch.tensor([launches, float(sum(ss["folds"][i] for i in mine))], device=dev, dtype=torch.float64)
    dist.all_reduce(ll, op=dist.ReduceOp.SUM)
    total = float(sum(ss["wave_len"]))
    line = None
    long_form = None
    if not args.no_long_form:
        # configs[4]: one 10-minute utterance, folds sharded over the ranks, ONE all_gather of the overlap edges, spans gathered to rank 0
        try:
            sr, hop, _ = GEOMETRY[args.geometry]
            T = int(round(600.0 * sr / hop)) + 1
            mel = torch.rand(1, 80, T, generator=torch.Generator().manual_seed(99))
            times = []
            for it in range(3):
                barrier()
                w0 = time.perf_counter()
                wav = D.generate_sharded(model, mel, TARGET, OVERLAP, True, seed=7 + it)
                barrier()
                times.append(time.perf_counter() - w0)
            t_lf = torch.tensor([min(times[1:])], device=dev, dtype=torch.float64)
            dist.all_reduce(t_lf, op=dist.ReduceOp.MAX)
            long_form = {"workload": "configs[4]: one 10-minute synthetic utterance (%d frames, %d folds), folds sharded over %d GPUs, overlap "
                                     "edges all-gathered over NCCL, spans gathered to rank 0" % (T, fold_count(T * hop), world),
                         "seconds": float(t_lf.item()), "value": (T - 1) * hop / float(t_lf.item()), "unit": "samples/s",
                         "rtf": float(t_lf.item()) / 600.0, "wave_len": (T - 1) * hop,
                         "collective": "all_gather of one [overlap] fp32 edge per rank (2.2 KB) + gather of the float64 spans to rank 0"}
            del wav
        except Exception as e:  # pragma: no cover
            long_form = {"failed": repr(e)}
    if rank == 0:
        loads = [int(sum(ss["folds"][i] for i in p)) for p in plan]
        line = {
            "metric": "generated_samples_per_sec", "value": total / t_dev, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16 tensor-core products, f32 accumulation / state / sampling (precision 'auto' picks the dense kernel for %d folds per GPU)" % max(loads),
            "data": "synthetic", "rtf": t_dev / (total / ss["sr"]),
            "config": config_dict(args, world),
            "e2e": {"value": total / t_e2e, "unit": "samples/s", "h2d_bytes_per_step": int(sum(t * 80 * 4 for t in ss["T"])),
                    "d2h_bytes_per_step": int(total * 8), "ms_per_step": t_e2e * 1e3, "rtf": t_e2e / (total / ss["sr"])},
            "gpu_launches": int(round(ll[0].item() * args.steps)),
            "clocks": clocks,
            "folds_per_gpu": loads,
            "roofline": {"bound": "tensor", "kernel": "wavernn_dense_kernel",
                         "achieved": sum(ss["folds"]) * ss["S"] * FLOP_PER_FOLD_STEP[args.mode] / t_dev / 1e12,
                         "peak": peaks()[0]["bf16_tflops"] * world, "unit": "TFLOP/s",
                         "frac": sum(ss["folds"]) * ss["S"] * FLOP_PER_FOLD_STEP[args.mode] / t_dev / 1e12 / (peaks()[0]["bf16_tflops"] * world),
                         "traffic": None, "note": "whole-call time (conditioning, step loop, epilogue, copies) against the summed bf16 peak of the GPUs"},
            "long_form": long_form,
        }
    return line


if __name__ == "__main__":
    a = parse()
    sys.exit(run_reference(a) if a.impl == "reference" else run_b200(a))
