#!/usr/bin/env python
"""bench.py -- WaveRNN batched generation throughput (BASELINE.json metric) on N B200s.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (torch port of it)

A "step" is one pass of the hot path over one batch of synthetic input: one
`generate(mel, batched=True, target=11000, overlap=550, mu_law=True)` of a 10 s utterance
(BASELINE.json configs[1]: RAW 9-bit, random-init weights, 22.05 kHz / hop 275 -> 20 folds of
12100 steps).  N > 1 (torchrun): every rank vocodes its own utterance (weak scaling, no
data-path collective); value = samples of all ranks / max-over-ranks device time.
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
                    help="fp32 (headline) | bf16 resident weights of the FFMA kernel (BASELINE.json configs[2]) | bf16-dense: the "
                         "tcgen05 / tensor-memory kernel for large fold batches")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-regime (pooled folds, tcgen05) measurement")
    ap.add_argument("--dense-folds", type=int, default=480)
    ap.add_argument("--dense-steps", type=int, default=2000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-sample-steps", type=int, default=1200)
    return ap.parse_args()


def model_kwargs(mode, geometry):
    sr, hop, ups = GEOMETRY[geometry]
    return dict(rnn_dims=512, fc_dims=512, bits=9, pad=2, upsample_factors=ups, feat_dims=80, compute_dims=128,
                res_out_dims=128, res_blocks=10, hop_length=hop, sample_rate=sr, mode=mode)


def workload(args):
    sr, hop, _ = GEOMETRY[args.geometry]
    T = int(round(args.seconds * sr / hop)) + 1
    L = T * hop
    n = (L - OVERLAP) // (TARGET + OVERLAP)
    if L - (n * (TARGET + OVERLAP) + OVERLAP) != 0:
        n += 1
    return dict(T=T, hop=hop, sr=sr, L=L, folds=n, S=TARGET + 2 * OVERLAP, wave_len=(T - 1) * hop)


def config_dict(args, wl, n_gpus):
    return {"workload": "configs[1]: WaveRNN %s 9-bit batched generate, target=11000 overlap=550, one %.0f s synthetic utterance per GPU"
                        % (args.mode, args.seconds),
            "geometry": "%s (%d Hz, hop %d)" % (args.geometry, wl["sr"], wl["hop"]),
            "mel_frames": wl["T"], "folds": wl["folds"], "steps_per_fold": wl["S"], "wave_len": wl["wave_len"],
            "utterances_per_step": n_gpus, "parallelism": "utterance-per-gpu x%d" % n_gpus,
            "weights": "random-init (torch.manual_seed(0))",
            "l2": "per-step inputs exceed L2: %.0f MB of upsampled conditioning rewritten every step" % (wl["L"] * 208 * 4 / 1e6)}


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
# the reference's CPU implementation (torch port of fatchord_version.py::generate), bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(args, wl, state, sample_steps, repeats=1, warmup=0):
    """Times `sample_steps` of the step loop over ALL folds of the workload (plus the conditioning
    network and fold) on the host cores with torch's own threading; returns per-repeat useful
    samples/s extrapolated by fold-steps (S / sample_steps)."""
    import torch
    from oracle import torch_port
    _, _, ups = GEOMETRY[args.geometry]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = {k: v.detach().cpu() for k, v in state.items()}
    mel = torch.rand(1, 80, wl["T"], generator=torch.Generator().manual_seed(0))
    times = []
    for it in range(warmup + repeats):
        t0 = time.perf_counter()
        with torch.no_grad():
            m, aux = torch_port.conditioning(sd, mel, ups, 2)
            m = torch_port.fold_with_overlap(m, TARGET, OVERLAP)[:, :sample_steps].contiguous()
            aux = torch_port.fold_with_overlap(aux, TARGET, OVERLAP)[:, :sample_steps].contiguous()
            t1 = time.perf_counter()
            torch_port.step_loop(sd, args.mode, m, aux, generator=torch.Generator().manual_seed(it))
        t2 = time.perf_counter()
        full = (t1 - t0) + (t2 - t1) * wl["S"] / sample_steps          # conditioning once + all S steps
        if it >= warmup:
            times.append(full)
    desc = "%d of %d steps of the %d-fold step loop + conditioning/fold, torch %s CPU, extrapolated by steps" % (
        sample_steps, wl["S"], wl["folds"], torch.__version__)
    return [wl["wave_len"] / t for t in times], times, cores, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from expressive_speech_synthesis_research_b200 import WaveRNN
    wl = workload(args)
    torch.manual_seed(0)
    state = WaveRNN(**model_kwargs(args.mode, args.geometry)).state_dict()
    vals, times, cores, desc = cpu_reference_sample(args, wl, state, args.ref_sample_steps, repeats=args.steps,
                                                    warmup=args.warmup)
    t = sum(times) / len(times)
    value = wl["wave_len"] / t
    line = {"impl": "reference", "metric": "generated_samples_per_sec", "value": value, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "rtf": t / (wl["wave_len"] / wl["sr"]), "config": config_dict(args, wl, 1),
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def dense_regime(args, dev, state, pk):
    """BASELINE.json configs[3]/[4] regime: hundreds of pooled folds per GPU (sentence sets, long-form).  Times the
    step loop alone (wrnn_generate_folds, conditioning resident in HBM, in-kernel Philox draws) for the tcgen05 kernel
    (precision bf16-dense) and for the fp32 FFMA kernel on the same folds."""
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
        nb = B if prec == "bf16-dense" else min(B, 64)           # the FFMA kernel advances 64 folds per launch: time one launch
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
                       "frac": tf / pk["bf16_tflops"],
                       # ncu --set full of this kernel at 480 folds x 300 steps: dram read 127.8 MB + write 4.1 MB (profiles/r01_dense.md
                       # section 3; algorithmic 832 B per fold-step = 119.8 MB), scaled to this launch's fold-steps
                       "traffic": (127.8e6 + 4.1e6) / (480 * 300) * d["folds"] * args.dense_steps if d["folds"] == 480 else None,
                       "traffic_unit": "dram bytes per launch (ncu capture of 480 x 300, scaled by fold-steps)",
                       "note": "algorithmic FLOPs (8.14 MFLOP per fold-step); every tcgen05.mma is M=128 x N=32 x K=16 and is paced by "
                               "its shared-memory operand reads (40 clk measured, scripts/umma_rate.cu), the step by the shared-memory "
                               "port (weights cross it twice: TMA fill + MMA read) and the cluster exchange (DESIGN.md 9)"}
    return out


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
    from expressive_speech_synthesis_research_b200 import WaveRNN

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

    wl = workload(args)
    torch.manual_seed(0)
    model = WaveRNN(**model_kwargs(args.mode, args.geometry)).to(dev)
    model.precision = args.precision
    mel_host = torch.rand(1, 80, wl["T"], generator=torch.Generator().manual_seed(rank)).pin_memory()
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
    # ---- end-to-end leg (host buffers, H2D + D2H inside the timed region) ----------------------
    barrier()
    ev0.record()
    w0 = time.perf_counter()
    for i in range(args.steps):
        wav = e2e_step(200 + i)
    ev1.record()
    barrier()
    t_e2e = max(ev0.elapsed_time(ev1) / 1e3, 0.0)
    t_e2e_wall = time.perf_counter() - w0
    clocks = sampler.stop()
    model.train()

    if world > 1:
        tt = torch.tensor([t_dev, t_e2e, t_e2e_wall], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, t_e2e_wall = tt.tolist()
    n = world
    total_samples = n * wl["wave_len"] * args.steps
    value = total_samples / t_dev
    e2e_value = total_samples / max(t_e2e, t_e2e_wall)

    line = None
    if rank == 0:
        pk, pk_src = peaks()
        fold_steps = wl["folds"] * wl["S"]
        k_ms = sum(kernel_ms) / len(kernel_ms)
        alg_bytes = fold_steps * HBM_BYTES_PER_FOLD_STEP[args.mode]
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        t_sync = None
        try:
            t_sync = eng.measure_exchange(4000)
        except Exception as e:  # pragma: no cover
            t_sync = None
        us_per_step = k_ms * 1e3 / wl["S"]
        groups = (wl["folds"] + 7) // 8
        teams = min(groups, 3)
        fp32_peak = 148 * 128 * 2 * (clocks["sm_mhz"] or 1965.0) * 1e6 / 1e12
        lat_floor = 5 * t_sync if t_sync else None
        fma_floor = wl["folds"] * FLOP_PER_FOLD_STEP[args.mode] / (fp32_peak * 1e12) * 1e6
        dense = args.precision == "bf16-dense"
        # dram__bytes_read.sum + dram__bytes_write.sum of the one ncu --set full capture of this kernel on this workload
        # (20 folds x 3000 steps of configs[1]: 66.2 MB + 3.8 MB, profiles/r01_summary.md section 3), scaled to this launch's
        # fold-steps; other modes / kernels have no capture of their own and report null
        traffic = None
        if not dense and args.precision == "fp32" and args.mode == "RAW" and wl["folds"] == 20:
            traffic = (66.2e6 + 3.8e6) / (20 * 3000) * fold_steps
        line = {
            "metric": "generated_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": n,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "bf16": "f32 math, bf16 resident weights", "bf16-dense": "bf16 tensor-core products, f32 accumulation / state / sampling"}[args.precision],
            "data": "synthetic",
            "rtf": (t_dev / args.steps) / (wl["wave_len"] / wl["sr"]),
            "config": config_dict(args, wl, n),
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(mel_host.numel() * 4),
                    "d2h_bytes_per_step": int(wl["wave_len"] * 8), "ms_per_step": max(t_e2e, t_e2e_wall) / args.steps * 1e3,
                    "rtf": (max(t_e2e, t_e2e_wall) / args.steps) / (wl["wave_len"] / wl["sr"])},
            "gpu_launches": int((info1.launches - info0.launches) + (info1.epilogue_launches - info0.epilogue_launches)),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "wavernn_dense_kernel" if dense else "wavernn_persistent_kernel", "achieved": achieved, "peak": pk["hbm_gbs"],
                         "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk_src,
                         "kernel_ms_per_launch": k_ms, "algorithmic_bytes_per_launch": alg_bytes,
                         "traffic_unit": "bytes per launch (ncu dram read + write of a 20 x 3000 capture, scaled by fold-steps)",
                         "note": "HBM is not what bounds this kernel (840 B per fold-step, measured dram traffic ~0.7 GB/s); "
                                 "the binding terms are the step-latency model below (SURVEY.md 8d, DESIGN.md 7): five "
                                 "dependent grid-level exchanges per step plus the FFMA2 / shared-memory floors"},
            "step_latency_model": {"us_per_step": us_per_step, "exchanges_per_step": 5, "groups": groups, "teams_per_cta": teams,
                                   "t_exchange_us_measured": t_sync, "latency_floor_us": lat_floor,
                                   "fp32_fma_floor_us": fma_floor, "fp32_peak_tflops_at_clock": fp32_peak,
                                   "frac_of_floor": (max(lat_floor or 0.0, fma_floor) / us_per_step) if us_per_step else None,
                                   "achieved_tflops": fold_steps * FLOP_PER_FOLD_STEP[args.mode] / (k_ms * 1e-3) / 1e12},
        }
        if dense:
            line["step_latency_model"] = {"us_per_step": us_per_step, "folds": wl["folds"], "clusters": (wl["folds"] + 31) // 32,
                                          "achieved_tflops": fold_steps * FLOP_PER_FOLD_STEP[args.mode] / (k_ms * 1e-3) / 1e12}
        if n == 1 and not args.no_dense and not dense and args.mode == "RAW":
            try:
                line["dense_regime"] = dense_regime(args, dev, model.state_dict(), pk)
            except Exception as e:  # pragma: no cover
                line["dense_regime"] = {"failed": repr(e)}
        if n == 1 and not args.no_cpu_baseline:
            try:
                vals, times, cores, desc = cpu_reference_sample(args, wl, model.state_dict(), 600, repeats=1, warmup=0)
                line["cpu_baseline"] = {"value": vals[0], "unit": "samples/s", "cores": cores, "kind": "port", "sample": desc}
            except Exception as e:  # pragma: no cover
                line["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": "failed: %r" % (e,)}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse()
    sys.exit(run_reference(a) if a.impl == "reference" else run_b200(a))
