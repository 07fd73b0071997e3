"""bench.py's output contract, checked on CPU through the reference arm (the b200 arm needs a GPU and is run by
tests/test_gpu_parity.py's box): stdout carries exactly one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "cpu_baseline", "impl")


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--seconds", "1.0"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout[-2000:]
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "generated_samples_per_sec" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"] and "model" not in d["config"]
    assert "nothing extrapolated" in d["cpu_baseline"]["sample"]            # the reference arm runs the workload in full
    assert d["cpu_baseline_1thread"]["cores"] == 1 and d["cpu_baseline_1thread"]["value"] > 0


def test_multi_gpu_config_names_the_sentence_set():
    """N > 1 runs BASELINE.json configs[3] (strong scaling): both arms print the same config for a given N."""
    sys.path.insert(0, ROOT)
    import bench
    args = bench.argparse.Namespace(geometry="fatchord", mode="RAW", seconds=10.0, utterances=256)
    c1, c8 = bench.config_dict(args, 1), bench.config_dict(args, 8)
    assert c1["workload"].startswith("configs[1]") and c1["folds"] == 20 and c1["wave_len"] == 220550
    assert c8["workload"].startswith("configs[3]") and c8["utterances"] == 256
    assert c8["folds"] == 3719 and c8["wave_len"] == 41561850                # T = round(D sr / hop) + 1, D = default_rng(0).uniform(2, 12, 256)
    args.geometry = "ref"
    c8 = bench.config_dict(args, 8)
    assert c8["folds"] == 2734 and c8["wave_len"] == 30158000                # SURVEY.md 8d, config 4 (ref geometry)
