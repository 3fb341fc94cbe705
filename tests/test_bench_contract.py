"""bench.py contract on CPU: the reference arm (the CPU oracle port timed on the host cores) prints exactly ONE JSON line on
stdout with the keys the driver reads, and the GPU arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          timeout=600, env={**os.environ, **(env or {})})


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--rays", "64")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("train rays/sec") and d["value"] > 0 and d["n_gpus"] == 1
    for k in ("steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k
    assert d["steps"] == 1 and d["warmup"] == 1                     # the arm honours --steps / --warmup
    assert "workload" in d["config"] and "model" not in d["config"]
    # the config names the workload only (identical in both arms): nothing GPU-specific is claimed by the CPU run
    assert "mlp_precision" not in d["config"] and "launch" not in d["config"] and "CPU" in d["launch"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == dict(value=d["value"], unit="rays/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_reference_arm_uses_all_host_threads_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm must still use every core it may run on."""
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--rays", "64", "--gpus", "2",
             env=dict(RANK="0", WORLD_SIZE="2", LOCAL_RANK="0", OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) and d["n_gpus"] == 2


def test_reference_arm_other_configs():
    for cfg, metric in (("c4", "eval rays/sec"), ("c5", "train rays/sec")):
        r = _run("--impl", "reference", "--config", cfg, "--steps", "1", "--warmup", "0", *(["--rays", "64"] if cfg == "c5" else []))
        assert r.returncode == 0, r.stderr[-2000:]
        d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
        assert d["metric"].startswith(metric) and d["value"] > 0 and d["config"]["workload"].startswith(cfg)
        assert d["scaling"] == ("strong" if cfg == "c4" else "weak")


def test_reference_arm_other_ranks_print_nothing():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--rays", "64", "--gpus", "2",
             env=dict(RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
