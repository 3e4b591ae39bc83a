"""bench.py contract on the CPU side: the reference arm prints exactly one JSON
line with the keys the driver reads; the GPU arm refuses to run without a GPU
(no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ, **(env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          env=e, timeout=600)


def test_reference_arm_json_line():
    p = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"],
             {"HRT_REF_PATHS": "64", "HRT_REF_SPLIT": "1"})
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ray-bounces/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("ray-bounces/s on street_canyon_with_cars")
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["workload"].startswith("BASELINE configs[3]") and d["steps"] == 1 and d["n_gpus"] == 1


def test_reference_arm_other_ranks_stay_silent():
    p = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], {"RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_gpu_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    p = _run(["--steps", "1", "--warmup", "0"])
    assert p.returncode != 0 and "no CPU fallback" in (p.stdout + p.stderr)
