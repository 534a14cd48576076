"""bench.py's reference arm runs on the host CPU (the C/OpenMP port of the reference's three-pass step): its JSON line has
the keys the driver reads, and the product arm refuses to run without a GPU instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                          cwd=ROOT, env=dict(os.environ, **(env or {})))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    res = _run("--impl", "reference", "--workload", "cylinder", "--steps", "3", "--warmup", "1")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MLUPS" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["steps"] == 3 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["gpu_launches"] == 0 and d["dtype"] == "f32"
    assert d["value"] > 0 and abs(d["value"] - 512 * 128 * 3 / (d["ms_per_step"] * 3e-3) / 1e6) < 1e-6 * d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["grid"] == [512, 128] and "workload" in d["config"]


def test_reference_arm_runs_on_rank_0_only():
    res = _run("--impl", "reference", "--workload", "cylinder", "--steps", "1", "--warmup", "1", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and not [l for l in res.stdout.splitlines() if l.startswith("{")]


def test_product_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    res = _run("--steps", "1", "--warmup", "1", "--workload", "cylinder")
    assert res.returncode != 0 and "no CPU fallback" in (res.stderr + res.stdout)
