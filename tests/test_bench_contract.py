"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the agreed keys, and the
product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tests/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("ray-triangle tests/s") and d["value"] > 1e6
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "2048x2048" in d["config"]["workload"]
    assert d["gpu_launches"] == 0 and d["n_gpus"] == 1 and d["steps"] == 1


def test_reference_arm_other_ranks_exit_quietly():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=str(ROOT), env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_product_arm_fails_loudly_without_a_gpu():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
    assert not any(l.startswith("{") for l in r.stdout.splitlines())      # no number is ever printed from a CPU fallback
