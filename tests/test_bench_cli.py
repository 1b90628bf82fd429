"""bench.py contract checks that need no GPU: the reference arm (the reference's CPU path as restated by oracle/) prints
one JSON line with the contract's keys, and the product arm fails loudly when there is no CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _run(*args, timeout=600):
    env = dict(os.environ, OMP_NUM_THREADS="4")
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["n_gpus"] == 1 and d["dtype"] == "f32"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "MME U-Net 64x64" in d["config"]["workload"] and "model" not in d["config"]
    # the driver compares the two arms' `config` and `warmup`: both come from the same helper / rule as the product arm's
    sys.path.insert(0, str(ROOT))
    import bench
    assert d["config"] == bench.line_config(16, 1, 4096)
    assert d["warmup"] == max(1, 3)


def test_reference_arm_config_follows_the_requested_gpu_count():
    import bench
    c1, c8 = bench.line_config(16, 1, 4096), bench.line_config(16, 8, 4096)
    assert c1["parallelism"] == "single" and c8["parallelism"] == "dp8" and c8["global_batch"] == 128
    assert set(c1) == set(c8) and "l2" in c1


def test_product_arm_fails_loudly_without_a_gpu():
    from s2s_ismr_unet_b200.runtime import device_count
    if device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline", timeout=300)
    assert r.returncode != 0
    assert "S2SError" in r.stderr or "CUDA" in r.stderr or "cuda" in r.stderr
