"""bench.py prints ONE JSON line with the contracted keys (both arms), on a small batch so the test stays quick."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *extra], capture_output=True, text=True, timeout=600,
                       cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    return json.loads(lines[0])


def test_b200_arm_json_contract():
    d = _run("--gpus", "1", "--steps", "3", "--warmup", "3", "--batch", "1024", "--ref-per-core", "1")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "sweeps/s" and d["dtype"] == "f64" and d["scaling"] == "weak" and d["n_gpus"] == 1
    assert d["value"] > 0 and d["steps"] == 3 and d["gpu_launches"] == 9 and d["vs_baseline"] is None
    assert "workload" in d["config"] and d["config"]["parity_max_rel_err_vs_oracle_first4"] < 1e-6
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["kernel"] == "pdp_k_aux_lqr_bwd" and 0 < r["kernel_share_of_step"] < 1
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert d["config"]["e2e_matches_device_path"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_reference_arm_json_contract():
    d = _run("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1", "--ref-per-core", "1")
    assert d["impl"] == "reference" and d["unit"] == "sweeps/s" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["gpu_launches"] == 0
