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


def _common(d, launches):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "sweeps/s" and d["dtype"] == "f64" and d["scaling"] == "weak" and d["n_gpus"] == 1
    assert d["value"] > 0 and d["steps"] == 3 and d["gpu_launches"] == launches and d["vs_baseline"] is None
    assert "workload" in d["config"] and d["config"]["parity"]["tolerance"] == 1e-6
    errs = [v for k, v in d["config"]["parity"].items() if k.startswith("max_rel_err")]
    assert errs and max(errs) < 1e-6, d["config"]["parity"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert 0 < r["kernel_share_of_step"] <= 1 and r["kernel"] in r["kernels"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert d["config"]["e2e_matches_device_path"] is True
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_b200_arm_json_contract():
    d = _run("--gpus", "1", "--steps", "3", "--warmup", "3", "--batch", "1024", "--ref-per-core", "1")
    _common(d, 4 * 3)                       # rollout/costate + bwd + fwd + batch reduction (no sub-batches at B = 1024)
    assert d["roofline"]["kernel"] == "pdp_k_aux_lqr_bwd"
    par = d["config"]["parity"]
    assert par["shipped_demos"]["max_rel_err"] < 1e-6 and par["shipped_demos"]["status"] == [0, 0]
    assert par["status_bit0_nonfinite"] == par["trajectories_with_nonfinite_outputs"]
    x = d["e2e"]["with_sensitivities_to_host"]
    assert x["matches_device_path"] is True and x["d2h_bytes_per_step"] > 1024 * 51 * 13 * 9 * 8


@pytest.mark.parametrize("cfg,launches,kernel", [("c5", 2, "pdp_k_sens_fwd"), ("c4", 1, "pdp_k_rollout_costate_tma"),
                                                 ("c2", 1, "pdp_k_sens_fwd")])
def test_secondary_config_arms_json_contract(cfg, launches, kernel):
    """bench.py --config c2|c4|c5: the same contract line for the other BASELINE configs (H = 100 for C4 / C5)."""
    d = _run("--config", cfg, "--gpus", "1", "--steps", "3", "--warmup", "3", "--batch", "512", "--ref-per-core", "1")
    _common(d, launches * 3)
    assert d["roofline"]["kernel"] == kernel and cfg.upper() in d["config"]["workload"]


def test_reference_arm_json_contract():
    d = _run("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1", "--ref-per-core", "1")
    assert d["impl"] == "reference" and d["unit"] == "sweeps/s" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    staged = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "reference_src", "PDP.py"))       # oracle/stage_reference.py ran in build()
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port") and d["gpu_launches"] == 0
