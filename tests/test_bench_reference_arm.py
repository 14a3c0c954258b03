"""`bench.py --impl reference` needs no GPU: the contract line of the CPU arm, checked here on a tiny sample."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("cfg", ["c3", "c5"])
def test_reference_arm_line_without_a_gpu(cfg):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", cfg, "--gpus", "1",
                        "--steps", "1", "--warmup", "0", "--ref-per-core", "1"], capture_output=True, text=True, timeout=600,
                       cwd=ROOT, env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "sweeps/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["e2e"] == {"value": d["value"], "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    staged = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "reference_src", "PDP.py")) or \
        os.path.isfile("/root/reference/PDP/PDP.py")
    cb = d["cpu_baseline"]
    assert cb["kind"] == ("reference" if staged else "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert cfg.upper() in d["config"]["workload"] and d["gpu_launches"] == 0


def test_reference_arm_under_torchrun_only_rank_zero_prints():
    """N > 1: the driver launches the reference arm under torchrun too; rank 0 alone runs and prints, the others exit 0."""
    env = {**os.environ, "CUDA_VISIBLE_DEVICES": "", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29577"}
    outs = []
    for rank in (0, 1):
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                            "--warmup", "0", "--ref-per-core", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT,
                           env={**env, "RANK": str(rank), "LOCAL_RANK": str(rank)})
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append([ln for ln in p.stdout.splitlines() if ln.startswith("{")])
    assert len(outs[0]) == 1 and outs[1] == []
    assert json.loads(outs[0][0])["n_gpus"] == 2
