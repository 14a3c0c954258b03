"""Sharded outer loops on >= 2 GPUs (skipped on a single-GPU box): tools/multigpu_check.py under torchrun + NCCL."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_outer_loops_two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tools", "multigpu_check.py")],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, (p.stdout[-1500:], p.stderr[-3000:])
    rep = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
    assert rep["world"] == 2 and rep["sysid_sharded_vs_single_gpu_rel"] < 1e-12
    assert rep["sysid_graph_vs_eager_gd"] < 1e-12 and rep["sysid_graph_vs_eager_adam"] < 1e-12
    assert rep["irl_graph_vs_eager_sharded"] < 1e-6


def test_single_gpu_graph_iterations():
    """The same script on ONE rank: graph-captured SysID / IRL iterations (gradient descent and Adam) equal the eager ones."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "1", "--master-addr",
                        "127.0.0.1", "--master-port", "29518", os.path.join(ROOT, "tools", "multigpu_check.py")],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, (p.stdout[-1500:], p.stderr[-3000:])
    rep = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
    assert rep["world"] == 1 and rep["sysid_graph_vs_eager_adam"] < 1e-12 and rep["irl_graph_vs_eager_sharded"] < 1e-6
