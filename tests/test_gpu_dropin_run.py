"""Drop-in proof on hardware: the reference's own Examples scripts, UNMODIFIED, executed on the GPU against this repo's
PDP / JinEnv / casadi packages (north star: "the Examples/ scripts drop in unchanged").

The scripts come from baseline/_ref/Examples (staged from the read-only reference tree by tools/stage_reference_examples.py
in the build container; git-ignored, shipped to the GPU box with the snapshot) or from $PDP_REFERENCE_EXAMPLES.  They are run
by tools/run_unmodified_script.py, which controls only their environment (import path, a stop after a few printed
iterations, and the values their initial np.random draw returns so that a stored trial can be reproduced).  Checked:
the losses the scripts print against the loss_trace of the reference's shipped result file (IRL pendulum, K3) and
against the oracle at the same parameters (SysID quadrotor, OC cart-pole)."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import scipy.io as sio
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.environ.get("PDP_REFERENCE_EXAMPLES", os.path.join(ROOT, "baseline", "_ref", "Examples"))
G = os.path.join(ROOT, "tests", "golden")
NUM = r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?"


def _run(rel, seconds, max_prints, random=None, stub=False):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    path = os.path.join(EX, rel)
    if not os.path.isfile(path):
        pytest.skip("reference Examples not staged (run tools/stage_reference_examples.py in the build container)")
    cmd = [sys.executable, os.path.join(ROOT, "tools", "run_unmodified_script.py"), path, "--seconds", str(seconds),
           "--max-prints", str(max_prints)]
    if random:
        cmd += ["--random", json.dumps(random)]
    if stub:
        cmd += ["--stub-matplotlib"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=seconds + 240)
    assert p.returncode == 0, p.stderr[-3000:]
    return p.stdout


def test_unmodified_irl_pendulum_script_reproduces_the_shipped_loss_trace():
    """Examples/IRL/pendulum/pendulum_PDP.py: ocSolver -> getAuxSys -> lqrSolver -> chain rule through the legacy API.
    With the shipped trial's initial parameter the printed losses equal the reference's own loss_trace (K3)."""
    d = os.path.join(EX, "IRL", "pendulum", "data")
    if not os.path.isfile(os.path.join(d, "PDP_results_trial_0.mat")):
        pytest.skip("reference Examples not staged")
    res = sio.loadmat(os.path.join(d, "PDP_results_trial_0.mat"))["results"][0, 0]
    true = np.asarray(sio.loadmat(os.path.join(d, "pendulum_demos.mat"))["true_parameter"], dtype=np.float64)
    init = np.asarray(res["initial_parameter"], dtype=np.float64).reshape(true.shape)
    trace = np.asarray(res["loss_trace"], dtype=np.float64).ravel()
    lr_stored = float(np.asarray(res["learning_rate"]).ravel()[0])
    sigma = 0.9                                                     # pendulum_PDP.py:38
    delta = init - true
    assert np.ptp(delta) < 1e-12                                    # one scalar draw (len((1,r) array) == 1)
    u = (float(delta.ravel()[0]) + sigma / 2) / sigma
    out = _run("IRL/pendulum/pendulum_PDP.py", 120, 3, random={"random": [[u]]})
    losses = [float(m.group(1)) for m in re.finditer(r"trial # 0 iter:\s+\d+\s+loss:\s+\[?(%s)" % NUM, out)]
    assert len(losses) >= 3, out[-2000:]
    n_cmp = 3 if abs(lr_stored - 1e-5) < 1e-12 else 1              # later iterates depend on the learning rate used
    for k in range(n_cmp):
        assert abs(losses[k] - trace[k]) <= 1e-6 * abs(trace[k]), (k, losses[k], trace[k])


def test_unmodified_sysid_quadrotor_script_matches_oracle():
    """Examples/SysID/quadrotor/uav_PDP.py: SysID.step in its gradient-descent loop; the losses printed at iterations 0 and
    100 equal the oracle's for the same initial parameter and the script's learning rate."""
    from oracle import envs, pdp_oracle
    g = np.load(os.path.join(G, "k1_iodata.npz"))
    inputs, states, true = g["quadrotor_inputs"], g["quadrotor_states"], g["quadrotor_true_parameter"]
    u, sigma, lr = 0.25, 0.6, 1e-4                                 # uav_PDP.py:37-40
    out = _run("SysID/quadrotor/uav_PDP.py", 120, 2, random={"rand": [[u]]})
    losses = {int(m.group(1)): float(m.group(2)) for m in re.finditer(r"Trial: 0 Iter: (\d+) loss: (%s)" % NUM, out)}
    assert 0 in losses and 100 in losses, out[-2000:]
    e = envs.quadrotor(c=0.01)
    sid = pdp_oracle.OracleSysID(e["X"], e["U"], e["dyn_params"], e["X"] + 0.1 * e["f"])
    th = true + sigma * u - sigma / 2
    ref = {}
    for k in range(101):
        loss, dp = sid.step(list(inputs), list(states), th)
        ref[k] = loss
        th = th - lr * dp
    for k in (0, 100):
        assert abs(losses[k] - ref[k]) <= 1e-8 * abs(ref[k]), (k, losses[k], ref[k])


def test_unmodified_oc_cartpole_script_matches_oracle():
    """Examples/OC/cartpole/cartpole_PDP_poly.py (imports matplotlib at the top: import-only stub): OCSys.ocSolver for the
    ground truth, then ControlPlanning.init_step / step in its loop; printed losses at iterations 0 and 100 vs the oracle."""
    from oracle import envs, pdp_oracle
    th0 = np.array([0.3, -0.2, 0.5, 0.1, -0.4, 0.25])
    out = _run("OC/cartpole/cartpole_PDP_poly.py", 180, 3, random={"randn": [th0.tolist()]}, stub=True)
    losses = {int(m.group(1)): float(m.group(2)) for m in re.finditer(r"Trial: 0 Iter: (\d+) loss: (%s)" % NUM, out)}
    assert 0 in losses and 100 in losses, out[-2000:]
    true_cost = float(re.search(r"\[\[(%s)\]\]" % NUM, out).group(1))
    e = envs.cartpole(mc=0.1, mp=0.1, l=1, wx=0.1, wq=0.6, wdx=0.1, wdq=0.1, wu=0.3)
    cp = pdp_oracle.OracleCP(e["X"], e["U"], e["X"] + 0.05 * e["f"], e["path_cost"], e["final_cost"])
    H = 25
    cp.set_poly(np.linspace(0, H, 6))
    th, ref = th0.copy(), {}
    for k in range(101):
        loss, dp = cp.step(np.zeros(4), H, th)
        ref[k] = loss
        th = th - 1e-3 * dp
    for k in (0, 100):
        assert abs(losses[k] - ref[k]) <= 1e-8 * abs(ref[k]), (k, losses[k], ref[k])
    assert true_cost <= min(ref.values()) + 1e-9                   # the OC optimum bounds every policy's cost from below
