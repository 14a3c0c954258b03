"""The reference's own Examples scripts, UNMODIFIED, against the drop-in packages (build container only: the
reference tree is not present on the GPU box, and this container has no GPU).  Each script must get through its
imports, environment / OCSys / SysID construction, data loading and symbolic differentiation with this repo's
`PDP`, `JinEnv` and `casadi` packages, and stop exactly where the first hot-path method needs the CUDA device
(``PDPBackendError``) -- i.e. nothing in the scripts' use of the class surface is missing."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = "/root/reference/Examples"
# the OC scripts import matplotlib.pyplot at the top (absent here): they get an import-only stub through the harness
OC_SCRIPTS = ["OC/cartpole/cartpole_PDP_poly.py", "OC/cartpole/cartpole_PDP_neural.py", "OC/rocket/rocket_PDP_Recmat.py",
              "OC/quadrotor/uav_PDP.py", "OC/robotarm/robotarm_PDP_Recmat.py"]
SCRIPTS = ["IRL/quadrotor/uav_PDP.py", "IRL/pendulum/pendulum_PDP.py", "IRL/cartpole/cartpole_PDP.py",
           "IRL/robotarm/robotarm_PDP.py", "IRL/rocket/rocket_PDP.py", "SysID/quadrotor/uav_PDP.py",
           "SysID/cartpole/cartpole_PDP.py", "SysID/pendulum/pendulum_PDP.py", "SysID/robotarm/robotarm_PDP.py",
           "SysID/rocket/rocket_PDP.py", "SysID/robotarm/robotarm_PDP_neural.py"]


@pytest.mark.skipif(not os.path.isdir(EX), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("script", SCRIPTS)
def test_unmodified_reference_script_reaches_the_cuda_boundary(script):
    import torch
    if torch.cuda.is_available():
        pytest.skip("meant for the CPU-only build container")
    path = os.path.join(EX, script)
    env = dict(os.environ, PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable, path], cwd=os.path.dirname(path), env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode != 0
    err = p.stderr
    assert "PDPBackendError" in err, err[-1500:]
    assert "the PDP B200 engine needs a CUDA device" in err
    # the failure must come from a hot-path call made by the script itself, not from an import / setup problem
    assert any(k in err for k in ("ocSolver", "step(", "integrateDyn", "getAuxSys")), err[-1500:]


@pytest.mark.skipif(not os.path.isdir(EX), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("script", OC_SCRIPTS)
def test_unmodified_reference_oc_script_reaches_the_cuda_boundary(script):
    """Examples/OC scripts through tools/run_unmodified_script.py (import-only matplotlib stub): everything up to the first
    hot-path call (the ground-truth OCSys.ocSolver) works with the drop-in packages."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("meant for the CPU-only build container")
    path = os.path.join(EX, script)
    if not os.path.isfile(path):
        pytest.skip("script not in this reference snapshot")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_unmodified_script.py"), path, "--seconds", "250",
                        "--stub-matplotlib"], capture_output=True, text=True, timeout=300)
    assert p.returncode != 0
    err = p.stderr
    assert "PDPBackendError" in err, err[-1500:]
    assert "the PDP B200 engine needs a CUDA device" in err
    assert any(k in err for k in ("ocSolver", "step(", "integrateSys")), err[-1500:]
