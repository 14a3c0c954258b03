"""Cross-compile (no GPU needed) the modules the GPU tests create through the drop-in classes, so the GPU box
does not spend its minutes in nvcc.  Run after __graft_entry__.build()."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from PDP import PDP  # noqa: E402
from JinEnv import JinEnv  # noqa: E402
from casadi import vertcat  # noqa: E402
from pontryagin_differentiable_programming_b200 import engine, ocsolver  # noqa: E402

dt11 = np.array([[0.1]])
for env in ("pendulum", "quadrotor", "robotarm", "rocket", "cartpole"):
    if env == "pendulum":
        e = JinEnv.SinglePendulum(); e.initDyn(); e.initCost()
    elif env == "quadrotor":
        e = JinEnv.Quadrotor(); e.initDyn(c=0.01); e.initCost(wthrust=0.1)
    elif env == "robotarm":
        e = JinEnv.RobotArm(); e.initDyn(g=0); e.initCost(wu=0.01)
    elif env == "cartpole":
        e = JinEnv.CartPole(); e.initDyn(); e.initCost(wu=0.1)
    else:
        e = JinEnv.Rocket(); e.initDyn(); e.initCost(wthrust=0.1)
    oc = PDP.OCSys()
    oc.setAuxvarVariable(vertcat(e.dyn_auxvar, e.cost_auxvar))
    oc.setControlVariable(e.U)
    oc.setStateVariable(e.X)
    oc.setDyn(e.X + dt11 * e.f)
    oc.setPathCost(e.path_cost)
    oc.setFinalCost(e.final_cost)
    print(env, ocsolver.newton_system(oc._system()).module_path)

rocket = JinEnv.Rocket()
rocket.initDyn(Jx=0.5, Jy=1., Jz=1., mass=1., l=1.)
rocket.initCost(wr=1, wv=1, wtilt=50, ww=1, wsidethrust=1, wthrust=0.4)
cp = PDP.ControlPlanning()
cp.setStateVariable(rocket.X)
cp.setControlVariable(rocket.U)
cp.setDyn(rocket.X + 0.1 * rocket.f)
cp.setPathCost(rocket.path_cost)
cp.setFinalCost(rocket.final_cost)
print("rocket recmat", cp._oc_system().module_path)

# neural-dynamics SysID module of tests/test_gpu_modes.py::test_sysid_neural_dynamics_matches_oracle
from casadi import SX, mtimes, tanh, vcat  # noqa: E402
arm = JinEnv.RobotArm()
arm.initDyn(g=0)
inp = vertcat(arm.X, arm.U)
M1, b1 = SX.sym('M1', 6, 6), SX.sym('b1', 6)
M2, b2 = SX.sym('M2', 4, 6), SX.sym('b2', 4)
net = mtimes(M2, tanh(mtimes(M1, inp) + b1)) + b2
sid = PDP.SysID()
sid.setAuxvarVariable(vcat([M1.reshape((-1, 1)), b1.reshape((-1, 1)), M2.reshape((-1, 1)), b2.reshape((-1, 1))]))
sid.setStateVariable(arm.X)
sid.setControlVariable(arm.U)
sid.setDyn(net)
print("neural sysid", sid._system().module_path)

# modules of tests/test_gpu_modes.py::test_legacy_getauxsys_of_sysid_and_controlplanning
uav = JinEnv.Quadrotor()
uav.initDyn(c=0.01)
sid2 = PDP.SysID()
sid2.setAuxvarVariable(uav.dyn_auxvar)
sid2.setStateVariable(uav.X)
sid2.setControlVariable(uav.U)
sid2.setDyn(uav.X + 0.1 * uav.f)
sid2._system()
for fn in (sid2.dfx_fn, sid2.dfe_fn):
    engine.GpuFunction(fn)
engine.DenseLQR.get(13, 1, 5)
engine.DenseLQR.get(4, 1, 6)
cart = JinEnv.CartPole()
cart.initDyn(mc=0.1, mp=0.1, l=1)
cart.initCost(wx=0.1, wq=0.6, wdx=0.1, wdq=0.1, wu=0.3)
cpl = PDP.ControlPlanning()
cpl.setStateVariable(cart.X)
cpl.setControlVariable(cart.U)
cpl.setDyn(cart.X + 0.05 * cart.f)
cpl.setPathCost(cart.path_cost)
cpl.setFinalCost(cart.final_cost)
cpl.init_step(15)
cpl._cp_system()
for nm in ("dfx_fn", "dfu_fn", "dpolicy_dx_fn", "dpolicy_de_fn"):
    engine.GpuFunction(getattr(cpl, nm))
print("legacy-chain modules built")

# the one-trajectory-per-warp build of the quadrotor module (tests/test_gpu_fullsize.py compares the two layouts)
from tools.tune_aux_lqr import make, V1  # noqa: E402
print("quadrotor, one trajectory per warp", make(**V1).module_path)

# modules of tests/test_gpu_dropin_run.py (the unmodified reference scripts executed on the GPU box)
cpl25 = PDP.ControlPlanning()
cpl25.setStateVariable(cart.X)
cpl25.setControlVariable(cart.U)
cpl25.setDyn(cart.X + 0.05 * cart.f)
cpl25.setPathCost(cart.path_cost)
cpl25.setFinalCost(cart.final_cost)
cpl25.init_step(25)
print("cartpole ControlPlanning poly H=25", cpl25._cp_system().module_path)
toc = PDP.OCSys()
toc.setStateVariable(cart.X)
toc.setControlVariable(cart.U)
toc.setDyn(cart.X + 0.05 * cart.f)
toc.setPathCost(cart.path_cost)
toc.setFinalCost(cart.final_cost)
toc.setAuxvarVariable()
print("cartpole ground-truth OCSys", ocsolver.newton_system(toc._system()).module_path)
pen = JinEnv.SinglePendulum()
pen.initDyn()
pen.initCost()
poc = PDP.OCSys()
poc.setAuxvarVariable(vertcat(pen.dyn_auxvar, pen.cost_auxvar))
poc.setStateVariable(pen.X)
poc.setControlVariable(pen.U)
poc.setDyn(pen.X + dt11 * pen.f)
poc.setPathCost(pen.path_cost)
poc.setFinalCost(pen.final_cost)
print("pendulum IRL script", ocsolver.newton_system(poc._system()).module_path)
engine.DenseLQR.get(2, 1, 5)
