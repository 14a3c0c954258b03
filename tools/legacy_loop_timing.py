"""Time the quadrotor IRL outer loop (reference Examples/IRL/quadrotor/uav_PDP.py:40-83) on the GPU box:
(a) through the drop-in legacy API exactly as the script calls it (ocSolver -> getAuxSys -> LQR.lqrSolver -> chain
rule in NumPy, B = 1 per demo), (b) through the device-resident IRLTrainer.  The reference's shipped traces record
3974-4384 s per 10 000 iterations (2.3-2.5 it/s) on the author's machine."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from PDP import PDP  # noqa: E402
from JinEnv import JinEnv  # noqa: E402
from casadi import vertcat  # noqa: E402
from pontryagin_differentiable_programming_b200 import irl  # noqa: E402

g2 = np.load(os.path.join(ROOT, "tests", "golden", "k2_demos.npz"))
g3 = np.load(os.path.join(ROOT, "tests", "golden", "k3_irl_traces.npz"))
uav = JinEnv.Quadrotor()
uav.initDyn(c=0.01)
uav.initCost(wthrust=0.1)
oc = PDP.OCSys()
oc.setAuxvarVariable(vertcat(uav.dyn_auxvar, uav.cost_auxvar))
oc.setControlVariable(uav.U)
oc.setStateVariable(uav.X)
oc.setDyn(uav.X + g2["quadrotor_dt"].reshape(1, 1) * uav.f)
oc.setPathCost(uav.path_cost)
oc.setFinalCost(uav.final_cost)
oc.diffPMP()
lqr_solver = PDP.LQR()
demos = [(g2["quadrotor_%d_X" % i], g2["quadrotor_%d_U" % i]) for i in range(2)]
lr = 1e-4
theta0 = g3["quadrotor_0_theta"][0].reshape(1, -1)


def legacy_iteration(current_parameter, n_starts):
    loss, dp = 0, np.zeros(current_parameter.shape)
    for Xd, Ud in demos:
        H = Ud.shape[0]
        traj = oc.ocSolver(Xd[0, :], H, current_parameter, n_starts=n_starts)
        aux_sys = oc.getAuxSys(state_traj_opt=traj['state_traj_opt'], control_traj_opt=traj['control_traj_opt'],
                               costate_traj_opt=traj['costate_traj_opt'], auxvar_value=current_parameter)
        lqr_solver.setDyn(dynF=aux_sys['dynF'], dynG=aux_sys['dynG'], dynE=aux_sys['dynE'])
        lqr_solver.setPathCost(Hxx=aux_sys['Hxx'], Huu=aux_sys['Huu'], Hxu=aux_sys['Hxu'], Hux=aux_sys['Hux'],
                               Hxe=aux_sys['Hxe'], Hue=aux_sys['Hue'])
        lqr_solver.setFinalCost(hxx=aux_sys['hxx'], hxe=aux_sys['hxe'])
        aux_sol = lqr_solver.lqrSolver(np.zeros((oc.n_state, oc.n_auxvar)), H)
        dldx, dldu = traj['state_traj_opt'] - Xd, traj['control_traj_opt'] - Ud
        loss = loss + np.linalg.norm(dldx) ** 2 + np.linalg.norm(dldu) ** 2
        for t in range(H):
            dp = dp + np.matmul(dldx[t, :], aux_sol['state_traj_opt'][t]) + np.matmul(dldu[t, :], aux_sol['control_traj_opt'][t])
        dp = dp + np.dot(dldx[-1, :], aux_sol['state_traj_opt'][-1])
    return loss / 2, dp / 2


rows = []
for n_starts in (1,):
    theta = theta0.copy()
    legacy_iteration(theta, n_starts)                       # compile / warm
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 15
    for _ in range(K):
        loss, dp = legacy_iteration(theta, n_starts)
        theta = theta - lr * dp
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K
    rows.append({"path": "legacy drop-in API (script loop), n_starts=%d" % n_starts, "s_per_iter": dt, "iters_per_s": 1 / dt,
                 "loss_after": float(loss)})
dev = torch.device("cuda:0")
Xd = torch.as_tensor(np.stack([d[0] for d in demos]), device=dev)
Ud = torch.as_tensor(np.stack([d[1] for d in demos]), device=dev)
trainer = irl.IRLTrainer(oc._system(), Xd, Ud, lr)
theta = torch.as_tensor(theta0.ravel(), device=dev)
trainer.step(theta)
torch.cuda.synchronize()
t0 = time.perf_counter()
K = 100
for _ in range(K):
    loss, theta = trainer.step(theta)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / K
rows.append({"path": "IRLTrainer (device-resident, warm-started solver, 2 demos)", "s_per_iter": dt, "iters_per_s": 1 / dt,
             "loss_after": float(loss)})
for n_newton in (3, 10):
    gtr = irl.IRLTrainer(oc._system(), Xd, Ud, lr)
    th_g = torch.as_tensor(theta0.ravel(), device=dev)
    for _ in range(3):                                                 # capture + the large first updates
        th_g = gtr.step_graph(th_g, n_newton=n_newton)[1].clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        out = gtr.step_graph(th_g, n_newton=n_newton)
        th_g.copy_(out[1])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K
    rows.append({"path": "IRLTrainer.step_graph (one CUDA graph per iteration, n_newton=%d)" % n_newton, "s_per_iter": dt,
                 "iters_per_s": 1 / dt, "loss_after": float(out[0]), "resid": float(out[2])})
rows.append({"path": "reference as shipped (author's machine, PDP_results_trial_*.mat time_passed)", "iters_per_s": 2.4})
for r in rows:
    print(json.dumps(r))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "legacy_loop_timing.json"), "w"), indent=1)
