"""GPU parity of the SysID / ControlPlanning / generic-LQR / adjoint paths against the oracle and goldens."""
import os

import numpy as np
import pytest
import torch

from oracle import envs, pdp_oracle

G = os.path.join(os.path.dirname(__file__), "golden")
pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _t(a, dev):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b)))


def test_sysid_quadrotor_matches_oracle_and_k1_golden():
    from pontryagin_differentiable_programming_b200 import systems
    dev = _dev()
    g = np.load(os.path.join(G, "k1_iodata.npz"))
    sys_ = systems.quadrotor_sysid(0.1)
    inputs, states = g["quadrotor_inputs"], g["quadrotor_states"]
    theta_true = g["quadrotor_true_parameter"]
    # K1: rollout at the true parameter reproduces the shipped states (loss ~ 0)
    out = sys_.step(_t(inputs, dev), _t(states, dev), _t(theta_true, dev), want_traj=True, want_sens=True)
    assert np.max(np.abs(out["X"].cpu().numpy() - states)) < 1e-12
    # off the optimum: loss / half-gradient / sensitivities vs the oracle restatement of SysID.step
    theta = theta_true + np.array([0.1, -0.2, 0.15, 0.2, -0.05])
    e = envs.quadrotor(c=0.01)
    sid = pdp_oracle.OracleSysID(e["X"], e["U"], e["dyn_params"], e["X"] + 0.1 * e["f"])
    out = sys_.step(_t(inputs, dev), _t(states, dev), _t(theta, dev), want_traj=True, want_sens=True)
    ldp = out["loss_dp"].cpu().numpy()
    loss_ref, dp_ref = sid.step(list(inputs), list(states), theta)
    assert abs(ldp[:, 0].mean() - loss_ref) < 1e-12 * loss_ref
    assert _rel(ldp[:, 1:].mean(axis=0), dp_ref) < 1e-11
    for b in range(inputs.shape[0]):
        X = sid.integrateDyn(states[b, 0], inputs[b], theta)
        S = np.stack(sid.sens(X, inputs[b], theta))
        assert _rel(out["X"][b].cpu().numpy(), X) < 1e-13
        assert _rel(out["dX"][b].cpu().numpy(), S) < 1e-12


@pytest.mark.parametrize("policy", ["poly", "neural"])
def test_controlplanning_cartpole_step_matches_oracle(policy):
    from pontryagin_differentiable_programming_b200 import systems
    dev = _dev()
    H, dt = 50, 0.05
    sys_ = systems.cartpole_cp(policy, H, dt)
    e = envs.cartpole(mc=0.1, mp=0.1, l=1, wx=0.1, wq=0.6, wdx=0.1, wdq=0.1, wu=0.3)
    cp = pdp_oracle.OracleCP(e["X"], e["U"], e["X"] + dt * e["f"], e["path_cost"], e["final_cost"])
    if policy == "poly":
        cp.set_poly(np.linspace(0, H, 6))
    else:
        cp.set_neural([4, 4])
    assert cp.r == sys_.r
    rng = np.random.default_rng(5)
    B = 5
    x0 = 0.1 * rng.standard_normal((B, 4))
    theta = (1.0 if policy == "poly" else 0.5) * rng.standard_normal((B, cp.r))
    out = sys_.step(_t(x0, dev), H, _t(theta, dev), want_traj=True, want_sens=True)
    ldp = out["loss_dp"].cpu().numpy()
    for b in range(B):
        cost, g, X, U, dX, dU = cp.step(x0[b], H, theta[b], return_traj=True)
        assert _rel(out["X"][b].cpu().numpy(), X) < 1e-11
        assert _rel(out["U"][b].cpu().numpy(), U) < 1e-11
        assert abs(ldp[b, 0] - cost) < 1e-11 * abs(cost)
        assert _rel(out["dX"][b].cpu().numpy(), dX) < 1e-10
        assert _rel(out["dU"][b].cpu().numpy(), dU) < 1e-10
        assert _rel(ldp[b, 1:], g) < 1e-10


def test_k5_neural_policy_rollout_golden():
    """Shipped final neural policy of the reference reproduces its stored rollout (column-major packing)."""
    from pontryagin_differentiable_programming_b200 import systems
    dev = _dev()
    g = np.load(os.path.join(G, "k5_neural.npz"))
    H, dt = int(g["cartpole_horizon"][0]), float(g["cartpole_dt"][0])
    sys_ = systems.cartpole_cp("neural", H, dt)
    out = sys_.step(_t(g["cartpole_X"][:1], dev), H, _t(g["cartpole_theta"], dev), want_traj=True)
    assert np.max(np.abs(out["X"][0].cpu().numpy() - g["cartpole_X"])) < 1e-10
    assert np.max(np.abs(out["U"][0].cpu().numpy() - g["cartpole_U"])) < 1e-10


def test_k4_rocket_adjoint_gradient_golden():
    """recmat semantics through the drop-in class: J(U_k) and dJ/dU(U_k) of the shipped rocket OC trace."""
    from PDP import PDP
    from JinEnv import JinEnv
    _dev()
    g = np.load(os.path.join(G, "k4_rocket_oc.npz"))
    rocket = JinEnv.Rocket()
    rocket.initDyn(Jx=0.5, Jy=1., Jz=1., mass=1., l=1.)
    rocket.initCost(wr=1, wv=1, wtilt=50, ww=1, wsidethrust=1, wthrust=0.4)
    dt, H = float(g["dt"][0]), int(g["horizon"][0])
    oc = PDP.ControlPlanning()
    oc.setStateVariable(rocket.X)
    oc.setControlVariable(rocket.U)
    oc.setDyn(rocket.X + dt * rocket.f)
    oc.setPathCost(rocket.path_cost)
    oc.setFinalCost(rocket.final_cost)
    oc.recmat_init_step(H, -1)
    assert oc.n_auxvar == 3 * H
    x0 = [10, -8, 5., -.1, 0, 0] + JinEnv.toQuaternion(1.5, [0, 0, 1]) + [0, 0, 0]
    lr = float(g["lr"][0])
    for U, Un, loss in zip(g["U"], g["U_next"], g["loss"]):
        J, grad = oc.recmat_step(x0, H, U)
        assert abs(J - loss) <= 1e-12 * abs(loss)
        gref = (U - Un) / lr
        assert np.max(np.abs(grad - gref)) <= 1e-9 * np.max(np.abs(gref)) + 1e-7
    sol = oc.recmat_unwarp(x0, H, g["solved_U"].reshape(-1))
    assert np.max(np.abs(sol["state_traj"] - g["solved_X"])) < 1e-11


@pytest.mark.parametrize("env", ["quadrotor", "pendulum"])
def test_k6_dropin_lqr_matches_reference_numpy(env):
    """PDP.LQR (generic dense module) on the exact matrices the reference's own lqrSolver was run on."""
    from PDP import PDP
    _dev()
    g = np.load(os.path.join(G, "k6_reference_lqr.npz"))
    aux = {k: list(g["%s_%s" % (env, k)]) for k in
           ("dynF", "dynG", "dynE", "Hxx", "Hxu", "Hxe", "Hux", "Huu", "Hue", "hxx", "hxe")}
    H = len(aux["dynF"])
    n, r = aux["dynE"][0].shape
    lqr = PDP.LQR()
    lqr.setDyn(dynF=aux["dynF"], dynG=aux["dynG"], dynE=aux["dynE"])
    lqr.setPathCost(Hxx=aux["Hxx"], Huu=aux["Huu"], Hxu=aux["Hxu"], Hux=aux["Hux"], Hxe=aux["Hxe"], Hue=aux["Hue"])
    lqr.setFinalCost(hxx=aux["hxx"], hxe=aux["hxe"])
    sol = lqr.lqrSolver(np.zeros((n, r)), H)
    assert _rel(np.stack(sol["state_traj_opt"]), g[env + "_dX"]) < 1e-9
    assert _rel(np.stack(sol["control_traj_opt"]), g[env + "_dU"]) < 1e-9
    assert _rel(np.stack(sol["costate_traj_opt"]), g[env + "_dL"]) < 1e-8


def test_k6_forward_recursions_match_reference_numpy():
    from PDP import PDP
    _dev()
    g = np.load(os.path.join(G, "k6_reference_lqr.npz"))
    F, Gm, Ux, Ue, E = (list(g["fs_" + k]) for k in ("F", "G", "Ux", "Ue", "E"))
    n, r = E[0].shape
    cp = PDP.ControlPlanning().integrateAuxSys(F, Gm, Ux, Ue, np.zeros((n, r)))
    assert np.allclose(np.stack(cp["state_traj"]), g["fs_cp_X"], rtol=0, atol=1e-12)
    assert np.allclose(np.stack(cp["control_traj"]), g["fs_cp_U"], rtol=0, atol=1e-12)
    sid = PDP.SysID().integrateAuxSys(F, E, np.zeros((n, r)))
    assert np.allclose(np.stack(sid["state_traj"]), g["fs_sysid_X"], rtol=0, atol=1e-12)


def test_dropin_getauxsys_and_legacy_calls():
    """OCSys.getAuxSys -> LQR.lqrSolver through the legacy list-of-ndarray API equals the fused kernel."""
    from PDP import PDP
    from JinEnv import JinEnv
    from casadi import vertcat
    dev = _dev()
    g2 = np.load(os.path.join(G, "k2_demos.npz"))
    g6 = np.load(os.path.join(G, "k6_reference_lqr.npz"))
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
    X, U, L = g2["quadrotor_0_X"], g2["quadrotor_0_U"], g2["quadrotor_0_L"]
    theta = g6["quadrotor_theta"].reshape(1, -1)          # (1, r) like the reference scripts pass it
    aux = oc.getAuxSys(state_traj_opt=X, control_traj_opt=U, costate_traj_opt=L, auxvar_value=theta)
    for k in ("dynF", "dynG", "dynE", "Hxx", "Hxu", "Hxe", "Hux", "Huu", "Hue", "hxx", "hxe"):
        assert np.max(np.abs(np.stack(aux[k]) - g6["quadrotor_" + k])) < 1e-11
    lqr = PDP.LQR()
    lqr.setDyn(dynF=aux["dynF"], dynG=aux["dynG"], dynE=aux["dynE"])
    lqr.setPathCost(Hxx=aux["Hxx"], Huu=aux["Huu"], Hxu=aux["Hxu"], Hux=aux["Hux"], Hxe=aux["Hxe"], Hue=aux["Hue"])
    lqr.setFinalCost(hxx=aux["hxx"], hxe=aux["hxe"])
    sol = lqr.lqrSolver(np.zeros((13, 9)), 50)
    assert _rel(np.stack(sol["state_traj_opt"]), g6["quadrotor_dX"]) < 1e-9
