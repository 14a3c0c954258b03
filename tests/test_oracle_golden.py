"""Pin the CPU oracle against the golden vectors the reference ships."""
import os

import numpy as np
import pytest

from oracle import envs, pdp_oracle, ref_loader

G = os.path.join(os.path.dirname(__file__), "golden")

SYSID_SETUP = {  # reference Examples/SysID/*/generate_traj.py
    "pendulum": (envs.pendulum, {}, 0.05), "cartpole": (envs.cartpole, {}, 0.05),
    "robotarm": (envs.robotarm, dict(g=0), 0.1), "quadrotor": (envs.quadrotor, dict(c=0.01), 0.1),
    "rocket": (envs.rocket, {}, 0.2)}
IRL_SETUP = {  # reference Examples/IRL/*/generate_demos.py
    "pendulum": (envs.pendulum, {}), "cartpole": (envs.cartpole, dict(wu=0.1)),
    "robotarm": (envs.robotarm, dict(g=0, wu=0.01)), "quadrotor": (envs.quadrotor, dict(c=0.01, wthrust=0.1)),
    "rocket": (envs.rocket, dict(wthrust=0.1))}


@pytest.mark.parametrize("env", list(SYSID_SETUP))
def test_k1_rollout_matches_shipped_iodata(env):
    g = np.load(os.path.join(G, "k1_iodata.npz"))
    builder, kw, dt = SYSID_SETUP[env]
    e = builder(**kw)
    sid = pdp_oracle.OracleSysID(e["X"], e["U"], e["dyn_params"], e["X"] + dt * e["f"])
    for inp, st in zip(g[env + "_inputs"], g[env + "_states"]):
        X = sid.integrateDyn(st[0], inp, g[env + "_true_parameter"])
        assert np.max(np.abs(X - st)) < 1e-12


@pytest.mark.parametrize("env", list(IRL_SETUP))
def test_k2_demos_satisfy_dynamics_cost_and_pmp(env):
    g = np.load(os.path.join(G, "k2_demos.npz"))
    builder, kw = IRL_SETUP[env]
    oc = pdp_oracle.build_oc(builder(**kw), float(g[env + "_dt"][0]))
    oc.diffPMP()
    theta = g[env + "_true_parameter"]
    for i in range(int(g[env + "_n"])):
        X, U, L = g["%s_%d_X" % (env, i)], g["%s_%d_U" % (env, i)], g["%s_%d_L" % (env, i)]
        Xr, cost = oc.rollout(X[0], U, theta)
        scale = max(1.0, np.max(np.abs(X)))
        assert np.max(np.abs(Xr - X)) < 2e-6 * scale           # IPOPT constraint tolerance
        assert abs(cost - g["%s_%d_cost" % (env, i)][0]) < 1e-6 * max(1.0, abs(cost))
        # costate convention: costate[t] = lambda_{t+1}, PMP recursion and dHu = 0 (IPOPT tol)
        Lr = oc.costate(X, U, theta)
        lscale = max(1.0, np.max(np.abs(L)))
        assert np.max(np.abs(Lr - L)) < 5e-5 * lscale
        dHu = oc.dHu_traj(X, U, L, theta)
        assert np.max(np.abs(dHu)) < 5e-5 * lscale


def test_k4_rocket_oc_adjoint_gradient():
    g = np.load(os.path.join(G, "k4_rocket_oc.npz"))
    e = envs.rocket(Jx=0.5, Jy=1., Jz=1., mass=1., l=1., wr=1, wv=1, wtilt=50, ww=1, wsidethrust=1, wthrust=0.4)
    dt, H = float(g["dt"][0]), int(g["horizon"][0])
    cp = pdp_oracle.OracleCP(e["X"], e["U"], e["X"] + dt * e["f"], e["path_cost"], e["final_cost"])
    x0 = np.array([10, -8, 5., -.1, 0, 0] + envs.to_quaternion(1.5, [0, 0, 1]).tolist() + [0, 0, 0])
    lr = float(g["lr"][0])
    for U, Un, loss in zip(g["U"], g["U_next"], g["loss"]):
        J, grad, _ = cp.adjoint_grad(x0, U.reshape(H, 3))
        assert abs(J - loss) <= 1e-12 * abs(loss)
        gref = (U - Un) / lr
        assert np.max(np.abs(grad.ravel() - gref)) <= 1e-9 * np.max(np.abs(gref)) + 1e-7  # trace stored after fp subtract
    # stored final rollout
    _, _, Xs = cp.adjoint_grad(x0, g["solved_U"])
    assert np.max(np.abs(Xs - g["solved_X"])) < 1e-11


@pytest.mark.parametrize("env", ["cartpole", "robotarm"])
def test_k5_neural_policy_layout(env):
    g = np.load(os.path.join(G, "k5_neural.npz"))
    dt, H = float(g[env + "_dt"][0]), int(g[env + "_horizon"][0])
    if env == "cartpole":
        e = envs.cartpole(mc=0.1, mp=0.1, l=1, wx=0.1, wq=0.6, wdx=0.1, wdq=0.1, wu=0.3)
        x0 = np.zeros(4)
    else:  # reference Examples/OC/robotarm/robotarm_PDP_neural.py
        e = envs.robotarm(l1=1, m1=1, l2=1, m2=1, g=0, wq1=0.1, wq2=0.1, wdq1=0.1, wdq2=0.1, wu=0.01)
        x0 = np.array([-np.pi / 2, 3 * np.pi / 4, 0, 0])
    cp = pdp_oracle.OracleCP(e["X"], e["U"], e["X"] + dt * e["f"], e["path_cost"], e["final_cost"])
    cp.set_neural([cp.n, cp.n])
    assert cp.r == g[env + "_theta"].size
    X, U, cost = cp.integrateSys(g[env + "_X"][0], H, g[env + "_theta"])
    assert np.max(np.abs(X - g[env + "_X"])) < 1e-10
    assert np.max(np.abs(U - g[env + "_U"])) < 1e-10


@pytest.mark.parametrize("env", ["quadrotor", "pendulum"])
def test_k6_lqr_restatement_matches_reference_numpy(env):
    g = np.load(os.path.join(G, "k6_reference_lqr.npz"))
    aux = {k: list(g["%s_%s" % (env, k)]) for k in
           ("dynF", "dynG", "dynE", "Hxx", "Hxu", "Hxe", "Hux", "Huu", "Hue", "hxx", "hxe")}
    H = len(aux["dynF"])
    n, r = aux["dynE"][0].shape
    sol = pdp_oracle.lqr_solve(aux, np.zeros((n, r)), H)
    for key, ref in (("state_traj_opt", g[env + "_dX"]), ("control_traj_opt", g[env + "_dU"]),
                     ("costate_traj_opt", g[env + "_dL"])):
        mine = np.stack(sol[key])
        assert np.max(np.abs(mine - ref)) <= 1e-11 * max(1.0, np.max(np.abs(ref)))


def test_k6_oracle_aux_matrices_regenerate_golden():
    """The aux matrices stored in K6 came from this oracle + shipped demos; re-evaluating must agree."""
    g6 = np.load(os.path.join(G, "k6_reference_lqr.npz"))
    g2 = np.load(os.path.join(G, "k2_demos.npz"))
    oc = pdp_oracle.build_oc(envs.quadrotor(c=0.01, wthrust=0.1), float(g2["quadrotor_dt"][0]))
    aux = oc.getAuxSys(g2["quadrotor_0_X"], g2["quadrotor_0_U"], g2["quadrotor_0_L"], g6["quadrotor_theta"])
    for k in ("dynF", "dynG", "dynE", "Hxx", "Hxu", "Hxe", "Hux", "Huu", "Hue", "hxx", "hxe"):
        assert np.max(np.abs(np.stack(aux[k]) - g6["quadrotor_" + k])) < 1e-12


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree only exists in the build container")
def test_k6_live_reference_lqr_and_sensitivity_recursions():
    ref = ref_loader.load_reference_pdp()
    g = np.load(os.path.join(G, "k6_reference_lqr.npz"))
    F, Gm, Ux, Ue, E = (list(g["fs_" + k]) for k in ("F", "G", "Ux", "Ue", "E"))
    n, r = E[0].shape
    cp = ref.ControlPlanning().integrateAuxSys(F, Gm, Ux, Ue, np.zeros((n, r)))
    assert np.allclose(np.stack(cp["state_traj"]), g["fs_cp_X"], rtol=0, atol=1e-13)
    sid = ref.SysID().integrateAuxSys(F, E, np.zeros((n, r)))
    assert np.allclose(np.stack(sid["state_traj"]), g["fs_sysid_X"], rtol=0, atol=1e-13)


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference PDP.py neither mounted nor staged")
def test_k6_live_reference_lqr_solver_matches_the_restatement_on_a_quadrotor_sweep():
    """The CPU arm's 'reference' kind (bench.py): the unmodified LQR.lqrSolver fed by the oracle's getAuxSys gives the
    restated lqr_solve's trajectories (1e-10 relative)."""
    oc = pdp_oracle.build_oc(envs.quadrotor(c=0.01, wthrust=0.1), 0.1)
    oc.diffPMP()
    rng = np.random.default_rng(5)
    x0 = np.concatenate([rng.uniform(-3, 3, 3), np.zeros(3), [1, 0, 0, 0], np.zeros(3)])
    theta = np.array([1, 1, 1, 1, 0.4, 1, 1, 5, 1.0]) + rng.uniform(-0.2, 0.2, 9)
    U = 2.5 + 0.3 * rng.standard_normal((12, 4))
    X, _ = oc.rollout(x0, U, theta)
    aux = oc.getAuxSys(X, U, oc.costate(X, U, theta), theta)
    ours = pdp_oracle.lqr_solve(aux, np.zeros((13, 9)), 12)
    ref = ref_loader.reference_lqr_solver(aux, np.zeros((13, 9)), 12)
    for k in ("state_traj_opt", "control_traj_opt", "costate_traj_opt"):
        a, b = np.stack(ours[k]), np.stack(ref[k])
        assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b)), k


@pytest.mark.skipif(not os.path.isfile("/root/reference/PDP/PDP.py"), reason="build container only")
def test_staged_reference_copy_is_byte_identical():
    import filecmp
    from oracle import stage_reference
    assert stage_reference.stage(verbose=False)
    assert filecmp.cmp(stage_reference.SRC, stage_reference.DST, shallow=False)


@pytest.mark.parametrize("env,demo", [("pendulum", 0), ("pendulum", 3), ("quadrotor", 0), ("robotarm", 1), ("cartpole", 2)])
def test_k2_oracle_oc_solver_reproduces_shipped_ipopt_demos(env, demo):
    """The oracle's OC solve from the cold start U = 0 lands on the demo IPOPT found (K2)."""
    from oracle import oc_solve
    g = np.load(os.path.join(G, "k2_demos.npz"))
    builder, kw = IRL_SETUP[env]
    oc = pdp_oracle.build_oc(builder(**kw), float(g[env + "_dt"][0]))
    Xd, Ud, Ld = (g["%s_%d_%s" % (env, demo, k)] for k in ("X", "U", "L"))
    X, U, L, cost, it = oc_solve.solve(oc, Xd[0], Ud.shape[0], g[env + "_true_parameter"])
    assert abs(cost - g["%s_%d_cost" % (env, demo)][0]) < 1e-7 * abs(cost)
    assert np.max(np.abs(X - Xd)) < 2e-5 * max(1.0, np.max(np.abs(Xd)))
    assert np.max(np.abs(U - Ud)) < 2e-5 * max(1.0, np.max(np.abs(Ud)))
    assert np.max(np.abs(L - Ld)) < 5e-5 * max(1.0, np.max(np.abs(Ld)))


@pytest.mark.parametrize("env,trial", [("pendulum", 0), ("pendulum", 1), ("quadrotor", 0), ("cartpole", 0), ("robotarm", 0),
                                       ("rocket", 0)])
def test_k3_irl_loss_and_gradient_trace(env, trial):
    """End-to-end hot path: ocSolver -> getAuxSys -> lqrSolver -> chain rule reproduces the shipped traces:
    loss(theta_k) = loss_trace[k+1], dp(theta_k) = (theta_k - theta_{k+1}) / lr."""
    from oracle import oc_solve
    g2 = np.load(os.path.join(G, "k2_demos.npz"))
    g3 = np.load(os.path.join(G, "k3_irl_traces.npz"))
    builder, kw = IRL_SETUP[env]
    oc = pdp_oracle.build_oc(builder(**kw), float(g2[env + "_dt"][0]))
    lr = float(g3["%s_%d_lr" % (env, trial)][0])
    nd = int(g2[env + "_n"])
    for k in range({"pendulum": 2, "cartpole": 2, "robotarm": 2}.get(env, 1)):
        theta = g3["%s_%d_theta" % (env, trial)][k]
        loss, dp = 0.0, np.zeros(oc.r)
        for i in range(nd):
            Xd, Ud = g2["%s_%d_X" % (env, i)], g2["%s_%d_U" % (env, i)]
            X, U, L, cost, it = oc_solve.solve(oc, Xd[0], Ud.shape[0], theta)
            aux = oc.getAuxSys(X, U, L, theta)
            sol = pdp_oracle.lqr_solve(aux, np.zeros((oc.n, oc.r)), Ud.shape[0])
            l_i, dp_i = pdp_oracle.irl_loss_grad(X, U, Xd, Ud, sol["state_traj_opt"], sol["control_traj_opt"])
            loss += l_i
            dp += dp_i
        loss, dp = loss / nd, dp / nd
        dp_ref = (theta - g3["%s_%d_theta_next" % (env, trial)][k]) / lr
        assert abs(loss - g3["%s_%d_loss" % (env, trial)][k]) < 1e-6 * abs(loss)
        assert np.max(np.abs(dp - dp_ref)) < 1e-5 * np.max(np.abs(dp_ref))       # late cart-pole iterates: IPOPT floor
