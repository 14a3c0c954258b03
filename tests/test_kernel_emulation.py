"""CPU checks of the GENERATED Riccati kernels through the warp emulator (tools/warp_emu.py, test infrastructure).

The emulator compiles the very CUDA text that nvcc compiles for sm_100a (kernels + device functions) with g++ and
runs one warp as 32 host threads, so the lane layout, shared-memory indexing and synchronisation of both backward
kernels (one / two trajectories per warp) and of the packed forward kernel are checked against the oracle without a
GPU.  The `-m gpu` suite repeats the same comparisons on the real device through the C ABI."""
import os

import numpy as np
import pytest

from oracle import envs, pdp_oracle
from tools import warp_emu

G = os.path.join(os.path.dirname(__file__), "golden")

ORACLE_ENVS = {
    "quadrotor": (envs.quadrotor, dict(c=0.01, wthrust=0.1)), "pendulum": (envs.pendulum, {}),
    "rocket": (envs.rocket, dict(wthrust=0.1)), "cartpole": (envs.cartpole, dict(wu=0.1)),
    "robotarm": (envs.robotarm, dict(g=0, wu=0.01))}


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b)))


def _variant(base, **kw):
    from pontryagin_differentiable_programming_b200 import codegen
    cfg = dict(chunk=base.chunk, warps_per_block=base.wpb, min_blocks=base.min_blocks, fwd_warps_per_block=base.wpbf,
               fwd_min_blocks=base.min_blocks_f, keep_fg=base.keep_fg, fast_rcp=base.fast_rcp,
               early_solve=base.early_solve, fwd_pack=base.fwd_pack, fwd_chunk=base.fwd_chunk, bwd_pack=base.bwd_pack)
    cfg.update(fwd_vec=base.fwd_vec, prefetch=base.prefetch, prefetch_dist=base.prefetch_dist, h_group=base.h_group,
               rollout_tma=base.rollout_tma, tma_chunk=base.tma_chunk)
    cfg.update(kw)
    return codegen.OCModuleSource(base.x, base.u, base.th, base.dyn, base.c, base.h, **cfg)


@pytest.mark.parametrize("env", ["quadrotor", "pendulum", "rocket", "cartpole", "robotarm"])
def test_emulated_kernels_match_oracle_on_the_shipped_demos(env):
    """Same data as tests/test_gpu_oc.py::test_demo_trajectories_backward_sweep_matches_oracle: the shipped IPOPT
    demos (X, U, Lam) with a perturbed theta; both backward-kernel layouts must reproduce the oracle's literal
    reference-form lqrSolver (PDP.py:446-615) to 1e-9 and agree with each other to round-off."""
    from pontryagin_differentiable_programming_b200 import systems
    g2 = np.load(os.path.join(G, "k2_demos.npz"))
    dt = float(g2[env + "_dt"][0])
    base = systems.OC_BUILDERS[env](dt).src
    builder, kw = ORACLE_ENVS[env]
    oc = pdp_oracle.build_oc(builder(**kw), dt)
    nd = int(g2[env + "_n"])
    theta = g2[env + "_true_parameter"] * 1.1
    X = np.stack([g2["%s_%d_X" % (env, i)] for i in range(nd)])
    U = np.stack([g2["%s_%d_U" % (env, i)] for i in range(nd)])
    L = np.stack([g2["%s_%d_L" % (env, i)] for i in range(nd)])
    if nd % 2 == 0:                      # an odd batch exercises the shadowing tail half-warp
        X, U, L = (np.concatenate([a, a[:1]]) for a in (X, U, L))
    assert base.bwd_pack == 2            # all five shipped systems fit the two-trajectory layout
    gains = {}
    for pack in (2, 1):
        src = base if pack == 2 else _variant(base, bwd_pack=1, chunk=5, warps_per_block=2, min_blocks=1, keep_fg=True)
        emu = warp_emu.Emulator(src)
        gains[pack], status = emu.backward(X, U, L, theta)
        assert not np.isnan(gains[pack]).any() and int(status.max()) == 0
        dX, dU, _, st = emu.forward(X, U, theta, gains[pack])
        assert int(st.max()) == 0
        for i in range(X.shape[0]):
            aux = oc.getAuxSys(X[i], U[i], L[i], theta)
            sol = pdp_oracle.lqr_solve(aux, np.zeros((oc.n, oc.r)), U.shape[1])
            assert _rel(dX[i], np.stack(sol["state_traj_opt"])) < 1e-9
            assert _rel(dU[i], np.stack(sol["control_traj_opt"])) < 1e-9
    assert _rel(gains[1], gains[2]) < 1e-11       # same algebra, row-for-row; only the operand order may differ


def test_emulated_two_trajectory_kernel_chunk_tails_and_per_trajectory_theta():
    """Horizon not a multiple of the chunk, chunk of one step, per-trajectory theta, fused IRL loss / chain rule."""
    from pontryagin_differentiable_programming_b200 import systems
    base = systems.quadrotor_irl(0.1).src
    oc = pdp_oracle.build_oc(envs.quadrotor(c=0.01, wthrust=0.1), 0.1)
    rng = np.random.default_rng(3)
    B, H = 3, 11
    x0 = np.tile(np.array([-8, -6, 9., 0, 0, 0, 1, 0, 0, 0, 0, 0, 0]), (B, 1)) + 0.05 * rng.standard_normal((B, 13))
    theta = np.array([1, 1, 1, 1, 0.4, 1, 1, 5, 1.]) * (1 + 0.1 * rng.uniform(-1, 1, (B, 9)))
    U = 2.5 + 0.5 * rng.standard_normal((B, H, 4))
    ref = [pdp_oracle.pdp_sweep(oc, x0[b], U[b], theta[b]) for b in range(B)]
    X = np.stack([r[0] for r in ref])
    L = np.stack([r[1] for r in ref])
    Xd = X + 0.1 * rng.standard_normal(X.shape)
    Ud = U + 0.1 * rng.standard_normal(U.shape)
    for kw in (dict(chunk=4), dict(chunk=1, early_solve=False, fast_rcp=False), dict(chunk=16, warps_per_block=2)):
        emu = warp_emu.Emulator(_variant(base, **kw))
        gains, _ = emu.backward(X, U, L, theta)
        dX, dU, ldp, _ = emu.forward(X, U, theta, gains, Xref=Xd, Uref=Ud)
        for b in range(B):
            assert _rel(dX[b], np.asarray(ref[b][3])) < 1e-10
            assert _rel(dU[b], np.asarray(ref[b][4])) < 1e-10
            loss, dp = pdp_oracle.irl_loss_grad(X[b], U[b], Xd[b], Ud[b], ref[b][3], ref[b][4])
            assert abs(ldp[b, 0] - loss) < 1e-12 * abs(loss)
            assert _rel(ldp[b, 1:], np.asarray(dp).ravel()) < 1e-10


@pytest.mark.parametrize("env,B,H", [("quadrotor", 37, 9), ("quadrotor", 5, 1), ("pendulum", 33, 21), ("cartpole", 3, 8),
                                     ("rocket", 2, 6)])
def test_emulated_rollout_costate_kernel_matches_oracle(env, B, H):
    """pdp_k_rollout_costate (thread per trajectory; two-stage prefetched row loads with both address parities: rows of
    n = 13 doubles alternate between 16-byte aligned and misaligned) vs the oracle's rollout / PMP costate recursion /
    dH/du (PDP.py:158-175, 203-209): odd and even horizons, one-step horizon, batches beyond one block."""
    from pontryagin_differentiable_programming_b200 import systems
    src = systems.OC_BUILDERS[env](0.1).src
    builder, kw = ORACLE_ENVS[env]
    oc = pdp_oracle.build_oc(builder(**kw), 0.1)
    rng = np.random.default_rng(2)
    x0 = 0.3 * rng.standard_normal((B, src.n))
    if env in ("quadrotor", "rocket"):
        x0[:, 6] += 1.0
    theta = 1.0 + 0.2 * rng.uniform(-1, 1, (B, src.r))
    U = 0.5 * rng.standard_normal((B, H, src.m)) + (2.5 if env == "quadrotor" else 0.0)
    emu = warp_emu.Emulator(src)
    X, L, cost, dHu = emu.rollout(x0, theta, U, want_dHu=True)
    assert not np.isnan(X).any() and not np.isnan(L).any() and not np.isnan(dHu).any()
    for b in sorted({0, 1, B // 2, B - 1}):
        Xr, c = oc.rollout(x0[b], U[b], theta[b])
        Lr = oc.costate(Xr, U[b], theta[b])
        assert np.max(np.abs(X[b] - Xr)) < 1e-12
        assert _rel(L[b], Lr) < 1e-12
        assert abs(cost[b] - float(c)) < 1e-11 * max(1.0, abs(float(c)))
        assert _rel(dHu[b], oc.dHu_traj(Xr, U[b], Lr, theta[b])) < 1e-11
    X2, L2, cost2, _ = emu.rollout(x0, theta, U)                      # without dH/du: same X / Lam / cost
    assert np.array_equal(X, X2) and np.array_equal(L, L2) and np.array_equal(cost, cost2)


def test_emulated_closed_loop_rollout_with_candidate_groups():
    """Closed-loop mode of the rollout kernel (the batched ocSolver's line search): u_t = U[t] + alpha k_t + K_t (x_t - Xref[t])
    with `group` candidates per source trajectory in one launch, vs a NumPy restatement on the oracle's dynamics."""
    from pontryagin_differentiable_programming_b200 import systems
    src = systems.OC_BUILDERS["pendulum"](0.1).src
    oc = pdp_oracle.build_oc(envs.pendulum(), 0.1)
    rng = np.random.default_rng(3)
    Bs, H, group = 13, 11, 3
    n, m = src.n, src.m
    x0 = 0.3 * rng.standard_normal((Bs, n))
    theta = 1.0 + 0.2 * rng.uniform(-1, 1, (Bs, src.r))
    U = 0.5 * rng.standard_normal((Bs, H, m))
    Xref = np.stack([oc.rollout(x0[b], U[b], theta[b])[0] for b in range(Bs)])
    gains = 0.1 * rng.standard_normal((Bs, H, (n + 1) * m))
    alpha = rng.uniform(0.1, 1.0, Bs * group)
    emu = warp_emu.Emulator(src)
    X, L, cost, dHu, Uout = emu.rollout(x0, theta, U, want_dHu=True, feedback=dict(gains=gains, X=Xref, alpha=alpha, group=group))
    for b in (0, 1, 17, Bs * group - 1):
        bs = b // group
        x = x0[bs].copy()
        Ua = np.zeros((H, m))
        for t in range(H):
            K = gains[bs, t, :n * m].reshape(n, m)               # record layout: rows 0..n-1 = columns of K, then k
            k = gains[bs, t, n * m:]
            Ua[t] = U[bs, t] + alpha[b] * k + (x - Xref[bs, t]) @ K
            assert np.max(np.abs(X[b, t] - x)) < 1e-12
            x = np.asarray(oc.dyn_fn(x, Ua[t], theta[bs]), dtype=np.float64).reshape(n)
        assert np.max(np.abs(Uout[b] - Ua)) < 1e-12
        Xr, c = oc.rollout(x0[bs], Ua, theta[bs])
        assert abs(cost[b] - float(c)) < 1e-11 * max(1.0, abs(float(c)))
        assert _rel(L[b], oc.costate(Xr, Ua, theta[bs])) < 1e-11


@pytest.mark.parametrize("env", ["quadrotor", "pendulum"])
def test_emulated_newton_module_layouts_agree(env):
    """The one-column "Newton module" of the batched ocSolver (Hue := dH/du, E = Hxe = 0, theta extended by the
    Hessian switch s and the Levenberg shift mu) through both backward-kernel layouts: same gains, same (dx, du)."""
    from pontryagin_differentiable_programming_b200 import codegen, systems
    s = systems.OC_BUILDERS[env](0.1).src
    rng = np.random.default_rng(1)
    B, H = 3, 13
    res = {}
    for pack, kw in ((2, dict(chunk=8, warps_per_block=1, min_blocks=8, keep_fg=True)),
                     (1, dict(chunk=5, warps_per_block=2, min_blocks=1, keep_fg=True))):
        src = codegen.NewtonModuleSource(s.x, s.u, s.th, s.dyn, s.c, s.h, fast_rcp=True, early_solve=True, bwd_pack=pack, **kw)
        assert src.bwd_pack == pack and src.r == 1 and src.nth == s.r + 2
        if pack == 2:
            X = rng.standard_normal((B, H + 1, src.n)) * 0.3
            if env == "quadrotor":
                X[:, :, 6] += 1.0
            U = rng.standard_normal((B, H, src.m)) * 0.3 + (2.5 if env == "quadrotor" else 0.0)
            L = rng.standard_normal((B, H, src.n)) * 0.1
            th = np.concatenate([np.tile(np.abs(rng.standard_normal(src.nth - 2)) + 0.5, (B, 1)), np.zeros((B, 1)),
                                 np.full((B, 1), 0.5)], axis=1)         # Gauss-Newton Hessians, mu = 0.5
        emu = warp_emu.Emulator(src)
        g, st = emu.backward(X, U, L, th)
        dX, dU, _, _ = emu.forward(X, U, th, g)
        assert int(st.max()) == 0 and not np.isnan(g).any() and np.isfinite(dU).all()
        res[pack] = (g, dX, dU)
    for a, b in zip(res[1], res[2]):
        assert _rel(a, b) < 1e-11


@pytest.mark.parametrize("env,dims", [("quadrotor", (13, 4, 9)), ("pendulum", (2, 1, 5))])
def test_emulated_dense_lqr_kernels_reproduce_the_reference_lqrsolver(env, dims):
    """K6 on the CPU: the generic dense LQR module (the kernels behind the drop-in ``LQR.lqrSolver``) fed with the
    auxiliary matrices of the golden file must reproduce the output of the reference's OWN ``LQR.lqrSolver``
    (PDP.py:446-615, run unmodified by tests/golden/make_golden.py) -- literal inv(I+PR) form there, stacked standard
    form with LDL^T here."""
    from pontryagin_differentiable_programming_b200 import codegen
    n, m, r = dims
    g = np.load(os.path.join(G, "k6_reference_lqr.npz"))
    H = g[env + "_dynF"].shape[0]
    aux = np.concatenate([g[env + "_" + k].reshape(H, -1)
                          for k in ("dynF", "dynG", "dynE", "Hxx", "Hxu", "Hxe", "Hux", "Huu", "Hue")], axis=1)[None]
    term = np.concatenate([g[env + "_hxx"].reshape(-1), g[env + "_hxe"].reshape(-1)])[None]
    emu = warp_emu.Emulator(codegen.LQRModuleSource(n, m, r))
    gains, st = emu.backward_dense(aux, term)
    dX, dU, st2 = emu.forward_dense(aux, gains)
    assert int(st.max()) == 0 and int(st2.max()) == 0
    assert _rel(dX[0], g[env + "_dX"]) < 1e-11
    assert _rel(dU[0], g[env + "_dU"]) < 1e-11


def test_emulated_sysid_kernel_matches_k1_golden_and_oracle():
    """pdp_k_sens_fwd (SysID kind) on the CPU: the rollout at the true parameter reproduces the shipped
    uav_iodata.mat states (K1), and off the optimum loss / half-gradient / sensitivities equal the oracle's
    restatement of SysID.step (PDP.py:1261-1296)."""
    from pontryagin_differentiable_programming_b200 import systems
    g = np.load(os.path.join(G, "k1_iodata.npz"))
    sys_ = systems.quadrotor_sysid(0.1)
    inputs, states = g["quadrotor_inputs"][:3], g["quadrotor_states"][:3]
    theta_true = g["quadrotor_true_parameter"]
    emu = warp_emu.SensEmulator(sys_.src)
    H = inputs.shape[1]
    out = emu.run(states[:, 0], theta_true, H, inputs=inputs, Xobs=states)
    assert np.max(np.abs(out["X"] - states)) < 1e-12 and np.max(np.abs(out["loss_dp"][:, 0])) < 1e-20
    theta = theta_true + np.array([0.1, -0.2, 0.15, 0.2, -0.05])
    e = envs.quadrotor(c=0.01)
    sid = pdp_oracle.OracleSysID(e["X"], e["U"], e["dyn_params"], e["X"] + 0.1 * e["f"])
    out = emu.run(states[:, 0], theta, H, inputs=inputs, Xobs=states)
    loss_ref, dp_ref = sid.step(list(inputs), list(states), theta)
    assert abs(out["loss_dp"][:, 0].mean() - loss_ref) < 1e-12 * loss_ref
    assert _rel(out["loss_dp"][:, 1:].mean(axis=0), dp_ref) < 1e-11
    for b in range(inputs.shape[0]):
        X = sid.integrateDyn(states[b, 0], inputs[b], theta)
        assert _rel(out["X"][b], X) < 1e-13
        assert _rel(out["dX"][b], np.stack(sid.sens(X, inputs[b], theta))) < 1e-12
    fused = emu.run(states[:, 0], theta, H, inputs=inputs, Xobs=states, fused_only=True)     # pdp_k_sens_fwd (no outputs)
    assert np.array_equal(fused["loss_dp"], out["loss_dp"])


@pytest.mark.parametrize("groups", [1, 2, 12])
def test_emulated_sysid_kernel_ragged_batch_and_column_groups(groups):
    """70 trajectories (a second block with six live threads), H = 11, column groups of 1 / 2 / all columns (grid
    dimension x): every output vs the oracle, the fused-only call identical, rollout-only mode without observations."""
    from JinEnv import JinEnv
    from pontryagin_differentiable_programming_b200 import codegen_sens
    env = JinEnv.Quadrotor()
    env.initDyn(c=0.01)
    src = codegen_sens.SensModuleSource(codegen_sens.KIND_SYSID, env.X, env.U, env.dyn_auxvar, env.X + 0.1 * env.f,
                                        max_group_cols=groups)
    assert len(src.groups) == {1: 5, 2: 3, 12: 1}[groups]
    rng = np.random.default_rng(4)
    B, H = 70, 11
    inputs = rng.uniform(-3, 3, (B, H, 4))
    x0 = np.tile(np.array([-8, -6, 9., 0, 0, 0, 1, 0, 0, 0, 0, 0, 0]), (B, 1)) + 0.05 * rng.standard_normal((B, 13))
    th_true = np.array([1, 1, 1, 1, 0.4])
    theta = th_true + np.array([0.1, -0.2, 0.15, 0.2, -0.05])
    e = envs.quadrotor(c=0.01)
    sid = pdp_oracle.OracleSysID(e["X"], e["U"], e["dyn_params"], e["X"] + 0.1 * e["f"])
    Xobs = np.stack([sid.integrateDyn(x0[b], inputs[b], th_true) for b in range(B)])
    emu = warp_emu.SensEmulator(src)
    out = emu.run(x0, theta, H, inputs=inputs, Xobs=Xobs)
    assert not np.isnan(out["X"]).any() and not np.isnan(out["dX"]).any()
    for b in (0, 1, 63, 64, 69):
        X = sid.integrateDyn(x0[b], inputs[b], theta)
        S = np.stack(sid.sens(X, inputs[b], theta))
        loss, dp = sid.step([inputs[b]], [Xobs[b]], theta)
        assert _rel(out["X"][b], X) < 1e-13 and _rel(out["dX"][b], S) < 1e-12
        assert abs(out["loss_dp"][b, 0] - loss) < 1e-12 * loss and _rel(out["loss_dp"][b, 1:], dp) < 1e-11
    fused = emu.run(x0, theta, H, inputs=inputs, Xobs=Xobs, fused_only=True)
    assert np.array_equal(fused["loss_dp"], out["loss_dp"])
    roll = emu.run(x0, th_true, H, inputs=inputs)                     # rollout only (no observations): X = Xobs
    assert np.max(np.abs(roll["X"] - Xobs)) < 1e-13


@pytest.mark.parametrize("policy", ["poly", "neural"])
def test_emulated_controlplanning_kernel_matches_oracle(policy):
    """pdp_k_sens_fwd (ControlPlanning kind) on the CPU vs the oracle's ControlPlanning.step (PDP.py:850-878): policy
    rollout, forward sensitivities of states and controls, cost and its gradient (r = 6 polynomial / r = 45 MLP)."""
    from pontryagin_differentiable_programming_b200 import systems
    H, dt = 20, 0.05
    sys_ = systems.cartpole_cp(policy, H, dt)
    e = envs.cartpole(mc=0.1, mp=0.1, l=1, wx=0.1, wq=0.6, wdx=0.1, wdq=0.1, wu=0.3)
    cp = pdp_oracle.OracleCP(e["X"], e["U"], e["X"] + dt * e["f"], e["path_cost"], e["final_cost"])
    if policy == "poly":
        cp.set_poly(np.linspace(0, H, 6))
    else:
        cp.set_neural([4, 4])
    assert cp.r == sys_.r
    rng = np.random.default_rng(5)
    B = 3
    x0 = 0.1 * rng.standard_normal((B, 4))
    theta = (1.0 if policy == "poly" else 0.5) * rng.standard_normal((B, cp.r))
    out = warp_emu.SensEmulator(sys_.src).run(x0, theta, H, policy=True)
    for b in range(B):
        cost, gref, X, U, dX, dU = cp.step(x0[b], H, theta[b], return_traj=True)
        assert _rel(out["X"][b], X) < 1e-11 and _rel(out["U"][b], U) < 1e-11
        assert abs(out["loss_dp"][b, 0] - cost) < 1e-11 * abs(cost)
        assert _rel(out["dX"][b], dX) < 1e-10 and _rel(out["dU"][b], dU) < 1e-10
        assert _rel(out["loss_dp"][b, 1:], gref) < 1e-10


def test_emulator_detects_a_missing_barrier():
    """Negative control of the emulator itself: the lanes are real host threads, so a kernel whose Z^T staging
    barrier is removed must give wrong gains (it does, by orders of magnitude and differently on every run) -- the
    value comparisons of this file therefore also guard the shared-memory choreography, not only the arithmetic."""
    from pontryagin_differentiable_programming_b200 import codegen, systems
    base = systems.quadrotor_irl(0.1).src

    class Broken(codegen.OCModuleSource):
        def source(self):
            t = super().source()
            i = t.index("// B: transpose through shared memory")
            j = t.index("__syncwarp();", i)
            return t[:j] + "/* barrier removed */" + t[j + len("__syncwarp();"):]

    broken = Broken(base.x, base.u, base.th, base.dyn, base.c, base.h, chunk=base.chunk, warps_per_block=base.wpb,
                    min_blocks=base.min_blocks, keep_fg=base.keep_fg, fast_rcp=base.fast_rcp, early_solve=base.early_solve,
                    bwd_pack=base.bwd_pack)
    rng = np.random.default_rng(0)
    B, H = 4, 12
    X = 0.3 * rng.standard_normal((B, H + 1, 13))
    X[:, :, 6] += 1.0
    U = 2.5 + 0.3 * rng.standard_normal((B, H, 4))
    L = 0.1 * rng.standard_normal((B, H, 13))
    th = np.array([1, 1, 1, 1, 0.4, 1, 1, 5, 1.]) * np.ones((B, 9))
    good, _ = warp_emu.Emulator(base).backward(X, U, L, th)
    again, _ = warp_emu.Emulator(base).backward(X, U, L, th)
    assert np.array_equal(good, again)                       # the intact kernel is deterministic
    bad, _ = warp_emu.Emulator(broken).backward(X, U, L, th)
    assert not np.allclose(np.nan_to_num(bad), good, rtol=1e-6, atol=0)


@pytest.mark.parametrize("env,B,H,chunk", [("quadrotor", 37, 9, 3), ("quadrotor", 3, 1, 3), ("rocket", 5, 7, 4), ("pendulum", 34, 21, 2),
                                           ("cartpole", 3, 8, 8)])
def test_emulated_tma_rollout_kernel_equals_the_register_prefetch_kernel(env, B, H, chunk):
    """pdp_k_rollout_costate_tma (the open-loop kernel for small batches: rows moved by thread-private 1-D bulk copies; in the emulator a bulk
    copy is an immediate memcpy, so this checks slot layout, the 0 / 8-byte parity shifts of loads and stores -- rows of
    13 / 3 / 1 doubles --, the head / tail elements of the stores, chunk tails and the end-of-tensor guard): bit-identical
    to the register-prefetch kernel, with and without dH/du."""
    from pontryagin_differentiable_programming_b200 import systems
    base = systems.OC_BUILDERS[env](0.1).src
    src = _variant(base, rollout_tma=1, tma_chunk=chunk)
    assert "pdp_k_rollout_costate_tma" in src.source() and "#define PDP_TC %d\n" % chunk in src.source()
    assert "pdp_k_rollout_costate_tma" not in _variant(base, rollout_tma=0).source()
    rng = np.random.default_rng(6)
    x0 = 0.3 * rng.standard_normal((B, src.n))
    if env in ("quadrotor", "rocket"):
        x0[:, 6] += 1.0
    theta = 1.0 + 0.2 * rng.uniform(-1, 1, (B, src.r))
    U = 0.5 * rng.standard_normal((B, H, src.m)) + (2.5 if env == "quadrotor" else 0.0)
    emu = warp_emu.Emulator(src)
    for want in (True, False):
        ref = emu.rollout(x0, theta, U, want_dHu=want)
        got = emu.rollout(x0, theta, U, want_dHu=want, tma=True)
        for a, b_ in zip(ref, got):
            assert (a is None and b_ is None) or np.array_equal(a, b_)
