"""BASELINE-size checks through size-independent properties (the oracle only finishes small cases in seconds):
batch independence, fused-vs-materialised chain rule, linearity of the auxiliary system in its initial condition,
determinism, ragged horizons / batch sizes, and sharded == unsharded."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _inputs(B, H, dev, seed=0):
    import bench
    return [torch.as_tensor(np.ascontiguousarray(a), device=dev) for a in bench.synth_quadrotor(B, H, seed=seed)]


def test_c3_full_size_batch_independence_and_chain_rule():
    """B = 16384, H = 50 (config C3): (a) trajectory b's result does not depend on the rest of the batch (bitwise),
    (b) the fused (loss, dp) equals the contraction of the materialised dX, dU with the residuals, (c) two
    launches are bitwise identical."""
    from pontryagin_differentiable_programming_b200 import systems
    dev = _dev()
    sys_ = systems.quadrotor_irl(0.1)
    B, H = 16384, 50
    x0, th, U, Xr, Ur = _inputs(B, H, dev)
    full = sys_.sweep(x0, th, U, Xref=Xr, Uref=Ur)
    again = sys_.sweep(x0, th, U, Xref=Xr, Uref=Ur)
    for k in ("X", "Lam", "dX", "dU", "loss_dp"):
        assert torch.equal(full[k], again[k]), k
    idx = torch.tensor([0, 1, 31, 32, 33, 4095, 8191, 16383], device=dev)
    sub = sys_.sweep(x0[idx].contiguous(), th[idx].contiguous(), U[idx].contiguous(), Xref=Xr[idx].contiguous(), Uref=Ur[idx].contiguous())
    for k in ("X", "Lam", "dX", "dU", "loss_dp"):
        assert torch.equal(full[k][idx], sub[k]), k
    # chain rule (reference uav_PDP.py:67-75) from the materialised sensitivities, in torch float64
    dlx, dlu = full["X"] - Xr, U - Ur
    loss = (dlx ** 2).sum(dim=(1, 2)) + (dlu ** 2).sum(dim=(1, 2))
    dp = torch.einsum("bti,btir->br", dlx, full["dX"]) + torch.einsum("bta,btar->br", dlu, full["dU"])
    ok = torch.isfinite(full["loss_dp"]).all(dim=1) & torch.isfinite(dp).all(dim=1)
    assert bool(ok.all())          # the oracle finds no non-finite row on these seeded batches (max |dX/dtheta| 1e7)
    scale = dp[ok].abs().amax(dim=1, keepdim=True).clamp_min(1e-300)
    assert ((full["loss_dp"][ok, 1:] - dp[ok]).abs() / scale).max() < 1e-9
    assert ((full["loss_dp"][ok, 0] - loss[ok]).abs() / loss[ok]).max() < 1e-12


def test_aux_system_is_linear_in_its_initial_condition():
    """X_aux(X0 = a A + b B) - X_aux(0) = a (X_aux(A) - X_aux(0)) + b (X_aux(B) - X_aux(0)) for the fused LQR."""
    from pontryagin_differentiable_programming_b200 import systems
    dev = _dev()
    sys_ = systems.quadrotor_irl(0.1)
    B, H = 64, 50
    g = torch.Generator().manual_seed(1)
    import os
    g2 = np.load(os.path.join(os.path.dirname(__file__), "golden", "k2_demos.npz"))
    X = torch.as_tensor(np.repeat(g2["quadrotor_0_X"][None], B, 0), device=dev)
    U = torch.as_tensor(np.repeat(g2["quadrotor_0_U"][None], B, 0), device=dev)
    L = torch.as_tensor(np.repeat(g2["quadrotor_0_L"][None], B, 0), device=dev)
    th = torch.as_tensor(g2["quadrotor_true_parameter"] * 1.05, device=dev)
    A = torch.randn((B, 13, 9), dtype=torch.float64, generator=g).to(dev)
    Bm = torch.randn((B, 13, 9), dtype=torch.float64, generator=g).to(dev)
    run = lambda X0: sys_.aux_lqr(X, U, L, th, X0aux=X0)
    z, a_, b_, c_ = run(None), run(A), run(Bm), run((0.7 * A - 1.9 * Bm).contiguous())
    for k in ("dX", "dU"):
        lhs = c_[k] - z[k]
        rhs = 0.7 * (a_[k] - z[k]) - 1.9 * (b_[k] - z[k])
        assert (lhs - rhs).abs().max() < 1e-9 * max(1.0, rhs.abs().max().item())


@pytest.mark.parametrize("B,H", [(1, 1), (3, 2), (5, 17), (130, 33)])
def test_ragged_sizes_match_oracle(B, H):
    """Edge sizes: single step, single trajectory, batch not a multiple of the block, horizon not a multiple of the chunk."""
    from oracle import envs, pdp_oracle
    from pontryagin_differentiable_programming_b200 import systems
    dev = _dev()
    sys_ = systems.quadrotor_irl(0.1)
    oc = pdp_oracle.build_oc(envs.quadrotor(c=0.01, wthrust=0.1), 0.1)
    x0, th, U, Xr, Ur = _inputs(B, H, dev, seed=5)
    res = sys_.sweep(x0, th, U, Xref=Xr, Uref=Ur)
    for b in sorted({0, B - 1}):
        X, L, cost, dX, dU = pdp_oracle.pdp_sweep(oc, x0[b].cpu().numpy(), U[b].cpu().numpy(), th[b].cpu().numpy())
        rel = lambda a, r: np.max(np.abs(a - r)) / max(1e-300, np.max(np.abs(r)))
        assert rel(res["X"][b].cpu().numpy(), X) < 1e-12
        assert rel(res["Lam"][b].cpu().numpy(), L) < 1e-10
        assert rel(res["dX"][b].cpu().numpy(), dX) < 1e-7
        assert rel(res["dU"][b].cpu().numpy(), dU) < 1e-7


def test_sharded_equals_unsharded():
    """Weak-scaling layout: running two contiguous shards separately gives exactly the unsharded per-trajectory results,
    and the reduction of (loss, dp) over shards equals the global mean (what the one all-reduce computes)."""
    from pontryagin_differentiable_programming_b200 import distributed, systems
    dev = _dev()
    sys_ = systems.quadrotor_irl(0.1)
    B, H = 1000, 50
    x0, th, U, Xr, Ur = _inputs(B, H, dev, seed=2)
    full = sys_.sweep(x0, th, U, Xref=Xr, Uref=Ur, want_traj=False)["loss_dp"]
    parts = []
    for rank in range(2):
        lo, hi = distributed.shard_bounds(B, rank, 2)
        sl = lambda t: t[lo:hi].contiguous()
        parts.append(sys_.sweep(sl(x0), sl(th), sl(U), Xref=sl(Xr), Uref=sl(Ur), want_traj=False)["loss_dp"])
    assert torch.equal(torch.cat(parts), full)


def test_empty_batch_and_status_flags():
    """B = 0 is a no-op that returns correctly shaped empty tensors; a non-finite input raises status bit 0 for that
    trajectory only; an indefinite Quu raises bit 1 (the C ABI never throws for data-dependent conditions)."""
    from pontryagin_differentiable_programming_b200 import systems
    dev = _dev()
    sys_ = systems.quadrotor_irl(0.1)
    H = 7
    e = lambda *s: torch.empty(s, dtype=torch.float64, device=dev)
    res = sys_.sweep(e(0, 13), e(0, 9), e(0, H, 4))
    assert res["dX"].shape == (0, H + 1, 13, 9) and res["X"].shape == (0, H + 1, 13)
    x0, th, U, _, _ = _inputs(6, H, dev, seed=8)
    U[2, 3, 1] = float("nan")
    status = torch.zeros(6, dtype=torch.int32, device=dev)
    res = sys_.sweep(x0, th, U, status=status)
    torch.cuda.synchronize()
    st = status.cpu().numpy()
    assert st[2] & 1 and not (st[[0, 1, 3, 4, 5]] & 1).any()
    assert torch.isfinite(res["dX"][[0, 1, 3, 4, 5]]).all() or (st[[0, 1, 3, 4, 5]] & 2).any()


def test_pipelined_sweep_equals_the_unsplit_sweep_bit_for_bit():
    """pdp_sweep cuts the aux-LQR phase of large batches into sub-batches on two internal streams; every output must
    equal the unsplit call exactly (odd batch: uneven sub-batches), and the call must stay ordered in its stream."""
    import bench
    from pontryagin_differentiable_programming_b200 import systems
    dev = _dev()
    sys_ = systems.quadrotor_irl(0.1)
    B, H = 16384 + 3, 20
    x0, th, U, Xr, Ur = [torch.as_tensor(np.ascontiguousarray(a), device=dev) for a in bench.synth_quadrotor(B, H, seed=5)]
    res = {}
    try:
        for parts in (1, 0, 3, 7):
            sys_.set_sweep_parts(parts)
            status = torch.zeros(B, dtype=torch.int32, device=dev)
            out = sys_.sweep(x0, th, U, Xref=Xr, Uref=Ur, status=status)
            # no synchronisation: the host copy below is ordered behind the sweep in the current stream
            res[parts] = {k: v.clone() for k, v in out.items()}
            res[parts]["status"] = status.clone()
    finally:
        sys_.set_sweep_parts(0)
    torch.cuda.synchronize()
    for parts in (0, 3, 7):
        for k, v in res[1].items():
            assert torch.equal(v, res[parts][k]), (parts, k)


def test_both_backward_kernel_layouts_agree_on_the_gpu():
    """The one-trajectory-per-warp backward kernel (fallback for systems whose stack does not fit two rows per team
    lane, and the kernel of the generic dense LQR module) stays covered: same gains-driven outputs as the shipped
    two-trajectory kernel on a mid-size batch with an odd count."""
    import bench
    from tools.tune_aux_lqr import make, V1
    dev = _dev()
    B, H = 1027, 50
    x0, th, U, Xr, Ur = [torch.as_tensor(np.ascontiguousarray(a), device=dev) for a in bench.synth_quadrotor(B, H, seed=9)]
    outs = []
    for kw in ({}, V1):
        s = make(**kw)
        assert s.src.bwd_pack == (1 if kw else 2)
        outs.append(s.sweep(x0, th, U, Xref=Xr, Uref=Ur))
    torch.cuda.synchronize()
    for k in ("dX", "dU", "loss_dp", "X", "Lam"):
        a, b = outs[0][k], outs[1][k]
        assert float((a - b).abs().max()) <= 1e-11 * float(b.abs().max()), k


def test_c5_full_size_zero_residual_batch_independence_and_chain_rule():
    """B = 32768, H = 100 (config C5 per GPU, reference SysID.step PDP/PDP.py:1261-1296): (a) at theta_true the residual
    against observations rolled out by the same kernel is exactly zero, so loss = 0 and dp = 0 bit for bit; (b) two launches
    agree bitwise and a trajectory's result does not depend on the rest of the batch; (c) the fused (loss, dp) equals the
    contraction of the materialised dX/dtheta with the residual (PDP.py:1285-1290); (d) the reduction kernel's batch sums
    equal a float64 torch sum."""
    import bench
    from pontryagin_differentiable_programming_b200 import engine, systems
    dev = _dev()
    B, H = 32768, 100
    inputs, x0, th_true, theta = [torch.as_tensor(np.ascontiguousarray(a), device=dev) for a in bench.synth_sysid(B, H, seed=0)]
    sid = systems.quadrotor_sysid(0.1)
    Xobs = sid.step(inputs, None, th_true, x0=x0, want_traj=True)["X"]
    zero = sid.step(inputs, Xobs, th_true)["loss_dp"]
    fin = torch.isfinite(Xobs).all(dim=(1, 2))
    assert bool(fin.all()) and bool((zero == 0).all())
    full = sid.step(inputs, Xobs, theta, want_traj=True, want_sens=True)
    again = sid.step(inputs, Xobs, theta)
    assert torch.equal(full["loss_dp"], again["loss_dp"])
    idx = torch.tensor([0, 1, 31, 32, 63, 64, 4095, 16384, 32767], device=dev)
    sub = sid.step(inputs[idx].contiguous(), Xobs[idx].contiguous(), theta, want_traj=True, want_sens=True)
    for k in ("X", "dX", "loss_dp"):
        assert torch.equal(full[k][idx], sub[k]), k
    d = full["X"] - Xobs
    loss = (d ** 2).sum(dim=(1, 2))
    dp = torch.einsum("bti,btir->br", d, full["dX"])
    ok = torch.isfinite(full["loss_dp"]).all(dim=1) & torch.isfinite(dp).all(dim=1) & (loss > 0)
    assert bool(ok.all())          # the oracle finds no non-finite row on these seeded batches (max |dX/dtheta| 1e7)
    scale = dp[ok].abs().amax(dim=1, keepdim=True).clamp_min(1e-300)
    assert ((full["loss_dp"][ok, 1:] - dp[ok]).abs() / scale).max() < 1e-9
    assert ((full["loss_dp"][ok, 0] - loss[ok]).abs() / loss[ok]).max() < 1e-12
    rows = full["loss_dp"][ok].contiguous()
    sums = engine.reduce_loss_dp(rows)
    ref = rows.sum(dim=0)
    assert ((sums[:-1] - ref).abs() / ref.abs().clamp_min(1e-300)).max() < 1e-11 and int(sums[-1]) == rows.shape[0]


def test_c4_full_size_both_rollout_kernels_agree_and_gradient_matches_finite_differences():
    """H = 100 rocket (config C4, recmat semantics PDP/PDP.py:1100-1114): B = 8192 per GPU runs on the per-thread TMA kernel,
    B = 16384 on the register-prefetch kernel -- the same trajectories give bitwise the same X, Lam, J, dJ/dU on both;
    launches are deterministic; and dJ/dU agrees with central differences of J along a random direction (every trajectory)."""
    import bench
    from pontryagin_differentiable_programming_b200 import systems
    dev = _dev()
    H = 100
    x0, U = [torch.as_tensor(np.ascontiguousarray(a), device=dev) for a in bench.synth_rocket(16384, H, seed=0)]
    ro = systems.rocket_oc_adjoint(0.1)
    th = torch.zeros((1, 1), dtype=torch.float64, device=dev)
    big = ro.rollout_costate(x0, th, U, want_dHu=True)
    small = ro.rollout_costate(x0[:8192].contiguous(), th, U[:8192].contiguous(), want_dHu=True)
    again = ro.rollout_costate(x0[:8192].contiguous(), th, U[:8192].contiguous(), want_dHu=True)
    for k in ("X", "Lam", "cost", "dHu"):
        assert torch.equal(small[k], again[k]), k
        assert torch.equal(small[k], big[k][:8192]), k
    g = torch.Generator(device="cpu").manual_seed(1)
    D = torch.randn(U.shape, generator=g, dtype=torch.float64).to(dev)
    eps = 1e-5
    Jp = ro.rollout_costate(x0, th, U + eps * D)["cost"]
    Jm = ro.rollout_costate(x0, th, U - eps * D)["cost"]
    fd = (Jp - Jm) / (2 * eps)
    an = (big["dHu"] * D).sum(dim=(1, 2))
    ok = torch.isfinite(fd) & torch.isfinite(an)
    assert bool(ok.all())
    # scaled by |dJ/dU| |D| (Cauchy-Schwarz), not by the directional derivative itself, which can be ~0 by chance; the CPU
    # oracle gives 1e-11 on this scale at eps = 1e-5
    scale = big["dHu"].flatten(1).norm(dim=1) * D.flatten(1).norm(dim=1)
    err = ((fd - an).abs() / scale.clamp_min(1e-300))[ok]
    assert err.max() < 1e-6, float(err.max())
