"""GPU parity at the BASELINE.json shapes of the secondary configs (SURVEY 8(d) generators, the same ones bench.py uses):
C5 quadrotor SysID H = 100, C4 rocket adjoint gradient H = 100, C2 cartpole ControlPlanning at B = 4096 -- each against
the oracle on sampled trajectories -- plus the batch-reduction kernel and the host-buffer entry points of the C ABI."""
import numpy as np
import pytest
import torch

import bench
from oracle import envs, pdp_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-9        # the north star asks for 1e-6 relative; the deterministic sweeps agree far better


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _t(a, dev):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b)))


def _sample(B, k=40, seed=0):
    """k sampled trajectory indices incl. the first / last and the warp / block edges."""
    rng = np.random.default_rng(seed)
    fixed = [0, 1, 31, 32, 63, 64, B - 1]
    return sorted(set(fixed) | set(int(i) for i in rng.integers(0, B, k - len(fixed))))


def test_c5_sysid_h100_matches_oracle():
    """SysID.step (reference PDP/PDP.py:1261-1296) at the C5 shape: H = 100, several hundred trajectories of the C5
    generator, shared theta; rollout, loss and half-gradient per trajectory vs the oracle on 40 samples, the batch mean
    through the reduction kernel vs the oracle's own mean."""
    from pontryagin_differentiable_programming_b200 import distributed, systems
    dev = _dev()
    B, H = 700, 100
    inputs, x0, th_true, theta = bench.synth_sysid(B, H, seed=3)
    sys_ = systems.quadrotor_sysid(0.1)
    Xobs = sys_.step(_t(inputs, dev), None, _t(th_true, dev), x0=_t(x0, dev), want_traj=True)["X"]
    st = torch.zeros(B, dtype=torch.int32, device=dev)
    out = sys_.step(_t(inputs, dev), Xobs, _t(theta, dev), want_traj=True, want_sens=True, status=st)
    fused = sys_.step(_t(inputs, dev), Xobs, _t(theta, dev))
    assert torch.equal(out["loss_dp"], fused["loss_dp"])
    assert int(st.max()) == 0
    e = envs.quadrotor(c=0.01)
    sid = pdp_oracle.OracleSysID(e["X"], e["U"], e["dyn_params"], e["X"] + 0.1 * e["f"])
    idx = _sample(B)
    Xo_h = Xobs.cpu().numpy()
    ldp = out["loss_dp"].cpu().numpy()
    obs_list, in_list = [], []
    for b in idx:
        Xo = sid.integrateDyn(x0[b], inputs[b], th_true)
        assert _rel(Xo_h[b], Xo) < 1e-12
        X = sid.integrateDyn(x0[b], inputs[b], theta)
        S = np.stack(sid.sens(X, inputs[b], theta))
        assert _rel(out["X"][b].cpu().numpy(), X) < 1e-12
        assert _rel(out["dX"][b].cpu().numpy(), S) < TOL
        loss, dp = sid.step([inputs[b]], [Xo], theta)
        assert abs(ldp[b, 0] - loss) < TOL * abs(loss) and _rel(ldp[b, 1:], dp) < TOL
        obs_list.append(Xo)
        in_list.append(inputs[b])
    # batch mean (PDP.py:1293-1294) over the sampled trajectories through the reduction kernel
    loss_m, dp_m = distributed.reduce_loss_dp(out["loss_dp"][idx].contiguous())
    loss_ref, dp_ref = sid.step(in_list, obs_list, theta)
    assert abs(float(loss_m) - loss_ref) < TOL * abs(loss_ref) and _rel(dp_m.cpu().numpy(), dp_ref) < TOL


def test_c4_rocket_adjoint_h100_matches_oracle():
    """recmat semantics (PDP/PDP.py:1100-1114) at the C4 shape: rocket n = 13 m = 3 H = 100, C4 generator; X, J and dJ/dU."""
    from pontryagin_differentiable_programming_b200 import systems
    dev = _dev()
    B, H = 600, 100
    x0, U = bench.synth_rocket(B, H, seed=9)
    sys_ = systems.rocket_oc_adjoint(0.1)
    th = torch.zeros((1, 1), dtype=torch.float64, device=dev)
    st = torch.zeros(B, dtype=torch.int32, device=dev)
    out = sys_.rollout_costate(_t(x0, dev), th, _t(U, dev), want_dHu=True, status=st)
    e = envs.rocket(Jx=0.5, Jy=1., Jz=1., mass=1., l=1., wr=1, wv=1, wtilt=50, ww=1, wsidethrust=1, wthrust=0.4)
    cp = pdp_oracle.OracleCP(e["X"], e["U"], e["X"] + 0.1 * e["f"], e["path_cost"], e["final_cost"])
    checked = 0
    for b in _sample(B):
        cost, g, X = cp.adjoint_grad(x0[b], U[b])
        if not np.isfinite(cost):
            assert int(st[b]) & 1
            continue
        assert _rel(out["X"][b].cpu().numpy(), X) < 1e-11
        assert abs(float(out["cost"][b]) - cost) < 1e-11 * abs(cost)
        assert _rel(out["dHu"][b].cpu().numpy(), g) < TOL
        checked += 1
    assert checked >= 32


@pytest.mark.parametrize("policy", ["poly", "neural"])
def test_c2_cartpole_b4096_matches_oracle(policy):
    """ControlPlanning.step (PDP/PDP.py:850-878) at the C2 shape: B = 4096, H = 50, C2 generator; every output."""
    from pontryagin_differentiable_programming_b200 import systems
    dev = _dev()
    B, H = 4096, 50
    sys_ = systems.cartpole_cp(policy, H, 0.05)
    x0, theta = bench.synth_cartpole(B, sys_.r, seed=11)
    if policy == "neural":
        theta = 0.5 * theta
    e = envs.cartpole(mc=0.1, mp=0.1, l=1, wx=0.1, wq=0.6, wdx=0.1, wdq=0.1, wu=0.3)
    cp = pdp_oracle.OracleCP(e["X"], e["U"], e["X"] + 0.05 * e["f"], e["path_cost"], e["final_cost"])
    cp.set_poly(np.linspace(0, H, 6)) if policy == "poly" else cp.set_neural([4, 4])
    out = sys_.step(_t(x0, dev), H, _t(theta, dev), want_traj=True, want_sens=True)
    fused = sys_.step(_t(x0, dev), H, _t(theta, dev))
    assert torch.equal(out["loss_dp"], fused["loss_dp"])
    ldp = out["loss_dp"].cpu().numpy()
    checked = 0
    for b in _sample(B, 36):
        cost, g, X, U, dX, dU = cp.step(x0[b], H, theta[b], return_traj=True)
        if not (np.isfinite(cost) and np.isfinite(dX).all()) or np.max(np.abs(X)) > 1e6:
            continue                        # random polynomial policies can blow the cart-pole up within 50 steps
        assert _rel(out["X"][b].cpu().numpy(), X) < TOL and _rel(out["U"][b].cpu().numpy(), U) < TOL
        assert abs(ldp[b, 0] - cost) < TOL * abs(cost)
        assert _rel(out["dX"][b].cpu().numpy(), dX) < 1e-8 and _rel(out["dU"][b].cpu().numpy(), dU) < 1e-8
        assert _rel(ldp[b, 1:], g) < 1e-8
        checked += 1
    assert checked >= 32


@pytest.mark.parametrize("B,r", [(1, 0), (7, 5), (64, 9), (1000, 9), (16384, 9), (333, 45), (50, 300)])
def test_batch_reduction_kernel(B, r):
    """pdp_reduce_loss_dp: (sum loss, sum dp, count) equals torch's float64 column sums, is deterministic, and its scratch
    is reusable call after call (the ticket word returns to zero)."""
    from pontryagin_differentiable_programming_b200 import engine
    dev = _dev()
    g = torch.Generator().manual_seed(B + r)
    ldp = torch.randn((B, r + 1), dtype=torch.float64, generator=g).to(dev)
    a = engine.reduce_loss_dp(ldp).clone()
    b = engine.reduce_loss_dp(ldp).clone()
    assert torch.equal(a, b)
    ref = ldp.sum(dim=0)
    assert float(a[-1]) == B
    assert torch.allclose(a[:-1], ref, rtol=1e-12, atol=1e-12 * float(ldp.abs().sum()))


def test_batch_reduction_is_graph_capturable():
    from pontryagin_differentiable_programming_b200 import engine
    dev = _dev()
    ldp = torch.randn((500, 6), dtype=torch.float64, device=dev)
    out = torch.empty(7, dtype=torch.float64, device=dev)
    engine.reduce_loss_dp(ldp, out=out)
    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        engine.reduce_loss_dp(ldp, out=out)              # creates this stream's scratch outside the capture
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            engine.reduce_loss_dp(ldp, out=out)
    torch.cuda.current_stream(dev).wait_stream(s)
    for k in range(3):
        ldp.mul_(1.5)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.allclose(out[:-1], ldp.sum(dim=0), rtol=1e-12)


def test_host_buffer_entry_points_match_device_path():
    """pdp_sweep_host_traj / pdp_rollout_costate_host / pdp_sens_fwd_host (pinned host buffers, sub-batches on two
    streams, odd batch) return exactly what the device-resident calls produce."""
    from pontryagin_differentiable_programming_b200 import systems
    dev = _dev()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    # --- IRL sweep with the sensitivities copied back
    B, H = 37, 19
    host = bench.synth_quadrotor(B, H, seed=4)
    s3 = systems.quadrotor_irl(0.1)
    ref = s3.sweep(*[_t(a, dev) for a in host[:3]], Xref=_t(host[3], dev), Uref=_t(host[4], dev))
    p = [pin(a) for a in host]
    mk = lambda *shape: torch.empty(shape, dtype=torch.float64).pin_memory()
    ldp, cost, X_h, L_h, dX_h, dU_h = mk(B, 10), mk(B), mk(B, H + 1, 13), mk(B, H, 13), mk(B, H + 1, 13, 9), mk(B, H, 4, 9)
    s3.sweep_host(p[0], p[1], p[2], p[3], p[4], ldp, cost_h=cost, n_chunks=3, X_h=X_h, Lam_h=L_h, dX_h=dX_h, dU_h=dU_h)
    torch.cuda.synchronize()
    for got, key in ((ldp, "loss_dp"), (cost, "cost"), (X_h, "X"), (L_h, "Lam"), (dX_h, "dX"), (dU_h, "dU")):
        assert torch.equal(got, ref[key].cpu()), key
    with pytest.raises(ValueError):
        s3.sweep_host(p[0], p[1], p[2], p[3][:, :-1].contiguous().pin_memory(), p[4], ldp)
    # --- adjoint gradient
    B, H = 45, 100
    x0, U = bench.synth_rocket(B, H, seed=2)
    s4 = systems.rocket_oc_adjoint(0.1)
    th = torch.zeros((1, 1), dtype=torch.float64)
    ref = s4.rollout_costate(_t(x0, dev), th.to(dev), _t(U, dev), want_dHu=True)
    cost, dHu = mk(B), mk(B, H, 3)
    s4.rollout_costate_host(pin(x0), th.pin_memory(), pin(U), cost_h=cost, dHu_h=dHu, n_chunks=4)
    torch.cuda.synchronize()
    assert torch.equal(cost, ref["cost"].cpu()) and torch.equal(dHu, ref["dHu"].cpu())
    # --- SysID step with the per-sub-batch reductions
    B, H = 101, 100
    inputs, x0, th_true, theta = bench.synth_sysid(B, H, seed=8)
    s5 = systems.quadrotor_sysid(0.1)
    Xobs = s5.step(_t(inputs, dev), None, _t(th_true, dev), x0=_t(x0, dev), want_traj=True)["X"]
    ref = s5.step(_t(inputs, dev), Xobs, _t(theta, dev))["loss_dp"]
    ldp, sums = mk(B, 6), torch.zeros((4, 7), dtype=torch.float64).pin_memory()
    for _ in range(2):                                      # twice: the reduction scratch must be reusable
        s5.step_host(pin(x0), pin(theta.reshape(1, -1)), H, inputs_h=pin(inputs), Xobs_h=Xobs.cpu().pin_memory(),
                     loss_dp_h=ldp, sums_h=sums, n_chunks=4)
        torch.cuda.synchronize()
        assert torch.equal(ldp, ref.cpu())
        tot = sums.sum(dim=0)
        assert float(tot[-1]) == B and torch.allclose(tot[:-1], ref.sum(dim=0).cpu(), rtol=1e-12)
