"""Batched optimal-control solver on the GPU (replaces IPOPT inside ``OCSys.ocSolver``).

The reference transcribes the problem by multiple shooting and calls IPOPT once per trajectory
(``PDP/PDP.py:131-182``).  Here all B problems are solved together by a globalised Newton method
on the controls (single shooting):

  repeat
    1. ``pdp_k_rollout_costate``: X, cost, costates Lam and the gradient g_t = dH/du_t;
    2. Newton direction from the SAME fused Riccati kernel as the PDP backward sweep, run on the
       ``NewtonModuleSource`` variant (one column, "Hue" := g): exact second-order Hessians
       (s = 1) with a per-trajectory fall-back to Gauss-Newton / iLQR Hessians (s = 0) and a
       Levenberg shift mu when Quu is not positive definite or the line search fails;
    3. batched Armijo back-tracking with CLOSED-LOOP trial rollouts (DDP style):
       u_t = u_t + alpha k_t + K_t (x_t - x_t^old), the gains being the ones the Riccati sweep spilled
       (every trial is one launch of the rollout kernel in feedback mode).
  until max_t |dH/du_t| <= tol for every trajectory.

Everything stays on the device; the host only reads one convergence scalar per iteration.
The converged costates are the PMP costates, ``Lam[:, t] = lambda_{t+1}`` -- the convention of the
reference's ``costate_traj_opt`` (PDP.py:195, 203-209).
"""
from __future__ import annotations

import torch

from . import backend, build, codegen
from .engine import OCSystem, _ptr, require_cuda


class _NewtonSystem:
    def __init__(self, oc: OCSystem):
        s = oc.src
        pack = 2 if (s.bwd_pack == 2 or (s.n <= 16 and s.m + 1 <= 16)) else 1     # the Newton stack has one column
        d = OCSystem.BWD_DEFAULTS[pack]
        self.src = codegen.NewtonModuleSource(s.x, s.u, s.th, s.dyn, s.c, s.h, chunk=d["chunk"],
                                              warps_per_block=d["warps_per_block"], min_blocks=d["min_blocks"],
                                              fwd_warps_per_block=s.wpbf, fwd_min_blocks=s.min_blocks_f,
                                              keep_fg=d["keep_fg"], fast_rcp=s.fast_rcp, early_solve=s.early_solve,
                                              bwd_pack=pack)
        self.module_path = build.compile_module(self.src.source(), self.src.key())
        self._handle = None
        self.n, self.m, self.nth = self.src.n, self.src.m, self.src.nth
        self._ws = None

    @property
    def handle(self):
        if self._handle is None:
            self._handle = backend.SystemHandle(self.module_path)
        return self._handle

    def direction(self, X, U, Lam, theta_ext, status):
        """-> dU[B,H,m] (Newton step), dX[B,H+1,n]."""
        dev = X.device
        B, H = U.shape[0], U.shape[1]
        dX = torch.empty((B, H + 1, self.n, 1), dtype=torch.float64, device=dev)
        dU = torch.empty((B, H, self.m, 1), dtype=torch.float64, device=dev)
        need = self.handle.workspace_bytes(backend.OP_AUX_LQR, B, H)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            backend.check(self.handle.lib.pdp_aux_lqr(self.handle.ptr, B, H, _ptr(X), _ptr(U), _ptr(Lam), _ptr(theta_ext),
                                                      self.nth, None, 0, _ptr(dX), _ptr(dU), None, None, None,
                                                      _ptr(self._ws), self._ws.numel(), _ptr(status), st), "pdp_aux_lqr(newton)")
        return dU.squeeze(-1), dX.squeeze(-1), self._ws


def newton_system(oc: OCSystem) -> _NewtonSystem:
    if getattr(oc, "_newton", None) is None:
        oc._newton = _NewtonSystem(oc)
    return oc._newton


def solve(oc: OCSystem, x0, horizon, theta, control_init=None, tol=1e-8, max_iter=300, max_backtrack=25,
          verbose=False):
    """Solve B optimal-control problems.  x0[B,n], theta[B|1,r] CUDA float64; control_init[B,H,m] or None
    (zeros, like the reference's NLP initial guess).  Returns dict X, U, Lam, cost, iters, converged, grad_norm."""
    require_cuda()
    dev = x0.device
    B, H = x0.shape[0], int(horizon)
    nt = newton_system(oc)
    if theta.dim() == 1:
        theta = theta.unsqueeze(0)
    if theta.shape[0] == 1 and B > 1:
        theta = theta.expand(B, -1)
    theta = theta.contiguous()
    U = torch.zeros((B, H, oc.m), dtype=torch.float64, device=dev) if control_init is None else control_init.clone()
    ext = torch.empty((B, oc.r + 2), dtype=torch.float64, device=dev)
    ext[:, :oc.r] = theta
    s_newton = torch.ones(B, dtype=torch.float64, device=dev)     # 1 = exact Hessians, 0 = Gauss-Newton
    mu = torch.zeros(B, dtype=torch.float64, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    cur = oc.rollout_costate(x0, theta, U, want_dHu=True)
    it = 0
    gnorm = cur["dHu"].abs().amax(dim=(1, 2))
    for it in range(1, max_iter + 1):
        scale = 1.0 + cur["Lam"].abs().amax(dim=(1, 2))
        active = gnorm > tol * scale
        if not bool(active.any()):
            it -= 1
            break
        ext[:, oc.r] = s_newton
        ext[:, oc.r + 1] = mu
        status.zero_()
        dU, _, gains = nt.direction(cur["X"], U, cur["Lam"], ext, status)
        slope = (cur["dHu"] * dU).sum(dim=(1, 2))
        bad_dir = (status != 0) | ~torch.isfinite(slope) | (slope >= 0)
        # back-tracking line search on the true cost, per trajectory
        alpha = torch.where(active & ~bad_dir, 1.0, 0.0).to(torch.float64)
        accepted = ~active | bad_dir
        best = {k: v.clone() for k, v in cur.items()}
        Unew = U.clone()
        for _ in range(max_backtrack):
            if bool(accepted.all()):
                break
            trial = oc.rollout_feedback(x0, theta, U, cur["X"], gains, alpha, want_dHu=True)
            Utry = trial["U"]
            ok = (~accepted) & torch.isfinite(trial["cost"]) & (trial["cost"] <= cur["cost"] + 1e-4 * alpha * slope)
            if bool(ok.any()):
                sel = ok.view(B, 1, 1)
                Unew = torch.where(sel, Utry, Unew)
                for k in ("X", "Lam", "dHu"):
                    best[k] = torch.where(sel, trial[k], best[k])
                best["cost"] = torch.where(ok, trial["cost"], best["cost"])
            accepted = accepted | ok
            alpha = torch.where(accepted, alpha, alpha * 0.5)
        failed = active & (bad_dir | ~accepted)
        # failed trajectories: first drop to Gauss-Newton Hessians, then raise the Levenberg shift
        was_gn = s_newton == 0
        mu = torch.where(failed & was_gn, torch.clamp(mu * 10.0, min=1e-6), torch.where(failed, mu, mu * 0.1))
        mu = torch.where(mu < 1e-12, torch.zeros_like(mu), mu)
        s_newton = torch.where(failed, torch.zeros_like(s_newton), s_newton)
        # successful full Newton-like steps re-enable exact Hessians
        s_newton = torch.where(~failed & (alpha >= 0.5) & active, torch.ones_like(s_newton), s_newton)
        U, cur = Unew, best
        gnorm = cur["dHu"].abs().amax(dim=(1, 2))
        if verbose:
            print("ocsolver iter %3d  cost[0]=%.10g  max|dHu|=%.3e  alpha[0]=%.3g  gn=%d  mu_max=%.1e"
                  % (it, cur["cost"][0].item(), gnorm.max().item(), alpha[0].item(), int((s_newton == 0).sum()), mu.max().item()))
    scale = 1.0 + cur["Lam"].abs().amax(dim=(1, 2))
    return {"X": cur["X"], "U": U, "Lam": cur["Lam"], "cost": cur["cost"], "iters": it,
            "converged": gnorm <= tol * scale, "grad_norm": gnorm}


START_SCALES = (0.0, 0.1, 1.0, 1.0, 3.0, 3.0, 10.0, 10.0)


def solve_multistart(oc: OCSystem, x0, horizon, theta, n_starts=8, seed=0, **opts):
    """Several initial control guesses per problem solved as ONE batch; the best stationary point wins.

    Non-convex problems (rocket landing, swing-ups) have several local minima and no local solver -- IPOPT
    included -- is guaranteed the best one; the reference always cold-starts IPOPT from zero (PDP.py:146-167).
    Start 0 is that same all-zero guess; the others are seeded Gaussian controls of growing scale."""
    dev = x0.device
    B, H, S_ = x0.shape[0], int(horizon), int(n_starts)
    if S_ <= 1:
        return solve(oc, x0, H, theta, **opts)
    if theta.dim() == 1:
        theta = theta.unsqueeze(0)
    if theta.shape[0] == 1:
        theta = theta.expand(B, -1)
    gen = torch.Generator(device="cpu").manual_seed(int(seed))
    noise = torch.randn((S_, B, H, oc.m), dtype=torch.float64, generator=gen).to(dev)
    scales = torch.tensor([START_SCALES[i % len(START_SCALES)] for i in range(S_)], dtype=torch.float64, device=dev)
    U0 = (noise * scales.view(S_, 1, 1, 1)).reshape(S_ * B, H, oc.m).contiguous()
    sol = solve(oc, x0.repeat(S_, 1), H, theta.repeat(S_, 1).contiguous(), control_init=U0, **opts)
    cost = torch.where(sol["converged"] & torch.isfinite(sol["cost"]), sol["cost"], torch.full_like(sol["cost"], float("inf")))
    cost = cost.view(S_, B)
    fallback = torch.where(torch.isfinite(sol["cost"]), sol["cost"], torch.full_like(sol["cost"], float("inf"))).view(S_, B)
    cost = torch.where(torch.isfinite(cost).any(dim=0, keepdim=True), cost, fallback)
    best = cost.argmin(dim=0)                                   # [B]
    idx = best * B + torch.arange(B, device=dev)
    out = {k: sol[k][idx] for k in ("X", "U", "Lam", "cost", "converged", "grad_norm")}
    out["iters"] = sol["iters"]
    out["start"] = best
    return out


DEFAULT_ALPHAS = (1.0, 0.5, 0.25, 0.125, 0.0625, 0.03125, 0.015625, 0.0078125)


class FixedSolverState:
    """Per-problem solver state that persists across calls of :func:`solve_fixed` (warm start, Hessian mode, shift)."""

    def __init__(self, B, H, m, device, alphas=None):
        z = lambda *s: torch.zeros(s, dtype=torch.float64, device=device)
        self.U = z(B, H, m)
        self.s_newton = torch.ones(B, dtype=torch.float64, device=device)
        self.mu = z(B)
        # constants created here (outside any graph capture: host -> device copies cannot be captured)
        self.alphas = tuple(DEFAULT_ALPHAS if alphas is None else alphas)
        self.a_row = torch.tensor(self.alphas, dtype=torch.float64, device=device)
        self.a_tab = self.a_row.repeat(B).contiguous()
        self.rows = torch.arange(B, device=device)


def solve_fixed(oc: OCSystem, x0, horizon, theta, state: FixedSolverState, n_iter=3, tol=1e-8):
    """Fixed-shape variant of :func:`solve` for CUDA-graph capture: ``n_iter`` Newton/DDP iterations, each with the
    WHOLE back-tracking line search evaluated in one launch (``len(alphas)`` closed-loop candidates per problem),
    no host synchronisation and no data-dependent control flow.  Warm-starts from and updates ``state`` in place.
    Returns the final dict X, U, Lam, cost, dHu, grad_norm (all device tensors)."""
    require_cuda()
    dev = x0.device
    B, H = x0.shape[0], int(horizon)
    nt = newton_system(oc)
    if theta.dim() == 1:
        theta = theta.unsqueeze(0)
    if theta.shape[0] == 1 and B > 1:
        theta = theta.expand(B, -1)
    theta = theta.contiguous()
    A = len(state.alphas)
    a_row, a_tab, rows = state.a_row, state.a_tab, state.rows
    ext = torch.empty((B, oc.r + 2), dtype=torch.float64, device=dev)
    ext[:, :oc.r] = theta
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    U, s_newton, mu = state.U, state.s_newton, state.mu
    for _ in range(n_iter):
        cur = oc.rollout_costate(x0, theta, U, want_dHu=True)
        gnorm = cur["dHu"].abs().amax(dim=(1, 2))
        active = gnorm > tol * (1.0 + cur["Lam"].abs().amax(dim=(1, 2)))
        ext[:, oc.r] = s_newton
        ext[:, oc.r + 1] = mu
        status.zero_()
        dU, _, gains = nt.direction(cur["X"], U, cur["Lam"], ext, status)
        slope = (cur["dHu"] * dU).sum(dim=(1, 2))
        bad = (status != 0) | ~torch.isfinite(slope) | (slope >= 0)
        trial = oc.rollout_feedback(x0, theta, U, cur["X"], gains, a_tab, want_dHu=False, want_costate=False, group=A)
        costs = trial["cost"].view(B, A)
        ok = torch.isfinite(costs) & (costs <= cur["cost"][:, None] + 1e-4 * a_row[None, :] * slope[:, None])
        ok = ok & active[:, None] & ~bad[:, None]
        any_ok = ok.any(dim=1)
        first = torch.argmax(ok.to(torch.int8), dim=1)                 # alphas are descending: the largest accepted step
        Usel = trial["U"].view(B, A, H, oc.m)[rows, first]
        U = torch.where(any_ok[:, None, None], Usel, U)
        failed = active & ~any_ok
        was_gn = s_newton == 0
        mu = torch.where(failed & was_gn, torch.clamp(mu * 10.0, min=1e-6), torch.where(failed, mu, mu * 0.1))
        mu = torch.where(mu < 1e-12, torch.zeros_like(mu), mu)
        s_newton = torch.where(failed, torch.zeros_like(s_newton), s_newton)
        s_newton = torch.where(any_ok & (a_row[first] >= 0.5), torch.ones_like(s_newton), s_newton)
    state.U.copy_(U)
    state.s_newton.copy_(s_newton)
    state.mu.copy_(mu)
    final = oc.rollout_costate(x0, theta, state.U, want_dHu=True)
    final["U"] = state.U
    final["grad_norm"] = final["dHu"].abs().amax(dim=(1, 2))
    return final
