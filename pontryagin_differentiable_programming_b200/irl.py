"""Device-resident outer loops of the PDP learning modes.

The reference runs ``theta <- theta - lr * dp`` in Python with one IPOPT solve + getAuxSys + lqrSolver per
demonstration per iteration (``Examples/IRL/quadrotor/uav_PDP.py:40-83``,
``Examples/SysID/quadrotor/uav_PDP.py:42-48``).  Here one iteration is a handful of launches for the whole
(sharded) demonstration batch: batched ocSolver (warm-started from the previous iterate) -> fused sweep with the
IRL loss / chain rule -> batch reduction kernel -> ONE all-reduce of (sum loss, sum dp, count) -> parameter update
(plain gradient descent like the reference scripts, or Adam).  Everything but the final scalar read-outs stays on the
GPU, and the whole iteration -- the NCCL all-reduce included -- can be replayed from one captured CUDA graph."""
from __future__ import annotations

import torch

from . import distributed, ocsolver


class _Update:
    """Parameter update on the device: ``gd`` = the reference scripts' ``theta - lr * dp``; ``adam`` = Adam on the same
    half-gradient ``dp`` (state m, v and the step counter are device tensors, so the update is graph-capturable)."""

    def __init__(self, kind, lr, r, device, betas=(0.9, 0.999), eps=1e-8):
        if kind not in ("gd", "adam"):
            raise ValueError("optimizer must be 'gd' or 'adam'")
        self.kind, self.lr, self.betas, self.eps = kind, float(lr), betas, float(eps)
        if kind == "adam":
            z = lambda: torch.zeros(r, dtype=torch.float64, device=device)
            self.m, self.v, self.t = z(), z(), torch.zeros((), dtype=torch.float64, device=device)

    def state(self):
        return (self.m, self.v, self.t) if self.kind == "adam" else ()

    def __call__(self, theta, dp):
        if self.kind == "gd":
            return theta - self.lr * dp
        b1, b2 = self.betas
        self.t.add_(1.0)
        self.m.mul_(b1).add_(dp, alpha=1.0 - b1)
        self.v.mul_(b2).addcmul_(dp, dp, value=1.0 - b2)
        mhat = self.m / (1.0 - torch.pow(torch.full_like(self.t, b1), self.t))
        vhat = self.v / (1.0 - torch.pow(torch.full_like(self.t, b2), self.t))
        return theta - self.lr * mhat / (vhat.sqrt() + self.eps)


def _capture(owner, fn, dev, restore):
    """Warm ``fn`` up twice on a side stream (NCCL communicators, lazily created workspaces and streams must exist before
    the capture), put ``restore()``'s tensors back, then capture one call of ``fn`` into a CUDA graph."""
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        saved = [(t, t.clone()) for t in restore()]
        for _ in range(2):
            fn()
        for dst, src in saved:
            dst.copy_(src)
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            out = fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    owner._graph_stream = side
    return graph, out


class _GraphOwner:
    def release_graph(self):
        """Drop the captured CUDA graph (and the tensors it pins).  In a sharded run the graph holds a captured NCCL
        all-reduce: release it on every rank BEFORE ``torch.distributed.destroy_process_group()``, which otherwise
        waits forever for the communicator's outstanding (captured) work."""
        if getattr(self, "_graph", None) is not None:
            torch.cuda.synchronize()
            self._graph = self._g_out = self._graph_keepalive = self._graph_stream = None


class IRLTrainer(_GraphOwner):
    """Inverse-RL / inverse-OC mode for a compiled ``OCSystem`` and a (per-rank shard of a) batch of demonstrations."""

    def __init__(self, system, demo_states, demo_controls, lr, warm_start=True, group=None, optimizer="gd"):
        self.sys = system
        self.Xd, self.Ud = demo_states.contiguous(), demo_controls.contiguous()
        self.x0 = self.Xd[:, 0, :].contiguous()
        self.H = self.Ud.shape[1]
        self.lr = float(lr)
        self.warm_start = warm_start
        self.group = group
        self.update = _Update(optimizer, lr, system.r, self.Xd.device)
        self.status = torch.zeros(self.Xd.shape[0], dtype=torch.int32, device=self.Xd.device)
        self._U = None
        self.last = None

    def gradient(self, theta):
        """-> (mean loss, mean dp[r]) over the GLOBAL demonstration batch; dp is the reference's half-gradient."""
        th = theta.reshape(1, -1).to(self.Xd.device, torch.float64)
        sol = ocsolver.solve(self.sys, self.x0, self.H, th, control_init=self._U if self.warm_start else None)
        if self.warm_start:
            self._U = sol["U"]
        self.status.zero_()
        res = self.sys.sweep(self.x0, th, sol["U"], Xref=self.Xd, Uref=self.Ud, want_traj=False, status=self.status)
        self.last = {"solution": sol, "loss_dp": res["loss_dp"]}
        return distributed.reduce_loss_dp(res["loss_dp"], self.group)

    def diagnostics(self):
        """Health of the last iteration (one host read): fraction of inner solves that converged, worst residual
        |dH/du|, and how many trajectories raised a kernel status flag (bit 0 non-finite, bit 1 Quu not positive definite
        -- the auxiliary-LQR gradient is only meaningful at a stationary point with Quu > 0)."""
        out = {"status_bit0_nonfinite": int((self.status & 1).ne(0).sum().item()),
               "status_bit1_quu_not_pd": int((self.status & 2).ne(0).sum().item())}
        sol = (self.last or {}).get("solution")
        if sol is not None:
            out["converged_fraction"] = float(sol["converged"].double().mean().item()) if torch.is_tensor(sol["converged"]) \
                else float(bool(sol["converged"]))
            out["max_grad_norm"] = float(sol["grad_norm"].max().item())
        return out

    def step(self, theta):
        """One iteration: returns (loss, theta_next)."""
        loss, dp = self.gradient(theta)
        return loss, self.update(theta.reshape(-1).to(dp.device, torch.float64), dp)

    # ------------------------------------------------------------------ CUDA-graph path
    def _iteration_fixed(self, theta, n_newton):
        th = theta.reshape(1, -1)
        sol = ocsolver.solve_fixed(self.sys, self.x0, self.H, th, self._fixed_state, n_iter=n_newton)
        res = self.sys.sweep(self.x0, th, sol["U"], Xref=self.Xd, Uref=self.Ud, want_traj=False, status=self.status)
        loss, dp = distributed.reduce_loss_dp(res["loss_dp"], self.group)
        return loss, self.update(theta.reshape(-1), dp), sol["grad_norm"].max()

    def step_graph(self, theta, n_newton=3):
        """The same iteration as :meth:`step` replayed from ONE captured CUDA graph (fixed ``n_newton`` Newton
        iterations with single-launch line searches, fused sweep, batch reduction, the all-reduce of a sharded run, the
        update): no host round-trips inside.  The first call brings the warm start in with the adaptive solver, warms
        the kernels (and the NCCL communicator) up and captures; every rank of a sharded run must call it the same number
        of times.  Returns (loss, theta_next, max residual |dH/du| of this rank's inner solves) as device tensors valid
        until the next call.  ``n_newton`` is baked into the graph: pick it for the largest parameter step the run will
        take (the early, large updates of the quadrotor example need ~10; 3 is enough once the loss has settled) and
        monitor the returned residual - converged problems cost nothing extra because their iterations are no-ops."""
        dev = self.Xd.device
        if getattr(self, "_graph", None) is None:
            self._theta_in = theta.detach().reshape(-1).to(dev, torch.float64).clone()
            sol = ocsolver.solve(self.sys, self.x0, self.H, self._theta_in.reshape(1, -1))     # cold start, adaptive
            self._fixed_state = ocsolver.FixedSolverState(self.x0.shape[0], self.H, self.sys.m, dev)
            self._fixed_state.U.copy_(sol["U"])
            fs = self._fixed_state
            self._graph, self._g_out = _capture(
                self, lambda: self._iteration_fixed(self._theta_in, n_newton), dev,
                lambda: (fs.U, fs.s_newton, fs.mu) + tuple(self.update.state()))
            # the captured launches hold raw pointers into the systems' workspaces: keep them alive with the graph
            self._graph_keepalive = (tuple(self.sys._ws.values()), ocsolver.newton_system(self.sys)._ws)
        self._theta_in.copy_(theta.reshape(-1))
        self._graph.replay()
        return self._g_out


class SysIDTrainer(_GraphOwner):
    """System-identification mode for a compiled ``SysIDSystem`` (reference PDP.py:1261-1296 + the GD loop of
    Examples/SysID/quadrotor/uav_PDP.py:42-48) on a (per-rank shard of a) batch of input / state trajectories."""

    def __init__(self, system, inputs, states, lr, group=None, optimizer="gd"):
        self.sys, self.inputs, self.states, self.lr, self.group = system, inputs.contiguous(), states.contiguous(), float(lr), group
        self.update = _Update(optimizer, lr, system.r, self.inputs.device)
        self.status = torch.zeros(self.inputs.shape[0], dtype=torch.int32, device=self.inputs.device)
        self._x0 = self.states[:, 0, :].contiguous()          # once, not one strided-copy kernel per iteration

    def gradient(self, theta):
        th = theta.reshape(1, -1).to(self.inputs.device, torch.float64)
        res = self.sys.step(self.inputs, self.states, th, x0=self._x0, status=self.status)
        return distributed.reduce_loss_dp(res["loss_dp"], self.group)

    def step(self, theta):
        loss, dp = self.gradient(theta)
        return loss, self.update(theta.reshape(-1).to(dp.device, torch.float64), dp)

    def step_graph(self, theta):
        """One SysID iteration (sweep kernel, batch reduction, all-reduce of a sharded run, update) replayed from a
        captured CUDA graph.  Returns (loss, theta_next) as device tensors valid until the next call."""
        dev = self.inputs.device
        if getattr(self, "_graph", None) is None:
            self._theta_in = theta.detach().reshape(-1).to(dev, torch.float64).clone()

            def it():
                res = self.sys.step(self.inputs, self.states, self._theta_in.reshape(1, -1), x0=self._x0, status=self.status)
                loss, dp = distributed.reduce_loss_dp(res["loss_dp"], self.group)
                return loss, self.update(self._theta_in, dp)

            self._graph, self._g_out = _capture(self, it, dev, lambda: tuple(self.update.state()))
        self._theta_in.copy_(theta.reshape(-1))
        self._graph.replay()
        return self._g_out
