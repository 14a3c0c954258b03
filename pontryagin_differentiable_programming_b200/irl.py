"""Device-resident outer loops of the PDP learning modes.

The reference runs ``theta <- theta - lr * dp`` in Python with one IPOPT solve + getAuxSys + lqrSolver per
demonstration per iteration (``Examples/IRL/quadrotor/uav_PDP.py:40-83``,
``Examples/SysID/quadrotor/uav_PDP.py:42-48``).  Here one iteration is a handful of launches for the whole
(sharded) demonstration batch: batched ocSolver (warm-started from the previous iterate) -> fused sweep with the
IRL loss / chain rule -> one all-reduce of (sum loss, sum dp, count) -> parameter update.  Everything but the
final scalar read-outs stays on the GPU."""
from __future__ import annotations

import torch

from . import distributed, ocsolver


class IRLTrainer:
    """Inverse-RL / inverse-OC mode for a compiled ``OCSystem`` and a batch of demonstrations."""

    def __init__(self, system, demo_states, demo_controls, lr, warm_start=True, group=None):
        self.sys = system
        self.Xd, self.Ud = demo_states.contiguous(), demo_controls.contiguous()
        self.x0 = self.Xd[:, 0, :].contiguous()
        self.H = self.Ud.shape[1]
        self.lr = float(lr)
        self.warm_start = warm_start
        self.group = group
        self._U = None
        self.last = None

    def gradient(self, theta):
        """-> (mean loss, mean dp[r]) over the GLOBAL demonstration batch; dp is the reference's half-gradient."""
        th = theta.reshape(1, -1).to(self.Xd.device, torch.float64)
        sol = ocsolver.solve(self.sys, self.x0, self.H, th, control_init=self._U if self.warm_start else None)
        if self.warm_start:
            self._U = sol["U"]
        res = self.sys.sweep(self.x0, th, sol["U"], Xref=self.Xd, Uref=self.Ud, want_traj=False)
        self.last = {"solution": sol, "loss_dp": res["loss_dp"]}
        return distributed.reduce_loss_dp(res["loss_dp"], self.group)

    def step(self, theta):
        """One gradient-descent iteration: returns (loss, theta_next)."""
        loss, dp = self.gradient(theta)
        return loss, theta.reshape(-1) - self.lr * dp

    # ------------------------------------------------------------------ CUDA-graph path
    def _iteration_fixed(self, theta, n_newton):
        th = theta.reshape(1, -1)
        sol = ocsolver.solve_fixed(self.sys, self.x0, self.H, th, self._fixed_state, n_iter=n_newton)
        res = self.sys.sweep(self.x0, th, sol["U"], Xref=self.Xd, Uref=self.Ud, want_traj=False)
        loss, dp = distributed.reduce_loss_dp(res["loss_dp"], self.group)
        return loss, theta.reshape(-1) - self.lr * dp, sol["grad_norm"].max()

    def step_graph(self, theta, n_newton=3):
        """The same iteration as :meth:`step` replayed from ONE captured CUDA graph (fixed ``n_newton`` Newton
        iterations with single-launch line searches, fused sweep, update): no host round-trips inside.  The first call
        brings the warm start in with the adaptive solver, warms the kernels up and captures.  Returns
        (loss, theta_next, max residual |dH/du| of the inner solves) as device tensors valid until the next call.
        ``n_newton`` is baked into the graph: pick it for the largest parameter step the run will take (the early,
        large updates of the quadrotor example need ~10; 3 is enough once the loss has settled) and monitor the returned
        residual - converged problems cost nothing extra because their iterations are no-ops."""
        dev = self.Xd.device
        if getattr(self, "_graph", None) is None:
            if torch.distributed.is_available() and torch.distributed.is_initialized() and \
                    torch.distributed.get_world_size(self.group) > 1:
                raise RuntimeError("step_graph captures a single-GPU iteration; use step() for sharded runs")
            self._theta_in = theta.detach().reshape(-1).to(dev, torch.float64).clone()
            sol = ocsolver.solve(self.sys, self.x0, self.H, self._theta_in.reshape(1, -1))     # cold start, adaptive
            self._fixed_state = ocsolver.FixedSolverState(self.x0.shape[0], self.H, self.sys.m, dev)
            self._fixed_state.U.copy_(sol["U"])
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                saved = (self._fixed_state.U.clone(), self._fixed_state.s_newton.clone(), self._fixed_state.mu.clone())
                for _ in range(2):                                                               # warm-up outside capture
                    self._iteration_fixed(self._theta_in, n_newton)
                for dst, src in zip((self._fixed_state.U, self._fixed_state.s_newton, self._fixed_state.mu), saved):
                    dst.copy_(src)
            torch.cuda.current_stream(dev).wait_stream(side)
            # the captured launches hold raw pointers into the system's workspaces: keep them alive with the graph
            self._graph_keepalive = (tuple(self.sys._ws.values()), ocsolver.newton_system(self.sys)._ws)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._g_out = self._iteration_fixed(self._theta_in, n_newton)
        self._theta_in.copy_(theta.reshape(-1))
        self._graph.replay()
        return self._g_out


class SysIDTrainer:
    """System-identification mode for a compiled ``SysIDSystem`` (reference PDP.py:1261-1296 + the GD loop)."""

    def __init__(self, system, inputs, states, lr, group=None):
        self.sys, self.inputs, self.states, self.lr, self.group = system, inputs.contiguous(), states.contiguous(), float(lr), group

    def gradient(self, theta):
        th = theta.reshape(1, -1).to(self.inputs.device, torch.float64)
        return distributed.reduce_loss_dp(self.sys.step(self.inputs, self.states, th)["loss_dp"], self.group)

    def step(self, theta):
        loss, dp = self.gradient(theta)
        return loss, theta.reshape(-1) - self.lr * dp
