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
