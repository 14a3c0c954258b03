"""Smallest case of the per-thread TMA rollout kernel for `compute-sanitizer --tool racecheck --racecheck-report hazard`:
3 rocket trajectories, H = 19 (three chunks of 8: both slots and both mbarriers are reused), outputs compared with the
register-prefetch kernel (closed-loop entry with zero gains takes that kernel)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from pontryagin_differentiable_programming_b200 import systems

dev = torch.device("cuda:0")
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3      # B = 1: one active lane per warp, nothing for the tool to mis-attribute
x0, U = bench.synth_rocket(B, 19, seed=3)
ro = systems.rocket_oc_adjoint(0.1)
o = ro.rollout_costate(t(x0), torch.zeros((1, 1), dtype=torch.float64, device=dev), t(U), want_dHu=True)
torch.cuda.synchronize()
print("tma rollout ok", float(o["cost"][0]), bool(torch.isfinite(o["Lam"]).all()))
