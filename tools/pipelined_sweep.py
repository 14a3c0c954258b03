"""Measure OCSystem.sweep vs sweep_pipelined (2-4 batch parts on two streams) on the GPU box."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pontryagin_differentiable_programming_b200 import systems

dev = torch.device("cuda:0")
s = systems.quadrotor_irl(0.1)
B, H = 16384, 50
x0, th, U, Xr, Ur = [torch.as_tensor(np.ascontiguousarray(a), device=dev) for a in bench.synth_quadrotor(B, H)]
mk = lambda *sh: torch.empty(sh, dtype=torch.float64, device=dev)
out = {"X": mk(B, H + 1, 13), "Lam": mk(B, H, 13), "cost": mk(B), "dX": mk(B, H + 1, 13, 9), "dU": mk(B, H, 4, 9), "loss_dp": mk(B, 10)}
ref = s.sweep(x0, th, U, Xref=Xr, Uref=Ur)
for parts in (1, 2, 3, 4, 6, 8):
    f = lambda: s.sweep_pipelined(x0, th, U, out, Xref=Xr, Uref=Ur, n_parts=parts)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    same = all(torch.equal(out[k], ref[k]) for k in ("X", "Lam", "dX", "dU", "loss_dp"))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(json.dumps({"n_parts": parts, "ms": ms, "sweeps_per_s": B / ms * 1e3, "bitwise_equal_to_sweep": same}), flush=True)
