"""Pick the sub-batch count of OCSystem.sweep_host (copy/compute overlap) on the GPU box."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pontryagin_differentiable_programming_b200 import systems  # noqa: E402

dev = torch.device("cuda:0")
s = systems.quadrotor_irl(0.1)
B, H = 16384, 50
pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in bench.synth_quadrotor(B, H)]
ldp = torch.empty((B, 10), dtype=torch.float64).pin_memory()
cost = torch.empty((B,), dtype=torch.float64).pin_memory()
for nch in (1, 2, 4, 8, 16, 32, 64):
    for keep in (1, 0):
        f = lambda: s.sweep_host(pinned[0], pinned[1], pinned[2], pinned[3], pinned[4], ldp, cost_h=cost, keep_dtraj=bool(keep), n_chunks=nch, device=dev)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            f()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(json.dumps({"n_chunks": nch, "keep_dtraj": keep, "ms": ms, "sweeps_per_s": B / ms * 1e3}), flush=True)
