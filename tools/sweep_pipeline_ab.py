"""A/B of pdp_sweep's sub-batch pipeline at C3 (16 384 trajectories): number of sub-batches.  (The variant that also ran
the rollout / costate kernel per sub-batch was measured in round 2 -- profiles/r2k_sweep_pipeline_ab.json, slower at every
split -- and removed.)

  python tools/sweep_pipeline_ab.py      # on the GPU box; writes gpurun_out/sweep_pipeline_ab.json
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pontryagin_differentiable_programming_b200 import systems  # noqa: E402
from tools.tune_aux_lqr import make  # noqa: E402

# module variants to put through the pipelined sweep (keyword overrides of tune_aux_lqr.BASE); {} = the shipped module
VARIANTS = [dict()]


def main():
    dev = torch.device("cuda:0")
    B, H = 16384, 50
    d = [torch.as_tensor(np.ascontiguousarray(a), device=dev) for a in bench.synth_quadrotor(B, H, seed=(0, 0))]
    n, m, r = 13, 4, 9
    mk = lambda *sh: torch.empty(sh, dtype=torch.float64, device=dev)
    out = {"X": mk(B, H + 1, n), "Lam": mk(B, H, n), "cost": mk(B), "dX": mk(B, H + 1, n, r), "dU": mk(B, H, m, r), "loss_dp": mk(B, r + 1)}
    rows, ref = [], None
    for var in VARIANTS:
        s = make(**var)
        for parts in (1, 2, 3, 4, 5, 6, 8):
            s.set_sweep_parts(parts)
            fn = lambda: s.sweep(d[0], d[1], d[2], Xref=d[3], Uref=d[4], out=out)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / 20)
            chk = out["dX"][::257].clone()
            if ref is None:
                ref = chk
            rows.append({"variant": var, "parts": parts, "ms_per_sweep": best, "sweeps_per_s": B / best * 1e3,
                         "identical_to_first": bool(torch.equal(chk, ref))})
            print(json.dumps(rows[-1]), flush=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "sweep_pipeline_ab.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
