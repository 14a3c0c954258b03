"""Secondary BASELINE.json configs (C2, C4, C5) at their per-GPU sizes: sweeps/s and algorithmic GB/s.

  python tools/bench_configs.py            # on the GPU box; writes gpurun_out/configs.json
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pontryagin_differentiable_programming_b200 import systems  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    rows = []
    # ---- C2 cartpole ControlPlanning, H=50, B=4096
    for policy in ("poly", "neural"):
        s = systems.cartpole_cp(policy, 50, 0.05)
        B, H = 4096, 50
        x0 = (0.1 * torch.randn((B, 4), dtype=torch.float64, generator=g)).to(dev)
        th = ((1.0 if policy == "poly" else 0.3) * torch.randn((B, s.r), dtype=torch.float64, generator=g)).to(dev)
        ms = timeit(lambda: s.step(x0, H, th))
        n, m, r = s.n, s.m, s.r
        alg = 8 * (n + r + r + 1)                                 # fused: read x0, theta; write (loss, dtheta)
        alg_full = 8 * (n + r + (H + 1) * n + H * m + (H + 1) * n * r + H * m * r + r + 1)
        ms_full = timeit(lambda: s.step(x0, H, th, want_traj=True, want_sens=True))
        rows.append({"config": "C2 cartpole ControlPlanning %s r=%d H=50 B=4096" % (policy, r), "ms_fused": ms,
                     "sweeps_per_s_fused": B / ms * 1e3, "ms_full_outputs": ms_full, "sweeps_per_s_full": B / ms_full * 1e3,
                     "alg_GBps_full": alg_full * B / ms_full / 1e6, "alg_bytes_fused": alg, "alg_bytes_full": alg_full})
    # ---- C4 rocket OC adjoint gradient, H=100, B=8192 per GPU
    s = systems.rocket_oc_adjoint(0.1)
    B, H = 8192, 100
    q = np.array([np.cos(0.75), 0, 0, np.sin(0.75)])
    x0 = torch.tensor([10, -8, 5., -.1, 0, 0, *q, 0, 0, 0], dtype=torch.float64).repeat(B, 1)
    x0 = (x0 + 0.5 * torch.randn(x0.shape, dtype=torch.float64, generator=g)).to(dev)
    U = (torch.tensor([10., 0, 0], dtype=torch.float64) + torch.randn((B, H, 3), dtype=torch.float64, generator=g)).to(dev)
    th = torch.zeros((1, 1), dtype=torch.float64, device=dev)
    ms = timeit(lambda: s.rollout_costate(x0, th, U, want_dHu=True))
    alg = 8 * (13 + H * 3 + (H + 1) * 13 + H * 13 + H * 3 + 1)
    rows.append({"config": "C4 rocket OC adjoint n=13 m=3 H=100 B=8192/GPU", "ms": ms, "sweeps_per_s": B / ms * 1e3,
                 "alg_bytes": alg, "alg_GBps": alg * B / ms / 1e6})
    # ---- C5 quadrotor SysID, H=100, B=32768 per GPU
    s = systems.quadrotor_sysid(0.1)
    B, H = 32768, 100
    inputs = (20 * torch.rand((B, H, 4), dtype=torch.float64, generator=g) - 10).to(dev)
    x0 = torch.tensor([-8, -6, 9., 0, 0, 0, 1, 0, 0, 0, 0, 0, 0], dtype=torch.float64).repeat(B, 1).to(dev)
    th_true = torch.tensor([1, 1, 1, 1, 0.4], dtype=torch.float64, device=dev)
    Xobs = s.step(inputs, None, th_true, x0=x0, want_traj=True)["X"]
    th = th_true + 0.3 * (2 * torch.rand(5, dtype=torch.float64, generator=g) - 1).to(dev)
    ms = timeit(lambda: s.step(inputs, Xobs, th))
    alg = 8 * (H * 4 + (H + 1) * 13 + 5 + 1)
    rows.append({"config": "C5 quadrotor SysID n=13 r=5 H=100 B=32768/GPU (fused loss)", "ms": ms,
                 "sweeps_per_s": B / ms * 1e3, "alg_bytes": alg, "alg_GBps": alg * B / ms / 1e6})
    # ---- C5 variants: column-group size x block size of pdp_k_sens_fwd
    from JinEnv import JinEnv
    from pontryagin_differentiable_programming_b200 import engine
    env = JinEnv.Quadrotor()
    env.initDyn(c=0.01)
    for gc, blk in ((12, 64), (12, 128), (3, 64), (2, 64), (1, 64), (1, 128), (2, 128)):
        sv = engine.SysIDSystem(env.X, env.U, env.dyn_auxvar, env.X + 0.1 * env.f, max_group_cols=gc, block=blk)
        ms = timeit(lambda: sv.step(inputs, Xobs, th))
        rows.append({"config": "C5 variant group_cols=%d block=%d" % (gc, blk), "ms": ms, "sweeps_per_s": B / ms * 1e3,
                     "alg_GBps": alg * B / ms / 1e6})
    for r_ in rows:
        print(json.dumps(r_))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
