"""Compile (CPU, here) and time (GPU box) variants of the forward-sensitivity and rollout kernels at the C5 / C2 / C4 / C3 shapes.

  python tools/tune_sens.py --build      # cross-compile every variant in-tree
  python tools/tune_sens.py --run        # on the GPU: CUDA-event timings -> gpurun_out/tune_sens.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SYSID_VARIANTS = [dict(max_group_cols=3), dict(max_group_cols=2), dict(max_group_cols=5)]
CP_VARIANTS = [dict(max_group_cols=12), dict(max_group_cols=3)]


def sysid(**kw):
    from JinEnv import JinEnv
    from pontryagin_differentiable_programming_b200 import engine
    env = JinEnv.Quadrotor()
    env.initDyn(c=0.01)
    return engine.SysIDSystem(env.X, env.U, env.dyn_auxvar, env.X + 0.1 * env.f, **kw)


def cartpole(**kw):
    import numpy as np
    from JinEnv import JinEnv
    from pontryagin_differentiable_programming_b200 import engine, systems
    from pontryagin_differentiable_programming_b200.symbolic import SX
    env = JinEnv.CartPole()
    env.initDyn(mc=0.1, mp=0.1, l=1)
    env.initCost(wx=0.1, wq=0.6, wdx=0.1, wdq=0.1, wu=0.3)
    t = SX.sym('t')
    pol, th = systems.lagrange_policy(1, np.linspace(0, 50, 6), t)
    return engine.CPSystem(env.X, env.U, th, env.X + 0.05 * env.f, pol, t, env.path_cost, env.final_cost, **kw)


def timeit(torch, fn, iters=20, warm=3):
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--run", action="store_true")
    args = ap.parse_args()
    if args.build:
        for v in SYSID_VARIANTS:
            print("sysid", v, sysid(**v).module_path)
        for v in CP_VARIANTS:
            print("cartpole", v, cartpole(**v).module_path)
    if args.run:
        import numpy as np
        import torch
        import bench
        from pontryagin_differentiable_programming_b200 import systems
        dev = torch.device("cuda:0")
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
        rows = []
        # ---- C5
        B, H = 32768, 100
        inputs, x0, th_true, theta = bench.synth_sysid(B, H, seed=(5, 0))
        inputs, x0, th_true, theta = t(inputs), t(x0), t(th_true), t(theta)
        ref = None
        for v in SYSID_VARIANTS:
            s = sysid(**v)
            Xobs = s.step(inputs, None, th_true, x0=x0, want_traj=True)["X"]
            ldp = s.step(inputs, Xobs, theta)["loss_dp"]
            if ref is None:
                ref = ldp.clone()
            err = float((ldp - ref).abs().max() / ref.abs().max())
            ms = timeit(torch, lambda: s.step(inputs, Xobs, theta))
            ms_full = timeit(torch, lambda: s.step(inputs, Xobs, theta, want_traj=True, want_sens=True), iters=5)
            rows.append({"kernel": "sens C5", "variant": v, "groups": len(s.src.groups), "ms_fused": ms,
                         "ms_full_outputs": ms_full, "alg_GBps": bench.alg_bytes_sysid(13, 4, 5, H) * B / ms / 1e6, "rel_diff_vs_first": err})
            print(json.dumps(rows[-1]), flush=True)
        # ---- C2
        B, H = 4096, 50
        x0c, thc = bench.synth_cartpole(B, 6, seed=(2, 0))
        x0c, thc = t(x0c), t(thc)
        for v in CP_VARIANTS:
            s = cartpole(**v)
            ms = timeit(torch, lambda: s.step(x0c, H, thc))
            ms_full = timeit(torch, lambda: s.step(x0c, H, thc, want_traj=True, want_sens=True))
            rows.append({"kernel": "sens C2 poly", "variant": v, "groups": len(s.src.groups), "ms_fused": ms, "ms_full_outputs": ms_full,
                         "alg_GBps_full": bench.alg_bytes_cp_full(4, 1, 6, H) * B / ms_full / 1e6})
            print(json.dumps(rows[-1]), flush=True)
        # ---- rollout / costate kernel: C3 and C4 shapes
        s3 = systems.quadrotor_irl(0.1)
        d3 = [t(a) for a in bench.synth_quadrotor(16384, 50, seed=(0, 0))]
        ms = timeit(torch, lambda: s3.rollout_costate(d3[0], d3[1], d3[2]))
        rows.append({"kernel": "rollout C3", "ms": ms, "alg_GBps": bench.alg_bytes_rollout(13, 4, 9, 50) * 16384 / ms / 1e6})
        print(json.dumps(rows[-1]), flush=True)
        s4 = systems.rocket_oc_adjoint(0.1)
        x04, U4 = bench.synth_rocket(8192, 100, seed=(4, 0))
        x04, U4 = t(x04), t(U4)
        th4 = torch.zeros((1, 1), dtype=torch.float64, device=dev)
        for Bq in (8192, 65536):
            xq, Uq = x04.repeat(Bq // 8192, 1), U4.repeat(Bq // 8192, 1, 1)
            ms = timeit(torch, lambda: s4.rollout_costate(xq, th4, Uq, want_dHu=True))
            rows.append({"kernel": "rollout C4 B=%d" % Bq, "ms": ms, "alg_GBps": bench.alg_bytes_adjoint(13, 3, 100) * Bq / ms / 1e6})
            print(json.dumps(rows[-1]), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "tune_sens.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
