"""Compile (CPU, here) and check + time (GPU box) variants of the rollout / costate kernel at the C3 and C4 shapes.

  python tools/tune_rollout.py --build
  python tools/tune_rollout.py --run        # -> gpurun_out/tune_rollout.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = [dict(), dict(rollout_tma=1, tma_chunk=2), dict(rollout_tma=1, tma_chunk=3), dict(rollout_tma=1, tma_chunk=4), dict(rollout_tma=1, tma_chunk=6)]


def quad(**kw):
    from tools.tune_aux_lqr import make
    return make(**kw)


def rocket(**kw):
    from JinEnv import JinEnv
    from pontryagin_differentiable_programming_b200 import engine
    from pontryagin_differentiable_programming_b200.symbolic import SX
    env = JinEnv.Rocket()
    env.initDyn(Jx=0.5, Jy=1., Jz=1., mass=1., l=1.)
    env.initCost(wr=1, wv=1, wtilt=50, ww=1, wsidethrust=1, wthrust=0.4)
    return engine.OCSystem(env.X, env.U, SX.sym('unused_auxvar'), env.X + 0.1 * env.f, env.path_cost, env.final_cost, **kw)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--run", action="store_true")
    args = ap.parse_args()
    if args.build:
        for v in VARIANTS:
            print(v, quad(**v).module_path, rocket(**v).module_path)
    if args.run:
        import numpy as np
        import torch
        import bench
        from tools.tune_sens import timeit
        dev = torch.device("cuda:0")
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
        rows = []
        d3 = [t(a) for a in bench.synth_quadrotor(16384, 50, seed=(0, 0))]
        x04, U4 = bench.synth_rocket(8192, 100, seed=(4, 0))
        x04, U4 = t(x04), t(U4)
        th4 = torch.zeros((1, 1), dtype=torch.float64, device=dev)
        ref3 = ref4 = None
        for v in VARIANTS:
            s3, s4 = quad(**v), rocket(**v)
            # odd sizes first (ragged tails, both address parities), then the full shapes
            o3s = s3.rollout_costate(d3[0][:37].contiguous(), d3[1][:37].contiguous(), d3[2][:37, :19].contiguous(), want_dHu=True)
            o3 = s3.rollout_costate(d3[0], d3[1], d3[2], want_dHu=True)
            o4 = s4.rollout_costate(x04, th4, U4, want_dHu=True)
            torch.cuda.synchronize()
            if ref3 is None:
                ref3, ref4, ref3s = o3, o4, o3s
            same = all(torch.equal(o3[k], ref3[k]) for k in ("X", "Lam", "cost", "dHu")) and \
                all(torch.equal(o4[k], ref4[k]) for k in ("X", "Lam", "cost", "dHu")) and \
                all(torch.equal(o3s[k], ref3s[k]) for k in ("X", "Lam", "cost", "dHu"))
            ms3 = timeit(torch, lambda: s3.rollout_costate(d3[0], d3[1], d3[2]))
            ms4 = timeit(torch, lambda: s4.rollout_costate(x04, th4, U4, want_dHu=True))
            rows.append({"variant": v, "identical_to_shipped": bool(same), "c3_rollout_ms": ms3, "c4_adjoint_ms": ms4,
                         "c3_alg_GBps": bench.alg_bytes_rollout(13, 4, 9, 50) * 16384 / ms3 / 1e6,
                         "c4_alg_GBps": bench.alg_bytes_adjoint(13, 3, 100) * 8192 / ms4 / 1e6})
            print(json.dumps(rows[-1]), flush=True)
        json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "tune_rollout.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
