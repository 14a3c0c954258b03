"""Compile (CPU, here) and time (GPU box) variants of the quadrotor OC module: chunk x warps/block x min blocks.

  python tools/tune_aux_lqr.py --build          # cross-compile every variant in-tree (prints regs / smem)
  python tools/tune_aux_lqr.py --run            # on the GPU: time pdp_k_aux_lqr for every variant
"""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (chunk, bwd warps/block, bwd min blocks, fwd warps/block, fwd min blocks)
# (chunk, bwd warps/block, bwd min blocks, fwd warps/block, fwd min blocks, keep [F|G] slots in registers A->C)
# (chunk, bwd warps/block, bwd min blocks, fwd warps/block, fwd min blocks, keep_fg, fast_rcp, early_solve)
# ... + (fwd_pack = trajectories per warp of the forward kernel, fwd_chunk)
VARIANTS = [(17, 4, 3, 1, 8, True, True, True, 3, 10), (17, 4, 3, 1, 6, True, True, True, 3, 10),
            (17, 4, 3, 2, 4, True, True, True, 3, 10), (17, 4, 3, 4, 2, True, True, True, 3, 10),
            (17, 4, 3, 1, 12, True, True, True, 3, 10), (17, 4, 3, 4, 3, True, True, True, 1, 17)]


def make(ch, wpb, mb, wpbf=4, mbf=1, kf=True, frcp=False, early=False, fpack=0, fchunk=0, verbose=False):
    from JinEnv import JinEnv
    from pontryagin_differentiable_programming_b200 import engine
    from pontryagin_differentiable_programming_b200.symbolic import vertcat
    env = JinEnv.Quadrotor()
    env.initDyn(c=0.01)
    env.initCost(wthrust=0.1)
    return engine.OCSystem(env.X, env.U, vertcat(env.dyn_auxvar, env.cost_auxvar), env.X + 0.1 * env.f, env.path_cost,
                           env.final_cost, chunk=ch, warps_per_block=wpb, min_blocks=mb, fwd_warps_per_block=wpbf,
                           fwd_min_blocks=mbf, keep_fg=kf, fast_rcp=frcp, early_solve=early, verbose=verbose,
                           fwd_pack=fpack, fwd_chunk=fchunk)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--run", action="store_true")
    ap.add_argument("--batch", type=int, default=16384)
    args = ap.parse_args()
    if args.build:
        for v in VARIANTS:
            print("== variant", v)
            make(*v, verbose=True)
    if args.run:
        import numpy as np
        import torch
        import bench
        dev = torch.device("cuda:0")
        B, H = args.batch, 50
        x0, theta, U, Xr, Ur = [torch.as_tensor(a, device=dev) for a in bench.synth_quadrotor(B, H)]
        rows = []
        for v in VARIANTS:
            s = make(*v)
            ro = s.rollout_costate(x0, theta, U)
            out = {"dX": torch.empty((B, H + 1, 13, 9), dtype=torch.float64, device=dev),
                   "dU": torch.empty((B, H, 4, 9), dtype=torch.float64, device=dev),
                   "loss_dp": torch.empty((B, 10), dtype=torch.float64, device=dev)}
            ms = {}
            for phase in ("backward", "forward"):
                for _ in range(3):
                    s.aux_lqr(ro["X"], U, ro["Lam"], theta, Xref=Xr, Uref=Ur, out=out, phase=phase)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    s.aux_lqr(ro["X"], U, ro["Lam"], theta, Xref=Xr, Uref=Ur, out=out, phase=phase)
                e1.record()
                torch.cuda.synchronize()
                ms[phase] = e0.elapsed_time(e1) / 10
            rows.append({"chunk": v[0], "wpb": v[1], "minb": v[2], "wpbf": v[3], "minbf": v[4], "keep_fg": v[5], "fast_rcp": v[6], "early_solve": v[7], "fwd_pack": v[8], "fwd_chunk": v[9], "bwd_ms": ms["backward"],
                         "fwd_ms": ms["forward"], "sweeps_per_s": B / (ms["backward"] + ms["forward"]) * 1e3})
            print(json.dumps(rows[-1]), flush=True)
            del out
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "tune_aux_lqr.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
