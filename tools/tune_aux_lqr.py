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

# every variant: keyword overrides of BASE (= the shipped configuration of engine.OCSystem)
BASE = dict(chunk=8, warps_per_block=1, min_blocks=8, fwd_warps_per_block=1, fwd_min_blocks=8, keep_fg=True,
            fast_rcp=True, early_solve=True, fwd_pack=3, fwd_chunk=10, bwd_pack=2, fwd_vec=-1, prefetch=2, prefetch_dist=2,
            inline_eval=-1, h_group=1, prefetch_l1_lead=0)
V1 = dict(bwd_pack=1, chunk=17, warps_per_block=4, min_blocks=3, keep_fg=True)     # one trajectory per warp
VARIANTS = [
    V1,
    {},                                                                    # shipped: two trajectories per warp
]
# Measured and removed in round 2 (profiles/r2a_tune_pending_variants.json, B200, C3 at 16 384 trajectories; shipped path
# rollout 0.101 / bwd 0.650 / fwd 0.403 ms): cp.async staging of the backward kernel's chunk rows (bwd 0.757), of the
# forward kernel's chunk rows (fwd 0.605), multi-warp rollout with 2 / 4 warps per 32 trajectories (0.120 / 0.178),
# backward + forward of a warp's trajectories fused in one kernel (1.096 vs 1.054 for the two launches).  Streaming
# stores of dX / dU (fwd 0.400) were adopted unconditionally.
# Also measured and removed (profiles/r2f_tune_lean_backward.json): a register-lean backward step (columns of Z streamed from
# the staging tile, [F|G] re-read from the chunk buffer) -- 0.667 ms at the same 8 warps per SM (+40 shared-memory wavefronts
# per step), and SLOWER with every extra resident warp the smaller register budget allows: 10 warps/SM 0.770, 12: 0.957,
# 14: 1.232, 16: 1.369 ms.
# TMA-staged chunk rows (lane 0 brings the evaluation's rows in with one cp.async.bulk per array and trajectory, one chunk
# ahead; profiles/r2t_tune_tma_staged_chunk_rows.json, r2w_sweep_pipeline_fwd_variants.json): forward kernel alone 0.3975 ->
# 0.3805 ms, but its slots (+8 KB per warp) cost the co-residency the pipelined sweep lives on: whole sweep 1.068 -> 1.082 ms;
# backward kernel 0.6058 -> 0.6054 ms (its evaluator's loads are not a limiter).  Both removed.  The kernel is not latency-bound: two warps per scheduler already saturate what the shared-memory
# data path and the FP64 pipe deliver together; more warps only shrink the chunk (fewer evaluation lanes) and add spills.


def make(verbose=False, **kw):
    from JinEnv import JinEnv
    from pontryagin_differentiable_programming_b200 import engine
    from pontryagin_differentiable_programming_b200.symbolic import vertcat
    cfg = dict(BASE)
    cfg.update(kw)
    env = JinEnv.Quadrotor()
    env.initDyn(c=0.01)
    env.initCost(wthrust=0.1)
    return engine.OCSystem(env.X, env.U, vertcat(env.dyn_auxvar, env.cost_auxvar), env.X + 0.1 * env.f, env.path_cost,
                           env.final_cost, verbose=verbose, **cfg)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--run", action="store_true")
    ap.add_argument("--batch", type=int, default=16384)
    args = ap.parse_args()
    if args.build:
        for v in VARIANTS:
            print("== variant", v)
            make(verbose=True, **v)
    if args.run:
        import numpy as np
        import torch
        import bench
        dev = torch.device("cuda:0")
        B, H = args.batch, 50
        x0, theta, U, Xr, Ur = [torch.as_tensor(a, device=dev) for a in bench.synth_quadrotor(B, H)]
        rows = []
        ref_dx = None
        for v in VARIANTS:
            s = make(**v)
            ro = s.rollout_costate(x0, theta, U)
            out = {"dX": torch.empty((B, H + 1, 13, 9), dtype=torch.float64, device=dev),
                   "dU": torch.empty((B, H, 4, 9), dtype=torch.float64, device=dev),
                   "loss_dp": torch.empty((B, 10), dtype=torch.float64, device=dev)}
            ms = {}
            for _ in range(3):
                s.rollout_costate(x0, theta, U, out=ro)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                s.rollout_costate(x0, theta, U, out=ro)
            e1.record()
            torch.cuda.synchronize()
            ms["rollout"] = e0.elapsed_time(e1) / 10
            for phase in ("backward", "forward", "both"):     # "both": one call (the fused kernel where the module has one)
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
            dx = out["dX"][:256].clone()
            if ref_dx is None:
                ref_dx = dx
            err = float((dx - ref_dx).abs().max() / ref_dx.abs().max())       # parity against the first variant
            row = dict(BASE)
            row.update(v)
            row.update({"rollout_ms": ms["rollout"], "bwd_ms": ms["backward"], "fwd_ms": ms["forward"], "both_ms": ms["both"],
                        "rel_diff_vs_first": err,
                        "sweeps_per_s": B / (ms["rollout"] + ms["backward"] + ms["forward"]) * 1e3})
            rows.append(row)
            print(json.dumps(rows[-1]), flush=True)
            del out
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "tune_aux_lqr.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
