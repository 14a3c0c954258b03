"""Experiment: does cutting the sweep into sub-batches (bwd -> fwd of one sub-batch back to back, a few sub-batches in
flight on separate streams) let the forward kernel read the gain spill from L2 instead of HBM?

  python tools/subbatch_l2.py            # on the GPU box; writes gpurun_out/subbatch_l2.json
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tools.tune_aux_lqr import make  # noqa: E402


PARTS = ((1, 1), (2, 2), (3, 2), (4, 2), (5, 2), (6, 2), (7, 2), (3, 3), (6, 3), (8, 2))     # (sub-batches, streams)


def main():
    dev = torch.device("cuda:0")
    B, H = 16384, 50
    x0, th, U, Xr, Ur = [torch.as_tensor(np.ascontiguousarray(a), device=dev) for a in bench.synth_quadrotor(B, H)]
    n, m, r = 13, 4, 9
    out = {"X": torch.empty((B, H + 1, n), dtype=torch.float64, device=dev), "Lam": torch.empty((B, H, n), dtype=torch.float64, device=dev),
           "cost": torch.empty((B,), dtype=torch.float64, device=dev), "dX": torch.empty((B, H + 1, n, r), dtype=torch.float64, device=dev),
           "dU": torch.empty((B, H, m, r), dtype=torch.float64, device=dev), "loss_dp": torch.empty((B, r + 1), dtype=torch.float64, device=dev)}
    rows = []
    ref = None
    for parts, nslots in PARTS:
        systems_ = [make() for _ in range(nslots)]
        for s_ in systems_:
            s_._handle = None                                  # separate handles -> separate workspaces
        streams = [torch.cuda.Stream(device=dev) for _ in range(nslots)]
        bounds = [(B * i) // parts for i in range(parts + 1)]

        def sweep():
            cur = torch.cuda.current_stream(dev)
            # the rollout kernel is latency-bound (0.1 ms whatever the batch): run it once for the whole batch
            systems_[0].rollout_costate(x0, th, U, out=out)
            ev = torch.cuda.Event()
            ev.record(cur)
            for i in range(parts):
                lo, hi = bounds[i], bounds[i + 1]
                k = i % nslots
                st, s_ = streams[k], systems_[k]
                if i < nslots:
                    st.wait_event(ev)
                with torch.cuda.stream(st):
                    o = {kk: v[lo:hi] for kk, v in out.items()}
                    s_.aux_lqr(o["X"], U[lo:hi], o["Lam"], th[lo:hi], phase="backward")
                    s_.aux_lqr(o["X"], U[lo:hi], o["Lam"], th[lo:hi], Xref=Xr[lo:hi], Uref=Ur[lo:hi], out=o, phase="forward")
            for st in streams:
                e2 = torch.cuda.Event()
                e2.record(st)
                cur.wait_event(e2)

        for _ in range(3):
            sweep()
        torch.cuda.synchronize()
        # replay as ONE CUDA graph: the Python / ctypes launch cost (3 launches per part) must not decide the outcome
        cap = torch.cuda.Stream(device=dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(cap):
            with torch.cuda.graph(graph, stream=cap):
                sweep()
        torch.cuda.synchronize()
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        dx = out["dX"][::257].clone()
        if ref is None:
            ref = dx
        rows.append({"parts": parts, "streams": nslots, "ms_per_sweep": ms, "sweeps_per_s": B / ms * 1e3,
                     "max_abs_diff_vs_unsplit": float((dx - ref).abs().max())})
        print(json.dumps(rows[-1]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "subbatch_l2.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
