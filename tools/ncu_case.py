"""One full-size launch of every hot-path kernel of every BASELINE config between cudaProfilerStart / Stop, for

  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_prof -f python tools/ncu_case.py

(C3: rollout/costate, backward Riccati, forward pass, batch reduction as four serial launches on 16 384 trajectories;
C5 / C2: the sensitivity kernel; C4: rollout/costate with dH/du).  Numbers printed under ncu are never bench values."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c3,c5,c4,c2")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    args = argparse.Namespace(batch=0, horizon=0, e2e_chunks=4)
    todo = []
    for c in a.configs.split(","):
        wl = bench.WORKLOADS[c](args, 0, 1, dev)
        if c == "c3":
            d, out = wl.d, wl.out

            def run(wl=wl, d=d, out=out):
                wl.sys.rollout_costate(d[0], d[1], d[2], status=wl.status, out=out)
                wl.sys.aux_lqr(out["X"], d[2], out["Lam"], d[1], status=wl.status, phase="backward")
                wl.sys.aux_lqr(out["X"], d[2], out["Lam"], d[1], Xref=d[3], Uref=d[4], status=wl.status, out=out, phase="forward")
                wl.finish(out["loss_dp"])
        else:
            run = wl.step
        for _ in range(2):
            run()
        todo.append(run)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for run in todo:
        run()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
