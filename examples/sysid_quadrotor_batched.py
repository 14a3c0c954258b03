"""Batched quadrotor system identification on the GPU(s) (config C5 of BASELINE.json): the loop of reference
Examples/SysID/quadrotor/uav_PDP.py:33-51 with B random-input trajectories sharded over the ranks and ONE
all-reduce of (sum loss, sum dp, count) per iteration.

  python examples/sysid_quadrotor_batched.py --traj 32768 --iters 200
  torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/sysid_quadrotor_batched.py --traj 262144
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pontryagin_differentiable_programming_b200 import distributed, irl, systems  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--traj", type=int, default=32768, help="global number of recorded trajectories")
    ap.add_argument("--horizon", type=int, default=100)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--lr", type=float, default=1e-4, help="initial step; halved whenever the loss rises")
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sys_ = systems.quadrotor_sysid(0.1)
    lo, hi = distributed.shard_bounds(args.traj, rank, world)
    B = hi - lo
    gen = torch.Generator().manual_seed(1000 + rank)
    inputs = (20 * torch.rand((B, args.horizon, 4), dtype=torch.float64, generator=gen) - 10).to(dev)   # U(-10, 10)
    x0 = torch.tensor([-8, -6, 9., 0, 0, 0, 1, 0, 0, 0, 0, 0, 0], dtype=torch.float64).repeat(B, 1).to(dev)
    theta_true = torch.tensor([1, 1, 1, 1, 0.4], dtype=torch.float64, device=dev)
    states = sys_.step(inputs, None, theta_true, x0=x0, want_traj=True)["X"]           # the "recorded" data
    trainer = irl.SysIDTrainer(sys_, inputs, states, args.lr)
    theta = theta_true + torch.tensor([0.1, -0.08, 0.06, 0.08, -0.04], dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prev = float("inf")
    for k in range(args.iters):
        loss, dp = trainer.gradient(theta)               # same on every rank (all-reduced), so is the step control
        lv = loss.item()
        if not (lv <= prev * 1.0001):                     # the plain GD of the reference diverges for long horizons:
            trainer.lr *= 0.5                             # halve the step and retry from the last good iterate
            theta = good
            continue
        prev, good = lv, theta
        theta = theta - trainer.lr * dp
        if rank == 0 and (k % 50 == 0 or k == args.iters - 1):
            print("iter %4d  loss %.6e  |theta - true| %.5f  lr %.2e" % (k, lv, (good - theta_true).norm().item(), trainer.lr))
    torch.cuda.synchronize()
    if rank == 0:
        dt = (time.perf_counter() - t0) / args.iters
        print("%.3f ms per iteration, %.1f M trajectory-sweeps/s over %d GPU(s)" % (dt * 1e3, args.traj / dt / 1e6, world))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
