"""Multi-GPU checks of the sharded outer loops, run under torchrun (one rank per GPU, NCCL):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/multigpu_check.py

(a) SysID: the sharded gradient (per-rank sweep kernel -> batch reduction kernel -> ONE all-reduce) equals the
    single-GPU gradient of the whole batch; the CUDA-graph iteration (all-reduce captured inside) equals the eager one,
    for gradient descent and Adam.
(b) IRL: the graph-captured iteration of a sharded demonstration batch equals the eager sharded iteration.
Prints one JSON line from rank 0; exit code != 0 on any mismatch."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pontryagin_differentiable_programming_b200 import distributed, irl, systems  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    report = {"world": world}

    def stage(msg):
        if rank == 0:
            print("[multigpu_check] %s (%.1f s)" % (msg, time.perf_counter() - t_start), file=sys.stderr, flush=True)
    t_start = time.perf_counter()

    # ---------------- (a) SysID, global batch 4096 + 3 (ragged split), H = 100
    Bg, H = 4099, 100
    inputs, x0, th_true, theta = bench.synth_sysid(Bg, H, seed=21)
    s = systems.quadrotor_sysid(0.1)
    Xobs = s.step(t(inputs), None, t(th_true), x0=t(x0), want_traj=True)["X"]
    full = s.step(t(inputs), Xobs, t(theta))["loss_dp"]
    loss_ref, dp_ref = full[:, 0].mean(), full[:, 1:].mean(dim=0)
    lo, hi = distributed.shard_bounds(Bg, rank, world)
    tr = irl.SysIDTrainer(s, t(inputs[lo:hi]), Xobs[lo:hi].contiguous(), lr=1e-5)
    loss, dp = tr.gradient(t(theta))
    e1 = max(float((loss - loss_ref).abs() / loss_ref.abs()), float((dp - dp_ref).abs().max() / dp_ref.abs().max()))
    report["sysid_sharded_vs_single_gpu_rel"] = e1
    stage("sharded SysID gradient ok")
    assert e1 < 1e-12, e1
    lr_gd = 0.01 / float(dp_ref.abs().max())      # random +-10 inputs over 100 steps make the loss (and dp) huge: scale the step
    for opt in ("gd", "adam"):
        eager = irl.SysIDTrainer(s, t(inputs[lo:hi]), Xobs[lo:hi].contiguous(), lr=lr_gd if opt == "gd" else 1e-3, optimizer=opt)
        graph = irl.SysIDTrainer(s, t(inputs[lo:hi]), Xobs[lo:hi].contiguous(), lr=lr_gd if opt == "gd" else 1e-3, optimizer=opt)
        th_e = th_g = t(theta)
        for k in range(6):
            le, th_e = eager.step(th_e)
            lg, th_g = graph.step_graph(th_g)
            lg, th_g = lg.clone(), th_g.clone()
            err = max(float((le - lg).abs() / le.abs()), float((th_e - th_g).abs().max()))
            assert err < 1e-12, (opt, k, err, float(le), float(lg))
        report["sysid_graph_vs_eager_%s" % opt] = err
        stage("SysID graph vs eager (%s) ok" % opt)
        report["sysid_loss_after_6_%s" % opt] = float(lg)
    # iterations per second of the captured iteration at the C5 per-GPU size
    B5 = 32768
    inputs5, x05, _, _ = bench.synth_sysid(B5, H, seed=(5, rank))
    X5 = s.step(t(inputs5), None, t(th_true), x0=t(x05), want_traj=True)["X"]
    tr5 = irl.SysIDTrainer(s, t(inputs5), X5, lr=1e-3, optimizer="adam")
    th = t(theta)
    for _ in range(5):
        th = tr5.step_graph(th)[1].clone()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_it = 200
    for _ in range(n_it):
        th = tr5.step_graph(th)[1]
    torch.cuda.synchronize(); dist.barrier()
    dt = (time.perf_counter() - t0) / n_it
    stage("C5 outer loop timed")
    report["c5_outer_loop_graph"] = {"global_batch": B5 * world, "iters_per_s": 1 / dt, "ms_per_iter": dt * 1e3,
                                     "traj_sweeps_per_s": B5 * world / dt}

    # ---------------- (b) IRL on the shipped quadrotor demos, replicated to 2 per rank with different theta paths
    g = np.load(os.path.join(ROOT, "tests", "golden", "k2_demos.npz"))
    Xd = t(np.stack([g["quadrotor_%d_X" % i] for i in range(2)]))
    Ud = t(np.stack([g["quadrotor_%d_U" % i] for i in range(2)]))
    sl = slice(rank % 2, rank % 2 + 1) if world > 1 else slice(0, 2)
    oc = systems.quadrotor_irl(0.1)
    th0 = t(np.asarray(g["quadrotor_true_parameter"]).reshape(-1) * 1.2)
    eager, graph = irl.IRLTrainer(oc, Xd[sl], Ud[sl], 1e-4), irl.IRLTrainer(oc, Xd[sl], Ud[sl], 1e-4)
    th_e = th_g = th0
    for k in range(4):
        le, th_e = eager.step(th_e)
        stage("IRL eager step %d" % k)
        lg, th_g, resid = graph.step_graph(th_g, n_newton=12)
        lg, th_g = lg.clone(), th_g.clone()
        stage("IRL graph step %d" % k)
    err = max(float((le - lg).abs() / le.abs()), float((th_e - th_g).abs().max() / th_e.abs().max()))
    report["irl_graph_vs_eager_sharded"] = err
    report["irl_diagnostics"] = eager.diagnostics()
    assert err < 1e-6, err
    if rank == 0:
        print(json.dumps(report), flush=True)
    # captured NCCL collectives must be gone before the communicator is torn down (destroy_process_group hangs otherwise)
    for tr_ in (tr5, eager, graph):
        tr_.release_graph()
    del tr5, eager, graph
    import gc
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    import threading
    threading.Timer(60.0, lambda: os._exit(0)).start()      # belt and braces: never let a teardown problem hang the box
    dist.destroy_process_group()
    os._exit(0)


if __name__ == "__main__":
    main()
