#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on B200.

metric : PDP sweeps/sec (fwd + aux-LQR bwd), quadrotor n_x=13 n_u=4 r=9 H=50, batch 16384 per GPU.
A "step" = one PDP sweep of the whole (per-rank) batch at given controls:
  pdp_k_rollout_costate (rollout + cost + costate recursion)  ->  pdp_k_aux_lqr (aux evaluation +
  Riccati sweep + aux forward pass: dX/dtheta, dU/dtheta, fused IRL loss/dp).
Inputs are synthetic (seed 0, float64).  `value` times the step (OCSystem.sweep -> C-ABI pdp_sweep)
with inputs resident in HBM; `e2e` times the C-ABI host-buffer call (pdp_sweep_host) incl. H2D of the inputs and D2H
of (loss, dp).  `roofline` is for the dominant kernel pdp_k_aux_lqr_bwd (CUDA events around that launch, live, in a
second loop that issues the three kernels serially);
`cpu_baseline` times the oracle port (reference-shaped NumPy loops) on the box's host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--horizon H]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_X, N_U, N_R = 13, 4, 9
TRUE_THETA = np.array([1, 1, 1, 1, 0.4, 1, 1, 5, 1.0])


def synth_quadrotor(B, H, seed=0):
    """Synthetic C3 batch: x0 positions U([-8,8]^2 x [3,9]), random small-angle attitude,
    theta = true + U(-0.4,0.4) clipped >= 0.1, U = 2.5 + 0.5 N(0,1) (hover thrust), demos = noisy copies."""
    rng = np.random.default_rng(seed)
    x0 = np.zeros((B, N_X))
    x0[:, 0:2] = rng.uniform(-8, 8, (B, 2))
    x0[:, 2] = rng.uniform(3, 9, B)
    ang = rng.uniform(-0.5, 0.5, B)
    axis = rng.standard_normal((B, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    x0[:, 6] = np.cos(ang / 2)
    x0[:, 7:10] = np.sin(ang / 2)[:, None] * axis
    theta = np.maximum(TRUE_THETA + rng.uniform(-0.4, 0.4, (B, N_R)), 0.1)
    U = 2.5 + 0.5 * rng.standard_normal((B, H, N_U))
    Xref = rng.standard_normal((B, H + 1, N_X))
    Uref = 2.5 + 0.5 * rng.standard_normal((B, H, N_U))
    return x0, theta, U, Xref, Uref


def alg_bytes_aux_lqr(n, m, r, H):
    """Compulsory bytes of ONE trajectory through pdp_k_aux_lqr: read X,U,Lam,theta once,
    write dX,dU once, gain spill written once + read once."""
    return 8 * (((H + 1) * n + H * m + H * n + r) + ((H + 1) * n * r + H * m * r) + 2 * H * m * (n + r))


def alg_bytes_bwd(n, m, r, H):
    """pdp_k_aux_lqr_bwd: read X, U, Lam, theta once; write the gain spill (K_t|k_t) once."""
    return 8 * (((H + 1) * n + H * m + H * n + r) + H * m * (n + r))


def alg_bytes_fwd(n, m, r, H):
    """pdp_k_aux_lqr_fwd: read X, U, theta, the gain spill, Xref, Uref once; write dX, dU, (loss, dp) once."""
    return 8 * (((H + 1) * n + H * m + r) + H * m * (n + r) + ((H + 1) * n + H * m) + ((H + 1) * n * r + H * m * r) + r + 1)


def alg_flops_bwd(n, m, r, H):
    return H * (4 * n ** 3 + 6 * n * n * m + 4 * n * m * m + m ** 3 / 3 + 4 * n * n * r + 4 * n * m * r + 2 * m * m * r + 400)


def alg_bytes_sweep(n, m, r, H):
    """SURVEY 8(d) figure for the whole sweep (C3: 144 816 B)."""
    return 8 * ((n + r + H * m) + ((H + 1) * n + H * n + (H + 1) * n * r + H * m * r) + 2 * H * m * (n + r))


def alg_flops_sweep(n, m, r, H):
    back = 4 * n ** 3 + 6 * n * n * m + 4 * n * m * m + m ** 3 / 3 + 4 * n * n * r + 4 * n * m * r + 2 * m * m * r
    fwd = 2 * n * n * r + 4 * n * m * r
    return H * (back + fwd + 400)


def ncu_traffic(kernel):
    """DRAM bytes (read + write) per launch of ``kernel`` from the newest committed `ncu --set full` capture of this
    same command (profiles/*_ncu_traffic.json, C3 at 16 384 trajectories); None if no capture is committed."""
    for name in TRAFFIC_FILES:
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))
            return float(d[kernel]["dram_bytes_read"]) + float(d[kernel]["dram_bytes_write"]), d.get("source")
        except Exception:
            continue
    return None, None


TRAFFIC_FILES = ("r1m_ncu_traffic.json", "r1_final_ncu_traffic.json")     # newest first


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML by a polling thread DURING the timed region."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.maxclk = index, [], set(), None
        self._stop = threading.Event()
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.maxclk = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as exc:  # pragma: no cover
            self.nv, self.err = None, repr(exc)

    def _poll(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml unavailable: " + self.err]}
        self._stop.set()
        self.thread.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.maxclk,
                "samples": len(self.samples), "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------- CPU arms
_ORACLE = {}


def _oracle_oc():
    if "oc" not in _ORACLE:
        from oracle import envs, pdp_oracle
        oc = pdp_oracle.build_oc(envs.quadrotor(c=0.01, wthrust=0.1), 0.1)
        oc.diffPMP()
        _ORACLE["oc"] = (oc, pdp_oracle)
    return _ORACLE["oc"]


def _cpu_worker(args):
    x0, theta, U = args
    oc, po = _oracle_oc()
    t0 = time.perf_counter()
    for b in range(x0.shape[0]):
        po.pdp_sweep(oc, x0[b], U[b], theta[b])
    return time.perf_counter() - t0


def cpu_sweeps_per_s(H, per_core, cores):
    """Oracle port (reference-shaped per-step NumPy / lambdified loops), one process per host core."""
    import multiprocessing as mp
    _oracle_oc()  # build before forking so children inherit the lambdified functions
    x0, theta, U, _, _ = synth_quadrotor(per_core * cores, H, seed=1)
    chunks = [(x0[i::cores], theta[i::cores], U[i::cores]) for i in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        _cpu_worker(chunks[0])
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_cpu_worker, chunks)
    wall = time.perf_counter() - t0
    return per_core * cores / wall, wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    H = args.horizon
    per_core = args.ref_per_core
    for _ in range(min(args.warmup, 1)):
        cpu_sweeps_per_s(H, 1, cores)
    vals = []
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, wall = cpu_sweeps_per_s(H, per_core, cores)
        vals.append(v)
    total = time.perf_counter() - t_all
    value = float(np.mean(vals))
    sample = "%d trajectories of the C3 workload per step (%d per core x %d cores), oracle port" % (per_core * cores, per_core, cores)
    print(json.dumps({
        "impl": "reference", "metric": "PDP sweeps/sec (fwd+aux-LQR bwd)", "value": value, "unit": "sweeps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3 quadrotor IRL sweep n=13 m=4 r=9 H=%d" % H, "batch_per_step": per_core * cores},
        "cpu_baseline": {"value": value, "unit": "sweeps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from pontryagin_differentiable_programming_b200 import systems

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the PDP B200 engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout for the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    B, H = args.batch, args.horizon
    sys_ = systems.quadrotor_irl(0.1)
    n, m, r = sys_.n, sys_.m, sys_.r

    # the global batch is the concatenation of per-rank blocks of B trajectories, block g seeded with (0, g): every
    # GPU count sees the same trajectories in block g, and no rank has to materialise the other ranks' blocks
    host = [np.ascontiguousarray(a) for a in synth_quadrotor(B, H, seed=(0, rank))]
    pinned = [torch.from_numpy(a).pin_memory() for a in host]
    d_x0, d_th, d_U, d_Xr, d_Ur = [p.to(dev) for p in pinned]
    out = {"X": torch.empty((B, H + 1, n), dtype=torch.float64, device=dev),
           "Lam": torch.empty((B, H, n), dtype=torch.float64, device=dev),
           "cost": torch.empty((B,), dtype=torch.float64, device=dev),
           "dX": torch.empty((B, H + 1, n, r), dtype=torch.float64, device=dev),
           "dU": torch.empty((B, H, m, r), dtype=torch.float64, device=dev),
           "loss_dp": torch.empty((B, r + 1), dtype=torch.float64, device=dev)}
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step(ev=None):
        # identical to OCSystem.sweep, split in two so the dominant kernel can be bracketed by events
        ro = sys_.rollout_costate(d_x0, d_th, d_U, status=status, out=out)
        if ev is not None:
            ev[0].record(stream)
        sys_.aux_lqr(out["X"], d_U, out["Lam"], d_th, status=status, phase="backward")
        if ev is not None:
            ev[1].record(stream)
        sys_.aux_lqr(out["X"], d_U, out["Lam"], d_th, Xref=d_Xr, Uref=d_Ur, status=status, out=out, phase="forward")
        if ev is not None:
            ev[2].record(stream)
        return ro

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def sweep():
        # THE timed step: OCSystem.sweep -> C-ABI pdp_sweep (rollout/costate on the whole batch, then the aux-LQR phase
        # in sub-batches on the library's two internal streams so that bwd of one overlaps fwd of the other)
        sys_.sweep(d_x0, d_th, d_U, Xref=d_Xr, Uref=d_Ur, status=status, out=out)

    for _ in range(max(args.warmup, 3)):
        sweep()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for k in range(args.steps):
        sweep()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    # ---- the same work as three serial launches, each bracketed by CUDA events: per-kernel durations for `roofline`
    for _ in range(2):
        step()
    kev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(4)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        kev[k][3].record(stream)
        step(kev[k])
    barrier()
    kr_ms = float(np.mean([d.elapsed_time(a) for a, _, _, d in kev]))     # pdp_k_rollout_costate
    k_ms = float(np.mean([a.elapsed_time(b) for a, b, _, _ in kev]))      # pdp_k_aux_lqr_bwd (dominant kernel)
    kf_ms = float(np.mean([b.elapsed_time(c) for _, b, c, _ in kev]))     # pdp_k_aux_lqr_fwd
    parts = 4 if B >= 16384 else (2 if B >= 8192 else 1)                   # pdp_sweep's automatic split

    # ---- parity subset against the oracle every run (first 4 trajectories of rank 0)
    parity = None
    if rank == 0:
        oc, po = _oracle_oc()
        worst = 0.0
        for b in range(4):
            X, L, cost, dX, dU = po.pdp_sweep(oc, host[0][b], host[2][b], host[1][b])
            for nm_, ref in (("X", X), ("Lam", L), ("dX", dX), ("dU", dU)):
                got = out[nm_][b].cpu().numpy()
                worst = max(worst, float(np.max(np.abs(got - ref)) / max(1e-300, np.max(np.abs(ref)))))
        parity = worst

    # ---- e2e: the public host-buffer API (OCSystem.sweep_host -> C-ABI pdp_sweep_host per sub-batch on two
    #      streams): H2D of every input + all kernels + D2H of (loss, dp, cost) inside the timed region
    ldp_host = torch.empty((B, r + 1), dtype=torch.float64).pin_memory()
    cost_host = torch.empty((B,), dtype=torch.float64).pin_memory()

    def e2e_step():
        sys_.sweep_host(pinned[0], pinned[1], pinned[2], pinned[3], pinned[4], ldp_host, cost_h=cost_host,
                        keep_dtraj=True, n_chunks=args.e2e_chunks, device=dev)

    for _ in range(3):
        e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for _ in range(args.steps):
        e2e_step()
    f1.record(stream)
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    e2e_ok = bool(torch.allclose(ldp_host.to(dev), out["loss_dp"], rtol=1e-12, atol=0))

    times = torch.tensor([ms_total, e2e_ms, k_ms, kf_ms, kr_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, k_ms, kf_ms, kr_ms = (float(v) for v in times.cpu())
    nbad = int((status != 0).sum().item())

    if rank == 0:
        value = B * world * args.steps / (ms_total * 1e-3)
        e2e_val = B * world * args.steps / (e2e_ms * 1e-3)
        peak, peak_src = measured_peaks()
        kbytes = alg_bytes_bwd(n, m, r, H) * B
        fbytes = alg_bytes_fwd(n, m, r, H) * B
        achieved = kbytes / (k_ms * 1e-3) / 1e9
        h2d = sum(int(p.numel()) * 8 for p in pinned)
        d2h = int(ldp_host.numel() + cost_host.numel()) * 8
        traffic, traffic_src = ncu_traffic("pdp_k_aux_lqr_bwd") if B == 16384 and H == 50 else (None, None)
        line = {
            "metric": "PDP sweeps/sec (fwd+aux-LQR bwd)", "value": value, "unit": "sweeps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C3 quadrotor IRL PDP sweep n_x=13 n_u=4 r=9 H=%d" % H, "batch_per_gpu": B,
                       "global_batch": B * world, "parallelism": "batch-sharded x%d, no data-path collective" % world,
                       "step": "OCSystem.sweep -> pdp_sweep: 1 rollout/costate launch + %d sub-batches x (bwd, fwd) on two streams" % parts,
                       "outputs": "X, Lam, cost, dX/dtheta, dU/dtheta, fused (loss, dp)",
                       "l2": "per-step working set %.2f GB per GPU >> 126 MB L2 (no flush needed)"
                             % ((alg_bytes_sweep(n, m, r, H) * B) / 1e9),
                       "parity_max_rel_err_vs_oracle_first4": parity, "status_flagged_trajectories": nbad,
                       "status_note": "flag bit 1 = Quu not positive definite: expected for random (non-optimal) controls, "
                                      "the sweep is still the reference's algebra (parity above); 0 at OC optima (tests)",
                       "e2e_matches_device_path": e2e_ok},
            "roofline": {"kernel": "pdp_k_aux_lqr_bwd", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "alg_bytes_per_launch": kbytes, "kernel_ms": k_ms,
                         "kernel_share_of_step": k_ms / (kr_ms + k_ms + kf_ms),
                         "kernel_timing": "three serial launches bracketed by CUDA events (rollout %.4f / bwd %.4f / fwd %.4f ms); "
                                          "the timed step overlaps bwd and fwd of different sub-batches, so it is shorter "
                                          "than their sum" % (kr_ms, k_ms, kf_ms),
                         "fp64_tflops_alg_bwd": alg_flops_bwd(n, m, r, H) * B / (k_ms * 1e-3) / 1e12,
                         "fp64_peak_tflops_nominal": 37.0,
                         "second_kernel": {"kernel": "pdp_k_aux_lqr_fwd", "kernel_ms": kf_ms, "alg_bytes_per_launch": fbytes,
                                           "achieved": fbytes / (kf_ms * 1e-3) / 1e9,
                                           "frac": fbytes / (kf_ms * 1e-3) / 1e9 / peak},
                         "sweep_alg_bytes": alg_bytes_sweep(n, m, r, H) * B,
                         "sweep_achieved_GBps": alg_bytes_sweep(n, m, r, H) * B / (ms_total / args.steps * 1e-3) / 1e9},
            "e2e": {"value": e2e_val, "unit": "sweeps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "OCSystem.sweep_host -> pdp_sweep_host (C ABI, pinned host buffers), %d sub-batches on 2 streams" % args.e2e_chunks},
            "gpu_launches": (1 + 2 * parts) * args.steps, "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            per_core = args.ref_per_core
            v, wall = cpu_sweeps_per_s(H, per_core, cores)
            line["cpu_baseline"] = {"value": v, "unit": "sweeps/s", "cores": cores, "kind": "port",
                                    "sample": "%d trajectories of the same workload (%d per core), %.1f s wall"
                                              % (per_core * cores, per_core, wall)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16384, help="trajectories per GPU")
    ap.add_argument("--horizon", type=int, default=50)
    ap.add_argument("--ref-per-core", type=int, default=48,
                    help="trajectories per host core in one CPU-baseline step (about 2 s of work per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=4)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
