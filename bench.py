#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on B200, for each of its GPU configs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3|c2|c4|c5] [--impl reference]

--config c3 (default, the headline)  quadrotor IRL PDP sweep, n=13 m=4 r=9 H=50, 16 384 trajectories per GPU:
      pdp_k_rollout_costate -> pdp_k_aux_lqr_bwd -> pdp_k_aux_lqr_fwd (dX/dtheta, dU/dtheta, fused loss / chain rule)
      -> pdp_k_reduce_loss_dp (batch sums of (loss, dp)) -> [N > 1] ONE NCCL all-reduce of r+2 doubles
--config c5  quadrotor SysID step, n=13 m=4 r=5 H=100, 32 768 per GPU (= 262 144 on 8): pdp_k_sens_fwd with the fused loss
      -> pdp_k_reduce_loss_dp -> [N > 1] the all-reduce of the outer-loop gradient (reference PDP/PDP.py:1293-1294)
--config c4  rocket powered-landing OC, n=13 m=3 H=100, 8 192 per GPU (= 65 536 on 8): rollout + costate + adjoint
      gradient dJ/dU (recmat semantics, PDP/PDP.py:1100-1114); outputs stay sharded, no collective
--config c2  cartpole ControlPlanning (Lagrange-polynomial policy r=6), n=4 m=1 H=50, 4 096 per GPU: pdp_k_sens_fwd with
      every output written (X, U, dX/dtheta, dU/dtheta, (loss, dtheta))

A "step" = one pass of that path over the per-rank batch.  Inputs are synthetic (seeded, float64).
`value` times the step with inputs resident in HBM; `e2e` times the C-ABI host-buffer call (pinned host inputs -> H2D ->
kernels -> D2H of the step's result) of the same step; `roofline` is for the step's dominant kernel, timed live with CUDA
events on the launching stream; `cpu_baseline` / `--impl reference` time the path on the box's host cores: the reference's
own NumPy half (unmodified `LQR.lqrSolver` / `integrateAuxSys`, staged under `baseline/_ref/reference_src` by `oracle/stage_reference.py`,
kind "reference") fed by the oracle's restatement of the CasADi half, or the oracle alone (kind "port") where that copy is
absent or the path has no NumPy half (C4).  A parity subset is checked against the oracle in every run.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_X, N_U, N_R = 13, 4, 9
TRUE_THETA = np.array([1, 1, 1, 1, 0.4, 1, 1, 5, 1.0])
METRIC = "PDP sweeps/sec (fwd+aux-LQR bwd)"


# ------------------------------------------------------------------------------------------- synthetic inputs
def synth_quadrotor(B, H, seed=0):
    """C3 batch: x0 positions U([-8,8]^2 x [3,9]), random small-angle attitude, theta = true + U(-0.4,0.4) clipped >= 0.1,
    U = 2.5 + 0.5 N(0,1) (hover thrust), demos = noisy copies."""
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


def synth_sysid(B, H, seed=0):
    """C5 batch (reference Examples/SysID/quadrotor/generate_traj.py:24-33, uav_PDP.py:39-40): inputs U(-10,10), the
    script's initial state, theta = theta_true + U(-0.3,0.3) shared by the batch."""
    rng = np.random.default_rng(seed)
    inputs = rng.uniform(-10, 10, (B, H, 4))
    q = np.array([1.0, 0, 0, 0])                                 # toQuaternion(0, [1,-1,1])
    x0 = np.tile(np.concatenate([[-8, -6, 9.], [0, 0, 0], q, [0, 0, 0]]), (B, 1))
    theta_true = np.array([1, 1, 1, 1, 0.4])
    theta = theta_true + np.random.default_rng(12345).uniform(-0.3, 0.3, 5)      # one draw, the same on every rank
    return inputs, x0, theta_true, theta


def synth_rocket(B, H, seed=0):
    """C4 batch (reference Examples/OC/rocket/rocket_PDP_Recmat.py:11-28): x0 = script state + 0.5 N, U = [10,0,0] + N."""
    rng = np.random.default_rng(seed)
    q = np.array([np.cos(0.75), 0, 0, np.sin(0.75)])            # toQuaternion(1.5, [0,0,1])
    x0 = np.concatenate([[10, -8, 5.], [-.1, 0, 0], q, [0, 0, 0]]) + 0.5 * rng.standard_normal((B, 13))
    U = np.array([10., 0, 0]) + rng.standard_normal((B, H, 3))
    return x0, U


def synth_cartpole(B, r, seed=0):
    """C2 batch (reference Examples/OC/cartpole/cartpole_PDP_poly.py:13-19,50): x0 ~ 0.1 N, theta ~ N(0, I_r)."""
    rng = np.random.default_rng(seed)
    return 0.1 * rng.standard_normal((B, 4)), rng.standard_normal((B, r))


# ------------------------------------------------------------------------------------------- algorithmic bytes / flops
def alg_bytes_bwd(n, m, r, H):
    """pdp_k_aux_lqr_bwd: read X, U, Lam, theta once; write the gain spill (K_t|k_t) once."""
    return 8 * (((H + 1) * n + H * m + H * n + r) + H * m * (n + r))


def alg_bytes_fwd(n, m, r, H):
    """pdp_k_aux_lqr_fwd: read X, U, theta, the gain spill, Xref, Uref once; write dX, dU, (loss, dp) once."""
    return 8 * (((H + 1) * n + H * m + r) + H * m * (n + r) + ((H + 1) * n + H * m) + ((H + 1) * n * r + H * m * r) + r + 1)


def alg_bytes_rollout(n, m, r, H):
    """pdp_k_rollout_costate: read x0, theta, U; write X, Lam, cost."""
    return 8 * ((n + r + H * m) + ((H + 1) * n + H * n + 1))


def alg_flops_bwd(n, m, r, H):
    return H * (4 * n ** 3 + 6 * n * n * m + 4 * n * m * m + m ** 3 / 3 + 4 * n * n * r + 4 * n * m * r + 2 * m * m * r + 400)


def alg_bytes_sweep(n, m, r, H):
    """SURVEY 8(d) figure for the whole IRL sweep (C3: 144 816 B)."""
    return 8 * ((n + r + H * m) + ((H + 1) * n + H * n + (H + 1) * n * r + H * m * r) + 2 * H * m * (n + r))


def alg_bytes_sysid(n, m, r, H):
    """SURVEY 8(d), SysID sweep with the fused loss (C5: 13 752 B)."""
    return 8 * (H * m + (H + 1) * n + r + 1)


def alg_bytes_adjoint(n, m, H):
    """SURVEY 8(d), adjoint-gradient sweep (C4: 25 816 B)."""
    return 8 * (n + H * m + (H + 1) * n + H * n + H * m + 1)


def alg_bytes_cp_full(n, m, r, H):
    """SURVEY 8(d), ControlPlanning forward-sensitivity sweep with every output written (C2 r=6: 14 360 B)."""
    return 8 * (n + r + (H + 1) * n + H * m + (H + 1) * n * r + H * m * r + r + 1)


def ncu_traffic(kernel, files):
    """DRAM bytes (read + write) per launch of ``kernel`` from the newest committed `ncu --set full` capture of the same
    workload (profiles/*_ncu_traffic.json); None if no capture is committed."""
    for name in files:
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))
            return float(d[kernel]["dram_bytes_read"]) + float(d[kernel]["dram_bytes_write"]), d.get("source")
        except Exception:
            continue
    return None, None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measure_fp64_peak(torch, dev):
    """FP64 FMA peak of this GPU in TFLOP/s from tools/microbench (8 independent DFMA chains per thread, 4 blocks of 256
    threads per SM, CUDA events); None if the microbenchmark library is not built."""
    import ctypes
    so = os.path.join(ROOT, "tools", "microbench", "libpdp_microbench.so")
    if not os.path.isfile(so):
        return None
    lib = ctypes.CDLL(so)
    lib.pdp_fp64_peak.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    blocks, iters = sms * 4, 4096
    out = torch.zeros(blocks * 256, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream(dev)
    best = 0.0
    for k in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        if lib.pdp_fp64_peak(iters, blocks, out.data_ptr(), st.cuda_stream) != 0:
            return None
        e1.record(st)
        torch.cuda.synchronize(dev)
        if k:
            best = max(best, 2.0 * 8 * 32 * iters * blocks * 256 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


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
            time.sleep(0.0005)

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


# ------------------------------------------------------------------------------------------- CPU arms (oracle port)
_ORACLE = {}


def _oracle(kind):
    """The oracle objects (test infrastructure; bench.py uses them for cpu_baseline / --impl reference / parity only)."""
    if kind not in _ORACLE:
        from oracle import envs, pdp_oracle
        if kind == "c3":
            oc = pdp_oracle.build_oc(envs.quadrotor(c=0.01, wthrust=0.1), 0.1)
            oc.diffPMP()
            _ORACLE[kind] = (oc, pdp_oracle)
        elif kind == "c5":
            e = envs.quadrotor(c=0.01)
            _ORACLE[kind] = (pdp_oracle.OracleSysID(e["X"], e["U"], e["dyn_params"], e["X"] + 0.1 * e["f"]), pdp_oracle)
        elif kind == "c4":
            e = envs.rocket(Jx=0.5, Jy=1., Jz=1., mass=1., l=1., wr=1, wv=1, wtilt=50, ww=1, wsidethrust=1, wthrust=0.4)
            _ORACLE[kind] = (pdp_oracle.OracleCP(e["X"], e["U"], e["X"] + 0.1 * e["f"], e["path_cost"], e["final_cost"]),
                             pdp_oracle)
        elif kind == "c2":
            e = envs.cartpole(mc=0.1, mp=0.1, l=1, wx=0.1, wq=0.6, wdx=0.1, wdq=0.1, wu=0.3)
            cp = pdp_oracle.OracleCP(e["X"], e["U"], e["X"] + 0.05 * e["f"], e["path_cost"], e["final_cost"])
            cp.set_poly(np.linspace(0, 50, 6))
            _ORACLE[kind] = (cp, pdp_oracle)
    return _ORACLE[kind]


def _ref_loader():
    """oracle/ref_loader when the unmodified reference PDP.py is reachable (build container, or staged under baseline/_ref/reference_src by
    oracle/stage_reference.py), else None: the CPU arms then time the oracle's restatement of the NumPy half as well."""
    from oracle import ref_loader
    return ref_loader if ref_loader.reference_available() else None


def cpu_kind(kind):
    """'reference' = the reference's own NumPy half (LQR.lqrSolver / SysID.integrateAuxSys / ControlPlanning.integrateAuxSys,
    unmodified) fed by the oracle's restatement of the CasADi half; 'port' = oracle only.  C4's reference
    path (recmat) is one CasADi function with no NumPy half."""
    return "reference" if kind != "c4" and _ref_loader() is not None else "port"


def _asmat(fn, shape, *a):
    return np.asarray(fn(*a), dtype=np.float64).reshape(shape)


def _cpu_worker(job):
    kind, H, idx, cores, data = job
    if cores > 1:                                   # one process per core: keep each worker's BLAS on its own core
        try:                                        # (pool over trajectories with OMP_NUM_THREADS=1)
            from threadpoolctl import threadpool_limits
            threadpool_limits(1)
        except ImportError:
            pass
    obj, po = _oracle(kind)
    rl = _ref_loader() if cpu_kind(kind) == "reference" else None
    t0 = time.perf_counter()
    if kind == "c3":
        x0, theta, U, _, _ = data
        for b in range(x0.shape[0]):
            if rl is None:
                po.pdp_sweep(obj, x0[b], U[b], theta[b])
            else:                                   # Examples/IRL/quadrotor/uav_PDP.py:52-65 with ocSolver's outputs given
                X, _ = obj.rollout(x0[b], U[b], theta[b])
                L = obj.costate(X, U[b], theta[b])
                rl.reference_lqr_solver(obj.getAuxSys(X, U[b], L, theta[b]), np.zeros((obj.n, obj.r)), H)
    elif kind == "c5":
        inputs, states, theta = data
        if rl is None:
            obj.step(list(inputs), list(states), theta)
        else:                                       # PDP.py:1261-1296 with the reference's own integrateAuxSys
            sid, n, r = rl.load_reference_pdp().SysID(), obj.n, obj.r
            for u, obs in zip(inputs, states):
                X = obj.integrateDyn(obs[0], u, theta)
                F = [_asmat(obj.dfx_fn, (n, n), X[t], u[t], theta) for t in range(H)]
                E = [_asmat(obj.dfe_fn, (n, r), X[t], u[t], theta) for t in range(H)]
                S = sid.integrateAuxSys(F, E, np.zeros((n, r)))["state_traj"]
                d, dp = X - obs, np.zeros(r)
                for t in range(H + 1):
                    dp += d[t] @ S[t]
    elif kind == "c4":
        x0, U = data
        for b in range(x0.shape[0]):
            obj.adjoint_grad(x0[b], U[b])
    elif kind == "c2":
        x0, theta = data
        if rl is not None:
            cpr, n, m, r = rl.load_reference_pdp().ControlPlanning(), obj.n, obj.m, obj.r
        for b in range(x0.shape[0]):
            if rl is None:
                obj.step(x0[b], H, theta[b])
            else:                                   # PDP.py:850-878 with the reference's own integrateAuxSys
                X, U, _ = obj.integrateSys(x0[b], H, theta[b])
                F = [_asmat(obj.dfx_fn, (n, n), X[t], U[t]) for t in range(H)]
                G = [_asmat(obj.dfu_fn, (n, m), X[t], U[t]) for t in range(H)]
                Ux = [_asmat(obj.dpolicy_dx_fn, (m, n), [t], X[t], theta[b]) for t in range(H)]
                Ue = [_asmat(obj.dpolicy_de_fn, (m, r), [t], X[t], theta[b]) for t in range(H)]
                aux = cpr.integrateAuxSys(F, G, Ux, Ue, np.zeros((n, r)))
                g = np.zeros(r)
                for t in range(H):
                    g += (_asmat(obj.dcx_fn, (1, n), X[t], U[t]) @ aux["state_traj"][t] +
                          _asmat(obj.dcu_fn, (1, m), X[t], U[t]) @ aux["control_traj"][t]).ravel()
                g += (_asmat(obj.dhx_fn, (1, n), X[H]) @ aux["state_traj"][H]).ravel()
    return time.perf_counter() - t0


def _cpu_jobs(kind, obj, n, H, cores):
    """``n`` trajectories of the config's synthetic workload dealt over ``cores`` workers."""
    if kind == "c3":
        data = synth_quadrotor(n, H, seed=1)
        jobs = [tuple(a[i::cores] for a in data) for i in range(cores)]
    elif kind == "c5":
        inputs, x0, th_true, theta = synth_sysid(n, H, seed=1)
        # observed states: a short oracle rollout would dominate the sample, so the (discarded) loss uses the inputs' own
        # rollout start; the arithmetic per trajectory is identical whatever the observations are
        states = np.zeros((n, H + 1, 13))
        states[:, 0] = x0
        jobs = [(inputs[i::cores], states[i::cores], theta) for i in range(cores)]
    elif kind == "c4":
        x0, U = synth_rocket(n, H, seed=1)
        jobs = [(x0[i::cores], U[i::cores]) for i in range(cores)]
    else:
        x0, theta = synth_cartpole(n, obj.r, seed=1)
        jobs = [(x0[i::cores], theta[i::cores]) for i in range(cores)]
    return jobs


def cpu_sweeps_per_s(kind, H, per_core, cores):
    """Oracle port (reference-shaped per-step NumPy / lambdified loops), one process per host core, ``per_core``
    trajectories of the config's workload each."""
    import multiprocessing as mp
    obj, po = _oracle(kind)                      # build before forking so the children inherit the lambdified functions
    if cpu_kind(kind) == "reference":
        _ref_loader().load_reference_pdp()       # likewise the imported reference module
    _cpu_worker((kind, H, 0, 1, _cpu_jobs(kind, obj, 1, H, 1)[0]))       # one-off costs (lazy imports, BLAS start-up) untimed
    n = per_core * cores
    jobs = _cpu_jobs(kind, obj, n, H, cores)
    jobs = [(kind, H, i, cores, j) for i, j in enumerate(jobs)]
    t0 = time.perf_counter()
    if cores == 1:
        _cpu_worker(jobs[0])
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    return n / wall, wall


CPU_KIND_NOTE = {"reference": "unmodified reference NumPy half (baseline/_ref/reference_src/PDP.py) + oracle restatement of the CasADi half",
                 "port": "oracle port"}
CPU_PER_CORE = {"c3": 48, "c5": 24, "c4": 24, "c2": 48}     # ~1-2 s of oracle work per core and step


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    H = cfg["H"]
    per_core = args.ref_per_core or CPU_PER_CORE[args.config]
    for _ in range(min(args.warmup, 1)):
        cpu_sweeps_per_s(args.config, H, 1, cores)
    vals = []
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, wall = cpu_sweeps_per_s(args.config, H, per_core, cores)
        vals.append(v)
    total = time.perf_counter() - t_all
    value = float(np.mean(vals))
    sample = "%d trajectories of the %s workload per step (%d per core x %d cores), %s" % (
        per_core * cores, args.config.upper(), per_core, cores, CPU_KIND_NOTE[cpu_kind(args.config)])
    print(json.dumps({
        "impl": "reference", "metric": cfg["metric"], "value": value, "unit": "sweeps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "batch_per_step": per_core * cores},
        "cpu_baseline": {"value": value, "unit": "sweeps/s", "cores": cores, "kind": cpu_kind(args.config),
                         "sample": sample},
        "e2e": {"value": value, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------- GPU workloads
def _rel(got, ref):
    return float(np.max(np.abs(got - ref)) / max(1e-300, np.max(np.abs(ref))))


class Workload:
    """One BASELINE config on one rank.  Sub-classes fill in the device step, the per-kernel split, parity, e2e."""

    collective = False           # does the step end in the (loss, dp) all-reduce?

    def __init__(self, args, rank, world, dev):
        import torch
        self.torch, self.args, self.rank, self.world, self.dev = torch, args, rank, world, dev
        self.stream = torch.cuda.current_stream(dev)

    def pin(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).pin_memory()

    def finish(self, loss_dp):
        """Batch sums of (loss, dp) in one kernel + the single all-reduce of an outer iteration."""
        from pontryagin_differentiable_programming_b200 import distributed, engine
        engine.reduce_loss_dp(loss_dp, out=self.sums)
        distributed.all_reduce_sums(self.sums)

    def timed(self, fn, steps):
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            fn()
        e1.record(self.stream)
        return e0, e1


class C3(Workload):
    name, H, B_default, collective = "c3", 50, 16384, True

    def __init__(self, args, rank, world, dev, B=None, seed_rank=None):
        super().__init__(args, rank, world, dev)
        torch = self.torch
        from pontryagin_differentiable_programming_b200 import systems
        self.B = B = int(B or args.batch or self.B_default)
        H = self.H = args.horizon or self.H
        self.sys = systems.quadrotor_irl(0.1)
        n, m, r = self.sys.n, self.sys.m, self.sys.r
        # the global batch is the concatenation of per-rank blocks, block g seeded with (0, g): every GPU count sees the
        # same trajectories in block g, and no rank has to materialise the other ranks' blocks
        self.host = [np.ascontiguousarray(a) for a in synth_quadrotor(B, H, seed=(0, rank if seed_rank is None else seed_rank))]
        self.pinned = [self.pin(a) for a in self.host]
        self.d = [p.to(dev) for p in self.pinned]
        mk = lambda *s: torch.empty(s, dtype=torch.float64, device=dev)
        self.out = {"X": mk(B, H + 1, n), "Lam": mk(B, H, n), "cost": mk(B), "dX": mk(B, H + 1, n, r), "dU": mk(B, H, m, r),
                    "loss_dp": mk(B, r + 1)}
        self.status = torch.zeros(B, dtype=torch.int32, device=dev)
        self.sums = mk(r + 2)
        self.parts = 4 if B >= 16384 else (2 if B >= 8192 else 1)            # pdp_sweep's automatic split
        self.launches_per_step = 1 + 2 * self.parts + 1
        self.metric = METRIC
        self.workload = "C3 quadrotor IRL PDP sweep n_x=13 n_u=4 r=9 H=%d" % H
        self.dominant = "pdp_k_aux_lqr_bwd"

    def step(self):
        # THE timed step: OCSystem.sweep -> C-ABI pdp_sweep (rollout/costate on the whole batch, then the aux-LQR phase in
        # sub-batches on the library's two internal streams), then the batch reduction (+ all-reduce at N > 1)
        d = self.d
        self.sys.sweep(d[0], d[1], d[2], Xref=d[3], Uref=d[4], status=self.status, out=self.out)
        self.finish(self.out["loss_dp"])

    def kernel_split(self, steps):
        """The same work as serial launches, each bracketed by CUDA events -> per-kernel durations for `roofline`."""
        torch, d, out, st = self.torch, self.d, self.out, self.stream
        def serial(ev=None):
            self.sys.rollout_costate(d[0], d[1], d[2], status=self.status, out=out)
            if ev: ev[0].record(st)
            self.sys.aux_lqr(out["X"], d[2], out["Lam"], d[1], status=self.status, phase="backward")
            if ev: ev[1].record(st)
            self.sys.aux_lqr(out["X"], d[2], out["Lam"], d[1], Xref=d[3], Uref=d[4], status=self.status, out=out, phase="forward")
            if ev: ev[2].record(st)
            self.finish(out["loss_dp"])
            if ev: ev[3].record(st)
        for _ in range(2):
            serial()
        kev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(5)) for _ in range(steps)]
        torch.cuda.synchronize(self.dev)
        for k in range(steps):
            kev[k][4].record(st)
            serial(kev[k])
        torch.cuda.synchronize(self.dev)
        mean = lambda f: float(np.mean([f(e) for e in kev]))
        return {"pdp_k_rollout_costate": mean(lambda e: e[4].elapsed_time(e[0])),
                "pdp_k_aux_lqr_bwd": mean(lambda e: e[0].elapsed_time(e[1])),
                "pdp_k_aux_lqr_fwd": mean(lambda e: e[1].elapsed_time(e[2])),
                "pdp_k_reduce_loss_dp%s" % (" + ncclAllReduce" if self.world > 1 else ""): mean(lambda e: e[2].elapsed_time(e[3]))}

    def alg_bytes(self):
        n, m, r, H = self.sys.n, self.sys.m, self.sys.r, self.H
        return {"pdp_k_rollout_costate": alg_bytes_rollout(n, m, r, H), "pdp_k_aux_lqr_bwd": alg_bytes_bwd(n, m, r, H),
                "pdp_k_aux_lqr_fwd": alg_bytes_fwd(n, m, r, H), "step": alg_bytes_sweep(n, m, r, H)}

    def parity(self):
        """First 64 trajectories of the timed batch + the two shipped quadrotor demos with their IPOPT U
        against the oracle; status bits reported separately."""
        torch = self.torch
        oc, po = _oracle("c3")
        h = self.host
        worst, nonfinite = 0.0, 0
        nb = min(64, self.B)
        got = {k: self.out[k][:nb].cpu().numpy() for k in ("X", "Lam", "dX", "dU")}
        for b in range(nb):
            X, L, cost, dX, dU = po.pdp_sweep(oc, h[0][b], h[2][b], h[1][b])
            if not all(np.isfinite(a).all() for a in (X, L, dX, dU)):
                nonfinite += 1
                continue
            for nm_, ref in (("X", X), ("Lam", L), ("dX", dX), ("dU", dU)):
                worst = max(worst, _rel(got[nm_][b], ref))
        demo = None
        gpath = os.path.join(ROOT, "tests", "golden", "k2_demos.npz")
        if os.path.isfile(gpath) and self.H == 50:
            g = np.load(gpath)
            th = np.asarray(g["quadrotor_true_parameter"], dtype=np.float64).reshape(-1)
            Xd = np.stack([g["quadrotor_%d_X" % i] for i in range(2)])
            Ud = np.stack([g["quadrotor_%d_U" % i] for i in range(2)])
            t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=self.dev)
            st = torch.zeros(2, dtype=torch.int32, device=self.dev)
            res = self.sys.sweep(t(Xd[:, 0]), t(th), t(Ud), status=st)
            demo = 0.0
            for i in range(2):
                X, L, cost, dX, dU = po.pdp_sweep(oc, Xd[i, 0], Ud[i], th)
                for nm_, ref in (("X", X), ("Lam", L), ("dX", dX), ("dU", dU)):
                    demo = max(demo, _rel(res[nm_][i].cpu().numpy(), ref))
            demo = {"max_rel_err": demo, "status": [int(v) for v in st.cpu()],
                    "X_vs_shipped_ipopt_demo": float(np.max(np.abs(res["X"].cpu().numpy() - Xd)))}
        stt = self.status
        fin = torch.isfinite(self.out["dX"]).all(dim=(1, 2, 3)) & torch.isfinite(self.out["dU"]).all(dim=(1, 2, 3))
        return {"max_rel_err_vs_oracle_first64": worst, "oracle_nonfinite_skipped": nonfinite, "tolerance": 1e-6,
                "shipped_demos": demo,
                "status_bit0_nonfinite": int((stt & 1).ne(0).sum().item()),
                "status_bit1_quu_not_pd": int((stt & 2).ne(0).sum().item()),
                "trajectories_with_nonfinite_outputs": int((~fin).sum().item()),
                "status_note": "bit 1 = Quu not positive definite: expected for random (non-optimal) controls, the sweep is still "
                               "the reference's algebra (parity above); 0 on the shipped demos (OC optima)"}

    def e2e_setup(self):
        torch, B, r = self.torch, self.B, self.sys.r
        self.ldp_host = torch.empty((B, r + 1), dtype=torch.float64).pin_memory()
        self.cost_host = torch.empty((B,), dtype=torch.float64).pin_memory()
        self.h2d = sum(int(p.numel()) * 8 for p in self.pinned)
        self.d2h = int(self.ldp_host.numel() + self.cost_host.numel()) * 8
        self.e2e_api = "OCSystem.sweep_host -> pdp_sweep_host (C ABI, pinned host buffers), %d sub-batches on 2 streams" % self.args.e2e_chunks

    def e2e_step(self):
        p = self.pinned
        self.sys.sweep_host(p[0], p[1], p[2], p[3], p[4], self.ldp_host, cost_h=self.cost_host, keep_dtraj=True,
                            n_chunks=self.args.e2e_chunks, device=self.dev)

    def e2e_check(self):
        return bool(self.torch.allclose(self.ldp_host.to(self.dev), self.out["loss_dp"], rtol=1e-12, atol=0))

    def e2e_extra(self, steps):
        """Second end-to-end figure: the sensitivities themselves (dX/dtheta, dU/dtheta) copied back to the host."""
        torch, B, H = self.torch, self.B, self.H
        n, m, r = self.sys.n, self.sys.m, self.sys.r
        dX_h = torch.empty((B, H + 1, n, r), dtype=torch.float64).pin_memory()
        dU_h = torch.empty((B, H, m, r), dtype=torch.float64).pin_memory()
        p = self.pinned
        fn = lambda: self.sys.sweep_host(p[0], p[1], p[2], p[3], p[4], self.ldp_host, cost_h=self.cost_host,
                                         n_chunks=max(self.args.e2e_chunks, 8), device=self.dev, dX_h=dX_h, dU_h=dU_h)
        for _ in range(2):
            fn()
        torch.cuda.synchronize(self.dev)
        e0, e1 = self.timed(fn, steps)
        torch.cuda.synchronize(self.dev)
        ms = e0.elapsed_time(e1) / steps
        ok = bool(torch.equal(dX_h[:256].to(self.dev), self.out["dX"][:256]))
        return {"what": "same call with dX/dtheta and dU/dtheta copied back to pinned host memory (pdp_sweep_host_traj)",
                "value_per_gpu": B / (ms * 1e-3), "unit": "sweeps/s", "ms_per_step": ms,
                "d2h_bytes_per_step": self.d2h + (dX_h.numel() + dU_h.numel()) * 8, "matches_device_path": ok}


class C5(Workload):
    name, H, B_default, collective = "c5", 100, 32768, True

    def __init__(self, args, rank, world, dev):
        super().__init__(args, rank, world, dev)
        torch = self.torch
        from pontryagin_differentiable_programming_b200 import systems
        self.B = B = int(args.batch or self.B_default)
        H = self.H = args.horizon or self.H
        self.sys = systems.quadrotor_sysid(0.1)
        n, m, r = self.sys.n, self.sys.m, self.sys.r
        inputs, x0, th_true, theta = synth_sysid(B, H, seed=(5, rank))
        self.host = {"inputs": inputs, "x0": x0, "theta_true": th_true, "theta": theta}
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
        self.inputs, self.x0, self.theta = t(inputs), t(x0), t(theta)
        # observations = rollout at the true parameter (reference generate_traj.py:34-36), produced by the same kernel;
        # the parity subset below re-derives them with the oracle
        self.Xobs = self.sys.step(self.inputs, None, t(th_true), x0=self.x0, want_traj=True)["X"]
        self.sums = torch.empty(r + 2, dtype=torch.float64, device=dev)
        self.status = torch.zeros(B, dtype=torch.int32, device=dev)
        self.launches_per_step = 2
        self.metric = "PDP sweeps/sec (SysID forward-sensitivity sweep, fused loss)"
        self.workload = "C5 quadrotor SysID step n_x=13 n_u=4 r=5 H=%d" % H
        self.dominant = "pdp_k_sens_fwd"

    def step(self):
        self.res = self.sys.step(self.inputs, self.Xobs, self.theta, x0=self.x0, status=self.status)
        self.finish(self.res["loss_dp"])

    def kernel_split(self, steps):
        torch, st = self.torch, self.stream
        kev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(steps)]
        for _ in range(2):
            self.step()
        torch.cuda.synchronize(self.dev)
        for k in range(steps):
            kev[k][0].record(st)
            res = self.sys.step(self.inputs, self.Xobs, self.theta, x0=self.x0, status=self.status)
            kev[k][1].record(st)
            self.finish(res["loss_dp"])
            kev[k][2].record(st)
        torch.cuda.synchronize(self.dev)
        mean = lambda f: float(np.mean([f(e) for e in kev]))
        return {"pdp_k_sens_fwd": mean(lambda e: e[0].elapsed_time(e[1])),
                "pdp_k_reduce_loss_dp%s" % (" + ncclAllReduce" if self.world > 1 else ""): mean(lambda e: e[1].elapsed_time(e[2]))}

    def alg_bytes(self):
        b = alg_bytes_sysid(self.sys.n, self.sys.m, self.sys.r, self.H)
        return {"pdp_k_sens_fwd": b, "step": b}

    def parity(self):
        sid, po = _oracle("c5")
        h = self.host
        nb = min(64, self.B)
        Xobs = self.Xobs[:nb].cpu().numpy()
        ldp = self.res["loss_dp"][:nb].cpu().numpy()
        worst_obs = worst = 0.0
        for b in range(nb):
            Xo = sid.integrateDyn(h["x0"][b], h["inputs"][b], h["theta_true"])
            worst_obs = max(worst_obs, _rel(Xobs[b], Xo))
            loss, dp = sid.step([h["inputs"][b]], [Xo], h["theta"])
            worst = max(worst, abs(ldp[b, 0] - loss) / max(abs(loss), 1e-300), _rel(ldp[b, 1:], dp))
        sums = self.sums.cpu().numpy()
        tot = self.res["loss_dp"].sum(dim=0).cpu().numpy() if self.world == 1 else None
        return {"max_rel_err_loss_dp_vs_oracle_first64": worst, "max_rel_err_observations_vs_oracle_first64": worst_obs,
                "tolerance": 1e-6, "status_bit0_nonfinite": int((self.status & 1).ne(0).sum().item()),
                "reduction_vs_torch_sum": None if tot is None else _rel(sums[:-1], tot), "count": float(sums[-1])}

    def e2e_setup(self):
        torch, B, r = self.torch, self.B, self.sys.r
        self.p_in, self.p_x0 = self.pin(self.host["inputs"]), self.pin(self.host["x0"])
        self.p_Xobs = self.Xobs.cpu().pin_memory()
        self.p_th = self.pin(self.host["theta"].reshape(1, -1))
        self.sums_host = torch.zeros((self.args.e2e_chunks, r + 2), dtype=torch.float64).pin_memory()
        self.h2d = 8 * int(self.p_in.numel() + self.p_x0.numel() + self.p_Xobs.numel() + self.p_th.numel() * self.args.e2e_chunks)
        self.d2h = 8 * int(self.sums_host.numel())
        self.e2e_api = "SysIDSystem.step_host -> pdp_sens_fwd_host (C ABI, pinned host buffers), %d sub-batches on 2 streams; " \
                       "result = per-sub-batch (sum loss, sum dp, count)" % self.args.e2e_chunks

    def e2e_step(self):
        self.sys.step_host(self.p_x0, self.p_th, self.H, inputs_h=self.p_in, Xobs_h=self.p_Xobs, sums_h=self.sums_host,
                           n_chunks=self.args.e2e_chunks, device=self.dev)

    def e2e_check(self):
        tot = self.sums_host.sum(dim=0).numpy()
        ref = self.res["loss_dp"].sum(dim=0).cpu().numpy()
        return bool(_rel(tot[:-1], ref) < 1e-12 and tot[-1] == self.B)


class C4(Workload):
    name, H, B_default = "c4", 100, 8192

    def __init__(self, args, rank, world, dev):
        super().__init__(args, rank, world, dev)
        torch = self.torch
        from pontryagin_differentiable_programming_b200 import systems
        self.B = B = int(args.batch or self.B_default)
        H = self.H = args.horizon or self.H
        self.sys = systems.rocket_oc_adjoint(0.1)
        x0, U = synth_rocket(B, H, seed=(4, rank))
        self.host = {"x0": x0, "U": U}
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
        self.x0, self.U = t(x0), t(U)
        self.th = torch.zeros((1, 1), dtype=torch.float64, device=dev)
        mk = lambda *s: torch.empty(s, dtype=torch.float64, device=dev)
        self.out = {"X": mk(B, H + 1, 13), "Lam": mk(B, H, 13), "cost": mk(B), "dHu": mk(B, H, 3)}
        self.status = torch.zeros(B, dtype=torch.int32, device=dev)
        self.launches_per_step = 1
        self.metric = "PDP sweeps/sec (OC adjoint-gradient sweep: rollout + costate + dJ/dU)"
        self.workload = "C4 rocket powered-landing OC n_x=13 n_u=3 H=%d (recmat semantics)" % H
        # open-loop rollouts of up to 2 x 32 x #SM trajectories run as the TMA kernel (thread-private bulk copies), larger ones
        # as the register-prefetch kernel (the module's launcher decides; see kernel_templates.K_LAUNCH_ROLLOUT_TMA_BRANCH)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        self.dominant = "pdp_k_rollout_costate_tma" if (B + 31) // 32 <= 2 * sms else "pdp_k_rollout_costate"

    def step(self):
        self.sys.rollout_costate(self.x0, self.th, self.U, want_dHu=True, status=self.status, out=self.out)

    def kernel_split(self, steps):
        torch = self.torch
        torch.cuda.synchronize(self.dev)
        e0, e1 = self.timed(self.step, steps)
        torch.cuda.synchronize(self.dev)
        return {self.dominant: e0.elapsed_time(e1) / steps}

    def alg_bytes(self):
        b = alg_bytes_adjoint(13, 3, self.H)
        return {self.dominant: b, "step": b}

    def parity(self):
        cp, po = _oracle("c4")
        h = self.host
        nb = min(64, self.B)
        got = {k: self.out[k][:nb].cpu().numpy() for k in ("X", "cost", "dHu")}
        worst = 0.0
        for b in range(nb):
            cost, g, X = cp.adjoint_grad(h["x0"][b], h["U"][b])
            worst = max(worst, _rel(got["X"][b], X), abs(got["cost"][b] - cost) / abs(cost), _rel(got["dHu"][b], g))
        return {"max_rel_err_X_cost_dJdU_vs_oracle_first64": worst, "tolerance": 1e-6,
                "status_bit0_nonfinite": int((self.status & 1).ne(0).sum().item())}

    def e2e_setup(self):
        torch, B, H = self.torch, self.B, self.H
        self.p_x0, self.p_U = self.pin(self.host["x0"]), self.pin(self.host["U"])
        self.p_th = torch.zeros((1, 1), dtype=torch.float64).pin_memory()
        self.cost_host = torch.empty((B,), dtype=torch.float64).pin_memory()
        self.dHu_host = torch.empty((B, H, 3), dtype=torch.float64).pin_memory()
        self.h2d = 8 * int(self.p_x0.numel() + self.p_U.numel() + self.args.e2e_chunks)
        self.d2h = 8 * int(self.cost_host.numel() + self.dHu_host.numel())
        self.e2e_api = "OCSystem.rollout_costate_host -> pdp_rollout_costate_host (C ABI, pinned host buffers), %d sub-batches " \
                       "on 2 streams; result = cost[B], dJ/dU[B,H,m]" % self.args.e2e_chunks

    def e2e_step(self):
        self.sys.rollout_costate_host(self.p_x0, self.p_th, self.p_U, cost_h=self.cost_host, dHu_h=self.dHu_host,
                                      n_chunks=self.args.e2e_chunks, device=self.dev)

    def e2e_check(self):
        return bool(self.torch.equal(self.dHu_host.to(self.dev), self.out["dHu"]))


class C2(Workload):
    name, H, B_default = "c2", 50, 4096

    def __init__(self, args, rank, world, dev):
        super().__init__(args, rank, world, dev)
        torch = self.torch
        from pontryagin_differentiable_programming_b200 import systems
        self.B = B = int(args.batch or self.B_default)
        H = self.H = 50
        self.sys = systems.cartpole_cp("poly", H, 0.05)
        x0, theta = synth_cartpole(B, self.sys.r, seed=(2, rank))
        self.host = {"x0": x0, "theta": theta}
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
        self.x0, self.theta = t(x0), t(theta)
        self.status = torch.zeros(B, dtype=torch.int32, device=dev)
        self.launches_per_step = 1
        self.metric = "PDP sweeps/sec (ControlPlanning forward-sensitivity sweep, all outputs)"
        self.workload = "C2 cartpole ControlPlanning n_x=4 n_u=1 poly policy r=%d H=%d" % (self.sys.r, H)
        self.dominant = "pdp_k_sens_fwd"

    def step(self):
        self.res = self.sys.step(self.x0, self.H, self.theta, want_traj=True, want_sens=True, status=self.status)

    def kernel_split(self, steps):
        torch = self.torch
        torch.cuda.synchronize(self.dev)
        e0, e1 = self.timed(self.step, steps)
        f0, f1 = self.timed(lambda: self.sys.step(self.x0, self.H, self.theta), steps)
        torch.cuda.synchronize(self.dev)
        self.fused_ms = f0.elapsed_time(f1) / steps
        return {"pdp_k_sens_fwd": e0.elapsed_time(e1) / steps}

    def alg_bytes(self):
        b = alg_bytes_cp_full(self.sys.n, self.sys.m, self.sys.r, self.H)
        return {"pdp_k_sens_fwd": b, "step": b}

    def parity(self):
        cp, po = _oracle("c2")
        h = self.host
        nb = min(64, self.B)
        got = {k: self.res[k][:nb].cpu().numpy() for k in ("X", "U", "dX", "dU", "loss_dp")}
        worst = 0.0
        for b in range(nb):
            cost, g, X, U, dX, dU = cp.step(h["x0"][b], self.H, h["theta"][b], return_traj=True)
            if not np.isfinite(cost):
                continue
            worst = max(worst, _rel(got["X"][b], X), _rel(got["U"][b], U), _rel(got["dX"][b], dX), _rel(got["dU"][b], dU),
                        abs(got["loss_dp"][b, 0] - cost) / abs(cost), _rel(got["loss_dp"][b, 1:], g))
        return {"max_rel_err_all_outputs_vs_oracle_first64": worst, "tolerance": 1e-6,
                "status_bit0_nonfinite": int((self.status & 1).ne(0).sum().item()),
                "fused_loss_only_ms": getattr(self, "fused_ms", None)}

    def e2e_setup(self):
        torch, B, r = self.torch, self.B, self.sys.r
        self.p_x0, self.p_th = self.pin(self.host["x0"]), self.pin(self.host["theta"])
        self.ldp_host = torch.empty((B, r + 1), dtype=torch.float64).pin_memory()
        self.h2d = 8 * int(self.p_x0.numel() + self.p_th.numel())
        self.d2h = 8 * int(self.ldp_host.numel())
        self.e2e_api = "CPSystem.step_host -> pdp_sens_fwd_host (C ABI, pinned host buffers), %d sub-batches on 2 streams; " \
                       "result = (loss, dtheta)[B, r+1]" % self.args.e2e_chunks

    def e2e_step(self):
        self.sys.step_host(self.p_x0, self.p_th, self.H, loss_dp_h=self.ldp_host, n_chunks=self.args.e2e_chunks, device=self.dev)

    def e2e_check(self):
        return bool(self.torch.allclose(self.ldp_host.to(self.dev), self.res["loss_dp"], rtol=1e-12, atol=0))


WORKLOADS = {"c3": C3, "c5": C5, "c4": C4, "c2": C2}
CONFIGS = {
    "c3": {"H": 50, "metric": METRIC, "workload": "C3 quadrotor IRL sweep n=13 m=4 r=9 H=50"},
    "c5": {"H": 100, "metric": "PDP sweeps/sec (SysID forward-sensitivity sweep, fused loss)",
           "workload": "C5 quadrotor SysID step n=13 m=4 r=5 H=100"},
    "c4": {"H": 100, "metric": "PDP sweeps/sec (OC adjoint-gradient sweep: rollout + costate + dJ/dU)",
           "workload": "C4 rocket OC n=13 m=3 H=100 (recmat semantics)"},
    "c2": {"H": 50, "metric": "PDP sweeps/sec (ControlPlanning forward-sensitivity sweep, all outputs)",
           "workload": "C2 cartpole ControlPlanning n=4 m=1 r=6 H=50"},
}
TRAFFIC_FILES = {"c3": ("r2_ncu_traffic.json", "r1m_ncu_traffic.json"), "c5": ("r2_c5_ncu_traffic.json",),
                 "c4": ("r2_c4_ncu_traffic.json",), "c2": ("r2_c2_ncu_traffic.json",)}


def bind_numa(local, world):
    """Give each rank its own slice of the host cores (and, where the box exposes NUMA nodes per GPU, its GPU's node) so
    that the pinned staging buffers and the copy-issuing threads of the ranks do not share cores."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        node = None
        p = "/sys/bus/pci/devices"
        try:
            import torch
            bus = torch.cuda.get_device_properties(local).pci_bus_id
            dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
            cand = "%s/%04x:%02x:00.0/numa_node" % (p, dom, bus)
            if os.path.isfile(cand):
                node = int(open(cand).read().strip())
        except Exception:
            node = None
        if node is not None and node >= 0:
            lst = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
            nodec = set()
            for part in lst.split(","):
                a, _, b = part.partition("-")
                nodec.update(range(int(a), int(b or a) + 1))
            mine = [c for c in cores if c in nodec]
            if mine:
                cores = mine
        if world > 1 and len(cores) >= world:
            per = len(cores) // world
            cores = cores[local * per:(local + 1) * per]
        os.sched_setaffinity(0, cores)
        return {"numa_node": node, "cores": "%d-%d (%d)" % (cores[0], cores[-1], len(cores))}
    except Exception as exc:  # pragma: no cover
        return {"error": repr(exc)}


def run_gpu(args, cfg):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the PDP B200 engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_numa(local, world) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout for the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.config](args, rank, world, dev)
    B, H = wl.B, wl.H
    steps, warm = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(warm):
        wl.step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = wl.timed(wl.step, steps)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    split = wl.kernel_split(steps)
    barrier()
    wl.step()
    torch.cuda.synchronize(dev)
    parity = wl.parity() if rank == 0 else None

    # ---- e2e: the public host-buffer API: H2D of every input + all kernels + D2H of the result inside the timed region
    wl.e2e_setup()
    for _ in range(3):
        wl.e2e_step()
    barrier()
    f0, f1 = wl.timed(wl.e2e_step, steps)
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    e2e_ok = wl.e2e_check()
    extra = wl.e2e_extra(max(2, steps // 4)) if (hasattr(wl, "e2e_extra") and world == 1 and not args.no_e2e_extra) else None

    # ---- strong scaling beside the weak line (N > 1, C3 only): the SAME global batch of one GPU's size cut over the ranks
    strong_ms = None
    if world > 1 and args.config == "c3" and B % world == 0:
        ws = C3(args, rank, world, dev, B=B // world, seed_rank=1000 + rank)
        for _ in range(warm):
            ws.step()
        barrier()
        s0, s1 = ws.timed(ws.step, steps)
        barrier()
        strong_ms = s0.elapsed_time(s1)
        del ws

    names = list(split)
    times = torch.tensor([ms_total, e2e_ms, strong_ms or 0.0] + [split[k] for k in names], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    tv = [float(v) for v in times.cpu()]
    ms_total, e2e_ms, strong_ms = tv[0], tv[1], tv[2]
    split = dict(zip(names, tv[3:]))

    if rank == 0:
        value = B * world * steps / (ms_total * 1e-3)
        e2e_val = B * world * steps / (e2e_ms * 1e-3)
        peak, peak_src = measured_peaks()
        ab = wl.alg_bytes()
        k_ms = split[wl.dominant]
        kbytes = ab[wl.dominant] * B
        achieved = kbytes / (k_ms * 1e-3) / 1e9
        full = (args.batch in (0, wl.B_default)) and (args.horizon in (0, cfg["H"]))
        traffic, traffic_src = ncu_traffic(wl.dominant, TRAFFIC_FILES[args.config]) if full else (None, None)
        kernels = {}
        for k, ms in split.items():
            kernels[k] = {"ms": ms}
            if k in ab:
                kernels[k].update({"alg_bytes_per_launch": ab[k] * B, "achieved_GBps": ab[k] * B / (ms * 1e-3) / 1e9,
                                   "frac_of_hbm_peak": ab[k] * B / (ms * 1e-3) / 1e9 / peak})
        par = "batch-sharded x%d" % world
        if wl.collective:
            par += (", per step ONE NCCL all-reduce of %d float64 (sum loss, sum dp, count) after the in-library batch reduction"
                    % int(wl.sums.numel())) if world > 1 else ", batch reduction of (loss, dp) in the step (the all-reduce joins at N > 1)"
        else:
            par += ", outputs stay sharded: no data-path collective (OC mode)"
        line = {
            "metric": wl.metric, "value": value, "unit": "sweeps/s", "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": wl.workload, "batch_per_gpu": B, "global_batch": B * world, "parallelism": par,
                       "l2": "per-step working set %.2f GB per GPU > 126 MB L2 (no flush needed)" % (ab["step"] * B / 1e9)
                             if ab["step"] * B > 4 * 126e6 else
                             "per-step working set %.0f MB per GPU: inputs and outputs of consecutive steps may be L2-resident "
                             "(the config's own size; L2 is not flushed between steps)" % (ab["step"] * B / 1e6),
                       "parity": parity, "e2e_matches_device_path": e2e_ok, "host_affinity": affinity},
            "roofline": {"kernel": wl.dominant, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "alg_bytes_per_launch": kbytes, "kernel_ms": k_ms,
                         "kernel_share_of_step": k_ms / sum(split.values()),
                         "kernel_timing": "serial launches bracketed by CUDA events on the launching stream",
                         "kernels": kernels,
                         "step_alg_bytes": ab["step"] * B,
                         "step_achieved_GBps": ab["step"] * B / (ms_total / steps * 1e-3) / 1e9,
                         "step_frac_of_hbm_peak": ab["step"] * B / (ms_total / steps * 1e-3) / 1e9 / peak},
            "e2e": {"value": e2e_val, "unit": "sweeps/s", "h2d_bytes_per_step": wl.h2d, "d2h_bytes_per_step": wl.d2h,
                    "api": wl.e2e_api},
            "gpu_launches": wl.launches_per_step * steps, "clocks": clocks,
        }
        if args.config == "c3":
            n, m, r = wl.sys.n, wl.sys.m, wl.sys.r
            line["roofline"]["fp64_tflops_alg_bwd"] = alg_flops_bwd(n, m, r, H) * B / (k_ms * 1e-3) / 1e12
            line["roofline"]["fp64_peak_tflops_measured"] = measure_fp64_peak(torch, dev)
            line["roofline"]["fp64_note"] = "fp64_tflops_alg_bwd is a dense-equivalent flop count (SURVEY 8(d)); the kernel exploits " \
                                            "sparsity, so its executed-instruction FP64 pipe utilisation is the ncu figure in " \
                                            "profiles/ (sm__inst_executed_pipe_fp64), not this number over the measured peak"
            line["config"]["step"] = "OCSystem.sweep -> pdp_sweep: 1 rollout/costate launch + %d sub-batches x (bwd, fwd) on two " \
                                     "streams, then pdp_reduce_loss_dp" % wl.parts
        if extra is not None:
            line["e2e"]["with_sensitivities_to_host"] = extra
        if strong_ms:
            line["strong_scaling"] = {"global_batch": B, "batch_per_gpu": B // world, "n_gpus": world,
                                      "value": B * steps / (strong_ms * 1e-3), "unit": "sweeps/s",
                                      "ms_per_step": strong_ms / steps,
                                      "note": "same step, one GPU's batch cut over the ranks (strong scaling)"}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            per_core = args.ref_per_core or CPU_PER_CORE[args.config]
            v, wall = cpu_sweeps_per_s(args.config, H, per_core, cores)
            line["cpu_baseline"] = {"value": v, "unit": "sweeps/s", "cores": cores, "kind": cpu_kind(args.config),
                                    "sample": "%d trajectories of the same workload (%d per core), %.1f s wall, %s"
                                              % (per_core * cores, per_core, wall, CPU_KIND_NOTE[cpu_kind(args.config)])}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="trajectories per GPU (0 = the config's per-GPU size)")
    ap.add_argument("--horizon", type=int, default=0)
    ap.add_argument("--ref-per-core", type=int, default=0,
                    help="trajectories per host core in one CPU-baseline step (0 = about 1-2 s of work per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e-extra", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=4)
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.horizon:
        cfg["H"] = args.horizon
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_gpu(args, cfg)


if __name__ == "__main__":
    main()
