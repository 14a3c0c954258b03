"""Batched entry points on torch CUDA float64 tensors (the additive API of SURVEY 8(b)).

PyTorch is plumbing only (device memory, streams); all arithmetic runs in the generated
sm_100a kernels behind the C ABI.  Every method raises ``PDPBackendError`` when no CUDA device
is available -- there is no CPU fallback.
"""
from __future__ import annotations

import torch

from . import backend, build, codegen
from .backend import PDPBackendError


def require_cuda():
    if not torch.cuda.is_available():
        raise PDPBackendError(
            "the PDP B200 engine needs a CUDA device: hot-path methods (rollout / costate / aux-LQR / "
            "sensitivity sweeps) run only as sm_100a kernels and have no CPU fallback")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _chk(t, shape, name, device):
    if t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous():
        raise ValueError("%s must be a contiguous CUDA float64 tensor" % name)
    if tuple(t.shape) != tuple(shape):
        raise ValueError("%s has shape %s, expected %s" % (name, tuple(t.shape), tuple(shape)))
    if t.device != device:
        raise ValueError("%s is on %s, expected %s" % (name, t.device, device))


class OCSystem:
    """Compiled optimal-control system: rollout/costate, fused getAuxSys+lqrSolver, dense aux eval."""

    # measured defaults of the backward Riccati kernel (profiles/README.md): the two-trajectories-per-warp kernel
    # wherever the stack fits two rows per team lane (n <= 16, m + r <= 16), else one trajectory per warp
    BWD_DEFAULTS = {2: dict(chunk=8, warps_per_block=1, min_blocks=8, keep_fg=True),
                    1: dict(chunk=17, warps_per_block=4, min_blocks=3, keep_fg=True)}

    def __init__(self, state, control, auxvar, dyn, path_cost, final_cost, verbose=False, **kernel_options):
        """``kernel_options``: keyword options of :class:`codegen.OCModuleSource` (``chunk``, ``warps_per_block``,
        ``min_blocks``, ``keep_fg``, ``bwd_pack``, ``fwd_pack``, ``prefetch``, ... -- the A/B switches of
        ``tools/tune_aux_lqr.py``); anything not given takes the measured default."""
        fits = state.numel() <= 16 and control.numel() + auxvar.numel() <= 16
        opts = dict(fwd_warps_per_block=1, fwd_min_blocks=8, fast_rcp=True, early_solve=True)
        opts.update({k: v for k, v in kernel_options.items() if v is not None})
        opts["bwd_pack"] = 2 if (int(opts.get("bwd_pack", 2)) == 2 and fits) else 1
        for k, v in self.BWD_DEFAULTS[opts["bwd_pack"]].items():
            opts.setdefault(k, v)
        self.src = codegen.OCModuleSource(state, control, auxvar, dyn, path_cost, final_cost, **opts)
        self.n, self.m, self.r = self.src.n, self.src.m, self.src.r
        self.module_path = build.compile_module(self.src.source(), self.src.key(), verbose=verbose)
        self._handle = None
        self._ws = {}

    # the handle is created lazily so that systems can be *built* on a CPU-only box
    @property
    def handle(self):
        if self._handle is None:
            self._handle = backend.SystemHandle(self.module_path)
        return self._handle

    def _theta(self, theta, B, device):
        if theta.dim() == 1:
            theta = theta.unsqueeze(0)
        if theta.shape[0] == 1:
            _chk(theta, (1, self.r), "theta", device)
            return theta, 0
        _chk(theta, (B, self.r), "theta", device)
        return theta, self.r

    def _workspace(self, op, B, H, device):
        need = self.handle.workspace_bytes(op, B, H)
        key = (op, device)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(max(need, 256), dtype=torch.uint8, device=device)
            self._ws[key] = ws
        return ws, need

    def rollout_costate(self, x0, theta, U, want_costate=True, want_dHu=False, status=None, out=None):
        """x0[B,n], theta[B|1,r], U[B,H,m] -> X[B,H+1,n], Lam[B,H,n] (Lam[:,t] = lambda_{t+1}), cost[B][, dHu[B,H,m]]."""
        require_cuda()
        dev = x0.device
        B, H = U.shape[0], U.shape[1]
        _chk(x0, (B, self.n), "x0", dev)
        _chk(U, (B, H, self.m), "U", dev)
        theta, ts = self._theta(theta, B, dev)
        def buf(name, shape):
            if out is not None and name in out:
                return out[name]
            return torch.empty(shape, dtype=torch.float64, device=dev)

        X = buf("X", (B, H + 1, self.n))
        Lam = buf("Lam", (B, H, self.n)) if (want_costate or want_dHu) else None
        cost = buf("cost", (B,))
        dHu = buf("dHu", (B, H, self.m)) if want_dHu else None
        lib = self.handle.lib
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            backend.check(lib.pdp_rollout_costate(self.handle.ptr, B, H, _ptr(x0), _ptr(theta), ts, _ptr(U), _ptr(X),
                                                  _ptr(Lam), _ptr(cost), _ptr(dHu), _ptr(status), st),
                          "pdp_rollout_costate")
        out = {"X": X, "Lam": Lam, "cost": cost}
        if want_dHu:
            out["dHu"] = dHu
        return out

    def rollout_feedback(self, x0, theta, Uref, Xref, gains, alpha, want_dHu=True, want_costate=True, status=None,
                         group=1):
        """Closed-loop rollout u_t = Uref[t] + alpha k_t + K_t (x_t - Xref[t]); gains[B,H,n+1,m] (or the raw
        workspace of a one-column Riccati sweep).  ``group`` > 1: ``alpha`` has B*group entries and candidate
        ``b*group + j`` rolls trajectory ``b`` out with ``alpha[b*group + j]`` (a whole line search in one launch).
        -> dict U (applied), X, cost [, Lam, dHu], leading dimension B*group."""
        require_cuda()
        dev = x0.device
        B, H = Uref.shape[0], Uref.shape[1]
        theta, ts = self._theta(theta, B, dev)
        Bx = B * int(group)
        mk = lambda *shape: torch.empty(shape, dtype=torch.float64, device=dev)
        X, cost, Uout = mk(Bx, H + 1, self.n), mk(Bx), mk(Bx, H, self.m)
        Lam = mk(Bx, H, self.n) if (want_costate or want_dHu) else None
        dHu = mk(Bx, H, self.m) if want_dHu else None
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            backend.check(self.handle.lib.pdp_rollout_feedback(self.handle.ptr, Bx, H, _ptr(x0), _ptr(theta), ts, _ptr(Uref),
                                                               _ptr(Xref), _ptr(gains), _ptr(alpha), _ptr(Uout), _ptr(X),
                                                               _ptr(Lam), _ptr(cost), _ptr(dHu), int(group), _ptr(status), st),
                          "pdp_rollout_feedback")
        out = {"U": Uout, "X": X, "cost": cost}
        if Lam is not None:
            out["Lam"] = Lam
        if want_dHu:
            out["dHu"] = dHu
        return out

    def aux_lqr(self, X, U, Lam, theta, X0aux=None, want_traj=True, Xref=None, Uref=None, status=None, out=None,
                phase="both"):
        """Fused getAuxSys + lqrSolver: -> dX[B,H+1,n,r], dU[B,H,m,r] and/or loss_dp[B,r+1].
        ``phase`` = "both" | "backward" (Riccati sweep only, gains stay in the workspace) | "forward"."""
        require_cuda()
        dev = X.device
        B, H = U.shape[0], U.shape[1]
        _chk(X, (B, H + 1, self.n), "X", dev)
        _chk(U, (B, H, self.m), "U", dev)
        _chk(Lam, (B, H, self.n), "Lam", dev)
        theta, ts = self._theta(theta, B, dev)
        x0s = 0
        if X0aux is not None:
            if X0aux.dim() == 2:
                X0aux = X0aux.unsqueeze(0)
            x0s = 0 if X0aux.shape[0] == 1 else 1
            _chk(X0aux, (B if x0s else 1, self.n, self.r), "X0aux", dev)
        res = {}
        dX = dU = None
        if want_traj:
            if out is not None:
                dX, dU = out["dX"], out["dU"]
            else:
                dX = torch.empty((B, H + 1, self.n, self.r), dtype=torch.float64, device=dev)
                dU = torch.empty((B, H, self.m, self.r), dtype=torch.float64, device=dev)
            res["dX"], res["dU"] = dX, dU
        ldp = None
        if Xref is not None:
            _chk(Xref, (B, H + 1, self.n), "Xref", dev)
            if Uref is not None:
                _chk(Uref, (B, H, self.m), "Uref", dev)
            ldp = out["loss_dp"] if (out is not None and "loss_dp" in out) else \
                torch.empty((B, self.r + 1), dtype=torch.float64, device=dev)
            res["loss_dp"] = ldp
        ws, need = self._workspace(backend.OP_AUX_LQR, B, H, dev)
        lib = self.handle.lib
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            if phase == "backward":
                backend.check(lib.pdp_aux_lqr_backward(self.handle.ptr, B, H, _ptr(X), _ptr(U), _ptr(Lam), _ptr(theta), ts,
                                                       _ptr(ws), ws.numel(), _ptr(status), st), "pdp_aux_lqr_backward")
            elif phase == "forward":
                backend.check(lib.pdp_aux_lqr_forward(self.handle.ptr, B, H, _ptr(X), _ptr(U), _ptr(theta), ts,
                                                      _ptr(X0aux), x0s, _ptr(dX), _ptr(dU), _ptr(Xref), _ptr(Uref),
                                                      _ptr(ldp), _ptr(ws), ws.numel(), _ptr(status), st),
                              "pdp_aux_lqr_forward")
            else:
                backend.check(lib.pdp_aux_lqr(self.handle.ptr, B, H, _ptr(X), _ptr(U), _ptr(Lam), _ptr(theta), ts,
                                              _ptr(X0aux), x0s, _ptr(dX), _ptr(dU), _ptr(Xref), _ptr(Uref), _ptr(ldp),
                                              _ptr(ws), ws.numel(), _ptr(status), st), "pdp_aux_lqr")
        return res

    def set_sweep_parts(self, parts: int):
        """Sub-batches ``pdp_sweep`` cuts its aux-LQR phase into (two internal streams): 0 = automatic, 1 = never."""
        backend.check(self.handle.lib.pdp_set_sweep_parts(self.handle.ptr, int(parts)), "pdp_set_sweep_parts")

    def sweep(self, x0, theta, U, Xref=None, Uref=None, want_traj=True, status=None, out=None):
        """One PDP sweep (the BASELINE metric's unit): rollout + costate + fused aux-LQR.  Large batches are pipelined
        inside the C ABI (sub-batches on two internal streams; identical results, see include/pdp_b200.h)."""
        require_cuda()
        dev = x0.device
        B, H = U.shape[0], U.shape[1]
        _chk(x0, (B, self.n), "x0", dev)
        _chk(U, (B, H, self.m), "U", dev)
        theta, ts = self._theta(theta, B, dev)

        def buf(name, shape):
            if out is not None and name in out:
                return out[name]
            return torch.empty(shape, dtype=torch.float64, device=dev)

        X = buf("X", (B, H + 1, self.n))
        Lam = buf("Lam", (B, H, self.n))
        cost = buf("cost", (B,))
        dX = buf("dX", (B, H + 1, self.n, self.r)) if want_traj else None
        dU = buf("dU", (B, H, self.m, self.r)) if want_traj else None
        ldp = buf("loss_dp", (B, self.r + 1)) if Xref is not None else None
        ws, need = self._workspace(backend.OP_SWEEP, B, H, dev)
        lib = self.handle.lib
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            backend.check(lib.pdp_sweep(self.handle.ptr, B, H, _ptr(x0), _ptr(theta), ts, _ptr(U), _ptr(X), _ptr(Lam),
                                        _ptr(cost), _ptr(dX), _ptr(dU), _ptr(Xref), _ptr(Uref), _ptr(ldp), _ptr(ws),
                                        ws.numel(), _ptr(status), st), "pdp_sweep")
        res = {"X": X, "Lam": Lam, "cost": cost}
        if want_traj:
            res["dX"], res["dU"] = dX, dU
        if ldp is not None:
            res["loss_dp"] = ldp
        return res

    def sweep_host(self, x0_h, theta_h, U_h, Xref_h, Uref_h, loss_dp_h, cost_h=None, keep_dtraj=True, n_chunks=4,
                   device=None):
        """End-to-end sweep from PINNED HOST tensors: per sub-batch H2D of (x0, theta, U, Xref, Uref) -> rollout /
        costate / fused aux-LQR kernels -> D2H of (loss, dp) [and cost], through the C-ABI ``pdp_sweep_host``.
        The batch is cut into ``n_chunks`` sub-batches issued alternately on two side streams so that the copies
        of one sub-batch overlap the kernels of the other; the call returns with the work ordered into the
        current stream (synchronise that stream before reading ``loss_dp_h``)."""
        require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        B, H = U_h.shape[0], U_h.shape[1]
        for t_, nm_ in ((x0_h, "x0"), (theta_h, "theta"), (U_h, "U"), (Xref_h, "Xref"), (loss_dp_h, "loss_dp")):
            if not (t_.is_pinned() and t_.is_contiguous() and t_.dtype == torch.float64):
                raise ValueError("sweep_host: %s must be a pinned contiguous float64 host tensor" % nm_)
        ts = 0 if theta_h.shape[0] == 1 else self.r
        n_chunks = max(1, min(int(n_chunks), B))
        bounds = [(B * c) // n_chunks for c in range(n_chunks + 1)]
        sizes = [bounds[c + 1] - bounds[c] for c in range(n_chunks)]
        per = [self.handle.workspace_bytes(backend.OP_SWEEP_HOST, sz, H) for sz in sizes]
        key = ("host", dev)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < sum(per):
            ws = torch.empty(sum(per), dtype=torch.uint8, device=dev)
            self._ws[key] = ws
        if getattr(self, "_side", None) is None:
            self._side = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
        cur = torch.cuda.current_stream(dev)
        start = torch.cuda.Event()
        start.record(cur)
        lib = self.handle.lib
        off = 0
        el = 8
        with torch.cuda.device(dev):
            for c in range(n_chunks):
                lo, sz = bounds[c], sizes[c]
                st = self._side[c % 2] if n_chunks > 1 else cur
                if n_chunks > 1 and c < 2:
                    st.wait_event(start)
                backend.check(lib.pdp_sweep_host(
                    self.handle.ptr, sz, H, x0_h.data_ptr() + lo * self.n * el,
                    theta_h.data_ptr() + (lo * self.r * el if ts else 0), ts, U_h.data_ptr() + lo * H * self.m * el,
                    Xref_h.data_ptr() + lo * (H + 1) * self.n * el,
                    (Uref_h.data_ptr() + lo * H * self.m * el) if Uref_h is not None else None,
                    loss_dp_h.data_ptr() + lo * (self.r + 1) * el,
                    (cost_h.data_ptr() + lo * el) if cost_h is not None else None, 1 if keep_dtraj else 0,
                    ws.data_ptr() + off, per[c], st.cuda_stream), "pdp_sweep_host")
                off += per[c]
            if n_chunks > 1:
                for st in self._side:
                    done = torch.cuda.Event()
                    done.record(st)
                    cur.wait_event(done)

    def aux_eval(self, X, U, Lam, theta):
        """Dense auxiliary matrices (legacy getAuxSys return value) as a dict of [B,H,...] tensors."""
        require_cuda()
        dev = X.device
        B, H = U.shape[0], U.shape[1]
        n, m, r = self.n, self.m, self.r
        theta, ts = self._theta(theta, B, dev)
        nd = 2 * (n * n + n * m + n * r) + m * n + m * m + m * r
        aux = torch.empty((B, H, nd), dtype=torch.float64, device=dev)
        term = torch.empty((B, n * n + n * r), dtype=torch.float64, device=dev)
        lib = self.handle.lib
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            backend.check(lib.pdp_aux_eval(self.handle.ptr, B, H, _ptr(X), _ptr(U), _ptr(Lam), _ptr(theta), ts,
                                           _ptr(aux), _ptr(term), st), "pdp_aux_eval")
        out, o = {}, 0
        for name, (a, b_) in (("dynF", (n, n)), ("dynG", (n, m)), ("dynE", (n, r)), ("Hxx", (n, n)), ("Hxu", (n, m)),
                              ("Hxe", (n, r)), ("Hux", (m, n)), ("Huu", (m, m)), ("Hue", (m, r))):
            out[name] = aux[:, :, o:o + a * b_].reshape(B, H, a, b_)
            o += a * b_
        out["hxx"] = term[:, :n * n].reshape(B, n, n)
        out["hxe"] = term[:, n * n:].reshape(B, n, r)
        return out


class _SensSystem:
    """Common driver of the forward-sensitivity modules (SysID / ControlPlanning)."""

    def __init__(self, src, verbose=False):
        self.src = src
        self.n, self.m, self.r = src.n, src.m, src.r
        self.module_path = build.compile_module(src.source(), src.key(), verbose=verbose)
        self._handle = None

    @property
    def handle(self):
        if self._handle is None:
            self._handle = backend.SystemHandle(self.module_path)
        return self._handle

    def _run(self, B, H, x0, theta, inputs, Xobs, want_traj, want_sens, want_loss, status, dev, has_policy):
        theta2 = theta.unsqueeze(0) if theta.dim() == 1 else theta
        ts = 0 if theta2.shape[0] == 1 else self.r
        _chk(theta2, (B if ts else 1, self.r), "theta", dev)
        _chk(x0, (B, self.n), "x0", dev)
        mk = lambda *shape: torch.empty(shape, dtype=torch.float64, device=dev)
        X = mk(B, H + 1, self.n) if want_traj else None
        Uout = mk(B, H, self.m) if (want_traj and has_policy) else None
        dX = mk(B, H + 1, self.n, self.r) if want_sens else None
        dU = mk(B, H, self.m, self.r) if (want_sens and has_policy) else None
        ldp = mk(B, self.r + 1) if want_loss else None
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            backend.check(self.handle.lib.pdp_sens_fwd(self.handle.ptr, B, H, _ptr(x0), _ptr(theta2), ts, _ptr(inputs),
                                                       _ptr(Xobs), _ptr(X), _ptr(Uout), _ptr(dX), _ptr(dU), _ptr(ldp),
                                                       _ptr(status), st), "pdp_sens_fwd")
        out = {}
        for k, v in (("X", X), ("U", Uout), ("dX", dX), ("dU", dU), ("loss_dp", ldp)):
            if v is not None:
                out[k] = v
        return out


class SysIDSystem(_SensSystem):
    """Fused SysID.step (reference PDP/PDP.py:1261-1296): rollout + X+ = F X + E + loss / half-gradient."""

    def __init__(self, state, control, auxvar, dyn, verbose=False, **kw):
        from . import codegen_sens
        kw.setdefault("max_group_cols", 3)     # measured on B200 (profiles/r1d_secondary_configs.json): 2 groups of <= 3
        super().__init__(codegen_sens.SensModuleSource(codegen_sens.KIND_SYSID, state, control, auxvar, dyn, **kw), verbose)

    def step(self, inputs, Xobs, theta, x0=None, want_traj=False, want_sens=False, status=None):
        """inputs[B,H,m], Xobs[B,H+1,n] (x0 defaults to Xobs[:,0]) -> dict with loss_dp[B,r+1] (+X, dX)."""
        require_cuda()
        dev = inputs.device
        B, H = inputs.shape[0], inputs.shape[1]
        _chk(inputs, (B, H, self.m), "inputs", dev)
        if Xobs is not None:
            _chk(Xobs, (B, H + 1, self.n), "Xobs", dev)
        if x0 is None:
            x0 = Xobs[:, 0, :].contiguous()
        return self._run(B, H, x0, theta, inputs, Xobs, want_traj, want_sens, Xobs is not None, status, dev, False)


class CPSystem(_SensSystem):
    """Fused ControlPlanning.step (reference PDP/PDP.py:850-878) for a parameterised policy."""

    def __init__(self, state, control, auxvar, dyn, policy, tvar, path_cost, final_cost, verbose=False, **kw):
        from . import codegen_sens
        super().__init__(codegen_sens.SensModuleSource(codegen_sens.KIND_CP, state, control, auxvar, dyn, policy=policy,
                                                       tvar=tvar, path_cost=path_cost, final_cost=final_cost, **kw), verbose)

    def step(self, x0, H, theta, want_traj=False, want_sens=False, status=None):
        """x0[B,n], theta[B|1,r] -> loss_dp[B,r+1] = (cost, dcost/dtheta) (+ X, U, dX, dU)."""
        require_cuda()
        dev = x0.device
        B = x0.shape[0]
        return self._run(B, int(H), x0, theta, None, None, want_traj, want_sens, True, status, dev, True)


class DenseLQR:
    """Generic (n, m, r) matrix LQR on caller-supplied matrices (the drop-in ``LQR.lqrSolver``)."""

    _cache = {}

    def __init__(self, n, m, r, verbose=False):
        self.n, self.m, self.r = int(n), int(m), int(r)
        if self.n + self.m + self.r > 32:
            raise ValueError("DenseLQR handles n+m+r <= 32 per launch; split the auxvar columns (see LQR.lqrSolver)")
        self.src = codegen.LQRModuleSource(self.n, self.m, self.r)
        self.module_path = build.compile_module(self.src.source(), self.src.key(), verbose=verbose)
        self._handle = None
        self._ws = None

    @classmethod
    def get(cls, n, m, r):
        key = (int(n), int(m), int(r))
        if key not in cls._cache:
            cls._cache[key] = cls(*key)
        return cls._cache[key]

    @property
    def handle(self):
        if self._handle is None:
            self._handle = backend.SystemHandle(self.module_path)
        return self._handle

    @property
    def ndense(self):
        n, m, r = self.n, self.m, self.r
        return 2 * (n * n + n * m + n * r) + m * n + m * m + m * r

    def solve(self, aux, term, X0aux=None, status=None, gains=None):
        """aux[B,H,NDENSE], term[B,n*n+n*r] -> Xaux[B,H+1,n,r], Uaux[B,H,m,r].
        ``gains[B,H,n+r,m]`` given => forward-only recursion with those gains (no Riccati sweep)."""
        require_cuda()
        dev = aux.device
        B, H = aux.shape[0], aux.shape[1]
        n, m, r = self.n, self.m, self.r
        _chk(aux, (B, H, self.ndense), "aux", dev)
        if term is not None:
            _chk(term, (B, n * n + n * r), "term", dev)
        x0s = 0
        if X0aux is not None:
            if X0aux.dim() == 2:
                X0aux = X0aux.unsqueeze(0)
            x0s = 0 if X0aux.shape[0] == 1 else 1
            _chk(X0aux, (B if x0s else 1, n, r), "X0aux", dev)
        Xa = torch.empty((B, H + 1, n, r), dtype=torch.float64, device=dev)
        Ua = torch.empty((B, H, m, r), dtype=torch.float64, device=dev)
        need = self.handle.workspace_bytes(backend.OP_AUX_LQR, B, H)
        if gains is not None:
            _chk(gains, (B, H, n + r, m), "gains", dev)
            ws = gains.view(torch.uint8).reshape(-1)
            if ws.numel() < need:
                ws = torch.cat([ws, torch.zeros(need - ws.numel(), dtype=torch.uint8, device=dev)])
        else:
            if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
                self._ws = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
            ws = self._ws
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            backend.check(self.handle.lib.pdp_lqr_dense(self.handle.ptr, B, H, _ptr(aux), _ptr(term), _ptr(X0aux), x0s,
                                                        _ptr(Xa), _ptr(Ua), 1 if gains is not None else 0, _ptr(ws),
                                                        ws.numel(), _ptr(status), st), "pdp_lqr_dense")
        return Xa, Ua


class GpuFunction:
    """A symbolic ``Function`` compiled to a batched CUDA kernel (one thread per sample)."""

    def __init__(self, fn, verbose=False):
        self.fn = fn
        self.src = codegen.FunctionModuleSource(fn)
        self.module_path = build.compile_module(self.src.source(), self.src.key(), verbose=verbose)
        self._handle = None

    @property
    def handle(self):
        if self._handle is None:
            self._handle = backend.SystemHandle(self.module_path)
        return self._handle

    def __call__(self, *args):
        """args[k]: CUDA float64 tensor [B, numel_k] or [numel_k] (shared).  Returns a list of
        [B, rows, cols] tensors.  Input elements are in the Function's column-major element order."""
        import ctypes
        require_cuda()
        fn = self.fn
        if len(args) != fn.n_in():
            raise TypeError("GpuFunction: expected %d inputs" % fn.n_in())
        B = max([a.shape[0] for a in args if a.dim() == 2] + [1])
        dev = args[0].device
        ins, strides = [], []
        for k, a in enumerate(args):
            ne = fn.numel_in(k)
            a2 = a if a.dim() == 2 else a.unsqueeze(0)
            _chk(a2, (a2.shape[0], ne), "input %d" % k, dev)
            if a2.shape[0] not in (1, B):
                raise ValueError("GpuFunction: inconsistent batch sizes")
            ins.append(a2)
            strides.append(ne if a2.shape[0] == B and B > 1 else (ne if B == 1 else 0))
        outs = [torch.empty((B,) + tuple(fn.size_out(k)), dtype=torch.float64, device=dev) for k in range(fn.n_out())]
        PtrArr = ctypes.c_void_p * max(len(ins), 1)
        OutArr = ctypes.c_void_p * max(len(outs), 1)
        StrArr = ctypes.c_int * max(len(ins), 1)
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            backend.check(self.handle.lib.pdp_eval_function(self.handle.ptr, B, PtrArr(*[t.data_ptr() for t in ins]),
                                                            StrArr(*strides), OutArr(*[t.data_ptr() for t in outs]), st),
                          "pdp_eval_function")
        return outs
