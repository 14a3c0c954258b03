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


_RED_WS = {}


def reduce_loss_dp(loss_dp, out=None):
    """loss_dp[B, r+1] (CUDA float64) -> sums[r+2] = (sum loss, sum dp[r], B): the batch reduction of the outer loops
    (reference PDP/PDP.py:1293-1294, Examples/IRL/quadrotor/uav_PDP.py:78-81) as ONE deterministic kernel of the C ABI
    (``pdp_reduce_loss_dp``); the result is what a multi-GPU run all-reduces.  Stream-ordered, graph-capturable."""
    require_cuda()
    dev = loss_dp.device
    B, r1 = loss_dp.shape
    _chk(loss_dp, (B, r1), "loss_dp", dev)
    lib = backend.load_library()
    st = torch.cuda.current_stream(dev)
    key = (dev, r1, st.cuda_stream)
    ws = _RED_WS.get(key)
    if ws is None:
        ws = torch.zeros(int(lib.pdp_reduce_workspace_bytes(r1 - 1)), dtype=torch.uint8, device=dev)   # ticket word = 0
        _RED_WS[key] = ws
    if out is None:
        out = torch.empty(r1 + 1, dtype=torch.float64, device=dev)
    else:
        _chk(out, (r1 + 1,), "out", dev)
    with torch.cuda.device(dev):
        backend.check(lib.pdp_reduce_loss_dp(B, r1 - 1, _ptr(loss_dp), _ptr(out), _ptr(ws), ws.numel(), st.cuda_stream),
                      "pdp_reduce_loss_dp")
    return out


def _host_chk(t, shape, name):
    if not (t.is_pinned() and t.is_contiguous() and t.dtype == torch.float64):
        raise ValueError("%s must be a pinned contiguous float64 host tensor" % name)
    if tuple(t.shape) != tuple(shape):
        raise ValueError("%s has shape %s, expected %s" % (name, tuple(t.shape), tuple(shape)))


class _ChunkedHostCall:
    """Shared driver of the *_host entry points: the batch is cut into ``n_chunks`` sub-batches issued alternately on two
    side streams (copies of one sub-batch overlap the kernels of the other); the call returns with the work ordered
    into the current stream."""

    def __init__(self):
        self._side = {}

    def run(self, dev, B, n_chunks, issue):
        n_chunks = max(1, min(int(n_chunks), B))
        bounds = [(B * c) // n_chunks for c in range(n_chunks + 1)]
        cur = torch.cuda.current_stream(dev)
        if n_chunks == 1:
            with torch.cuda.device(dev):
                issue(0, 0, B, cur)
            return
        side = self._side.get(dev)
        if side is None:
            side = self._side[dev] = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
        start = torch.cuda.Event()
        start.record(cur)
        with torch.cuda.device(dev):
            for c in range(n_chunks):
                st = side[c % 2]
                if c < 2:
                    st.wait_event(start)
                issue(c, bounds[c], bounds[c + 1] - bounds[c], st)
            for st in side:
                done = torch.cuda.Event()
                done.record(st)
                cur.wait_event(done)


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
        key = (op, device, torch.cuda.current_stream(device).cuda_stream)
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
        _chk(x0, (B, self.n), "x0", dev)
        _chk(Uref, (B, H, self.m), "Uref", dev)
        _chk(Xref, (B, H + 1, self.n), "Xref", dev)
        _chk(alpha, (Bx,), "alpha", dev)
        if not (gains.is_cuda and gains.is_contiguous() and gains.device == dev and
                gains.numel() * gains.element_size() >= B * H * (self.n + 1) * self.m * 8):
            raise ValueError("gains must be a contiguous CUDA buffer of at least B*H*(n+1)*m float64 values")
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
        if Xref is not None:
            _chk(Xref, (B, H + 1, self.n), "Xref", dev)
        if Uref is not None:
            if Xref is None:
                raise ValueError("sweep: Uref needs Xref")
            _chk(Uref, (B, H, self.m), "Uref", dev)

        def buf(name, shape):
            if out is not None and name in out:
                _chk(out[name], shape, name, dev)
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

    def _host_ws(self, op, sizes, H, dev, zero=False):
        per = [self.handle.workspace_bytes(op, sz, H) for sz in sizes]
        key = ("host", op, dev, torch.cuda.current_stream(dev).cuda_stream)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < sum(per):
            ws = (torch.zeros if zero else torch.empty)(sum(per), dtype=torch.uint8, device=dev)
            self._ws[key] = ws
        return ws, per

    def sweep_host(self, x0_h, theta_h, U_h, Xref_h, Uref_h, loss_dp_h, cost_h=None, keep_dtraj=True, n_chunks=4,
                   device=None, X_h=None, Lam_h=None, dX_h=None, dU_h=None):
        """End-to-end sweep from PINNED HOST tensors: per sub-batch H2D of (x0, theta, U, Xref, Uref) -> rollout /
        costate / fused aux-LQR kernels -> D2H of (loss, dp) [and cost], through the C-ABI ``pdp_sweep_host``.
        With any of ``X_h, Lam_h, dX_h, dU_h`` (pinned host tensors) the trajectories / sensitivities themselves come
        back as well (``pdp_sweep_host_traj``).  The batch is cut into ``n_chunks`` sub-batches issued alternately on two
        side streams so that the copies of one sub-batch overlap the kernels of the other; the call returns with the
        work ordered into the current stream (synchronise that stream before reading the host outputs)."""
        require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        B, H = U_h.shape[0], U_h.shape[1]
        n, m, r = self.n, self.m, self.r
        ts = 0 if (theta_h.dim() == 1 or theta_h.shape[0] == 1) else r
        _host_chk(x0_h, (B, n), "x0")
        _host_chk(theta_h, (B, r) if ts else tuple(theta_h.shape), "theta")
        if theta_h.numel() != (B * r if ts else r):
            raise ValueError("theta has %d elements, expected %d" % (theta_h.numel(), B * r if ts else r))
        _host_chk(U_h, (B, H, m), "U")
        _host_chk(Xref_h, (B, H + 1, n), "Xref")
        _host_chk(loss_dp_h, (B, r + 1), "loss_dp")
        for t_, shp, nm_ in ((Uref_h, (B, H, m), "Uref"), (cost_h, (B,), "cost"), (X_h, (B, H + 1, n), "X"),
                             (Lam_h, (B, H, n), "Lam"), (dX_h, (B, H + 1, n, r), "dX"), (dU_h, (B, H, m, r), "dU")):
            if t_ is not None:
                _host_chk(t_, shp, nm_)
        traj = any(t_ is not None for t_ in (X_h, Lam_h, dX_h, dU_h))
        n_chunks = max(1, min(int(n_chunks), B))
        sizes = [(B * (c + 1)) // n_chunks - (B * c) // n_chunks for c in range(n_chunks)]
        ws, per = self._host_ws(backend.OP_SWEEP_HOST, sizes, H, dev)
        offs = [sum(per[:c]) for c in range(n_chunks)]
        lib, el = self.handle.lib, 8
        at = lambda t_, lo, row: None if t_ is None else t_.data_ptr() + lo * row * el

        def issue(c, lo, sz, st):
            common = (self.handle.ptr, sz, H, at(x0_h, lo, n), theta_h.data_ptr() + (lo * r * el if ts else 0), ts,
                      at(U_h, lo, H * m), at(Xref_h, lo, (H + 1) * n), at(Uref_h, lo, H * m), at(loss_dp_h, lo, r + 1),
                      at(cost_h, lo, 1))
            if traj:
                backend.check(lib.pdp_sweep_host_traj(*common, at(X_h, lo, (H + 1) * n), at(Lam_h, lo, H * n),
                                                      at(dX_h, lo, (H + 1) * n * r), at(dU_h, lo, H * m * r),
                                                      ws.data_ptr() + offs[c], per[c], st.cuda_stream), "pdp_sweep_host_traj")
            else:
                backend.check(lib.pdp_sweep_host(*common, 1 if keep_dtraj else 0, ws.data_ptr() + offs[c], per[c],
                                                 st.cuda_stream), "pdp_sweep_host")

        if getattr(self, "_hostcall", None) is None:
            self._hostcall = _ChunkedHostCall()
        self._hostcall.run(dev, B, n_chunks, issue)

    def rollout_costate_host(self, x0_h, theta_h, U_h, cost_h=None, dHu_h=None, X_h=None, Lam_h=None, n_chunks=4, device=None):
        """End-to-end rollout + costate (+ adjoint gradient dH/du) from PINNED HOST tensors through the C-ABI
        ``pdp_rollout_costate_host`` (the batched ``ControlPlanning.recmat_step``, reference PDP/PDP.py:1100-1114);
        same sub-batch / two-stream scheme as :meth:`sweep_host`."""
        require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        B, H = U_h.shape[0], U_h.shape[1]
        n, m = self.n, self.m
        nth = theta_h.shape[-1]
        ts = 0 if (theta_h.dim() == 1 or theta_h.shape[0] == 1) else nth
        _host_chk(x0_h, (B, n), "x0")
        _host_chk(theta_h, tuple(theta_h.shape), "theta")
        _host_chk(U_h, (B, H, m), "U")
        for t_, shp, nm_ in ((cost_h, (B,), "cost"), (dHu_h, (B, H, m), "dHu"), (X_h, (B, H + 1, n), "X"), (Lam_h, (B, H, n), "Lam")):
            if t_ is not None:
                _host_chk(t_, shp, nm_)
        n_chunks = max(1, min(int(n_chunks), B))
        sizes = [(B * (c + 1)) // n_chunks - (B * c) // n_chunks for c in range(n_chunks)]
        ws, per = self._host_ws(backend.OP_ROLLOUT_HOST, sizes, H, dev)
        offs = [sum(per[:c]) for c in range(n_chunks)]
        lib, el = self.handle.lib, 8
        at = lambda t_, lo, row: None if t_ is None else t_.data_ptr() + lo * row * el

        def issue(c, lo, sz, st):
            backend.check(lib.pdp_rollout_costate_host(
                self.handle.ptr, sz, H, at(x0_h, lo, n), theta_h.data_ptr() + (lo * nth * el if ts else 0), ts,
                at(U_h, lo, H * m), at(cost_h, lo, 1), at(dHu_h, lo, H * m), at(X_h, lo, (H + 1) * n), at(Lam_h, lo, H * n),
                ws.data_ptr() + offs[c], per[c], st.cuda_stream), "pdp_rollout_costate_host")

        if getattr(self, "_hostcall", None) is None:
            self._hostcall = _ChunkedHostCall()
        self._hostcall.run(dev, B, n_chunks, issue)

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


    def step_host(self, x0_h, theta_h, H, inputs_h=None, Xobs_h=None, loss_dp_h=None, sums_h=None, n_chunks=4, device=None):
        """End-to-end fused step from PINNED HOST tensors through the C-ABI ``pdp_sens_fwd_host``: H2D of x0 / theta (and
        inputs / observed states for SysID) -> ``pdp_k_sens_fwd`` -> D2H of ``loss_dp_h[B,r+1]`` and / or of the per-sub-batch
        reductions ``sums_h[n_chunks, r+2]`` = (sum loss, sum dp, count) -- add the rows and divide by the count for the
        batch mean of reference PDP/PDP.py:1293-1294."""
        require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        B = x0_h.shape[0]
        n, m, r = self.n, self.m, self.r
        ts = 0 if (theta_h.dim() == 1 or theta_h.shape[0] == 1) else r
        _host_chk(x0_h, (B, n), "x0")
        _host_chk(theta_h, (B, r) if ts else tuple(theta_h.shape), "theta")
        for t_, shp, nm_ in ((inputs_h, (B, H, m), "inputs"), (Xobs_h, (B, H + 1, n), "Xobs"), (loss_dp_h, (B, r + 1), "loss_dp")):
            if t_ is not None:
                _host_chk(t_, shp, nm_)
        n_chunks = max(1, min(int(n_chunks), B))
        if sums_h is not None:
            _host_chk(sums_h, (n_chunks, r + 2), "sums")
        if loss_dp_h is None and sums_h is None:
            raise ValueError("step_host: give loss_dp_h and / or sums_h")
        sizes = [(B * (c + 1)) // n_chunks - (B * c) // n_chunks for c in range(n_chunks)]
        per = [self.handle.workspace_bytes(backend.OP_SENS_HOST, sz, H) for sz in sizes]
        key = (dev, torch.cuda.current_stream(dev).cuda_stream)
        if not hasattr(self, "_hws"):
            self._hws, self._hostcall = {}, _ChunkedHostCall()
        ws = self._hws.get(key)
        if ws is None or ws.numel() < sum(per):
            ws = self._hws[key] = torch.zeros(sum(per), dtype=torch.uint8, device=dev)      # reduction tickets start at zero
        offs = [sum(per[:c]) for c in range(n_chunks)]
        lib, el = self.handle.lib, 8
        at = lambda t_, lo, row: None if t_ is None else t_.data_ptr() + lo * row * el

        def issue(c, lo, sz, st):
            backend.check(lib.pdp_sens_fwd_host(
                self.handle.ptr, sz, H, at(x0_h, lo, n), theta_h.data_ptr() + (lo * r * el if ts else 0), ts,
                at(inputs_h, lo, H * m), at(Xobs_h, lo, (H + 1) * n), at(loss_dp_h, lo, r + 1),
                None if sums_h is None else sums_h.data_ptr() + c * (r + 2) * el,
                ws.data_ptr() + offs[c], per[c], st.cuda_stream), "pdp_sens_fwd_host")

        self._hostcall.run(dev, B, n_chunks, issue)


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
        # column groups of <= 3 (measured at C2, B = 4096: 0.174 ms vs 0.201 ms in one group, all outputs written); large
        # parameter vectors (neural policies) keep groups of <= 12 so that a thread's columns fit its shared-memory slice
        kw.setdefault("max_group_cols", 3 if auxvar.numel() <= 12 else 12)
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
