"""Forward-sensitivity modules: ``SysID.step`` and ``ControlPlanning.step`` as fused CUDA kernels.

Reference semantics restated:

* SysID (``PDP/PDP.py:1209-1296``): rollout with given inputs, ``X_{t+1} = F_t X_t + E_t`` with
  ``F = df/dx, E = df/dtheta``, ``X_0 = 0``; ``loss = ||x - xobs||^2``, ``dp = sum_t (x_t - xobs_t) X_t``
  (half the gradient, like the reference).
* ControlPlanning (``PDP/PDP.py:763-878``): rollout under the policy ``u_t = pi(t, x_t, theta)``,
  ``U_t = Ux_t X_t + Ue_t``, ``X_{t+1} = F_t X_t + G_t U_t``; ``loss = sum c + h``,
  ``dtheta = sum_t (c_x X_t + c_u U_t) + h_x X_H``.

Kernel design (DESIGN.md "pdp_k_sens_fwd"): a block owns 32 trajectories (lane = trajectory) and has one WARP PER
COLUMN GROUP.  The columns of the sensitivity matrix ``X_t (n x r)`` evolve independently, so warp ``g`` keeps its
group's columns in shared memory laid out ``[element][lane]`` and streams them through registers one column at a
time; the (cheap) rollout is repeated by every warp, everything else (Jacobian entries, policy derivatives) is
straight-line generated code over structural non-zeros.  All warps of a block share ONE set of staged input rows
(SysID: ``[inputs_t | xobs_t]``, brought in one chunk of time steps ahead by warp-cooperative coalesced ``cp.async``
copies -- see ``kernel_templates.K_STAGING``) and, when trajectories / sensitivities are requested, one output tile
that leaves through coalesced write-outs, so column groups add parallelism (small batches, large r) without
re-reading anything from HBM.  Two entry points are generated from one body: ``pdp_k_sens_fwd`` (fused loss /
gradient only) and ``pdp_k_sens_fwd_out`` (also X, U, dX/dtheta, dU/dtheta).
"""
from __future__ import annotations

import hashlib
from typing import Dict, List, Optional, Sequence

from . import symbolic as S
from .symbolic import SX, Node
from .codegen import _lit

KIND_SYSID, KIND_CP = 2, 3

_BODY_HEAD = r"""
// One block = 32 trajectories (lane = trajectory) x PDP_NG warps (warp = column group of the sensitivity matrix).
template <bool OUT>
__device__ __forceinline__ void pdp_sens_body(int B, int H, const double* __restrict__ x0, const double* __restrict__ theta,
               int theta_stride, const double* __restrict__ inputs, const double* __restrict__ Xobs, double* __restrict__ X,
               double* __restrict__ Uout, double* __restrict__ dX, double* __restrict__ dU, double* __restrict__ loss_dp,
               int* __restrict__ status)
{
  extern __shared__ __align__(16) double pdp_smem[];
  constexpr int RC = OUT ? PDP_RCO : PDP_RCF;
  // tile rows: input tile [inputs RC*m | xobs RC*n]; output tile [X (RC+1)*n | U RC*m | dX (RC+1)*n*r | dU RC*m*r]
  constexpr int PDP_IU = 0, PDP_IX = RC * PDP_M, IN_ROWS = RC * PDP_RIN;
  constexpr int PDP_OX = 0, PDP_OU = (RC + 1) * PDP_N, PDP_ODX = PDP_OU + (PDP_CP ? RC * PDP_M : 0),
                PDP_ODU = PDP_ODX + (RC + 1) * PDP_N * PDP_R;
  (void)PDP_IU; (void)PDP_IX; (void)PDP_OX; (void)PDP_OU; (void)PDP_ODX; (void)PDP_ODU;
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int b0 = blockIdx.x * 32;
  if (b0 >= B) return;
  const bool live = b0 + lane < B;
  const int b = live ? b0 + lane : B - 1;                  // tail lanes shadow a valid trajectory
  const int nvalid = B - b0 < 32 ? B - b0 : 32;
  double* ST = pdp_smem + (size_t)grp * (PDP_GMAX * PDP_N * 32) + lane;         // this warp's columns, [element][lane]
  double* in_tiles = pdp_smem + PDP_STATE_DOUBLES;
  double* out_tile = in_tiles + (size_t)2 * IN_ROWS * PDP_TLD;
  double* OT = out_tile + lane;
  (void)OT; (void)nvalid;
  const double* th = theta + (size_t)b * theta_stride;
  (void)th;
  for (int k = 0; k < PDP_GMAX * PDP_N; ++k) ST[k * 32] = 0.0;
  double loss = 0.0;"""

_FIRST_CHUNK = r"""  // first chunk of input rows; the copies of a chunk are split over the warps of the block (trajectory j by warp j % NG)
  {
    const int nn = H < RC ? H : RC;
    pdp_stage_in(in_tiles, PDP_IU, inputs, (size_t)H * PDP_M, 0, nn * PDP_M, b0, B, 1, lane, grp, PDP_NG);
    if (Xobs) pdp_stage_in(in_tiles, PDP_IX, Xobs, (size_t)(H + 1) * PDP_N, 0, nn * PDP_N, b0, B, 1, lane, grp, PDP_NG);
  }"""

_LOOP_HEAD = r"""  int buf = 0;
  #pragma unroll 1
  for (int t0 = 0; t0 < H; t0 += RC, buf ^= 1) {
    const int nst = H - t0 < RC ? H - t0 : RC;
    const bool last = t0 + RC >= H;
    const double* INT = in_tiles + (size_t)buf * IN_ROWS * PDP_TLD + lane;
    (void)INT;
    pdp_cp_async_wait_all();
    __syncthreads();       // this chunk's rows have landed (all warps' copies); the previous chunk's write-out is finished"""

_NEXT_CHUNK = r"""    if (!last) {
      const int nn = H - t0 - RC < RC ? H - t0 - RC : RC;
      double* nt = in_tiles + (size_t)(buf ^ 1) * IN_ROWS * PDP_TLD;
      pdp_stage_in(nt, PDP_IU, inputs, (size_t)H * PDP_M, (size_t)(t0 + RC) * PDP_M, nn * PDP_M, b0, B, 1, lane, grp, PDP_NG);
      if (Xobs) pdp_stage_in(nt, PDP_IX, Xobs, (size_t)(H + 1) * PDP_N, (size_t)(t0 + RC) * PDP_N, nn * PDP_N, b0, B, 1, lane, grp, PDP_NG);
    }"""

_WRITE_OUT = r"""    if (OUT) {
      __syncthreads();     // every warp has put its columns of this chunk into the output tile
      const int nrow = nst + (last ? 1 : 0);
      if (X) pdp_stage_out(out_tile, PDP_OX, X, (size_t)(H + 1) * PDP_N, (size_t)t0 * PDP_N, nrow * PDP_N, b0, nvalid, lane, grp, PDP_NG);
      if (dX) pdp_stage_out(out_tile, PDP_ODX, dX, (size_t)(H + 1) * PDP_N * PDP_R, (size_t)t0 * PDP_N * PDP_R, nrow * PDP_N * PDP_R, b0, nvalid, lane, grp, PDP_NG);
#if PDP_CP
      if (Uout) pdp_stage_out(out_tile, PDP_OU, Uout, (size_t)H * PDP_M, (size_t)t0 * PDP_M, nst * PDP_M, b0, nvalid, lane, grp, PDP_NG);
      if (dU) pdp_stage_out(out_tile, PDP_ODU, dU, (size_t)H * PDP_M * PDP_R, (size_t)t0 * PDP_M * PDP_R, nst * PDP_M * PDP_R, b0, nvalid, lane, grp, PDP_NG);
#endif
    }
  }"""

_ENTRY_POINTS = r"""
extern "C" __global__ void __launch_bounds__(PDP_NG * 32, PDP_MINB)
pdp_k_sens_fwd(int B, int H, const double* __restrict__ x0, const double* __restrict__ theta, int theta_stride,
               const double* __restrict__ inputs, const double* __restrict__ Xobs, double* __restrict__ loss_dp,
               int* __restrict__ status)
{
  pdp_sens_body<false>(B, H, x0, theta, theta_stride, inputs, Xobs, nullptr, nullptr, nullptr, nullptr, loss_dp, status);
}

extern "C" __global__ void __launch_bounds__(PDP_NG * 32, PDP_MINB)
pdp_k_sens_fwd_out(int B, int H, const double* __restrict__ x0, const double* __restrict__ theta, int theta_stride,
                   const double* __restrict__ inputs, const double* __restrict__ Xobs, double* __restrict__ X,
                   double* __restrict__ Uout, double* __restrict__ dX, double* __restrict__ dU, double* __restrict__ loss_dp,
                   int* __restrict__ status)
{
  pdp_sens_body<true>(B, H, x0, theta, theta_stride, inputs, Xobs, X, Uout, dX, dU, loss_dp, status);
}

extern "C" void pdpmod_info(int* out) {
  out[0] = PDP_KIND; out[1] = PDP_N; out[2] = PDP_M; out[3] = PDP_R; out[4] = PDP_NG; out[5] = PDP_GMAX;
  out[6] = 0; out[7] = 0; out[8] = PDP_RCF; out[9] = PDP_RCO; out[10] = 0;
}
"""

_LAUNCHER = r"""
extern "C" int pdpmod_sens_fwd(int B, int H, const double* x0, const double* theta, int theta_stride, const double* inputs,
                               const double* Xobs, double* X, double* Uout, double* dX, double* dU, double* loss_dp,
                               int* status, cudaStream_t st) {
  if (B <= 0) return 0;
  const bool out = X || Uout || dX || dU;
  const int rc = out ? PDP_RCO : PDP_RCF;
  size_t rows = (size_t)2 * rc * PDP_RIN;
  if (out) rows += (size_t)(rc + 1) * PDP_N * (1 + PDP_R) + (PDP_CP ? (size_t)rc * PDP_M * (1 + PDP_R) : 0);
  const size_t smem = ((size_t)PDP_STATE_DOUBLES + rows * PDP_TLD) * sizeof(double);
  cudaError_t e = pdp_opt_in_smem(out ? (const void*)pdp_k_sens_fwd_out : (const void*)pdp_k_sens_fwd, smem);
  if (e != cudaSuccess) return (int)e;
  const dim3 grid((B + 31) / 32);
  if (out)
    pdp_k_sens_fwd_out<<<grid, PDP_NG * 32, smem, st>>>(B, H, x0, theta, theta_stride, inputs, Xobs, X, Uout, dX, dU, loss_dp, status);
  else
    pdp_k_sens_fwd<<<grid, PDP_NG * 32, smem, st>>>(B, H, x0, theta, theta_stride, inputs, Xobs, loss_dp, status);
  return (int)cudaGetLastError();
}
"""


class SensModuleSource:
    def __init__(self, kind: int, state: SX, control: SX, auxvar: SX, dyn: SX,
                 policy: Optional[SX] = None, tvar: Optional[SX] = None,
                 path_cost: Optional[SX] = None, final_cost: Optional[SX] = None,
                 max_group_cols: int = 0, max_groups: int = 8, min_blocks: int = 1, tile_kb_per_warp: int = 12):
        self.min_blocks = max(1, int(min_blocks))
        self.kind = kind
        self.x, self.u, self.th = state, control, auxvar
        self.n, self.m, self.r = state.numel(), control.numel(), auxvar.numel()
        n, m, r = self.n, self.m, self.r
        self.dyn = SX(dyn).reshape((n, 1))
        self.t = tvar if tvar is not None else SX.sym("t")
        self.F = S.jacobian(self.dyn, self.x)
        self.G = S.jacobian(self.dyn, self.u)
        self.E = S.jacobian(self.dyn, self.th)
        if kind == KIND_CP:
            assert policy is not None and path_cost is not None and final_cost is not None
            self.policy = SX(policy).reshape((m, 1))
            self.Ux = S.jacobian(self.policy, self.x)
            self.Ue = S.jacobian(self.policy, self.th)
            self.c = SX(path_cost)
            self.h = SX(final_cost)
            self.cx = S.jacobian(self.c, self.x)
            self.cu = S.jacobian(self.c, self.u)
            self.hx = S.jacobian(self.h, self.x)
            if any(e is not S.ZERO for e in self.E.elements()):
                raise ValueError("ControlPlanning dynamics must not depend on the policy parameters")
        else:
            self.policy = None
        # column groups = warps of a block.  Default: one column per warp up to ``max_groups`` warps; an explicit
        # ``max_group_cols`` asks for ceil(r / max_group_cols) groups.
        if max_group_cols and max_group_cols > 0:
            ng = max(1, -(-r // int(max_group_cols)))
        else:
            ng = min(max(r, 1), int(max_groups))
        ng = max(1, min(ng, 16, max(r, 1)))
        per = -(-max(r, 1) // ng)
        self.groups: List[List[int]] = [list(range(g * per, min(r, (g + 1) * per))) for g in range(ng)]
        self.groups = [g for g in self.groups if g] or [[]]
        self.gmax = max(1, max(len(g) for g in self.groups))
        cp = kind == KIND_CP
        # rows per time step of the staging tiles, and the chunk lengths that keep them within the budget
        self.rin = 0 if cp else (m + n)
        self.rout = n + n * r + ((m + m * r) if cp else 0)
        fin = n + n * r                                         # final rows x_H, dX_H ride with the last chunk
        # staging tiles of a block: ~tile_kb_per_warp KB per warp of the block, at least 18 KB (the state columns come on top)
        rows_budget = max(18 * 1024, int(tile_kb_per_warp) * 1024 * ng) // (33 * 8)
        self.rcf = max(1, min(8, rows_budget // (2 * self.rin))) if self.rin else 8
        self.rco = max(1, min(8, (rows_budget - fin) // max(2 * self.rin + self.rout, 1)))

    # ------------------------------------------------------------------------------------------
    def _th_in_regs(self):
        return self.r <= 16

    def _leaf(self) -> Dict[int, str]:
        leaf = {}
        for k, e in enumerate(self.x.elements()):
            leaf[e.uid] = "xs%d" % k
        for k, e in enumerate(self.u.elements()):
            leaf[e.uid] = "u%d" % k
        for k, e in enumerate(self.th.elements()):
            leaf[e.uid] = ("th%d" % k) if self._th_in_regs() else ("th[%d]" % k)
        leaf[self.t.elements()[0].uid] = "tt"
        return leaf

    def _ent(self, node: Node, names: Dict[int, str]):
        if node is S.ZERO:
            return None
        if node.op == "const":
            return _lit(node.val)
        return names[node.uid]

    def _group_step(self, g: int, cols: Sequence[int]) -> str:
        """Straight-line code of one time step (chunk slot ``s``, time ``t``) for column group ``g``."""
        n, m, r = self.n, self.m, self.r
        cp = self.kind == KIND_CP
        L: List[str] = []
        ind = "          "
        leaf = self._leaf()
        outs: List[Node] = []
        if cp:
            # the policy must be evaluated first (u is an input of everything else)
            plines, pnames = S.emit_c(self.policy.elements(), leaf, prefix="p", indent=ind)
            L += plines
            for a in range(m):
                L.append(ind + "const double u%d = %s;" % (a, pnames[a]))
        else:
            for a in range(m):
                L.append(ind + "const double u%d = INT[(PDP_IU + s * PDP_M + %d) * PDP_TLD];" % (a, a))
        outs += self.dyn.elements()
        for M in (self.F, self.G):
            outs += [e for e in M.elements() if e is not S.ZERO and e.op != "const"]
        if cp:
            outs += [e for e in self.Ux.elements() if e is not S.ZERO and e.op != "const"]
            outs += [self.Ue.at(a, c) for a in range(m) for c in cols if self.Ue.at(a, c).op not in ("const",)]
            outs += [e for e in self.cx.elements() + self.cu.elements() if e.op != "const"]
            outs += self.c.elements()
        else:
            outs += [self.E.at(i, c) for i in range(n) for c in cols if self.E.at(i, c).op != "const"]
        lines, names_list = S.emit_c(outs, leaf, prefix="w", indent=ind)
        L += lines
        names = {o.uid: nm for o, nm in zip(outs, names_list)}
        dyn_names = [names[e.uid] for e in self.dyn.elements()]
        # ---- weights of the chain rule at step t
        if cp:
            if g == 0:
                L.append(ind + "loss += %s;" % names[self.c.elements()[0].uid])
            wx = [self._ent(self.cx.at(0, i), names) for i in range(n)]
            wu = [self._ent(self.cu.at(0, a), names) for a in range(m)]
        else:
            L.append(ind + "double " + ", ".join("d%d = 0.0" % i for i in range(n)) + ";")
            L.append(ind + "if (Xobs) {")
            for i in range(n):
                L.append(ind + "  d%d = xs%d - INT[(PDP_IX + s * PDP_N + %d) * PDP_TLD];" % (i, i, i))
            if g == 0:
                L.append(ind + "  " + " ".join("loss = fma(d%d, d%d, loss);" % (i, i) for i in range(n)))
            L.append(ind + "}")
            wx = ["d%d" % i for i in range(n)]
            wu = [None] * m
        # ---- per column
        for k, c in enumerate(cols):
            L.append(ind + "{ // column %d" % c)
            for i in range(n):
                L.append(ind + "  const double s%d = ST[%d * 32];" % (i, k * n + i))
            acc = "g%d" % k
            for i in range(n):
                if wx[i] is not None:
                    L.append(ind + "  %s = fma(%s, s%d, %s);" % (acc, wx[i], i, acc))
            du = [None] * m
            if cp:
                for a in range(m):
                    terms = []
                    for l in range(n):
                        e = self._ent(self.Ux.at(a, l), names)
                        if e is not None:
                            terms.append("%s * s%d" % (e, l))
                    e = self._ent(self.Ue.at(a, c), names)
                    if e is not None:
                        terms.append(e)
                    if terms:
                        L.append(ind + "  const double v%d = %s;" % (a, " + ".join(terms)))
                        du[a] = "v%d" % a
                        if wu[a] is not None:
                            L.append(ind + "  %s = fma(%s, v%d, %s);" % (acc, wu[a], a, acc))
            L.append(ind + "  if (OUT && dX) {")
            for i in range(n):
                L.append(ind + "    OT[(PDP_ODX + s * %d + %d) * PDP_TLD] = s%d;" % (n * r, i * r + c, i))
            L.append(ind + "  }")
            if cp:
                L.append(ind + "  if (OUT && dU) {")
                for a in range(m):
                    L.append(ind + "    OT[(PDP_ODU + s * %d + %d) * PDP_TLD] = %s;" % (m * r, a * r + c, du[a] if du[a] else "0.0"))
                L.append(ind + "  }")
            for i in range(n):
                terms = []
                for l in range(n):
                    e = self._ent(self.F.at(i, l), names)
                    if e is not None:
                        terms.append("s%d" % l if e == "1.0" else "%s * s%d" % (e, l))
                if cp:
                    for a in range(m):
                        e = self._ent(self.G.at(i, a), names)
                        if e is not None and du[a] is not None:
                            terms.append("%s * %s" % (e, du[a]))
                else:
                    e = self._ent(self.E.at(i, c), names)
                    if e is not None:
                        terms.append(e)
                L.append(ind + "  ST[%d * 32] = %s;" % (k * n + i, " + ".join(terms) if terms else "0.0"))
            L.append(ind + "}")
        # ---- outputs (group 0) and state advance
        if g == 0:
            L.append(ind + "if (OUT && X) {")
            for i in range(n):
                L.append(ind + "  OT[(PDP_OX + s * PDP_N + %d) * PDP_TLD] = xs%d;" % (i, i))
            L.append(ind + "}")
            if cp:
                L.append(ind + "if (OUT && Uout) {")
                for a in range(m):
                    L.append(ind + "  OT[(PDP_OU + s * PDP_M + %d) * PDP_TLD] = u%d;" % (a, a))
                L.append(ind + "}")
        for i in range(n):
            L.append(ind + "const double xn%d = %s;" % (i, dyn_names[i]))
        for i in range(n):
            L.append(ind + "xs%d = xn%d;" % (i, i))
        return "\n".join(L)

    def _group_terminal(self, g: int, cols: Sequence[int]) -> str:
        """Terminal step t = H (slot ``s = nst`` of the last chunk: the final rows ride with it)."""
        n, r = self.n, self.r
        cp = self.kind == KIND_CP
        L: List[str] = []
        ind = "          "
        leaf = self._leaf()
        if cp:
            outs = self.h.elements() + [e for e in self.hx.elements() if e.op != "const"]
            lines, nl = S.emit_c(outs, leaf, prefix="h", indent=ind)
            L += lines
            names = {o.uid: nm for o, nm in zip(outs, nl)}
            if g == 0:
                L.append(ind + "loss += %s;" % names[self.h.elements()[0].uid])
            wx = [self._ent(self.hx.at(0, i), names) for i in range(n)]
        else:
            L.append(ind + "double " + ", ".join("d%d = 0.0" % i for i in range(n)) + ";")
            L.append(ind + "if (Xobs) {")
            for i in range(n):
                L.append(ind + "  d%d = xs%d - Xobs[((size_t)b * (H + 1) + H) * %d + %d];" % (i, i, n, i))
            if g == 0:
                L.append(ind + "  " + " ".join("loss = fma(d%d, d%d, loss);" % (i, i) for i in range(n)))
            L.append(ind + "}")
            wx = ["d%d" % i for i in range(n)]
        for k, c in enumerate(cols):
            L.append(ind + "{")
            for i in range(n):
                L.append(ind + "  const double s%d = ST[%d * 32];" % (i, k * n + i))
                if wx[i] is not None:
                    L.append(ind + "  g%d = fma(%s, s%d, g%d);" % (k, wx[i], i, k))
            L.append(ind + "  if (OUT && dX) {")
            for i in range(n):
                L.append(ind + "    OT[(PDP_ODX + s * %d + %d) * PDP_TLD] = s%d;" % (n * r, i * r + c, i))
            L.append(ind + "  }")
            L.append(ind + "}")
        if g == 0:
            L.append(ind + "if (OUT && X) {")
            for i in range(n):
                L.append(ind + "  OT[(PDP_OX + s * PDP_N + %d) * PDP_TLD] = xs%d;" % (i, i))
            L.append(ind + "}")
        return "\n".join(L)

    def source(self) -> str:
        from .kernel_templates import K_OPT_IN, K_STAGING
        n, m, r = self.n, self.m, self.r
        cp = self.kind == KIND_CP
        ng = len(self.groups)
        hdr = ["// GENERATED by pontryagin_differentiable_programming_b200/codegen_sens.py -- do not edit",
               "#include <cuda_runtime.h>", "#include <math.h>", "#include <stdint.h>",
               "#define PDP_N %d" % n, "#define PDP_M %d" % m, "#define PDP_R %d" % r, "#define PDP_KIND %d" % self.kind,
               "#define PDP_NG %d" % ng, "#define PDP_GMAX %d" % self.gmax, "#define PDP_RCF %d" % self.rcf,
               "#define PDP_RCO %d" % self.rco, "#define PDP_RIN %d" % self.rin, "#define PDP_CP %d" % (1 if cp else 0),
               "#define PDP_STATE_DOUBLES (PDP_NG * PDP_GMAX * PDP_N * 32)", "#define PDP_MINB %d" % self.min_blocks]
        body = [K_STAGING, _BODY_HEAD]
        if self._th_in_regs() and r > 0:
            body.append("  " + " ".join("const double th%d = th[%d];" % (k, k) for k in range(r)))
            body.append("  " + " ".join("(void)th%d;" % k for k in range(r)))
        body.append("  double " + ", ".join("xs%d = x0[(size_t)b * %d + %d]" % (i, n, i) for i in range(n)) + ";")
        body.append("  double " + ", ".join("g%d = 0.0" % k for k in range(self.gmax)) + ";")
        body.append("  " + " ".join("(void)g%d;" % k for k in range(self.gmax)))
        if not cp:
            body.append(_FIRST_CHUNK)
        body.append(_LOOP_HEAD)
        if not cp:
            body.append(_NEXT_CHUNK)
        body.append("    switch (grp) {")
        for g, cols in enumerate(self.groups):
            body.append("    case %d: {" % g)
            body.append("      #pragma unroll 1")
            body.append("      for (int s = 0; s < nst; ++s) {")
            body.append("          const int t = t0 + s; const double tt = (double)t; (void)tt; (void)t;")
            body.append(self._group_step(g, cols))
            body.append("      }")
            body.append("      if (last) {")
            body.append("          const int s = nst; (void)s;")
            body.append(self._group_terminal(g, cols))
            body.append("      }")
            body.append("    } break;")
        body.append("    default: break;")
        body.append("    }")
        body.append(_WRITE_OUT)
        body.append("  if (loss_dp && live) {")
        body.append("    switch (grp) {")
        for g, cols in enumerate(self.groups):
            st = ["      loss_dp[(size_t)b * %d + %d] = g%d;" % (r + 1, 1 + c, k) for k, c in enumerate(cols)]
            if g == 0:
                st.append("      loss_dp[(size_t)b * %d] = loss;" % (r + 1))
            body.append("    case %d:\n%s\n      break;" % (g, "\n".join(st)))
        body.append("    default: break;")
        body.append("    }")
        body.append("  }")
        body.append("  if (status && live && grp == 0 && !isfinite(loss)) atomicOr(&status[b], 1);")
        body.append("}")
        body.append(_ENTRY_POINTS)
        body.append(K_OPT_IN)
        body.append(_LAUNCHER)
        return "\n".join(hdr) + "\n" + "\n".join(body)

    def key(self) -> str:
        return hashlib.sha256(self.source().encode()).hexdigest()[:20]
