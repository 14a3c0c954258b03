"""Forward-sensitivity modules: ``SysID.step`` and ``ControlPlanning.step`` as fused CUDA kernels.

Reference semantics restated:

* SysID (``PDP/PDP.py:1209-1296``): rollout with given inputs, ``X_{t+1} = F_t X_t + E_t`` with
  ``F = df/dx, E = df/dtheta``, ``X_0 = 0``; ``loss = ||x - xobs||^2``, ``dp = sum_t (x_t - xobs_t) X_t``
  (half the gradient, like the reference).
* ControlPlanning (``PDP/PDP.py:763-878``): rollout under the policy ``u_t = pi(t, x_t, theta)``,
  ``U_t = Ux_t X_t + Ue_t``, ``X_{t+1} = F_t X_t + G_t U_t``; ``loss = sum c + h``,
  ``dtheta = sum_t (c_x X_t + c_u U_t) + h_x X_H``.

Kernel design (DESIGN.md "pdp_k_sens_fwd"): ONE THREAD PER (trajectory, column group).  The columns
of the sensitivity matrix ``X_t (n x r)`` evolve independently, so each thread keeps its group's
columns in shared memory laid out ``[element][thread]`` (conflict-free) and streams them through
registers one column at a time; everything else (state, Jacobian entries, policy derivatives) is
straight-line generated code over structural non-zeros.  Large ``r`` (neural policies, r = 45)
is split into column groups, which also restores parallelism when B is small.  The group index is the FAST grid
dimension (``blockIdx.x``), so the groups of one block of trajectories are scheduled next to each other and the
second group's reads of the shared input rows hit L2 instead of HBM (measured round 2: the round-1 order
``(trajectory block, group)`` read everything once per group from DRAM, 1.86x the algorithmic bytes at C5; C5
0.321 -> 0.300 ms).  Measured and rejected in round 2 (profiles/r2d_*, r2e_*, r2f_*): fetching the next step's rows into
registers (spills: 0.381 ms), cache-hint prefetches (0.415 ms), rows as 16-byte pairs (0.314 ms), register caps for more
resident warps (168 / 128 registers: 0.70 / 1.16 ms -- the spills cost far more than the occupancy gives) and a
warp-cooperative staged variant with shared tiles (0.47 ms).
"""
from __future__ import annotations

import hashlib
from typing import Dict, List, Optional, Sequence

from . import symbolic as S
from .symbolic import SX, Node
from .codegen import _lit

KIND_SYSID, KIND_CP = 2, 3


class SensModuleSource:
    def __init__(self, kind: int, state: SX, control: SX, auxvar: SX, dyn: SX,
                 policy: Optional[SX] = None, tvar: Optional[SX] = None,
                 path_cost: Optional[SX] = None, final_cost: Optional[SX] = None,
                 max_group_cols: int = 12, block: int = 64):
        self.kind = kind
        self.x, self.u, self.th = state, control, auxvar
        self.n, self.m, self.r = state.numel(), control.numel(), auxvar.numel()
        n, m, r = self.n, self.m, self.r
        self.dyn = SX(dyn).reshape((n, 1))
        self.block = block
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
        # column groups
        ng = max(1, -(-r // max_group_cols))
        per = -(-r // ng)
        self.groups: List[List[int]] = [list(range(g * per, min(r, (g + 1) * per))) for g in range(ng)]
        self.groups = [g for g in self.groups if g]
        self.gmax = max(len(g) for g in self.groups)

    # ------------------------------------------------------------------------------------------
    def _param_recips(self):
        """Divisors that depend on theta only: reciprocals computed once per thread before the time loop (see
        ``symbolic.param_divisors``)."""
        if getattr(self, "_recips", None) is None:
            outs = list(self.dyn.elements())
            for M in (self.F, self.G, self.E):
                outs += M.elements()
            if self.kind == KIND_CP:
                outs += self.policy.elements() + self.Ux.elements() + self.Ue.elements() + self.cx.elements() + \
                    self.cu.elements() + self.hx.elements() + self.c.elements() + self.h.elements()
            self._recips = S.param_divisors(outs, {e.uid for e in self.th.elements()})
        return self._recips

    def _rmap(self):
        return {d.uid: "rcp%d" % k for k, d in enumerate(self._param_recips())}

    def _leaf(self) -> Dict[int, str]:
        leaf = {}
        for k, e in enumerate(self.x.elements()):
            leaf[e.uid] = "xs%d" % k
        for k, e in enumerate(self.u.elements()):
            leaf[e.uid] = "u%d" % k
        for k, e in enumerate(self.th.elements()):
            leaf[e.uid] = "th[%d]" % k
        leaf[self.t.elements()[0].uid] = "tt"
        return leaf

    def _ent(self, node: Node, names: Dict[int, str]):
        if node is S.ZERO:
            return None
        if node.op == "const":
            return _lit(node.val)
        return names[node.uid]

    def _group_step(self, g: int, cols: Sequence[int]) -> str:
        """Straight-line code of one time step for column group ``g``."""
        n, m, r = self.n, self.m, self.r
        cp = self.kind == KIND_CP
        L: List[str] = []
        ind = "        "
        leaf = self._leaf()
        # ---- expressions needed this step
        outs: List[Node] = []
        if cp:
            # the policy must be evaluated first (u is an input of everything else)
            plines, pnames = S.emit_c(self.policy.elements(), leaf, prefix="p", indent=ind, recip=self._rmap())
            L += plines
            for a in range(m):
                L.append(ind + "const double u%d = %s;" % (a, pnames[a]))
        outs += self.dyn.elements()
        mats = {"F": self.F, "G": self.G}
        for M in mats.values():
            outs += [e for e in M.elements() if e is not S.ZERO and e.op != "const"]
        if cp:
            outs += [e for e in self.Ux.elements() if e is not S.ZERO and e.op != "const"]
            outs += [self.Ue.at(a, c) for a in range(m) for c in cols if self.Ue.at(a, c).op not in ("const",)]
            outs += [e for e in self.cx.elements() + self.cu.elements() if e.op != "const"]
            outs += self.c.elements()
        else:
            outs += [self.E.at(i, c) for i in range(n) for c in cols if self.E.at(i, c).op != "const"]
        lines, names_list = S.emit_c(outs, leaf, prefix="w", indent=ind, recip=self._rmap())
        L += lines
        names = {o.uid: nm for o, nm in zip(outs, names_list)}
        dyn_names = [names[e.uid] for e in self.dyn.elements()]
        # ---- weights of the chain rule at step t
        if cp:
            L.append(ind + "if (grp == 0) loss += %s;" % names[self.c.elements()[0].uid])
            wx = [self._ent(self.cx.at(0, i), names) for i in range(n)]
            wu = [self._ent(self.cu.at(0, a), names) for a in range(m)]
        else:
            L.append(ind + "double " + ", ".join("d%d = 0.0" % i for i in range(n)) + ";")
            L.append(ind + "if (Xobs) {")
            for i in range(n):
                L.append(ind + "  d%d = xs%d - xo%d;" % (i, i, i))
            L.append(ind + "  if (grp == 0) { " + " ".join("loss = fma(d%d, d%d, loss);" % (i, i) for i in range(n)) + " }")
            L.append(ind + "}")
            wx = ["d%d" % i for i in range(n)]
            wu = [None] * m
        # ---- per column
        for k, c in enumerate(cols):
            L.append(ind + "{ // column %d" % c)
            for i in range(n):
                L.append(ind + "  const double s%d = SM[%d * PDP_BLOCK];" % (i, k * n + i))
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
            L.append(ind + "  if (dX) { double* o = dX + (((size_t)b * (H + 1) + t) * %d) * %d + %d;" % (n, r, c))
            for i in range(n):
                L.append(ind + "    o[%d] = s%d;" % (i * r, i))
            L.append(ind + "  }")
            if cp:
                L.append(ind + "  if (dU) { double* o = dU + (((size_t)b * H + t) * %d) * %d + %d;" % (m, r, c))
                for a in range(m):
                    L.append(ind + "    o[%d] = %s;" % (a * r, du[a] if du[a] else "0.0"))
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
                L.append(ind + "  SM[%d * PDP_BLOCK] = %s;" % (k * n + i, " + ".join(terms) if terms else "0.0"))
            L.append(ind + "}")
        # ---- outputs and state advance
        L.append(ind + "if (X && grp == 0) { double* o = X + ((size_t)b * (H + 1) + t) * %d;" % n)
        for i in range(n):
            L.append(ind + "  o[%d] = xs%d;" % (i, i))
        L.append(ind + "}")
        if cp:
            L.append(ind + "if (Uout && grp == 0) { double* o = Uout + ((size_t)b * H + t) * %d;" % m)
            for a in range(m):
                L.append(ind + "  o[%d] = u%d;" % (a, a))
            L.append(ind + "}")
        for i in range(n):
            L.append(ind + "const double xn%d = %s;" % (i, dyn_names[i]))
        for i in range(n):
            L.append(ind + "xs%d = xn%d;" % (i, i))
        return "\n".join(L)

    def _group_terminal(self, g: int, cols: Sequence[int]) -> str:
        n, r = self.n, self.r
        cp = self.kind == KIND_CP
        L: List[str] = []
        ind = "      "
        leaf = self._leaf()
        if cp:
            outs = self.h.elements() + [e for e in self.hx.elements() if e.op != "const"]
            lines, nl = S.emit_c(outs, leaf, prefix="h", indent=ind, recip=self._rmap())
            L += lines
            names = {o.uid: nm for o, nm in zip(outs, nl)}
            L.append(ind + "if (grp == 0) loss += %s;" % names[self.h.elements()[0].uid])
            wx = [self._ent(self.hx.at(0, i), names) for i in range(n)]
        else:
            L.append(ind + "double " + ", ".join("d%d = 0.0" % i for i in range(n)) + ";")
            L.append(ind + "if (Xobs) {")
            for i in range(n):
                L.append(ind + "  d%d = xs%d - nxo%d;" % (i, i, i))
            L.append(ind + "  if (grp == 0) { " + " ".join("loss = fma(d%d, d%d, loss);" % (i, i) for i in range(n)) + " }")
            L.append(ind + "}")
            wx = ["d%d" % i for i in range(n)]
        for k, c in enumerate(cols):
            L.append(ind + "{")
            for i in range(n):
                L.append(ind + "  const double s%d = SM[%d * PDP_BLOCK];" % (i, k * n + i))
                if wx[i] is not None:
                    L.append(ind + "  g%d = fma(%s, s%d, g%d);" % (k, wx[i], i, k))
            L.append(ind + "  if (dX) { double* o = dX + (((size_t)b * (H + 1) + H) * %d) * %d + %d;" % (n, r, c))
            for i in range(n):
                L.append(ind + "    o[%d] = s%d;" % (i * r, i))
            L.append(ind + "  }")
            L.append(ind + "}")
        L.append(ind + "if (X && grp == 0) { double* o = X + ((size_t)b * (H + 1) + H) * %d;" % n)
        for i in range(n):
            L.append(ind + "  o[%d] = xs%d;" % (i, i))
        L.append(ind + "}")
        for k, c in enumerate(cols):
            L.append(ind + "if (loss_dp) loss_dp[(size_t)b * %d + %d] = g%d;" % (r + 1, 1 + c, k))
        L.append(ind + "if (loss_dp && grp == 0) loss_dp[(size_t)b * %d] = loss;" % (r + 1))
        return "\n".join(L)

    def source(self) -> str:
        n, m, r = self.n, self.m, self.r
        cp = self.kind == KIND_CP
        hdr = ["// GENERATED by pontryagin_differentiable_programming_b200/codegen_sens.py -- do not edit",
               "#include <cuda_runtime.h>", "#include <math.h>", "#include <stdint.h>",
               "#define PDP_N %d" % n, "#define PDP_M %d" % m, "#define PDP_R %d" % r, "#define PDP_KIND %d" % self.kind,
               "#define PDP_NG %d" % len(self.groups), "#define PDP_GMAX %d" % self.gmax, "#define PDP_BLOCK %d" % self.block]
        body = []
        body.append(r'''
// One thread per (trajectory, column group); the group's columns of X_t (n x r) live in shared memory.
extern "C" __global__ void __launch_bounds__(PDP_BLOCK)
pdp_k_sens_fwd(int B, int H, const double* __restrict__ x0, const double* __restrict__ theta, int theta_stride,
               const double* __restrict__ inputs, const double* __restrict__ Xobs, double* __restrict__ X,
               double* __restrict__ Uout, double* __restrict__ dX, double* __restrict__ dU, double* __restrict__ loss_dp,
               int* __restrict__ status)
{
  extern __shared__ __align__(16) double pdp_smem[];
  const int b = blockIdx.y * PDP_BLOCK + threadIdx.x;
  const int grp = blockIdx.x;
  if (b >= B) return;
  double* SM = pdp_smem + threadIdx.x;
  const double* th = theta + (size_t)b * theta_stride;
  for (int k = 0; k < PDP_GMAX * PDP_N; ++k) SM[k * PDP_BLOCK] = 0.0;
  double loss = 0.0;''')
        rc = self._param_recips()
        if rc:
            rl, rn = S.emit_c([S.div(S.ONE, d) for d in rc], self._leaf(), prefix="q", indent="  ")
            body += rl
            body.append("  " + " ".join("const double rcp%d = %s;" % (k, nm) for k, nm in enumerate(rn)))
        body.append("  double " + ", ".join("xs%d = x0[(size_t)b * %d + %d]" % (i, n, i) for i in range(n)) + ";")
        body.append("  double " + ", ".join("g%d = 0.0" % k for k in range(self.gmax)) + ";")
        if not cp:
            body.append("  " + " ".join("double nxo%d = 0.0;" % i for i in range(n)))
        body.append("  switch (grp) {")
        for g, cols in enumerate(self.groups):
            body.append("  case %d: {" % g)
            body.append("    #pragma unroll 1")
            body.append("    for (int t = 0; t < H; ++t) {")
            body.append("      const double tt = (double)t; (void)tt;")
            if not cp:
                body.append("      " + " ".join("const double u%d = inputs[((size_t)b * H + t) * %d + %d];" % (a, m, a) for a in range(m)))
                body.append("      " + " ".join("const double xo%d = Xobs ? Xobs[((size_t)b * (H + 1) + t) * %d + %d] : 0.0;" % (i, n, i) for i in range(n)))
            body.append(self._group_step(g, cols))
            body.append("    }")
            body.append("    {")
            if not cp:
                body.append("      if (Xobs) { " + " ".join("nxo%d = Xobs[((size_t)b * (H + 1) + H) * %d + %d];" % (i, n, i) for i in range(n)) + " }")
            body.append(self._group_terminal(g, cols))
            body.append("    }")
            body.append("  } break;")
        body.append("  default: break;")
        body.append("  }")
        body.append("  if (status && grp == 0 && !isfinite(loss)) atomicOr(&status[b], 1);")
        body.append("}")
        from .kernel_templates import K_OPT_IN
        body.append(r'''
extern "C" void pdpmod_info(int* out) {
  out[0] = PDP_KIND; out[1] = PDP_N; out[2] = PDP_M; out[3] = PDP_R; out[4] = PDP_NG; out[5] = PDP_GMAX;
  out[6] = 0; out[7] = 0; out[8] = 0; out[9] = 0; out[10] = 0;
}
''' + K_OPT_IN + r'''

extern "C" int pdpmod_sens_fwd(int B, int H, const double* x0, const double* theta, int theta_stride, const double* inputs,
                               const double* Xobs, double* X, double* Uout, double* dX, double* dU, double* loss_dp,
                               int* status, cudaStream_t st) {
  if (B <= 0) return 0;
  const size_t smem = (size_t)PDP_GMAX * PDP_N * PDP_BLOCK * sizeof(double);
  cudaError_t e = pdp_opt_in_smem((const void*)pdp_k_sens_fwd, smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(PDP_NG, (B + PDP_BLOCK - 1) / PDP_BLOCK);
  pdp_k_sens_fwd<<<grid, PDP_BLOCK, smem, st>>>(B, H, x0, theta, theta_stride, inputs, Xobs, X, Uout, dX, dU, loss_dp, status);
  return (int)cudaGetLastError();
}
''')
        return "\n".join(hdr) + "\n" + "\n".join(body)

    def key(self) -> str:
        return hashlib.sha256(self.source().encode()).hexdigest()[:20]
