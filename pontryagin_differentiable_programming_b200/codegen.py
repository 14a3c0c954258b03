"""Expressions -> one sm_100a CUDA translation unit per (system, mode).

This is the "diffPMP" step of the engine (reference ``PDP/PDP.py:222-270`` builds 14 CasADi
``Function`` objects and ``getAuxSys`` :272-314 calls them per time step): here the same
derivatives are taken symbolically ONCE and emitted as straight-line CUDA that is compiled into
the kernels -- the auxiliary-system matrices never exist in HBM.

Three module kinds are generated (see DESIGN.md for the kernel designs and rooflines):

``oc``     OCSys path: ``pdp_k_rollout_costate`` (thread / trajectory), ``pdp_k_aux_lqr``
           (warp / trajectory: chunked aux evaluation with lanes = time steps, structured-sparse
           Riccati sweep in the stacked form, gain spill, aux forward pass), ``pdp_k_aux_eval``
           (thread / (trajectory, step): dense aux matrices for the legacy ``getAuxSys`` API).
``sysid``  SysID path: rollout + forward sensitivity + fused loss / gradient.
``cp``     ControlPlanning path: policy rollout + forward sensitivity + fused chain rule, and the
           adjoint gradient that replaces the reference's recovery matrix.

The sparsity pattern of every derivative matrix is known at generation time, so all small-matrix
products are emitted fully unrolled over structural non-zeros only; constants are folded into the
instruction stream, shared expression nodes are stored once per step ("slots").
"""
from __future__ import annotations

import hashlib
from typing import Dict, List, Sequence, Tuple

from . import symbolic as S
from .symbolic import SX, Node

WARP = 32


# ---------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------

def _lit(v: float) -> str:
    return S._c_literal(float(v))


class SlotTable:
    """Distinct non-constant expression nodes that must be materialised per time step."""

    def __init__(self):
        self.nodes: List[Node] = []
        self.index: Dict[int, int] = {}

    def entry(self, node: Node):
        """Classify a matrix entry: ('z',) | ('c', value) | ('v', slot, sign)."""
        if node is S.ZERO:
            return ("z",)
        if node.op == "const":
            return ("c", node.val)
        sign = 1.0
        if node.op == "neg":
            sign, node = -1.0, node.args[0]
        k = self.index.get(node.uid)
        if k is None:
            k = len(self.nodes)
            self.index[node.uid] = k
            self.nodes.append(node)
        return ("v", k, sign)

    def exact(self, node: Node) -> int:
        """Slot holding exactly ``node`` (sign and constants included) -- used for per-lane indexed loads."""
        k = self.index.get(("x", node.uid))
        if k is None:
            if node.op not in ("neg", "const"):
                k = self.index.get(node.uid)          # already stored for the S block with sign +1
            if k is None:
                k = len(self.nodes)
                self.nodes.append(node)
            self.index[("x", node.uid)] = k
        return k

    def __len__(self):
        return len(self.nodes)


def _pad_ld(n):
    """Row stride (doubles) of a per-step slot row: >= n and = 2 (mod 16), so that the 8..16 evaluation
    lanes (one row each) hit distinct 16-byte bank groups when they store their slots."""
    v = max(int(n), 2)
    while v % 16 != 2:
        v += 1
    return v


def _emit_function(name: str, inputs: Sequence[Tuple[str, SX]], outputs: Sequence[Node], out_expr,
                   qualifiers="__device__ __forceinline__", recips=None) -> str:
    """``void name(const double* in0, ..., double* out)`` computing ``outputs``.

    ``out_expr(i)`` gives the C lvalue for output ``i``.  Inputs are read into locals first so the
    stores to ``out`` can never alias the loads.  ``recips = (param name, first index, [divisor nodes])``: the reciprocals
    of these parameter-only divisors sit behind the parameters in the same array (``param[first + k]``, filled once per
    trajectory by ``pdp_f_recips``) and divisions by them are emitted as multiplications."""
    leaf: Dict[int, str] = {}
    order = S.topo_order(outputs)
    used = {n.uid for n in order if n.op == "sym"}
    loads = []
    for pname, sx in inputs:
        for k, e in enumerate(sx.elements()):
            if e.uid in used:
                loads.append("  const double %s_%d = %s[%d];" % (pname, k, pname, k))
                leaf[e.uid] = "%s_%d" % (pname, k)
    rmap = None
    if recips and recips[2]:
        pname, first, nodes = recips
        divisors = {n.args[1].uid for n in order if n.op == "div"}
        rmap = {}
        for k, d in enumerate(nodes):
            if d.uid in divisors:
                loads.append("  const double rcp_%d = %s[%d];" % (k, pname, first + k))
                rmap[d.uid] = "rcp_%d" % k
    lines, names = S.emit_c(outputs, leaf, recip=rmap)
    sig = ", ".join("const double* __restrict__ %s" % p for p, _ in inputs)
    body = ["%s void %s(%s, double* __restrict__ out) {" % (qualifiers, name, sig)]
    body += loads + lines
    for i, nm in enumerate(names):
        body.append("  %s = %s;" % (out_expr(i), nm))
    body.append("}")
    return "\n".join(body)


class _Acc:
    """Tracks first-touch of accumulator registers so products start with a mul, not 0 + fma."""

    def __init__(self, prefix, n, lines, indent="      "):
        self.prefix, self.n, self.lines, self.indent = prefix, n, lines, indent
        self.touched = [False] * n

    def name(self, j):
        return "%s%d" % (self.prefix, j)

    def add(self, j, a: str, ent, load):
        """acc_j += a * entry, with entry classified by SlotTable.entry; ``load(slot)`` -> C name."""
        if ent[0] == "z":
            return
        acc = self.name(j)
        if ent[0] == "c":
            c = ent[1]
            if not self.touched[j]:
                self.lines.append("%s%s = %s;" % (self.indent, acc, a if c == 1.0 else ("-%s" % a if c == -1.0 else "%s * %s" % (a, _lit(c)))))
            elif c == 1.0:
                self.lines.append("%s%s += %s;" % (self.indent, acc, a))
            elif c == -1.0:
                self.lines.append("%s%s -= %s;" % (self.indent, acc, a))
            else:
                self.lines.append("%s%s = fma(%s, %s, %s);" % (self.indent, acc, a, _lit(c), acc))
        else:
            s = load(ent[1])
            neg = ent[2] < 0
            if not self.touched[j]:
                self.lines.append("%s%s = %s%s * %s;" % (self.indent, acc, "-" if neg else "", a, s))
            else:
                self.lines.append("%s%s = fma(%s%s, %s, %s);" % (self.indent, acc, "-" if neg else "", a, s, acc))
        self.touched[j] = True

    def finish(self):
        for j in range(self.n):
            if not self.touched[j]:
                self.lines.append("%s%s = 0.0;" % (self.indent, self.name(j)))


class _SlotLoader:
    """Emits broadcast shared-memory loads of aux slots, pairing neighbours into one 16-byte load."""

    def __init__(self, lines, base="ar", needed=(), indent="      ", tag="s"):
        self.lines, self.base, self.indent, self.tag = lines, base, indent, tag
        self.needed = set(needed)
        self.loaded: Dict[int, str] = {}

    def __call__(self, e: int) -> str:
        nm = self.loaded.get(e)
        if nm is not None:
            return nm
        pair = e ^ 1
        if pair in self.needed and pair not in self.loaded:
            lo = min(e, pair)
            v = "%s%d_%d" % (self.tag, lo, len(self.lines))
            self.lines.append("%sconst double2 %s = *reinterpret_cast<const double2*>(%s + %d);" % (self.indent, v, self.base, lo))
            self.loaded[lo] = v + ".x"
            self.loaded[lo + 1] = v + ".y"
        else:
            v = "%s%d_%d" % (self.tag, e, len(self.lines))
            self.lines.append("%sconst double %s = %s[%d];" % (self.indent, v, self.base, e))
            self.loaded[e] = v
        return self.loaded[e]


def _odd(n):
    return n if n % 2 == 1 else n + 1


def _even(n):
    return n if n % 2 == 0 else n + 1


# ---------------------------------------------------------------------------------------------
# OC module
# ---------------------------------------------------------------------------------------------

class OCModuleSource:
    """Generates the CUDA source of an ``oc`` module from the symbolic optimal-control system."""

    def __init__(self, state: SX, control: SX, auxvar: SX, dyn: SX, path_cost: SX, final_cost: SX,
                 chunk: int = 8, warps_per_block: int = 4, min_blocks: int = 1, fwd_warps_per_block: int = 4,
                 fwd_min_blocks: int = 1, keep_fg: bool = True, fast_rcp: bool = False, early_solve: bool = False,
                 fwd_pack: int = 0, fwd_chunk: int = 0, bwd_pack: int = 1, fwd_vec: int = -1, prefetch: int = 2,
                 prefetch_dist: int = 2, inline_eval: int = -1, h_group: int = 1, prefetch_l1_lead: int = 0, rollout_tma: int = 1, tma_chunk: int = 0):
        self.keep_fg = bool(keep_fg)
        self.rollout_tma, self.tma_chunk = int(rollout_tma), max(0, int(tma_chunk))
        self.prefetch_l1_lead = int(prefetch_l1_lead)
        self.h_group = int(h_group)
        self.inline_eval = int(inline_eval)
        self.prefetch, self.prefetch_dist = int(prefetch), int(prefetch_dist)
        self.fwd_vec = int(fwd_vec)
        self.fwd_pack, self.fwd_chunk = int(fwd_pack), int(fwd_chunk)
        self.bwd_pack = int(bwd_pack)
        self.fast_rcp, self.early_solve = bool(fast_rcp), bool(early_solve)
        self.min_blocks = int(min_blocks)
        self.wpbf, self.min_blocks_f = int(fwd_warps_per_block), int(fwd_min_blocks)
        self.x, self.u, self.th = state, control, auxvar
        self.n, self.m, self.r = state.numel(), control.numel(), auxvar.numel()
        self.nth = self.r
        self.ns = self.n + self.m + self.r
        self.chunk = int(chunk)
        self.wpb = int(warps_per_block)
        n, m, r = self.n, self.m, self.r
        self.dyn = SX(dyn).reshape((n, 1))
        self.c = SX(path_cost)
        self.h = SX(final_cost)
        self.lam = SX.sym("lam", n)
        # --- diffPMP (reference PDP.py:229-270)
        Hm = self.c + S.dot(self.dyn, self.lam)
        self.dfx = S.jacobian(self.dyn, self.x)
        self.dfu = S.jacobian(self.dyn, self.u)
        self.dfe = S.jacobian(self.dyn, self.th)
        self.dHx = S.jacobian(Hm, self.x).T
        self.dHu = S.jacobian(Hm, self.u).T
        self.ddHxx = S.jacobian(self.dHx, self.x)
        self.ddHxu = S.jacobian(self.dHx, self.u)
        self.ddHxe = S.jacobian(self.dHx, self.th)
        self.ddHux = S.jacobian(self.dHu, self.x)
        self.ddHuu = S.jacobian(self.dHu, self.u)
        self.ddHue = S.jacobian(self.dHu, self.th)
        self.dhx = S.jacobian(self.h, self.x).T
        self.ddhxx = S.jacobian(self.dhx, self.x)
        self.ddhxe = S.jacobian(self.dhx, self.th)
        self._customise()
        self.ns = self.n + self.m + self.r
        if self.bwd_pack == 2 and not (self.n <= 16 and self.m + self.r <= 16):
            self.bwd_pack = 1          # the two-rows-per-lane layout needs n <= 16 and m + r <= 16
        if self.bwd_pack == 2:
            self.chunk = min(self.chunk, 16)
        elif self.ns > WARP:
            # one trajectory per warp: lane j owns stack row j of [P ; . ; W^T] -- there is no lane for a row beyond 31
            raise ValueError(
                "the fused aux-LQR kernels hold the stack [P; control rows; W^T] with one row per lane: n + m + r = %d + %d + %d "
                "= %d exceeds the 32 rows of a warp.  Differentiate with respect to at most %d auxiliary variables per system "
                "(split the auxvar vector and build one system per group; the columns of dX/dtheta are independent), or use "
                "the generic LQR.lqrSolver, which splits the columns itself" % (self.n, self.m, self.r, self.ns, WARP - self.n - self.m))
        self._layout()

    def _customise(self):
        """Hook for variants that replace some auxiliary matrices (see NewtonModuleSource)."""

    # ---- slot layout ---------------------------------------------------------------------------
    def _layout(self):
        n, m, r, ns = self.n, self.m, self.r, self.ns
        slots = SlotTable()
        # S = [F | G | E]  (n x ns), row-major walk so slot order follows the use order
        self.S_ent = [[None] * ns for _ in range(n)]
        for k in range(n):
            for j in range(ns):
                if j < n:
                    node = self.dfx.at(k, j)
                elif j < n + m:
                    node = self.dfu.at(k, j - n)
                else:
                    node = self.dfe.at(k, j - n - m)
                self.S_ent[k][j] = slots.entry(node)
        self.nvar_s = len(slots)
        # stacked Hamiltonian Hessian  [[Hxx Hxu],[Hxu^T Huu],[Hxe^T Hue^T]]   (ns x (n+m))
        # (the reference never uses Hux in arithmetic, only transpose(Hxu): PDP.py:569,572,598)
        nm = n + m
        self.H_idx = [[0] * nm for _ in range(ns)]
        for j in range(ns):
            for l in range(nm):
                if j < n and l < n:
                    node = self.ddHxx.at(j, l)
                elif j < n:
                    node = self.ddHxu.at(j, l - n)
                elif j < nm and l < n:
                    node = self.ddHxu.at(l, j - n)
                elif j < nm:
                    node = self.ddHuu.at(j - n, l - n)
                elif l < n:
                    node = self.ddHxe.at(l, j - nm)
                else:
                    node = self.ddHue.at(l - n, j - nm)
                self.H_idx[j][l] = slots.exact(node)
        self.zero_slot = slots.exact(S.ZERO)
        self.h_groups = self._h_column_groups() if getattr(self, "bwd_pack", 1) == 2 else None
        self._place_h_slots(slots)
        self.slots = slots
        self.nvar = len(slots)
        self.auxld = _pad_ld(self.nvar)
        self.ldz = _odd(n)
        self.ldk = _even(n)

    def _h_slot_rows(self):
        """Stack rows served by slot 0 / slot 1 of the two-trajectory kernel."""
        return [list(range(self.n)), list(range(self.n, self.ns))]

    def _h_column_groups(self):
        if not getattr(self, "h_group", 1):       # A/B switch: one load per column, as in the one-trajectory kernel
            return [[[l] for l in range(self.n + self.m)] for _ in self._h_slot_rows()]
        return self._h_column_groups_coloured()

    def _h_column_groups_coloured(self):
        """Two-trajectory kernel: the Hamiltonian stack is mostly structural zeros (quadrotor: 127 of 442 entries), and
        every per-lane indexed load costs two shared-memory wavefronts however many lanes fetch a zero.  Columns of a
        slot in which no row has more than one non-zero share ONE load (each lane fetches its own entry, then selects
        the column it belongs to): greedy column colouring as for sparse Jacobians.  -> per slot a list of column
        groups; columns without any non-zero in the slot are in no group."""
        nm = self.n + self.m
        out = []
        for rows in self._h_slot_rows():
            nzr = {l: {j for j in rows if self.H_idx[j][l] != self.zero_slot} for l in range(nm)}
            groups: List[List[int]] = []
            for l in sorted(range(nm), key=lambda c: (-len(nzr[c]), c)):
                if not nzr[l]:
                    continue
                for g in groups:
                    if len(g) < 15 and all(not (nzr[l] & nzr[o]) for o in g):
                        g.append(l)
                        break
                else:
                    groups.append([l])
            out.append([sorted(g) for g in groups])
        return out

    def _h_group_entry(self, row, group):
        """(slot index, member position) of ``row``'s entry within a column group; (zero slot, 15) if it has none."""
        for pos, l in enumerate(group):
            if self.H_idx[row][l] != self.zero_slot:
                return self.H_idx[row][l], pos
        return self.zero_slot, 15

    def _h_conflict_cost(self, phys):
        """Extra shared-memory wavefronts of the per-lane indexed Hamiltonian loads: a 64-bit warp load is served
        per half-warp, one wavefront per distinct word that shares a 16-way bank-pair with another distinct word."""
        ns, nm = self.ns, self.n + self.m
        if getattr(self, "h_groups", None) is not None:
            # two-trajectory kernel: one load per (slot, column group) serves the slot's rows of a half-warp; idle team
            # lanes repeat the slot's first row
            cost = 0
            for rows, cgroups in zip(self._h_slot_rows(), self.h_groups):
                for g in cgroups:
                    banks: Dict[int, set] = {}
                    for j in rows:
                        a = phys[self._h_group_entry(j, g)[0]]
                        banks.setdefault(a % 16, set()).add(a)
                    cost += max(len(v) for v in banks.values()) - 1
            return cost
        if getattr(self, "bwd_pack", 1) == 2:
            groups = [list(range(self.n)), list(range(self.n, ns))]
        else:
            groups = [list(range(16)), list(range(16, 32))]
        cost = 0
        for l in range(nm):
            for rows in groups:
                banks: Dict[int, set] = {}
                for j in rows:
                    a = phys[self.H_idx[j][l]] if j < ns else phys[self.zero_slot]
                    banks.setdefault(a % 16, set()).add(a)
                cost += max(len(v) for v in banks.values()) - 1
        return cost

    def _place_h_slots(self, slots: "SlotTable", pad: int = 6, sweeps: int = 40):
        """Renumber the slots used only by the Hamiltonian stack (the [F|G|E] slots keep their order: the forward
        kernel evaluates just that prefix) so that the lanes of one indexed load hit distinct bank-pairs.
        Deterministic pairwise-swap descent; a few padding slots give it room."""
        lo = self.nvar_s
        movable = list(range(lo, len(slots.nodes)))
        if len(movable) < 2:
            return
        positions = list(range(lo, len(slots.nodes) + pad))
        phys = {e: e for e in range(len(slots.nodes))}          # slot id -> physical index
        occupant = {p: (p if p < len(slots.nodes) else None) for p in positions}
        best = self._h_conflict_cost(phys)
        for _ in range(sweeps):
            improved = False
            for e in movable:
                if best == 0:
                    break
                for p in positions:
                    q = phys[e]
                    if p == q:
                        continue
                    other = occupant[p]
                    phys[e] = p
                    if other is not None:
                        phys[other] = q
                    c = self._h_conflict_cost(phys)
                    if c < best:
                        best, improved = c, True
                        occupant[p], occupant[q] = e, other
                    else:
                        phys[e] = q
                        if other is not None:
                            phys[other] = p
            if not improved or best == 0:
                break
        # apply: rebuild the node list in physical order (padding slots hold 0.0) and remap the index tables
        size = max(phys.values()) + 1
        nodes = [S.ZERO] * size
        for e, p in phys.items():
            nodes[p] = slots.nodes[e]
        slots.nodes[:] = nodes
        self.H_idx = [[phys[e] for e in row] for row in self.H_idx]
        self.zero_slot = phys[self.zero_slot]
        self.h_conflicts = best

    # ---- device functions ----------------------------------------------------------------------
    def _function_table(self):
        """(name, inputs, outputs, qualifiers) of every generated device function."""
        x, u, th, lam = ("x", self.x), ("u", self.u), ("th", self.th), ("lam", self.lam)
        # the two slot evaluators are deliberately NOT inlined: they run once per chunk on a few lanes and would
        # otherwise dictate the register allocation (hence the occupancy) of the per-step hot loops
        # (option inline_eval: the two-trajectory backward kernel runs at the 255-register cap anyway, so inlining costs
        # no occupancy there -- but it does not pay either)
        inl = getattr(self, "inline_eval", -1)
        inl = False if inl < 0 else bool(inl)      # measured on the two-trajectory kernel: 0.6695 vs 0.6689 ms -- no gain
        fi, ni = "__device__ __forceinline__", "__device__ __noinline__"
        # terminal Hessians, dense row-major [hxx (n*n) | hxe (n*r)]
        term = [self.ddhxx.at(i, j) for i in range(self.n) for j in range(self.n)] + \
               [self.ddhxe.at(i, j) for i in range(self.n) for j in range(self.r)]
        # dense aux matrices for the legacy API: F G E Hxx Hxu Hxe Hux Huu Hue, each row-major
        dense = []
        for M in (self.dfx, self.dfu, self.dfe, self.ddHxx, self.ddHxu, self.ddHxe, self.ddHux, self.ddHuu, self.ddHue):
            dense += [M.at(i, j) for i in range(M.shape[0]) for j in range(M.shape[1])]
        return [("pdp_f_dyn", [x, u, th], self.dyn.elements(), fi), ("pdp_f_path_cost", [x, u, th], self.c.elements(), fi),
                ("pdp_f_final_cost", [x, th], self.h.elements(), fi), ("pdp_f_dHx", [x, u, lam, th], self.dHx.elements(), fi),
                ("pdp_f_dHu", [x, u, lam, th], self.dHu.elements(), fi), ("pdp_f_dhx", [x, th], self.dhx.elements(), fi),
                # all aux slots / only the dynamics-Jacobian slots
                ("pdp_f_aux_slots", [x, u, lam, th], self.slots.nodes, fi if inl else ni),
                ("pdp_f_dyn_slots", [x, u, th], self.slots.nodes[:self.nvar_s] or [S.ZERO], ni),
                ("pdp_f_terminal", [x, th], term, fi), ("pdp_f_aux_dense", [x, u, lam, th], dense, fi)]

    def _param_recips(self):
        """Divisors that depend on the parameters only (masses, inertias, lengths ...): their reciprocals are computed once
        per trajectory (``pdp_f_recips``) and kept behind the parameters in the same array, so the per-step code multiplies."""
        if getattr(self, "_recips", None) is None:
            outs = [e for _, _, o, _ in self._function_table() for e in o]
            self._recips = S.param_divisors(outs, {e.uid for e in self.th.elements()})
        return self._recips

    def _device_functions(self) -> str:
        rc = self._param_recips()
        recips = ("th", self.nth, rc)
        parts = [_emit_function(nm, ins, outs, lambda i: "out[%d]" % i, qualifiers=q, recips=recips)
                 for nm, ins, outs, q in self._function_table()]
        if rc:
            parts.append(_emit_function("pdp_f_recips", [("th", self.th)], [S.div(S.ONE, d) for d in rc],
                                        lambda i: "out[%d]" % i))
        return "\n\n".join(parts)

    # ---- the Riccati step body ---------------------------------------------------------------------
    def _backward_step(self) -> str:
        n, m, r, ns = self.n, self.m, self.r, self.ns
        nm = n + m
        L: List[str] = []
        ind = "      "
        # phase A: z = y . S   (lanes < n own row i of P in y0..y{n-1})
        L.append(ind + "// A: Z(i,:) = P(i,:) * [F|G|E]  -- structural non-zeros only, S broadcast from the chunk buffer")
        L.append(ind + "double " + ", ".join("z%d" % j for j in range(ns)) + ";")
        needed = {e[1] for row in self.S_ent for e in row if e[0] == "v"}
        load = _SlotLoader(L, "ar", needed, ind, "sa")
        acc = _Acc("z", ns, L, ind)
        for k in range(n):
            for j in range(ns):
                acc.add(j, "y%d" % k, self.S_ent[k][j], load)
        acc.finish()
        # phase B: transpose through shared memory
        L.append(ind + "// B: lane j picks up column j of Z; the auxvar lanes add their column of W")
        L.append(ind + "if (lane < %d) {" % n)
        for j in range(ns):
            L.append(ind + "  ZT[%d + lane] = z%d;" % (j * self.ldz, j))
        L.append(ind + "}")
        L.append(ind + "__syncwarp();")
        L.append(ind + "double " + ", ".join("c%d" % k for k in range(n)) + ";")
        L.append(ind + "{ const double* zr = ZT + lrow * %d;" % self.ldz)
        for k in range(n):
            L.append(ind + "  c%d = zr[%d];" % (k, k))
        L.append(ind + "}")
        L.append(ind + "if (lane >= %d) {" % nm)
        for k in range(n):
            L.append(ind + "  c%d += y%d;" % (k, k))
        L.append(ind + "}")
        # phase C: q = Hrow + c . [F|G]
        L.append(ind + "// C: Q(j,:) = Hstack(j,:) + Z(:,j)^T [F|G]   (H: per-lane indexed loads; [F|G] still in registers from A)")
        L.append(ind + "double " + ", ".join("q%d" % l for l in range(nm)) + ";")
        for l in range(nm):
            L.append(ind + "q%d = ar[ho%d];" % (l, l))
        if not getattr(self, "keep_fg", True):
            needed = {self.S_ent[k][l][1] for k in range(n) for l in range(nm) if self.S_ent[k][l][0] == "v"}
            load = _SlotLoader(L, "ar", needed, ind, "sc")
        acc = _Acc("q", nm, L, ind)
        acc.touched = [True] * nm
        early = getattr(self, "early_solve", False)
        # with early_solve the control columns (Quu / Qxu / Qeu, needed by the factorisation) are accumulated first and
        # the state columns afterwards, in the same basic block as the serial LDL^T chain so the scheduler can overlap them
        first_cols = range(n, nm) if early else range(nm)
        for k in range(n):
            for l in first_cols:
                acc.add(l, "c%d" % k, self.S_ent[k][l], load)
        late_lines: List[str] = []
        if early:
            acc_late = _Acc("q", nm, late_lines, ind)
            acc_late.touched = [True] * nm
            for k in range(n):
                for l in range(n):
                    acc_late.add(l, "c%d" % k, self.S_ent[k][l], load if getattr(self, "keep_fg", True) else load)
        # phase D: Quu to smem, LDL^T in every lane, solve for own right-hand side
        L.append(ind + "// D: Quu = rows n..n+m-1; every lane factors it (LDL^T, uniform) and solves for its own column")
        L.append(ind + "if (lane >= %d && lane < %d) {" % (n, nm))
        for a in range(m):
            L.append(ind + "  QUU[(lane - %d) * %d + %d] = q%d;" % (n, m, a, n + a))
        L.append(ind + "}")
        L.append(ind + "__syncwarp();")
        # LDL^T: A = L D L^T, unit lower L.  d_i, l_ij (i>j)
        for i in range(m):
            for j in range(i + 1):
                L.append(ind + "double a%d%d = QUU[%d];" % (i, j, i * m + j))
        for j in range(m):
            # d_j = a_jj - sum_k l_jk^2 d_k
            expr = "a%d%d" % (j, j)
            for k in range(j):
                expr = "fma(-l%d%d * l%d%d, d%d, %s)" % (j, k, j, k, k, expr)
            L.append(ind + "const double d%d = %s;" % (j, expr))
            L.append(ind + "bad |= !(d%d > 0.0);" % j)
            if getattr(self, "fast_rcp", False):
                # reciprocal = hardware seed + two Newton steps (<= 1 ulp; no special-case slow path in the chain)
                L.append(ind + "double r%d; asm(\"rcp.approx.ftz.f64 %%0, %%1;\" : \"=d\"(r%d) : \"d\"(d%d));" % (j, j, j))
                L.append(ind + "r%d = fma(r%d, fma(-d%d, r%d, 1.0), r%d);" % (j, j, j, j, j))
                L.append(ind + "r%d = fma(r%d, fma(-d%d, r%d, 1.0), r%d);" % (j, j, j, j, j))
            else:
                L.append(ind + "const double r%d = 1.0 / d%d;" % (j, j))
            for i in range(j + 1, m):
                expr = "a%d%d" % (i, j)
                for k in range(j):
                    expr = "fma(-l%d%d * l%d%d, d%d, %s)" % (i, k, j, k, k, expr)
                L.append(ind + "const double l%d%d = (%s) * r%d;" % (i, j, expr, j))
            if late_lines:   # a slice of the independent state-column FMAs after every pivot
                take = (len(late_lines) + (m - j) - 1) // (m - j)
                L.extend(late_lines[:take])
                del late_lines[:take]
        # solve L D L^T v = -rhs ; rhs = q[n..n+m-1]
        for i in range(m):
            expr = "-q%d" % (n + i)
            for k in range(i):
                expr = "fma(-l%d%d, w%d, %s)" % (i, k, k, expr)
            L.append(ind + "const double w%d = %s;" % (i, expr))
        for i in reversed(range(m)):
            expr = "w%d * r%d" % (i, i)
            for k in range(i + 1, m):
                expr = "fma(-l%d%d, v%d, %s)" % (k, i, k, expr)
            L.append(ind + "const double v%d = %s;" % (i, expr))
        # phase E: spill gains, share K, update
        L.append(ind + "// E: spill (K|k) for the forward pass, broadcast K, rank-m update of the stack")
        L.append(ind + "if (lane < %d) {" % n)
        for a in range(m):
            L.append(ind + "  KS[%d + lane] = v%d;" % (a * self.ldk, a))
        L.append(ind + "}")
        L.append(ind + "if (gslot >= 0) {")
        L.append(ind + "  double* gp = gains + ((size_t)b * H + t) * %d + gslot * %d;" % ((n + r) * m, m))
        L.append(self._vec_store("gp", ["v%d" % a for a in range(m)], ind + "  "))
        L.append(ind + "}")
        L.append(ind + "__syncwarp();")
        for a in range(m):
            if self.ldk % 2 == 0:
                for l in range(0, n - 1, 2):
                    L.append(ind + "{ const double2 kk = *reinterpret_cast<const double2*>(KS + %d); q%d = fma(q%d, kk.x, q%d); q%d = fma(q%d, kk.y, q%d); }"
                             % (a * self.ldk + l, l, n + a, l, l + 1, n + a, l + 1))
                if n % 2 == 1:
                    L.append(ind + "q%d = fma(q%d, KS[%d], q%d);" % (n - 1, n + a, a * self.ldk + n - 1, n - 1))
            else:
                for l in range(n):
                    L.append(ind + "q%d = fma(q%d, KS[%d], q%d);" % (l, n + a, a * self.ldk + l, l))
        for l in range(n):
            L.append(ind + "y%d = q%d;" % (l, l))
        return "\n".join(L)

    def _ldlt_lines(self, L, ind, late_lines):
        """Uniform LDL^T of Quu (read from QUU) with the reciprocal pivots r_j; ``late_lines`` (independent FMAs)
        are sliced in after every pivot so the scheduler can overlap them with the serial chain."""
        m = self.m
        for i in range(m):
            for j in range(i + 1):
                L.append(ind + "double a%d%d = QUU[%d];" % (i, j, i * m + j))
        for j in range(m):
            expr = "a%d%d" % (j, j)
            for k in range(j):
                expr = "fma(-l%d%d * l%d%d, d%d, %s)" % (j, k, j, k, k, expr)
            L.append(ind + "const double d%d = %s;" % (j, expr))
            L.append(ind + "bad |= !(d%d > 0.0);" % j)
            if getattr(self, "fast_rcp", False):
                L.append(ind + "double r%d; asm(\"rcp.approx.ftz.f64 %%0, %%1;\" : \"=d\"(r%d) : \"d\"(d%d));" % (j, j, j))
                L.append(ind + "r%d = fma(r%d, fma(-d%d, r%d, 1.0), r%d);" % (j, j, j, j, j))
                L.append(ind + "r%d = fma(r%d, fma(-d%d, r%d, 1.0), r%d);" % (j, j, j, j, j))
            else:
                L.append(ind + "const double r%d = 1.0 / d%d;" % (j, j))
            for i in range(j + 1, m):
                expr = "a%d%d" % (i, j)
                for k in range(j):
                    expr = "fma(-l%d%d * l%d%d, d%d, %s)" % (i, k, j, k, k, expr)
                L.append(ind + "const double l%d%d = (%s) * r%d;" % (i, j, expr, j))
            if late_lines:
                take = (len(late_lines) + (m - j) - 1) // (m - j)
                L.extend(late_lines[:take])
                del late_lines[:take]

    def _solve_lines(self, L, ind, rhs, tag):
        """v = -(L D L^T)^{-1} rhs for the right-hand side names ``rhs``; results in v<tag>_i."""
        m = self.m
        for i in range(m):
            expr = "-%s" % rhs[i]
            for k in range(i):
                expr = "fma(-l%d%d, w%s_%d, %s)" % (i, k, tag, k, expr)
            L.append(ind + "const double w%s_%d = %s;" % (tag, i, expr))
        for i in reversed(range(m)):
            expr = "w%s_%d * r%d" % (tag, i, i)
            for k in range(i + 1, m):
                expr = "fma(-l%d%d, v%s_%d, %s)" % (k, i, tag, k, expr)
            L.append(ind + "const double v%s_%d = %s;" % (tag, i, expr))

    def _zero_z_columns(self):
        """Columns j of Z = P [F|G|E] that are structurally zero (no entry of column j of [F|G|E]); ns <= 32."""
        return {j for j in range(self.ns) if all(self.S_ent[k][j][0] == "z" for k in range(self.n))}

    def _h_loads(self):
        """[(slot, column group)] in load order (two-trajectory kernel)."""
        return [(sl, g) for sl, cg in enumerate(self.h_groups) for g in cg]

    def _h_multi_index(self, sl, g):
        """Running index of a multi-column group among all multi-column groups (its 4-bit member code lives there)."""
        multi = [(a, b) for a, b in self._h_loads() if len(b) > 1]
        return multi.index((sl, g))

    def _hidx8(self) -> bool:
        """Two-trajectory kernel: pack four 8-bit Hamiltonian slot indices per register (needs < 256 slots)."""
        return self.nvar <= 255

    def _h_init_lines(self, L, ind, slots):
        """q<slot>_<l> = Hstack entry of the lane's row (column-coloured loads, see _h_column_groups)."""
        nm = self.n + self.m
        if self._hidx8() and self.h_groups is not None:
            assigned = set()
            for k, (sl, g) in enumerate(self._h_loads()):
                if sl not in slots:
                    continue
                idx = "(ho%d >> %d) & 0xffu" % (k // 4, 8 * (k % 4))
                if len(g) == 1:
                    L.append(ind + "q%d_%d = ar[%s];" % (sl, g[0], idx))
                else:
                    c = self._h_multi_index(sl, g)
                    L.append(ind + "{ const double hv = ar[%s]; const unsigned cd = (hc%d >> %d) & 0xfu;" % (idx, c // 8, 4 * (c % 8)))
                    for pos, l in enumerate(g):
                        L.append(ind + "  q%d_%d = (cd == %du) ? hv : 0.0;" % (sl, l, pos))
                    L.append(ind + "}")
                assigned |= {(sl, l) for l in g}
            for sl in slots:
                for l in range(nm):
                    if (sl, l) not in assigned:
                        L.append(ind + "q%d_%d = 0.0;" % (sl, l))
        else:
            for l in range(nm):
                if 0 in slots:
                    L.append(ind + "q0_%d = ar[ho%d & 0xffffu];" % (l, l))
                if 1 in slots:
                    L.append(ind + "q1_%d = ar[ho%d >> 16];" % (l, l))

    def _phase_c_joint(self, L, ind):
        """Phases B(pick-up) / C / first half of D for both slots together: one operand load feeds two FMAs."""
        n, m, r, ns = self.n, self.m, self.r, self.ns
        nm = n + m
        load = self._a_loader
        L.append(ind + "double " + ", ".join("c0_%d, c1_%d" % (k, k) for k in range(n)) + ";")
        L.append(ind + "{ const double* zr = ZT + zr0 * %d;" % self.ldz)
        for k in range(n):
            L.append(ind + "  c0_%d = zr[%d];" % (k, k))
        L.append(ind + "}")
        L.append(ind + "{ const double* zr = ZT + zr1 * %d;" % self.ldz)
        for k in range(n):
            L.append(ind + "  c1_%d = zr[%d];" % (k, k))
        L.append(ind + "}")
        L.append(ind + "if (wrow) {")
        for k in range(n):
            L.append(ind + "  c1_%d += y1_%d;" % (k, k))
        L.append(ind + "}")
        L.append(ind + "// C: Q(j,:) = Hstack(j,:) + Z(:,j)^T [F|G] for both rows (one operand load, two FMAs)")
        L.append(ind + "double " + ", ".join("q0_%d, q1_%d" % (l, l) for l in range(nm)) + ";")
        self._h_init_lines(L, ind, (0, 1))
        if not getattr(self, "keep_fg", True):
            needed = {self.S_ent[k][l][1] for k in range(n) for l in range(nm) if self.S_ent[k][l][0] == "v"}
            load = _SlotLoader(L, "ar", needed, ind, "sc")
        acc0, acc1 = _Acc("q0_", nm, L, ind), _Acc("q1_", nm, L, ind)
        acc0.touched = [True] * nm
        acc1.touched = [True] * nm
        early = getattr(self, "early_solve", False)
        first_cols = range(n, nm) if early else range(nm)
        for k in range(n):
            for l in first_cols:
                acc0.add(l, "c0_%d" % k, self.S_ent[k][l], load)
                acc1.add(l, "c1_%d" % k, self.S_ent[k][l], load)
        late_lines: List[str] = []
        if early:
            late0, late1 = _Acc("q0_", nm, late_lines, ind), _Acc("q1_", nm, late_lines, ind)
            late0.touched = [True] * nm
            late1.touched = [True] * nm
            lload = load       # operand loads are appended to L here, i.e. ahead of the pivots the FMAs are sliced between
            for k in range(n):
                for l in range(n):
                    late0.add(l, "c0_%d" % k, self.S_ent[k][l], lload)
                    late1.add(l, "c1_%d" % k, self.S_ent[k][l], lload)
        L.append(ind + "// D: Quu = slot-1 rows of team lanes < m; every lane factors it (LDL^T, uniform per half) and solves for its two columns")
        L.append(ind + "if (tl < %d) {" % m)
        for a in range(m):
            L.append(ind + "  QUU[tl * %d + %d] = q1_%d;" % (m, a, n + a))
        L.append(ind + "}")
        L.append(ind + "__syncwarp();")
        return late_lines

    def _backward_step2(self) -> str:
        """Riccati step of the two-trajectories-per-warp kernel: team lane tl owns stack rows tl (slot 0: P) and
        n + tl (slot 1: control rows, then the columns of W).  Same phases A-E as :meth:`_backward_step`."""
        n, m, r, ns = self.n, self.m, self.r, self.ns
        nm = n + m
        L: List[str] = []
        ind = "      "
        L.append(ind + "// A: Z(i,:) = P(i,:) * [F|G|E]  -- structural non-zeros only; each half-warp reads its own trajectory's slots")
        L.append(ind + "double " + ", ".join("z%d" % j for j in range(ns)) + ";")
        needed = {e[1] for row in self.S_ent for e in row if e[0] == "v"}
        load = _SlotLoader(L, "ar", needed, ind, "sa")
        self._a_loader = load
        acc = _Acc("z", ns, L, ind)
        for k in range(n):
            for j in range(ns):
                acc.add(j, "y0_%d" % k, self.S_ent[k][j], load)
        acc.finish()
        L.append(ind + "// B: transpose through shared memory: slot 0 picks up column tl of Z, slot 1 column n + tl (+ its W column)")
        zero_cols = self._zero_z_columns()
        L.append(ind + "if (tl < %d) {" % n)
        for j in range(ns):
            if j not in zero_cols:
                L.append(ind + "  ZT[%d + tl] = z%d;" % (j * self.ldz, j))
        L.append(ind + "}")
        L.append(ind + "__syncwarp();")
        late_lines = self._phase_c_joint(L, ind)
        self._ldlt_lines(L, ind, late_lines)
        self._solve_lines(L, ind, ["q0_%d" % (n + i) for i in range(m)], "0")
        self._solve_lines(L, ind, ["q1_%d" % (n + i) for i in range(m)], "1")
        L.append(ind + "// E: spill (K|k) for the forward pass, broadcast K, rank-m update of both rows")
        L.append(ind + "if (tl < %d) {" % n)
        for a in range(m):
            L.append(ind + "  KS[%d + tl] = v0_%d;" % (a * self.ldk, a))
        L.append(ind + "}")
        L.append(ind + "if (live) {")
        L.append(ind + "  double* gp = gains + ((size_t)b * H + t) * %d;" % ((n + r) * m))
        L.append(ind + "  if (gslot0 >= 0) {")
        L.append(self._vec_store("(gp + gslot0 * %d)" % m, ["v0_%d" % a for a in range(m)], ind + "    "))
        L.append(ind + "  }")
        L.append(ind + "  if (gslot1 >= 0) {")
        L.append(self._vec_store("(gp + gslot1 * %d)" % m, ["v1_%d" % a for a in range(m)], ind + "    "))
        L.append(ind + "  }")
        L.append(ind + "}")
        L.append(ind + "__syncwarp();")
        for a in range(m):
            if self.ldk % 2 == 0:
                for l in range(0, n - 1, 2):
                    L.append(ind + "{ const double2 kk = *reinterpret_cast<const double2*>(KS + %d); "
                             "q0_%d = fma(q0_%d, kk.x, q0_%d); q1_%d = fma(q1_%d, kk.x, q1_%d); "
                             "q0_%d = fma(q0_%d, kk.y, q0_%d); q1_%d = fma(q1_%d, kk.y, q1_%d); }"
                             % (a * self.ldk + l, l, n + a, l, l, n + a, l, l + 1, n + a, l + 1, l + 1, n + a, l + 1))
                if n % 2 == 1:
                    L.append(ind + "{ const double kk = KS[%d]; q0_%d = fma(q0_%d, kk, q0_%d); q1_%d = fma(q1_%d, kk, q1_%d); }"
                             % (a * self.ldk + n - 1, n - 1, n + a, n - 1, n - 1, n + a, n - 1))
            else:
                for l in range(n):
                    L.append(ind + "{ const double kk = KS[%d]; q0_%d = fma(q0_%d, kk, q0_%d); q1_%d = fma(q1_%d, kk, q1_%d); }"
                             % (a * self.ldk + l, l, n + a, l, l, n + a, l))
        for l in range(n):
            L.append(ind + "y0_%d = q0_%d; y1_%d = q1_%d;" % (l, l, l, l))
        return "\n".join(L)

    def _vec_store(self, ptr, names, ind):
        m = len(names)
        out = []
        if m % 2 == 0:
            for a in range(0, m, 2):
                out.append("%s*reinterpret_cast<double2*>(%s + %d) = make_double2(%s, %s);" % (ind, ptr, a, names[a], names[a + 1]))
        else:
            for a in range(m):
                out.append("%s%s[%d] = %s;" % (ind, ptr, a, names[a]))
        return "\n".join(out)

    fwd_smem_budget = 18 * 1024

    def _nthx(self):
        """Length of a trajectory's parameter array in the kernels: theta followed by the hoisted reciprocals."""
        return self.nth + len(self._param_recips())

    def _fwd_group_stride(self):
        """Lanes per trajectory group of the forward kernel: r rounded up to even (see the kernel template)."""
        r = max(self.r, 1)
        return min(r + (r & 1), WARP) if r < WARP else WARP

    def _fwd_shape(self):
        """(trajectories per warp, chunk length) of the forward kernel: lane g*r + c owns column c of trajectory g.
        The chunk is capped so that one warp's regions stay within ``fwd_smem_budget`` bytes of shared memory
        (18 KB: twelve warps per SM)."""
        gs = self._fwd_group_stride()
        fg = getattr(self, "fwd_pack", 0) or max(1, min(WARP // gs, 4))
        fg = max(1, min(fg, WARP // gs, WARP))
        ch = getattr(self, "fwd_chunk", 0)
        if not ch:
            per_step = _pad_ld(self.nvar_s) + self.n + self.m
            fit = (self.fwd_smem_budget // (8 * fg) - self.n * self.m - max(self._nthx(), 1) - 16) // per_step
            ch = min(WARP // fg if fg > 1 else self.chunk, max(fit, 1))
        ch = max(1, min(ch, WARP // fg))
        return fg, ch

    def _forward_step(self) -> str:
        n, m, r, ns = self.n, self.m, self.r, self.ns
        L: List[str] = []
        ind = "      "
        L.append(ind + "// U(:,c) = k(:,c) + K X(:,c); K staged as [l][a] per trajectory; two partial sums per row shorten the chains")
        half = (n + 1) // 2
        # operand loads: with several trajectories per warp every group of lanes reads its own trajectory's word; a
        # 64-bit load serves up to four such groups in ONE shared-memory wavefront, a 128-bit load needs one pass per
        # quarter-warp (measured: tools/microbench/smem_wavefronts.cu), so the packed kernel uses scalar loads
        fv = getattr(self, "fwd_vec", -1)
        vec = (self._fwd_shape()[0] == 1) if fv < 0 else bool(fv)
        # (volatile pointers: nvcc would otherwise prove the alignment and fuse neighbouring loads back into LDS.128)
        ks, arn = ("KS", "ar") if vec else ("vKS", "var")
        if not vec:
            L.append(ind + "const volatile double* vKS = KS; const volatile double* var = ar;")
        for a in range(m):
            L.append(ind + "double u%d = g%d, ub%d = 0.0;" % (a, a, a))
        for l in range(n):
            a = 0
            while a < m:
                if m % 2 == 0 and vec:
                    L.append(ind + "{ const double2 kk = *reinterpret_cast<const double2*>(KS + %d);" % (l * m + a))
                    for d, comp in ((0, "x"), (1, "y")):
                        tgt = "u%d" % (a + d) if l < half else "ub%d" % (a + d)
                        L.append(ind + "  %s = fma(kk.%s, x%d, %s);" % (tgt, comp, l, tgt))
                    L.append(ind + "}")
                    a += 2
                else:
                    tgt = "u%d" % a if l < half else "ub%d" % a
                    L.append(ind + "%s = fma(%s[%d], x%d, %s);" % (tgt, ks, l * m + a, l, tgt))
                    a += 1
        for a in range(m):
            L.append(ind + "u%d += ub%d;" % (a, a))
        L.append(ind + "// X+(:,c) = F X(:,c) + G U(:,c) + E(:,c)")
        L.append(ind + "double " + ", ".join("n%d" % i for i in range(n)) + ";")
        needed = {e[1] for row in self.S_ent for e in row if e[0] == "v"} if vec else set()
        load = _SlotLoader(L, arn, needed, ind, "sf")
        acc = _Acc("n", n, L, ind)
        for i in range(n):
            for l in range(n):
                acc.add(i, "x%d" % l, self.S_ent[i][l], load)
            for a in range(m):
                acc.add(i, "u%d" % a, self.S_ent[i][n + a], load)
        acc.finish()
        # E column: predicated adds (E is sparse for mechanical systems)
        for i in range(n):
            for c in range(r):
                ent = self.S_ent[i][n + m + c]
                if ent[0] == "z":
                    continue
                if ent[0] == "c":
                    val = _lit(ent[1])
                else:
                    val = ("-" if ent[2] < 0 else "") + load(ent[1])
                L.append(ind + "n%d += (col == %d) ? %s : 0.0;" % (i, c, val))
        return "\n".join(L)

    def _fwd_gain_prefetch(self, fg):
        """K of the FG trajectories is fetched cooperatively (unit q = lane + 32 j), the k column by its owner lane."""
        n, m, r = self.n, self.m, self.r
        kw = n * m
        vec = (m % 2 == 0)
        unit = 2 if vec else 1
        per = kw // unit
        nq = (fg * per + WARP - 1) // WARP
        ty = "double2" if vec else "double"
        setup, load, store = [], [], []
        for j in range(nq):
            setup.append("  const int kq%d = lane + %d;" % (j, WARP * j))
            setup.append("  const bool kv%d = kq%d < %d;" % (j, j, fg * per))
            setup.append("  const int kg%d = kv%d ? kq%d / %d : 0, ke%d = kq%d - (kq%d / %d) * %d;" % (j, j, j, per, j, j, j, per, per))
            setup.append("  const double* kp%d = gains + (size_t)((b0 + kg%d < B) ? b0 + kg%d : B - 1) * H * PDP_GREC + ke%d * %d;"
                         % (j, j, j, j, unit))
            setup.append("  double* kd%d = wbase + kg%d * PDP_FTS + PDP_FOFF_KS + ke%d * %d;" % (j, j, j, unit))
            setup.append("  %s kn%d = %s;" % (ty, j, "make_double2(0.0, 0.0)" if vec else "0.0"))
            load.append("        if (kv%d) kn%d = *reinterpret_cast<const %s*>(kp%d + ro);" % (j, j, ty, j))
            store.append("      if (kv%d) *reinterpret_cast<%s*>(kd%d) = kn%d;" % (j, ty, j, j))
        gl = ["        const size_t ro = (size_t)(t + 1) * PDP_GREC;", "        if (col >= 0) {",
              "          const double* gp = gains + (size_t)bg * H * PDP_GREC + ro + (PDP_N + col) * PDP_M;"]
        if vec:
            gl += ["          { const double2 gg = *reinterpret_cast<const double2*>(gp + %d); gn%d = gg.x; gn%d = gg.y; }" % (a, a, a + 1)
                   for a in range(0, m, 2)]
        else:
            gl += ["          gn%d = gp[%d];" % (a, a) for a in range(m)]
        gl.append("        }")
        return "\n".join(setup), "\n".join(gl + load), "\n".join(store)

    # ---- whole translation unit ----------------------------------------------------------------
    def source(self) -> str:
        n, m, r, ns = self.n, self.m, self.r, self.ns
        nm = n + m
        zt_size = _even((ns + 1) * self.ldz)      # + one all-zero row (two-trajectory kernel, structurally zero columns)
        ks_size = _even(m * self.ldk)
        auxc_size = _even(max(self.chunk * self.auxld, n * n + n * r))
        off_zt = auxc_size
        off_ks = off_zt + zt_size
        off_quu = off_ks + ks_size
        off_th = off_quu + _even(m * m)
        warp_doubles = _even(off_th + max(self._nthx(), 1))
        bp = getattr(self, "bwd_pack", 1)
        half_stride = _pad_ld(warp_doubles)       # two-trajectory kernel: per-trajectory regions = 2 (mod 16) doubles apart
        if bp == 2:
            warp_doubles = 2 * half_stride
        # forward kernel, per trajectory region: [CHF][FLD] dynamics slots | K [l][a] | TH | residuals [CHF][n+m]
        fg, chf = self._fwd_shape()
        fld = _pad_ld(self.nvar_s)
        foff_ks = _even(chf * fld)
        foff_th = foff_ks + _even(n * m)
        foff_dl = foff_th + _even(max(self._nthx(), 1))           # residuals x - xref [CHF*n], u - uref [CHF*m]
        foff_du = foff_dl + chf * n
        fts = _pad_ld(foff_du + chf * m)
        fwarp_doubles = max(fg * fts, WARP)
        defs = {
            "N": n, "M": m, "R": r, "NS": ns, "NM": nm, "NVAR": self.nvar, "NVAR_S": self.nvar_s,
            "AUXLD": self.auxld, "CH": self.chunk, "WPB": self.wpb, "LDZ": self.ldz,
            "LDK": self.ldk, "OFF_ZT": off_zt, "OFF_KS": off_ks, "OFF_QUU": off_quu, "OFF_TH": off_th,
            "FLD": fld, "FOFF_KS": foff_ks, "FOFF_TH": foff_th, "FOFF_DL": foff_dl, "FG": fg, "CHF": chf, "FTS": fts,
            "FOFF_DU": foff_du, "FGS": self._fwd_group_stride(),
            "FWARP_DOUBLES": fwarp_doubles, "WPBF": getattr(self, "wpbf", 4), "MINBF": getattr(self, "min_blocks_f", 1),
            "WARP_DOUBLES": warp_doubles, "NTH": self.nth, "NRCP": self._nthx() - self.nth, "NTHX": max(self._nthx(), 1), "MINB": getattr(self, "min_blocks", 1), "GREC": (n + r) * m,
            "NDENSE": n * n + n * m + n * r + n * n + n * m + n * r + m * n + m * m + m * r,
            "BP": bp, "HS": half_stride,
            "ZMASK": sum(1 << j for j in self._zero_z_columns()) if (bp == 2 and hasattr(self, "S_ent")) else 0,
            "PF": getattr(self, "prefetch", 0), "PFD": max(1, getattr(self, "prefetch_dist", 3)),
            "PFL": max(0, getattr(self, "prefetch_l1_lead", 0)),
        }
        if getattr(self, "rollout_tma", 0):
            # thread-private slots of the TMA rollout kernel (doubles): x rows / u rows in (double-buffered), x-or-lambda rows and
            # dH/du rows out, two mbarriers; +2 per slot for the parity shift and the rounding of a copy to 16 bytes.  The chunk
            # is the longest whose slots leave room for two warps per SM (<= 110 KB per warp), at most 16 steps.
            def tma_layout(tc):
                txs, tus, tls = _even(tc * n + 2), _even(tc * m + 2), _even((tc + 1) * n + 2)
                tstride = 2 * (txs + tus) + tls + tus + 2
                while tstride % 4 != 2:          # 2 (mod 4) doubles: the 16 lanes of a half-warp spread over 8 bank pairs
                    tstride += 2
                return txs, tus, tls, tstride
            tc = self.tma_chunk
            if not tc:
                tc = 1
                while tc < 16 and 32 * 8 * tma_layout(tc + 1)[3] <= 110 * 1024:
                    tc += 1
            txs, tus, tls, tstride = tma_layout(tc)
            defs.update({"TC": tc, "TB": 32, "TXS": txs, "TUS": tus, "TLS": tls, "TSTRIDE": tstride})
        header = ["// GENERATED by pontryagin_differentiable_programming_b200/codegen.py -- do not edit",
                  "#include <cuda_runtime.h>", "#include <math.h>", "#include <stdint.h>"]
        header += ["#define PDP_%s %d" % kv for kv in defs.items()]
        tables = []
        hid = []
        if bp == 2:
            # packed per-lane slot indices: low half = row of slot 0, high half = row of slot 1 (idle lanes repeat a row)
            def pair(l, lane):
                tl = lane & 15
                ra, rb = (tl if tl < n else 0), n + (tl if tl < m + r else 0)
                return self.H_idx[ra][l], self.H_idx[rb][l]
            hcode = []
            if self._hidx8():
                loads = self._h_loads()
                multi = [lg for lg in loads if len(lg[1]) > 1]

                def lane_row(sl, lane):
                    tl = lane & 15
                    return (tl if tl < n else 0) if sl == 0 else n + (tl if tl < m + r else 0)
                nw = (len(loads) + 3) // 4
                for w in range(nw):
                    for lane in range(WARP):
                        word = 0
                        for q, (sl, g) in enumerate(loads[4 * w:4 * w + 4]):
                            word |= self._h_group_entry(lane_row(sl, lane), g)[0] << (8 * q)
                        hid.append(word)
                for w in range((len(multi) + 7) // 8):
                    for lane in range(WARP):
                        word = 0
                        for q, (sl, g) in enumerate(multi[8 * w:8 * w + 8]):
                            word |= self._h_group_entry(lane_row(sl, lane), g)[1] << (4 * q)
                        hcode.append(word)
            else:
                nw = nm
                for l in range(nm):
                    for lane in range(WARP):
                        a0, a1 = pair(l, lane)
                        hid.append(a0 | (a1 << 16))
            tables.append("__device__ const unsigned int pdp_hidx[%d] = {%s};" % (len(hid), ", ".join(map(str, hid))))
            tabload = "\n".join("  const unsigned int ho%d = pdp_hidx[%d + lane];" % (w, w * WARP) for w in range(nw))
            if hcode:
                tables.append("__device__ const unsigned int pdp_hcode[%d] = {%s};" % (len(hcode), ", ".join(map(str, hcode))))
                tabload += "\n" + "\n".join("  const unsigned int hc%d = pdp_hcode[%d + lane];" % (w, w * WARP)
                                              for w in range(len(hcode) // WARP))
            ydecl = "double " + ", ".join("y0_%d = 0.0, y1_%d = 0.0" % (k, k) for k in range(n)) + ";"
            term_init = []
            for k in range(n):
                term_init.append("    if (tl < %d) y0_%d = TB[tl * %d + %d];" % (n, k, n, k))
                term_init.append("    if (wrow) y1_%d = TB[%d + %d * %d + (tl - %d)];" % (k, n * n, k, r, m))
        else:
            for l in range(nm):
                hid += [self.H_idx[j][l] if j < ns else self.zero_slot for j in range(WARP)]
            tables.append("__device__ const unsigned short pdp_hidx[%d] = {%s};" % (len(hid), ", ".join(map(str, hid))))
            tabload = "\n".join("  const int ho%d = pdp_hidx[%d + lane];" % (l, l * WARP) for l in range(nm))
            ydecl = "double " + ", ".join("y%d = 0.0" % k for k in range(n)) + ";"
            # terminal init: lane i<n takes row i of hxx, lane n+m+c takes column c of hxe
            term_init = []
            for k in range(n):
                term_init.append("    y%d = (lane < %d) ? TB[lane * %d + %d] : ((lane >= %d && lane < %d) ? TB[%d + %d * %d + (lane - %d)] : 0.0);"
                                 % (k, n, n, k, nm, ns, n * n, k, r, nm))
        xdecl = "double " + ", ".join("x%d" % k for k in range(n)) + ";"
        xinit = "\n".join("    x%d = (X0a != nullptr && col >= 0) ? X0a[(size_t)(x0a_stride ? bg : 0) * %d + %d * %d + col] : 0.0;"
                          % (k, n * r, k, r) for k in range(n))
        gndecl = "double " + ", ".join("gn%d = 0.0" % a for a in range(m)) + ";"
        gcur = "      const double " + ", ".join("g%d = gn%d" % (a, a) for a in range(m)) + ";"
        kq_setup, gnload, ks_store = self._fwd_gain_prefetch(fg)
        # streaming (evict-first) stores for dX / dU: nothing on the device re-reads them, so they should not displace
        # the gain records / chunk rows in L2 (measured r2a: forward kernel 0.4033 -> 0.4002 ms)
        xstore = "\n".join("          __stcs(o + %d, n%d);" % (i * r, i) for i in range(n))
        ustore = "\n".join("          __stcs(o + %d, u%d);" % (a * r, a) for a in range(m))
        xcopy = "\n".join("      x%d = n%d;" % (k, k) for k in range(n))
        x0store = "\n".join("    o[%d] = x%d;" % (i * r, i) for i in range(n))

        rep = {
            "@@TABLOAD@@": tabload, "@@YDECL@@": ydecl,
            "@@TERM_INIT@@": "\n".join(term_init), "@@BACKWARD_STEP@@": self._backward_step2() if bp == 2 else self._backward_step(),
            "@@XDECL@@": xdecl, "@@XINIT@@": xinit, "@@GNDECL@@": gndecl, "@@GNLOAD@@": gnload, "@@GCUR@@": gcur, "@@KQ_SETUP@@": kq_setup,
            "@@KS_STORE@@": ks_store, "@@FORWARD_STEP@@": self._forward_step(), "@@XSTORE@@": xstore, "@@USTORE@@": ustore,
            "@@XCOPY@@": xcopy, "@@X0STORE@@": x0store,
            "@@DPACC@@": "\n".join(["        dpacc = fma(dlx[%d], x%d, dpacc);" % (i, i) for i in range(n)] +
                                    ["        dpacc = fma(dlu[%d], u%d, dpacc);" % (a, a) for a in range(m)]),
            "@@DPTERM@@": "\n".join("      { const double d = xh[%d] - xrh[%d]; dpacc = fma(d, x%d, dpacc); lt = fma(d, d, lt); }"
                                     % (i, i, i) for i in range(n)),
            "@@XCHK@@": "\n".join("    chk += x%d;" % k for k in range(n)),
        }
        rep.update(self._eval_macros())
        kernels = self._kernel_text()
        for k, v in rep.items():
            kernels = kernels.replace(k, v)
        header.append("#define PDP_KIND %d" % self.kind_id)
        return ("\n".join(header) + "\n\n" + "\n".join(tables + self._extra_tables()) + "\n\n" +
                self._device_functions() + "\n\n" + kernels)

    kind_id = 1

    def _extra_tables(self):
        return []

    def _kernel_text(self):
        bwd = _K_AUX_LQR_BWD2 if getattr(self, "bwd_pack", 1) == 2 else _K_AUX_LQR_BWD
        tma, launch_common = "", _K_LAUNCH_COMMON
        if getattr(self, "rollout_tma", 0):
            from .kernel_templates import K_ROLLOUT_TMA, rollout_tma_launcher
            tma, launch_common = K_ROLLOUT_TMA, rollout_tma_launcher(_K_LAUNCH_COMMON)
        return _K_PRELUDE + _K_TMA_PRIMS + _K_ROLLOUT_AUXEVAL + tma + _K_AUX_LQR_HEAD + bwd + _K_AUX_LQR_FWD + launch_common + _K_LAUNCH_LQR

    def _eval_macros(self):
        el = "tl" if getattr(self, "bwd_pack", 1) == 2 else "lane"       # evaluation lane = time step of the chunk
        return {
            "@@EVAL_TERM@@": "  if (%s == 0) pdp_f_terminal(Xb + (size_t)H * PDP_N, TH, TB);" % el,
            "@@EVAL_AUX_CHUNK@@": """    {
      const int te = tc + %(el)s;
      if (%(el)s < PDP_CH && te < H)
        pdp_f_aux_slots(Xb + (size_t)te * PDP_N, Ub + (size_t)te * PDP_M, Lb + (size_t)te * PDP_N, TH, auxc + %(el)s * PDP_AUXLD);
    }""" % {"el": el},
            "@@EVAL_DYN@@": "        pdp_f_dyn_slots(X + ((size_t)be * (H + 1) + te) * PDP_N, U + ((size_t)be * H + te) * PDP_M, the, eo);",
            "@@EVAL_DYN_COOP@@": "",
            "@@PREFETCH_AUX_CHUNK@@": "",      # measured: no gain (two-trajectory kernel) / a loss (one-trajectory kernel)
            "@@PREFETCH_DYN_CHUNK@@": """#if PDP_PF
    {
      const int tp = tc + PDP_CHF + se;
      if (evl && tp < H) {
        const double* xp = X + ((size_t)be * (H + 1) + tp) * PDP_N;
        const double* up = U + ((size_t)be * H + tp) * PDP_M;
        pdp_prefetch(xp); pdp_prefetch(xp + (PDP_N - 1)); pdp_prefetch(up); pdp_prefetch(up + (PDP_M - 1));
        if (fused) {
          const double* xr = Xref + ((size_t)be * (H + 1) + tp) * PDP_N;
          pdp_prefetch(xr); pdp_prefetch(xr + (PDP_N - 1));
          if (Uref) { pdp_prefetch(Uref + ((size_t)be * H + tp) * PDP_M); pdp_prefetch(Uref + ((size_t)be * H + tp) * PDP_M + (PDP_M - 1)); }
        }
      }
    }
#endif""",
            "@@PREFETCH_DYN_CHUNK_L1@@": """    {
      const int tp = tc + PDP_CHF + se;
      if (evl && tp < H) {
        const double* xp = X + ((size_t)be * (H + 1) + tp) * PDP_N;
        const double* up = U + ((size_t)be * H + tp) * PDP_M;
        pdp_prefetch_l1(xp); pdp_prefetch_l1(xp + (PDP_N - 1)); pdp_prefetch_l1(up); pdp_prefetch_l1(up + (PDP_M - 1));
        if (fused) {
          const double* xr = Xref + ((size_t)be * (H + 1) + tp) * PDP_N;
          pdp_prefetch_l1(xr); pdp_prefetch_l1(xr + (PDP_N - 1));
          if (Uref) { pdp_prefetch_l1(Uref + ((size_t)be * H + tp) * PDP_M); pdp_prefetch_l1(Uref + ((size_t)be * H + tp) * PDP_M + (PDP_M - 1)); }
        }
      }
    }""",
        }

    def key(self) -> str:
        return hashlib.sha256(self.source().encode()).hexdigest()[:20]



class NewtonModuleSource(OCModuleSource):
    """Newton / iLQR direction for the optimal-control problem itself (the batched ``ocSolver``).

    The exact Newton step of  min_U J(U)  at a dynamically consistent (X, U, lambda) solves an LQ problem
    with the same structure as the PDP auxiliary system with ONE column: E = 0, Hxe = 0, hxe = 0 and the
    "Hue" column replaced by the gradient dH/du.  So the fused aux-LQR kernel is reused unchanged; its
    forward pass returns (dx_t, du_t).  The module's parameter vector is [auxvar, s, mu]:
    ``s`` scales the costate inside the Hessians (s = 1: exact Newton / DDP Hessians, s = 0: Gauss-Newton =
    iLQR, always positive definite for convex costs) and ``mu`` is a Levenberg shift added to Huu."""

    def _customise(self):
        n, m = self.n, self.m
        s_, mu = SX.sym("newton_s"), SX.sym("newton_mu")
        lam_scaled = self.lam * s_
        def scaled(M):
            return S.substitute(M, self.lam, lam_scaled)
        self.ddHxx = scaled(self.ddHxx)
        self.ddHxu = scaled(self.ddHxu)
        self.ddHux = scaled(self.ddHux)
        self.ddHuu = scaled(self.ddHuu) + mu * SX.eye(m)
        self.dfe = SX.zeros(n, 1)
        self.ddHxe = SX.zeros(n, 1)
        self.ddHue = self.dHu
        self.ddhxe = SX.zeros(n, 1)
        self.th = S.vertcat(self.th, s_, mu)
        self.nth = self.th.numel()
        self.r = 1


class FunctionModuleSource:
    """Any symbolic ``Function`` as a batched CUDA kernel (one thread per sample).

    Used for the legacy ``getAuxSys`` return values of ControlPlanning / SysID (reference
    PDP/PDP.py:788-811, 1225-1239) and for evaluating user-visible ``*_fn`` objects on device.
    Inputs ``in_k[B or 1, numel_k]`` (stride 0 = shared), outputs ``out_k[B, rows*cols]`` row-major."""

    kind_id = 5
    MAX_IN, MAX_OUT = 6, 12

    def __init__(self, fn: "S.Function"):
        self.fn = fn
        if fn.n_in() > self.MAX_IN or fn.n_out() > self.MAX_OUT:
            raise ValueError("FunctionModuleSource: too many inputs / outputs")

    def source(self) -> str:
        fn = self.fn
        ins, outs = fn.sx_in(), fn.sx_out()
        leaf, loads = {}, []
        used = {n.uid for o in outs for n in S.topo_order(o.elements()) if n.op == "sym"}
        for k, m in enumerate(ins):
            for q, e in enumerate(m.elements()):  # column-major element order of the input matrix
                if e.uid in used:
                    loads.append("  const double i%d_%d = p.in[%d][(size_t)b * p.in_stride[%d] + %d];" % (k, q, k, k, q))
                    leaf[e.uid] = "i%d_%d" % (k, q)
        flat, where = [], []
        for k, o in enumerate(outs):
            rows, cols = o.shape
            for i in range(rows):
                for j in range(cols):
                    flat.append(o.at(i, j))
                    where.append((k, i * cols + j, rows * cols))
        lines, names = S.emit_c(flat, leaf)
        stores = ["  p.out[%d][(size_t)b * %d + %d] = %s;" % (k, tot, off, nm) for (k, off, tot), nm in zip(where, names)]
        return "\n".join([
            "// GENERATED by pontryagin_differentiable_programming_b200/codegen.py (FunctionModuleSource)",
            "#include <cuda_runtime.h>", "#include <math.h>",
            "#define PDP_KIND 5", "#define PDP_NIN %d" % len(ins), "#define PDP_NOUT %d" % len(outs),
            "struct pdp_fn_args { const double* in[%d]; int in_stride[%d]; double* out[%d]; };" % (self.MAX_IN, self.MAX_IN, self.MAX_OUT),
            'extern "C" __global__ void __launch_bounds__(128) pdp_k_fn(int B, pdp_fn_args p) {',
            "  const int b = blockIdx.x * blockDim.x + threadIdx.x;", "  if (b >= B) return;"] + loads + lines + stores + [
            "}",
            'extern "C" void pdpmod_info(int* out) { out[0] = PDP_KIND; out[1] = PDP_NIN; out[2] = PDP_NOUT; for (int i = 3; i < 11; ++i) out[i] = 0; }',
            'extern "C" int pdpmod_fn(int B, const double* const* ins, const int* strides, double* const* outs, cudaStream_t st) {',
            "  if (B <= 0) return 0;", "  pdp_fn_args p;",
            "  for (int i = 0; i < PDP_NIN; ++i) { p.in[i] = ins[i]; p.in_stride[i] = strides[i]; }",
            "  for (int i = 0; i < PDP_NOUT; ++i) p.out[i] = outs[i];",
            "  pdp_k_fn<<<(B + 127) / 128, 128, 0, st>>>(B, p);", "  return (int)cudaGetLastError();", "}", ""])

    def key(self) -> str:
        return hashlib.sha256(self.source().encode()).hexdigest()[:20]


class LQRModuleSource(OCModuleSource):
    """Generic time-varying matrix LQR of size (n, m, r) whose auxiliary matrices are READ from HBM
    (the drop-in ``LQR.lqrSolver``, reference PDP/PDP.py:446-615, for user-supplied matrices).

    Same kernel as the fused OC path with every entry treated as a structural non-zero and the
    per-chunk "evaluation" replaced by a cooperative gather from the dense per-step record
    ``[F|G|E|Hxx|Hxu|Hxe|Hux|Huu|Hue]`` (the layout ``pdp_aux_eval`` writes).  ``Hux`` is carried but
    unused, exactly like the reference (PDP.py:569,572,598 use transpose(Hxu))."""

    kind_id = 4
    fwd_smem_budget = 40 * 1024

    def __init__(self, n: int, m: int, r: int, chunk: int = 2, warps_per_block: int = 4):
        self.n, self.m, self.r = int(n), int(m), int(r)
        self.ns = self.n + self.m + self.r
        self.nth = 0
        self.chunk, self.wpb = int(chunk), int(warps_per_block)
        self.min_blocks, self.wpbf, self.min_blocks_f, self.keep_fg = 1, int(warps_per_block), 1, False
        n, m, r, ns = self.n, self.m, self.r, self.ns
        nm = n + m
        oF, oG, oE = 0, n * n, n * n + n * m
        oHxx = oE + n * r
        oHxu = oHxx + n * n
        oHxe = oHxu + n * m
        oHux = oHxe + n * r
        oHuu = oHux + m * n
        oHue = oHuu + m * m
        self.ndense = oHue + m * r
        src: List[int] = []
        seen: Dict[int, int] = {}

        def slot(off):
            if off not in seen:
                seen[off] = len(src)
                src.append(off)
            return ("v", seen[off], 1.0)

        self.S_ent = [[None] * ns for _ in range(n)]
        for k in range(n):
            for j in range(ns):
                if j < n:
                    self.S_ent[k][j] = slot(oF + k * n + j)
                elif j < nm:
                    self.S_ent[k][j] = slot(oG + k * m + (j - n))
                else:
                    self.S_ent[k][j] = slot(oE + k * r + (j - nm))
        self.nvar_s = len(src)
        self.H_idx = [[0] * nm for _ in range(ns)]
        for j in range(ns):
            for l in range(nm):
                if j < n and l < n:
                    off = oHxx + j * n + l
                elif j < n:
                    off = oHxu + j * m + (l - n)
                elif j < nm and l < n:
                    off = oHxu + l * m + (j - n)          # transpose(Hxu)
                elif j < nm:
                    off = oHuu + (j - n) * m + (l - n)
                elif l < n:
                    off = oHxe + l * r + (j - nm)
                else:
                    off = oHue + (l - n) * r + (j - nm)
                self.H_idx[j][l] = slot(off)[1]
        self.zero_slot = 0                                 # idle lanes read any valid slot (result unused)
        self.src_off = src
        self.nvar = len(src)
        self.auxld = _pad_ld(self.nvar)
        self.ldz, self.ldk = _odd(n), _even(n)

    def _device_functions(self) -> str:
        return ""

    def _param_recips(self):
        return []

    def _extra_tables(self):
        return ["__device__ const int pdp_slot_src[%d] = {%s};" % (len(self.src_off), ", ".join(map(str, self.src_off)))]

    def _kernel_text(self):
        info = _K_LAUNCH_COMMON[:_K_LAUNCH_COMMON.index('extern "C" int pdpmod_rollout_costate')]
        return _K_PRELUDE + _K_TMA_PRIMS + _K_AUX_LQR + info + _K_LAUNCH_LQR

    def _eval_macros(self):
        gather = """    {
      const int nst = (tc + PDP_CH < H ? PDP_CH : H - tc);
      for (int idx = lane; idx < nst * %(NV)s; idx += 32) {
        const int l = idx / %(NV)s, e = idx - l * %(NV)s;
        auxc[l * PDP_AUXLD + e] = auxrec[((size_t)b * H + tc + l) * PDP_NDENSE + pdp_slot_src[e]];
      }
    }"""
        return {
            "@@EVAL_TERM@@": "  for (int i = lane; i < PDP_N * PDP_N + PDP_N * PDP_R; i += 32) TB[i] = termrec[(size_t)b * (PDP_N * PDP_N + PDP_N * PDP_R) + i];",
            "@@EVAL_AUX_CHUNK@@": gather % {"NV": "PDP_NVAR"},
            "@@EVAL_DYN@@": "",
            "@@PREFETCH_AUX_CHUNK@@": "",
            "@@PREFETCH_DYN_CHUNK@@": "",
            "@@PREFETCH_DYN_CHUNK_L1@@": "",
            "@@EVAL_DYN_COOP@@": """    {
      const int nst = (tc + PDP_CHF < H ? PDP_CHF : H - tc);
      for (int idx = lane; idx < PDP_FG * nst * PDP_NVAR_S; idx += 32) {
        const int gg = idx / (nst * PDP_NVAR_S), rem = idx - gg * (nst * PDP_NVAR_S);
        const int l = rem / PDP_NVAR_S, e = rem - l * PDP_NVAR_S;
        const int bb = (b0 + gg < B) ? b0 + gg : B - 1;
        wbase[gg * PDP_FTS + l * PDP_FLD + e] = auxrec[((size_t)bb * H + tc + l) * PDP_NDENSE + pdp_slot_src[e]];
      }
    }""",
        }


from .kernel_templates import (  # noqa: E402
    K_AUX_LQR_BWD as _K_AUX_LQR_BWD, K_AUX_LQR_BWD2 as _K_AUX_LQR_BWD2, K_AUX_LQR_FWD as _K_AUX_LQR_FWD,
    K_AUX_LQR_HEAD as _K_AUX_LQR_HEAD, K_AUX_LQR as _K_AUX_LQR, K_LAUNCH_COMMON as _K_LAUNCH_COMMON, K_LAUNCH_LQR as _K_LAUNCH_LQR,
    K_PRELUDE as _K_PRELUDE, K_ROLLOUT_AUXEVAL as _K_ROLLOUT_AUXEVAL, K_TMA_PRIMS as _K_TMA_PRIMS)
