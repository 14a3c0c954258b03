"""Scalar-expression symbolic engine (the code-generation front-end of the B200 PDP engine).

The reference builds every dynamics / cost / Hamiltonian derivative with CasADi ``SX``
(reference ``PDP/PDP.py:96-119,222-270``; ``JinEnv/JinEnv.py``).  CasADi is not available in
this image, and in this design symbolic algebra is *only* a code generator (expressions ->
CUDA ``__device__`` straight-line code compiled into the sm_100a kernels), never something
evaluated on the hot path.  This module therefore implements the subset of the CasADi Python
API the reference uses, on top of a small hash-consed scalar expression DAG:

* ``SX`` dense matrices of scalar expressions (column-major semantics like CasADi),
* ``jacobian`` by reverse-mode differentiation on the DAG,
* ``Function`` objects callable with numbers (-> ``DM`` with ``.full()``) or with ``SX``
  arguments (-> symbolic substitution),
* emitters that turn a set of expressions into straight-line Python or C/CUDA source with
  common sub-expression sharing (``emit_python`` / ``emit_c``).

Nothing here is copied from CasADi or from the reference; it is a from-scratch engine.
"""
from __future__ import annotations

import hashlib
import math
import numbers
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

__all__ = [
    "SX", "MX", "DM", "Function", "jacobian", "gradient", "hessian", "vertcat", "horzcat", "vcat", "hcat",
    "mtimes", "dot", "transpose", "inv", "diag", "trace", "sin", "cos", "tan", "tanh", "exp", "log",
    "sqrt", "fabs", "sumsqr", "norm_2", "substitute", "pi", "inf", "nlpsol", "emit_python", "emit_c",
    "symvar", "is_constant",
]

pi = math.pi
inf = math.inf

# ---------------------------------------------------------------------------------------------
# Expression DAG
# ---------------------------------------------------------------------------------------------


class Node:
    """One vertex of the expression DAG (immutable, hash-consed)."""

    __slots__ = ("op", "args", "val", "uid", "skey", "height")

    def __init__(self, op, args, val, uid, skey):
        self.op = op
        self.args = args
        self.val = val
        self.uid = uid
        self.skey = skey          # structural key: depends on the expression's content only, never on creation order
        self.height = 1 + max(a.height for a in args) if args else 0

    def __repr__(self):
        return _fmt(self)


_TABLE: Dict[tuple, Node] = {}
_COUNTER = [0]
_SYM_COUNTER = [0]



def _skey(op, args, val) -> int:
    """64-bit content hash of a node: operator, the operands' keys, and the constant value or the symbol NAME (not its
    creation index).  Stable across processes (no use of Python's randomised ``hash``)."""
    if op == "sym":
        text = "sym:%s" % (val[0],)
    elif op == "const":
        text = "const:%r" % (val,)
    else:
        text = "%s:%s:%r" % (op, ",".join("%x" % a.skey for a in args), val)
    return int.from_bytes(hashlib.blake2b(text.encode(), digest_size=8).digest(), "big")


def _mk(op, args=(), val=None):
    key = (op, tuple(a.uid for a in args), val)
    node = _TABLE.get(key)
    if node is None:
        _COUNTER[0] += 1
        node = Node(op, tuple(args), val, _COUNTER[0], _skey(op, args, val))
        _TABLE[key] = node
    return node


def const(v) -> Node:
    v = float(v)
    if v == 0.0:
        v = 0.0  # fold -0.0
    return _mk("const", (), v)


ZERO = const(0.0)
ONE = const(1.0)
MINUS_ONE = const(-1.0)
TWO = const(2.0)


def sym(name: str) -> Node:
    # every call creates a distinct symbol even if the name repeats (CasADi semantics)
    _SYM_COUNTER[0] += 1
    return _mk("sym", (), (name, _SYM_COUNTER[0]))


def _isc(n: Node) -> bool:
    return n.op == "const"


def _swap(a: Node, b: Node) -> bool:
    """Canonical operand order of commutative ops: constants first, then the shallower operand (height in the DAG: it is
    ready earlier, which is also what creation order tended to give), then the STRUCTURAL key; creation order only breaks
    ties between distinct symbols of the same name.  Measured on B200 (profiles/r2f_order_ab.txt): identical kernel times
    to creation order; ordering by the content hash alone cost the latency-bound C2 / C4 kernels 7-14 %.  Ordering by creation index made the generated source -- and with it
    the module cache key -- depend on what else had been built from the same symbols earlier in the process (e.g.
    ``OCSys.diffPMP()`` before the first sweep): same mathematics, different operand order, a needless recompile."""
    ca, cb = a.op == "const", b.op == "const"
    if ca != cb:
        return cb
    return (a.height, a.skey, a.uid) > (b.height, b.skey, b.uid)


def add(a: Node, b: Node) -> Node:
    if _isc(a) and _isc(b):
        return const(a.val + b.val)
    if a is ZERO:
        return b
    if b is ZERO:
        return a
    if b.op == "neg":
        return sub(a, b.args[0])
    if a.op == "neg":
        return sub(b, a.args[0])
    if _swap(a, b):
        a, b = b, a
    return _mk("add", (a, b))


def sub(a: Node, b: Node) -> Node:
    if _isc(a) and _isc(b):
        return const(a.val - b.val)
    if b is ZERO:
        return a
    if a is ZERO:
        return neg(b)
    if a is b:
        return ZERO
    if b.op == "neg":
        return add(a, b.args[0])
    return _mk("sub", (a, b))


def neg(a: Node) -> Node:
    if _isc(a):
        return const(-a.val)
    if a.op == "neg":
        return a.args[0]
    if a.op == "sub":
        return sub(a.args[1], a.args[0])
    return _mk("neg", (a,))


def mul(a: Node, b: Node) -> Node:
    if _isc(a) and _isc(b):
        return const(a.val * b.val)
    if a is ZERO or b is ZERO:
        return ZERO
    if a is ONE:
        return b
    if b is ONE:
        return a
    if a is MINUS_ONE:
        return neg(b)
    if b is MINUS_ONE:
        return neg(a)
    if a is b:
        return sq(a)
    if a.op == "neg" and b.op == "neg":
        return mul(a.args[0], b.args[0])
    if a.op == "neg":
        return neg(mul(a.args[0], b))
    if b.op == "neg":
        return neg(mul(a, b.args[0]))
    if _swap(a, b):
        a, b = b, a
    return _mk("mul", (a, b))


def div(a: Node, b: Node) -> Node:
    if _isc(a) and _isc(b):
        return const(a.val / b.val)
    if a is ZERO:
        return ZERO
    if b is ONE:
        return a
    if b is MINUS_ONE:
        return neg(a)
    if a is b:
        return ONE
    if a.op == "neg":
        return neg(div(a.args[0], b))
    if b.op == "neg":
        return neg(div(a, b.args[0]))
    return _mk("div", (a, b))


def sq(a: Node) -> Node:
    if _isc(a):
        return const(a.val * a.val)
    if a.op == "neg":
        return sq(a.args[0])
    return _mk("sq", (a,))


def power(a: Node, b: Node) -> Node:
    if _isc(b):
        e = b.val
        if e == 0.0:
            return ONE
        if e == 1.0:
            return a
        if e == 2.0:
            return sq(a)
        if e == -1.0:
            return div(ONE, a)
        if e == 0.5:
            return unary("sqrt", a)
        if _isc(a):
            return const(a.val ** e)
    return _mk("pow", (a, b))


_PYFUN = {"sin": math.sin, "cos": math.cos, "tan": math.tan, "tanh": math.tanh, "exp": math.exp,
          "log": math.log, "sqrt": math.sqrt, "fabs": abs}


def unary(op: str, a: Node) -> Node:
    if op == "neg":
        return neg(a)
    if op == "sq":
        return sq(a)
    if _isc(a):
        return const(_PYFUN[op](a.val))
    if op in ("cos", "fabs") and a.op == "neg":
        return unary(op, a.args[0])
    if op in ("sin", "tan", "tanh") and a.op == "neg":
        return neg(unary(op, a.args[0]))
    return _mk(op, (a,))


def _fmt(n: Node, depth=0) -> str:
    if n.op == "const":
        return repr(n.val)
    if n.op == "sym":
        return n.val[0]
    if depth > 6:
        return "..."
    a = [_fmt(x, depth + 1) for x in n.args]
    if n.op == "add":
        return "(%s+%s)" % tuple(a)
    if n.op == "sub":
        return "(%s-%s)" % tuple(a)
    if n.op == "mul":
        return "(%s*%s)" % tuple(a)
    if n.op == "div":
        return "(%s/%s)" % tuple(a)
    if n.op == "neg":
        return "(-%s)" % a[0]
    if n.op == "pow":
        return "pow(%s,%s)" % tuple(a)
    return "%s(%s)" % (n.op, a[0])


def topo_order(outputs: Iterable[Node]) -> List[Node]:
    """Nodes reachable from ``outputs`` in dependency order (iterative DFS)."""
    seen = set()
    order: List[Node] = []
    for root in outputs:
        if root.uid in seen:
            continue
        stack = [(root, 0)]
        while stack:
            node, i = stack.pop()
            if i == 0 and node.uid in seen:
                continue
            if i < len(node.args):
                stack.append((node, i + 1))
                child = node.args[i]
                if child.uid not in seen:
                    stack.append((child, 0))
            else:
                if node.uid not in seen:
                    seen.add(node.uid)
                    order.append(node)
    return order


def _partials(n: Node) -> Tuple[Node, ...]:
    """d n / d arg_k for each argument, as expressions."""
    op = n.op
    a = n.args
    if op == "add":
        return (ONE, ONE)
    if op == "sub":
        return (ONE, MINUS_ONE)
    if op == "mul":
        return (a[1], a[0])
    if op == "div":
        # d(a/b) = 1/b, -(a/b)/b
        return (div(ONE, a[1]), neg(div(n, a[1])))
    if op == "neg":
        return (MINUS_ONE,)
    if op == "sq":
        return (mul(TWO, a[0]),)
    if op == "sin":
        return (unary("cos", a[0]),)
    if op == "cos":
        return (neg(unary("sin", a[0])),)
    if op == "tan":
        return (add(ONE, sq(n)),)
    if op == "tanh":
        return (sub(ONE, sq(n)),)
    if op == "exp":
        return (n,)
    if op == "log":
        return (div(ONE, a[0]),)
    if op == "sqrt":
        return (div(const(0.5), n),)
    if op == "fabs":
        return (div(a[0], n),)
    if op == "pow":
        x, y = a
        dx = mul(y, power(x, sub(y, ONE)))
        if _isc(y):
            return (dx, ZERO)
        return (dx, mul(n, unary("log", x)))
    raise ValueError("no derivative rule for op %r" % op)


def reverse_gradient(out: Node, wrt: Sequence[Node]) -> List[Node]:
    """Symbolic gradient of scalar node ``out`` w.r.t. the symbol nodes in ``wrt``."""
    order = topo_order([out])
    adj: Dict[int, Node] = {out.uid: ONE}
    for n in reversed(order):
        bar = adj.get(n.uid)
        if bar is None or bar is ZERO or not n.args:
            continue
        for arg, p in zip(n.args, _partials(n)):
            if arg.op == "const":
                continue
            contrib = mul(bar, p)
            if contrib is ZERO:
                continue
            prev = adj.get(arg.uid)
            adj[arg.uid] = contrib if prev is None else add(prev, contrib)
    return [adj.get(w.uid, ZERO) for w in wrt]


def depends_on(out: Node, wrt_ids: set) -> bool:
    for n in topo_order([out]):
        if n.uid in wrt_ids:
            return True
    return False


def substitute_nodes(outs: Sequence[Node], mapping: Dict[int, Node]) -> List[Node]:
    """Rebuild ``outs`` with symbol uid -> replacement node."""
    cache: Dict[int, Node] = {}
    for n in topo_order(outs):
        if n.uid in mapping:
            cache[n.uid] = mapping[n.uid]
        elif not n.args:
            cache[n.uid] = n
        else:
            a = [cache[x.uid] for x in n.args]
            op = n.op
            if op == "add":
                r = add(a[0], a[1])
            elif op == "sub":
                r = sub(a[0], a[1])
            elif op == "mul":
                r = mul(a[0], a[1])
            elif op == "div":
                r = div(a[0], a[1])
            elif op == "pow":
                r = power(a[0], a[1])
            else:
                r = unary(op, a[0])
            cache[n.uid] = r
    return [cache[o.uid] for o in outs]


# ---------------------------------------------------------------------------------------------
# Source emitters (shared by Function evaluation and by the CUDA code generator)
# ---------------------------------------------------------------------------------------------

def _c_literal(v: float) -> str:
    if v == math.inf:
        return "INFINITY"
    if v == -math.inf:
        return "(-INFINITY)"
    r = repr(float(v))
    if "e" not in r and "." not in r and "n" not in r:
        r += ".0"
    return r


def param_divisors(outputs: Iterable[Node], param_uids) -> List[Node]:
    """Divisors of ``outputs``' divisions that depend on nothing but the symbols ``param_uids`` (and constants), each once,
    in dependency order.  They are loop-invariant along a trajectory (parameters such as masses and inertias), so a
    kernel computes their reciprocals ONCE per trajectory and the generated per-step code multiplies (``emit_c(...,
    recip=...)``) instead of running a ~30-instruction FP64 division per occurrence and time step."""
    param_uids = set(param_uids)
    order = topo_order(outputs)
    only_params: Dict[int, bool] = {}
    for n in order:
        if n.op == "const":
            only_params[n.uid] = True
        elif n.op == "sym":
            only_params[n.uid] = n.uid in param_uids
        else:
            only_params[n.uid] = all(only_params[a.uid] for a in n.args)
    out, seen = [], set()
    for n in order:
        if n.op == "div":
            d = n.args[1]
            if d.op != "const" and only_params[d.uid] and d.uid not in seen:
                seen.add(d.uid)
                out.append(d)
    return out


def emit_c(outputs: Sequence[Node], leaf_names: Dict[int, str], prefix: str = "w", real: str = "double",
           indent: str = "  ", recip: Dict[int, str] = None) -> Tuple[List[str], List[str]]:
    """Straight-line C for ``outputs``.

    ``leaf_names`` maps symbol uid -> C expression.  Returns ``(lines, names)`` where ``names[i]``
    is the C expression holding ``outputs[i]`` after ``lines`` have run.  Shared sub-expressions
    are computed once; constants and leaves are inlined.  ``recip`` maps the uid of a divisor node to a C
    expression holding its precomputed reciprocal: divisions by it become multiplications.
    """
    order = topo_order(outputs)
    uses: Dict[int, int] = {}
    for n in order:
        for a in n.args:
            uses[a.uid] = uses.get(a.uid, 0) + 1
    for o in outputs:
        uses[o.uid] = uses.get(o.uid, 0) + 2  # outputs always materialised
    name: Dict[int, str] = {}
    lines: List[str] = []
    k = 0
    for n in order:
        if n.op == "const":
            name[n.uid] = _c_literal(n.val) if n.val >= 0 else "(%s)" % _c_literal(n.val)
            continue
        if n.op == "sym":
            if n.uid not in leaf_names:
                raise KeyError("free symbol %s is not an input of the generated function" % n.val[0])
            name[n.uid] = leaf_names[n.uid]
            continue
        a = [name[x.uid] for x in n.args]
        op = n.op
        if op == "add":
            e = "%s + %s" % (a[0], a[1])
        elif op == "sub":
            e = "%s - %s" % (a[0], a[1])
        elif op == "mul":
            e = "%s * %s" % (a[0], a[1])
        elif op == "div":
            d = n.args[1]
            if recip and d.uid in recip:
                e = "%s * %s" % (a[0], recip[d.uid])
            elif d.op == "const" and d.val != 0.0 and math.isfinite(1.0 / d.val):
                # division by a compile-time constant -> multiplication by its reciprocal (an FP64 division is ~30
                # instructions on the GPU; the Lagrange-polynomial policies of ControlPlanning divide by pivot differences
                # 30 times per step).  At most one ulp from the quotient, far inside every parity tolerance.
                r_ = 1.0 / d.val
                e = "%s * %s" % (a[0], _c_literal(r_) if r_ >= 0 else "(%s)" % _c_literal(r_))
            else:
                e = "%s / %s" % (a[0], a[1])
        elif op == "neg":
            e = "-%s" % a[0]
        elif op == "sq":
            e = "%s * %s" % (a[0], a[0])
        elif op == "pow":
            e = "pow(%s, %s)" % (a[0], a[1])
        else:
            e = "%s(%s)" % (op, a[0])
        # single-use cheap nodes are inlined to keep the source compact
        if uses.get(n.uid, 0) <= 1 and op in ("neg",):
            name[n.uid] = "(%s)" % e
            continue
        v = "%s%d" % (prefix, k)
        k += 1
        lines.append("%sconst %s %s = %s;" % (indent, real, v, e))
        name[n.uid] = v
    return lines, [name[o.uid] for o in outputs]


def emit_python(outputs: Sequence[Node], leaf_names: Dict[int, str], prefix: str = "w") -> Tuple[List[str], List[str]]:
    """Same as :func:`emit_c` but Python/numpy source (works on floats and on ndarrays)."""
    order = topo_order(outputs)
    name: Dict[int, str] = {}
    lines: List[str] = []
    k = 0
    for n in order:
        if n.op == "const":
            name[n.uid] = "(%r)" % n.val
            continue
        if n.op == "sym":
            if n.uid not in leaf_names:
                raise KeyError("free symbol %s is not an input of the function" % n.val[0])
            name[n.uid] = leaf_names[n.uid]
            continue
        a = [name[x.uid] for x in n.args]
        op = n.op
        if op == "add":
            e = "%s + %s" % (a[0], a[1])
        elif op == "sub":
            e = "%s - %s" % (a[0], a[1])
        elif op == "mul":
            e = "%s * %s" % (a[0], a[1])
        elif op == "div":
            e = "%s / %s" % (a[0], a[1])
        elif op == "neg":
            e = "-%s" % a[0]
        elif op == "sq":
            e = "%s * %s" % (a[0], a[0])
        elif op == "pow":
            e = "%s ** %s" % (a[0], a[1])
        elif op == "fabs":
            e = "_np.abs(%s)" % a[0]
        else:
            e = "_np.%s(%s)" % (op, a[0])
        v = "%s%d" % (prefix, k)
        k += 1
        lines.append("%s = %s" % (v, e))
        name[n.uid] = v
    return lines, [name[o.uid] for o in outputs]


def count_flops(outputs: Sequence[Node]) -> int:
    return sum(1 for n in topo_order(outputs) if n.args)


# ---------------------------------------------------------------------------------------------
# Dense matrix of expressions with CasADi-like semantics
# ---------------------------------------------------------------------------------------------

def _to_node(x) -> Node:
    if isinstance(x, Node):
        return x
    if isinstance(x, SX):
        if x.numel() != 1:
            raise ValueError("expected a scalar expression, got shape %s" % (x.shape,))
        return x._e[0]
    if isinstance(x, DM):
        return const(float(x))
    if isinstance(x, np.ndarray):
        if x.size != 1:
            raise ValueError("expected a scalar")
        return const(float(x.reshape(-1)[0]))
    if isinstance(x, numbers.Real):
        return const(float(x))
    raise TypeError("cannot convert %r to an expression" % type(x))


class SX:
    """Dense matrix of scalar expressions, stored column-major (``_e[i + j*rows]``)."""

    __array_ufunc__ = None  # make ndarray <op> SX defer to SX.__r<op>__
    __array_priority__ = 1000

    __slots__ = ("_e", "_r", "_c")

    def __init__(self, *args):
        if len(args) == 0:
            self._e, self._r, self._c = [], 0, 0
        elif len(args) == 1:
            other = _as_sx(args[0])
            self._e, self._r, self._c = list(other._e), other._r, other._c
        elif len(args) == 2:  # SX(n, m): zeros
            r, c = int(args[0]), int(args[1])
            self._e, self._r, self._c = [ZERO] * (r * c), r, c
        else:
            raise TypeError("SX(): unsupported constructor arguments")

    # ---- construction helpers -------------------------------------------------------------
    @staticmethod
    def _make(elems, r, c):
        m = SX.__new__(SX)
        m._e, m._r, m._c = list(elems), r, c
        return m

    @staticmethod
    def sym(name, n=1, m=1):
        if isinstance(n, (tuple, list)):
            n, m = n
        n, m = int(n), int(m)
        if n * m == 1:
            return SX._make([sym(name)], 1, 1)
        return SX._make([sym("%s_%d" % (name, k)) for k in range(n * m)], n, m)

    @staticmethod
    def zeros(n=1, m=1):
        return SX._make([ZERO] * (int(n) * int(m)), int(n), int(m))

    @staticmethod
    def ones(n=1, m=1):
        return SX._make([ONE] * (int(n) * int(m)), int(n), int(m))

    @staticmethod
    def eye(n):
        n = int(n)
        return SX._make([ONE if i == j else ZERO for j in range(n) for i in range(n)], n, n)

    # ---- shape ----------------------------------------------------------------------------
    @property
    def shape(self):
        return (self._r, self._c)

    def size(self, axis=None):
        if axis is None:
            return (self._r, self._c)
        return self._r if axis == 0 else self._c

    def size1(self):
        return self._r

    def size2(self):
        return self._c

    def numel(self):
        return self._r * self._c

    def rows(self):
        return self._r

    def columns(self):
        return self._c

    def is_scalar(self):
        return self._r * self._c == 1

    def is_empty(self):
        return self._r * self._c == 0

    def nnz(self):
        return sum(1 for e in self._e if e is not ZERO)

    def elements(self) -> List[Node]:
        """Column-major list of the element expressions."""
        return list(self._e)

    def at(self, i, j=0) -> Node:
        return self._e[i + j * self._r]

    def is_constant(self):
        return all(_const_only(e) for e in self._e)

    @property
    def T(self):
        return transpose(self)

    def reshape(self, *shape):
        if len(shape) == 1:
            shape = shape[0]
        r, c = int(shape[0]), int(shape[1])
        n = self.numel()
        if r == -1:
            r = n // c
        if c == -1:
            c = n // r
        if r * c != n:
            raise ValueError("reshape size mismatch")
        return SX._make(self._e, r, c)  # column-major storage => plain reinterpretation

    def __len__(self):
        return self._r

    def __iter__(self):
        raise TypeError("SX is not iterable (CasADi semantics); index it explicitly")

    # ---- indexing -------------------------------------------------------------------------
    @staticmethod
    def _idx(k, n):
        if isinstance(k, slice):
            return list(range(*k.indices(n))), False
        if isinstance(k, (list, tuple, np.ndarray)):
            return [int(i) + (n if int(i) < 0 else 0) for i in k], False
        k = int(k)
        if k < 0:
            k += n
        if not 0 <= k < n:
            raise IndexError("index %d out of range for size %d" % (k, n))
        return [k], True

    def __getitem__(self, key):
        if isinstance(key, tuple):
            ri, _ = self._idx(key[0], self._r)
            ci, _ = self._idx(key[1], self._c)
            return SX._make([self._e[i + j * self._r] for j in ci for i in ri], len(ri), len(ci))
        li, scalar = self._idx(key, self.numel())
        if self._r == 1 and self._c > 1 and not scalar:
            return SX._make([self._e[i] for i in li], 1, len(li))
        return SX._make([self._e[i] for i in li], len(li), 1)

    def __setitem__(self, key, value):
        value = _as_sx(value)
        if isinstance(key, tuple):
            ri, _ = self._idx(key[0], self._r)
            ci, _ = self._idx(key[1], self._c)
            tgt = [i + j * self._r for j in ci for i in ri]
        else:
            tgt, _ = self._idx(key, self.numel())
        if value.numel() == 1:
            for t in tgt:
                self._e[t] = value._e[0]
        else:
            if value.numel() != len(tgt):
                raise ValueError("assignment size mismatch")
            for t, v in zip(tgt, value._e):
                self._e[t] = v

    # ---- arithmetic (element-wise with scalar broadcasting, like CasADi) ---------------------
    def _bin(self, other, fn, swap=False):
        try:
            other = _as_sx(other)
        except TypeError:
            return NotImplemented
        a, b = (other, self) if swap else (self, other)
        if a.numel() == 1 and b.numel() != 1:
            x = a._e[0]
            return SX._make([fn(x, y) for y in b._e], b._r, b._c)
        if b.numel() == 1 and a.numel() != 1:
            y = b._e[0]
            return SX._make([fn(x, y) for x in a._e], a._r, a._c)
        if a.shape != b.shape:
            if a.numel() == b.numel() and (1 in a.shape and 1 in b.shape):
                b = SX._make(b._e, a._r, a._c)  # row/column vector mix-up tolerated
            else:
                raise ValueError("dimension mismatch: %s vs %s" % (a.shape, b.shape))
        return SX._make([fn(x, y) for x, y in zip(a._e, b._e)], a._r, a._c)

    def __add__(self, o):
        return self._bin(o, add)

    def __radd__(self, o):
        return self._bin(o, add, True)

    def __sub__(self, o):
        return self._bin(o, sub)

    def __rsub__(self, o):
        return self._bin(o, sub, True)

    def __mul__(self, o):
        return self._bin(o, mul)

    def __rmul__(self, o):
        return self._bin(o, mul, True)

    def __truediv__(self, o):
        return self._bin(o, div)

    def __rtruediv__(self, o):
        return self._bin(o, div, True)

    def __pow__(self, o):
        return self._bin(o, power)

    def __rpow__(self, o):
        return self._bin(o, power, True)

    def __matmul__(self, o):
        return mtimes(self, o)

    def __rmatmul__(self, o):
        return mtimes(o, self)

    def __neg__(self):
        return SX._make([neg(x) for x in self._e], self._r, self._c)

    def __pos__(self):
        return self

    def __float__(self):
        if self.numel() != 1 or not _isc(self._e[0]):
            raise TypeError("only constant scalar SX can be converted to float")
        return self._e[0].val

    def __repr__(self):
        if self.numel() == 1:
            return "SX(%s)" % _fmt(self._e[0])
        if self.numel() <= 16:
            rows = [[_fmt(self._e[i + j * self._r]) for j in range(self._c)] for i in range(self._r)]
            return "SX(%s)" % rows
        return "SX(%dx%d)" % self.shape

    __hash__ = None


MX = SX  # the reference only uses MX inside ocSolver's NLP transcription; one expression type is enough


def _const_only(n: Node) -> bool:
    return all(x.op != "sym" for x in topo_order([n]))


def is_constant(x) -> bool:
    return all(_const_only(e) for e in _as_sx(x)._e)


def _as_sx(x) -> SX:
    if isinstance(x, SX):
        return x
    if isinstance(x, Node):
        return SX._make([x], 1, 1)
    if isinstance(x, DM):
        x = x.full()
    if isinstance(x, numbers.Real):
        return SX._make([const(x)], 1, 1)
    if isinstance(x, np.ndarray):
        if x.dtype == object:
            flat = [_to_node(v) for v in x.reshape(-1, order="F")]
        else:
            flat = None
        if x.ndim == 0:
            return SX._make([const(float(x))], 1, 1)
        if x.ndim == 1:
            r, c = x.shape[0], 1
        elif x.ndim == 2:
            r, c = x.shape
        else:
            raise TypeError("only 0/1/2-D arrays convert to SX")
        if flat is None:
            flat = [const(v) for v in np.asarray(x, dtype=np.float64).reshape(-1, order="F")]
        return SX._make(flat, r, c)
    if isinstance(x, (list, tuple)):
        if len(x) == 0:
            return SX._make([], 0, 1)
        if all(isinstance(v, (list, tuple)) for v in x):
            return _as_sx(np.array(x, dtype=np.float64))
        return SX._make([_to_node(v) for v in x], len(x), 1)
    raise TypeError("cannot convert %r to SX" % type(x))


# ---- free functions ---------------------------------------------------------------------------

def vertcat(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)) and not _looks_numeric_vector(args[0]):
        args = tuple(args[0])
    mats = [_as_sx(a) for a in args]
    mats = [m for m in mats if m.numel() > 0]
    if not mats:
        return SX._make([], 0, 1)
    c = mats[0]._c
    for m in mats:
        if m._c != c:
            raise ValueError("vertcat: column mismatch %s" % ([m.shape for m in mats],))
    r = sum(m._r for m in mats)
    elems = []
    for j in range(c):
        for m in mats:
            elems.extend(m._e[j * m._r:(j + 1) * m._r])
    return SX._make(elems, r, c)


def _looks_numeric_vector(x):
    return all(isinstance(v, numbers.Real) for v in x)


def horzcat(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)) and not _looks_numeric_vector(args[0]):
        args = tuple(args[0])
    mats = [_as_sx(a) for a in args]
    mats = [m for m in mats if m.numel() > 0]
    if not mats:
        return SX._make([], 1, 0)
    r = mats[0]._r
    for m in mats:
        if m._r != r:
            raise ValueError("horzcat: row mismatch %s" % ([m.shape for m in mats],))
    elems = []
    for m in mats:
        elems.extend(m._e)
    return SX._make(elems, r, sum(m._c for m in mats))


def vcat(items):
    return vertcat(*list(items))


def hcat(items):
    return horzcat(*list(items))


def transpose(a):
    a = _as_sx(a)
    return SX._make([a._e[i + j * a._r] for i in range(a._r) for j in range(a._c)], a._c, a._r)


def mtimes(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = tuple(args[0])
    out = _as_sx(args[0])
    for nxt in args[1:]:
        out = _mtimes2(out, _as_sx(nxt))
    return out


def _mtimes2(a: SX, b: SX) -> SX:
    if a.numel() == 1 or b.numel() == 1:
        return a * b
    if a._c != b._r:
        raise ValueError("mtimes: inner dimension mismatch %s x %s" % (a.shape, b.shape))
    r, c, kk = a._r, b._c, a._c
    elems = []
    for j in range(c):
        for i in range(r):
            acc = ZERO
            for k in range(kk):
                acc = add(acc, mul(a._e[i + k * r], b._e[k + j * kk]))
            elems.append(acc)
    return SX._make(elems, r, c)


def dot(a, b):
    a, b = _as_sx(a), _as_sx(b)
    if a.numel() != b.numel():
        raise ValueError("dot: size mismatch")
    acc = ZERO
    for x, y in zip(a._e, b._e):
        acc = add(acc, mul(x, y))
    return SX._make([acc], 1, 1)


def sumsqr(a):
    a = _as_sx(a)
    return dot(a, a)


def norm_2(a):
    return sqrt(sumsqr(a))


def trace(a):
    a = _as_sx(a)
    if a._r != a._c:
        raise ValueError("trace of a non-square matrix")
    acc = ZERO
    for i in range(a._r):
        acc = add(acc, a._e[i + i * a._r])
    return SX._make([acc], 1, 1)


def diag(a):
    a = _as_sx(a)
    if a._c == 1 or a._r == 1:
        n = a.numel()
        return SX._make([a._e[i] if i == j else ZERO for j in range(n) for i in range(n)], n, n)
    n = min(a._r, a._c)
    return SX._make([a._e[i + i * a._r] for i in range(n)], n, 1)


def inv(a):
    """Symbolic inverse by Gauss-Jordan elimination with structural pivot choice (small matrices)."""
    a = _as_sx(a)
    n = a._r
    if n != a._c:
        raise ValueError("inv of a non-square matrix")
    if n == 1:
        return SX._make([div(ONE, a._e[0])], 1, 1)
    if n == 2:
        p, q, r_, s = a._e[0], a._e[2], a._e[1], a._e[3]  # [[p q],[r s]]
        det = sub(mul(p, s), mul(q, r_))
        return SX._make([div(s, det), div(neg(r_), det), div(neg(q), det), div(p, det)], 2, 2)
    # augmented rows
    rows = [[a._e[i + j * n] for j in range(n)] + [ONE if i == j else ZERO for j in range(n)] for i in range(n)]
    for col in range(n):
        piv = next((i for i in range(col, n) if rows[i][col] is not ZERO), None)
        if piv is None:
            raise ValueError("inv: structurally singular matrix")
        rows[col], rows[piv] = rows[piv], rows[col]
        p = rows[col][col]
        rows[col] = [div(v, p) for v in rows[col]]
        for i in range(n):
            if i != col and rows[i][col] is not ZERO:
                f = rows[i][col]
                rows[i] = [sub(v, mul(f, w)) for v, w in zip(rows[i], rows[col])]
    return SX._make([rows[i][n + j] for j in range(n) for i in range(n)], n, n)


def _map_unary(op, a):
    if isinstance(a, numbers.Real):
        return _PYFUN[op](a)
    if isinstance(a, np.ndarray) and a.dtype != object:
        return getattr(np, "abs" if op == "fabs" else op)(a)
    a = _as_sx(a)
    return SX._make([unary(op, x) for x in a._e], a._r, a._c)


def sin(a):
    return _map_unary("sin", a)


def cos(a):
    return _map_unary("cos", a)


def tan(a):
    return _map_unary("tan", a)


def tanh(a):
    return _map_unary("tanh", a)


def exp(a):
    return _map_unary("exp", a)


def log(a):
    return _map_unary("log", a)


def sqrt(a):
    return _map_unary("sqrt", a)


def fabs(a):
    return _map_unary("fabs", a)


def symvar(x) -> List[Node]:
    return [n for n in topo_order(_as_sx(x)._e) if n.op == "sym"]


def jacobian(expr, wrt):
    """d vec(expr) / d vec(wrt) -> (numel(expr) x numel(wrt)), reverse mode per output row."""
    expr, wrt = _as_sx(expr), _as_sx(wrt)
    wnodes = wrt._e
    for w in wnodes:
        if w.op != "sym":
            raise ValueError("jacobian: second argument must be purely symbolic")
    p, q = expr.numel(), wrt.numel()
    cols = [[ZERO] * p for _ in range(q)]
    for i, e in enumerate(expr._e):
        g = reverse_gradient(e, wnodes)
        for j in range(q):
            cols[j][i] = g[j]
    elems = []
    for j in range(q):
        elems.extend(cols[j])
    return SX._make(elems, p, q)


def gradient(expr, wrt):
    return transpose(jacobian(expr, wrt))


def hessian(expr, wrt):
    g = gradient(expr, wrt)
    return jacobian(g, wrt), g


def substitute(expr, old, new):
    expr, old, new = _as_sx(expr), _as_sx(old), _as_sx(new)
    if old.numel() != new.numel():
        raise ValueError("substitute: size mismatch")
    mapping = {o.uid: nn for o, nn in zip(old._e, new._e)}
    return SX._make(substitute_nodes(expr._e, mapping), expr._r, expr._c)


# ---------------------------------------------------------------------------------------------
# Numeric matrix result type
# ---------------------------------------------------------------------------------------------

class DM:
    """Numeric dense matrix returned by :class:`Function` calls (subset of ``casadi.DM``)."""

    __array_priority__ = 900

    def __init__(self, a=0.0):
        if isinstance(a, DM):
            a = a._a
        a = np.array(a, dtype=np.float64)
        if a.ndim == 0:
            a = a.reshape(1, 1)
        elif a.ndim == 1:
            a = a.reshape(-1, 1)
        self._a = a

    def full(self):
        return self._a.copy()

    def toarray(self, simplify=False):
        return self._a.copy()

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    @property
    def shape(self):
        return self._a.shape

    def numel(self):
        return self._a.size

    def size1(self):
        return self._a.shape[0]

    def size2(self):
        return self._a.shape[1]

    @property
    def T(self):
        return DM(self._a.T)

    def __float__(self):
        if self._a.size != 1:
            raise TypeError("only 1x1 DM converts to float")
        return float(self._a.reshape(-1)[0])

    def __int__(self):
        return int(float(self))

    def __getitem__(self, k):
        if isinstance(k, tuple):
            return DM(np.atleast_2d(self._a[k]))
        return DM(self._a.reshape(-1, order="F")[k])

    def _bin(self, o, fn, swap=False):
        if isinstance(o, SX):
            return NotImplemented
        b = o._a if isinstance(o, DM) else np.asarray(o, dtype=np.float64)
        a = self._a
        if b.ndim == 1 and a.shape[1] == 1 and b.size == a.size:
            b = b.reshape(a.shape)
        return DM(fn(b, a) if swap else fn(a, b))

    def __add__(self, o):
        return self._bin(o, np.add)

    def __radd__(self, o):
        return self._bin(o, np.add, True)

    def __sub__(self, o):
        return self._bin(o, np.subtract)

    def __rsub__(self, o):
        return self._bin(o, np.subtract, True)

    def __mul__(self, o):
        return self._bin(o, np.multiply)

    def __rmul__(self, o):
        return self._bin(o, np.multiply, True)

    def __truediv__(self, o):
        return self._bin(o, np.divide)

    def __rtruediv__(self, o):
        return self._bin(o, np.divide, True)

    def __neg__(self):
        return DM(-self._a)

    def __repr__(self):
        return "DM(%s)" % (self._a.tolist(),)


# ---------------------------------------------------------------------------------------------
# Function
# ---------------------------------------------------------------------------------------------

class Function:
    """Callable built from symbolic inputs/outputs (subset of ``casadi.Function``).

    Numeric calls run generated straight-line Python (compiled once, lazily); calls with ``SX``
    arguments substitute symbolically.  ``sx_in``/``sx_out`` expose the defining expressions so
    the CUDA code generator can emit the same function as ``__device__`` code.
    """

    def __init__(self, name, inputs, outputs, *_ignored, **_ignored_kw):
        self._name = str(name)
        self._in = [_as_sx(i) for i in inputs]
        self._out = [_as_sx(o) for o in outputs]
        for i in self._in:
            for e in i._e:
                if e.op != "sym":
                    raise ValueError("Function %s: inputs must be purely symbolic" % name)
        self._compiled = None

    def name(self):
        return self._name

    def n_in(self):
        return len(self._in)

    def n_out(self):
        return len(self._out)

    def sx_in(self, k=None):
        return list(self._in) if k is None else self._in[k]

    def sx_out(self, k=None):
        return list(self._out) if k is None else self._out[k]

    def size_in(self, k):
        return self._in[k].shape

    def size_out(self, k):
        return self._out[k].shape

    def numel_in(self, k):
        return self._in[k].numel()

    def numel_out(self, k):
        return self._out[k].numel()

    def _compile(self):
        leaf = {}
        for a, m in enumerate(self._in):
            for k, e in enumerate(m._e):
                leaf[e.uid] = "a%d[%d]" % (a, k)
        outs = [e for o in self._out for e in o._e]
        lines, names = emit_python(outs, leaf)
        src = ["def _f(%s):" % ", ".join("a%d" % a for a in range(len(self._in)))]
        src += ["    " + ln for ln in lines]
        src.append("    return (%s)" % "".join(n + ", " for n in names))
        ns = {"_np": np}
        exec(compile("\n".join(src), "<Function %s>" % self._name, "exec"), ns)
        self._compiled = ns["_f"]

    def _flat_arg(self, k, a):
        want = self._in[k].numel()
        if isinstance(a, DM):
            a = a._a
        arr = np.asarray(a, dtype=np.float64)
        if arr.ndim == 2:
            flat = arr.reshape(-1, order="F")
        else:
            flat = arr.reshape(-1)
        if flat.size != want:
            if flat.size == 1:
                flat = np.full(want, flat[0])
            else:
                raise ValueError("Function %s: argument %d has %d elements, expected %d"
                                 % (self._name, k, flat.size, want))
        return flat

    def eval_flat(self, flats):
        """Evaluate with already-flattened (column-major) float inputs; entries may be ndarrays
        (batched evaluation broadcasts).  Returns the flat tuple of output element values."""
        if self._compiled is None:
            self._compile()
        return self._compiled(*flats)

    def __call__(self, *args, **kwargs):
        if kwargs:
            raise TypeError("Function %s: keyword calls are not supported" % self._name)
        if len(args) != len(self._in):
            raise TypeError("Function %s takes %d arguments (%d given)" % (self._name, len(self._in), len(args)))
        if any(isinstance(a, SX) for a in args):
            mapping = {}
            for m, a in zip(self._in, args):
                a = _as_sx(a)
                if a.numel() != m.numel():
                    if a.numel() == 1:
                        a = SX._make([a._e[0]] * m.numel(), m._r, m._c)
                    else:
                        raise ValueError("Function %s: symbolic argument size mismatch" % self._name)
                for s, v in zip(m._e, a._e):
                    mapping[s.uid] = v
            res = []
            for o in self._out:
                res.append(SX._make(substitute_nodes(o._e, mapping), o._r, o._c))
            return res[0] if len(res) == 1 else tuple(res)
        flats = [self._flat_arg(k, a) for k, a in enumerate(args)]
        vals = self.eval_flat(flats)
        res = []
        pos = 0
        for o in self._out:
            n = o.numel()
            block = np.array([float(v) for v in vals[pos:pos + n]], dtype=np.float64)
            pos += n
            res.append(DM(block.reshape((o._r, o._c), order="F")))
        return res[0] if len(res) == 1 else tuple(res)

    def __repr__(self):
        return "Function(%s: %s -> %s)" % (self._name, [i.shape for i in self._in], [o.shape for o in self._out])


def nlpsol(*_a, **_k):
    raise NotImplementedError(
        "nlpsol/IPOPT is not part of the B200 PDP engine: OCSys.ocSolver solves the optimal control "
        "problem with the batched CUDA Newton/iLQR solver instead (see DESIGN.md)")
