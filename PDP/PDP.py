"""Drop-in ``PDP`` module backed by the B200-native batched engine.

Keeps the class surface of the reference's ``PDP/PDP.py`` -- ``OCSys`` (:57), ``LQR`` (:334),
``ControlPlanning`` (:640), ``SysID`` (:1157) -- so ``Examples/`` scripts run unchanged, but every
numeric hot-path method executes as sm_100a CUDA kernels behind the C ABI of
``include/pdp_b200.h``.  Legacy single-trajectory calls are the B = 1 case of the batched kernels;
the additional ``*_batched`` methods take/return torch CUDA float64 tensors with a leading
trajectory dimension.  There is no CPU fallback: without a CUDA device the hot-path methods raise
``PDPBackendError``.

Quirks of the reference that are preserved on purpose ("traps"): ``dp`` is half
the gradient of the printed loss; ``LQR`` uses ``transpose(Hxu)`` wherever ``Hux`` would appear;
``LQR.lqrSolver`` without ``hxe`` fails; time-invariant matrices given as a bare ndarray / 1-element
list are broadcast over the horizon; ``getAuxSys`` calls ``diffPMP`` on demand.
"""
import numpy
import numpy as np  # noqa: F401  (the reference leaks ``np`` through ``from casadi import *``)

from pontryagin_differentiable_programming_b200.symbolic import (  # noqa: F401
    SX, MX, DM, Function, dot, jacobian, mtimes, tanh, transpose, vcat, vertcat)


# ------------------------------------------------------------------------------------------------ helpers
def _torch():
    import torch
    return torch


def _engine():
    from pontryagin_differentiable_programming_b200 import engine
    return engine


def _device():
    eng = _engine()
    eng.require_cuda()
    torch = _torch()
    return torch.device("cuda", torch.cuda.current_device())


def _flat(value, size, what):
    """list / ndarray / DM / scalar of any orientation -> float64 vector of length ``size``."""
    if isinstance(value, DM):
        value = value.full()
    arr = numpy.asarray(value, dtype=numpy.float64).reshape(-1)
    if arr.size == 1 and size != 1:
        arr = numpy.full(size, arr[0])
    assert arr.size == size, "%s has %d elements, expected %d" % (what, arr.size, size)
    return arr


def _dev_tensor(a, dev):
    torch = _torch()
    return torch.as_tensor(numpy.ascontiguousarray(a, dtype=numpy.float64), device=dev)


def _bounds(given, count, default):
    return given if len(given) == count else count * [default]


def _matrix_list(value, message, optional=False):
    """ndarray -> [ndarray]; list of ndarray -> itself; None allowed when ``optional``."""
    if value is None and optional:
        return None
    if type(value) is numpy.ndarray:
        return [value]
    if type(value[0]) is numpy.ndarray:
        return value
    assert False, message


# ================================================================================================== OCSys
class OCSys:
    """Optimal control system  x+ = f(x,u,auxvar),  J = sum c(x,u,auxvar) + h(x,auxvar)."""

    def __init__(self, project_name="my optimal control system"):
        self.project_name = project_name
        self._compiled = None

    # -------------------------------------------------------------------------------- definition
    def setAuxvarVariable(self, auxvar=None):
        if auxvar is None or auxvar.numel() == 0:
            auxvar = SX.sym('auxvar')
        self.auxvar = auxvar
        self.n_auxvar = self.auxvar.numel()
        self._compiled = None

    def setStateVariable(self, state, state_lb=[], state_ub=[]):
        self.state = state
        self.n_state = self.state.numel()
        self.state_lb = _bounds(state_lb, self.n_state, -1e20)
        self.state_ub = _bounds(state_ub, self.n_state, 1e20)
        self._compiled = None

    def setControlVariable(self, control, control_lb=[], control_ub=[]):
        self.control = control
        self.n_control = self.control.numel()
        self.control_lb = _bounds(control_lb, self.n_control, -1e20)
        self.control_ub = _bounds(control_ub, self.n_control, 1e20)
        self._compiled = None

    def _need_auxvar(self):
        if not hasattr(self, 'auxvar'):
            self.setAuxvarVariable()

    def setDyn(self, ode):
        self._need_auxvar()
        self.dyn = SX(ode)
        self.dyn_fn = Function('dynamics', [self.state, self.control, self.auxvar], [self.dyn])
        self._compiled = None

    def setPathCost(self, path_cost):
        self._need_auxvar()
        assert path_cost.numel() == 1, "path_cost must be a scalar function"
        self.path_cost = path_cost
        self.path_cost_fn = Function('path_cost', [self.state, self.control, self.auxvar], [self.path_cost])
        self._compiled = None

    def setFinalCost(self, final_cost):
        self._need_auxvar()
        assert final_cost.numel() == 1, "final_cost must be a scalar function"
        self.final_cost = final_cost
        self.final_cost_fn = Function('final_cost', [self.state, self.auxvar], [self.final_cost])
        self._compiled = None

    def _check_defined(self, cost_word="running cost"):
        assert hasattr(self, 'state'), "Define the state variable first!"
        assert hasattr(self, 'control'), "Define the control variable first!"
        assert hasattr(self, 'dyn'), "Define the system dynamics first!"
        assert hasattr(self, 'path_cost'), "Define the %s function first!" % cost_word
        assert hasattr(self, 'final_cost'), "Define the final cost function first!"

    # -------------------------------------------------------------------------------- differentiation
    def diffPMP(self):
        """Differentiate the Pontryagin conditions symbolically (reference PDP.py:222-270) and expose the
        same ``*_fn`` attributes.  The CUDA module holding these derivatives is compiled on first use."""
        self._check_defined("running cost/reward")
        x, u, e = self.state, self.control, self.auxvar
        self.costate = SX.sym('lambda', self.state.numel())
        lam = self.costate
        self.path_Hamil = self.path_cost + dot(self.dyn, lam)
        self.final_Hamil = self.final_cost
        xue, xule, xe = [x, u, e], [x, u, lam, e], [x, e]
        self.dfx, self.dfu, self.dfe = jacobian(self.dyn, x), jacobian(self.dyn, u), jacobian(self.dyn, e)
        self.dfx_fn, self.dfu_fn, self.dfe_fn = (Function(nm, xue, [ex]) for nm, ex in
                                                 (('dfx', self.dfx), ('dfu', self.dfu), ('dfe', self.dfe)))
        self.dHx = jacobian(self.path_Hamil, x).T
        self.dHu = jacobian(self.path_Hamil, u).T
        self.dHx_fn = Function('dHx', xule, [self.dHx])
        self.dHu_fn = Function('dHu', xule, [self.dHu])
        for nm, first, wrt in (('ddHxx', self.dHx, x), ('ddHxu', self.dHx, u), ('ddHxe', self.dHx, e),
                               ('ddHux', self.dHu, x), ('ddHuu', self.dHu, u), ('ddHue', self.dHu, e)):
            expr = jacobian(first, wrt)
            setattr(self, nm, expr)
            setattr(self, nm + '_fn', Function(nm, xule, [expr]))
        self.dhx = jacobian(self.final_Hamil, x).T
        self.dhx_fn = Function('dhx', xe, [self.dhx])
        self.ddhxx = jacobian(self.dhx, x)
        self.ddhxx_fn = Function('ddhxx', xe, [self.ddhxx])
        self.ddhxe = jacobian(self.dhx, e)
        self.ddhxe_fn = Function('ddhxe', xe, [self.ddhxe])

    def _system(self):
        """The compiled engine object (code generation + nvcc happen here, once per definition)."""
        if self._compiled is None:
            self._check_defined()
            self._compiled = _engine().OCSystem(self.state, self.control, self.auxvar, self.dyn,
                                                self.path_cost, self.final_cost)
        return self._compiled

    # -------------------------------------------------------------------------------- batched API (new)
    def rollout_batched(self, x0, auxvar_value, control_traj, want_costate=True, want_dHu=False):
        """x0[B,n], theta[B|1,r], U[B,H,m] (CUDA float64) -> dict X, Lam, cost[, dHu]."""
        return self._system().rollout_costate(x0, auxvar_value, control_traj, want_costate, want_dHu)

    def pdp_sweep_batched(self, x0, auxvar_value, control_traj, state_ref=None, control_ref=None, want_traj=True):
        """One PDP sweep per trajectory at given controls: rollout, costate, fused getAuxSys + lqrSolver.
        -> dict X[B,H+1,n], Lam[B,H,n], cost[B], dX[B,H+1,n,r], dU[B,H,m,r][, loss_dp[B,r+1]]."""
        return self._system().sweep(x0, auxvar_value, control_traj, Xref=state_ref, Uref=control_ref,
                                    want_traj=want_traj)

    def aux_lqr_batched(self, state_traj, control_traj, costate_traj, auxvar_value, **kw):
        return self._system().aux_lqr(state_traj, control_traj, costate_traj, auxvar_value, **kw)

    def _check_unbounded(self):
        """The reference hands state / control bounds to IPOPT as lbw / ubw (PDP.py:147-168).  The batched Newton / DDP
        solver that replaces IPOPT here is unconstrained: refuse loudly instead of returning an unconstrained optimum
        for a constrained problem.  (None of the reference's Examples scripts passes a bound.)"""
        for name in ("state_lb", "state_ub", "control_lb", "control_ub"):
            b = numpy.asarray(getattr(self, name, []), dtype=numpy.float64).ravel()
            if b.size and bool(numpy.any(numpy.abs(b) < 1e19)):
                raise NotImplementedError(
                    "OCSys.ocSolver: finite %s = %s was set, but the CUDA Newton/DDP solver of this engine handles "
                    "unconstrained problems only (the reference passes bounds to IPOPT, PDP.py:147-168); remove the "
                    "bound or enforce it through a penalty in the path cost" % (name, b.tolist()))

    def ocSolver_batched(self, ini_state, horizon, auxvar_value, control_init=None, n_starts=1, **opts):
        """Batched optimal-control solve (CUDA Newton / DDP, see ocsolver.py).  ``n_starts`` > 1 tries several
        seeded initial guesses per problem in the same batch and keeps the best stationary point."""
        from pontryagin_differentiable_programming_b200 import ocsolver
        self._check_unbounded()
        if control_init is None and n_starts > 1:
            return ocsolver.solve_multistart(self._system(), ini_state, int(horizon), auxvar_value, n_starts, **opts)
        return ocsolver.solve(self._system(), ini_state, int(horizon), auxvar_value, control_init, **opts)

    # -------------------------------------------------------------------------------- legacy API
    def ocSolver(self, ini_state, horizon, auxvar_value=1, print_level=0, costate_option=0, control_init=None,
                 n_starts=1):
        """Solve the OC problem for one initial state (reference PDP.py:121-220 used IPOPT; here the batched
        CUDA Newton/DDP solver).  Returns the same dict; ``costate_traj_opt[t] = lambda_{t+1}`` for either
        ``costate_option`` (the PMP recursion and the NLP multipliers coincide at a stationary point).
        Additive keyword arguments: ``control_init`` (H x m warm start) and ``n_starts`` (> 1: seeded multi-start in
        one batch, best stationary point wins; the default 1 is the reference's cold start from all-zero controls)."""
        self._check_defined()
        self._check_unbounded()
        dev = _device()
        x0 = _dev_tensor(_flat(ini_state, self.n_state, "ini_state")[None, :], dev)
        theta = _dev_tensor(_flat(auxvar_value, self.n_auxvar, "auxvar_value")[None, :], dev)
        if control_init is not None:
            control_init = _dev_tensor(numpy.asarray(control_init, dtype=numpy.float64).reshape(1, horizon, self.n_control), dev)
        sol = self.ocSolver_batched(x0, horizon, theta, control_init=control_init, n_starts=n_starts,
                                    verbose=print_level > 0)
        converged = bool(sol["converged"][0].item()) if "converged" in sol else True
        if not converged:
            # the reference never looks at IPOPT's return status (PDP.py:182-183); here a solve that stopped short of a
            # stationary point is at least announced: the auxiliary-system gradient is only meaningful at dH/du = 0
            import warnings
            warnings.warn("OCSys.ocSolver: the Newton/DDP solve did not reach the stationarity tolerance "
                          "(max |dH/du| = %.3g); try control_init= or n_starts=" % float(sol["grad_norm"][0].item()), RuntimeWarning)
        return {"state_traj_opt": sol["X"][0].cpu().numpy(),
                "control_traj_opt": sol["U"][0].cpu().numpy(),
                "costate_traj_opt": sol["Lam"][0].cpu().numpy(),
                'auxvar_value': auxvar_value,
                "time": numpy.arange(horizon + 1),
                "horizon": horizon,
                "cost": sol["cost"][0:1].cpu().numpy().reshape(1, 1),
                "converged": converged}        # additive key (not in the reference's dict)

    def getAuxSys(self, state_traj_opt, control_traj_opt, costate_traj_opt, auxvar_value=1):
        """Matrices of the auxiliary control system along a trajectory (reference PDP.py:272-314), evaluated
        by the ``pdp_k_aux_eval`` kernel and returned as lists of ndarrays like the reference."""
        if not all(hasattr(self, a) for a in ('dfx_fn', 'ddHxx_fn', 'ddhxe_fn')):
            self.diffPMP()
        dev = _device()
        U = numpy.asarray(control_traj_opt, dtype=numpy.float64).reshape(-1, self.n_control)
        H = U.shape[0]
        X = numpy.asarray(state_traj_opt, dtype=numpy.float64).reshape(H + 1, self.n_state)
        L = numpy.asarray(costate_traj_opt, dtype=numpy.float64).reshape(H, self.n_state)
        theta = _flat(auxvar_value, self.n_auxvar, "auxvar_value")
        aux = self._system().aux_eval(_dev_tensor(X[None], dev), _dev_tensor(U[None], dev), _dev_tensor(L[None], dev),
                                      _dev_tensor(theta[None], dev))
        out = {}
        for key in ("dynF", "dynG", "dynE", "Hxx", "Hxu", "Hxe", "Hux", "Huu", "Hue"):
            mats = aux[key][0].cpu().numpy()
            out[key] = [mats[t] for t in range(H)]
        out["hxx"] = [aux["hxx"][0].cpu().numpy()]
        out["hxe"] = [aux["hxe"][0].cpu().numpy()]
        return out


# ==================================================================================================== LQR
class LQR:
    """Time-varying matrix-valued LQR (reference PDP.py:334-615):
        X+ = F X + G U + E,   cost = tr(1/2 X'Hxx X + 1/2 U'Huu U + X'Hxu U + Hue'U + Hxe'X) + final."""

    def __init__(self, project_name="LQR system"):
        self.project_name = project_name

    def setDyn(self, dynF, dynG, dynE=None):
        self.dynF = _matrix_list(dynF, "Type of dynF matrix should be numpy.ndarray  or list of numpy.ndarray")
        self.n_state = numpy.size(self.dynF[0], 0)
        self.dynG = _matrix_list(dynG, "Type of dynG matrix should be numpy.ndarray  or list of numpy.ndarray")
        self.n_control = numpy.size(self.dynG[0], 1)
        self.dynE = _matrix_list(dynE, "Type of dynE matrix should be numpy.ndarray, list of numpy.ndarray, or None", True)
        self.n_batch = None if self.dynE is None else numpy.size(self.dynE[0], 1)

    def setPathCost(self, Hxx, Huu, Hxu=None, Hux=None, Hxe=None, Hue=None):
        msg = "Type of path cost %s matrix should be numpy.ndarray or list of numpy.ndarray, or None"
        self.Hxx = _matrix_list(Hxx, msg % "Hxx")
        self.Huu = _matrix_list(Huu, msg % "Huu")
        self.Hxu = _matrix_list(Hxu, msg % "Hxu", True)
        self.Hux = _matrix_list(Hux, msg % "Hux", True)
        self.Hxe = _matrix_list(Hxe, msg % "Hxe", True)
        self.Hue = _matrix_list(Hue, msg % "Hue", True)

    def setFinalCost(self, hxx, hxe=None):
        self.hxx = _matrix_list(hxx, "Type of final cost hxx matrix should be numpy.ndarray or list of numpy.ndarray")
        self.hxe = _matrix_list(hxe, "Type of final cost hxe matrix should be numpy.ndarray, list of numpy.ndarray, or None", True)

    def _over_horizon(self, mats, name, default_shape=None):
        H = self.horizon
        if mats is None:
            return numpy.zeros((H,) + default_shape)
        if len(mats) > 1 and len(mats) != H:
            assert False, "time-varying %s is not consistent with given horizon" % name
        if len(mats) == 1:
            return numpy.broadcast_to(numpy.asarray(mats[0], dtype=numpy.float64), (H,) + mats[0].shape)
        return numpy.stack([numpy.asarray(m, dtype=numpy.float64) for m in mats])

    def lqrSolver(self, ini_state, horizon):
        n_state = numpy.size(self.dynF[0], 1)
        if type(ini_state) is list:
            self.ini_x = numpy.array(ini_state, numpy.float64)
        elif type(ini_state) is numpy.ndarray:
            self.ini_x = ini_state
        else:
            assert False, "Initial state should be of numpy.ndarray type or list!"
        if self.ini_x.ndim == 2:
            self.n_batch = numpy.size(self.ini_x, 1)
        else:
            self.n_batch = 1
            self.ini_x = self.ini_x.reshape(n_state, -1)
        self.horizon = horizon
        if self.dynE is not None:
            assert self.n_batch == numpy.size(self.dynE[0], 1), "Number of data batch is not consistent with column of dynE"
        n, m, r, H = self.n_state, self.n_control, self.n_batch, horizon
        F = self._over_horizon(self.dynF, "dynF")
        G = self._over_horizon(self.dynG, "dynG")
        E = self._over_horizon(self.dynE, "dynE", (n, r))
        Hxx = self._over_horizon(self.Hxx, "Hxx")
        Huu = self._over_horizon(self.Huu, "Huu")
        Hxu = self._over_horizon(self.Hxu, "Hxu", (n, m))
        Hux = self._over_horizon(self.Hux, "Hux", (m, n))
        Hxe = self._over_horizon(self.Hxe, "Hxe", (n, r))
        Hue = self._over_horizon(self.Hue, "Hue", (m, r))
        hxx = numpy.asarray(self.hxx[0], dtype=numpy.float64)
        hxe = numpy.asarray(self.hxe[0], dtype=numpy.float64)  # like the reference (PDP.py:562) hxe=None fails here

        torch, eng = _torch(), _engine()
        dev = _device()
        rmax = 32 - n - m
        assert rmax >= 1, "LQR kernel supports n_state + n_control <= 31"
        Xs, Us = [], []
        for c0 in range(0, r, rmax):          # the columns of the matrix-valued state are independent
            cols = slice(c0, min(r, c0 + rmax))
            rc = cols.stop - cols.start
            rec = numpy.concatenate([a.reshape(H, -1) for a in
                                     (F, G, E[:, :, cols], Hxx, Hxu, Hxe[:, :, cols], Hux, Huu, Hue[:, :, cols])], axis=1)
            term = numpy.concatenate([hxx.reshape(-1), hxe[:, cols].reshape(-1)])
            solver = eng.DenseLQR.get(n, m, rc)
            Xa, Ua = solver.solve(_dev_tensor(rec[None], dev), _dev_tensor(term[None], dev),
                                  X0aux=_dev_tensor(numpy.ascontiguousarray(self.ini_x[:, cols])[None], dev))
            Xs.append(Xa[0])
            Us.append(Ua[0])
        Xa = torch.cat(Xs, dim=2)
        Ua = torch.cat(Us, dim=2)
        # costate of the auxiliary LQ problem by its own PMP recursion (equals P X + W at the optimum)
        tF, tHxx, tHxu, tHxe = (_dev_tensor(a, dev) for a in (F, Hxx, Hxu, Hxe))
        lam = _dev_tensor(hxx, dev) @ Xa[H] + _dev_tensor(hxe, dev)
        lams = [None] * H
        for t in range(H - 1, -1, -1):
            lams[t] = lam
            if t > 0:
                lam = tHxx[t] @ Xa[t] + tHxu[t] @ Ua[t] + tHxe[t] + tF[t].T @ lam
        Xn, Un = Xa.cpu().numpy(), Ua.cpu().numpy()
        return {'state_traj_opt': [Xn[t] for t in range(H + 1)],
                'control_traj_opt': [Un[t] for t in range(H)],
                'costate_traj_opt': [l.cpu().numpy() for l in lams],
                'time': [k for k in range(H + 1)]}


# ======================================================================================== ControlPlanning
class ControlPlanning:
    """Control / planning mode: x+ = f(x,u), J = sum c(x,u) + h(x), parameterised control policy."""

    def __init__(self, project_name="planner"):
        self.project_name = project_name
        self._cp = None
        self._oc = None

    def setStateVariable(self, state, state_lb=[], state_ub=[]):
        self.state = state
        self.n_state = self.state.numel()
        self.state_lb = _bounds(state_lb, self.n_state, -1e20)
        self.state_ub = _bounds(state_ub, self.n_state, 1e20)

    def setControlVariable(self, control, control_lb=[], control_ub=[]):
        self.control = control
        self.n_control = self.control.numel()
        self.control_lb = _bounds(control_lb, self.n_control, -1e20)
        self.control_ub = _bounds(control_ub, self.n_control, 1e20)

    def setDyn(self, ode):
        self.dyn = SX(ode)
        xu = [self.state, self.control]
        self.dyn_fn = Function('dynFun', xu, [self.dyn])
        self.dfx = jacobian(self.dyn, self.state)
        self.dfx_fn = Function('dfx', xu, [self.dfx])
        self.dfu = jacobian(self.dyn, self.control)
        self.dfu_fn = Function('dfu', xu, [self.dfu])
        self._cp = self._oc = None

    def setPathCost(self, path_cost):
        self.path_cost = path_cost
        xu = [self.state, self.control]
        self.path_cost_fn = Function('pathCost', xu, [self.path_cost])
        self.dcx_fn = Function('dcx', xu, [jacobian(self.path_cost, self.state)])
        self.dcu_fn = Function('dcx', xu, [jacobian(self.path_cost, self.control)])
        self._cp = self._oc = None

    def setFinalCost(self, final_cost):
        self.final_cost = final_cost
        self.final_cost_fn = Function('finalCost', [self.state], [self.final_cost])
        self.dhx_fn = Function('dhx', [self.state], [jacobian(self.final_cost, self.state)])
        self._cp = self._oc = None

    # -------------------------------------------------------------------------------- policies
    def _set_policy(self, policy, params):
        self.auxvar = params
        self.n_auxvar = self.auxvar.numel()
        self.policy = policy
        txe = [self.t, self.state, self.auxvar]
        self.policy_fn = Function('policy_fn', txe, [policy])
        self.dpolicy_dx_fn = Function('dpolicy_dx', txe, [jacobian(policy, self.state)])
        self.dpolicy_de_fn = Function('dpolicy_de', txe, [jacobian(policy, self.auxvar)])
        self._cp = None
        self._gpu_fns = {}

    def setPolyControl(self, pivots):
        """u(t) = Lagrange polynomial through control values at the pivot steps (reference PDP.py:699-725)."""
        from pontryagin_differentiable_programming_b200.systems import lagrange_policy
        self.t = SX.sym('t')
        policy, params = lagrange_policy(self.n_control, pivots, self.t)
        self._set_policy(policy, params)

    def setNeuralPolicy(self, hidden_layers):
        """u = tanh-MLP(x) with column-major packed weights (reference PDP.py:727-759)."""
        from pontryagin_differentiable_programming_b200.systems import neural_policy
        self.t = SX.sym('t')
        policy, params = neural_policy(self.state, self.n_control, hidden_layers)
        self._set_policy(policy, params)

    def init_step(self, horizon, n_poly=5):
        self.setPolyControl(numpy.linspace(0, horizon, n_poly + 1))

    def init_step_neural_policy(self, hidden_layers=None):
        if hidden_layers is None:
            hidden_layers = [self.n_state]
        self.setNeuralPolicy(hidden_layers)

    def _cp_system(self):
        if self._cp is None:
            self._cp = _engine().CPSystem(self.state, self.control, self.auxvar, self.dyn, self.policy, self.t,
                                          self.path_cost, self.final_cost)
        return self._cp

    # -------------------------------------------------------------------------------- batched API (new)
    def step_batched(self, ini_state, horizon, auxvar_value, want_traj=False, want_sens=False):
        """x0[B,n], theta[B|1,r] (CUDA float64) -> dict loss_dp[B,r+1] = (cost, dcost/dtheta) [+X,U,dX,dU]."""
        assert hasattr(self, 'policy_fn'), 'please set the control policy by running the init_step method first!'
        return self._cp_system().step(ini_state, horizon, auxvar_value, want_traj=want_traj, want_sens=want_sens)

    # -------------------------------------------------------------------------------- legacy API
    def integrateSys(self, ini_state, horizon, auxvar_value):
        assert hasattr(self, 'dyn_fn'), "Set the dynamics first!"
        assert hasattr(self, 'policy_fn'), "Set the control policy first, you may use [setPolicy_polyControl] "
        dev = _device()
        x0 = _dev_tensor(_flat(ini_state, self.n_state, "ini_state")[None], dev)
        th = _dev_tensor(_flat(auxvar_value, self.n_auxvar, "auxvar_value")[None], dev)
        out = self._cp_system().step(x0, horizon, th, want_traj=True)
        return {'state_traj': out["X"][0].cpu().numpy(), 'control_traj': out["U"][0].cpu().numpy(),
                'cost': float(out["loss_dp"][0, 0].item())}

    def _gpu_fn(self, name):
        if name not in self._gpu_fns:
            self._gpu_fns[name] = _engine().GpuFunction(getattr(self, name))
        return self._gpu_fns[name]

    def getAuxSys(self, state_traj, control_traj, auxvar_value):
        assert hasattr(self, 'dfx_fn'), "Set the dynamics equation first!"
        assert hasattr(self, 'dpolicy_de_fn'), "Set the policy first, you may want to use method [setPolicy_]"
        assert hasattr(self, 'dpolicy_dx_fn'), "Set the policy first, you may want to use method [setPolicy_]"
        dev = _device()
        U = numpy.asarray(control_traj, dtype=numpy.float64).reshape(-1, self.n_control)
        H = U.shape[0]
        X = _dev_tensor(numpy.asarray(state_traj, dtype=numpy.float64)[:H], dev)
        Ud = _dev_tensor(U, dev)
        th = _dev_tensor(_flat(auxvar_value, self.n_auxvar, "auxvar_value"), dev)
        tt = _dev_tensor(numpy.arange(H, dtype=numpy.float64)[:, None], dev)
        res = {"dynF": self._gpu_fn("dfx_fn")(X, Ud)[0], "dynG": self._gpu_fn("dfu_fn")(X, Ud)[0],
               "dUx": self._gpu_fn("dpolicy_dx_fn")(tt, X, th)[0], "dUe": self._gpu_fn("dpolicy_de_fn")(tt, X, th)[0]}
        return {k: [m for m in v.cpu().numpy()] for k, v in res.items()}

    def integrateAuxSys(self, dynF, dynG, dUx, dUe, ini_condition):
        if type(dynF) != list or type(dynG) != list or type(dUx) != list or type(dUe) != list:
            assert False, "The input dynF, dynE, dUx, and dUe should be list of numpy.array!"
        if len(dynG) != len(dynF) or len(dUe) != len(dUx) or len(dUe) != len(dynG):
            assert False, "The length of dynF, dynE, dUx, and dUe should be the same"
        if type(ini_condition) is not numpy.ndarray:
            assert False, "The initial condition should be numpy.array"
        Xn, Un = _forward_recursion(dynF, dynG, dUx, dUe, None, ini_condition)
        return {'state_traj': [Xn[t] for t in range(Xn.shape[0])], 'control_traj': [Un[t] for t in range(Un.shape[0])]}

    def step(self, ini_state, horizon, auxvar_value):
        assert hasattr(self, 'policy_fn'), 'please set the control policy by running the init_step method first!'
        dev = _device()
        x0 = _dev_tensor(_flat(ini_state, self.n_state, "ini_state")[None], dev)
        th = _dev_tensor(_flat(auxvar_value, self.n_auxvar, "auxvar_value")[None], dev)
        ldp = self._cp_system().step(x0, horizon, th)["loss_dp"][0].cpu().numpy()
        return float(ldp[0]), ldp[1:].copy()

    # -------------------------------------------------------------------------------- recovery-matrix mode
    # The reference builds one giant symbolic dJ/dU "recovery matrix" (PDP.py:1039-1079).  dJ/du_t equals the
    # adjoint expression c_u + f_u' lambda_{t+1}, so here recmat_* run the rollout/costate kernel and sum
    # dH/du over every warped interval (piecewise-constant controls on the time grid).
    def _oc_system(self):
        if self._oc is None:
            dummy = SX.sym('unused_auxvar')
            self._oc = _engine().OCSystem(self.state, self.control, dummy, self.dyn, self.path_cost, self.final_cost)
        return self._oc

    def recmat_init_step(self, horizon, time_grid=None):
        assert hasattr(self, 'dyn_fn'), 'Please set the dynamics first!'
        assert hasattr(self, 'path_cost_fn'), 'Please set the path cost first!'
        assert hasattr(self, 'final_cost_fn'), 'Please set the final cost first!'
        if time_grid is None:
            time_grid = numpy.linspace(0, 1, numpy.amin([horizon + 1, 11]))
        if numpy.isscalar(time_grid) and time_grid == -1:
            time_grid = numpy.linspace(0, horizon, horizon + 1)
        time_grid = numpy.asarray(time_grid, dtype=numpy.float64)
        self.time_grid = numpy.rint(horizon * time_grid / time_grid[-1]).astype(int)
        self.whorizon = len(self.time_grid) - 1
        self.n_auxvar = self.whorizon * self.n_control
        self.auxvar = SX.sym('U', self.n_auxvar)
        self._interval = numpy.repeat(numpy.arange(self.whorizon), numpy.diff(self.time_grid))

    def _interval_sums(self, per_step, n_blocks):
        """Sum the per-step gradient dH/du over every interval of the time grid (on the device)."""
        torch = _torch()
        idx = torch.as_tensor(self._interval, device=per_step.device, dtype=torch.long)
        return torch.zeros((n_blocks, self.n_control), dtype=per_step.dtype, device=per_step.device).index_add_(0, idx, per_step)

    def _expand_controls(self, auxvar_value):
        Uw = _flat(auxvar_value, self.n_auxvar, "auxvar_value").reshape(self.whorizon, self.n_control)
        return Uw[self._interval]

    def recmat_step_batched(self, ini_state, control_traj):
        """x0[B,n], U[B,H,m] (CUDA float64) -> cost[B], dJ/dU[B,H,m] by the costate (adjoint) kernel."""
        torch = _torch()
        th = torch.zeros((1, 1), dtype=torch.float64, device=ini_state.device)
        out = self._oc_system().rollout_costate(ini_state, th, control_traj, want_dHu=True)
        return out["cost"], out["dHu"], out["X"]

    def recmat_step(self, ini_state, horizon, auxvar_value):
        dev = _device()
        U = self._expand_controls(auxvar_value)
        x0 = _dev_tensor(_flat(ini_state, self.n_state, "ini_state")[None], dev)
        cost, dHu, _ = self.recmat_step_batched(x0, _dev_tensor(U[None], dev))
        dw = self._interval_sums(dHu[0], self.whorizon)
        return float(cost[0].item()), dw.cpu().numpy().reshape(-1)

    def recmat_unwarp(self, ini_state, horizon, auxvar_value):
        dev = _device()
        U = self._expand_controls(auxvar_value)
        x0 = _dev_tensor(_flat(ini_state, self.n_state, "ini_state")[None], dev)
        cost, _, X = self.recmat_step_batched(x0, _dev_tensor(U[None], dev))
        return {'state_traj': X[0].cpu().numpy(), 'control_traj': U, 'cost': cost.cpu().numpy().reshape(1)}

    # -------------------------------------------------------------------------------- time-warped policies
    # Reference PDP.py:882-1035 composes the dynamics / cost over every interval of a time grid symbolically and
    # parameterises the control by a Lagrange polynomial whose pivots are the integer warped steps 0..whorizon.
    # At an integer warped step wt the basis is the identity, so the applied control on interval wt is simply the
    # wt-th parameter block (the last block never acts).  The warped problem is therefore the original rollout
    # with piecewise-constant controls, and d(cost)/d(block wt) = sum of dH/du over the interval -- the same
    # adjoint kernel as recmat_*; no symbolic composition is needed.
    def warp_init_step(self, horizon, time_grid=None):
        assert hasattr(self, 'dyn_fn'), 'Please set the dynamics first!'
        assert hasattr(self, 'path_cost_fn'), 'Please set the path cost first!'
        assert hasattr(self, 'final_cost_fn'), 'Please set the final cost first!'
        if time_grid is None:
            time_grid = numpy.linspace(0, 1, numpy.amin([horizon + 1, 11]))
        if type(time_grid) == list:
            time_grid = numpy.array(time_grid)
        if numpy.isscalar(time_grid) and time_grid == -1:
            time_grid = numpy.linspace(0, horizon - 1, horizon)      # (sic) the reference's grid for -1, PDP.py:967-968
        time_grid = numpy.asarray(time_grid, dtype=numpy.float64)
        self.time_grid = numpy.rint(horizon * time_grid / time_grid[-1]).astype(int)
        self.whorizon = len(self.time_grid) - 1
        self.setPolyControl(numpy.linspace(0, self.whorizon, self.whorizon + 1))   # n_auxvar = (whorizon + 1) * m
        self._interval = numpy.repeat(numpy.arange(self.whorizon), numpy.diff(self.time_grid))

    def _warp_controls(self, auxvar_value):
        blocks = _flat(auxvar_value, self.n_auxvar, "auxvar_value").reshape(self.whorizon + 1, self.n_control)
        return blocks[:self.whorizon]

    def warp_integrateSys(self, ini_state, whorizon, auxvar_value):
        assert hasattr(self, 'time_grid'), "Warp the dynamics first by runing the method of warp_init_step! "
        dev = _device()
        Uw = self._warp_controls(auxvar_value)
        x0 = _dev_tensor(_flat(ini_state, self.n_state, "ini_state")[None], dev)
        cost, _, X = self.recmat_step_batched(x0, _dev_tensor(Uw[self._interval][None], dev))
        return {'wstate_traj': X[0].cpu().numpy()[self.time_grid], 'wcontrol_traj': Uw,
                'wcost': cost.cpu().numpy().reshape(1, 1)}

    def warp_step(self, ini_state, horizon, auxvar_value):
        assert hasattr(self, 'time_grid'), "Run warp_init_step first!"
        dev = _device()
        Uw = self._warp_controls(auxvar_value)
        x0 = _dev_tensor(_flat(ini_state, self.n_state, "ini_state")[None], dev)
        cost, dHu, _ = self.recmat_step_batched(x0, _dev_tensor(Uw[self._interval][None], dev))
        dw = self._interval_sums(dHu[0], self.whorizon + 1)
        return cost.cpu().numpy().reshape(1, 1), dw.cpu().numpy().reshape(-1)

    def warp_unwarp(self, ini_state, horizon, auxvar_value):
        dev = _device()
        U = self._warp_controls(auxvar_value)[self._interval]
        x0 = _dev_tensor(_flat(ini_state, self.n_state, "ini_state")[None], dev)
        cost, _, X = self.recmat_step_batched(x0, _dev_tensor(U[None], dev))
        return {'state_traj': X[0].cpu().numpy(), 'control_traj': U, 'cost': cost.cpu().numpy().reshape(1)}

    # ---- the reference's symbolic H-fold compositions (kept for API completeness; host-side symbolic builders) ------
    # warp_step / recmat_step above do NOT use them: they get the same losses and gradients from the adjoint kernel.
    # The objects built here are ordinary ``Function``s of the symbolic front-end (numeric calls evaluate on the host
    # like every other user-visible ``*_fn`` attribute; symbolic calls substitute).

    def _compose_interval(self, first_step, last_step):
        """(state after steps first_step..last_step-1 under one constant control, path cost summed over them), symbolic."""
        x, cost = self.state, SX(0.0)
        for _ in range(int(first_step), int(last_step)):
            cost = cost + self.path_cost_fn(x, self.control)
            x = self.dyn_fn(x, self.control)
        return x, cost

    def warp_dynCost(self, time_grid):
        """Per interval of ``time_grid``: the composed dynamics and summed path cost as functions of (state at the
        interval start, the interval's control) and their Jacobians -- the attribute lists of reference PDP.py:882-915
        (``wdyn_fns, wdfx_fns, wdfu_fns, wpath_cost_fns, wdcx_fns, wdcu_fns``, ``wfinal_cost_fn``, ``wdhx_fn``)."""
        assert hasattr(self, 'dyn_fn'), 'Please set the dynamics first!'
        assert hasattr(self, 'path_cost_fn'), 'Please set the path cost first!'
        assert hasattr(self, 'final_cost_fn'), 'Please set the final cost first!'
        args = [self.state, self.control]
        built = {name: [] for name in ("wdyn", "wdfx", "wdfu", "wpath_cost", "wdcx", "wdcu")}
        for wt, (t0, t1) in enumerate(zip(time_grid[:-1], time_grid[1:])):
            x_end, cost = self._compose_interval(t0, t1)
            outputs = {"wdyn": x_end, "wdfx": jacobian(x_end, self.state), "wdfu": jacobian(x_end, self.control),
                       "wpath_cost": cost, "wdcx": jacobian(cost, self.state), "wdcu": jacobian(cost, self.control)}
            for name, expr in outputs.items():
                built[name].append(Function("%s_fn%d" % (name, wt), args, [expr]))
        for name, fns in built.items():
            setattr(self, name + "_fns", fns)
        self.wfinal_cost_fn, self.wdhx_fn = self.final_cost_fn, self.dhx_fn

    def warp_getAuxSys(self, wstate_traj, wcontrol_traj, auxvar_value):
        """Jacobians of the composed dynamics and of the policy along a warped trajectory (reference PDP.py:940-957)
        -> dict of lists ``wdynF, wdynG, wdUx, wdUe``; needs ``warp_dynCost(self.time_grid)``."""
        assert hasattr(self, 'wdfx_fns'), "Warp the dynamics first by running warp_dynCost(self.time_grid)!"
        xs, us = numpy.asarray(wstate_traj), numpy.asarray(wcontrol_traj)
        steps = range(us.shape[0])
        return {"wdynF": [self.wdfx_fns[k](xs[k], us[k]).full() for k in steps],
                "wdynG": [self.wdfu_fns[k](xs[k], us[k]).full() for k in steps],
                "wdUx": [self.dpolicy_dx_fn(k, xs[k], auxvar_value).full() for k in steps],
                "wdUe": [self.dpolicy_de_fn(k, xs[k], auxvar_value).full() for k in steps]}

    def recmat_recoveryMatrix(self, whorizon):
        """``recovery_matrix_fn(x0, stacked controls)`` = gradient of the warped problem's cost with respect to the
        stacked interval controls, as a column (the quantity reference PDP.py:1039-1079 assembles from products of the
        interval Jacobians).  Here the warped cost is composed symbolically from ``wdyn_fns`` / ``wpath_cost_fns`` and
        differentiated by the front-end's reverse mode.  Like the reference this re-declares ``auxvar`` as the stacked
        controls.  Needs :meth:`warp_dynCost`."""
        assert hasattr(self, 'wdyn_fns'), 'Please warp the dynamics and cost function first by running warp_init_step!'
        x0 = SX.sym('X0', self.n_state)
        controls = [SX.sym('U_%d' % k, self.n_control) for k in range(whorizon)]
        x, total = x0, SX(0.0)
        for k, u_k in enumerate(controls):
            total = total + self.wpath_cost_fns[k](x, u_k)
            x = self.wdyn_fns[k](x, u_k)
        total = total + self.wfinal_cost_fn(x)
        self.auxvar = vcat(controls)
        self.n_auxvar = self.auxvar.numel()
        self.recovery_matrix_fn = Function('recovery_matrix_fn', [x0, self.auxvar], [transpose(jacobian(total, self.auxvar))])


def _forward_recursion(dynF, dynG, dUx, dUe, dynE, ini_condition):
    """X+ = F X + G (Ux X + Ue) + E on the dense LQR module in forward-only mode."""
    torch, eng = _torch(), _engine()
    dev = _device()
    H = len(dynF)
    n = dynF[0].shape[0]
    r = ini_condition.shape[1]
    m = dynG[0].shape[1] if dynG is not None else 1
    F = numpy.stack(dynF)
    G = numpy.stack(dynG) if dynG is not None else numpy.zeros((H, n, m))
    E = numpy.stack(dynE) if dynE is not None else numpy.zeros((H, n, r))
    K = numpy.stack(dUx) if dUx is not None else numpy.zeros((H, m, n))
    k = numpy.stack(dUe) if dUe is not None else numpy.zeros((H, m, r))
    rmax = 32 - n - m
    Xs, Us = [], []
    for c0 in range(0, r, rmax):
        cols = slice(c0, min(r, c0 + rmax))
        rc = cols.stop - cols.start
        zeros = lambda a, b: numpy.zeros((H, a * b))
        rec = numpy.concatenate([F.reshape(H, -1), G.reshape(H, -1), E[:, :, cols].reshape(H, -1), zeros(n, n), zeros(n, m),
                                 zeros(n, rc), zeros(m, n), zeros(m, m), zeros(m, rc)], axis=1)
        gains = numpy.concatenate([K.transpose(0, 2, 1), k[:, :, cols].transpose(0, 2, 1)], axis=1)  # [H, n+rc, m]
        Xa, Ua = eng.DenseLQR.get(n, m, rc).solve(
            _dev_tensor(rec[None], dev), None, X0aux=_dev_tensor(numpy.ascontiguousarray(ini_condition[:, cols])[None], dev),
            gains=_dev_tensor(gains[None], dev))
        Xs.append(Xa[0])
        Us.append(Ua[0])
    return torch.cat(Xs, dim=2).cpu().numpy(), torch.cat(Us, dim=2).cpu().numpy()


# ================================================================================================== SysID
class SysID:
    """System identification mode: x+ = f(x,u,auxvar) fitted to recorded input/state sequences."""

    def __init__(self, project_name='my system identification'):
        self.project_name = project_name
        self._sys = None
        self._gpu_fns = {}

    def setAuxvarVariable(self, auxvar):
        self.auxvar = auxvar
        self.n_auxvar = self.auxvar.numel()
        self._sys = None

    def setStateVariable(self, state):
        self.state = state
        self.n_state = self.state.numel()
        self.state_lb = self.n_state * [-1e20]
        self.state_ub = self.n_state * [1e20]
        self._sys = None

    def setControlVariable(self, control):
        self.control = control
        self.n_control = self.control.numel()
        self.control_lb = self.n_control * [-1e20]
        self.control_ub = self.n_control * [1e20]
        self._sys = None

    def setDyn(self, ode):
        self.dyn = SX(ode)
        xue = [self.state, self.control, self.auxvar]
        self.dyn_fn = Function('dyn_fn', xue, [self.dyn])
        self.dfx = jacobian(self.dyn, self.state)
        self.dfx_fn = Function('dfx', xue, [self.dfx])
        self.dfu = jacobian(self.dyn, self.control)
        self.dfu_fn = Function('dfu', xue, [self.dfu])
        self.dfe = jacobian(self.dyn, self.auxvar)
        self.dfe_fn = Function('dfe', xue, [self.dfe])
        self._sys = None
        self._gpu_fns = {}

    def _system(self):
        if self._sys is None:
            self._sys = _engine().SysIDSystem(self.state, self.control, self.auxvar, self.dyn)
        return self._sys

    def getRandomInputs(self, horizon=10, n_batch=1, lb=None, ub=None):
        lb = self.n_control * [-1] if lb is None else lb
        ub = self.n_control * [1] if ub is None else ub
        lo, hi = numpy.asarray(lb, dtype=numpy.float64), numpy.asarray(ub, dtype=numpy.float64)
        batch = []
        for _ in range(n_batch):
            cols = [(hi[i] - lo[i]) * numpy.random.random(horizon) + lo[i] for i in range(self.n_control)]
            batch.append(numpy.stack(cols, axis=1))
        return batch

    # -------------------------------------------------------------------------------- batched API (new)
    def step_batched(self, batch_inputs, batch_states, auxvar_value, want_traj=False, want_sens=False):
        """inputs[B,H,m], states[B,H+1,n], theta[B|1,r] (CUDA float64) -> per-trajectory loss_dp[B,r+1]
        (un-averaged; average over the GLOBAL batch, after any allreduce, to match reference PDP.py:1293)."""
        return self._system().step(batch_inputs, batch_states, auxvar_value, want_traj=want_traj, want_sens=want_sens)

    # -------------------------------------------------------------------------------- legacy API
    def integrateDyn(self, ini_state, inputs, auxvar_value):
        assert hasattr(self, 'dyn_fn'), "set the dynamics first!"
        dev = _device()
        U = numpy.asarray(inputs, dtype=numpy.float64).reshape(-1, self.n_control)
        x0 = _dev_tensor(_flat(ini_state, self.n_state, "ini_state")[None], dev)
        th = _dev_tensor(_flat(auxvar_value, self.n_auxvar, "auxvar_value")[None], dev)
        out = self._system().step(_dev_tensor(U[None], dev), None, th, x0=x0, want_traj=True)
        return out["X"][0].cpu().numpy()

    def getAuxSys(self, state_traj, control_traj, auxvar_value):
        dev = _device()
        U = numpy.asarray(control_traj, dtype=numpy.float64).reshape(-1, self.n_control)
        H = U.shape[0]
        X = _dev_tensor(numpy.asarray(state_traj, dtype=numpy.float64)[:H], dev)
        th = _dev_tensor(_flat(auxvar_value, self.n_auxvar, "auxvar_value"), dev)
        for nm in ("dfx_fn", "dfe_fn"):
            if nm not in self._gpu_fns:
                self._gpu_fns[nm] = _engine().GpuFunction(getattr(self, nm))
        Ud = _dev_tensor(U, dev)
        F = self._gpu_fns["dfx_fn"](X, Ud, th)[0].cpu().numpy()
        E = self._gpu_fns["dfe_fn"](X, Ud, th)[0].cpu().numpy()
        return {"dynF": [F[t] for t in range(H)], "dynE": [E[t] for t in range(H)]}

    def integrateAuxSys(self, dynF, dynE, ini_condition):
        if type(dynF) != list or type(dynE) != list:
            assert False, "The input dynF and dynE should be list of numpy.array!"
        if len(dynE) != len(dynF):
            assert False, "The length of dynF and dynE should be the same"
        if type(ini_condition) is not numpy.ndarray:
            assert False, "The initial condition should be numpy.array"
        Xn, _ = _forward_recursion(dynF, None, None, None, dynE, ini_condition)
        return {'state_traj': [Xn[t] for t in range(Xn.shape[0])]}

    def step(self, batch_inputs, batch_states, auxvar_value):
        """loss and half-gradient averaged over the batch (reference PDP.py:1261-1296)."""
        dev = _device()
        n_batch = len(batch_inputs)
        th = _dev_tensor(_flat(auxvar_value, self.n_auxvar, "auxvar_value")[None], dev)
        total = numpy.zeros(self.n_auxvar + 1)
        by_horizon = {}
        for inp, st in zip(batch_inputs, batch_states):
            by_horizon.setdefault(numpy.shape(inp)[0], []).append((inp, st))
        for H, items in by_horizon.items():
            U = _dev_tensor(numpy.stack([numpy.asarray(i, dtype=numpy.float64).reshape(H, self.n_control) for i, _ in items]), dev)
            Xo = _dev_tensor(numpy.stack([numpy.asarray(s, dtype=numpy.float64).reshape(H + 1, self.n_state) for _, s in items]), dev)
            total += self._system().step(U, Xo, th)["loss_dp"].sum(dim=0).cpu().numpy()
        return total[0] / n_batch, total[1:] / n_batch
