"""Named systems of the BASELINE.json configs, built from the drop-in ``JinEnv`` models.

Each builder returns a compiled engine object; ``__graft_entry__.build()`` calls all of them so
the generated modules are compiled in-tree and travel to the GPU box.
"""
from __future__ import annotations

import functools

from . import engine
from .symbolic import vertcat


def _jinenv():
    from JinEnv import JinEnv
    return JinEnv


@functools.lru_cache(maxsize=None)
def quadrotor_irl(dt: float = 0.1):
    """C3: quadrotor IRL, n=13 m=4 r=9 (reference Examples/IRL/quadrotor/uav_PDP.py:9-28)."""
    env = _jinenv().Quadrotor()
    env.initDyn(c=0.01)
    env.initCost(wthrust=0.1)
    return engine.OCSystem(env.X, env.U, vertcat(env.dyn_auxvar, env.cost_auxvar), env.X + dt * env.f,
                           env.path_cost, env.final_cost)


@functools.lru_cache(maxsize=None)
def pendulum_irl(dt: float = 0.1):
    """C1: pendulum IRL, n=2 m=1 r=5 (reference Examples/IRL/pendulum/pendulum_PDP.py:9-30)."""
    env = _jinenv().SinglePendulum()
    env.initDyn()
    env.initCost()
    return engine.OCSystem(env.X, env.U, vertcat(env.dyn_auxvar, env.cost_auxvar), env.X + dt * env.f,
                           env.path_cost, env.final_cost)


@functools.lru_cache(maxsize=None)
def rocket_irl(dt: float = 0.1):
    """Rocket IRL, n=13 m=3 r=10 (reference Examples/IRL/rocket/rocket_PDP.py)."""
    env = _jinenv().Rocket()
    env.initDyn()
    env.initCost(wthrust=0.1)
    return engine.OCSystem(env.X, env.U, vertcat(env.dyn_auxvar, env.cost_auxvar), env.X + dt * env.f,
                           env.path_cost, env.final_cost)


@functools.lru_cache(maxsize=None)
def cartpole_irl(dt: float = 0.1):
    env = _jinenv().CartPole()
    env.initDyn()
    env.initCost(wu=0.1)
    return engine.OCSystem(env.X, env.U, vertcat(env.dyn_auxvar, env.cost_auxvar), env.X + dt * env.f,
                           env.path_cost, env.final_cost)


@functools.lru_cache(maxsize=None)
def robotarm_irl(dt: float = 0.1):
    env = _jinenv().RobotArm()
    env.initDyn(g=0)
    env.initCost(wu=0.01)
    return engine.OCSystem(env.X, env.U, vertcat(env.dyn_auxvar, env.cost_auxvar), env.X + dt * env.f,
                           env.path_cost, env.final_cost)


OC_BUILDERS = {"quadrotor": quadrotor_irl, "pendulum": pendulum_irl, "rocket": rocket_irl,
               "cartpole": cartpole_irl, "robotarm": robotarm_irl}


def build_all(verbose=False):
    out = {}
    for name, fn in OC_BUILDERS.items():
        out["oc_" + name] = fn().module_path
        if verbose:
            print("built", name, out["oc_" + name])
    for name, fn in SENS_BUILDERS.items():
        out[name] = fn().module_path
        if verbose:
            print("built", name, out[name])
    out["oc_rocket_adjoint"] = rocket_oc_adjoint().module_path          # C4 (bench.py --config c4, tests/test_gpu_configs.py)
    if verbose:
        print("built rocket_adjoint", out["oc_rocket_adjoint"])
    # Newton (ocSolver) variants, generic dense-LQR sizes used by the drop-in LQR class on the shipped examples
    from . import ocsolver
    for name, fn in OC_BUILDERS.items():
        out["newton_" + name] = ocsolver.newton_system(fn()).module_path
    for dims in ((13, 4, 9), (2, 1, 5), (4, 1, 7), (4, 2, 8), (13, 3, 10), (5, 2, 4), (5, 1, 4)):
        out["lqr_%d_%d_%d" % dims] = engine.DenseLQR.get(*dims).module_path
    if verbose:
        print("built %d modules in total" % len(out))
    return out


# ------------------------------------------------------------------------------ SysID / ControlPlanning
@functools.lru_cache(maxsize=None)
def quadrotor_sysid(dt: float = 0.1):
    """C5: quadrotor SysID, n=13 m=4 r=5 (reference Examples/SysID/quadrotor/uav_PDP.py:9-19)."""
    env = _jinenv().Quadrotor()
    env.initDyn(c=0.01)
    return engine.SysIDSystem(env.X, env.U, env.dyn_auxvar, env.X + dt * env.f)


def lagrange_policy(n_control, pivots, tvar):
    """u(t, theta) = sum_i b_i(t) U_i, Lagrange basis over ``pivots`` (reference PDP/PDP.py:699-725)."""
    from .symbolic import SX, vcat
    pol = 0
    params = []
    for i in range(len(pivots)):
        Ui = SX.sym('U_' + str(i), n_control)
        params.append(Ui)
        bi = 1
        for j in range(len(pivots)):
            if j != i:
                bi = bi * (tvar - pivots[j]) / (pivots[i] - pivots[j])
        pol = pol + bi * Ui
    return pol, vcat(params)


def neural_policy(state, n_control, hidden_layers):
    """tanh MLP with column-major packed weights (reference PDP/PDP.py:727-759)."""
    from .symbolic import SX, mtimes, tanh, vcat
    layers = list(hidden_layers) + [n_control]
    a = state
    params = []
    n_in = state.numel()
    for li, n_out in enumerate(layers):
        Ak = SX.sym('Ak', n_out, n_in)
        bk = SX.sym('bk', n_out)
        params += [Ak.reshape((-1, 1)), bk]
        if li > 0:
            a = tanh(a)
        a = mtimes(Ak, a) + bk
        n_in = n_out
    return a, vcat(params)


@functools.lru_cache(maxsize=None)
def cartpole_cp(policy: str = "poly", horizon: int = 50, dt: float = 0.05):
    """C2: cartpole ControlPlanning n=4 m=1 (reference Examples/OC/cartpole/cartpole_PDP_poly.py:13-30,
    cartpole_PDP_neural.py:13-30,49): poly r=6 (n_poly=5), neural [4,4] r=45."""
    import numpy as np
    from .symbolic import SX
    env = _jinenv().CartPole()
    env.initDyn(mc=0.1, mp=0.1, l=1)
    env.initCost(wx=0.1, wq=0.6, wdx=0.1, wdq=0.1, wu=0.3)
    t = SX.sym('t')
    if policy == "poly":
        pol, th = lagrange_policy(1, np.linspace(0, horizon, 6), t)
    else:
        pol, th = neural_policy(env.X, 1, [4, 4])
    return engine.CPSystem(env.X, env.U, th, env.X + dt * env.f, pol, t, env.path_cost, env.final_cost)


SENS_BUILDERS = {"sysid_quadrotor": quadrotor_sysid,
                 "cp_cartpole_poly": lambda: cartpole_cp("poly"), "cp_cartpole_neural": lambda: cartpole_cp("neural")}


@functools.lru_cache(maxsize=None)
def rocket_oc_adjoint(dt: float = 0.1):
    """C4: rocket powered-landing OC, n=13 m=3, gradient dJ/dU by the costate kernel (reference
    Examples/OC/rocket/rocket_PDP_Recmat.py:10-23 parameters; recmat semantics PDP/PDP.py:1100-1114)."""
    from .symbolic import SX
    env = _jinenv().Rocket()
    env.initDyn(Jx=0.5, Jy=1., Jz=1., mass=1., l=1.)
    env.initCost(wr=1, wv=1, wtilt=50, ww=1, wsidethrust=1, wthrust=0.4)
    return engine.OCSystem(env.X, env.U, SX.sym('unused_auxvar'), env.X + dt * env.f, env.path_cost, env.final_cost)
