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
    return out
