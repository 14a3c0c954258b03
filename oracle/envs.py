"""sympy restatement of the five reference environments (TEST INFRASTRUCTURE).

Follows the *semantics* of reference ``JinEnv/JinEnv.py``: SinglePendulum :37-100, RobotArm
:176-278, CartPole :360-430, Quadrotor :537-670 (+ helpers :831-854), Rocket :883-1016,
``toQuaternion`` :1192-1199.  Independent of the product's own symbolic engine on purpose, so
that product-vs-oracle agreement is a real check.

Every builder returns a dict with sympy column ``Matrix`` objects ``X``, ``U``, ``f`` (continuous
dynamics), lists ``dyn_params`` / ``cost_params`` (symbols, in the reference's auxvar order) and
scalar ``path_cost`` / ``final_cost``.  Fixed numeric arguments are baked in like the reference
does when a value (not ``None``) is passed.
"""
import math

import numpy as np
import sympy as sp

G = 10  # reference uses g = 10 everywhere (JinEnv.py:39,362,539,885)


def _param(value, name, bag):
    if value is None:
        s = sp.Symbol(name, real=True)
        bag.append(s)
        return s
    return sp.Float(value) if isinstance(value, float) else sp.sympify(value)


def to_quaternion(angle, axis):
    axis = np.asarray(axis, dtype=float)
    axis = axis / np.linalg.norm(axis)
    return np.concatenate([[math.cos(angle / 2)], math.sin(angle / 2) * axis])


def dir_cosine(q):
    q0, q1, q2, q3 = q
    return sp.Matrix([
        [1 - 2 * (q2 ** 2 + q3 ** 2), 2 * (q1 * q2 + q0 * q3), 2 * (q1 * q3 - q0 * q2)],
        [2 * (q1 * q2 - q0 * q3), 1 - 2 * (q1 ** 2 + q3 ** 2), 2 * (q2 * q3 + q0 * q1)],
        [2 * (q1 * q3 + q0 * q2), 2 * (q2 * q3 - q0 * q1), 1 - 2 * (q1 ** 2 + q2 ** 2)]])


def skew(v):
    return sp.Matrix([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def omega(w):
    return sp.Matrix([[0, -w[0], -w[1], -w[2]],
                      [w[0], 0, w[2], -w[1]],
                      [w[1], -w[2], 0, w[0]],
                      [w[2], w[1], -w[0], 0]])


def pendulum(l=None, m=None, damping_ratio=None, wq=None, wdq=None, wu=0.001):
    dp, cp = [], []
    l = _param(l, 'l', dp); m = _param(m, 'm', dp); d = _param(damping_ratio, 'damping_ratio', dp)
    wq = _param(wq, 'wq', cp); wdq = _param(wdq, 'wdq', cp)
    q, dq, u = sp.symbols('q dq u', real=True)
    I = sp.Rational(1, 3) * m * l * l
    f = sp.Matrix([dq, (u - m * G * l * sp.sin(q) - d * dq) / I])
    final = wq * (q - math.pi) ** 2 + wdq * dq ** 2
    return dict(X=sp.Matrix([q, dq]), U=sp.Matrix([u]), f=f, dyn_params=dp, cost_params=cp,
                path_cost=final + wu * u * u, final_cost=final)


def robotarm(l1=None, m1=None, l2=None, m2=None, g=10, wq1=None, wq2=None, wdq1=None, wdq2=None, wu=0.1):
    dp, cp = [], []
    l1 = _param(l1, 'l1', dp); m1 = _param(m1, 'm1', dp); l2 = _param(l2, 'l2', dp); m2 = _param(m2, 'm2', dp)
    wq1 = _param(wq1, 'wq1', cp); wq2 = _param(wq2, 'wq2', cp)
    wdq1 = _param(wdq1, 'wdq1', cp); wdq2 = _param(wdq2, 'wdq2', cp)
    q1, dq1, q2, dq2, u1, u2 = sp.symbols('q1 dq1 q2 dq2 u1 u2', real=True)
    r1, r2 = l1 / 2, l2 / 2
    I1, I2 = l1 * l1 * m1 / 12, l2 * l2 * m2 / 12
    M11 = m1 * r1 * r1 + I1 + m2 * (l1 * l1 + r2 * r2 + 2 * l1 * r2 * sp.cos(q2)) + I2
    M12 = m2 * (r2 * r2 + l1 * r2 * sp.cos(q2)) + I2
    M22 = m2 * r2 * r2 + I2
    M = sp.Matrix([[M11, M12], [M12, M22]])
    h = m2 * l1 * r2 * sp.sin(q2)
    C = sp.Matrix([-h * dq2 * dq2 - 2 * h * dq1 * dq2, h * dq1 * dq1])
    Gv = sp.Matrix([m1 * r1 * g * sp.cos(q1) + m2 * g * (r2 * sp.cos(q1 + q2) + l1 * sp.cos(q1)),
                    m2 * g * r2 * sp.cos(q1 + q2)])
    U = sp.Matrix([u1, u2])
    det = M11 * M22 - M12 * M12
    Minv = sp.Matrix([[M22, -M12], [-M12, M11]]) / det
    ddq = Minv * (-C - Gv + U)
    f = sp.Matrix([dq1, dq2, ddq[0], ddq[1]])
    final = wq1 * (q1 - math.pi / 2) ** 2 + wq2 * q2 ** 2 + wdq1 * dq1 ** 2 + wdq2 * dq2 ** 2
    return dict(X=sp.Matrix([q1, q2, dq1, dq2]), U=U, f=f, dyn_params=dp, cost_params=cp,
                path_cost=final + wu * (u1 * u1 + u2 * u2), final_cost=final)


def cartpole(mc=None, mp=None, l=None, wx=None, wq=None, wdx=None, wdq=None, wu=0.001):
    dp, cp = [], []
    mc = _param(mc, 'mc', dp); mp = _param(mp, 'mp', dp); l = _param(l, 'l', dp)
    wx = _param(wx, 'wx', cp); wq = _param(wq, 'wq', cp); wdx = _param(wdx, 'wdx', cp); wdq = _param(wdq, 'wdq', cp)
    x, q, dx, dq, u = sp.symbols('x q dx dq u', real=True)
    ddx = (u + mp * sp.sin(q) * (l * dq * dq + G * sp.cos(q))) / (mc + mp * sp.sin(q) ** 2)
    ddq = (-u * sp.cos(q) - mp * l * dq * dq * sp.sin(q) * sp.cos(q) - (mc + mp) * G * sp.sin(q)) / (
        l * mc + l * mp * sp.sin(q) ** 2)
    final = wx * x ** 2 + wq * (q - math.pi) ** 2 + wdx * dx ** 2 + wdq * dq ** 2
    return dict(X=sp.Matrix([x, q, dx, dq]), U=sp.Matrix([u]), f=sp.Matrix([dx, dq, ddx, ddq]),
                dyn_params=dp, cost_params=cp, path_cost=final + wu * u * u, final_cost=final)


def _sixdof_symbols(unames):
    r = sp.Matrix(sp.symbols('rx ry rz', real=True))
    v = sp.Matrix(sp.symbols('vx vy vz', real=True))
    q = sp.Matrix(sp.symbols('q0 q1 q2 q3', real=True))
    w = sp.Matrix(sp.symbols('wx wy wz', real=True))
    u = sp.Matrix(sp.symbols(unames, real=True))
    return r, v, q, w, u


def quadrotor(Jx=None, Jy=None, Jz=None, mass=None, l=None, c=None,
              wr=None, wv=None, wq=None, ww=None, wthrust=0.1):
    dp, cp = [], []
    Jx = _param(Jx, 'Jx', dp); Jy = _param(Jy, 'Jy', dp); Jz = _param(Jz, 'Jz', dp)
    mass = _param(mass, 'mass', dp); l = _param(l, 'l', dp); c = _param(c, 'c', dp)
    wr = _param(wr, 'wr', cp); wv = _param(wv, 'wv', cp); wq = _param(wq, 'wq', cp); ww = _param(ww, 'ww', cp)
    r, v, q, w, T = _sixdof_symbols('f1 f2 f3 f4')
    J = sp.diag(Jx, Jy, Jz)
    thrust = sp.Matrix([0, 0, T[0] + T[1] + T[2] + T[3]])
    Mb = sp.Matrix([-T[1] * l / 2 + T[3] * l / 2, -T[0] * l / 2 + T[2] * l / 2, (T[0] - T[1] + T[2] - T[3]) * c])
    C_B_I = dir_cosine(q)
    dv = (C_B_I.T * thrust) / mass + sp.Matrix([0, 0, -G])
    dq = omega(w) * q / 2
    dw = sp.diag(1 / Jx, 1 / Jy, 1 / Jz) * (Mb - skew(w) * J * w)
    f = sp.Matrix.vstack(v, dv, dq, dw)
    goal_R = dir_cosine(sp.Matrix(to_quaternion(0, [0, 0, 1]).tolist()))
    cost_q = (sp.eye(3) - goal_R.T * C_B_I).trace()
    final = wr * r.dot(r) + wv * v.dot(v) + ww * w.dot(w) + wq * cost_q
    return dict(X=sp.Matrix.vstack(r, v, q, w), U=T, f=f, dyn_params=dp, cost_params=cp,
                path_cost=final + wthrust * T.dot(T), final_cost=final)


def rocket(Jx=None, Jy=None, Jz=None, mass=None, l=None,
           wr=None, wv=None, wtilt=None, ww=None, wsidethrust=None, wthrust=1.0):
    dp, cp = [], []
    Jx = _param(Jx, 'Jx', dp); Jy = _param(Jy, 'Jy', dp); Jz = _param(Jz, 'Jz', dp)
    mass = _param(mass, 'mass', dp); l = _param(l, 'l', dp)
    # auxvar order of the reference: wr, wv, wtilt, wsidethrust, ww (JinEnv.py:948-976)
    wr = _param(wr, 'wr', cp); wv = _param(wv, 'wv', cp); wtilt = _param(wtilt, 'wtilt', cp)
    wsidethrust = _param(wsidethrust, 'wsidethrust', cp); ww = _param(ww, 'ww', cp)
    r, v, q, w, T = _sixdof_symbols('ux uy uz')
    J = sp.diag(Jx, Jy, Jz)
    rT = sp.Matrix([-l / 2, 0, 0])
    C_I_B = dir_cosine(q).T
    dv = (C_I_B * T) / mass + sp.Matrix([-G, 0, 0])
    dq = omega(w) * q / 2
    dw = sp.diag(1 / Jx, 1 / Jy, 1 / Jz) * (skew(rT) * T - skew(w) * J * w)
    f = sp.Matrix.vstack(v, dv, dq, dw)
    nose = C_I_B * sp.Matrix([1, 0, 0])
    cost_tilt = nose[1] ** 2 + nose[2] ** 2
    final = wr * r.dot(r) + wv * v.dot(v) + ww * w.dot(w) + wtilt * cost_tilt
    path = final + wsidethrust * (T[1] ** 2 + T[2] ** 2) + wthrust * T.dot(T)
    return dict(X=sp.Matrix.vstack(r, v, q, w), U=T, f=f, dyn_params=dp, cost_params=cp,
                path_cost=path, final_cost=final)


BUILDERS = dict(pendulum=pendulum, robotarm=robotarm, cartpole=cartpole, quadrotor=quadrotor, rocket=rocket)
