"""Drop-in ``JinEnv`` environments for the B200 PDP engine.

Same class surface as the reference's ``JinEnv/JinEnv.py`` (``SinglePendulum`` :33, ``RobotArm``
:171, ``CartPole`` :356, ``Quadrotor`` :519, ``Rocket`` :865, ``toQuaternion`` :1192): every class
exposes symbolic ``X, U, f, dyn_auxvar, path_cost, final_cost, cost_auxvar`` after
``initDyn``/``initCost``.  Here those expressions are hot-path INPUT: ``PDP`` code-generates them
into CUDA ``__device__`` functions.  The models are re-derived from the physics the reference
states (gravity g = 10 everywhere, parameter ordering = order in which unspecified arguments
are declared), written against this repo's own symbolic front-end.

matplotlib is imported lazily inside the animation helpers only (the reference imports it at
module top, ``JinEnv.py:21-27``), so headless benchmark boxes can import this module.
"""
import math

import numpy as np

from pontryagin_differentiable_programming_b200.symbolic import (
    SX, cos, diag, dot, horzcat, inv, mtimes, sin, trace, transpose, vcat, vertcat)

GRAVITY = 10.0


class _ParamBag:
    """Collects the symbols created for arguments the caller left as ``None``."""

    def __init__(self):
        self.symbols = []

    def take(self, value, name):
        if value is None:
            s = SX.sym(name)
            self.symbols.append(s)
            return s
        return value

    def stacked(self):
        return vcat(self.symbols)


# ------------------------------------------------------------------------------ rigid-body algebra
def _dcm_body_from_inertial(q):
    """Direction-cosine matrix C_B<-I of a unit quaternion q = (q0, q1, q2, q3), scalar first."""
    a, b, c, d = q[0], q[1], q[2], q[3]
    return vertcat(
        horzcat(1 - 2 * (c ** 2 + d ** 2), 2 * (b * c + a * d), 2 * (b * d - a * c)),
        horzcat(2 * (b * c - a * d), 1 - 2 * (b ** 2 + d ** 2), 2 * (c * d + a * b)),
        horzcat(2 * (b * d + a * c), 2 * (c * d - a * b), 1 - 2 * (b ** 2 + c ** 2)))


def _cross_matrix(v):
    return vertcat(horzcat(0, -v[2], v[1]),
                   horzcat(v[2], 0, -v[0]),
                   horzcat(-v[1], v[0], 0))


def _quat_rate_matrix(w):
    """Omega(w) with dq/dt = 1/2 * Omega(w) q."""
    return vertcat(horzcat(0, -w[0], -w[1], -w[2]),
                   horzcat(w[0], 0, w[2], -w[1]),
                   horzcat(w[1], -w[2], 0, w[0]),
                   horzcat(w[2], w[1], -w[0], 0))


def _six_dof_state(u_names):
    r = vertcat(*[SX.sym(n) for n in ("rx", "ry", "rz")])
    v = vertcat(*[SX.sym(n) for n in ("vx", "vy", "vz")])
    q = vertcat(*[SX.sym(n) for n in ("q0", "q1", "q2", "q3")])
    w = vertcat(*[SX.sym(n) for n in ("wx", "wy", "wz")])
    u = vertcat(*[SX.sym(n) for n in u_names])
    return r, v, q, w, u


def _lazy_pyplot():
    try:
        import matplotlib.pyplot as plt
        import matplotlib.animation as animation
    except ImportError as exc:  # pragma: no cover - visual helper only
        raise ImportError("JinEnv animations need matplotlib, which is not part of the numeric path") from exc
    return plt, animation


def _animate_lines(frames, draw, interval, title, save_option, fname):  # pragma: no cover - visual helper
    plt, animation = _lazy_pyplot()
    fig = plt.figure()
    ax = draw(fig, None, init=True)
    ani = animation.FuncAnimation(fig, lambda k: draw(fig, k, ax=ax), frames, interval=interval)
    if title:
        fig.suptitle(title)
    if save_option != 0:
        ani.save(fname, writer="ffmpeg", fps=int(1000 / max(interval, 1)))
    plt.show()
    return ani


# ------------------------------------------------------------------------------ single pendulum
class SinglePendulum:
    def __init__(self, project_name='single pendlumn system'):
        self.project_name = project_name

    def initDyn(self, l=None, m=None, damping_ratio=None):
        bag = _ParamBag()
        self.l = bag.take(l, 'l')
        self.m = bag.take(m, 'm')
        self.damping_ratio = bag.take(damping_ratio, 'damping_ratio')
        self.dyn_auxvar = bag.stacked()

        self.q, self.dq = SX.sym('q'), SX.sym('dq')
        self.X = vertcat(self.q, self.dq)
        self.U = SX.sym('u')
        inertia = 1 / 3 * self.m * self.l * self.l
        torque = self.U - self.m * GRAVITY * self.l * sin(self.q) - self.damping_ratio * self.dq
        self.f = vertcat(self.dq, torque / inertia)

    def initCost(self, wq=None, wdq=None, wu=0.001):
        bag = _ParamBag()
        self.wq = bag.take(wq, 'wq')
        self.wdq = bag.take(wdq, 'wdq')
        self.cost_auxvar = bag.stacked()

        self.cost_q = (self.q - math.pi) ** 2
        self.cost_dq = (self.dq - 0) ** 2
        self.cost_u = dot(self.U, self.U)
        self.final_cost = self.wq * self.cost_q + self.wdq * self.cost_dq
        self.path_cost = self.final_cost + wu * self.cost_u

    def get_pendulum_position(self, len, state_traj):
        ang = np.asarray(state_traj)[:, 0]
        return np.stack([len * np.sin(ang), -len * np.cos(ang)], axis=1)

    def play_animation(self, len, dt, state_traj, state_traj_ref=None, save_option=0):  # pragma: no cover
        pos = self.get_pendulum_position(len, state_traj)
        ref = self.get_pendulum_position(len, state_traj_ref) if state_traj_ref is not None else None

        def draw(fig, k, ax=None, init=False):
            if init:
                ax = fig.add_subplot(111, autoscale_on=False, xlim=(-len - 1, len + 1), ylim=(-len - 1, len + 1))
                ax.set_aspect('equal')
                ax.lines_ = [ax.plot([], [], 'o-', lw=3)[0], ax.plot([], [], 'o-', color='gray', alpha=.4)[0]]
                return ax
            ax.lines_[0].set_data([0, pos[k, 0]], [0, pos[k, 1]])
            if ref is not None:
                ax.lines_[1].set_data([0, ref[min(k, ref.shape[0] - 1), 0]], [0, ref[min(k, ref.shape[0] - 1), 1]])
            return ax

        _animate_lines(pos.shape[0], draw, 50, 'Pendulum', save_option, 'Pendulum.mp4')


# ------------------------------------------------------------------------------ two-link arm
class RobotArm:
    def __init__(self, project_name='two-link robot arm'):
        self.project_name = project_name

    def initDyn(self, l1=None, m1=None, l2=None, m2=None, g=10):
        bag = _ParamBag()
        self.l1 = bag.take(l1, 'l1')
        self.m1 = bag.take(m1, 'm1')
        self.l2 = bag.take(l2, 'l2')
        self.m2 = bag.take(m2, 'm2')
        self.dyn_auxvar = bag.stacked()

        self.q1, self.dq1, self.q2, self.dq2 = SX.sym('q1'), SX.sym('dq1'), SX.sym('q2'), SX.sym('dq2')
        self.X = vertcat(self.q1, self.q2, self.dq1, self.dq2)
        self.U = vertcat(SX.sym('u1'), SX.sym('u2'))

        l1_, l2_, m1_, m2_ = self.l1, self.l2, self.m1, self.m2
        c1, c2 = l1_ / 2, l2_ / 2  # link centres of mass
        j1, j2 = l1_ * l1_ * m1_ / 12, l2_ * l2_ * m2_ / 12  # link inertias about the centres
        # joint-space inertia matrix
        m11 = m1_ * c1 * c1 + j1 + m2_ * (l1_ * l1_ + c2 * c2 + 2 * l1_ * c2 * cos(self.q2)) + j2
        m12 = m2_ * (c2 * c2 + l1_ * c2 * cos(self.q2)) + j2
        m22 = m2_ * c2 * c2 + j2
        inertia = vertcat(horzcat(m11, m12), horzcat(m12, m22))
        # Coriolis / centrifugal and gravity torques
        hh = m2_ * l1_ * c2 * sin(self.q2)
        coriolis = vertcat(-hh * self.dq2 * self.dq2 - 2 * hh * self.dq1 * self.dq2, hh * self.dq1 * self.dq1)
        grav = vertcat(
            m1_ * c1 * g * cos(self.q1) + m2_ * g * (c2 * cos(self.q1 + self.q2) + l1_ * cos(self.q1)),
            m2_ * g * c2 * cos(self.q1 + self.q2))
        ddq = mtimes(inv(inertia), -coriolis - grav + self.U)
        self.f = vertcat(self.dq1, self.dq2, ddq)

    def initCost(self, wq1=None, wq2=None, wdq1=None, wdq2=None, wu=0.1):
        bag = _ParamBag()
        self.wq1 = bag.take(wq1, 'wq1')
        self.wq2 = bag.take(wq2, 'wq2')
        self.wdq1 = bag.take(wdq1, 'wdq1')
        self.wdq2 = bag.take(wdq2, 'wdq2')
        self.cost_auxvar = bag.stacked()

        self.cost_q1 = (self.q1 - math.pi / 2) ** 2
        self.cost_q2 = (self.q2 - 0) ** 2
        self.cost_dq1 = (self.dq1 - 0) ** 2
        self.cost_dq2 = (self.dq2 - 0) ** 2
        self.cost_u = dot(self.U, self.U)
        self.final_cost = (self.wq1 * self.cost_q1 + self.wq2 * self.cost_q2 +
                           self.wdq1 * self.cost_dq1 + self.wdq2 * self.cost_dq2)
        self.path_cost = self.final_cost + wu * self.cost_u

    def get_arm_position(self, l1, l2, state_traj):
        s = np.asarray(state_traj)
        x1, y1 = l1 * np.cos(s[:, 0]), l1 * np.sin(s[:, 0])
        x2, y2 = x1 + l2 * np.cos(s[:, 0] + s[:, 1]), y1 + l2 * np.sin(s[:, 0] + s[:, 1])
        return np.stack([x1, y1, x2, y2], axis=1)

    def play_animation(self, l1, l2, dt, state_traj, state_traj_ref=None, save_option=0):  # pragma: no cover
        pos = self.get_arm_position(l1, l2, state_traj)

        def draw(fig, k, ax=None, init=False):
            if init:
                lim = l1 + l2 + 0.5
                ax = fig.add_subplot(111, autoscale_on=False, xlim=(-lim, lim), ylim=(-lim, lim))
                ax.set_aspect('equal')
                ax.lines_ = [ax.plot([], [], 'o-', lw=3)[0]]
                return ax
            ax.lines_[0].set_data([0, pos[k, 0], pos[k, 2]], [0, pos[k, 1], pos[k, 3]])
            return ax

        _animate_lines(pos.shape[0], draw, 100, 'Robot arm', save_option, 'robot_arm.mp4')


# ------------------------------------------------------------------------------ cart-pole
class CartPole:
    def __init__(self, project_name='cart-pole-system'):
        self.project_name = project_name

    def initDyn(self, mc=None, mp=None, l=None):
        bag = _ParamBag()
        self.mc = bag.take(mc, 'mc')
        self.mp = bag.take(mp, 'mp')
        self.l = bag.take(l, 'l')
        self.dyn_auxvar = bag.stacked()

        self.x, self.q, self.dx, self.dq = SX.sym('x'), SX.sym('q'), SX.sym('dx'), SX.sym('dq')
        self.X = vertcat(self.x, self.q, self.dx, self.dq)
        self.U = SX.sym('u')
        sq_, cq_ = sin(self.q), cos(self.q)
        denom = self.mc + self.mp * sq_ * sq_
        ddx = (self.U + self.mp * sq_ * (self.l * self.dq * self.dq + GRAVITY * cq_)) / denom
        ddq = (-self.U * cq_ - self.mp * self.l * self.dq * self.dq * sq_ * cq_
               - (self.mc + self.mp) * GRAVITY * sq_) / (self.l * self.mc + self.l * self.mp * sq_ * sq_)
        self.f = vertcat(self.dx, self.dq, ddx, ddq)

    def initCost(self, wx=None, wq=None, wdx=None, wdq=None, wu=0.001):
        bag = _ParamBag()
        self.wx = bag.take(wx, 'wx')
        self.wq = bag.take(wq, 'wq')
        self.wdx = bag.take(wdx, 'wdx')
        self.wdq = bag.take(wdq, 'wdq')
        self.cost_auxvar = bag.stacked()

        goal = (0.0, math.pi, 0.0, 0.0)
        self.final_cost = (self.wx * (self.x - goal[0]) ** 2 + self.wq * (self.q - goal[1]) ** 2 +
                           self.wdx * (self.dx - goal[2]) ** 2 + self.wdq * (self.dq - goal[3]) ** 2)
        self.path_cost = self.final_cost + wu * (self.U * self.U)

    def get_cartpole_position(self, pole_len, state_traj):
        s = np.asarray(state_traj)
        return np.stack([s[:, 0], np.zeros(s.shape[0]),
                         s[:, 0] + pole_len * np.sin(s[:, 1]), -pole_len * np.cos(s[:, 1])], axis=1)

    def play_animation(self, pole_len, dt, state_traj, state_traj_ref=None, save_option=0,
                       title='Cart-pole system'):  # pragma: no cover
        pos = self.get_cartpole_position(pole_len, state_traj)

        def draw(fig, k, ax=None, init=False):
            if init:
                ax = fig.add_subplot(111, autoscale_on=False, xlim=(-10, 10), ylim=(-5, 5))
                ax.set_aspect('equal')
                ax.lines_ = [ax.plot([], [], 's-', lw=3)[0]]
                return ax
            ax.lines_[0].set_data([pos[k, 0], pos[k, 2]], [pos[k, 1], pos[k, 3]])
            return ax

        _animate_lines(pos.shape[0], draw, 50, title, save_option, 'cartpole.mp4')


# ------------------------------------------------------------------------------ quadrotor
class Quadrotor:
    def __init__(self, project_name='my UAV'):
        self.project_name = 'my uav'
        self.r_I, self.v_I, self.q, self.w_B, self.T_B = _six_dof_state(("f1", "f2", "f3", "f4"))

    def initDyn(self, Jx=None, Jy=None, Jz=None, mass=None, l=None, c=None):
        bag = _ParamBag()
        self.Jx = bag.take(Jx, 'Jx')
        self.Jy = bag.take(Jy, 'Jy')
        self.Jz = bag.take(Jz, 'Jz')
        self.mass = bag.take(mass, 'mass')
        self.l = bag.take(l, 'l')
        self.c = bag.take(c, 'c')
        self.dyn_auxvar = bag.stacked()

        self.J_B = diag(vertcat(self.Jx, self.Jy, self.Jz))
        self.g_I = vertcat(0, 0, -GRAVITY)
        self.m = self.mass

        f1, f2, f3, f4 = self.T_B[0], self.T_B[1], self.T_B[2], self.T_B[3]
        self.thrust_B = vertcat(0, 0, f1 + f2 + f3 + f4)
        half_arm = self.l / 2
        self.M_B = vertcat(-f2 * half_arm + f4 * half_arm,
                           -f1 * half_arm + f3 * half_arm,
                           (f1 - f2 + f3 - f4) * self.c)

        body_to_inertial = transpose(self.dir_cosine(self.q))
        d_pos = self.v_I
        d_vel = 1 / self.m * mtimes(body_to_inertial, self.thrust_B) + self.g_I
        d_quat = 1 / 2 * mtimes(self.omega(self.w_B), self.q)
        gyro = mtimes(mtimes(self.skew(self.w_B), self.J_B), self.w_B)
        d_rate = mtimes(inv(self.J_B), self.M_B - gyro)

        self.X = vertcat(self.r_I, self.v_I, self.q, self.w_B)
        self.U = self.T_B
        self.f = vertcat(d_pos, d_vel, d_quat, d_rate)

    def initCost(self, wr=None, wv=None, wq=None, ww=None, wthrust=0.1):
        bag = _ParamBag()
        self.wr = bag.take(wr, 'wr')
        self.wv = bag.take(wv, 'wv')
        self.wq = bag.take(wq, 'wq')
        self.ww = bag.take(ww, 'ww')
        self.cost_auxvar = bag.stacked()

        self.cost_r_I = dot(self.r_I, self.r_I)
        self.cost_v_I = dot(self.v_I, self.v_I)
        # attitude error w.r.t. the identity attitude: trace(I - R_goal^T R)
        goal_dcm = self.dir_cosine(toQuaternion(0, [0, 0, 1]))
        self.cost_q = trace(np.identity(3) - mtimes(transpose(goal_dcm), self.dir_cosine(self.q)))
        self.cost_w_B = dot(self.w_B, self.w_B)
        self.cost_thrust = dot(self.T_B, self.T_B)

        self.final_cost = (self.wr * self.cost_r_I + self.wv * self.cost_v_I +
                           self.ww * self.cost_w_B + self.wq * self.cost_q)
        self.path_cost = self.final_cost + wthrust * self.cost_thrust

    def dir_cosine(self, q):
        return _dcm_body_from_inertial(q)

    def skew(self, v):
        return _cross_matrix(v)

    def omega(self, w):
        return _quat_rate_matrix(w)

    def quaternion_mul(self, p, q):
        return vertcat(p[0] * q[0] - p[1] * q[1] - p[2] * q[2] - p[3] * q[3],
                       p[0] * q[1] + p[1] * q[0] + p[2] * q[3] - p[3] * q[2],
                       p[0] * q[2] - p[1] * q[3] + p[2] * q[0] + p[3] * q[1],
                       p[0] * q[3] + p[1] * q[2] - p[2] * q[1] + p[3] * q[0])

    def get_quadrotor_position(self, wing_len, state_traj):
        s = np.asarray(state_traj)
        arms = np.array([[wing_len / 2, 0, 0], [0, -wing_len / 2, 0], [-wing_len / 2, 0, 0], [0, wing_len / 2, 0]])
        out = np.zeros((s.shape[0], 15))
        for t in range(s.shape[0]):
            c_ib = _numeric_dcm(s[t, 6:10]).T
            out[t, 0:3] = s[t, 0:3]
            for k in range(4):
                out[t, 3 + 3 * k:6 + 3 * k] = s[t, 0:3] + c_ib @ arms[k]
        return out

    def play_animation(self, wing_len, state_traj, state_traj_ref=None, dt=0.1, save_option=0,
                       title='UAV Maneuvering'):  # pragma: no cover
        pos = self.get_quadrotor_position(wing_len, state_traj)

        def draw(fig, k, ax=None, init=False):
            if init:
                ax = fig.add_subplot(111, projection='3d')
                ax.set_xlim(-10, 10), ax.set_ylim(-10, 10), ax.set_zlim(0, 10)
                ax.lines_ = [ax.plot([], [], [], lw=2)[0] for _ in range(3)]
                return ax
            ax.lines_[0].set_data(pos[:k + 1, 0], pos[:k + 1, 1]), ax.lines_[0].set_3d_properties(pos[:k + 1, 2])
            ax.lines_[1].set_data(pos[k, [3, 9]], pos[k, [4, 10]]), ax.lines_[1].set_3d_properties(pos[k, [5, 11]])
            ax.lines_[2].set_data(pos[k, [6, 12]], pos[k, [7, 13]]), ax.lines_[2].set_3d_properties(pos[k, [8, 14]])
            return ax

        _animate_lines(pos.shape[0], draw, 100, title, save_option, 'uav.mp4')


# ------------------------------------------------------------------------------ rocket
class Rocket:
    def __init__(self, project_name='rocket powered landing'):
        self.project_name = project_name
        self.r_I, self.v_I, self.q, self.w_B, self.T_B = _six_dof_state(("ux", "uy", "uz"))

    def initDyn(self, Jx=None, Jy=None, Jz=None, mass=None, l=None):
        bag = _ParamBag()
        self.Jx = bag.take(Jx, 'Jx')
        self.Jy = bag.take(Jy, 'Jy')
        self.Jz = bag.take(Jz, 'Jz')
        self.mass = bag.take(mass, 'mass')
        self.l = bag.take(l, 'l')
        self.dyn_auxvar = bag.stacked()

        self.J_B = diag(vertcat(self.Jx, self.Jy, self.Jz))
        self.g_I = vertcat(-GRAVITY, 0, 0)  # the rocket's "up" is the inertial x axis
        self.r_T_B = vertcat(-self.l / 2, 0, 0)  # gimbal point relative to the centre of mass
        self.m = self.mass

        body_to_inertial = transpose(self.dir_cosine(self.q))
        d_pos = self.v_I
        d_vel = 1 / self.m * mtimes(body_to_inertial, self.T_B) + self.g_I
        d_quat = 1 / 2 * mtimes(self.omega(self.w_B), self.q)
        torque = mtimes(self.skew(self.r_T_B), self.T_B)
        gyro = mtimes(mtimes(self.skew(self.w_B), self.J_B), self.w_B)
        d_rate = mtimes(inv(self.J_B), torque - gyro)

        self.X = vertcat(self.r_I, self.v_I, self.q, self.w_B)
        self.U = self.T_B
        self.f = vertcat(d_pos, d_vel, d_quat, d_rate)

    def initCost(self, wr=None, wv=None, wtilt=None, ww=None, wsidethrust=None, wthrust=1.0):
        # NB declaration order (=> cost_auxvar order) is wr, wv, wtilt, wsidethrust, ww
        bag = _ParamBag()
        self.wr = bag.take(wr, 'wr')
        self.wv = bag.take(wv, 'wv')
        self.wtilt = bag.take(wtilt, 'wtilt')
        self.wsidethrust = bag.take(wsidethrust, 'wsidethrust')
        self.ww = bag.take(ww, 'ww')
        self.cost_auxvar = bag.stacked()

        self.cost_r_I = dot(self.r_I, self.r_I)
        self.cost_v_I = dot(self.v_I, self.v_I)
        # tilt: the body x axis expressed in the inertial frame should have no y / z component
        nose_I = mtimes(transpose(self.dir_cosine(self.q)), np.array([1., 0., 0.]))
        self.cost_tilt = dot(np.array([0., 1., 0.]), nose_I) ** 2 + dot(np.array([0., 0., 1.]), nose_I) ** 2
        self.cost_side_thrust = self.T_B[1] ** 2 + self.T_B[2] ** 2
        self.cost_thrust = dot(self.T_B, self.T_B)
        self.cost_w_B = dot(self.w_B, self.w_B)

        self.final_cost = (self.wr * self.cost_r_I + self.wv * self.cost_v_I +
                           self.ww * self.cost_w_B + self.wtilt * self.cost_tilt)
        self.path_cost = (self.final_cost + self.wsidethrust * self.cost_side_thrust +
                          wthrust * self.cost_thrust)

    def dir_cosine(self, q):
        return _dcm_body_from_inertial(q)

    def skew(self, v):
        return _cross_matrix(v)

    def omega(self, w):
        return _quat_rate_matrix(w)

    def get_rocket_body_position(self, rocket_len, state_traj, control_traj):
        s, u = np.asarray(state_traj), np.asarray(control_traj)
        tail = np.array([-rocket_len / 2, 0, 0])
        fmax = np.amax(np.linalg.norm(u, axis=1))
        out = np.zeros((u.shape[0], 12))
        for t in range(u.shape[0]):
            c_ib = _numeric_dcm(s[t, 6:10]).T
            rg = s[t, 0:3] + c_ib @ tail
            out[t, 0:3] = s[t, 0:3]
            out[t, 3:6] = rg
            out[t, 6:9] = s[t, 0:3] - c_ib @ tail
            out[t, 9:12] = rg - c_ib @ u[t, 0:3] / fmax
        return out

    def play_animation(self, rocket_len, state_traj, control_traj, state_traj_ref=None, control_traj_ref=None,
                       save_option=0, dt=0.1, title='Rocket Powered Landing'):  # pragma: no cover
        pos = self.get_rocket_body_position(rocket_len, state_traj, control_traj)

        def draw(fig, k, ax=None, init=False):
            if init:
                ax = fig.add_subplot(111, projection='3d')
                ax.set_xlim(-10, 10), ax.set_ylim(-10, 10), ax.set_zlim(0, 12)
                ax.lines_ = [ax.plot([], [], [], lw=2)[0] for _ in range(2)]
                return ax
            # plot with the inertial x axis pointing up
            ax.lines_[0].set_data(pos[:k + 1, 1], pos[:k + 1, 2]), ax.lines_[0].set_3d_properties(pos[:k + 1, 0])
            ax.lines_[1].set_data(pos[k, [4, 7]], pos[k, [5, 8]]), ax.lines_[1].set_3d_properties(pos[k, [3, 6]])
            return ax

        _animate_lines(pos.shape[0], draw, 100, title, save_option, 'rocket.mp4')


# ------------------------------------------------------------------------------ helpers
def _numeric_dcm(q):
    a, b, c, d = (float(v) for v in q)
    return np.array([[1 - 2 * (c * c + d * d), 2 * (b * c + a * d), 2 * (b * d - a * c)],
                     [2 * (b * c - a * d), 1 - 2 * (b * b + d * d), 2 * (c * d + a * b)],
                     [2 * (b * d + a * c), 2 * (c * d - a * b), 1 - 2 * (b * b + c * c)]])


def toQuaternion(angle, dir):
    """(angle, axis) -> unit quaternion as a 4-element list, scalar part first."""
    axis = np.asarray(dir, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    return [math.cos(angle / 2)] + (math.sin(angle / 2) * axis).tolist()


def normalizeVec(vec):
    vec = np.asarray(vec, dtype=np.float64)
    return vec / np.linalg.norm(vec)


def quaternion_conj(q):
    # like the reference (JinEnv.py:1210-1215) this negates the vector part IN PLACE and returns q
    q[1], q[2], q[3] = -q[1], -q[2], -q[3]
    return q
