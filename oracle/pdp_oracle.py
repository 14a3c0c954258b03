"""NumPy/sympy restatement of the reference's PDP hot path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Each function cites the reference lines whose arithmetic it restates.  Per-trajectory, per-step
Python loops are kept on purpose: this mirrors how the reference executes (one CasADi call /
NumPy op per time step) and is what ``bench.py``'s ``cpu_baseline`` times.
"""
import numpy as np
import sympy as sp


def _lam(args, expr):
    return sp.lambdify(args, expr, modules="numpy", cse=True)


class OracleOC:
    """Restates ``OCSys`` (reference ``PDP/PDP.py:57-314``) for a discrete system
    ``x+ = dyn(x,u,theta)``, path cost ``c(x,u,theta)``, final cost ``h(x,theta)``."""

    def __init__(self, X, U, theta, dyn, path_cost, final_cost):
        self.X, self.U, self.theta = sp.Matrix(X), sp.Matrix(U), sp.Matrix(theta)
        self.n, self.m, self.r = len(self.X), len(self.U), len(self.theta)
        self.dyn = sp.Matrix(dyn)
        self.c = sp.sympify(path_cost)
        self.h = sp.sympify(final_cost)
        xs, us, ths = list(self.X), list(self.U), list(self.theta)
        self._a3 = (xs, us, ths)
        self.dyn_fn = _lam(self._a3, self.dyn)                                   # PDP.py:101
        self.path_cost_fn = _lam(self._a3, self.c)                               # PDP.py:110
        self.final_cost_fn = _lam((xs, ths), self.h)                             # PDP.py:119
        self._diff_done = False

    # reference PDP.py:222-270
    def diffPMP(self):
        xs, us, ths = self._a3
        lam = sp.Matrix(sp.symbols("lam0:%d" % self.n, real=True))
        ls = list(lam)
        Hm = self.c + (self.dyn.T * lam)[0, 0]                                   # :231
        a4 = (xs, us, ls, ths)
        self.dfx_fn = _lam(self._a3, self.dyn.jacobian(self.X))                  # :235-236
        self.dfu_fn = _lam(self._a3, self.dyn.jacobian(self.U))                  # :237-238
        self.dfe_fn = _lam(self._a3, self.dyn.jacobian(self.theta))              # :239-240
        dHx = sp.Matrix([Hm]).jacobian(self.X).T                                 # :243
        dHu = sp.Matrix([Hm]).jacobian(self.U).T                                 # :245
        self.dHx_fn = _lam(a4, dHx)
        self.dHu_fn = _lam(a4, dHu)
        self.ddHxx_fn = _lam(a4, dHx.jacobian(self.X))                           # :249
        self.ddHxu_fn = _lam(a4, dHx.jacobian(self.U))                           # :251
        self.ddHxe_fn = _lam(a4, dHx.jacobian(self.theta))                       # :253
        self.ddHux_fn = _lam(a4, dHu.jacobian(self.X))                           # :255
        self.ddHuu_fn = _lam(a4, dHu.jacobian(self.U))                           # :257
        self.ddHue_fn = _lam(a4, dHu.jacobian(self.theta))                       # :259
        dhx = sp.Matrix([self.h]).jacobian(self.X).T                             # :263
        self.dhx_fn = _lam((xs, ths), dhx)
        self.ddhxx_fn = _lam((xs, ths), dhx.jacobian(self.X))                    # :267
        self.ddhxe_fn = _lam((xs, ths), dhx.jacobian(self.theta))                # :269
        self.dcx_fn = _lam(self._a3, sp.Matrix([self.c]).jacobian(self.X))
        self.dcu_fn = _lam(self._a3, sp.Matrix([self.c]).jacobian(self.U))
        self._diff_done = True

    @staticmethod
    def _mat(v, shape):
        return np.asarray(v, dtype=np.float64).reshape(shape)

    def rollout(self, x0, U, theta):
        """x_{t+1} = dyn(x_t, u_t, theta); cost = sum c + h  (PDP.py:158-175 without the NLP)."""
        H = U.shape[0]
        X = np.zeros((H + 1, self.n))
        X[0] = x0
        cost = 0.0
        for t in range(H):
            X[t + 1] = self._mat(self.dyn_fn(X[t], U[t], theta), self.n)
            cost += float(np.asarray(self.path_cost_fn(X[t], U[t], theta)))
        cost += float(np.asarray(self.final_cost_fn(X[H], theta)))
        return X, cost

    def costate(self, X, U, theta):
        """costate[t] = lambda_{t+1}; lambda_H = dh/dx(x_H); lambda_t = dHx(x_t,u_t,lambda_{t+1})
        (PDP.py:203-209, the option-1 recursion; same convention as IPOPT's lam_g :195)."""
        if not self._diff_done:
            self.diffPMP()
        H = U.shape[0]
        L = np.zeros((H, self.n))
        L[H - 1] = self._mat(self.dhx_fn(X[H], theta), self.n)
        for k in range(H - 1, 0, -1):
            L[k - 1] = self._mat(self.dHx_fn(X[k], U[k], L[k], theta), self.n)
        return L

    def dHu_traj(self, X, U, L, theta):
        """dJ/du_t = dHu(x_t,u_t,lambda_{t+1}) -- the adjoint gradient equal to the reference's
        recovery-matrix product (PDP.py:1039-1079,1112)."""
        return np.stack([self._mat(self.dHu_fn(X[t], U[t], L[t], theta), self.m) for t in range(U.shape[0])])

    def getAuxSys(self, X, U, L, theta):
        """PDP.py:272-314."""
        if not self._diff_done:
            self.diffPMP()
        n, m, r = self.n, self.m, self.r
        out = {k: [] for k in ("dynF", "dynG", "dynE", "Hxx", "Hxu", "Hxe", "Hux", "Huu", "Hue")}
        for t in range(U.shape[0]):
            x, u, lam = X[t], U[t], L[t]
            out["dynF"].append(self._mat(self.dfx_fn(x, u, theta), (n, n)))
            out["dynG"].append(self._mat(self.dfu_fn(x, u, theta), (n, m)))
            out["dynE"].append(self._mat(self.dfe_fn(x, u, theta), (n, r)))
            out["Hxx"].append(self._mat(self.ddHxx_fn(x, u, lam, theta), (n, n)))
            out["Hxu"].append(self._mat(self.ddHxu_fn(x, u, lam, theta), (n, m)))
            out["Hxe"].append(self._mat(self.ddHxe_fn(x, u, lam, theta), (n, r)))
            out["Hux"].append(self._mat(self.ddHux_fn(x, u, lam, theta), (m, n)))
            out["Huu"].append(self._mat(self.ddHuu_fn(x, u, lam, theta), (m, m)))
            out["Hue"].append(self._mat(self.ddHue_fn(x, u, lam, theta), (m, r)))
        out["hxx"] = [self._mat(self.ddhxx_fn(X[-1], theta), (n, n))]
        out["hxe"] = [self._mat(self.ddhxe_fn(X[-1], theta), (n, r))]
        return out


def lqr_solve(aux, ini_state, horizon, return_gains=False):
    """Matrix-valued LQR, literal reference form (PDP.py:557-608): backward ``PP/WW`` with
    ``inv(Huu)`` and ``inv(I + P R)``, then the forward pass.  Returns lists like the reference."""
    F, G, E = aux["dynF"], aux["dynG"], aux["dynE"]
    Hxx, Huu, Hxu, Hxe, Hue = aux["Hxx"], aux["Huu"], aux["Hxu"], aux["Hxe"], aux["Hue"]
    n = F[0].shape[0]
    r = ini_state.shape[1]
    I = np.eye(n)
    PP = [None] * horizon
    WW = [None] * horizon
    PP[-1] = aux["hxx"][0]
    WW[-1] = aux["hxe"][0]
    for t in range(horizon - 1, 0, -1):
        P, W = PP[t], WW[t]
        iHuu = np.linalg.inv(Huu[t])
        GiH = G[t] @ iHuu
        HxuiH = Hxu[t] @ iHuu
        A = F[t] - GiH @ Hxu[t].T
        R = GiH @ G[t].T
        M = E[t] - GiH @ Hue[t]
        Q = Hxx[t] - HxuiH @ Hxu[t].T
        N = Hxe[t] - HxuiH @ Hue[t]
        T = A.T @ np.linalg.inv(I + P @ R)
        PP[t - 1] = Q + T @ (P @ A)
        WW[t - 1] = N + T @ (W + P @ M)
    Xs = [ini_state]
    Us, Ls, Ks, ks = [], [], [], []
    for t in range(horizon):
        P, W = PP[t], WW[t]
        iHuu = np.linalg.inv(Huu[t])
        GiH = G[t] @ iHuu
        A = F[t] - GiH @ Hxu[t].T
        M = E[t] - GiH @ Hue[t]
        R = GiH @ G[t].T
        x = Xs[t]
        T2 = np.linalg.inv(I + P @ R)
        u = -iHuu @ (Hxu[t].T @ x + Hue[t]) - iHuu @ G[t].T @ T2 @ (P @ A @ x + P @ M + W)
        if return_gains:  # u = K x + k  (same expression, split into its linear and affine parts)
            Ks.append(-iHuu @ Hxu[t].T - iHuu @ G[t].T @ T2 @ P @ A)
            ks.append(-iHuu @ Hue[t] - iHuu @ G[t].T @ T2 @ (P @ M + W))
        xn = F[t] @ x + G[t] @ u + E[t]
        Xs.append(xn)
        Us.append(u)
        Ls.append(P @ xn + W)
    out = {"state_traj_opt": Xs, "control_traj_opt": Us, "costate_traj_opt": Ls}
    if return_gains:
        out["K"], out["k"] = Ks, ks
    return out


def irl_loss_grad(X, U, Xd, Ud, dX, dU):
    """Loss and (half-)gradient of reference ``Examples/IRL/quadrotor/uav_PDP.py:67-75``."""
    dldx, dldu = X - Xd, U - Ud
    loss = np.linalg.norm(dldx) ** 2 + np.linalg.norm(dldu) ** 2
    dp = np.zeros(dX[0].shape[1])
    for t in range(U.shape[0]):
        dp = dp + dldx[t] @ dX[t] + dldu[t] @ dU[t]
    dp = dp + dldx[-1] @ dX[-1]
    return loss, dp


def pdp_sweep(oc: OracleOC, x0, U, theta):
    """One 'sweep' (SURVEY 8(d)): rollout, costate, aux system, aux-LQR -> X, Lam, dX/dtheta, dU/dtheta."""
    X, cost = oc.rollout(x0, U, theta)
    L = oc.costate(X, U, theta)
    aux = oc.getAuxSys(X, U, L, theta)
    sol = lqr_solve(aux, np.zeros((oc.n, oc.r)), U.shape[0])
    return X, L, cost, np.stack(sol["state_traj_opt"]), np.stack(sol["control_traj_opt"])


class OracleSysID:
    """Restates ``SysID`` (PDP.py:1157-1296)."""

    def __init__(self, X, U, theta, dyn):
        self.X, self.U, self.theta = sp.Matrix(X), sp.Matrix(U), sp.Matrix(theta)
        self.n, self.m, self.r = len(self.X), len(self.U), len(self.theta)
        a3 = (list(self.X), list(self.U), list(self.theta))
        dyn = sp.Matrix(dyn)
        self.dyn_fn = _lam(a3, dyn)                                  # :1180
        self.dfx_fn = _lam(a3, dyn.jacobian(self.X))                 # :1183
        self.dfe_fn = _lam(a3, dyn.jacobian(self.theta))             # :1187

    def integrateDyn(self, x0, inputs, theta):                       # :1209-1223
        H = inputs.shape[0]
        X = np.zeros((H + 1, self.n))
        X[0] = x0
        for t in range(H):
            X[t + 1] = np.asarray(self.dyn_fn(X[t], inputs[t], theta), dtype=np.float64).reshape(self.n)
        return X

    def sens(self, X, inputs, theta):                                # :1225-1259
        S = [np.zeros((self.n, self.r))]
        for t in range(inputs.shape[0]):
            Ft = np.asarray(self.dfx_fn(X[t], inputs[t], theta), dtype=np.float64).reshape(self.n, self.n)
            Et = np.asarray(self.dfe_fn(X[t], inputs[t], theta), dtype=np.float64).reshape(self.n, self.r)
            S.append(Ft @ S[t] + Et)
        return S

    def step(self, batch_inputs, batch_states, theta):               # :1261-1296
        loss, dp = 0.0, np.zeros(self.r)
        for inputs, obs in zip(batch_inputs, batch_states):
            X = self.integrateDyn(obs[0], inputs, theta)
            S = self.sens(X, inputs, theta)
            d = X - obs
            loss += np.linalg.norm(d) ** 2
            for t in range(inputs.shape[0]):
                dp += d[t] @ S[t]
            dp += d[-1] @ S[-1]
        nb = len(batch_inputs)
        return loss / nb, dp / nb


class OracleCP:
    """Restates ``ControlPlanning`` with a parameterised policy (PDP.py:640-878)."""

    def __init__(self, X, U, dyn, path_cost, final_cost):
        self.X, self.U = sp.Matrix(X), sp.Matrix(U)
        self.n, self.m = len(self.X), len(self.U)
        xs, us = list(self.X), list(self.U)
        dyn = sp.Matrix(dyn)
        c, h = sp.sympify(path_cost), sp.sympify(final_cost)
        self.dyn_fn = _lam((xs, us), dyn)                                        # :674
        self.dfx_fn = _lam((xs, us), dyn.jacobian(self.X))                       # :677
        self.dfu_fn = _lam((xs, us), dyn.jacobian(self.U))                       # :679
        self.path_cost_fn = _lam((xs, us), c)                                    # :684
        self.dcx_fn = _lam((xs, us), sp.Matrix([c]).jacobian(self.X))            # :689
        self.dcu_fn = _lam((xs, us), sp.Matrix([c]).jacobian(self.U))            # :690
        self.final_cost_fn = _lam((xs,), h)                                      # :694
        self.dhx_fn = _lam((xs,), sp.Matrix([h]).jacobian(self.X))               # :697

    def set_poly(self, pivots):                                                  # :699-725
        t = sp.Symbol("t", real=True)
        K = len(pivots)
        Us = [sp.Matrix(sp.symbols("U%d_0:%d" % (i, self.m), real=True)) for i in range(K)]
        pol = sp.zeros(self.m, 1)
        for i in range(K):
            b = 1
            for j in range(K):
                if j != i:
                    b = b * (t - pivots[j]) / (pivots[i] - pivots[j])
            pol = pol + b * Us[i]
        self.theta = sp.Matrix([s for Ui in Us for s in Ui])
        self._set_policy(t, pol)

    def set_neural(self, hidden_layers):                                         # :727-759 (column-major packing)
        t = sp.Symbol("t", real=True)
        layers = list(hidden_layers) + [self.m]
        a = self.X
        params = []
        n_in = self.n
        for li, n_out in enumerate(layers):
            A = sp.Matrix(n_out, n_in, lambda i, j: sp.Symbol("A%d_%d_%d" % (li, i, j), real=True))
            b = sp.Matrix(n_out, 1, lambda i, j: sp.Symbol("b%d_%d" % (li, i), real=True))
            params += [A[i, j] for j in range(n_in) for i in range(n_out)]       # column-major reshape((-1,1))
            params += list(b)
            if li > 0:
                a = a.applyfunc(sp.tanh)
            a = A * a + b
            n_in = n_out
        self.theta = sp.Matrix(params)
        self._set_policy(t, a)

    def _set_policy(self, t, pol):
        self.r = len(self.theta)
        a = ([t], list(self.X), list(self.theta))
        self.policy_fn = _lam(a, pol)
        self.dpolicy_dx_fn = _lam(a, pol.jacobian(self.X))
        self.dpolicy_de_fn = _lam(a, pol.jacobian(self.theta))

    def integrateSys(self, x0, H, theta):                                        # :763-786
        X = np.zeros((H + 1, self.n)); U = np.zeros((H, self.m)); X[0] = x0
        cost = 0.0
        for t in range(H):
            U[t] = np.asarray(self.policy_fn([t], X[t], theta), dtype=np.float64).reshape(self.m)
            X[t + 1] = np.asarray(self.dyn_fn(X[t], U[t]), dtype=np.float64).reshape(self.n)
            cost += float(np.asarray(self.path_cost_fn(X[t], U[t])))
        cost += float(np.asarray(self.final_cost_fn(X[H])))
        return X, U, cost

    def step(self, x0, H, theta, return_traj=False):                             # :850-878
        n, m, r = self.n, self.m, self.r
        X, U, cost = self.integrateSys(x0, H, theta)
        dX = [np.zeros((n, r))]
        dU = []
        for t in range(H):
            F = np.asarray(self.dfx_fn(X[t], U[t]), dtype=np.float64).reshape(n, n)
            Gm = np.asarray(self.dfu_fn(X[t], U[t]), dtype=np.float64).reshape(n, m)
            Ux = np.asarray(self.dpolicy_dx_fn([t], X[t], theta), dtype=np.float64).reshape(m, n)
            Ue = np.asarray(self.dpolicy_de_fn([t], X[t], theta), dtype=np.float64).reshape(m, r)
            Ut = Ux @ dX[t] + Ue                                                 # :832
            dX.append(F @ dX[t] + Gm @ Ut)                                       # :833
            dU.append(Ut)
        g = np.zeros(r)
        for t in range(H):
            g += (np.asarray(self.dcx_fn(X[t], U[t]), dtype=np.float64).reshape(1, n) @ dX[t] +
                  np.asarray(self.dcu_fn(X[t], U[t]), dtype=np.float64).reshape(1, m) @ dU[t]).ravel()
        g += (np.asarray(self.dhx_fn(X[H]), dtype=np.float64).reshape(1, n) @ dX[H]).ravel()
        if return_traj:
            return cost, g, X, U, np.stack(dX), np.stack(dU)
        return cost, g

    def adjoint_grad(self, x0, Useq):
        """recmat semantics (PDP.py:1100-1114 with time_grid=-1): J(U) and dJ/dU by the costate."""
        H = Useq.shape[0]
        n, m = self.n, self.m
        X = np.zeros((H + 1, n)); X[0] = x0
        cost = 0.0
        for t in range(H):
            cost += float(np.asarray(self.path_cost_fn(X[t], Useq[t])))
            X[t + 1] = np.asarray(self.dyn_fn(X[t], Useq[t]), dtype=np.float64).reshape(n)
        cost += float(np.asarray(self.final_cost_fn(X[H])))
        lam = np.asarray(self.dhx_fn(X[H]), dtype=np.float64).reshape(n)
        g = np.zeros((H, m))
        for t in range(H - 1, -1, -1):
            F = np.asarray(self.dfx_fn(X[t], Useq[t]), dtype=np.float64).reshape(n, n)
            Gm = np.asarray(self.dfu_fn(X[t], Useq[t]), dtype=np.float64).reshape(n, m)
            g[t] = np.asarray(self.dcu_fn(X[t], Useq[t]), dtype=np.float64).reshape(m) + Gm.T @ lam
            lam = np.asarray(self.dcx_fn(X[t], Useq[t]), dtype=np.float64).reshape(n) + F.T @ lam
        return cost, g, X


def warp_step(cp: "OracleCP", x0, horizon, time_grid, theta):
    """Restates ControlPlanning.warp_init_step/warp_step (PDP.py:882-1008) numerically: warped dynamics/cost =
    composition over each grid interval, Lagrange policy over the integer warped steps, forward sensitivity
    X_{w+1} = wF X_w + wG U_w with U_w = dUe (dUx = 0), chain rule with the warped cost gradients."""
    n, m = cp.n, cp.m
    time_grid = np.asarray(time_grid, dtype=np.float64)
    grid = np.rint(horizon * time_grid / time_grid[-1]).astype(int)
    wh = len(grid) - 1
    pivots = np.linspace(0, wh, wh + 1)
    r = (wh + 1) * m
    theta = np.asarray(theta, dtype=np.float64).reshape(wh + 1, m)

    def basis(wt):
        b = np.ones(wh + 1)
        for i in range(wh + 1):
            for j in range(wh + 1):
                if j != i:
                    b[i] *= (wt - pivots[j]) / (pivots[i] - pivots[j])
        return b

    X = np.asarray(x0, dtype=np.float64)
    dXdth = np.zeros((n, r))
    cost, grad = 0.0, np.zeros(r)
    for wt in range(wh):
        b = basis(wt)
        u = b @ theta                                               # policy_fn(wt, x, theta)
        dUe = np.kron(b[None, :], np.eye(m))                        # d u / d theta  (m x r)
        Sx, Su = np.eye(n), np.zeros((n, m))                        # d x_t / d X_wt, d x_t / d u inside the interval
        cx_w, cu_w = np.zeros(n), np.zeros(m)
        x = X.copy()
        for t in range(grid[wt], grid[wt + 1]):
            cost += float(np.asarray(cp.path_cost_fn(x, u)))
            cx = np.asarray(cp.dcx_fn(x, u), dtype=np.float64).reshape(n)
            cu = np.asarray(cp.dcu_fn(x, u), dtype=np.float64).reshape(m)
            cx_w += cx @ Sx
            cu_w += cx @ Su + cu
            F = np.asarray(cp.dfx_fn(x, u), dtype=np.float64).reshape(n, n)
            G = np.asarray(cp.dfu_fn(x, u), dtype=np.float64).reshape(n, m)
            Sx, Su = F @ Sx, F @ Su + G
            x = np.asarray(cp.dyn_fn(x, u), dtype=np.float64).reshape(n)
        grad += cx_w @ dXdth + cu_w @ dUe                           # PDP.py:1002-1005
        dXdth = Sx @ dXdth + Su @ dUe                               # integrateAuxSys with wdynF, wdynG
        X = x
    cost += float(np.asarray(cp.final_cost_fn(X)))
    grad += np.asarray(cp.dhx_fn(X), dtype=np.float64).reshape(n) @ dXdth
    return cost, grad


def build_oc(env: dict, dt, theta_syms=None):
    """OCSys on ``dyn = X + dt*f`` with auxvar = [dyn_params, cost_params] (reference
    ``Examples/IRL/quadrotor/uav_PDP.py:20-27``)."""
    theta = list(env["dyn_params"]) + list(env["cost_params"]) if theta_syms is None else theta_syms
    return OracleOC(env["X"], env["U"], theta, env["X"] + dt * env["f"], env["path_cost"], env["final_cost"])
