"""CPU optimal-control solve for the oracle (TEST INFRASTRUCTURE): stands in for the reference's
IPOPT call (PDP/PDP.py:131-182) when checking the end-to-end IRL goldens (K2/K3).

Globalised Newton on the controls; the Newton direction comes from the oracle's literal
reference-form ``lqr_solve`` (one column, Hue := dH/du).  Independent of the CUDA solver's code
(NumPy + sympy lambdas), same mathematics (closed-loop DDP-style trial rollouts), so agreement at convergence checks both against the
shipped IPOPT solutions."""
import numpy as np

from . import pdp_oracle


def _direction(oc, X, U, L, theta, s, mu):
    H = U.shape[0]
    aux = oc.getAuxSys(X, U, s * L, theta)
    g = oc.dHu_traj(X, U, L, theta)
    n, m = oc.n, oc.m
    aux["dynE"] = [np.zeros((n, 1))] * H
    aux["Hxe"] = [np.zeros((n, 1))] * H
    aux["Hue"] = [g[t].reshape(m, 1) for t in range(H)]
    aux["Huu"] = [aux["Huu"][t] + mu * np.eye(m) for t in range(H)]
    aux["hxe"] = [np.zeros((n, 1))]
    sol = pdp_oracle.lqr_solve(aux, np.zeros((n, 1)), H, return_gains=True)
    return np.stack(sol["control_traj_opt"])[:, :, 0], g, sol["K"], sol["k"]


def _closed_loop(oc, x0, X, U, K, k, alpha, theta):
    """DDP-style trial rollout u_t = u_t + alpha k_t + K_t (x_t - x_t_old)."""
    H = U.shape[0]
    Xn = np.zeros_like(X); Un = np.zeros_like(U)
    Xn[0] = x0
    cost = 0.0
    for t in range(H):
        Un[t] = U[t] + alpha * k[t][:, 0] + K[t] @ (Xn[t] - X[t])
        cost += float(np.asarray(oc.path_cost_fn(Xn[t], Un[t], theta)))
        Xn[t + 1] = np.asarray(oc.dyn_fn(Xn[t], Un[t], theta), dtype=np.float64).reshape(oc.n)
    cost += float(np.asarray(oc.final_cost_fn(Xn[H], theta)))
    return Xn, Un, cost


def solve(oc, x0, H, theta, U0=None, tol=1e-9, max_iter=400):
    U = np.zeros((H, oc.m)) if U0 is None else U0.copy()
    X, cost = oc.rollout(x0, U, theta)
    L = oc.costate(X, U, theta)
    s, mu = 1.0, 0.0
    for it in range(max_iter):
        g = oc.dHu_traj(X, U, L, theta)
        if np.max(np.abs(g)) <= tol * (1 + np.max(np.abs(L))):
            break
        try:
            dU, g, K, k = _direction(oc, X, U, L, theta, s, mu)
            slope = float(np.sum(g * dU))
            ok_dir = np.all(np.isfinite(dU)) and slope < 0
        except np.linalg.LinAlgError:
            ok_dir = False
        accepted = False
        if ok_dir:
            alpha = 1.0
            for _ in range(25):
                Xt, Ut, ct = _closed_loop(oc, x0, X, U, K, k, alpha, theta)
                if np.isfinite(ct) and ct <= cost + 1e-4 * alpha * slope:
                    U, X, cost = Ut, Xt, ct
                    L = oc.costate(X, U, theta)
                    accepted = True
                    break
                alpha *= 0.5
        if accepted:
            if alpha >= 0.5:
                s = 1.0
            mu *= 0.1
            if mu < 1e-12:
                mu = 0.0
        else:
            if s == 0.0:
                mu = max(mu * 10.0, 1e-6)
            s = 0.0
    return X, U, L, cost, it
