"""On-disk compatibility with the reference's example scripts: `.mat` files with exactly the
schema the reference writes and its `*_results_plot.py` / `*_validation.py` scripts read.

* results of a learning run -- ``{'results': {...}}`` with the reference's field names (including its ``trail_no``
  spelling): IRL ``Examples/IRL/pendulum/pendulum_PDP.py:91-97``, SysID ``Examples/SysID/quadrotor/uav_PDP.py:54-59``,
  OC ``Examples/OC/rocket/rocket_PDP_Recmat.py:68-78``;
* demonstrations -- ``{'trajectories': [ocSolver dicts], 'dt', 'true_parameter'}``
  (``Examples/IRL/quadrotor/generate_demos.py``);
* SysID input/output data -- ``{'<env>_iodata': {'batch_inputs', 'batch_states', 'true_parameter'}}``.

Host-side helpers only (numpy + scipy.io); nothing here is on the hot path."""
from __future__ import annotations

import numpy as np
import scipy.io as sio


def _host(a):
    """torch tensor / list / ndarray -> float64 ndarray on the host."""
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.asarray(a, dtype=np.float64)


def save_results(path, trial_no, loss_trace, parameter_trace, learning_rate, time_passed, initial_parameter=None,
                 row_vectors=True, **extra):
    """Write ``{'results': ...}`` as the reference's scripts do.  ``parameter_trace``: K parameter vectors; they are
    stored as (K, 1, r) like the reference's list of (1, r) arrays (T9), ``loss_trace`` as (1, K).  ``extra`` takes
    the mode-specific fields (``solved_solution``, ``true_solution``, ``dt``, ``horizon``, ...); the OC scripts keep
    1-D parameter vectors (``row_vectors=False`` -> (K, r))."""
    ptrace = [_host(p).reshape(1, -1) if row_vectors else _host(p).reshape(-1) for p in parameter_trace]
    data = {"trail_no": int(trial_no)}
    if initial_parameter is not None:
        data["initial_parameter"] = _host(initial_parameter).reshape(1, -1)
    data["loss_trace"] = [float(_host(v).reshape(-1)[0]) for v in loss_trace]
    data["parameter_trace"] = ptrace
    data["learning_rate"] = float(learning_rate)
    data["time_passed"] = float(time_passed)
    for k, v in extra.items():
        data[k] = v
    sio.savemat(path, {"results": data})


def load_results(path):
    """-> dict of host arrays: ``loss_trace`` (K,), ``parameter_trace`` (K, r), scalars as Python numbers; other
    fields as stored."""
    s = sio.loadmat(path)["results"][0, 0]
    out = {}
    for name in s.dtype.names:
        v = s[name]
        if name == "loss_trace":
            out[name] = np.asarray(v, dtype=np.float64).reshape(-1)
        elif name == "parameter_trace":
            v = np.asarray(v, dtype=np.float64)
            out[name] = v.reshape(v.shape[0], -1)
        elif name in ("trail_no", "horizon"):
            out[name] = int(np.asarray(v).reshape(-1)[0])
        elif name in ("learning_rate", "time_passed", "dt"):
            out[name] = float(np.asarray(v).reshape(-1)[0])
        elif name == "initial_parameter":
            out[name] = np.asarray(v, dtype=np.float64).reshape(-1)
        else:
            out[name] = v
    return out


_TRAJ_FIELDS = ("state_traj_opt", "control_traj_opt", "costate_traj_opt", "auxvar_value", "time", "horizon", "cost")


def save_demos(path, trajectories, dt, true_parameter):
    """``trajectories``: list of ocSolver-style dicts (reference PDP.py:212-218)."""
    trajs = []
    for t in trajectories:
        d = {}
        for k in _TRAJ_FIELDS:
            if k in t:
                d[k] = int(t[k]) if k == "horizon" else _host(t[k])
        trajs.append(d)
    sio.savemat(path, {"trajectories": trajs, "dt": float(dt), "true_parameter": _host(true_parameter).reshape(1, -1)})


def load_demos(path):
    """-> (list of dicts with ``state_traj_opt`` (H+1,n), ``control_traj_opt`` (H,m), ``costate_traj_opt`` (H,n), ...,
    dt, true_parameter (r,)) from a file written by the reference's ``generate_demos.py`` or by :func:`save_demos`."""
    d = sio.loadmat(path)
    raw = d["trajectories"]
    out = []
    for i in range(raw.shape[1]):
        s = raw[0, i]
        s = s[0, 0] if s.dtype.names and s.shape == (1, 1) else s
        t = {}
        for name in s.dtype.names:
            v = np.asarray(s[name])
            if name == "horizon":
                t[name] = int(v.reshape(-1)[0])
            elif name == "cost":
                t[name] = float(v.reshape(-1)[0])
            else:
                t[name] = np.asarray(v, dtype=np.float64)
        out.append(t)
    return out, float(np.asarray(d["dt"]).reshape(-1)[0]), np.asarray(d["true_parameter"], dtype=np.float64).reshape(-1)


def save_iodata(path, key, batch_inputs, batch_states, true_parameter):
    sio.savemat(path, {key: {"batch_inputs": [_host(u) for u in batch_inputs], "batch_states": [_host(x) for x in batch_states],
                             "true_parameter": _host(true_parameter).reshape(-1)}})


def load_iodata(path):
    """-> (inputs [B,H,m], states [B,H+1,n], true_parameter (r,)) from ``*_iodata.mat``."""
    d = sio.loadmat(path)
    key = [k for k in d if not k.startswith("__")][0]
    s = d[key][0, 0]
    return (np.asarray(s["batch_inputs"], dtype=np.float64), np.asarray(s["batch_states"], dtype=np.float64),
            np.asarray(s["true_parameter"], dtype=np.float64).reshape(-1))
