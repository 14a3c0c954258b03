"""results_io: the `.mat` files written here have the schema of the files the reference ships (tests/golden/
schema_mat.json, extracted from the reference's own data by tests/golden/make_golden.py --schema) and round-trip."""
import json
import os

import numpy as np
import scipy.io as sio

from pontryagin_differentiable_programming_b200 import results_io

G = os.path.join(os.path.dirname(__file__), "golden")
SCHEMA = json.load(open(os.path.join(G, "schema_mat.json")))


def _fields(struct):
    return {n: (np.asarray(struct[n]).dtype.kind, np.asarray(struct[n]).ndim) for n in struct.dtype.names}


def _expect(name, key):
    return {n: (np.dtype(e["dtype"]).kind if e["dtype"] != "struct" else "V", e["ndim"]) for n, e in SCHEMA[name][key]["fields"].items()}


def test_irl_results_schema_and_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    K, r = 7, 5
    ptrace = [rng.standard_normal((1, r)) for _ in range(K)]
    losses = [np.array([[float(k)]]) for k in range(K)]          # the scripts append (1,1) arrays and floats alike
    p = str(tmp_path / "PDP_results_trial_0.mat")
    results_io.save_results(p, 0, losses, ptrace, 1e-4, 12.5, initial_parameter=ptrace[0])
    s = sio.loadmat(p)["results"]
    assert s.shape == (1, 1)
    got, want = _fields(s[0, 0]), _expect("irl_results", "results")
    assert set(got) == set(want)
    for n in want:
        assert got[n][1] == want[n][1], n
        assert got[n][0] == want[n][0] or {got[n][0], want[n][0]} <= {"i", "u"}, n
    back = results_io.load_results(p)
    assert back["trail_no"] == 0 and back["learning_rate"] == 1e-4 and back["time_passed"] == 12.5
    assert np.array_equal(back["loss_trace"], np.arange(K, dtype=float))
    assert np.array_equal(back["parameter_trace"], np.concatenate(ptrace))
    assert np.array_equal(back["initial_parameter"], ptrace[0].ravel())


def test_sysid_and_oc_results_schema(tmp_path):
    rng = np.random.default_rng(1)
    p = str(tmp_path / "sysid.mat")
    results_io.save_results(p, 3, [1.0, 0.5], [rng.standard_normal(5), rng.standard_normal(5)], 1e-3, 1.0)
    got, want = _fields(sio.loadmat(p)["results"][0, 0]), _expect("sysid_results", "results")
    assert set(got) == set(want) and all(got[n][1] == want[n][1] for n in want)
    p = str(tmp_path / "oc.mat")
    sol = {"state_traj": rng.standard_normal((4, 3)), "control_traj": rng.standard_normal((3, 1)), "cost": 1.5}
    results_io.save_results(p, 0, [1.0, 0.5], [rng.standard_normal(3), rng.standard_normal(3)], 1e-3, 1.0, row_vectors=False,
                            solved_solution=sol, true_solution=sol, dt=0.1, horizon=3)
    got, want = _fields(sio.loadmat(p)["results"][0, 0]), _expect("oc_results", "results")
    assert set(got) == set(want) and all(got[n][1] == want[n][1] for n in want)
    back = results_io.load_results(p)
    assert back["horizon"] == 3 and back["dt"] == 0.1 and back["parameter_trace"].shape == (2, 3)


def test_demos_and_iodata_round_trip(tmp_path):
    rng = np.random.default_rng(2)
    H, n, m, r = 6, 3, 2, 4
    trajs = [{"state_traj_opt": rng.standard_normal((H + 1, n)), "control_traj_opt": rng.standard_normal((H, m)),
              "costate_traj_opt": rng.standard_normal((H, n)), "auxvar_value": rng.standard_normal(r), "time": np.arange(H + 1),
              "horizon": H, "cost": np.array([[2.5]])} for _ in range(3)]
    p = str(tmp_path / "demos.mat")
    results_io.save_demos(p, trajs, 0.1, rng.standard_normal((1, r)))
    raw = sio.loadmat(p)
    assert raw["trajectories"].shape == (1, 3) and raw["dt"].shape == (1, 1) and raw["true_parameter"].shape == (1, r)
    want = _expect("irl_demos", "trajectories")
    s = raw["trajectories"][0, 0]
    s = s[0, 0] if s.shape == (1, 1) else s
    assert set(s.dtype.names) == set(want)
    back, dt, theta = results_io.load_demos(p)
    assert dt == 0.1 and theta.shape == (r,) and len(back) == 3
    for a, b in zip(back, trajs):
        assert np.array_equal(a["state_traj_opt"], b["state_traj_opt"]) and np.array_equal(a["control_traj_opt"], b["control_traj_opt"])
        assert a["horizon"] == H and a["cost"] == 2.5
    p = str(tmp_path / "uav_iodata.mat")
    U, X = rng.standard_normal((4, H, m)), rng.standard_normal((4, H + 1, n))
    results_io.save_iodata(p, "uav_iodata", list(U), list(X), rng.standard_normal(r))
    got, want = _fields(sio.loadmat(p)["uav_iodata"][0, 0]), _expect("sysid_iodata", "uav_iodata")
    assert set(got) == set(want) and got["batch_inputs"][1] == 3
    U2, X2, th = results_io.load_iodata(p)
    assert np.array_equal(U2, U) and np.array_equal(X2, X) and th.shape == (r,)


def test_load_demos_reads_the_reference_layout():
    """K2 fixture cross-check: demos loaded through results_io from a file in the reference's layout equal the
    golden arrays extracted from the shipped pendulum_demos.mat (only where the reference tree is mounted)."""
    ref = "/root/reference/Examples/IRL/pendulum/data/pendulum_demos.mat"
    if not os.path.isfile(ref):
        import pytest
        pytest.skip("reference tree not mounted")
    g2 = np.load(os.path.join(G, "k2_demos.npz"))
    trajs, dt, theta = results_io.load_demos(ref)
    assert len(trajs) == int(g2["pendulum_n"]) and abs(dt - float(g2["pendulum_dt"][0])) < 1e-15
    assert np.array_equal(theta, g2["pendulum_true_parameter"].ravel())
    for i, t in enumerate(trajs):
        assert np.array_equal(t["state_traj_opt"], g2["pendulum_%d_X" % i])
        assert np.array_equal(t["control_traj_opt"], g2["pendulum_%d_U" % i])
        assert np.array_equal(t["costate_traj_opt"], g2["pendulum_%d_L" % i])
