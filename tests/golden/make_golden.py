"""Generate the committed golden fixtures in tests/golden/ (run once in the build container).

Sources (all under /root/reference, read-only, never copied as source):
  K1  Examples/SysID/*/data/*_iodata.mat            -> k1_iodata.npz
  K2  Examples/IRL/*/data/*_demos.mat               -> k2_demos.npz
  K3  Examples/IRL/{pendulum,quadrotor}/data/PDP_results_trial_*.mat (a few iterations) -> k3_irl_traces.npz
  K4  Examples/OC/rocket/data/PDP_OC_results_trial_0.mat (a few iterations) -> k4_rocket_oc.npz
  K5  Examples/OC/{cartpole,robotarm}/data/PDP_Neural_trial_0.mat -> k5_neural.npz
  K6  the reference's own LQR.lqrSolver / integrateAuxSys (imported unmodified under a casadi
      stub, oracle/ref_loader.py) run on auxiliary systems evaluated by the oracle -> k6_reference_lqr.npz
  schema of the shipped result / demo / iodata .mat files -> schema_mat.json  (--schema: only this one)

Usage:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import scipy.io as sio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
EX = "/root/reference/Examples"


def k1():
    out = {}
    for env, fname in [("pendulum", "pendulum"), ("cartpole", "cartpole"), ("robotarm", "robotarm"),
                       ("quadrotor", "uav"), ("rocket", "rocket")]:
        d = sio.loadmat("%s/SysID/%s/data/%s_iodata.mat" % (EX, env, fname))
        key = [k for k in d if not k.startswith("__")][0]
        s = d[key][0, 0]
        out[env + "_inputs"] = np.asarray(s["batch_inputs"], dtype=np.float64)
        out[env + "_states"] = np.asarray(s["batch_states"], dtype=np.float64)
        out[env + "_true_parameter"] = np.asarray(s["true_parameter"], dtype=np.float64).ravel()
    np.savez_compressed(os.path.join(HERE, "k1_iodata.npz"), **out)


def k2():
    out = {}
    for env, fname in [("pendulum", "pendulum"), ("cartpole", "cartpole"), ("robotarm", "robotarm"),
                       ("quadrotor", "uav"), ("rocket", "rocket")]:
        d = sio.loadmat("%s/IRL/%s/data/%s_demos.mat" % (EX, env, fname))
        tr = d["trajectories"]
        nd = tr.shape[1]
        out[env + "_n"] = np.array(nd)
        out[env + "_dt"] = np.asarray(d["dt"], dtype=np.float64).ravel()
        out[env + "_true_parameter"] = np.asarray(d["true_parameter"], dtype=np.float64).ravel()
        for i in range(nd):
            s = tr[0, i]
            out["%s_%d_X" % (env, i)] = np.asarray(s["state_traj_opt"][0, 0], dtype=np.float64)
            out["%s_%d_U" % (env, i)] = np.asarray(s["control_traj_opt"][0, 0], dtype=np.float64)
            out["%s_%d_L" % (env, i)] = np.asarray(s["costate_traj_opt"][0, 0], dtype=np.float64)
            out["%s_%d_cost" % (env, i)] = np.asarray(s["cost"][0, 0], dtype=np.float64).ravel()
    np.savez_compressed(os.path.join(HERE, "k2_demos.npz"), **out)


def k3():
    out = {}
    picks = {"pendulum": ([0, 1, 2], [0, 1, 500, 5000]), "quadrotor": ([0, 3], [0, 1, 2000, 6000]),
             "cartpole": ([0], [3000, 8000]), "robotarm": ([0], [0, 3000]), "rocket": ([0], [0, 2000])}
    for env, (trials, iters) in picks.items():
        for j in trials:
            r = sio.loadmat("%s/IRL/%s/data/PDP_results_trial_%d.mat" % (EX, env, j))["results"][0, 0]
            P = np.asarray(r["parameter_trace"], dtype=np.float64)
            P = P.reshape(P.shape[0], -1)
            Lt = np.asarray(r["loss_trace"], dtype=np.float64).ravel()
            out["%s_%d_lr" % (env, j)] = np.asarray(r["learning_rate"], dtype=np.float64).ravel()
            its = [k for k in iters if k + 1 < P.shape[0]]
            out["%s_%d_iters" % (env, j)] = np.array(its)
            # parameter_trace[k] is theta AFTER iteration k; loss_trace[k+1], dp_{k+1} are evaluated at it
            out["%s_%d_theta" % (env, j)] = np.stack([P[k] for k in its])
            out["%s_%d_theta_next" % (env, j)] = np.stack([P[k + 1] for k in its])
            out["%s_%d_loss" % (env, j)] = np.array([Lt[k + 1] for k in its])
    np.savez_compressed(os.path.join(HERE, "k3_irl_traces.npz"), **out)


def k4():
    r = sio.loadmat("%s/OC/rocket/data/PDP_OC_results_trial_0.mat" % EX)["results"][0, 0]
    P = np.asarray(r["parameter_trace"], dtype=np.float64)
    Lt = np.asarray(r["loss_trace"], dtype=np.float64).ravel()
    its = [0, 1, 1000, 49999]
    sol, tsol = r["solved_solution"][0, 0], r["true_solution"][0, 0]
    np.savez_compressed(
        os.path.join(HERE, "k4_rocket_oc.npz"),
        iters=np.array(its), U=np.stack([P[k] for k in its]), U_next=np.stack([P[k + 1] for k in its]),
        loss=np.array([Lt[k] for k in its]), lr=np.asarray(r["learning_rate"], dtype=np.float64).ravel(),
        dt=np.asarray(r["dt"], dtype=np.float64).ravel(), horizon=np.asarray(r["horizon"]).ravel(),
        solved_X=np.asarray(sol["state_traj"], dtype=np.float64), solved_U=np.asarray(sol["control_traj"], dtype=np.float64),
        solved_cost=np.asarray(sol["cost"], dtype=np.float64).ravel(),
        true_X=np.asarray(tsol["state_traj_opt"], dtype=np.float64), true_U=np.asarray(tsol["control_traj_opt"], dtype=np.float64),
        true_L=np.asarray(tsol["costate_traj_opt"], dtype=np.float64), true_cost=np.asarray(tsol["cost"], dtype=np.float64).ravel())


def k5():
    out = {}
    for env in ("cartpole", "robotarm"):
        r = sio.loadmat("%s/OC/%s/data/PDP_Neural_trial_0.mat" % (EX, env))["results"][0, 0]
        P = np.asarray(r["parameter_trace"], dtype=np.float64)
        sol = r["solved_solution"][0, 0]
        out[env + "_theta"] = P[-1]
        out[env + "_X"] = np.asarray(sol["state_traj"], dtype=np.float64)
        out[env + "_U"] = np.asarray(sol["control_traj"], dtype=np.float64)
        out[env + "_cost"] = np.asarray(sol["cost"], dtype=np.float64).ravel()
        out[env + "_dt"] = np.asarray(r["dt"], dtype=np.float64).ravel()
        out[env + "_horizon"] = np.asarray(r["horizon"]).ravel()
        if env in r.dtype.names:
            s = r[env][0, 0]
            for nm in s.dtype.names:
                out["%s_param_%s" % (env, nm)] = np.asarray(s[nm], dtype=np.float64).ravel()
    np.savez_compressed(os.path.join(HERE, "k5_neural.npz"), **out)


def k6():
    """Reference NumPy code run unmodified on oracle-evaluated auxiliary systems."""
    from oracle import envs, pdp_oracle, ref_loader
    ref = ref_loader.load_reference_pdp()
    g2 = np.load(os.path.join(HERE, "k2_demos.npz"))
    out = {}
    cfgs = {"quadrotor": dict(builder=envs.quadrotor, kw=dict(c=0.01, wthrust=0.1)),
            "pendulum": dict(builder=envs.pendulum, kw=dict())}
    for env, cfg in cfgs.items():
        e = cfg["builder"](**cfg["kw"])
        oc = pdp_oracle.build_oc(e, float(g2[env + "_dt"][0]))
        theta = g2[env + "_true_parameter"] * 1.1  # off the optimum on purpose
        X, U, L = g2[env + "_0_X"], g2[env + "_0_U"], g2[env + "_0_L"]
        aux = oc.getAuxSys(X, U, L, theta)
        H = U.shape[0]
        lqr = ref.LQR()
        lqr.setDyn(dynF=aux["dynF"], dynG=aux["dynG"], dynE=aux["dynE"])
        lqr.setPathCost(Hxx=aux["Hxx"], Huu=aux["Huu"], Hxu=aux["Hxu"], Hux=aux["Hux"], Hxe=aux["Hxe"], Hue=aux["Hue"])
        lqr.setFinalCost(hxx=aux["hxx"], hxe=aux["hxe"])
        sol = lqr.lqrSolver(np.zeros((oc.n, oc.r)), H)
        out[env + "_theta"] = theta
        for k in ("dynF", "dynG", "dynE", "Hxx", "Hxu", "Hxe", "Hux", "Huu", "Hue", "hxx", "hxe"):
            out["%s_%s" % (env, k)] = np.stack(aux[k])
        out[env + "_dX"] = np.stack(sol["state_traj_opt"])
        out[env + "_dU"] = np.stack(sol["control_traj_opt"])
        out[env + "_dL"] = np.stack(sol["costate_traj_opt"])
    # forward-sensitivity recursions of ControlPlanning / SysID on random small systems
    rng = np.random.default_rng(7)
    H, n, m, r = 12, 5, 2, 4
    F = [rng.standard_normal((n, n)) * 0.4 for _ in range(H)]
    G = [rng.standard_normal((n, m)) for _ in range(H)]
    Ux = [rng.standard_normal((m, n)) * 0.3 for _ in range(H)]
    Ue = [rng.standard_normal((m, r)) for _ in range(H)]
    E = [rng.standard_normal((n, r)) for _ in range(H)]
    cp = ref.ControlPlanning().integrateAuxSys(F, G, Ux, Ue, np.zeros((n, r)))
    sid = ref.SysID().integrateAuxSys(F, E, np.zeros((n, r)))
    out.update(fs_F=np.stack(F), fs_G=np.stack(G), fs_Ux=np.stack(Ux), fs_Ue=np.stack(Ue), fs_E=np.stack(E),
               fs_cp_X=np.stack(cp["state_traj"]), fs_cp_U=np.stack(cp["control_traj"]),
               fs_sysid_X=np.stack(sid["state_traj"]))
    np.savez_compressed(os.path.join(HERE, "k6_reference_lqr.npz"), **out)




def schema():
    """Field names / dtypes / shapes of the `.mat` files the reference ships (first K entries only where a dimension
    is the iteration count) -> schema_mat.json; pins pontryagin_differentiable_programming_b200/results_io.py."""
    import json

    def describe(path):
        d = sio.loadmat(path)
        out = {}
        for k, v in d.items():
            if k.startswith("__"):
                continue
            ent = {"dtype": str(v.dtype) if not v.dtype.names else "struct", "shape": list(v.shape)}
            if v.dtype.names:
                s = v[0, 0]
                ent["fields"] = {n: {"dtype": str(s[n].dtype) if not s[n].dtype.names else "struct", "ndim": int(np.asarray(s[n]).ndim)}
                                 for n in v.dtype.names}
            elif v.dtype == object:
                s = v[0, 0]
                s = s[0, 0] if s.dtype.names and s.shape == (1, 1) else s
                ent["fields"] = {n: {"dtype": str(np.asarray(s[n]).dtype), "ndim": int(np.asarray(s[n]).ndim)} for n in s.dtype.names}
            out[k] = ent
        return out
    files = {"irl_results": "IRL/pendulum/data/PDP_results_trial_0.mat", "irl_demos": "IRL/pendulum/data/pendulum_demos.mat",
             "sysid_results": "SysID/quadrotor/data/PDP_SysID_results_trial_0.mat", "sysid_iodata": "SysID/quadrotor/data/uav_iodata.mat",
             "oc_results": "OC/rocket/data/PDP_OC_results_trial_0.mat"}
    out = {}
    for name, rel in files.items():
        p = os.path.join(EX, rel)
        if os.path.isfile(p):
            out[name] = describe(p)
    with open(os.path.join(HERE, "schema_mat.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    if "--schema" in sys.argv:          # only the .mat schema fixture
        schema()
    else:
        k1(); k2(); k3(); k4(); k5(); k6(); schema()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz") or f.endswith(".json"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
