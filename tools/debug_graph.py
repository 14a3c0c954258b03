import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pontryagin_differentiable_programming_b200 import irl, systems, ocsolver, distributed
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
dev = torch.device("cuda:0")
g2 = np.load(os.path.join(G, "k2_demos.npz")); g3 = np.load(os.path.join(G, "k3_irl_traces.npz"))
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
sys_ = systems.quadrotor_irl(float(g2["quadrotor_dt"][0]))
Xd = t(np.stack([g2["quadrotor_%d_X" % i] for i in range(2)])); Ud = t(np.stack([g2["quadrotor_%d_U" % i] for i in range(2)]))
theta = t(g3["quadrotor_0_theta"][0]); th = theta.reshape(1, -1)
lr = float(g3["quadrotor_0_lr"][0])
tr = irl.IRLTrainer(sys_, Xd, Ud, lr)
x0 = tr.x0
print("eager step loss", tr.step(theta)[0].item())
sol = ocsolver.solve(sys_, x0, 50, th)

def capture(fn):
    side = torch.cuda.Stream(device=dev); side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(2): fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    return g, out

# (a) sweep alone
Uc = sol["U"].clone()
def fa():
    res = sys_.sweep(x0, th, Uc, Xref=Xd, Uref=Ud, want_traj=False)
    return distributed.reduce_loss_dp(res["loss_dp"], None)
print("sweep eager", fa()[0].item())
g, out = capture(fa); g.replay(); torch.cuda.synchronize(); print("sweep graph", out[0].item())
# (b) rollout_costate alone
def fb():
    return sys_.rollout_costate(x0, th.expand(2, -1).contiguous(), Uc, want_dHu=True)
e = fb(); g, out = capture(fb); g.replay(); torch.cuda.synchronize()
print("rollout graph dcost", (out["cost"] - e["cost"]).abs().max().item(), "dX", (out["X"] - e["X"]).abs().max().item(), "dHu", (out["dHu"] - e["dHu"]).abs().max().item())
# (c) solve_fixed alone
st = ocsolver.FixedSolverState(2, 50, 4, dev); st.U.copy_(sol["U"])
def fc():
    return ocsolver.solve_fixed(sys_, x0, 50, th, st, n_iter=3)
e = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in fc().items()}
g, out = capture(fc); g.replay(); torch.cuda.synchronize()
print("solve_fixed graph: cost", out["cost"].tolist(), "eager", e["cost"].tolist(), "dU", (out["U"] - sol["U"]).abs().max().item(), "gnorm", out["grad_norm"].tolist(), "s", st.s_newton.tolist(), "mu", st.mu.tolist())
# (d) whole iteration
tr2 = irl.IRLTrainer(sys_, Xd, Ud, lr)
o = tr2.step_graph(theta); torch.cuda.synchronize()
print("step_graph loss", o[0].item(), "resid", o[2].item(), "dU vs sol", (tr2._fixed_state.U - sol["U"]).abs().max().item())
o = tr2.step_graph(theta); torch.cuda.synchronize()
print("step_graph again loss", o[0].item())
le, _, _ = tr2._iteration_fixed(theta, 3)
print("iteration_fixed eager loss", le.item())
