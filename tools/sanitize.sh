#!/bin/bash
# compute-sanitizer passes (run on the GPU box): memcheck + racecheck over small cases of every kernel on the hot path --
# the pipelined sweep with an odd batch cut into 3 sub-batches on the library's side streams (round-1 advisor finding), the
# batch-reduction kernel, the SysID / ControlPlanning sensitivity kernel with column groups, the adjoint rollout and the
# host-buffer entry points.  Output -> gpurun_out/sanitizer_*.log
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/pdp_sanitize_case.py <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, '.')
import bench
from pontryagin_differentiable_programming_b200 import engine, systems
dev = torch.device('cuda:0')
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
# --- IRL sweep, odd batch, 3 sub-batches on the side streams, then the unsplit call: identical
s = systems.quadrotor_irl(0.1)
host = bench.synth_quadrotor(37, 19, seed=4)
x0, th, U, Xr, Ur = [t(a) for a in host]
s.set_sweep_parts(3)
r3 = s.sweep(x0, th, U, Xref=Xr, Uref=Ur)
s.set_sweep_parts(1)
r1 = s.sweep(x0, th, U, Xref=Xr, Uref=Ur)
s.set_sweep_parts(0)
torch.cuda.synchronize()
assert all(torch.equal(r3[k], r1[k]) for k in ("X", "Lam", "dX", "dU", "loss_dp"))
sums = engine.reduce_loss_dp(r1["loss_dp"])
assert torch.allclose(sums[:-1], r1["loss_dp"].sum(dim=0), rtol=1e-12)
ldp = torch.empty((37, 10), dtype=torch.float64).pin_memory()
s.sweep_host(*[pin(a) for a in host], ldp, n_chunks=3)
torch.cuda.synchronize()
assert torch.equal(ldp, r1["loss_dp"].cpu())
print('sweep ok', float(r1['loss_dp'][0, 0]))
# --- SysID step (two column groups), with trajectories and sensitivities
inputs, x0s, th_true, theta = bench.synth_sysid(70, 12, seed=1)
sid = systems.quadrotor_sysid(0.1)
Xobs = sid.step(t(inputs), None, t(th_true), x0=t(x0s), want_traj=True)["X"]
o = sid.step(t(inputs), Xobs, t(theta), want_traj=True, want_sens=True)
print('sysid ok', float(engine.reduce_loss_dp(o["loss_dp"])[0]))
# --- ControlPlanning step, adjoint rollout
cp = systems.cartpole_cp("poly", 50, 0.05)
x0c, thc = bench.synth_cartpole(45, cp.r, seed=2)
o = cp.step(t(x0c), 50, t(thc), want_traj=True, want_sens=True)
x0r, Ur = bench.synth_rocket(33, 21, seed=3)
ro = systems.rocket_oc_adjoint(0.1)
o = ro.rollout_costate(t(x0r), torch.zeros((1, 1), dtype=torch.float64, device=dev), t(Ur), want_dHu=True)
torch.cuda.synchronize()
print('cp / adjoint ok', float(o["cost"][0]))
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/pdp_sanitize_case.py > gpurun_out/sanitizer_$tool.log 2>&1
  tail -3 gpurun_out/sanitizer_$tool.log
done
