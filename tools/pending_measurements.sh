#!/bin/bash
# Measurements prepared at the end of round 1 but not taken (the round's GPU budget was spent): run on the GPU box,
# everything lands in gpurun_out/.
#   gpurun --timeout 900 -- 'bash tools/pending_measurements.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# 1. fused backward+forward kernel (module option fused=1) and streaming dX/dU stores against the shipped two-kernel
#    path: compare the both_ms column (one aux-LQR call) -- DESIGN.md section 9, item 1
python tools/tune_aux_lqr.py --run 2>&1 | tail -6 > gpurun_out/pending_tune_fused.log
cp gpurun_out/tune_aux_lqr.json gpurun_out/pending_tune_fused.json 2>/dev/null
# 1b. the same unmeasured kernels under compute-sanitizer (racecheck + memcheck) and against the shipped path, BEFORE
#     anything is adopted: small batch with a tail (odd count, last block mostly idle lanes)
cat > /tmp/pdp_new_kernels_case.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from tools.tune_aux_lqr import make
dev = torch.device('cuda:0')
x0, th, U, Xr, Ur = [torch.as_tensor(np.ascontiguousarray(a), device=dev) for a in bench.synth_quadrotor(37, 19, seed=4)]
ref = make().sweep(x0, th, U, Xref=Xr, Uref=Ur)
new = make(fused=1, stream_out=1, rollout_parts=4, stage_inputs=1, fwd_stage_inputs=1).sweep(x0, th, U, Xref=Xr, Uref=Ur)
torch.cuda.synchronize()
for k in ("X", "Lam", "cost", "dX", "dU", "loss_dp"):
    d = float((ref[k] - new[k]).abs().max() / ref[k].abs().max())
    print(k, "max rel diff fused/multi-warp vs shipped:", d)
    assert d < 1e-11, k
print("new kernels ok")
PY
python /tmp/pdp_new_kernels_case.py > gpurun_out/pending_new_kernels_parity.log 2>&1
for tool in racecheck memcheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/pdp_new_kernels_case.py > gpurun_out/pending_sanitizer_$tool.log 2>&1
  tail -2 gpurun_out/pending_sanitizer_$tool.log
done
# 2. first ncu capture of the SysID / ControlPlanning sensitivity kernel (C5 at its per-GPU size) -- item 5
cat > /tmp/pdp_c5_case.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from pontryagin_differentiable_programming_b200 import systems
dev = torch.device('cuda:0')
s = systems.quadrotor_sysid(0.1)
B, H = 32768, 100
g = torch.Generator().manual_seed(0)
inputs = (20 * torch.rand((B, H, 4), dtype=torch.float64, generator=g) - 10).to(dev)
x0 = torch.tensor([-8, -6, 9., 0, 0, 0, 1, 0, 0, 0, 0, 0, 0], dtype=torch.float64).repeat(B, 1).to(dev)
th_true = torch.tensor([1, 1, 1, 1, 0.4], dtype=torch.float64, device=dev)
Xobs = s.step(inputs, None, th_true, x0=x0, want_traj=True)["X"]
for _ in range(4):
    s.step(inputs, Xobs, th_true * 1.1)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:pdp_k_sens_fwd -s 3 -c 1 -o gpurun_out/prof_sens -f \
    python /tmp/pdp_c5_case.py > gpurun_out/pending_ncu_sens.log 2>&1
# 3. launch list + full capture of the shipped (pipelined) bench step
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
ls -la gpurun_out | tail -12
