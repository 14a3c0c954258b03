#!/bin/bash
# compute-sanitizer passes over one small quadrotor sweep (run on the GPU box): memcheck + racecheck on the
# shared-memory choreography of the warp-per-trajectory kernels.  Output -> gpurun_out/sanitizer_*.log
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/pdp_sanitize_case.py <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, '.')
import bench
from pontryagin_differentiable_programming_b200 import systems
dev = torch.device('cuda:0')
s = systems.quadrotor_irl(0.1)
x0, th, U, Xr, Ur = [torch.as_tensor(np.ascontiguousarray(a), device=dev) for a in bench.synth_quadrotor(10, 19, seed=4)]
r = s.sweep(x0, th, U, Xref=Xr, Uref=Ur)
torch.cuda.synchronize()
print('sweep ok', float(r['loss_dp'][0, 0]))
PY
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/pdp_sanitize_case.py > gpurun_out/sanitizer_$tool.log 2>&1
  tail -3 gpurun_out/sanitizer_$tool.log
done
