"""Stage a few UNMODIFIED reference Examples scripts and the small data files they read under baseline/_ref/Examples
(git-ignored, but it travels to the GPU box with the gpurun snapshot) so that tests/test_gpu_dropin_run.py can execute them
on real hardware against this repo's drop-in PDP / JinEnv / casadi packages.  Build container only: needs /root/reference.

  python tools/stage_reference_examples.py            # idempotent; prints what it staged
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/Examples"
DST = os.path.join(ROOT, "baseline", "_ref", "Examples")
FILES = [
    "IRL/pendulum/pendulum_PDP.py", "IRL/pendulum/data/pendulum_demos.mat", "IRL/pendulum/data/PDP_results_trial_0.mat",
    "IRL/quadrotor/uav_PDP.py", "IRL/quadrotor/data/uav_demos.mat",
    "SysID/quadrotor/uav_PDP.py", "SysID/quadrotor/data/uav_iodata.mat",
    "OC/cartpole/cartpole_PDP_poly.py",
    "OC/rocket/rocket_PDP_Recmat.py",
]


def stage(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print("reference tree not present; nothing staged")
        return []
    done = []
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.isfile(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.isfile(dst) or os.path.getsize(dst) != os.path.getsize(src):
            shutil.copyfile(src, dst)
        done.append(rel)
    for d in {os.path.dirname(os.path.join(DST, rel)) for rel in FILES if rel.endswith(".py")}:
        os.makedirs(os.path.join(d, "data"), exist_ok=True)      # the scripts save into ./data at the end of a trial
    if verbose:
        print("staged %d files under %s" % (len(done), DST))
    return done


if __name__ == "__main__":
    stage()
