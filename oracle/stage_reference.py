"""Recipe that stages the UNMODIFIED reference ``PDP/PDP.py`` under ``baseline/_ref/reference_src/`` (TEST INFRASTRUCTURE, build container
only -- needs /root/reference).  ``baseline/_ref/`` is git-ignored (no reference source in the history) but not
gpurun-ignored, so the copy travels to the GPU box, where ``oracle/ref_loader.py`` imports it under a ``casadi`` stub and
``bench.py``'s CPU arms time the reference's own NumPy half (``LQR.lqrSolver`` PDP.py:446-615, ``SysID.integrateAuxSys``
:1241-1259, ``ControlPlanning.integrateAuxSys`` :813-838) next to the oracle's restatement of the CasADi half.

  python oracle/stage_reference.py        # idempotent
"""
import filecmp
import os
import shutil

SRC = "/root/reference/PDP/PDP.py"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "reference_src", "PDP.py")


def stage(verbose=True):
    if not os.path.isfile(SRC):
        if verbose:
            print("reference tree not present; nothing staged")
        return False
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    if not (os.path.isfile(DST) and filecmp.cmp(SRC, DST, shallow=False)):
        tmp = DST + ".tmp%d" % os.getpid()
        shutil.copyfile(SRC, tmp)
        os.replace(tmp, DST)
    if verbose:
        print("staged %s" % DST)
    return True


if __name__ == "__main__":
    stage()
