"""Import the UNMODIFIED reference ``PDP/PDP.py`` under a stub ``casadi`` (TEST INFRASTRUCTURE).

Source of the file, in this order: ``/root/reference/PDP/PDP.py`` (the build container) or the byte-identical copy that
``oracle/stage_reference.py`` (the committed recipe, run by ``__graft_entry__.build()``) puts under ``baseline/_ref/reference_src/`` (next to the staged Examples scripts) --
git-ignored, so no reference source enters the history, but it travels to the GPU box with the snapshot so that
``bench.py``'s CPU arms can time the reference's OWN NumPy half there (``cpu_baseline.kind`` "reference").
The stub makes ``from casadi import *`` succeed so the pure-NumPy half of the
reference -- ``LQR.*`` (PDP.py:334-615), ``ControlPlanning.integrateAuxSys`` (:813-838),
``SysID.integrateAuxSys`` (:1241-1259) -- runs exactly as shipped.  Used by
``tests/golden/make_golden.py`` to generate K6 fixtures and by ``tests/test_oracle_golden.py``
(skipped when the reference tree is absent).
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"
STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "reference_src", "PDP.py")
_MOD = None


def reference_path():
    for p in (os.path.join(REFERENCE_ROOT, "PDP", "PDP.py"), STAGED):
        if os.path.isfile(p):
            return p
    return None


def reference_available():
    return reference_path() is not None


def load_reference_pdp():
    global _MOD
    if _MOD is not None:
        return _MOD
    path = reference_path()
    if path is None:
        raise FileNotFoundError("reference PDP.py neither under %s nor staged at %s" % (REFERENCE_ROOT, STAGED))
    import numpy
    saved = sys.modules.get("casadi")
    stub = types.ModuleType("casadi")
    stub.np = numpy
    stub.__all__ = ["np"]
    sys.modules["casadi"] = stub
    try:
        spec = importlib.util.spec_from_file_location("_reference_PDP", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            del sys.modules["casadi"]
        else:
            sys.modules["casadi"] = saved
    _MOD = mod
    return mod


def reference_lqr_solver(aux, ini_state, horizon):
    """The reference's own calling sequence for the auxiliary control system (Examples/IRL/quadrotor/uav_PDP.py:56-63):
    unmodified ``LQR.setDyn / setPathCost / setFinalCost / lqrSolver`` on a ``getAuxSys``-shaped dict of lists."""
    lqr = load_reference_pdp().LQR()
    lqr.setDyn(dynF=aux["dynF"], dynG=aux["dynG"], dynE=aux["dynE"])
    lqr.setPathCost(Hxx=aux["Hxx"], Huu=aux["Huu"], Hxu=aux["Hxu"], Hux=aux["Hux"], Hxe=aux["Hxe"], Hue=aux["Hue"])
    lqr.setFinalCost(hxx=aux["hxx"], hxe=aux["hxe"])
    return lqr.lqrSolver(ini_state, horizon)
