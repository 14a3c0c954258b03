"""Import the UNMODIFIED reference ``PDP/PDP.py`` under a stub ``casadi`` (TEST INFRASTRUCTURE).

Only works where ``/root/reference`` exists (the build container); nothing that runs on the GPU
box may call this.  The stub makes ``from casadi import *`` succeed so the pure-NumPy half of the
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


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "PDP", "PDP.py"))


def load_reference_pdp():
    if not reference_available():
        raise FileNotFoundError("reference tree not present at %s" % REFERENCE_ROOT)
    import numpy
    saved = sys.modules.get("casadi")
    stub = types.ModuleType("casadi")
    stub.np = numpy
    stub.__all__ = ["np"]
    sys.modules["casadi"] = stub
    try:
        spec = importlib.util.spec_from_file_location("_reference_PDP", os.path.join(REFERENCE_ROOT, "PDP", "PDP.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            del sys.modules["casadi"]
        else:
            sys.modules["casadi"] = saved
    return mod
