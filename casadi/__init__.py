"""``casadi`` name-compatible front-end for the reference's scripts (``from casadi import *``).

The reference's Examples build their models with CasADi ``SX`` (e.g. reference
``Examples/IRL/quadrotor/uav_PDP.py:3,20-28``).  CasADi cannot be installed in this image, and
this engine uses symbolic algebra only to code-generate CUDA device functions, so this package
re-exports the engine's own expression front-end under the names those scripts import.
``np`` is exported too because the reference relies on ``from casadi import *`` leaking NumPy
(reference ``PDP/PDP.py:994,1265`` use a bare ``np``).
"""
import numpy as np  # noqa: F401  (leaks through ``import *`` exactly like the real package)
import numpy  # noqa: F401
from pontryagin_differentiable_programming_b200 import symbolic as _sym
from pontryagin_differentiable_programming_b200.symbolic import *  # noqa: F401,F403
from pontryagin_differentiable_programming_b200.symbolic import SX, MX, DM, Function  # noqa: F401

__version__ = "pdp_b200-shim"
__all__ = list(_sym.__all__) + ["np", "numpy"]
