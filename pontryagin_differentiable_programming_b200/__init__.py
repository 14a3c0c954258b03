"""B200-native batched Pontryagin Differentiable Programming engine (hot path only).

Sub-modules: ``symbolic`` (expression DAG / CasADi-subset front-end used purely for code
generation), ``codegen`` (expressions -> sm_100a CUDA translation unit), ``backend`` (ctypes
binding of the C-ABI in ``include/pdp_b200.h``), ``engine`` (batched entry points on torch
CUDA float64 tensors).  The reference-facing class surface lives in the top-level drop-in
packages ``PDP`` and ``JinEnv``.
"""
__version__ = "0.1.0"
