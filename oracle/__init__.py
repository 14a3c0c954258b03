"""CPU oracle for the PDP hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product path (``PDP``, ``JinEnv``,
``pontryagin_differentiable_programming_b200``) never does and fails loudly without CUDA.

Parity status: PINNED.  The restatement is checked against the golden vectors the reference
ships (``Examples/**/data/*.mat``, extracted to ``tests/golden/*.npz`` by
``tests/golden/make_golden.py``) and against the reference's own NumPy code
(``PDP.LQR.lqrSolver`` etc.) imported unmodified under a ``casadi`` stub in the build
container (``oracle/ref_loader.py``); see ``tests/test_oracle_golden.py``.
"""
