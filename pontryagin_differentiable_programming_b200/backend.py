"""ctypes binding of ``libpdp_b200.so`` (the C ABI declared in ``include/pdp_b200.h``).

There is deliberately NO CPU fallback: if the library or a CUDA device is missing, hot-path calls
raise ``PDPBackendError``.
"""
import ctypes
import os
import threading

from . import build

c_dp = ctypes.c_void_p  # device pointers are passed as integers


class PDPBackendError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()

OP_AUX_LQR, OP_SWEEP, OP_SWEEP_HOST, OP_ROLLOUT_HOST, OP_SENS_HOST = 1, 2, 3, 4, 5
KIND_OC, KIND_SYSID, KIND_CP, KIND_LQR = 1, 2, 3, 4

EXPORTS = ["pdp_load_system", "pdp_free_system", "pdp_system_dims", "pdp_last_error", "pdp_version",
           "pdp_workspace_bytes", "pdp_rollout_costate", "pdp_aux_lqr", "pdp_sweep", "pdp_aux_eval",
           "pdp_sens_fwd", "pdp_sweep_host", "pdp_lqr_dense", "pdp_eval_function", "pdp_aux_lqr_backward",
           "pdp_aux_lqr_forward", "pdp_rollout_feedback", "pdp_set_sweep_parts", "pdp_sweep_host_traj",
           "pdp_reduce_workspace_bytes", "pdp_reduce_loss_dp", "pdp_rollout_costate_host", "pdp_sens_fwd_host"]


def library_path():
    return build.LIB_PATH


def load_library(build_if_missing=True):
    """dlopen libpdp_b200.so (building it in-tree with nvcc first if needed) and declare prototypes."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = build.LIB_PATH
        if not os.path.isfile(path) and not build_if_missing:
            raise PDPBackendError("libpdp_b200.so not built (run __graft_entry__.build())")
        if build_if_missing:
            build.build_library()      # no-op when the library is newer than csrc/pdp_b200.cu and include/pdp_b200.h
        lib = ctypes.CDLL(path)
        i, sz, vp, dp = ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, c_dp
        lib.pdp_load_system.argtypes = [ctypes.c_char_p, ctypes.POINTER(vp)]
        lib.pdp_load_system.restype = i
        lib.pdp_free_system.argtypes = [vp]
        lib.pdp_free_system.restype = None
        lib.pdp_system_dims.argtypes = [vp, ctypes.POINTER(i)]
        lib.pdp_system_dims.restype = i
        lib.pdp_last_error.restype = ctypes.c_char_p
        lib.pdp_version.restype = ctypes.c_char_p
        lib.pdp_workspace_bytes.argtypes = [vp, i, i, i]
        lib.pdp_workspace_bytes.restype = sz
        lib.pdp_rollout_costate.argtypes = [vp, i, i, dp, dp, i, dp, dp, dp, dp, dp, dp, vp]
        lib.pdp_rollout_costate.restype = i
        lib.pdp_aux_lqr.argtypes = [vp, i, i, dp, dp, dp, dp, i, dp, i, dp, dp, dp, dp, dp, dp, sz, dp, vp]
        lib.pdp_aux_lqr.restype = i
        lib.pdp_lqr_dense.argtypes = [vp, i, i, dp, dp, dp, i, dp, dp, i, dp, sz, dp, vp]
        lib.pdp_eval_function.argtypes = [vp, i, ctypes.POINTER(dp), ctypes.POINTER(i), ctypes.POINTER(dp), vp]
        lib.pdp_eval_function.restype = i
        lib.pdp_lqr_dense.restype = i
        lib.pdp_rollout_feedback.argtypes = [vp, i, i, dp, dp, i, dp, dp, dp, dp, dp, dp, dp, dp, dp, i, dp, vp]
        lib.pdp_rollout_feedback.restype = i
        lib.pdp_aux_lqr_backward.argtypes = [vp, i, i, dp, dp, dp, dp, i, dp, sz, dp, vp]
        lib.pdp_aux_lqr_backward.restype = i
        lib.pdp_aux_lqr_forward.argtypes = [vp, i, i, dp, dp, dp, i, dp, i, dp, dp, dp, dp, dp, dp, sz, dp, vp]
        lib.pdp_aux_lqr_forward.restype = i
        lib.pdp_sweep.argtypes = [vp, i, i, dp, dp, i, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, sz, dp, vp]
        lib.pdp_sweep.restype = i
        lib.pdp_set_sweep_parts.argtypes = [vp, i]
        lib.pdp_set_sweep_parts.restype = i
        lib.pdp_aux_eval.argtypes = [vp, i, i, dp, dp, dp, dp, i, dp, dp, vp]
        lib.pdp_aux_eval.restype = i
        lib.pdp_sens_fwd.argtypes = [vp, i, i, dp, dp, i, dp, dp, dp, dp, dp, dp, dp, dp, vp]
        lib.pdp_sens_fwd.restype = i
        lib.pdp_sweep_host.argtypes = [vp, i, i, dp, dp, i, dp, dp, dp, dp, dp, i, dp, sz, vp]
        lib.pdp_sweep_host.restype = i
        lib.pdp_sweep_host_traj.argtypes = [vp, i, i, dp, dp, i, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, sz, vp]
        lib.pdp_sweep_host_traj.restype = i
        lib.pdp_reduce_workspace_bytes.argtypes = [i]
        lib.pdp_reduce_workspace_bytes.restype = sz
        lib.pdp_reduce_loss_dp.argtypes = [i, i, dp, dp, dp, sz, vp]
        lib.pdp_reduce_loss_dp.restype = i
        lib.pdp_rollout_costate_host.argtypes = [vp, i, i, dp, dp, i, dp, dp, dp, dp, dp, dp, sz, vp]
        lib.pdp_rollout_costate_host.restype = i
        lib.pdp_sens_fwd_host.argtypes = [vp, i, i, dp, dp, i, dp, dp, dp, dp, dp, sz, vp]
        lib.pdp_sens_fwd_host.restype = i
        _lib = lib
        return lib


def check(code, what=""):
    if code != 0:
        msg = load_library().pdp_last_error().decode(errors="replace")
        raise PDPBackendError("%s failed with code %d: %s" % (what or "pdp call", code, msg))


class SystemHandle:
    """Owns a ``pdp_system_t*``."""

    def __init__(self, module_path: str):
        self.lib = load_library()
        h = ctypes.c_void_p()
        check(self.lib.pdp_load_system(module_path.encode(), ctypes.byref(h)), "pdp_load_system")
        self.ptr = h
        dims = (ctypes.c_int * 4)()
        check(self.lib.pdp_system_dims(self.ptr, dims), "pdp_system_dims")
        self.kind, self.n, self.m, self.r = (int(v) for v in dims)
        self.module_path = module_path

    def workspace_bytes(self, op, B, H):
        return int(self.lib.pdp_workspace_bytes(self.ptr, op, B, H))

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                self.lib.pdp_free_system(self.ptr)
                self.ptr = None
        except Exception:
            pass
