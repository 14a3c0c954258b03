"""CPU warp emulator for the generated Riccati kernels (TEST INFRASTRUCTURE, never on the product path).

The build container has no GPU, so a new kernel variant could only be checked after a `gpurun` round trip.  This tool
compiles the *generated CUDA source* of a module (`OCModuleSource.source()`, cut before the host launchers) with g++
and runs one warp as 32 host threads: `__syncwarp()` is a pthread barrier, `__shfl_xor_sync` goes through a small
exchange buffer, shared memory is a global array.  It executes exactly the statements the GPU executes (same
generated text, same lane predicates, same shared-memory indices), so layout / indexing / synchronisation mistakes
show up here in seconds: the lanes are real concurrent threads, so a missing barrier gives wrong, run-to-run different
results (negative control in tests/test_kernel_emulation.py).  It says nothing about performance, and hazards between
accesses that a CPU barrier orders but a GPU warp does not are left to `tools/sanitize.sh` (compute-sanitizer racecheck
on the GPU).

    from tools import warp_emu
    emu = warp_emu.Emulator(src)                       # src: an OCModuleSource / NewtonModuleSource / LQRModuleSource
    gains, status = emu.backward(X, U, Lam, theta)     # numpy float64 in / out
    dX, dU, loss_dp, status = emu.forward(X, U, theta, gains)
    X, Lam, cost, dHu = emu.rollout(x0, theta, U)      # thread-per-trajectory rollout / costate kernel (threads run one by one)
    emu.backward_dense(aux, term), emu.forward_dense(aux, gains)
    warp_emu.SensEmulator(sens_src).run(...)           # SysID / ControlPlanning sensitivity kernel
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import re
import subprocess
import tempfile

import numpy as np

PREAMBLE = r'''
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <pthread.h>
#include <cmath>
using std::isfinite;
#define __global__
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __shared__ static          /* statically sized tiles: one block runs at a time */
#define __align__(x)
#define __launch_bounds__(...)
struct emu_dim3 { unsigned x, y, z; };
static thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
struct double2 { double x, y; };
static inline double2 make_double2(double a, double b) { double2 v; v.x = a; v.y = b; return v; }
extern "C" { alignas(16) double pdp_smem[1 << 18]; }
static pthread_barrier_t emu_bar;
static double emu_xch[32];
static pthread_barrier_t emu_block_bar;
static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&emu_bar); }
static inline void __syncthreads() { pthread_barrier_wait(&emu_block_bar); }
static inline double __shfl_xor_sync(unsigned, double v, int o) {
  const int l = threadIdx.x & 31;
  emu_xch[l] = v; pthread_barrier_wait(&emu_bar);
  const double r = emu_xch[l ^ o]; pthread_barrier_wait(&emu_bar);
  return r;
}
static inline double __shfl_sync(unsigned, double v, int src) {
  const int l = threadIdx.x & 31;
  emu_xch[l] = v; pthread_barrier_wait(&emu_bar);
  const double r = emu_xch[src & 31]; pthread_barrier_wait(&emu_bar);
  return r;
}
static inline int atomicOr(int* p, int v) { return __sync_fetch_and_or(p, v); }
static inline void __stcs(double* p, double v) { *p = v; }
'''

DRIVER = r'''
struct emu_args {
  int kernel, B, H, theta_stride, x0a_stride;
  const double *X, *U, *Lam, *theta, *X0a, *Xref, *Uref, *auxrec, *termrec;
  double *gains, *dX, *dU, *loss_dp;
  int* status;
  unsigned bx, tx0, bdim;
};
static emu_args EA;
static void* emu_lane(void* p) {
  const unsigned lane = (unsigned)(uintptr_t)p;
  threadIdx.x = EA.tx0 + lane; threadIdx.y = threadIdx.z = 0;
  blockIdx.x = EA.bx; blockIdx.y = blockIdx.z = 0;
  blockDim.x = EA.bdim; blockDim.y = blockDim.z = 1;
  if (EA.kernel == 0)
    EMU_BWD_KERNEL(EA.B, EA.H, EA.X, EA.U, EA.Lam, EA.theta, EA.theta_stride, EA.gains, EA.auxrec, EA.termrec, EA.status);
  else
    pdp_k_aux_lqr_fwd(EA.B, EA.H, EA.X, EA.U, EA.theta, EA.theta_stride, EA.X0a, EA.x0a_stride, EA.dX, EA.dU, EA.gains,
                      EA.Xref, EA.Uref, EA.loss_dp, EA.auxrec, EA.status);
  return nullptr;
}
static void emu_run(unsigned nblocks, unsigned warps_per_block) {
  pthread_barrier_init(&emu_bar, nullptr, 32);
  for (unsigned bx = 0; bx < nblocks; ++bx)
    for (unsigned w = 0; w < warps_per_block; ++w) {
      EA.bx = bx; EA.tx0 = w * 32; EA.bdim = warps_per_block * 32;
      pthread_t th[32];
      for (unsigned l = 0; l < 32; ++l) pthread_create(&th[l], nullptr, emu_lane, (void*)(uintptr_t)l);
      for (unsigned l = 0; l < 32; ++l) pthread_join(th[l], nullptr);
    }
  pthread_barrier_destroy(&emu_bar);
}
extern "C" void emu_backward(int B, int H, const double* X, const double* U, const double* Lam, const double* theta,
                             int theta_stride, double* gains, int* status, const double* auxrec, const double* termrec) {
  memset(&EA, 0, sizeof(EA));
  EA.auxrec = auxrec; EA.termrec = termrec;
  EA.kernel = 0; EA.B = B; EA.H = H; EA.X = X; EA.U = U; EA.Lam = Lam; EA.theta = theta; EA.theta_stride = theta_stride;
  EA.gains = gains; EA.status = status;
  const unsigned per_block = EMU_BWD_TRAJ_PER_BLOCK;
  emu_run((B + per_block - 1) / per_block, EMU_BWD_WPB);
}
extern "C" void emu_forward(int B, int H, const double* X, const double* U, const double* theta, int theta_stride,
                            const double* X0a, int x0a_stride, double* dX, double* dU, const double* gains,
                            const double* Xref, const double* Uref, double* loss_dp, int* status, const double* auxrec) {
  memset(&EA, 0, sizeof(EA));
  EA.auxrec = auxrec;
  EA.kernel = 1; EA.B = B; EA.H = H; EA.X = X; EA.U = U; EA.theta = theta; EA.theta_stride = theta_stride;
  EA.X0a = X0a; EA.x0a_stride = x0a_stride; EA.dX = dX; EA.dU = dU; EA.gains = (double*)gains;
  EA.Xref = Xref; EA.Uref = Uref; EA.loss_dp = loss_dp; EA.status = status;
  const unsigned per_block = PDP_WPBF * PDP_FG;
  emu_run((B + per_block - 1) / per_block, PDP_WPBF);
}
extern "C" int emu_grec() { return PDP_GREC; }
#ifdef EMU_HAS_ROLLOUT
// thread-per-trajectory kernel without warp-level primitives: the threads run one after the other
extern "C" void emu_rollout(int B, int H, const double* x0, const double* theta, int theta_stride, const double* U,
                            double* X, double* Lam, double* cost, double* dHu, int* status,
                            const double* fb_gains, const double* fb_X, const double* fb_alpha, double* Uout, int fb_group) {
  for (int b = 0; b < B; ++b) {
    threadIdx.x = b % 128; blockIdx.x = b / 128; blockDim.x = 128;
    pdp_k_rollout_costate(B, H, x0, theta, theta_stride, U, X, Lam, cost, dHu, status, fb_gains, fb_X, fb_alpha, Uout, fb_group);
  }
}
#endif
#ifdef PDP_TSTRIDE
// TMA rollout kernel: thread-private shared slots, no inter-thread synchronisation -> the threads run one after the other
// (bulk copies are immediate memcpy's, mbarriers / fences are no-ops)
extern "C" void emu_rollout_tma(int B, int H, const double* x0, const double* theta, int theta_stride, const double* U,
                                double* X, double* Lam, double* cost, double* dHu, int* status) {
  for (int b = 0; b < B; ++b) {
    threadIdx.x = b % PDP_TB; blockIdx.x = b / PDP_TB; blockDim.x = PDP_TB;
    pdp_k_rollout_costate_tma(B, H, x0, theta, theta_stride, U, X, Lam, cost, dHu, status);
  }
}
#endif
'''

_RCP = re.compile(r'asm\("rcp\.approx\.ftz\.f64 %0, %1;" : "=d"\((\w+)\) : "d"\((\w+)\)\);')


def translate(cuda_source: str) -> str:
    """Generated CUDA translation unit -> C++ the emulator can compile (kernels + device functions only)."""
    cut = cuda_source.find("// Host-side launchers")
    if cut >= 0:
        cut = cuda_source.rfind("// ====", 0, cut)
    else:
        cut = cuda_source.find('extern "C" void pdpmod_info')      # sensitivity modules: launchers follow the kernels
    if cut < 0:
        raise ValueError("launcher marker not found in the generated source")
    body = cuda_source[:cut]
    body = body.replace("#include <cuda_runtime.h>", "")
    body = body.replace("extern __shared__ __align__(16) double pdp_smem[];", "extern double pdp_smem[];")
    body = _RCP.sub(lambda mo: "%s = 1.0 / %s;" % (mo.group(1), mo.group(2)), body)
    body = re.sub(r'asm volatile\("prefetch\.global\.L[12] \[%0\];" :: "l"\(([^;]*?)\)\);', r"(void)(\1);", body)
    if "asm(" in body:
        raise ValueError("untranslated inline asm in the generated source")
    return body


class Emulator:
    def __init__(self, src, bwd_kernel="pdp_k_aux_lqr_bwd", bwd_traj_per_block="PDP_WPB * PDP_BP", bwd_wpb="PDP_WPB",
                 keep=False):
        self.n, self.m, self.r = src.n, src.m, src.r
        text = src.source() if hasattr(src, "source") else str(src)
        has_ro = "pdp_k_rollout_costate" in text
        cpp = (PREAMBLE + translate(text) + ("\n#define EMU_HAS_ROLLOUT 1" if has_ro else "") + "\n#define EMU_BWD_KERNEL %s\n#define EMU_BWD_TRAJ_PER_BLOCK (%s)\n"
               "#define EMU_BWD_WPB (%s)\n" % (bwd_kernel, bwd_traj_per_block, bwd_wpb) + DRIVER)
        key = hashlib.sha256(cpp.encode()).hexdigest()[:16]
        d = os.path.join(tempfile.gettempdir(), "pdp_warp_emu")
        os.makedirs(d, exist_ok=True)
        so = os.path.join(d, "emu_%s.so" % key)
        if not os.path.isfile(so):
            cc = os.path.join(d, "emu_%s.cpp" % key)
            with open(cc, "w") as f:
                f.write(cpp)
            p = subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-w", "-ffp-contract=off",
                                "-o", so + ".tmp", cc], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if p.returncode != 0:
                raise RuntimeError("g++ failed on the emulated kernel source (%s):\n%s" % (cc, p.stdout[-6000:]))
            os.replace(so + ".tmp", so)
        self.lib = ctypes.CDLL(so)
        self.grec = int(self.lib.emu_grec())

    @staticmethod
    def _p(a):
        return None if a is None else a.ctypes.data_as(ctypes.c_void_p)

    def rollout(self, x0, theta, U, want_dHu=False, feedback=None, tma=False):
        """pdp_k_rollout_costate -> X, Lam, cost[, dHu].  ``feedback`` = dict(gains[Bs,H,(n+1)*m], X[Bs,H+1,n], alpha[B],
        group): closed-loop mode (B = group * Bs candidates); then the applied controls are returned as a fifth item."""
        x0, U = (np.ascontiguousarray(a, dtype=np.float64) for a in (x0, U))
        H = U.shape[1]
        theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        ts = 0 if theta.shape[0] == 1 else theta.shape[1]
        group = int(feedback["group"]) if feedback else 1
        B = U.shape[0] * group
        X = np.full((B, H + 1, self.n), np.nan)
        Lam = np.full((B, H, self.n), np.nan)
        cost = np.full(B, np.nan)
        dHu = np.full((B, H, self.m), np.nan) if want_dHu else None
        status = np.zeros(B, dtype=np.int32)
        fg = fx = fa = Uout = None
        if feedback:
            fg, fx, fa = (np.ascontiguousarray(feedback[k], dtype=np.float64) for k in ("gains", "X", "alpha"))
            Uout = np.full((B, H, self.m), np.nan)
        if tma:      # modules generated with rollout_tma=1: the open-loop kernel with bulk-copy row traffic
            self.lib.emu_rollout_tma(B, H, self._p(x0), self._p(theta), ts, self._p(U), self._p(X), self._p(Lam), self._p(cost),
                                     self._p(dHu), self._p(status))
            return X, Lam, cost, dHu
        self.lib.emu_rollout(B, H, self._p(x0), self._p(theta), ts, self._p(U), self._p(X), self._p(Lam), self._p(cost),
                             self._p(dHu), self._p(status), self._p(fg), self._p(fx), self._p(fa), self._p(Uout), group)
        return (X, Lam, cost, dHu, Uout) if feedback else (X, Lam, cost, dHu)

    def backward_dense(self, aux, term):
        """Generic dense LQR module: aux[B,H,NDENSE] (per step [F|G|E|Hxx|Hxu|Hxe|Hux|Huu|Hue] row-major), term[B,n*n+n*r]."""
        aux, term = np.ascontiguousarray(aux, dtype=np.float64), np.ascontiguousarray(term, dtype=np.float64)
        B, H = aux.shape[0], aux.shape[1]
        gains = np.full((B, H, self.grec), np.nan)
        status = np.zeros(B, dtype=np.int32)
        self.lib.emu_backward(B, H, None, None, None, None, 0, self._p(gains), self._p(status), self._p(aux), self._p(term))
        return gains, status

    def forward_dense(self, aux, gains, X0=None):
        aux, gains = np.ascontiguousarray(aux, dtype=np.float64), np.ascontiguousarray(gains, dtype=np.float64)
        B, H = aux.shape[0], aux.shape[1]
        dX = np.full((B, H + 1, self.n, self.r), np.nan)
        dU = np.full((B, H, self.m, self.r), np.nan)
        status = np.zeros(B, dtype=np.int32)
        X0 = None if X0 is None else np.ascontiguousarray(X0, dtype=np.float64)
        self.lib.emu_forward(B, H, None, None, None, 0, self._p(X0), 0 if X0 is None or X0.ndim == 2 else 1, self._p(dX),
                             self._p(dU), self._p(gains), None, None, None, self._p(status), self._p(aux))
        return dX, dU, status

    def backward(self, X, U, Lam, theta):
        B, H = U.shape[0], U.shape[1]
        X, U, Lam = (np.ascontiguousarray(a, dtype=np.float64) for a in (X, U, Lam))
        theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        ts = 0 if theta.shape[0] == 1 else theta.shape[1]
        gains = np.full((B, H, self.grec), np.nan)
        status = np.zeros(B, dtype=np.int32)
        self.lib.emu_backward(B, H, self._p(X), self._p(U), self._p(Lam), self._p(theta), ts, self._p(gains),
                              self._p(status), None, None)
        return gains, status

    def forward(self, X, U, theta, gains, Xref=None, Uref=None):
        B, H = U.shape[0], U.shape[1]
        n, m, r = self.n, self.m, self.r
        X, U, gains = (np.ascontiguousarray(a, dtype=np.float64) for a in (X, U, gains))
        theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        ts = 0 if theta.shape[0] == 1 else theta.shape[1]
        dX = np.full((B, H + 1, n, r), np.nan)
        dU = np.full((B, H, m, r), np.nan)
        status = np.zeros(B, dtype=np.int32)
        ldp = None
        if Xref is not None:
            Xref = np.ascontiguousarray(Xref, dtype=np.float64)
            Uref = None if Uref is None else np.ascontiguousarray(Uref, dtype=np.float64)
            ldp = np.full((B, r + 1), np.nan)
        self.lib.emu_forward(B, H, self._p(X), self._p(U), self._p(theta), ts, None, 0, self._p(dX), self._p(dU),
                             self._p(gains), self._p(Xref), self._p(Uref), self._p(ldp), self._p(status), None)
        return dX, dU, ldp, status


SENS_DRIVER = r'''
extern "C" void emu_sens(int B, int H, const double* x0, const double* theta, int theta_stride, const double* inputs,
                         const double* Xobs, double* X, double* Uout, double* dX, double* dU, double* loss_dp, int* status) {
  for (int g = 0; g < PDP_NG; ++g)
    for (int b = 0; b < B; ++b) {
      threadIdx.x = b % PDP_BLOCK; blockIdx.y = b / PDP_BLOCK; blockIdx.x = g; blockDim.x = PDP_BLOCK;
      pdp_k_sens_fwd(B, H, x0, theta, theta_stride, inputs, Xobs, X, Uout, dX, dU, loss_dp, status);
    }
}
'''


class SensEmulator:
    """SysID / ControlPlanning forward-sensitivity module (codegen_sens.SensModuleSource) on the CPU."""

    def __init__(self, src):
        self.n, self.m, self.r = src.n, src.m, src.r
        cpp = PREAMBLE + "static inline double atomicAdd(double* p, double v) { double o = *p; *p += v; return o; }\n" + \
            translate(src.source()) + SENS_DRIVER
        key = hashlib.sha256(cpp.encode()).hexdigest()[:16]
        d = os.path.join(tempfile.gettempdir(), "pdp_warp_emu")
        os.makedirs(d, exist_ok=True)
        so = os.path.join(d, "emu_%s.so" % key)
        if not os.path.isfile(so):
            cc = os.path.join(d, "emu_%s.cpp" % key)
            with open(cc, "w") as f:
                f.write(cpp)
            p = subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-w", "-ffp-contract=off",
                                "-o", so + ".tmp", cc], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if p.returncode != 0:
                raise RuntimeError("g++ failed on the emulated kernel source (%s):\n%s" % (cc, p.stdout[-6000:]))
            os.replace(so + ".tmp", so)
        self.lib = ctypes.CDLL(so)

    def run(self, x0, theta, H, inputs=None, Xobs=None, policy=False, fused_only=False):
        """-> X[B,H+1,n], U[B,H,m] (policy modules), dX[B,H+1,n,r], dU[B,H,m,r] (policy), loss_dp[B,r+1].
        ``fused_only``: the entry point without trajectory outputs (pdp_k_sens_fwd) -> loss_dp only."""
        p = Emulator._p
        x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
        B = x0.shape[0]
        theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        ts = 0 if theta.shape[0] == 1 else theta.shape[1]
        inputs = None if inputs is None else np.ascontiguousarray(inputs, dtype=np.float64)
        Xobs = None if Xobs is None else np.ascontiguousarray(Xobs, dtype=np.float64)
        X = None if fused_only else np.full((B, H + 1, self.n), np.nan)
        dX = None if fused_only else np.full((B, H + 1, self.n, self.r), np.nan)
        Uo = np.full((B, H, self.m), np.nan) if (policy and not fused_only) else None
        dU = np.full((B, H, self.m, self.r), np.nan) if (policy and not fused_only) else None
        ldp = np.zeros((B, self.r + 1))
        status = np.zeros(B, dtype=np.int32)
        self.lib.emu_sens(B, H, p(x0), p(theta), ts, p(inputs), p(Xobs), p(X), p(Uo), p(dX), p(dU), p(ldp), p(status))
        return {"X": X, "U": Uo, "dX": dX, "dU": dU, "loss_dp": ldp, "status": status}
