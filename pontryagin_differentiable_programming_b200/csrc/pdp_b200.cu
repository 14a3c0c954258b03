// libpdp_b200.so -- C ABI of the B200 PDP engine (see include/pdp_b200.h).
// Thin, allocation-free dispatch layer: a "system" is a generated CUDA module (one .so per symbolic
// optimal-control / sysid / planning system, produced by codegen.py + nvcc) opened with dlopen; the
// kernels themselves live in that module so they are fully specialised on (n, m, r) and on the
// sparsity pattern of the system's derivatives.
#include "pdp_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

using fn_info = void (*)(int*);
using fn_rollout = int (*)(int, int, const double*, const double*, int, const double*, double*, double*, double*, double*, int*,
                           const double*, const double*, const double*, double*, int, cudaStream_t);
using fn_aux_eval = int (*)(int, int, const double*, const double*, const double*, const double*, int, double*, double*,
                            cudaStream_t);
using fn_aux_lqr = int (*)(int, int, const double*, const double*, const double*, const double*, int, const double*, int,
                           double*, double*, double*, const double*, const double*, double*, const double*, const double*,
                           int, int*, cudaStream_t);
using fn_eval = int (*)(int, const double* const*, const int*, double* const*, cudaStream_t);
using fn_sens = int (*)(int, int, const double*, const double*, int, const double*, const double*, double*, double*, double*,
                        double*, double*, int*, cudaStream_t);

inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

}  // namespace

struct pdp_system {
  void* handle = nullptr;
  int info[16] = {0};
  fn_rollout rollout = nullptr;
  fn_aux_eval aux_eval = nullptr;
  fn_aux_lqr aux_lqr = nullptr;
  fn_sens sens = nullptr;
  fn_eval fneval = nullptr;
  // pdp_sweep pipelining: sub-batches of the aux-LQR phase alternate between two internal streams so that the
  // shared-memory-bound backward kernel of one sub-batch overlaps the HBM-bound forward kernel of the other
  std::mutex mu;                                  // serialises the fork / join enqueue sequence
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  int side_dev = -1;
  int sweep_parts = 0;                            // 0 = auto (by batch size), 1 = off, k = k sub-batches
  int kind() const { return info[0]; }
  int n() const { return info[1]; }
  int m() const { return info[2]; }
  int r() const { return info[3]; }
  int grec() const { return info[6]; }
};

extern "C" {

const char* pdp_last_error(void) { return g_err; }
const char* pdp_version(void) { return "pdp_b200 0.1 (sm_100a)"; }

int pdp_load_system(const char* module_path, pdp_system_t** out) {
  if (!module_path || !out) return fail(PDP_ERR_ARG, "pdp_load_system: null argument");
  void* h = dlopen(module_path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return fail(PDP_ERR_LOAD, "pdp_load_system: dlopen(%s) failed: %s", module_path, dlerror());
  auto info = reinterpret_cast<fn_info>(dlsym(h, "pdpmod_info"));
  if (!info) {
    dlclose(h);
    return fail(PDP_ERR_LOAD, "pdp_load_system: %s is not a PDP system module (no pdpmod_info)", module_path);
  }
  pdp_system* s = new (std::nothrow) pdp_system();
  if (!s) {
    dlclose(h);
    return fail(PDP_ERR_LOAD, "pdp_load_system: out of host memory");
  }
  s->handle = h;
  info(s->info);
  s->rollout = reinterpret_cast<fn_rollout>(dlsym(h, "pdpmod_rollout_costate"));
  s->aux_eval = reinterpret_cast<fn_aux_eval>(dlsym(h, "pdpmod_aux_eval"));
  s->aux_lqr = reinterpret_cast<fn_aux_lqr>(dlsym(h, "pdpmod_aux_lqr"));
  s->sens = reinterpret_cast<fn_sens>(dlsym(h, "pdpmod_sens_fwd"));
  s->fneval = reinterpret_cast<fn_eval>(dlsym(h, "pdpmod_fn"));
  *out = s;
  return PDP_OK;
}

static void destroy_side(pdp_system* s) {
  for (int k = 0; k < 2; ++k) {
    if (s->side[k]) cudaStreamDestroy(s->side[k]);
    if (s->ev_join[k]) cudaEventDestroy(s->ev_join[k]);
    s->side[k] = nullptr;
    s->ev_join[k] = nullptr;
  }
  if (s->ev_fork) cudaEventDestroy(s->ev_fork);
  s->ev_fork = nullptr;
  s->side_dev = -1;
}

// internal streams / events of the current device, created on first use (never during a stream capture)
static bool ensure_side(pdp_system* s, bool may_create) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (s->side_dev == dev) return true;
  if (!may_create) return false;
  destroy_side(s);
  bool ok = cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming) == cudaSuccess;
  for (int k = 0; k < 2 && ok; ++k)
    ok = cudaStreamCreateWithFlags(&s->side[k], cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&s->ev_join[k], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { destroy_side(s); cudaGetLastError(); return false; }
  s->side_dev = dev;
  return true;
}

void pdp_free_system(pdp_system_t* sys) {
  if (!sys) return;
  destroy_side(sys);
  if (sys->handle) dlclose(sys->handle);
  delete sys;
}

int pdp_set_sweep_parts(pdp_system_t* sys, int parts) {
  if (!sys || parts < 0) return fail(PDP_ERR_ARG, "pdp_set_sweep_parts: bad argument");
  sys->sweep_parts = parts;
  return PDP_OK;
}

int pdp_system_dims(const pdp_system_t* sys, int* dims) {
  if (!sys || !dims) return fail(PDP_ERR_ARG, "pdp_system_dims: null argument");
  dims[0] = sys->kind(); dims[1] = sys->n(); dims[2] = sys->m(); dims[3] = sys->r();
  return PDP_OK;
}

size_t pdp_workspace_bytes(const pdp_system_t* sys, int op, int B, int H) {
  if (!sys || B <= 0 || H <= 0) return 0;
  const size_t n = sys->n(), m = sys->m(), r = sys->r();
  const size_t gains = align256(size_t(B) * H * sys->grec() * sizeof(double));
  switch (op) {
    case PDP_OP_AUX_LQR:
    case PDP_OP_SWEEP:
      return gains;
    case PDP_OP_SWEEP_HOST: {
      size_t tot = gains;
      tot += align256(size_t(B) * n * 8);            // x0
      tot += align256(size_t(B) * r * 8);            // theta
      tot += align256(size_t(B) * H * m * 8);        // U
      tot += 2 * align256(size_t(B) * (H + 1) * n * 8);  // X, Xref
      tot += align256(size_t(B) * H * n * 8);        // Lam
      tot += align256(size_t(B) * H * m * 8);        // Uref
      tot += align256(size_t(B) * 8);                // cost
      tot += align256(size_t(B) * (r + 1) * 8);      // loss_dp
      tot += align256(size_t(B) * (H + 1) * n * r * 8);  // dX
      tot += align256(size_t(B) * H * m * r * 8);    // dU
      return tot;
    }
    default:
      return 0;
  }
}

int pdp_rollout_costate(pdp_system_t* sys, int B, int H, const double* x0, const double* theta, int theta_stride,
                        const double* U, double* X, double* Lam, double* cost, double* dHu, int* status,
                        pdp_stream_t stream) {
  if (!sys || !sys->rollout) return fail(PDP_ERR_UNSUPPORTED, "pdp_rollout_costate: module has no rollout kernel");
  if (B == 0) return PDP_OK;  /* empty batch: nothing to do, pointers may be NULL */
  if (B < 0 || H < 1 || !x0 || !theta || !U || !X) return fail(PDP_ERR_ARG, "pdp_rollout_costate: bad argument");
  if (dHu && !Lam) return fail(PDP_ERR_ARG, "pdp_rollout_costate: dHu needs Lam");
  int e = sys->rollout(B, H, x0, theta, theta_stride, U, X, Lam, cost, dHu, status, nullptr, nullptr, nullptr, nullptr, 1,
                       (cudaStream_t)stream);
  if (e) return fail(PDP_ERR_CUDA, "pdp_rollout_costate: CUDA error %d (%s)", e, cudaGetErrorString((cudaError_t)e));
  return PDP_OK;
}

int pdp_rollout_feedback(pdp_system_t* sys, int B, int H, const double* x0, const double* theta, int theta_stride,
                         const double* Uref, const double* Xref, const double* gains, const double* alpha, double* Uout,
                         double* X, double* Lam, double* cost, double* dHu, int group, int* status, pdp_stream_t stream) {
  if (!sys || !sys->rollout) return fail(PDP_ERR_UNSUPPORTED, "pdp_rollout_feedback: module has no rollout kernel");
  if (B == 0) return PDP_OK;  /* empty batch: nothing to do, pointers may be NULL */
  if (B < 0 || H < 1 || !x0 || !theta || !Uref || !Xref || !gains || !alpha || !Uout || !X)
    return fail(PDP_ERR_ARG, "pdp_rollout_feedback: bad argument");
  if (dHu && !Lam) return fail(PDP_ERR_ARG, "pdp_rollout_feedback: dHu needs Lam");
  if (group < 1 || B % group != 0) return fail(PDP_ERR_ARG, "pdp_rollout_feedback: B must be a multiple of group");
  int e = sys->rollout(B, H, x0, theta, theta_stride, Uref, X, Lam, cost, dHu, status, gains, Xref, alpha, Uout, group,
                       (cudaStream_t)stream);
  if (e) return fail(PDP_ERR_CUDA, "pdp_rollout_feedback: CUDA error %d (%s)", e, cudaGetErrorString((cudaError_t)e));
  return PDP_OK;
}

static int aux_lqr_phases(int phases, pdp_system_t* sys, int B, int H, const double* X, const double* U, const double* Lam,
                const double* theta, int theta_stride, const double* X0aux, int x0aux_stride,
                double* dXdtheta, double* dUdtheta, const double* Xref, const double* Uref, double* loss_dp,
                void* workspace, size_t ws_bytes, int* status, pdp_stream_t stream) {
  if (!sys || !sys->aux_lqr || sys->kind() != PDP_KIND_OC)
    return fail(PDP_ERR_UNSUPPORTED, "pdp_aux_lqr: module has no fused aux-LQR kernel");
  if (B == 0) return PDP_OK;  /* empty batch: nothing to do, pointers may be NULL */
  if (B < 0 || H < 1 || !X || !U || !Lam || !theta) return fail(PDP_ERR_ARG, "pdp_aux_lqr: bad argument");
  if (loss_dp && !Xref) return fail(PDP_ERR_ARG, "pdp_aux_lqr: loss_dp needs Xref");
  if (ws_bytes < pdp_workspace_bytes(sys, PDP_OP_AUX_LQR, B, H) || (B > 0 && !workspace))
    return fail(PDP_ERR_WORKSPACE, "pdp_aux_lqr: workspace too small (%zu < %zu)", ws_bytes,
                pdp_workspace_bytes(sys, PDP_OP_AUX_LQR, B, H));
  int e = sys->aux_lqr(B, H, X, U, Lam, theta, theta_stride, X0aux, x0aux_stride, dXdtheta, dUdtheta,
                       reinterpret_cast<double*>(workspace), Xref, Uref, loss_dp, nullptr, nullptr, phases, status,
                       (cudaStream_t)stream);
  if (e) return fail(PDP_ERR_CUDA, "pdp_aux_lqr: CUDA error %d (%s)", e, cudaGetErrorString((cudaError_t)e));
  return PDP_OK;
}

int pdp_aux_lqr(pdp_system_t* sys, int B, int H, const double* X, const double* U, const double* Lam,
                const double* theta, int theta_stride, const double* X0aux, int x0aux_stride,
                double* dXdtheta, double* dUdtheta, const double* Xref, const double* Uref, double* loss_dp,
                void* workspace, size_t ws_bytes, int* status, pdp_stream_t stream) {
  return aux_lqr_phases(3, sys, B, H, X, U, Lam, theta, theta_stride, X0aux, x0aux_stride, dXdtheta, dUdtheta, Xref, Uref,
                        loss_dp, workspace, ws_bytes, status, stream);
}

int pdp_aux_lqr_backward(pdp_system_t* sys, int B, int H, const double* X, const double* U, const double* Lam,
                         const double* theta, int theta_stride, void* workspace, size_t ws_bytes, int* status,
                         pdp_stream_t stream) {
  return aux_lqr_phases(1, sys, B, H, X, U, Lam, theta, theta_stride, nullptr, 0, nullptr, nullptr, nullptr, nullptr,
                        nullptr, workspace, ws_bytes, status, stream);
}

int pdp_aux_lqr_forward(pdp_system_t* sys, int B, int H, const double* X, const double* U, const double* theta,
                        int theta_stride, const double* X0aux, int x0aux_stride, double* dXdtheta, double* dUdtheta,
                        const double* Xref, const double* Uref, double* loss_dp, const void* workspace, size_t ws_bytes,
                        int* status, pdp_stream_t stream) {
  return aux_lqr_phases(2, sys, B, H, X, U, X /*unused*/, theta, theta_stride, X0aux, x0aux_stride, dXdtheta, dUdtheta,
                        Xref, Uref, loss_dp, const_cast<void*>(workspace), ws_bytes, status, stream);
}

int pdp_lqr_dense(pdp_system_t* sys, int B, int H, const double* aux, const double* term, const double* X0aux,
                  int x0aux_stride, double* Xaux, double* Uaux, int forward_only, void* workspace, size_t ws_bytes,
                  int* status, pdp_stream_t stream) {
  if (!sys || !sys->aux_lqr || sys->kind() != PDP_KIND_LQR)
    return fail(PDP_ERR_UNSUPPORTED, "pdp_lqr_dense: not a dense-LQR module");
  if (B == 0) return PDP_OK;  /* empty batch: nothing to do, pointers may be NULL */
  if (B < 0 || H < 1 || !aux || (!term && !forward_only)) return fail(PDP_ERR_ARG, "pdp_lqr_dense: bad argument");
  if (ws_bytes < pdp_workspace_bytes(sys, PDP_OP_AUX_LQR, B, H) || (B > 0 && !workspace))
    return fail(PDP_ERR_WORKSPACE, "pdp_lqr_dense: workspace too small");
  int e = sys->aux_lqr(B, H, nullptr, nullptr, nullptr, nullptr, 0, X0aux, x0aux_stride, Xaux, Uaux,
                       reinterpret_cast<double*>(workspace), nullptr, nullptr, nullptr, aux, term, forward_only ? 2 : 3,
                       status, (cudaStream_t)stream);
  if (e) return fail(PDP_ERR_CUDA, "pdp_lqr_dense: CUDA error %d (%s)", e, cudaGetErrorString((cudaError_t)e));
  return PDP_OK;
}

int pdp_sweep(pdp_system_t* sys, int B, int H, const double* x0, const double* theta, int theta_stride,
              const double* U, double* X, double* Lam, double* cost, double* dXdtheta, double* dUdtheta,
              const double* Xref, const double* Uref, double* loss_dp, void* workspace, size_t ws_bytes,
              int* status, pdp_stream_t stream) {
  if (B == 0) return PDP_OK;
  if (!Lam) return fail(PDP_ERR_ARG, "pdp_sweep: Lam buffer required");
  int e = pdp_rollout_costate(sys, B, H, x0, theta, theta_stride, U, X, Lam, cost, nullptr, status, stream);
  if (e) return e;
  int parts = sys->sweep_parts > 0 ? sys->sweep_parts : (B >= 16384 ? 4 : (B >= 8192 ? 2 : 1));
  if (parts > B) parts = B;
  cudaStream_t st = (cudaStream_t)stream;
  if (parts > 1) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) { cudaGetLastError(); parts = 1; }
    std::lock_guard<std::mutex> lock(sys->mu);
    if (parts > 1 && !ensure_side(sys, cap == cudaStreamCaptureStatusNone)) parts = 1;
    if (parts > 1) {
      // same checks as pdp_aux_lqr, once for the whole batch; the sub-batches use disjoint slices of every buffer
      if (!sys->aux_lqr || sys->kind() != PDP_KIND_OC)
        return fail(PDP_ERR_UNSUPPORTED, "pdp_sweep: module has no fused aux-LQR kernel");
      if (loss_dp && !Xref) return fail(PDP_ERR_ARG, "pdp_sweep: loss_dp needs Xref");
      if (!workspace || ws_bytes < pdp_workspace_bytes(sys, PDP_OP_SWEEP, B, H))
        return fail(PDP_ERR_WORKSPACE, "pdp_sweep: workspace too small");
      const size_t n = sys->n(), m = sys->m(), r = sys->r(), g = sys->grec();
      double* gains = reinterpret_cast<double*>(workspace);
      cudaError_t ce = cudaEventRecord(sys->ev_fork, st);
      for (int k = 0; k < 2 && ce == cudaSuccess; ++k) ce = cudaStreamWaitEvent(sys->side[k], sys->ev_fork, 0);
      if (ce != cudaSuccess) return fail(PDP_ERR_CUDA, "pdp_sweep: %s", cudaGetErrorString(ce));
      int err = 0;
      for (int i = 0; i < parts && !err; ++i) {
        const size_t lo = (size_t(B) * i) / parts, hi = (size_t(B) * (i + 1)) / parts;
        err = sys->aux_lqr(int(hi - lo), H, X + lo * (H + 1) * n, U + lo * H * m, Lam + lo * H * n,
                           theta + (theta_stride ? lo * size_t(theta_stride) : 0), theta_stride, nullptr, 0,
                           dXdtheta ? dXdtheta + lo * (H + 1) * n * r : nullptr, dUdtheta ? dUdtheta + lo * H * m * r : nullptr,
                           gains + lo * H * g, Xref ? Xref + lo * (H + 1) * n : nullptr, Uref ? Uref + lo * H * m : nullptr,
                           loss_dp ? loss_dp + lo * (r + 1) : nullptr, nullptr, nullptr, 3, status ? status + lo : nullptr,
                           sys->side[i & 1]);
      }
      // always join, also after a failed launch, so that the caller's stream stays ordered behind the side streams
      for (int k = 0; k < 2; ++k) {
        ce = cudaEventRecord(sys->ev_join[k], sys->side[k]);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(st, sys->ev_join[k], 0);
        if (ce != cudaSuccess && !err) err = (int)ce;
      }
      if (err) return fail(PDP_ERR_CUDA, "pdp_sweep: CUDA error %d (%s)", err, cudaGetErrorString((cudaError_t)err));
      return PDP_OK;
    }
  }
  return pdp_aux_lqr(sys, B, H, X, U, Lam, theta, theta_stride, nullptr, 0, dXdtheta, dUdtheta, Xref, Uref, loss_dp,
                     workspace, ws_bytes, status, stream);
}

int pdp_aux_eval(pdp_system_t* sys, int B, int H, const double* X, const double* U, const double* Lam,
                 const double* theta, int theta_stride, double* aux, double* term, pdp_stream_t stream) {
  if (!sys || !sys->aux_eval) return fail(PDP_ERR_UNSUPPORTED, "pdp_aux_eval: module has no aux-eval kernel");
  if (B == 0) return PDP_OK;  /* empty batch: nothing to do, pointers may be NULL */
  if (B < 0 || H < 1 || !X || !U || !Lam || !theta || !aux) return fail(PDP_ERR_ARG, "pdp_aux_eval: bad argument");
  int e = sys->aux_eval(B, H, X, U, Lam, theta, theta_stride, aux, term, (cudaStream_t)stream);
  if (e) return fail(PDP_ERR_CUDA, "pdp_aux_eval: CUDA error %d (%s)", e, cudaGetErrorString((cudaError_t)e));
  return PDP_OK;
}

int pdp_sens_fwd(pdp_system_t* sys, int B, int H, const double* x0, const double* theta, int theta_stride,
                 const double* inputs, const double* Xobs, double* X, double* Uout, double* dX, double* dU,
                 double* loss_dp, int* status, pdp_stream_t stream) {
  if (!sys || !sys->sens) return fail(PDP_ERR_UNSUPPORTED, "pdp_sens_fwd: module has no forward-sensitivity kernel");
  if (B == 0) return PDP_OK;  /* empty batch: nothing to do, pointers may be NULL */
  if (B < 0 || H < 1 || !x0 || !theta) return fail(PDP_ERR_ARG, "pdp_sens_fwd: bad argument");
  int e = sys->sens(B, H, x0, theta, theta_stride, inputs, Xobs, X, Uout, dX, dU, loss_dp, status, (cudaStream_t)stream);
  if (e) return fail(PDP_ERR_CUDA, "pdp_sens_fwd: CUDA error %d (%s)", e, cudaGetErrorString((cudaError_t)e));
  return PDP_OK;
}

int pdp_eval_function(pdp_system_t* sys, int B, const double* const* inputs, const int* input_strides,
                      double* const* outputs, pdp_stream_t stream) {
  if (!sys || !sys->fneval) return fail(PDP_ERR_UNSUPPORTED, "pdp_eval_function: not a function module");
  if (B == 0) return PDP_OK;  /* empty batch: nothing to do, pointers may be NULL */
  if (B < 0 || !inputs || !input_strides || !outputs) return fail(PDP_ERR_ARG, "pdp_eval_function: bad argument");
  int e = sys->fneval(B, inputs, input_strides, outputs, (cudaStream_t)stream);
  if (e) return fail(PDP_ERR_CUDA, "pdp_eval_function: CUDA error %d (%s)", e, cudaGetErrorString((cudaError_t)e));
  return PDP_OK;
}

int pdp_sweep_host(pdp_system_t* sys, int B, int H, const double* x0_host, const double* theta_host,
                   int theta_stride, const double* U_host, const double* Xref_host, const double* Uref_host,
                   double* loss_dp_host, double* cost_host, int keep_dtraj, void* workspace, size_t ws_bytes,
                   pdp_stream_t stream) {
  if (!sys || !sys->aux_lqr || !sys->rollout) return fail(PDP_ERR_UNSUPPORTED, "pdp_sweep_host: not an OC module");
  if (B < 1 || H < 1 || !x0_host || !theta_host || !U_host || !Xref_host || !loss_dp_host)
    return fail(PDP_ERR_ARG, "pdp_sweep_host: bad argument");
  if (!workspace || ws_bytes < pdp_workspace_bytes(sys, PDP_OP_SWEEP_HOST, B, H))
    return fail(PDP_ERR_WORKSPACE, "pdp_sweep_host: workspace too small");
  const size_t n = sys->n(), m = sys->m(), r = sys->r();
  cudaStream_t st = (cudaStream_t)stream;
  char* p = reinterpret_cast<char*>(workspace);
  auto take = [&](size_t bytes) { char* q = p; p += align256(bytes); return reinterpret_cast<double*>(q); };
  double* gains = take(size_t(B) * H * sys->grec() * 8);
  double* d_x0 = take(size_t(B) * n * 8);
  double* d_th = take(size_t(B) * r * 8);
  double* d_U = take(size_t(B) * H * m * 8);
  double* d_X = take(size_t(B) * (H + 1) * n * 8);
  double* d_Xr = take(size_t(B) * (H + 1) * n * 8);
  double* d_L = take(size_t(B) * H * n * 8);
  double* d_Ur = take(size_t(B) * H * m * 8);
  double* d_cost = take(size_t(B) * 8);
  double* d_ldp = take(size_t(B) * (r + 1) * 8);
  double* d_dX = take(size_t(B) * (H + 1) * n * r * 8);
  double* d_dU = take(size_t(B) * H * m * r * 8);
  const size_t thn = theta_stride ? size_t(B) * r : r;
  cudaError_t ce;
#define PDP_CK(x) if ((ce = (x)) != cudaSuccess) return fail(PDP_ERR_CUDA, "pdp_sweep_host: %s", cudaGetErrorString(ce))
  PDP_CK(cudaMemcpyAsync(d_x0, x0_host, size_t(B) * n * 8, cudaMemcpyHostToDevice, st));
  PDP_CK(cudaMemcpyAsync(d_th, theta_host, thn * 8, cudaMemcpyHostToDevice, st));
  PDP_CK(cudaMemcpyAsync(d_U, U_host, size_t(B) * H * m * 8, cudaMemcpyHostToDevice, st));
  PDP_CK(cudaMemcpyAsync(d_Xr, Xref_host, size_t(B) * (H + 1) * n * 8, cudaMemcpyHostToDevice, st));
  if (Uref_host) PDP_CK(cudaMemcpyAsync(d_Ur, Uref_host, size_t(B) * H * m * 8, cudaMemcpyHostToDevice, st));
  int e = pdp_sweep(sys, B, H, d_x0, d_th, theta_stride, d_U, d_X, d_L, d_cost, keep_dtraj ? d_dX : nullptr,
                    keep_dtraj ? d_dU : nullptr, d_Xr, Uref_host ? d_Ur : nullptr, d_ldp, gains,
                    size_t(B) * H * sys->grec() * 8 + 256, nullptr, stream);
  if (e) return e;
  PDP_CK(cudaMemcpyAsync(loss_dp_host, d_ldp, size_t(B) * (r + 1) * 8, cudaMemcpyDeviceToHost, st));
  if (cost_host) PDP_CK(cudaMemcpyAsync(cost_host, d_cost, size_t(B) * 8, cudaMemcpyDeviceToHost, st));
#undef PDP_CK
  return PDP_OK;
}

}  // extern "C"
