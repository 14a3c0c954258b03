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

// ---- batch reduction of the per-trajectory (loss, dp) rows: loss_dp[B, r1] -> sums[r1 + 1] = (column sums, B) ----------
// The outer loops of the IRL / SysID modes average over the batch (reference PDP/PDP.py:1293-1294,
// Examples/IRL/quadrotor/uav_PDP.py:78-81); with several GPUs this vector is what the one all-reduce carries.
// Deterministic (no floating-point atomics): block g sums its slab of rows column-wise with coalesced loads (thread
// t owns column t % r1 and row lane t / r1), publishes partial[g][.], and the block that draws the last ticket adds the
// partials in block order.  The ticket counter returns to zero, so the call can be replayed from a CUDA graph.
constexpr int kRedBlocks = 64, kRedThreads = 256;

// Column sums of rows [lo, hi) of a row-major [rows, r1] array by one block, in a fixed order: thread t owns column t % r1 and
// row lane t / r1 (consecutive threads read consecutive doubles), four independent chains per thread, then the row lanes are
// added in lane order.  `load_cg`: read through L2 (the partial sums other blocks have just published).
template <bool LOAD_CG>
__device__ __forceinline__ void pdp_block_colsum(const double* __restrict__ src, int lo, int hi, int r1, double* red_sm,
                                                 double* __restrict__ out) {
  const int t = threadIdx.x;
  auto ld = [&](size_t i) { return LOAD_CG ? __ldcg(src + i) : src[i]; };
  if (r1 >= kRedThreads) {
    for (int c = t; c < r1; c += kRedThreads) {      // thread per column, consecutive threads read consecutive doubles
      double a = 0.0;
      for (int row = lo; row < hi; ++row) a += ld((size_t)row * r1 + c);
      out[c] = a;
    }
    return;
  }
  const int k = kRedThreads / r1;                    // row lanes
  const int c = t % r1, rl = t / r1;
  if (rl < k) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int row = lo + rl;
    for (; row + 3 * k < hi; row += 4 * k) {
      a0 += ld((size_t)row * r1 + c);
      a1 += ld((size_t)(row + k) * r1 + c);
      a2 += ld((size_t)(row + 2 * k) * r1 + c);
      a3 += ld((size_t)(row + 3 * k) * r1 + c);
    }
    for (; row < hi; row += k) a0 += ld((size_t)row * r1 + c);
    red_sm[rl * r1 + c] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
  if (t < r1) {
    double a = 0.0;
    for (int j = 0; j < k; ++j) a += red_sm[j * r1 + t];
    out[t] = a;
  }
  __syncthreads();
}

extern "C" __global__ void __launch_bounds__(kRedThreads)
pdp_k_reduce_loss_dp(int B, int r1, const double* __restrict__ ldp, double* __restrict__ sums, double* __restrict__ partial,
                     unsigned int* __restrict__ ticket) {
  extern __shared__ double red_sm[];                 // [row lanes][r1] (r1 < blockDim) or unused
  const int g = blockIdx.x, t = threadIdx.x;
  const int lo = (int)(((long long)B * g) / gridDim.x), hi = (int)(((long long)B * (g + 1)) / gridDim.x);
  pdp_block_colsum<false>(ldp, lo, hi, r1, red_sm, partial + (size_t)g * r1);
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (t == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  pdp_block_colsum<true>(partial, 0, (int)gridDim.x, r1, red_sm, sums);
  if (t == 0) { sums[r1] = (double)B; *ticket = 0u; }
}


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
    case PDP_OP_ROLLOUT_HOST: {
      const size_t nth = sys->info[11] > 0 ? sys->info[11] : r;
      return align256(size_t(B) * n * 8) + align256(size_t(B) * nth * 8) + 2 * align256(size_t(B) * H * m * 8) +
             align256(size_t(B) * (H + 1) * n * 8) + align256(size_t(B) * H * n * 8) + align256(size_t(B) * 8);
    }
    case PDP_OP_SENS_HOST:
      return align256(pdp_reduce_workspace_bytes((int)r)) + align256(size_t(B) * n * 8) + align256(size_t(B) * r * 8) +
             align256(size_t(B) * H * m * 8) + align256(size_t(B) * (H + 1) * n * 8) + align256(size_t(B) * (r + 1) * 8) +
             align256((r + 2) * 8);
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
  // the rollout / costate kernel runs once for the whole batch: it is latency-bound (0.1 ms whatever the batch), and running
  // it per sub-batch on the side streams measured slower (profiles/r2k_sweep_pipeline_ab.json: 1.143 vs 1.067 ms at 4 parts)
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

size_t pdp_reduce_workspace_bytes(int r) {
  if (r < 0) return 0;
  return align256(size_t(kRedBlocks) * size_t(r + 1) * sizeof(double)) + 256;
}

int pdp_reduce_loss_dp(int B, int r, const double* loss_dp, double* sums, void* workspace, size_t ws_bytes,
                       pdp_stream_t stream) {
  if (B < 1 || r < 0 || !loss_dp || !sums) return fail(PDP_ERR_ARG, "pdp_reduce_loss_dp: bad argument");
  if (!workspace || ws_bytes < pdp_reduce_workspace_bytes(r))
    return fail(PDP_ERR_WORKSPACE, "pdp_reduce_loss_dp: workspace too small (%zu < %zu)", ws_bytes,
                pdp_reduce_workspace_bytes(r));
  const int r1 = r + 1;
  // layout: [ticket (256 B, zero on first use, left at zero by every call)] [partial sums kRedBlocks x r1]
  unsigned int* ticket = reinterpret_cast<unsigned int*>(workspace);
  double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 256);
  const int blocks = B < kRedBlocks ? B : kRedBlocks;
  const size_t smem = r1 < kRedThreads ? size_t(kRedThreads / r1) * r1 * sizeof(double) : 0;
  pdp_k_reduce_loss_dp<<<blocks, kRedThreads, smem, (cudaStream_t)stream>>>(B, r1, loss_dp, sums, partial, ticket);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return fail(PDP_ERR_CUDA, "pdp_reduce_loss_dp: %s", cudaGetErrorString(ce));
  return PDP_OK;
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

static int sweep_host_impl(pdp_system_t* sys, int B, int H, const double* x0_host, const double* theta_host,
                           int theta_stride, const double* U_host, const double* Xref_host, const double* Uref_host,
                           double* loss_dp_host, double* cost_host, int keep_dtraj, double* X_host, double* Lam_host,
                           double* dX_host, double* dU_host, void* workspace, size_t ws_bytes, pdp_stream_t stream) {
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
  if (X_host) PDP_CK(cudaMemcpyAsync(X_host, d_X, size_t(B) * (H + 1) * n * 8, cudaMemcpyDeviceToHost, st));
  if (Lam_host) PDP_CK(cudaMemcpyAsync(Lam_host, d_L, size_t(B) * H * n * 8, cudaMemcpyDeviceToHost, st));
  if (dX_host) PDP_CK(cudaMemcpyAsync(dX_host, d_dX, size_t(B) * (H + 1) * n * r * 8, cudaMemcpyDeviceToHost, st));
  if (dU_host) PDP_CK(cudaMemcpyAsync(dU_host, d_dU, size_t(B) * H * m * r * 8, cudaMemcpyDeviceToHost, st));
#undef PDP_CK
  return PDP_OK;
}

int pdp_sweep_host(pdp_system_t* sys, int B, int H, const double* x0_host, const double* theta_host,
                   int theta_stride, const double* U_host, const double* Xref_host, const double* Uref_host,
                   double* loss_dp_host, double* cost_host, int keep_dtraj, void* workspace, size_t ws_bytes,
                   pdp_stream_t stream) {
  return sweep_host_impl(sys, B, H, x0_host, theta_host, theta_stride, U_host, Xref_host, Uref_host, loss_dp_host, cost_host,
                         keep_dtraj, nullptr, nullptr, nullptr, nullptr, workspace, ws_bytes, stream);
}

int pdp_sweep_host_traj(pdp_system_t* sys, int B, int H, const double* x0_host, const double* theta_host,
                        int theta_stride, const double* U_host, const double* Xref_host, const double* Uref_host,
                        double* loss_dp_host, double* cost_host, double* X_host, double* Lam_host, double* dX_host,
                        double* dU_host, void* workspace, size_t ws_bytes, pdp_stream_t stream) {
  return sweep_host_impl(sys, B, H, x0_host, theta_host, theta_stride, U_host, Xref_host, Uref_host, loss_dp_host, cost_host,
                         1, X_host, Lam_host, dX_host, dU_host, workspace, ws_bytes, stream);
}

// ---- host-buffer variants of the two single-kernel modes (end-to-end use: copies in, kernel, copies out) ----------------
int pdp_rollout_costate_host(pdp_system_t* sys, int B, int H, const double* x0_host, const double* theta_host,
                             int theta_stride, const double* U_host, double* cost_host, double* dHu_host, double* X_host,
                             double* Lam_host, void* workspace, size_t ws_bytes, pdp_stream_t stream) {
  if (!sys || !sys->rollout) return fail(PDP_ERR_UNSUPPORTED, "pdp_rollout_costate_host: module has no rollout kernel");
  if (B < 1 || H < 1 || !x0_host || !theta_host || !U_host)
    return fail(PDP_ERR_ARG, "pdp_rollout_costate_host: bad argument");
  if (!workspace || ws_bytes < pdp_workspace_bytes(sys, PDP_OP_ROLLOUT_HOST, B, H))
    return fail(PDP_ERR_WORKSPACE, "pdp_rollout_costate_host: workspace too small");
  const size_t n = sys->n(), m = sys->m(), nth = sys->info[11] > 0 ? sys->info[11] : sys->r();
  cudaStream_t st = (cudaStream_t)stream;
  char* p = reinterpret_cast<char*>(workspace);
  auto take = [&](size_t bytes) { char* q = p; p += align256(bytes); return reinterpret_cast<double*>(q); };
  double* d_x0 = take(size_t(B) * n * 8);
  double* d_th = take(size_t(B) * nth * 8);
  double* d_U = take(size_t(B) * H * m * 8);
  double* d_X = take(size_t(B) * (H + 1) * n * 8);
  double* d_L = take(size_t(B) * H * n * 8);
  double* d_cost = take(size_t(B) * 8);
  double* d_dHu = take(size_t(B) * H * m * 8);
  const size_t thn = theta_stride ? size_t(B) * theta_stride : nth;
  cudaError_t ce;
#define PDP_CK(x) if ((ce = (x)) != cudaSuccess) return fail(PDP_ERR_CUDA, "pdp_rollout_costate_host: %s", cudaGetErrorString(ce))
  PDP_CK(cudaMemcpyAsync(d_x0, x0_host, size_t(B) * n * 8, cudaMemcpyHostToDevice, st));
  PDP_CK(cudaMemcpyAsync(d_th, theta_host, thn * 8, cudaMemcpyHostToDevice, st));
  PDP_CK(cudaMemcpyAsync(d_U, U_host, size_t(B) * H * m * 8, cudaMemcpyHostToDevice, st));
  const bool need_lam = dHu_host || Lam_host;
  int e = pdp_rollout_costate(sys, B, H, d_x0, d_th, theta_stride, d_U, d_X, need_lam ? d_L : nullptr, d_cost,
                              dHu_host ? d_dHu : nullptr, nullptr, stream);
  if (e) return e;
  if (cost_host) PDP_CK(cudaMemcpyAsync(cost_host, d_cost, size_t(B) * 8, cudaMemcpyDeviceToHost, st));
  if (dHu_host) PDP_CK(cudaMemcpyAsync(dHu_host, d_dHu, size_t(B) * H * m * 8, cudaMemcpyDeviceToHost, st));
  if (X_host) PDP_CK(cudaMemcpyAsync(X_host, d_X, size_t(B) * (H + 1) * n * 8, cudaMemcpyDeviceToHost, st));
  if (Lam_host) PDP_CK(cudaMemcpyAsync(Lam_host, d_L, size_t(B) * H * n * 8, cudaMemcpyDeviceToHost, st));
#undef PDP_CK
  return PDP_OK;
}

int pdp_sens_fwd_host(pdp_system_t* sys, int B, int H, const double* x0_host, const double* theta_host, int theta_stride,
                      const double* inputs_host, const double* Xobs_host, double* loss_dp_host, double* sums_host,
                      void* workspace, size_t ws_bytes, pdp_stream_t stream) {
  if (!sys || !sys->sens) return fail(PDP_ERR_UNSUPPORTED, "pdp_sens_fwd_host: module has no forward-sensitivity kernel");
  if (B < 1 || H < 1 || !x0_host || !theta_host || (!loss_dp_host && !sums_host))
    return fail(PDP_ERR_ARG, "pdp_sens_fwd_host: bad argument");
  if (sys->kind() == PDP_KIND_SYSID && (!inputs_host || !Xobs_host))
    return fail(PDP_ERR_ARG, "pdp_sens_fwd_host: a SysID module needs inputs and observed states");
  if (!workspace || ws_bytes < pdp_workspace_bytes(sys, PDP_OP_SENS_HOST, B, H))
    return fail(PDP_ERR_WORKSPACE, "pdp_sens_fwd_host: workspace too small");
  const size_t n = sys->n(), m = sys->m(), r = sys->r();
  cudaStream_t st = (cudaStream_t)stream;
  char* p = reinterpret_cast<char*>(workspace);
  auto take = [&](size_t bytes) { char* q = p; p += align256(bytes); return reinterpret_cast<double*>(q); };
  // the reduction scratch comes first: its ticket word must be zero when the workspace is first used
  void* red = take(pdp_reduce_workspace_bytes((int)r));
  double* d_x0 = take(size_t(B) * n * 8);
  double* d_th = take(size_t(B) * r * 8);
  double* d_in = take(size_t(B) * H * m * 8);
  double* d_Xo = take(size_t(B) * (H + 1) * n * 8);
  double* d_ldp = take(size_t(B) * (r + 1) * 8);
  double* d_sums = take((r + 2) * 8);
  const size_t thn = theta_stride ? size_t(B) * r : r;
  cudaError_t ce;
#define PDP_CK(x) if ((ce = (x)) != cudaSuccess) return fail(PDP_ERR_CUDA, "pdp_sens_fwd_host: %s", cudaGetErrorString(ce))
  PDP_CK(cudaMemcpyAsync(d_x0, x0_host, size_t(B) * n * 8, cudaMemcpyHostToDevice, st));
  PDP_CK(cudaMemcpyAsync(d_th, theta_host, thn * 8, cudaMemcpyHostToDevice, st));
  if (inputs_host) PDP_CK(cudaMemcpyAsync(d_in, inputs_host, size_t(B) * H * m * 8, cudaMemcpyHostToDevice, st));
  if (Xobs_host) PDP_CK(cudaMemcpyAsync(d_Xo, Xobs_host, size_t(B) * (H + 1) * n * 8, cudaMemcpyHostToDevice, st));
  int e = pdp_sens_fwd(sys, B, H, d_x0, d_th, theta_stride, inputs_host ? d_in : nullptr, Xobs_host ? d_Xo : nullptr,
                       nullptr, nullptr, nullptr, nullptr, d_ldp, nullptr, stream);
  if (e) return e;
  if (loss_dp_host) PDP_CK(cudaMemcpyAsync(loss_dp_host, d_ldp, size_t(B) * (r + 1) * 8, cudaMemcpyDeviceToHost, st));
  if (sums_host) {
    e = pdp_reduce_loss_dp(B, (int)r, d_ldp, d_sums, red, pdp_reduce_workspace_bytes((int)r), stream);
    if (e) return e;
    PDP_CK(cudaMemcpyAsync(sums_host, d_sums, (r + 2) * 8, cudaMemcpyDeviceToHost, st));
  }
#undef PDP_CK
  return PDP_OK;
}

}  // extern "C"
