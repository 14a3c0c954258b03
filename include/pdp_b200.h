/* pdp_b200.h -- C ABI of the B200-native batched PDP engine (libpdp_b200.so).
 *
 * The reference (wanxinjin/Pontryagin-Differentiable-Programming) has no FFI: its boundary is the
 * Python class surface of PDP/PDP.py.  Each entry point below is what a binding of that surface
 * calls for the hot path; the reference method it replaces is cited per function.
 *
 * Conventions
 *   - every array pointer is a DEVICE pointer (float64, row-major, contiguous, batch-major) unless
 *     the parameter name ends in _host (pinned or pageable host memory);
 *   - the caller owns every buffer including workspaces; the library allocates nothing per call;
 *   - launches are asynchronous on `stream`; no implicit device synchronisation;
 *   - return value 0 = ok, <0 = error (pdp_last_error() gives a thread-local message); no exceptions,
 *     no exit();
 *   - status[b] (int32, optional, caller zero-initialises) collects per-trajectory flags:
 *     bit 0 = non-finite value encountered, bit 1 = Quu not positive definite in the Riccati sweep.
 *   - theta_stride = r for per-trajectory parameters, 0 when one parameter vector is shared.
 */
#ifndef PDP_B200_H
#define PDP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pdp_system pdp_system_t;
typedef void* pdp_stream_t; /* cudaStream_t */

enum { PDP_OK = 0, PDP_ERR_ARG = -1, PDP_ERR_LOAD = -2, PDP_ERR_CUDA = -3, PDP_ERR_WORKSPACE = -4, PDP_ERR_UNSUPPORTED = -5 };
enum { PDP_KIND_OC = 1, PDP_KIND_SYSID = 2, PDP_KIND_CP = 3, PDP_KIND_LQR = 4, PDP_KIND_FUNCTION = 5 };
enum { PDP_OP_AUX_LQR = 1, PDP_OP_SWEEP = 2, PDP_OP_SWEEP_HOST = 3, PDP_OP_ROLLOUT_HOST = 4, PDP_OP_SENS_HOST = 5 };

/* Load a generated system module (the product of OCSys.setDyn/setPathCost/setFinalCost + diffPMP,
 * reference PDP/PDP.py:96-119,222-270: here "differentiate PMP" = code-generate + nvcc). */
int pdp_load_system(const char* module_path, pdp_system_t** out);
void pdp_free_system(pdp_system_t* sys);
/* dims[0..3] = kind, n_state, n_control, n_auxvar */
int pdp_system_dims(const pdp_system_t* sys, int* dims);
const char* pdp_last_error(void);
const char* pdp_version(void);

/* Bytes of caller-provided workspace an operation needs (0 if none). */
size_t pdp_workspace_bytes(const pdp_system_t* sys, int op, int B, int H);

/* Forward rollout + cost + costate recursion at given controls; optional dH/du (adjoint gradient).
 * Replaces the rollout/costate semantics of OCSys.ocSolver (PDP/PDP.py:158-175, 196-209) and, with
 * dHu != NULL, ControlPlanning.recmat_step's gradient (PDP/PDP.py:1100-1114).
 *   x0[B,n] theta[B|1,r] U[B,H,m] -> X[B,H+1,n] Lam[B,H,n] (Lam[t] = lambda_{t+1}) cost[B] dHu[B,H,m]
 * Lam, cost, dHu may be NULL. */
int pdp_rollout_costate(pdp_system_t* sys, int B, int H, const double* x0, const double* theta, int theta_stride,
                        const double* U, double* X, double* Lam, double* cost, double* dHu, int* status,
                        pdp_stream_t stream);

/* Closed-loop rollout used by the batched ocSolver's line search (replaces IPOPT's iterations inside
 * OCSys.ocSolver, PDP/PDP.py:178-182): u_t = Uref[t] + alpha_b k_t + K_t (x_t - Xref[t]) with the gains
 * (K_t | k_t) of a one-column Riccati sweep, laid out [B,H,n+1,m] as pdp_aux_lqr_backward leaves them in its
 * workspace.  Writes the applied controls to Uout[B,H,m] and X / Lam / cost / dHu like pdp_rollout_costate.
 * group > 1: B = group x (source trajectories); candidate b reads x0/theta/Uref/Xref/gains of trajectory b/group and
 * its own alpha[b] (a whole back-tracking line search in one launch); outputs are indexed by b. */
int pdp_rollout_feedback(pdp_system_t* sys, int B, int H, const double* x0, const double* theta, int theta_stride,
                         const double* Uref, const double* Xref, const double* gains, const double* alpha, double* Uout,
                         double* X, double* Lam, double* cost, double* dHu, int group, int* status, pdp_stream_t stream);

/* Fused OCSys.getAuxSys (PDP/PDP.py:272-314) + LQR.lqrSolver (PDP/PDP.py:446-615).
 *   X,U,Lam,theta as above; X0aux[B|1,n,r] initial condition of the auxiliary system (NULL = zeros,
 *   x0aux_stride 0 = shared); outputs dXdtheta[B,H+1,n,r], dUdtheta[B,H,m,r] (either may be NULL).
 *   Optional fused IRL loss/chain rule (Examples/IRL/quadrotor/uav_PDP.py:67-75): Xref[B,H+1,n],
 *   Uref[B,H,m] -> loss_dp[B,r+1] = (loss_b, dp_b[0..r-1]) per trajectory (dp is 1/2 grad, like the
 *   reference).  workspace: pdp_workspace_bytes(sys, PDP_OP_AUX_LQR, B, H). */
int pdp_aux_lqr(pdp_system_t* sys, int B, int H, const double* X, const double* U, const double* Lam,
                const double* theta, int theta_stride, const double* X0aux, int x0aux_stride,
                double* dXdtheta, double* dUdtheta, const double* Xref, const double* Uref, double* loss_dp,
                void* workspace, size_t ws_bytes, int* status, pdp_stream_t stream);

/* The two halves of pdp_aux_lqr as separate calls.  pdp_aux_lqr_backward runs the Riccati sweep only and
 * leaves the gains (K_t | k_t) in `workspace`; pdp_aux_lqr_forward consumes them (e.g. again with
 * different Xref/Uref or X0aux, without repeating the sweep). */
int pdp_aux_lqr_backward(pdp_system_t* sys, int B, int H, const double* X, const double* U, const double* Lam,
                         const double* theta, int theta_stride, void* workspace, size_t ws_bytes, int* status,
                         pdp_stream_t stream);
int pdp_aux_lqr_forward(pdp_system_t* sys, int B, int H, const double* X, const double* U, const double* theta,
                        int theta_stride, const double* X0aux, int x0aux_stride, double* dXdtheta, double* dUdtheta,
                        const double* Xref, const double* Uref, double* loss_dp, const void* workspace, size_t ws_bytes,
                        int* status, pdp_stream_t stream);

/* LQR.lqrSolver on caller-supplied matrices (PDP/PDP.py:446-615) for kind PDP_KIND_LQR modules of size
 * (n, m, r):  aux[B,H,NDENSE] in the per-step layout of pdp_aux_eval, term[B, n*n+n*r] = [hxx|hxe],
 * X0aux as in pdp_aux_lqr -> Xaux[B,H+1,n,r], Uaux[B,H,m,r].  Hessian blocks must be symmetric
 * (Hux is ignored and transpose(Hxu) used instead, exactly as the reference does). */
int pdp_lqr_dense(pdp_system_t* sys, int B, int H, const double* aux, const double* term, const double* X0aux,
                  int x0aux_stride, double* Xaux, double* Uaux, int forward_only, void* workspace, size_t ws_bytes,
                  int* status, pdp_stream_t stream);
/* forward_only != 0: skip the Riccati sweep and run only X+ = F X + G (K X + k) + E with gains the caller has
 * written into `workspace` as [B,H,(n+r),m] records (rows 0..n-1 = columns of K, rows n.. = columns of k):
 * ControlPlanning.integrateAuxSys (PDP/PDP.py:813-838, K = dUx, k = dUe, E = 0) and
 * SysID.integrateAuxSys (PDP/PDP.py:1241-1259, gains = 0). */

/* Evaluate a code-generated symbolic Function (kind 5 module) for B samples: inputs[k] -> [B|1, numel_k]
 * (input_strides[k] = numel_k or 0 when shared), outputs[k] -> [B, rows_k*cols_k] row-major.  Replaces the
 * per-step CasADi calls of ControlPlanning.getAuxSys / SysID.getAuxSys (PDP/PDP.py:788-811, 1225-1239). */
int pdp_eval_function(pdp_system_t* sys, int B, const double* const* inputs, const int* input_strides,
                      double* const* outputs, pdp_stream_t stream);

/* One full PDP sweep = pdp_rollout_costate + pdp_aux_lqr on device buffers (the BASELINE metric's unit).
 * For large batches the aux-LQR phase is cut into sub-batches that alternate between two internal streams of the
 * system (forked from / joined back into `stream` with events, so the call keeps its stream-ordered semantics and can
 * be captured into a CUDA graph once the internal streams exist): the shared-memory-bound backward kernel of one
 * sub-batch overlaps the HBM-bound forward kernel of the other.  Results are identical to the unsplit call. */
int pdp_sweep(pdp_system_t* sys, int B, int H, const double* x0, const double* theta, int theta_stride,
              const double* U, double* X, double* Lam, double* cost, double* dXdtheta, double* dUdtheta,
              const double* Xref, const double* Uref, double* loss_dp, void* workspace, size_t ws_bytes,
              int* status, pdp_stream_t stream);

/* Number of sub-batches pdp_sweep cuts its aux-LQR phase into: 0 = automatic (4 from 16384 trajectories, 2 from 8192,
 * else 1), 1 = never split.  Concurrent pdp_sweep calls on one system serialise their (short) enqueue sequence. */
int pdp_set_sweep_parts(pdp_system_t* sys, int parts);

/* Dense auxiliary matrices for the legacy OCSys.getAuxSys return value (PDP/PDP.py:303-313).
 *   aux[B,H,NDENSE] with per-step layout [F n*n|G n*m|E n*r|Hxx|Hxu|Hxe|Hux m*n|Huu|Hue] row-major,
 *   term[B, n*n + n*r] = [hxx|hxe]. */
int pdp_aux_eval(pdp_system_t* sys, int B, int H, const double* X, const double* U, const double* Lam,
                 const double* theta, int theta_stride, double* aux, double* term, pdp_stream_t stream);

/* Forward-sensitivity sweeps (ControlPlanning.step PDP/PDP.py:850-878; SysID.step :1261-1296):
 * kind SYSID: args = (x0[B,n], theta[B|1,r], inputs[B,H,m], Xobs[B,H+1,n] or NULL)
 * kind CP   : args = (x0[B,n], theta[B|1,r])  (policy parameters), rollout under the policy.
 * outputs X[B,H+1,n], Uout[B,H,m] (CP), dX[B,H+1,n,r], dU[B,H,m,r] (CP), loss_dp[B,r+1]; any may be NULL. */
int pdp_sens_fwd(pdp_system_t* sys, int B, int H, const double* x0, const double* theta, int theta_stride,
                 const double* inputs, const double* Xobs, double* X, double* Uout, double* dX, double* dU,
                 double* loss_dp, int* status, pdp_stream_t stream);

/* Host-buffer variant of pdp_sweep for end-to-end use: copies x0/theta/U (and Xref/Uref) from host,
 * runs the sweep with the fused loss, copies loss_dp[B,r+1] (and cost[B] if cost_host != NULL) back.
 * All device scratch comes from `workspace` (pdp_workspace_bytes(sys, PDP_OP_SWEEP_HOST, B, H)). */
int pdp_sweep_host(pdp_system_t* sys, int B, int H, const double* x0_host, const double* theta_host,
                   int theta_stride, const double* U_host, const double* Xref_host, const double* Uref_host,
                   double* loss_dp_host, double* cost_host, int keep_dtraj, void* workspace, size_t ws_bytes,
                   pdp_stream_t stream);

/* pdp_sweep_host that also returns the trajectories and the sensitivities themselves to the host (any of X_host[B,H+1,n],
 * Lam_host[B,H,n], dX_host[B,H+1,n,r], dU_host[B,H,m,r] may be NULL): the return values of OCSys.ocSolver's rollout and of
 * LQR.lqrSolver (PDP/PDP.py:212-218, 611-615) for a whole batch. */
int pdp_sweep_host_traj(pdp_system_t* sys, int B, int H, const double* x0_host, const double* theta_host,
                        int theta_stride, const double* U_host, const double* Xref_host, const double* Uref_host,
                        double* loss_dp_host, double* cost_host, double* X_host, double* Lam_host, double* dX_host,
                        double* dU_host, void* workspace, size_t ws_bytes, pdp_stream_t stream);

/* Batch reduction of per-trajectory (loss, dp) rows: loss_dp[B,r+1] -> sums[r+2] = (sum loss, sum dp[0..r-1], B).
 * This is the averaging step of the outer loops (reference PDP/PDP.py:1293-1294, Examples/IRL/quadrotor/uav_PDP.py:78-81)
 * up to the division, and the vector the single all-reduce of a multi-GPU run carries.  Deterministic (fixed summation
 * order, no floating-point atomics).  workspace: pdp_reduce_workspace_bytes(r) bytes, ZERO-INITIALISED by the caller
 * before its first use (every call leaves it reusable); one workspace per concurrently used stream. */
size_t pdp_reduce_workspace_bytes(int r);
int pdp_reduce_loss_dp(int B, int r, const double* loss_dp, double* sums, void* workspace, size_t ws_bytes,
                       pdp_stream_t stream);

/* Host-buffer variant of pdp_rollout_costate (ControlPlanning.recmat_step for a batch, PDP/PDP.py:1100-1114): copies
 * x0/theta/U from host, runs rollout + costate (+ dH/du), copies cost[B] / dHu[B,H,m] / X / Lam back (any may be NULL).
 * workspace: pdp_workspace_bytes(sys, PDP_OP_ROLLOUT_HOST, B, H). */
int pdp_rollout_costate_host(pdp_system_t* sys, int B, int H, const double* x0_host, const double* theta_host,
                             int theta_stride, const double* U_host, double* cost_host, double* dHu_host, double* X_host,
                             double* Lam_host, void* workspace, size_t ws_bytes, pdp_stream_t stream);

/* Host-buffer variant of pdp_sens_fwd with the fused loss (SysID.step / ControlPlanning.step for a batch, PDP/PDP.py:1261-1296,
 * 850-878): copies x0/theta (and inputs/Xobs for SysID) from host, runs the sweep, copies loss_dp[B,r+1] and/or its batch
 * reduction sums[r+2] back (either may be NULL).  workspace: pdp_workspace_bytes(sys, PDP_OP_SENS_HOST, B, H), zero-initialised
 * before its first use (it holds the reduction scratch). */
int pdp_sens_fwd_host(pdp_system_t* sys, int B, int H, const double* x0_host, const double* theta_host, int theta_stride,
                      const double* inputs_host, const double* Xobs_host, double* loss_dp_host, double* sums_host,
                      void* workspace, size_t ws_bytes, pdp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PDP_B200_H */
