"""Hand-written CUDA kernel templates of the OC / LQR modules.

``codegen.py`` fills the ``@@...@@`` holes with generated straight-line code (structural non-zeros of the
system at hand) and prepends the ``PDP_*`` constants.  See DESIGN.md for the kernel designs."""

K_PRELUDE = r'''
// Software prefetch of rows that a later chunk / step will read (PDP_PF: 0 off, 1 into L1, 2 into L2): used by the forward
// Riccati kernel for its gain records and chunk rows (measured: -6 %; in the rollout kernel the same instructions cost
// 10 % and in the backward kernel they change nothing, so those kernels do not prefetch).
__device__ __forceinline__ void pdp_prefetch(const void* p) {
#if PDP_PF == 1
  asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
#elif PDP_PF == 2
  asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
#endif
}
__device__ __forceinline__ void pdp_prefetch_l1(const void* p) {
#if PDP_PF
  asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
#endif
}

'''

K_ROW_IO = r'''
// Row I/O of the thread-per-trajectory kernel.  A thread's 8-byte accesses each cost one 32-byte sector request;
// rows are only 8-byte aligned (n doubles per row), so pick the 16-byte pairing that matches the row's parity.
template <int LEN>
__device__ __forceinline__ void pdp_row_store(double* __restrict__ g, const double* x) {
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    #pragma unroll
    for (int i = 0; i + 1 < LEN; i += 2) *reinterpret_cast<double2*>(g + i) = make_double2(x[i], x[i + 1]);
    if (LEN & 1) g[LEN - 1] = x[LEN - 1];
  } else {
    g[0] = x[0];
    #pragma unroll
    for (int i = 1; i + 1 < LEN; i += 2) *reinterpret_cast<double2*>(g + i) = make_double2(x[i], x[i + 1]);
    if (!(LEN & 1)) g[LEN - 1] = x[LEN - 1];
  }
}

template <int LEN>
__device__ __forceinline__ void pdp_row_load(double* x, const double* g) {
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    #pragma unroll
    for (int i = 0; i + 1 < LEN; i += 2) { const double2 v = *reinterpret_cast<const double2*>(g + i); x[i] = v.x; x[i + 1] = v.y; }
    if (LEN & 1) x[LEN - 1] = g[LEN - 1];
  } else {
    x[0] = g[0];
    #pragma unroll
    for (int i = 1; i + 1 < LEN; i += 2) { const double2 v = *reinterpret_cast<const double2*>(g + i); x[i] = v.x; x[i + 1] = v.y; }
    if (!(LEN & 1)) x[LEN - 1] = g[LEN - 1];
  }
}

'''

K_TMA_PRIMS = r'''
// ---- TMA (1-D bulk async copy) and mbarrier primitives; in the CPU emulator a bulk copy is an immediate memcpy ----------
__device__ __forceinline__ unsigned pdp_smem_u32(const void* p) {
#ifdef __CUDACC__
  return (unsigned)__cvta_generic_to_shared(p);
#else
  return 0u;
#endif
}
__device__ __forceinline__ void pdp_mbar_init(double* mbar) {
#ifdef __CUDACC__
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(pdp_smem_u32(mbar)) : "memory");
#endif
}
__device__ __forceinline__ void pdp_mbar_init_fence() {
#ifdef __CUDACC__
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdp_mbar_expect(double* mbar, unsigned bytes) {
#ifdef __CUDACC__
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(pdp_smem_u32(mbar)), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void pdp_mbar_wait(double* mbar, unsigned phase) {
#ifdef __CUDACC__
  unsigned done = 0;
  const unsigned a = pdp_smem_u32(mbar);
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(a), "r"(phase) : "memory");
  } while (!done);
#endif
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void pdp_bulk_g2s(double* dst_shared, const void* src, unsigned bytes, double* mbar) {
#ifdef __CUDACC__
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(pdp_smem_u32(dst_shared)), "l"(src), "r"(bytes), "r"(pdp_smem_u32(mbar)) : "memory");
#else
  memcpy(dst_shared, src, bytes);        /* CPU emulation: immediate copy */
#endif
}
__device__ __forceinline__ void pdp_bulk_s2g(void* dst, const double* src_shared, unsigned bytes) {
#ifdef __CUDACC__
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(pdp_smem_u32(src_shared)), "r"(bytes) : "memory");
#else
  memcpy(dst, src_shared, bytes);
#endif
}
__device__ __forceinline__ void pdp_bulk_store_fence() {     // generic-proxy writes to shared memory -> visible to the bulk store
#ifdef __CUDACC__
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdp_bulk_commit() {
#ifdef __CUDACC__
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdp_bulk_wait_read() {       // the committed stores have finished READING shared memory
#ifdef __CUDACC__
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}
// A planned load of `count` doubles starting at g (8-byte aligned) into a 16-byte aligned slot: the copy starts at the aligned
// address at or below g, so the data sits `off` (0 or 1) doubles into the slot.  A copy that would read past `end` (the end of
// the tensor) is shortened by 16 bytes; the one or two doubles it leaves out are fetched with plain loads at issue time.
struct pdp_tma_plan { const double* src; unsigned bytes; int off, total; };
__device__ __forceinline__ pdp_tma_plan pdp_tma_plan_load(const double* g, int count, const double* end) {
  pdp_tma_plan p;
  p.off = (int)((reinterpret_cast<uintptr_t>(g) >> 3) & 1);
  p.src = g - p.off;
  p.total = p.off + count;
  p.bytes = (unsigned)((p.total * 8 + 15) & ~15);
  if (p.src + p.bytes / 8 > end) p.bytes -= 16;
  return p;
}
__device__ __forceinline__ void pdp_tma_issue_load(double* slot, const pdp_tma_plan& p, double* mbar) {
  for (int e = (int)(p.bytes / 8); e < p.total; ++e) slot[e] = p.src[e];       // only at the very end of a tensor
  if (p.bytes) pdp_bulk_g2s(slot, p.src, p.bytes, mbar);
}
// Send `count` doubles that sit in `slot` at offset `off` (= parity of g) to g: aligned middle as one bulk copy, the
// possible first / last element with plain stores.
__device__ __forceinline__ void pdp_tma_send(const double* slot, int off, double* g, int count) {
  int first = 0;
  if (off) { g[0] = slot[off]; first = 1; }
  const int mid = (count - first) & ~1;
  if (mid > 0) pdp_bulk_s2g(g + first, slot + off + first, (unsigned)mid * 8);
  if (first + mid < count) g[count - 1] = slot[off + count - 1];
}

'''

K_ROLLOUT_AUXEVAL = r'''
// =====================================================================================================
// Kernel 1: forward rollout + cost + costate recursion (+ optional dH/du), one thread per trajectory.
//   restates reference OCSys.ocSolver's rollout semantics at given controls (PDP.py:158-175) and the PMP
//   costate recursion (PDP.py:203-209): Lam[t] = lambda_{t+1}, lambda_H = dh/dx(x_H).
// =====================================================================================================

''' + K_ROW_IO + r'''
// Two-stage form of pdp_row_load for PREFETCHED rows: `issue` puts the row into a raw buffer with the same loads for
// both address parities (16-byte pairs starting at element `par`, the one or two left-over elements as scalars), and
// `unpack` sorts the buffer into x[] with selects when the row is consumed one step later.  (With pdp_row_load the
// two parity paths fill different registers and nvcc merges them with MOVs right behind the loads -- which wait
// for the data at once: 60 % of the rollout kernel's stall samples in the ncu source view.)
template <int LEN>
struct pdp_row_raw { double2 p[(LEN - 1) / 2 > 0 ? (LEN - 1) / 2 : 1]; double s0, s1; int par; };

template <int LEN>
__device__ __forceinline__ void pdp_row_issue(pdp_row_raw<LEN>& r, const double* g) {
  constexpr int NP = (LEN - 1) / 2;
  const int par = (int)((reinterpret_cast<uintptr_t>(g) >> 3) & 1);
  r.par = par;
  #pragma unroll
  for (int k = 0; k < NP; ++k) r.p[k] = *reinterpret_cast<const double2*>(g + par + 2 * k);
  if (LEN & 1) { r.s0 = g[par ? 0 : LEN - 1]; r.s1 = 0.0; }
  else { r.s0 = g[par ? 0 : LEN - 2]; r.s1 = g[LEN - 1]; }
}

template <int LEN>
__device__ __forceinline__ void pdp_row_unpack(double* x, const pdp_row_raw<LEN>& r) {
  constexpr int NP = (LEN - 1) / 2;
  const bool odd = r.par != 0;
  #pragma unroll
  for (int i = 0; i < LEN; ++i) {
    // candidate of the aligned layout / of the layout shifted by one element
    double c0, c1;
    if (i < 2 * NP) c0 = (i & 1) ? r.p[i / 2].y : r.p[i / 2].x;
    else c0 = (i == LEN - 1 && !(LEN & 1)) ? r.s1 : r.s0;
    if (i == 0) c1 = r.s0;
    else if (i <= 2 * NP) c1 = ((i - 1) & 1) ? r.p[(i - 1) / 2].y : r.p[(i - 1) / 2].x;
    else c1 = r.s1;
    x[i] = odd ? c1 : c0;
  }
}

extern "C" __global__ void __launch_bounds__(128)
pdp_k_rollout_costate(int B, int H, const double* __restrict__ x0, const double* __restrict__ theta, int theta_stride,
                      const double* __restrict__ U, double* __restrict__ X, double* __restrict__ Lam,
                      double* __restrict__ cost, double* __restrict__ dHu, int* __restrict__ status,
                      const double* __restrict__ fb_gains, const double* __restrict__ fb_X,
                      const double* __restrict__ fb_alpha, double* __restrict__ Uout, int fb_group)
{
  // Optional closed-loop mode (batched ocSolver line search): with fb_gains != NULL the applied control is
  //   u_t = U[t] + alpha_b * k_t + K_t (x_t - fb_X[t])   (gains in the (K|k) record layout of the Riccati
  // sweep with one column) and is written to Uout.  With fb_group > 1 the launch holds fb_group candidates per source
  // trajectory (thread b reads x0 / theta / U / fb_X / fb_gains of trajectory b / fb_group and its own alpha): a whole
  // back-tracking line search in one launch.
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int bs = (fb_gains != nullptr && fb_group > 1) ? b / fb_group : b;     // source trajectory of this thread
  double x[PDP_N], xn[PDP_N], th[PDP_NTHX], u[PDP_M], tmp[1];
  #pragma unroll
  for (int i = 0; i < PDP_NTH; ++i) th[i] = theta[(size_t)bs * theta_stride + i];
#if PDP_NRCP > 0
  pdp_f_recips(th, th + PDP_NTH);       // reciprocals of the parameter-only divisors, once per trajectory
#endif
  #pragma unroll
  for (int i = 0; i < PDP_N; ++i) x[i] = x0[(size_t)bs * PDP_N + i];
  double J = 0.0;
  double* Xb = X + (size_t)b * (H + 1) * PDP_N;
  const double* Ub = U + (size_t)bs * H * PDP_M;
  const double fb_a = fb_gains ? fb_alpha[b] : 0.0;
  // software prefetch TWO steps ahead: two raw buffers used by alternate steps of a loop unrolled by two, so that no
  // register move (which would wait for the load) sits between the issue of a row and its use two steps later
  pdp_row_raw<PDP_M> una, unb;
  pdp_row_issue<PDP_M>(una, Ub);
  pdp_row_issue<PDP_M>(unb, Ub + (H > 1 ? 1 : 0) * PDP_M);
  auto fstep = [&](const int t, pdp_row_raw<PDP_M>& un) {
    pdp_row_unpack<PDP_M>(u, un);
    pdp_row_issue<PDP_M>(un, Ub + (t + 2 < H ? t + 2 : H - 1) * PDP_M);    // unconditional, index clamped
    if (fb_gains != nullptr) {
      const double* g = fb_gains + ((size_t)bs * H + t) * ((PDP_N + 1) * PDP_M);
      const double* xo = fb_X + ((size_t)bs * (H + 1) + t) * PDP_N;
      #pragma unroll
      for (int a = 0; a < PDP_M; ++a) u[a] = fma(fb_a, g[PDP_N * PDP_M + a], u[a]);
      #pragma unroll
      for (int l = 0; l < PDP_N; ++l) {
        const double dx = x[l] - xo[l];
        #pragma unroll
        for (int a = 0; a < PDP_M; ++a) u[a] = fma(g[l * PDP_M + a], dx, u[a]);
      }
      #pragma unroll
      for (int i = 0; i < PDP_M; ++i) Uout[((size_t)b * H + t) * PDP_M + i] = u[i];
    }
    pdp_row_store<PDP_N>(Xb + t * PDP_N, x);
    pdp_f_path_cost(x, u, th, tmp);
    J += tmp[0];
    pdp_f_dyn(x, u, th, xn);
    #pragma unroll
    for (int i = 0; i < PDP_N; ++i) x[i] = xn[i];
  };
  #pragma unroll 1
  for (int t = 0; t < H; t += 2) {
    fstep(t, una);
    if (t + 1 < H) fstep(t + 1, unb);
  }
  pdp_row_store<PDP_N>(Xb + H * PDP_N, x);
  pdp_f_final_cost(x, th, tmp);
  J += tmp[0];
  if (cost) cost[b] = J;
  bool bad = !isfinite(J);
  if (Lam != nullptr) {
    double lam[PDP_N], ln[PDP_N], gu[PDP_M];
    double* Lb = Lam + (size_t)b * H * PDP_N;
    pdp_f_dhx(x, th, lam);
    const double* Ua = fb_gains ? Uout + (size_t)b * H * PDP_M : Ub;      // the controls actually applied
    // (x_{t-2}, u_{t-2}) are fetched while step t is being processed: two buffer pairs, loop unrolled by two
    pdp_row_raw<PDP_N> xpa, xpb;
    pdp_row_raw<PDP_M> upa, upb;
    pdp_row_issue<PDP_N>(xpa, Xb + (H - 1) * PDP_N);
    pdp_row_issue<PDP_M>(upa, Ua + (H - 1) * PDP_M);
    pdp_row_issue<PDP_N>(xpb, Xb + (H > 1 ? H - 2 : 0) * PDP_N);
    pdp_row_issue<PDP_M>(upb, Ua + (H > 1 ? H - 2 : 0) * PDP_M);
    auto bstep = [&](const int t, pdp_row_raw<PDP_N>& xp, pdp_row_raw<PDP_M>& up) {
      pdp_row_store<PDP_N>(Lb + t * PDP_N, lam);
      pdp_row_unpack<PDP_N>(x, xp);
      pdp_row_unpack<PDP_M>(u, up);
      {
        const int tq = t > 1 ? t - 2 : 0;       // unconditional, index clamped
        pdp_row_issue<PDP_N>(xp, Xb + tq * PDP_N);
        pdp_row_issue<PDP_M>(up, Ua + tq * PDP_M);
      }
      if (dHu != nullptr) {
        pdp_f_dHu(x, u, lam, th, gu);
        #pragma unroll
        for (int i = 0; i < PDP_M; ++i) dHu[((size_t)b * H + t) * PDP_M + i] = gu[i];
      }
      if (t > 0) {
        pdp_f_dHx(x, u, lam, th, ln);
        #pragma unroll
        for (int i = 0; i < PDP_N; ++i) lam[i] = ln[i];
      }
    };
    #pragma unroll 1
    for (int t = H - 1; t >= 0; t -= 2) {
      bstep(t, xpa, upa);
      if (t > 0) bstep(t - 1, xpb, upb);
    }
  }
  if (status && bad) atomicOr(&status[b], 1);
}

// =====================================================================================================
// Kernel 2: dense auxiliary-system matrices (legacy getAuxSys API, PDP.py:272-314); one thread per (b, t).
//   out layout per (b,t): [F n*n | G n*m | E n*r | Hxx | Hxu | Hxe | Hux | Huu | Hue], each row-major.
//   term layout per b   : [hxx n*n | hxe n*r]
// =====================================================================================================
extern "C" __global__ void __launch_bounds__(128)
pdp_k_aux_eval(int B, int H, const double* __restrict__ X, const double* __restrict__ U, const double* __restrict__ Lam,
               const double* __restrict__ theta, int theta_stride, double* __restrict__ out, double* __restrict__ term)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * (H + 1)) return;
  const int b = idx / (H + 1), t = idx - b * (H + 1);
  double th[PDP_NTHX];
  #pragma unroll
  for (int i = 0; i < PDP_NTH; ++i) th[i] = theta[(size_t)b * theta_stride + i];
#if PDP_NRCP > 0
  pdp_f_recips(th, th + PDP_NTH);
#endif
  if (t == H) {
    if (term) pdp_f_terminal(X + ((size_t)b * (H + 1) + H) * PDP_N, th, term + (size_t)b * (PDP_N * PDP_N + PDP_N * PDP_R));
    return;
  }
  pdp_f_aux_dense(X + ((size_t)b * (H + 1) + t) * PDP_N, U + ((size_t)b * H + t) * PDP_M, Lam + ((size_t)b * H + t) * PDP_N, th,
                  out + ((size_t)b * H + t) * PDP_NDENSE);
}

'''

# ---- open-loop rollout / costate kernel with per-thread TMA (1-D bulk async copies) for all row traffic ------------------
K_ROLLOUT_TMA = r'''
// =====================================================================================================
// Kernel 1t: the open-loop rollout / costate kernel with all row traffic on the TMA engine.
//   Same thread-per-trajectory arithmetic as pdp_k_rollout_costate.  What changes is how rows move: the rows a thread
//   needs (u_t forward; x_t, u_t backward) and produces (x_t forward; lambda_{t+1}, dH/du_t backward) for PDP_TC consecutive
//   time steps are contiguous in HBM, so each thread moves them with ONE 1-D bulk async copy per array and chunk
//   (cp.async.bulk, completion on a thread-private mbarrier) between HBM and a thread-private shared-memory slot: no
//   per-lane sector requests in the load-store unit (the limiter of the register-prefetch kernel: 0.8 L1 sector requests
//   per cycle per SM, long_scoreboard 5.3 of 7.6 stall cycles per issue), no extra instructions on the critical path, loads
//   one chunk ahead.  Bulk copies need 16-byte aligned addresses and sizes while rows are only 8-byte aligned (n = 13
//   doubles): a load starts at the aligned address below the chunk (the data then sits 0 or 8 bytes into its slot), a
//   store sends the aligned middle as a bulk copy and the possible first / last element with plain stores.
// =====================================================================================================
extern "C" __global__ void __launch_bounds__(PDP_TB)
pdp_k_rollout_costate_tma(int B, int H, const double* __restrict__ x0, const double* __restrict__ theta, int theta_stride,
                          const double* __restrict__ U, double* __restrict__ X, double* __restrict__ Lam,
                          double* __restrict__ cost, double* __restrict__ dHu, int* __restrict__ status)
{
  extern __shared__ __align__(16) double pdp_smem[];
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;                                   // no block-level synchronisation anywhere: every thread is on its own
  // thread-private slots (PDP_TSTRIDE doubles apart): 2 x [x rows | u rows] in (the next chunk streams in while this one is
  // computed), [x / lambda rows] and [dH/du rows] out (double-buffering these as well measured no different), 2 mbarriers
  double* my = pdp_smem + (size_t)threadIdx.x * PDP_TSTRIDE;
  auto XS = [&](int k) { return my + k * (PDP_TXS + PDP_TUS); };
  auto US = [&](int k) { return my + k * (PDP_TXS + PDP_TUS) + PDP_TXS; };
  double* LS = my + 2 * (PDP_TXS + PDP_TUS);
  double* GS = LS + PDP_TLS;
  double* MB = GS + PDP_TUS;                            // two 8-byte mbarriers
  pdp_mbar_init(MB);
  pdp_mbar_init(MB + 1);
  pdp_mbar_init_fence();
  unsigned phase0 = 0, phase1 = 0;
  double x[PDP_N], xn[PDP_N], th[PDP_NTHX], u[PDP_M], tmp[1];
  #pragma unroll
  for (int i = 0; i < PDP_NTH; ++i) th[i] = theta[(size_t)b * theta_stride + i];
#if PDP_NRCP > 0
  pdp_f_recips(th, th + PDP_NTH);
#endif
  #pragma unroll
  for (int i = 0; i < PDP_N; ++i) x[i] = x0[(size_t)b * PDP_N + i];
  double J = 0.0;
  double* Xb = X + (size_t)b * (H + 1) * PDP_N;
  const double* Ub = U + (size_t)b * H * PDP_M;
  const double* Uend = U + (size_t)B * H * PDP_M;
  const double* Xend = X + (size_t)B * (H + 1) * PDP_N;
  // ---------------------------------------------------------------- forward: rollout and cost
  int offu0 = 0, offu1 = 0, offx0 = 0, offx1 = 0;
  {
    const pdp_tma_plan pu = pdp_tma_plan_load(Ub, (H < PDP_TC ? H : PDP_TC) * PDP_M, Uend);
    pdp_mbar_expect(MB, pu.bytes);
    pdp_tma_issue_load(US(0), pu, MB);
    offu0 = pu.off;
  }
  int buf = 0;
  #pragma unroll 1
  for (int t0 = 0; t0 < H; t0 += PDP_TC, buf ^= 1) {
    const int nst = H - t0 < PDP_TC ? H - t0 : PDP_TC;
    if (t0 + PDP_TC < H) {                               // next chunk's controls into the other slot
      const int nn = H - t0 - PDP_TC < PDP_TC ? H - t0 - PDP_TC : PDP_TC;
      const pdp_tma_plan pu = pdp_tma_plan_load(Ub + (size_t)(t0 + PDP_TC) * PDP_M, nn * PDP_M, Uend);
      pdp_mbar_expect(MB + (buf ^ 1), pu.bytes);
      pdp_tma_issue_load(US(buf ^ 1), pu, MB + (buf ^ 1));
      if (buf) offu0 = pu.off; else offu1 = pu.off;
    }
    if (buf == 0) { pdp_mbar_wait(MB, phase0); phase0 ^= 1; } else { pdp_mbar_wait(MB + 1, phase1); phase1 ^= 1; }
    pdp_bulk_wait_read();                                // the previous chunk's stores have left the output slot
    const double* us = US(buf) + (buf ? offu1 : offu0);
    double* Xg = Xb + (size_t)t0 * PDP_N;
    const int offl = (int)((reinterpret_cast<uintptr_t>(Xg) >> 3) & 1);
    double* ls = LS + offl;
    #pragma unroll 1
    for (int s = 0; s < nst; ++s) {
      #pragma unroll
      for (int i = 0; i < PDP_M; ++i) u[i] = us[s * PDP_M + i];
      #pragma unroll
      for (int i = 0; i < PDP_N; ++i) ls[s * PDP_N + i] = x[i];
      pdp_f_path_cost(x, u, th, tmp);
      J += tmp[0];
      pdp_f_dyn(x, u, th, xn);
      #pragma unroll
      for (int i = 0; i < PDP_N; ++i) x[i] = xn[i];
    }
    const bool last = t0 + PDP_TC >= H;
    if (last) {                                          // x_H rides with the last chunk (rows are contiguous in X)
      #pragma unroll
      for (int i = 0; i < PDP_N; ++i) ls[nst * PDP_N + i] = x[i];
    }
    pdp_bulk_store_fence();
    pdp_tma_send(LS, offl, Xg, (nst + (last ? 1 : 0)) * PDP_N);
    pdp_bulk_commit();
  }
  pdp_f_final_cost(x, th, tmp);
  J += tmp[0];
  if (cost) cost[b] = J;
  const bool bad = !isfinite(J);
  if (Lam != nullptr) {
    // ---------------------------------------------------------------- backward: costates (and dH/du)
    double lam[PDP_N], ln[PDP_N], gu[PDP_M];
    double* Lb = Lam + (size_t)b * H * PDP_N;
    pdp_f_dhx(x, th, lam);
    // the X rows written above are read back by this thread's own bulk loads: wait until the stores have completed
#ifdef __CUDACC__
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");      // the first / last elements went out as plain stores (generic proxy)
#endif
    const int tlast = ((H - 1) / PDP_TC) * PDP_TC;
    buf = 0;
    // which mbarrier phase each slot is in after the forward pass is tracked in phase0 / phase1; slot 0 is used first again
    {
      const int nn = H - tlast;
      const pdp_tma_plan px = pdp_tma_plan_load(Xb + (size_t)tlast * PDP_N, nn * PDP_N, Xend);
      const pdp_tma_plan pu = pdp_tma_plan_load(Ub + (size_t)tlast * PDP_M, nn * PDP_M, Uend);
      pdp_mbar_expect(MB, px.bytes + pu.bytes);
      pdp_tma_issue_load(XS(0), px, MB);
      pdp_tma_issue_load(US(0), pu, MB);
      offx0 = px.off; offu0 = pu.off;
    }
    #pragma unroll 1
    for (int tc = tlast; tc >= 0; tc -= PDP_TC, buf ^= 1) {
      const int nst = H - tc < PDP_TC ? H - tc : PDP_TC;
      if (tc > 0) {
        const pdp_tma_plan px = pdp_tma_plan_load(Xb + (size_t)(tc - PDP_TC) * PDP_N, PDP_TC * PDP_N, Xend);
        const pdp_tma_plan pu = pdp_tma_plan_load(Ub + (size_t)(tc - PDP_TC) * PDP_M, PDP_TC * PDP_M, Uend);
        pdp_mbar_expect(MB + (buf ^ 1), px.bytes + pu.bytes);
        pdp_tma_issue_load(XS(buf ^ 1), px, MB + (buf ^ 1));
        pdp_tma_issue_load(US(buf ^ 1), pu, MB + (buf ^ 1));
        if (buf) { offx0 = px.off; offu0 = pu.off; } else { offx1 = px.off; offu1 = pu.off; }
      }
      if (buf == 0) { pdp_mbar_wait(MB, phase0); phase0 ^= 1; } else { pdp_mbar_wait(MB + 1, phase1); phase1 ^= 1; }
      pdp_bulk_wait_read();
      const double* xs = XS(buf) + (buf ? offx1 : offx0);
      const double* us = US(buf) + (buf ? offu1 : offu0);
      double* Lg = Lb + (size_t)tc * PDP_N;
      double* Gg = dHu ? dHu + ((size_t)b * H + tc) * PDP_M : nullptr;
      const int offl = (int)((reinterpret_cast<uintptr_t>(Lg) >> 3) & 1);
      const int offg = (int)((reinterpret_cast<uintptr_t>(Gg) >> 3) & 1);
      double* ls = LS + offl;
      double* gs = GS + offg;
      #pragma unroll 1
      for (int s = nst - 1; s >= 0; --s) {
        #pragma unroll
        for (int i = 0; i < PDP_N; ++i) { ls[s * PDP_N + i] = lam[i]; x[i] = xs[s * PDP_N + i]; }
        #pragma unroll
        for (int i = 0; i < PDP_M; ++i) u[i] = us[s * PDP_M + i];
        if (dHu != nullptr) {
          pdp_f_dHu(x, u, lam, th, gu);
          #pragma unroll
          for (int i = 0; i < PDP_M; ++i) gs[s * PDP_M + i] = gu[i];
        }
        if (tc + s > 0) {
          pdp_f_dHx(x, u, lam, th, ln);
          #pragma unroll
          for (int i = 0; i < PDP_N; ++i) lam[i] = ln[i];
        }
      }
      pdp_bulk_store_fence();
      pdp_tma_send(LS, offl, Lg, nst * PDP_N);
      if (dHu != nullptr) pdp_tma_send(GS, offg, Gg, nst * PDP_M);
      pdp_bulk_commit();
    }
  }
#ifdef __CUDACC__
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // shared memory must outlive the stores that read it
#endif
  if (status && bad) atomicOr(&status[b], 1);
}
'''

K_LAUNCH_ROLLOUT_TMA_BRANCH = r'''  // Open-loop rollouts of SMALL batches go to the TMA kernel: per-lane bulk copies are bound by the TMA engine's operation rate
  // (~1 op / 40 cycles / SM measured), so they only pay with long chunks, and a long chunk's slots allow two warps per SM --
  // enough for batches of up to 2 x 32 x #SM trajectories (C4: 8 192 per GPU, 0.187 -> 0.151 ms); larger batches (C3: 16 384)
  // stay on the register-prefetch kernel (0.100 vs 0.139 ms).
  int pdp_dev = 0, pdp_sms = 0;
  cudaGetDevice(&pdp_dev);
  cudaDeviceGetAttribute(&pdp_sms, cudaDevAttrMultiProcessorCount, pdp_dev);
  if (fb_gains == nullptr && (B + PDP_TB - 1) / PDP_TB <= 2 * pdp_sms) {
    const size_t smem_t = (size_t)PDP_TB * PDP_TSTRIDE * sizeof(double);
    cudaError_t et = pdp_opt_in_smem((const void*)pdp_k_rollout_costate_tma, smem_t);
    if (et != cudaSuccess) return (int)et;
    pdp_k_rollout_costate_tma<<<(B + PDP_TB - 1) / PDP_TB, PDP_TB, smem_t, st>>>(B, H, x0, theta, theta_stride, U, X, Lam, cost, dHu, status);
    return (int)cudaGetLastError();
  }
'''


def rollout_tma_launcher(launch_common_text):
    old = "  pdp_k_rollout_costate<<<(B + 127) / 128, 128, 0, st>>>("
    assert launch_common_text.count(old) == 1
    return launch_common_text.replace(old, K_LAUNCH_ROLLOUT_TMA_BRANCH + old)


K_AUX_LQR = r'''
// =====================================================================================================
// Kernels 3a/3b: fused getAuxSys + LQR.lqrSolver (PDP.py:272-314 + 446-615), ONE WARP PER TRAJECTORY.
//   3a  pdp_k_aux_lqr_bwd: backward Riccati sweep in the stacked form (see DESIGN.md): lane j < NS owns
//       row j of the stack  Y = [P ; . ; W^T]  (rows 0..n-1: P, rows n+m..: columns of W).  Per step
//         Z      = P [F|G|E] (+ W on the E block)              (structural non-zeros only)
//         Q      = Hstack + Z^T [F|G]                           (n+m+r) x (n+m)
//         K|k    = -Quu^{-1} [Qux|Que]   (every lane solves for the column it owns; LDL^T in registers)
//         Y     <- Q(:,0:n) + Q(:,n:n+m) K
//       and the gains (K_t|k_t) are spilled to HBM.
//   3b  pdp_k_aux_lqr_fwd: forward pass  U_t = K_t X_t + k_t,  X_{t+1} = F_t X_t + G_t U_t + E_t
//       -> dX/dtheta, dU/dtheta (and/or the fused IRL loss / chain rule); PDP_FG trajectories per warp,
//       lane g*r+c owns column c of trajectory g.
//   The auxiliary matrices are evaluated in chunks of PDP_CH steps with lanes = time steps; they never
//   exist in HBM.  Two kernels (not one) so that each gets its own register allocation / occupancy.
// =====================================================================================================
static_assert(PDP_NS <= 32, "one stack row per lane: n + m + r must not exceed 32 (codegen raises before this)");
extern "C" __global__ void __launch_bounds__(PDP_WPB * 32, PDP_MINB)
pdp_k_aux_lqr_bwd(int B, int H, const double* __restrict__ X, const double* __restrict__ U, const double* __restrict__ Lam,
                  const double* __restrict__ theta, int theta_stride, double* __restrict__ gains,
                  const double* __restrict__ auxrec, const double* __restrict__ termrec, int* __restrict__ status)
{
  extern __shared__ __align__(16) double pdp_smem[];
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * PDP_WPB + (threadIdx.x >> 5);
  if (b >= B) return;
  double* auxc = pdp_smem + (size_t)(threadIdx.x >> 5) * PDP_WARP_DOUBLES;   // [CH][AUXLD]
  double* ZT = auxc + PDP_OFF_ZT;                                            // Z^T staging
  double* KS = auxc + PDP_OFF_KS;                                            // K (m x n)
  double* QUU = auxc + PDP_OFF_QUU;                                          // m x m
  double* TH = auxc + PDP_OFF_TH;                                            // theta
  double* TB = auxc;                                                         // terminal buffer aliases the chunk buffer
  const int lrow = lane < PDP_NS ? lane : 0;
  const int gslot = lane < PDP_N ? lane : ((lane >= PDP_NM && lane < PDP_NS) ? lane - PDP_M : -1);
  const double* Xb = X + (size_t)b * (H + 1) * PDP_N;
  const double* Ub = U + (size_t)b * H * PDP_M;
  const double* Lb = Lam + (size_t)b * H * PDP_N;
  (void)Xb; (void)Ub; (void)Lb;
  bool bad = false;
@@TABLOAD@@
  if (theta != nullptr) for (int i = lane; i < PDP_NTH; i += 32) TH[i] = theta[(size_t)b * theta_stride + i];
  __syncwarp();
#if PDP_NRCP > 0
  if (lane == 0) pdp_f_recips(TH, TH + PDP_NTH);
  __syncwarp();
#endif
  // ---- terminal condition P = hxx(x_H), W = hxe(x_H)  (PDP.py:561-562)
@@EVAL_TERM@@
  __syncwarp();
  @@YDECL@@
  {
@@TERM_INIT@@
  }
  __syncwarp();
  #pragma unroll 1
  for (int tc = ((H - 1) / PDP_CH) * PDP_CH; tc >= 0; tc -= PDP_CH) {
    __syncwarp();            // every lane is done reading the previous chunk's slots
@@EVAL_AUX_CHUNK@@
@@PREFETCH_AUX_CHUNK@@
    __syncwarp();
    const int thi = (tc + PDP_CH < H ? tc + PDP_CH : H) - 1;
    #pragma unroll 1
    for (int t = thi; t >= tc; --t) {
      const double* ar = auxc + (t - tc) * PDP_AUXLD;
@@BACKWARD_STEP@@
      __syncwarp();
    }
  }
  if (status && bad) { if (lane == 0) atomicOr(&status[b], 2); }
}

extern "C" __global__ void __launch_bounds__(PDP_WPBF * 32, PDP_MINBF)
pdp_k_aux_lqr_fwd(int B, int H, const double* __restrict__ X, const double* __restrict__ U,
                  const double* __restrict__ theta, int theta_stride, const double* __restrict__ X0a, int x0a_stride,
                  double* __restrict__ dX, double* __restrict__ dU, const double* __restrict__ gains,
                  const double* __restrict__ Xref, const double* __restrict__ Uref, double* __restrict__ loss_dp,
                  const double* __restrict__ auxrec, int* __restrict__ status)
{
  // One warp carries PDP_FG trajectories: lane g*r + c owns column c of trajectory b0+g, so one shared-memory
  // wavefront (K entry, Jacobian slot) feeds FG*r lanes instead of r.  Each trajectory has its own region
  // [CHF][FLD] slots | K | theta | [CHF][n+m] residuals, PDP_FTS doubles apart (= 2 mod 16: distinct bank groups).
  extern __shared__ __align__(16) double pdp_smem[];
  const int lane = threadIdx.x & 31;
  const int b0 = (blockIdx.x * PDP_WPBF + (threadIdx.x >> 5)) * PDP_FG;
  if (b0 >= B) return;
  double* wbase = pdp_smem + (size_t)(threadIdx.x >> 5) * PDP_FWARP_DOUBLES;
  // lane g*PDP_FGS + c owns column c; the group stride PDP_FGS is r rounded up to EVEN so that both lanes of every
  // even/odd lane pair read the same trajectory's operand: only then is a 64-bit shared-memory load with several
  // group addresses served in one wavefront (measured with tools/microbench/smem_wavefronts.cu)
  const int grp = lane / PDP_FGS;
  const bool owner = grp < PDP_FG && lane - grp * PDP_FGS < PDP_R;
  const int g = grp < PDP_FG ? grp : 0;
  const int col = owner ? lane - grp * PDP_FGS : -1;
  const bool live = owner && (b0 + g < B);
  const int bg = (b0 + g < B) ? b0 + g : B - 1;                 // idle / tail lanes shadow a valid trajectory
  double* reg = wbase + g * PDP_FTS;
  const double* KS = reg + PDP_FOFF_KS;
  // evaluation mapping: lane ge*CHF + se evaluates step tc+se of trajectory b0+ge
  const int ge_ = lane / PDP_CHF;
  const bool evl = ge_ < PDP_FG;
  const int ge = evl ? ge_ : 0;
  const int se = lane - ge_ * PDP_CHF;
  const int be = (b0 + ge < B) ? b0 + ge : B - 1;
  double* ereg = wbase + ge * PDP_FTS;
  (void)ereg; (void)be; (void)se;
#if PDP_NTH > 0
  if (theta != nullptr)
    for (int i = lane; i < PDP_FG * PDP_NTH; i += 32) {
      const int tg = i / PDP_NTH, ti = i - tg * PDP_NTH;
      const int tb = (b0 + tg < B) ? b0 + tg : B - 1;
      wbase[tg * PDP_FTS + PDP_FOFF_TH + ti] = theta[(size_t)tb * theta_stride + ti];
    }
#endif
#if PDP_NRCP > 0
  __syncwarp();
  if (lane < PDP_FG) pdp_f_recips(wbase + lane * PDP_FTS + PDP_FOFF_TH, wbase + lane * PDP_FTS + PDP_FOFF_TH + PDP_NTH);
#endif
  @@XDECL@@
  {
@@XINIT@@
  }
  double* dXb = dX ? dX + (size_t)bg * (H + 1) * PDP_N * PDP_R : nullptr;
  double* dUb = dU ? dU + (size_t)bg * H * PDP_M * PDP_R : nullptr;
  __syncwarp();
  if (dXb != nullptr && live) {
    double* o = dXb + col;
@@X0STORE@@
  }
  const bool fused = (loss_dp != nullptr) && (Xref != nullptr);
  double dpacc = 0.0, lossacc = 0.0;
  double lacc[PDP_FG];
  #pragma unroll
  for (int gg = 0; gg < PDP_FG; ++gg) lacc[gg] = 0.0;
  // software prefetch of the gain records (k column of this lane, K spread over the warp) one step ahead
  @@GNDECL@@
@@KQ_SETUP@@
  {
    const int t = -1;
    (void)t;
@@GNLOAD@@
  }
  #pragma unroll 1
  for (int tc = 0; tc < H; tc += PDP_CHF) {
    const int nst = (tc + PDP_CHF < H ? PDP_CHF : H - tc);
    (void)nst;
@@EVAL_DYN_COOP@@
    {
      const int te = tc + se;
      if (evl && te < H) {
        double* eo = ereg + se * PDP_FLD;
        const double* the = ereg + PDP_FOFF_TH;
        (void)the; (void)eo;
@@EVAL_DYN@@
        if (fused) {
          // loss / chain rule of the IRL scripts (reference Examples/IRL/quadrotor/uav_PDP.py:67-75), per evaluation lane
          const double* xe = X + ((size_t)be * (H + 1) + te) * PDP_N;
          const double* xr = Xref + ((size_t)be * (H + 1) + te) * PDP_N;
          const double* ue = U + ((size_t)be * H + te) * PDP_M;
          const double* ur = Uref ? Uref + ((size_t)be * H + te) * PDP_M : ue;
          #pragma unroll
          for (int i = 0; i < PDP_N; ++i) {
            const double d = xe[i] - xr[i];
            ereg[PDP_FOFF_DL + se * PDP_N + i] = d; lossacc = fma(d, d, lossacc);
          }
          #pragma unroll
          for (int i = 0; i < PDP_M; ++i) {
            const double d = Uref ? ue[i] - ur[i] : 0.0;
            ereg[PDP_FOFF_DU + se * PDP_M + i] = d; lossacc = fma(d, d, lossacc);
          }
        }
      }
    }
@@PREFETCH_DYN_CHUNK@@
    __syncwarp();
    const int tend = tc + PDP_CHF < H ? tc + PDP_CHF : H;
    #pragma unroll 1
    for (int t = tc; t < tend; ++t) {
      const double* ar = reg + (t - tc) * PDP_FLD;
#if PDP_PF && PDP_PFL
      if (t == tend - PDP_PFL) {      // the next chunk's rows were brought into L2 a chunk ago; now pull them into L1
@@PREFETCH_DYN_CHUNK_L1@@
      }
#endif
#if PDP_PF
      // gain record of step t + PDP_PFD of this lane's trajectory: lane c of the group touches 128-byte line c
      if (col >= 0 && col * 16 < PDP_GREC + 15 && t + PDP_PFD < H)
        pdp_prefetch(gains + ((size_t)bg * H + t + PDP_PFD) * PDP_GREC + col * 16);
#endif
@@GCUR@@
@@KS_STORE@@
      if (t + 1 < H) {
@@GNLOAD@@
      }
      __syncwarp();
@@FORWARD_STEP@@
      if (fused) {
        const double* dlx = reg + PDP_FOFF_DL + (t - tc) * PDP_N;
        const double* dlu = reg + PDP_FOFF_DU + (t - tc) * PDP_M;
@@DPACC@@
      }
      if (live) {
        // each lane stores its column straight from registers (8r-byte runs; L2 merges the partial sectors)
        if (dXb != nullptr) {
          double* o = dXb + (size_t)(t + 1) * (PDP_N * PDP_R) + col;
@@XSTORE@@
        }
        if (dUb != nullptr) {
          double* o = dUb + (size_t)t * (PDP_M * PDP_R) + col;
@@USTORE@@
        }
      }
      __syncwarp();
@@XCOPY@@
    }
  }
  {
    double chk = 0.0;
@@XCHK@@
    if (status && live && !isfinite(chk)) atomicOr(&status[bg], 1);
  }
  if (fused) {
    // terminal term of the chain rule and of the loss; the residual sums were accumulated lane-wise per trajectory
    #pragma unroll
    for (int gg = 0; gg < PDP_FG; ++gg) {
      if (evl && ge == gg) lacc[gg] += lossacc;
      #pragma unroll
      for (int o = 16; o > 0; o >>= 1) lacc[gg] += __shfl_xor_sync(0xffffffffu, lacc[gg], o);
    }
    if (owner) {
      const double* xh = X + ((size_t)bg * (H + 1) + H) * PDP_N;
      const double* xrh = Xref + ((size_t)bg * (H + 1) + H) * PDP_N;
      double lt = lacc[0];
      #pragma unroll
      for (int gg = 1; gg < PDP_FG; ++gg) if (g == gg) lt = lacc[gg];
@@DPTERM@@
      if (live) {
        if (col == 0) loss_dp[(size_t)bg * (PDP_R + 1)] = lt;
        loss_dp[(size_t)bg * (PDP_R + 1) + 1 + col] = dpacc;
      }
    }
  }
}
'''

# ---- two-trajectories-per-warp variant of the backward kernel (PDP_BP == 2) --------------------------------
K_AUX_LQR_BWD2 = r"""
// Backward Riccati sweep, TWO TRAJECTORIES PER WARP: half-warp `half` carries trajectory b0 + half, and team lane
// tl = lane & 15 owns TWO rows of the stack Y = [P ; . ; W^T]:
//     slot 0: row tl            (tl < n : row tl of P)
//     slot 1: row n + tl        (tl < m : control row,  m <= tl < m + r : column tl - m of W)
// so every broadcast operand (Jacobian slot, K entry) feeds two DFMAs per lane, and the two half-warps fetch their
// own trajectory's operand in the same shared-memory wavefront (the per-trajectory regions are PDP_HS doubles
// apart, PDP_HS = 2 mod 16: distinct bank pairs).  Same algebra, same gain records as the one-trajectory kernel.
extern "C" __global__ void __launch_bounds__(PDP_WPB * 32, PDP_MINB)
pdp_k_aux_lqr_bwd(int B, int H, const double* __restrict__ X, const double* __restrict__ U, const double* __restrict__ Lam,
                  const double* __restrict__ theta, int theta_stride, double* __restrict__ gains,
                  const double* __restrict__ auxrec, const double* __restrict__ termrec, int* __restrict__ status)
{
  extern __shared__ __align__(16) double pdp_smem[];
  const int lane = threadIdx.x & 31;
  const int half = lane >> 4, tl = lane & 15;
  const int b0 = (blockIdx.x * PDP_WPB + (threadIdx.x >> 5)) * 2;
  if (b0 >= B) return;
  const bool live = b0 + half < B;
  const int b = live ? b0 + half : B - 1;                                    // a tail half shadows a valid trajectory
  double* auxc = pdp_smem + (size_t)(threadIdx.x >> 5) * PDP_WARP_DOUBLES + half * PDP_HS;   // [CH][AUXLD]
  double* ZT = auxc + PDP_OFF_ZT;                                            // Z^T staging
  double* KS = auxc + PDP_OFF_KS;                                            // K (m x n)
  double* QUU = auxc + PDP_OFF_QUU;                                          // m x m
  double* TH = auxc + PDP_OFF_TH;                                            // theta
  double* TB = auxc;                                                         // terminal buffer aliases the chunk buffer
  const int r0 = tl < PDP_N ? tl : 0;                                        // stack row of slot 0
  const int r1 = PDP_N + (tl < PDP_M + PDP_R ? tl : 0);                      // stack row of slot 1
  const bool wrow = tl >= PDP_M && tl < PDP_M + PDP_R;                       // slot 1 holds a column of W
  const int gslot0 = tl < PDP_N ? tl : -1;
  const int gslot1 = wrow ? PDP_N + tl - PDP_M : -1;
  const double* Xb = X + (size_t)b * (H + 1) * PDP_N;
  const double* Ub = U + (size_t)b * H * PDP_M;
  const double* Lb = Lam + (size_t)b * H * PDP_N;
  // stack rows whose column of Z = P [F|G|E] is structurally zero (bit j of PDP_ZMASK) are never staged: their
  // pick-up reads the all-zero row PDP_NS of the staging tile instead
  const int zr0 = ((PDP_ZMASK >> r0) & 1u) ? PDP_NS : r0;
  const int zr1 = ((PDP_ZMASK >> r1) & 1u) ? PDP_NS : r1;
  (void)Xb; (void)Ub; (void)Lb; (void)zr0; (void)zr1;
  bool bad = false;
@@TABLOAD@@
  if (theta != nullptr) for (int i = tl; i < PDP_NTH; i += 16) TH[i] = theta[(size_t)b * theta_stride + i];
  if (tl < PDP_LDZ) ZT[PDP_NS * PDP_LDZ + tl] = 0.0;
  __syncwarp();
#if PDP_NRCP > 0
  if (tl == 0) pdp_f_recips(TH, TH + PDP_NTH);
  __syncwarp();
#endif
  // ---- terminal condition P = hxx(x_H), W = hxe(x_H)  (PDP.py:561-562)
@@EVAL_TERM@@
  __syncwarp();
  @@YDECL@@
  {
@@TERM_INIT@@
  }
  __syncwarp();
  #pragma unroll 1
  for (int tc = ((H - 1) / PDP_CH) * PDP_CH; tc >= 0; tc -= PDP_CH) {
    __syncwarp();            // every lane is done reading the previous chunk's slots
@@EVAL_AUX_CHUNK@@
@@PREFETCH_AUX_CHUNK@@
    __syncwarp();
    const int thi = (tc + PDP_CH < H ? tc + PDP_CH : H) - 1;
    #pragma unroll 1
    for (int t = thi; t >= tc; --t) {
      const double* ar = auxc + (t - tc) * PDP_AUXLD;
@@BACKWARD_STEP@@
      __syncwarp();
    }
  }
  if (status && bad && live && tl == 0) atomicOr(&status[b], 2);
}

"""

K_OPT_IN = r'''
// Opt in to more than 48 KB of dynamic shared memory, once per (kernel, device).  The flags are atomics: two host threads
// racing here both set the (idempotent) attribute; nothing else in a module is mutable global state.
#include <atomic>
static cudaError_t pdp_opt_in_smem(const void* kernel, size_t bytes) {
  struct Slot { std::atomic<const void*> fn{nullptr}; std::atomic<unsigned long long> devmask{0}; };
  static Slot slots[8];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  Slot* sl = nullptr;
  for (auto& c : slots) {
    const void* cur = c.fn.load(std::memory_order_acquire);
    if (cur == kernel) { sl = &c; break; }
    if (cur == nullptr) {
      const void* expect = nullptr;
      if (c.fn.compare_exchange_strong(expect, kernel) || expect == kernel) { sl = &c; break; }
    }
  }
  if (sl != nullptr && (sl->devmask.load(std::memory_order_acquire) & bit)) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess && sl != nullptr) sl->devmask.fetch_or(bit, std::memory_order_release);
  return e;
}

'''

K_LAUNCH_COMMON = r'''
// =====================================================================================================
// Host-side launchers (C ABI of the module; bound by csrc/pdp_b200.cpp through dlopen)
// =====================================================================================================
''' + K_OPT_IN + r'''
extern "C" void pdpmod_info(int* out) {
  out[0] = PDP_KIND;
  out[1] = PDP_N; out[2] = PDP_M; out[3] = PDP_R; out[4] = PDP_NVAR; out[5] = PDP_NVAR_S; out[11] = PDP_NTH;
  out[6] = PDP_GREC; out[7] = PDP_NDENSE; out[8] = PDP_CH; out[9] = PDP_WPB; out[10] = PDP_WARP_DOUBLES;
}

extern "C" int pdpmod_rollout_costate(int B, int H, const double* x0, const double* theta, int theta_stride, const double* U,
                                      double* X, double* Lam, double* cost, double* dHu, int* status,
                                      const double* fb_gains, const double* fb_X, const double* fb_alpha, double* Uout,
                                      int fb_group, cudaStream_t st) {
  if (B <= 0) return 0;
  pdp_k_rollout_costate<<<(B + 127) / 128, 128, 0, st>>>(B, H, x0, theta, theta_stride, U, X, Lam, cost, dHu, status,
                                                         fb_gains, fb_X, fb_alpha, Uout, fb_group);
  return (int)cudaGetLastError();
}

extern "C" int pdpmod_aux_eval(int B, int H, const double* X, const double* U, const double* Lam, const double* theta,
                               int theta_stride, double* out, double* term, cudaStream_t st) {
  if (B <= 0) return 0;
  const int total = B * (H + 1);
  pdp_k_aux_eval<<<(total + 127) / 128, 128, 0, st>>>(B, H, X, U, Lam, theta, theta_stride, out, term);
  return (int)cudaGetLastError();
}

'''

K_LAUNCH_LQR = r'''
extern "C" int pdpmod_aux_lqr(int B, int H, const double* X, const double* U, const double* Lam, const double* theta,
                              int theta_stride, const double* X0a, int x0a_stride, double* dX, double* dU, double* gains,
                              const double* Xref, const double* Uref, double* loss_dp,
                              const double* auxrec, const double* termrec, int phases, int* status, cudaStream_t st) {
  if (B <= 0) return 0;
  const size_t smem_b = (size_t)PDP_WPB * PDP_WARP_DOUBLES * sizeof(double);
  const size_t smem_f = (size_t)PDP_WPBF * PDP_FWARP_DOUBLES * sizeof(double);
  cudaError_t e = pdp_opt_in_smem((const void*)pdp_k_aux_lqr_bwd, smem_b);
  if (e == cudaSuccess) e = pdp_opt_in_smem((const void*)pdp_k_aux_lqr_fwd, smem_f);
  if (e != cudaSuccess) return (int)e;
  if (phases & 1)
    pdp_k_aux_lqr_bwd<<<(B + PDP_WPB * PDP_BP - 1) / (PDP_WPB * PDP_BP), PDP_WPB * 32, smem_b, st>>>(B, H, X, U, Lam, theta, theta_stride, gains,
                                                                               auxrec, termrec, status);
  if (phases & 2)
    pdp_k_aux_lqr_fwd<<<(B + PDP_WPBF * PDP_FG - 1) / (PDP_WPBF * PDP_FG), PDP_WPBF * 32, smem_f, st>>>(B, H, X, U, theta, theta_stride, X0a,
                                                                                  x0a_stride, dX, dU, gains, Xref, Uref,
                                                                                  loss_dp, auxrec, status);
  return (int)cudaGetLastError();
}
'''


_ib = K_AUX_LQR.index('extern "C" __global__ void __launch_bounds__(PDP_WPB * 32, PDP_MINB)')
_if = K_AUX_LQR.index('extern "C" __global__ void __launch_bounds__(PDP_WPBF * 32, PDP_MINBF)')
K_AUX_LQR_HEAD, K_AUX_LQR_BWD, K_AUX_LQR_FWD = K_AUX_LQR[:_ib], K_AUX_LQR[_ib:_if], K_AUX_LQR[_if:]


