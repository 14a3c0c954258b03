// FP64 FMA peak of the device (DFMA issue rate of the CUDA cores): the co-bound SURVEY 8(d) asks to report next to the HBM
// fraction.  Every thread runs 8 independent FMA chains; 148 x 4 blocks of 256 threads keep all four schedulers of every SM
// saturated.  C ABI: pdp_fp64_peak(iters, out) launches on the current device's NULL stream; the caller times it with
// CUDA events and converts (2 flop x 8 chains x 32 unrolled x iters x threads).  Test / bench infrastructure only.
#include <cuda_runtime.h>

extern "C" __global__ void __launch_bounds__(256) pdp_k_fp64_peak(int iters, double* out) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1e-9, a2 = a0 + 2e-9, a3 = a0 + 3e-9, a4 = a0 + 4e-9, a5 = a0 + 5e-9,
         a6 = a0 + 6e-9, a7 = a0 + 7e-9;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    #pragma unroll
    for (int k = 0; k < 32; ++k) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;     // keeps the chains alive, never true
}

extern "C" int pdp_fp64_peak(int iters, int blocks, double* out, void* stream) {
  pdp_k_fp64_peak<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, out);
  return (int)cudaGetLastError();
}
