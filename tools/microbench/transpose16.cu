// Phase B of the backward Riccati kernel transposes a 16 x 16 FP64 tile inside each 16-lane team (lane tl holds row tl and
// needs column tl).  The kernel does it through shared memory; this microbenchmark times that against the alternative the
// round-1 review asked about: a butterfly of warp shuffles (4 stages x 8 exchanges x 2 SHFL.32 + selects), at the kernel's
// occupancy (one-warp blocks, 8 per SM, 2 warps per scheduler) with a DFMA per element between transpositions so that the
// values stay live in registers like in the kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/transpose16 tools/microbench/transpose16.cu && /tmp/transpose16
// Output: one JSON line with ns and SM cycles per transposition per warp for {none, smem, shfl}.
#include <cstdio>
#include <cuda_runtime.h>

#define LD 17            // odd row pitch: column reads of one team hit 16 distinct bank pairs
#define HS (16 * LD + 2) // team region stride = 2 (mod 16) doubles, as in the kernel (PDP_HS)

template <int MODE>
__global__ void __launch_bounds__(32, 8) k_transpose(double* out, int iters, double seed) {
  __shared__ __align__(16) double sm[2 * HS];
  const int lane = threadIdx.x, tl = lane & 15, half = lane >> 4;
  double* reg = sm + half * HS;
  double z[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) z[k] = seed * (lane * 16 + k);
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MODE == 1) {
#pragma unroll
      for (int k = 0; k < 16; ++k) reg[k * LD + tl] = z[k];        // element (tl, k) -> row k of the staging tile
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 16; ++k) z[k] = reg[tl * LD + k];        // lane tl picks up (k, tl) for every k
      __syncwarp();
    }
    if (MODE == 2) {
#pragma unroll
      for (int s = 8; s >= 1; s >>= 1) {
        const bool up = (tl & s) != 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if (k & s) continue;
          const double send = up ? z[k] : z[k | s];
          const double recv = __shfl_xor_sync(0xffffffffu, send, s);
          if (up) z[k] = recv; else z[k | s] = recv;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) z[k] = fma(z[k], 1.0000001, seed);   // the consumer
  }
  double acc = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) acc += z[k];
  out[(size_t)blockIdx.x * 32 + lane] = acc;
}

template <int MODE>
static float time_mode(double* out, int blocks, int iters) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k_transpose<MODE><<<blocks, 32>>>(out, iters / 10, 1e-9);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k_transpose<MODE><<<blocks, 32>>>(out, iters, 1e-9);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int blocks = sms * 8, iters = 20000;
  double* out = nullptr;
  cudaMalloc(&out, (size_t)blocks * 32 * sizeof(double));
  const float t0 = time_mode<0>(out, blocks, iters), t1 = time_mode<1>(out, blocks, iters), t2 = time_mode<2>(out, blocks, iters);
  if (cudaGetLastError() != cudaSuccess) { printf("{\"error\": \"cuda\"}\n"); return 1; }
  const double ns = 1e6 / iters;                       // ms per launch -> ns per iteration (8 warps per SM run concurrently)
  printf("{\"sms\": %d, \"warps_per_sm\": 8, \"iters\": %d, \"ns_per_iter\": {\"none\": %.2f, \"smem\": %.2f, \"shfl\": %.2f}, "
         "\"ns_per_transposition_per_sm_of_8_warps\": {\"smem\": %.2f, \"shfl\": %.2f}, \"nominal_clock_khz\": %d}\n",
         sms, iters, t0 * ns, t1 * ns, t2 * ns, (t1 - t0) * ns, (t2 - t0) * ns, khz);
  cudaFree(out);
  return 0;
}
