// Micro-benchmarks that ground the kernel design (run on the GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_latency tools/microbench/fp64_latency.cu && /tmp/fp64_latency
// Prints cycles per dependent DFMA / DADD / DMUL / (1.0/x), DFMA issue rate with k independent chains in one warp,
// and the cost of a broadcast LDS.64 / LDS.128 with 1..16 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void k_dfma(double* out, long long* cyc, int iters) {
  double a[CHAINS];
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < CHAINS; ++i) a[i] = threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], b, c);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < CHAINS; ++i) s += a[i];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_div(double* out, long long* cyc, int iters) {
  double a = 1.0 + threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) a = 1.0 / (a + 0.5);
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_rcp_newton(double* out, long long* cyc, int iters) {
  double a = 1.0 + threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    double d = a + 0.5, r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    r = fma(r, fma(-d, r, 1.0), r);
    r = fma(r, fma(-d, r, 1.0), r);
    a = r;
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int VEC>
__global__ void k_lds_bcast(double* out, long long* cyc, int iters) {
  __shared__ __align__(16) double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * 1e-3;
  __syncthreads();
  double acc = 0;
  int idx = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (VEC == 2) {
        double2 v = *reinterpret_cast<const double2*>(&sm[(idx + 2 * k) & 1022]);
        acc += v.x + v.y;
      } else {
        acc += sm[(idx + k) & 1023];
      }
    }
    idx += 16 * VEC;
  }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int it = 4096;
#define RUN(K, G, B, W) K<<<G, B>>>(out, cyc, it); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); W
  RUN(k_dfma<1>, 1, 32, printf("dependent DFMA latency        : %.2f cycles\n", (double)h / it);)
  RUN(k_dfma<2>, 1, 32, printf("1 warp, 2 chains: cycles/DFMA  : %.2f\n", (double)h / it / 2);)
  RUN(k_dfma<4>, 1, 32, printf("1 warp, 4 chains: cycles/DFMA  : %.2f\n", (double)h / it / 4);)
  RUN(k_dfma<8>, 1, 32, printf("1 warp, 8 chains: cycles/DFMA  : %.2f\n", (double)h / it / 8);)
  RUN(k_dfma<16>, 1, 32, printf("1 warp, 16 chains: cycles/DFMA : %.2f\n", (double)h / it / 16);)
  RUN(k_dfma<8>, 1, 128, printf("4 warps (1/SMSP) x 8 chains: cycles/DFMA/warp : %.2f\n", (double)h / it / 8);)
  RUN(k_dfma<8>, 1, 256, printf("8 warps (2/SMSP) x 8 chains: cycles/DFMA/warp : %.2f\n", (double)h / it / 8);)
  RUN(k_dfma<8>, 1, 512, printf("16 warps (4/SMSP) x 8 chains: cycles/DFMA/warp: %.2f\n", (double)h / it / 8);)
  RUN(k_div, 1, 32, printf("dependent 1.0/(a+0.5)         : %.2f cycles\n", (double)h / it);)
  RUN(k_rcp_newton, 1, 32, printf("dependent rcp.approx+2 Newton : %.2f cycles\n", (double)h / it);)
  for (int warps = 1; warps <= 16; warps *= 2) {
    RUN(k_lds_bcast<1>, 1, 32 * warps, printf("LDS.64 broadcast, %2d warps/SM: %.2f cycles per load per warp\n", warps, (double)h / it / 16);)
    RUN(k_lds_bcast<2>, 1, 32 * warps, printf("LDS.128 broadcast, %2d warps/SM: %.2f cycles per load per warp\n", warps, (double)h / it / 16);)
  }
  return 0;
}
