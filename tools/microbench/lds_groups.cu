// SUPERSEDED by smem_wavefronts.cu (wavefronts read off ncu): this clock-based loop is bound by its own dependent DADD
// chain, so it cannot tell one wavefront from two.  Kept for the record.
// Micro-benchmark behind the multi-trajectory-per-warp kernels (run on the GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/lds_groups tools/microbench/lds_groups.cu && /tmp/lds_groups
// What does a shared-memory load cost when the warp's lanes form G groups and each group reads ITS OWN address
// (one operand per trajectory), compared with a full-warp broadcast?  16 warps per SM, so the number printed is the
// sustained cost per load instruction on the SM's shared-memory data path (SM cycles per warp-load).
//   stride = distance between the groups' regions in doubles (2 mod 16 -> distinct bank pairs, 0 mod 16 -> conflicts)
#include <cstdio>
#include <cuda_runtime.h>

template <int VEC, int GROUPS>
__global__ void k_lds_groups(double* out, long long* cyc, int iters, int stride) {
  __shared__ __align__(16) double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-3;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int grp = lane / (32 / GROUPS);
  const double* base = sm + grp * stride;
  double acc = 0;
  int idx = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (VEC == 2) {
        double2 v = *reinterpret_cast<const double2*>(&base[(idx + 2 * k) & 1022]);
        acc += v.x + v.y;
      } else {
        acc += base[(idx + k) & 1023];
      }
    }
    idx += 16 * VEC;
  }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// every lane its own address: row `lane % rows` of a [rows][ld] tile per half-warp (the Z^T pick-up pattern)
__global__ void k_lds_rows(double* out, long long* cyc, int iters, int rows, int ld, int stride) {
  __shared__ __align__(16) double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-3;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int tl = lane & 15;
  const double* base = sm + (lane >> 4) * stride + (tl < rows ? tl : 0) * ld;
  double acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 13; ++k) acc += base[k + (it & 1)];
  }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (t1 - t0);
}

int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int it = 4096, warps = 16;
#define RUN(K, S, W) K<<<1, 32 * warps>>>(out, cyc, it, S); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); W
  const int strides[3] = {1346, 1344, 1348};
  for (int si = 0; si < 3; ++si) {
    const int s = strides[si];
    printf("-- group stride %d doubles (%d mod 16)\n", s, s % 16);
    RUN((k_lds_groups<1, 1>), s, printf("LDS.64  1 address  / warp : %.2f SM cycles per load instruction\n", (double)h / it / 16 / warps);)
    RUN((k_lds_groups<1, 2>), s, printf("LDS.64  2 addresses/ warp : %.2f\n", (double)h / it / 16 / warps);)
    RUN((k_lds_groups<1, 4>), s, printf("LDS.64  4 addresses/ warp : %.2f\n", (double)h / it / 16 / warps);)
    RUN((k_lds_groups<2, 1>), s, printf("LDS.128 1 address  / warp : %.2f\n", (double)h / it / 16 / warps);)
    RUN((k_lds_groups<2, 2>), s, printf("LDS.128 2 addresses/ warp : %.2f\n", (double)h / it / 16 / warps);)
    RUN((k_lds_groups<2, 4>), s, printf("LDS.128 4 addresses/ warp : %.2f\n", (double)h / it / 16 / warps);)
  }
  k_lds_rows<<<1, 32 * warps>>>(out, cyc, it, 13, 13, 1346); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("LDS.64 13 rows x 2 halves (ld 13, stride 1346): %.2f SM cycles per load instruction\n", (double)h / it / 13 / warps);
  k_lds_rows<<<1, 32 * warps>>>(out, cyc, it, 16, 13, 1346); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("LDS.64 16 rows x 2 halves (ld 13, stride 1346): %.2f SM cycles per load instruction\n", (double)h / it / 13 / warps);
  printf("(%d resident warps; numbers = elapsed cycles / loads issued by all warps)\n", warps);
  return 0;
}
