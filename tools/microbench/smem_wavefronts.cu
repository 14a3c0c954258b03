// Shared-memory wavefront cost of the access patterns used by the Riccati kernels, read off ncu's own counter
// (the cycle-based loop in lds_groups.cu turned out to be bound by its dependent DADD chain, not by the data path):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --extended-lambda -o /tmp/smem_wavefronts tools/microbench/smem_wavefronts.cu
//   ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum \
//       --csv /tmp/smem_wavefronts          -> wavefronts per instruction = column 1 / (column 2 or 3), per pattern
// One warp, LOADS instructions per launch; kernel name = pattern.  half = lane >> 4, tl = lane & 15, HS = 1346 doubles
// (the per-trajectory region stride of the two-trajectory kernel, = 2 mod 16).
#include <cstdio>
#include <cuda_runtime.h>

#define HS 1346
#define LOADS 1024

__device__ __forceinline__ double lds64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ double lds128(unsigned a) { double x, y; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a)); return x + y; }
__device__ __forceinline__ void sts64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v)); }
__device__ __forceinline__ void sts128(unsigned a, double v) { asm volatile("st.shared.v2.f64 [%0], {%1, %1};" :: "r"(a), "d"(v)); }

template <int KIND, typename F>
__device__ void run(double* out, F addr_of) {
  extern __shared__ __align__(16) double sm[];
  for (int i = threadIdx.x; i < 2 * HS + 96; i += 32) sm[i] = i;
  __syncwarp();
  const int lane = threadIdx.x;
  const int idx = addr_of(lane);
  if (idx < 0) { out[lane] = 0; return; }
  const unsigned a = (unsigned)__cvta_generic_to_shared(sm + idx);
  double acc = 0;
  // every access of the unrolled body carries its own immediate offset (a multiple of 16 bytes, the same in all
  // lanes, so the pattern is preserved) -- identical volatile asm statements would be merged by the compiler
#pragma unroll 8
  for (int i = 0; i < LOADS; ++i) {
    const unsigned ai = a + ((i & 7) << 4);
    if (KIND == 0) acc += lds64(ai);
    if (KIND == 1) acc += lds128(ai);
    if (KIND == 2) sts64(ai, acc + i);
    if (KIND == 3) sts128(ai, acc + i);
  }
  out[lane] = acc;
}

#define PATTERN(NAME, KIND, EXPR) \
  __global__ void NAME(double* out) { run<KIND>(out, [](int lane) { const int half = lane >> 4, tl = lane & 15; (void)half; (void)tl; return (int)(EXPR); }); }

PATTERN(baseline_no_access,                4, 8)
PATTERN(lds64_warp_broadcast,              0, 8)
PATTERN(lds64_broadcast_per_half,          0, half * HS + 8)
PATTERN(lds64_broadcast_per_half_stride0,  0, half * 1344 + 8)
PATTERN(lds64_broadcast_per_quarter,       0, (lane >> 3) * 674 + 8)
PATTERN(lds64_one_plus_zero_per_half,      0, half * HS + (tl == 0 ? 8 : 42))
PATTERN(lds64_four_distinct_per_half,      0, half * HS + (tl < 4 ? 8 + tl : 20))
PATTERN(lds64_13rows_ld13_per_half,        0, half * HS + (tl < 13 ? tl : 0) * 13)
PATTERN(lds64_13rows_half0_only,           0, half == 0 ? (tl < 13 ? tl : 0) * 13 : -1)
PATTERN(lds64_13consecutive_per_half,      0, half * HS + (tl < 13 ? tl : 0))
PATTERN(lds64_16consecutive_same_in_both,  0, tl)
PATTERN(lds64_32consecutive,               0, lane)
PATTERN(lds64_8consecutive_per_half,       0, half * HS + (tl & 7))
PATTERN(lds64_two_words_per_half,          0, half * HS + (tl < 8 ? 8 : 42))
PATTERN(lds64_5lane_groups,                0, (lane / 5) * 226 + 8)
PATTERN(lds64_10lane_groups,               0, (lane / 10) * 450 + 8)
PATTERN(lds128_5lane_groups,               1, (lane / 5) * 226 + 8)
PATTERN(lds128_9lane_groups,               1, (lane / 9) * 450 + 8)
PATTERN(lds64_9lane_groups_idle_to_g0,     0, (lane < 27 ? lane / 9 : 0) * 450 + 8)
PATTERN(lds64_9lane_groups_idle_to_g2,     0, (lane < 27 ? lane / 9 : 2) * 450 + 8)
PATTERN(lds64_9lane_groups_idle_own_word,  0, (lane / 9) * 450 + 8)
PATTERN(lds64_10lane_groups_idle_to_g0,    0, (lane < 30 ? lane / 10 : 0) * 450 + 8)
PATTERN(lds64_11lane_groups,               0, (lane / 11) * 450 + 8)
PATTERN(lds64_8lane_groups_3used,          0, (lane < 24 ? lane / 8 : 0) * 450 + 8)
PATTERN(lds64_12lane_groups,               0, (lane / 12) * 450 + 8)
PATTERN(lds64_3words_interleaved,          0, (lane % 3) * 450 + 8)
PATTERN(lds64_5words_5lane_groups,         0, (lane < 25 ? lane / 5 : 4) * 226 + 8)
PATTERN(lds64_6words_5lane_groups,         0, (lane < 30 ? lane / 5 : 5) * 226 + 8)
PATTERN(lds128_warp_broadcast,             1, 8)
PATTERN(lds128_broadcast_per_half,         1, half * HS + 8)
PATTERN(lds128_broadcast_per_quarter,      1, (lane >> 3) * 674 + 8)
PATTERN(lds128_13rows_ld14_per_half,       1, half * HS + (tl < 13 ? tl : 0) * 14)
PATTERN(lds128_8consecutive_per_half,      1, half * HS + 2 * (tl & 7))
PATTERN(lds128_32consecutive,              1, 2 * lane)
PATTERN(sts64_13consecutive_per_half,      2, tl < 13 ? half * HS + tl : -1)
PATTERN(sts64_13consecutive_half0_only,    2, (tl < 13 && half == 0) ? tl : -1)
PATTERN(sts64_8consecutive_per_half,       2, tl < 8 ? half * HS + tl : -1)
PATTERN(sts128_13consecutive_per_half,     3, tl < 13 ? half * HS + 2 * tl : -1)
PATTERN(sts128_8rows_ld114_per_half,       3, tl < 8 ? half * HS + tl * 114 : -1)
PATTERN(sts128_16rows_ld114_per_half,      3, half * HS + (tl * 114) % 1300)

int main() {
  double* out;
  cudaMalloc(&out, 4096);
  const size_t smem = (2 * HS + 96) * sizeof(double);
#define LAUNCH(NAME) NAME<<<1, 32, smem>>>(out);
  LAUNCH(baseline_no_access) LAUNCH(lds64_warp_broadcast) LAUNCH(lds64_broadcast_per_half) LAUNCH(lds64_broadcast_per_half_stride0)
  LAUNCH(lds64_broadcast_per_quarter) LAUNCH(lds64_one_plus_zero_per_half) LAUNCH(lds64_four_distinct_per_half)
  LAUNCH(lds64_13rows_ld13_per_half) LAUNCH(lds64_13rows_half0_only) LAUNCH(lds64_13consecutive_per_half)
  LAUNCH(lds64_16consecutive_same_in_both) LAUNCH(lds64_32consecutive) LAUNCH(lds64_8consecutive_per_half)
  LAUNCH(lds64_two_words_per_half) LAUNCH(lds64_5lane_groups) LAUNCH(lds64_10lane_groups) LAUNCH(lds128_5lane_groups) LAUNCH(lds128_9lane_groups)
  LAUNCH(lds64_9lane_groups_idle_to_g0) LAUNCH(lds64_9lane_groups_idle_to_g2) LAUNCH(lds64_9lane_groups_idle_own_word) LAUNCH(lds64_10lane_groups_idle_to_g0)
  LAUNCH(lds64_11lane_groups) LAUNCH(lds64_8lane_groups_3used) LAUNCH(lds64_12lane_groups) LAUNCH(lds64_3words_interleaved) LAUNCH(lds64_5words_5lane_groups) LAUNCH(lds64_6words_5lane_groups)
  LAUNCH(lds128_warp_broadcast) LAUNCH(lds128_broadcast_per_half) LAUNCH(lds128_broadcast_per_quarter)
  LAUNCH(lds128_13rows_ld14_per_half) LAUNCH(lds128_8consecutive_per_half) LAUNCH(lds128_32consecutive)
  LAUNCH(sts64_13consecutive_per_half) LAUNCH(sts64_13consecutive_half0_only) LAUNCH(sts64_8consecutive_per_half)
  LAUNCH(sts128_13consecutive_per_half) LAUNCH(sts128_8rows_ld114_per_half) LAUNCH(sts128_16rows_ld114_per_half)
  cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
