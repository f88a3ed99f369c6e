// exact.cuh -- device functions that reproduce the reference's f32 arithmetic ORDER bit for bit.
//
// The reference's L2 / dot are scalar Rust loops with 16 independent lane accumulators
// (lance-linalg/src/distance/l2.rs:57-91, dot.rs:30-58); rustc never contracts a*b+c, so every
// multiply and add is separately rounded.  We therefore use __fsub_rn/__fmul_rn/__fadd_rn, which
// nvcc never fuses into FFMA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lb2 {

constexpr int METRIC_L2 = 0;
constexpr int METRIC_COSINE = 1;
constexpr int METRIC_DOT = 2;

__device__ __forceinline__ float f_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sq_diff(float x, float y) {
  float d = __fsub_rn(x, y);
  return __fmul_rn(d, d);
}
template <int METRIC>
__device__ __forceinline__ float term(float x, float y) {
  return METRIC == METRIC_DOT ? __fmul_rn(x, y) : sq_diff(x, y);
}
// dot_distance = 1 - dot (dot.rs:68-70); L2 is the raw squared sum
template <int METRIC>
__device__ __forceinline__ float finish(float acc) {
  return METRIC == METRIC_DOT ? __fsub_rn(1.0f, acc) : acc;
}

// One thread computes the whole reference-order reduction (used for LUTs, re-ranking, flat scan).
template <int METRIC>
__device__ inline float dist_exact_thread(const float* __restrict__ x, const float* __restrict__ y,
                                          int d) {
  const int n16 = d & ~15;
  float s = 0.0f;
  for (int i = n16; i < d; ++i) s = f_add(s, term<METRIC>(x[i], y[i]));
  if (n16 == 0) return finish<METRIC>(f_add(s, 0.0f));
  float sums[16];
#pragma unroll
  for (int l = 0; l < 16; ++l) sums[l] = 0.0f;
  for (int c = 0; c < n16; c += 16) {
#pragma unroll
    for (int l = 0; l < 16; ++l) sums[l] = f_add(sums[l], term<METRIC>(x[c + l], y[c + l]));
  }
  float t = 0.0f;
#pragma unroll
  for (int l = 0; l < 16; ++l) t = f_add(t, sums[l]);
  return finish<METRIC>(f_add(s, t));
}

// f32::total_cmp as a signed-integer key (lance-index/src/vector/graph.rs:80-84)
__device__ __forceinline__ int32_t total_order_key(float f) {
  int32_t b = __float_as_int(f);
  return b ^ (int32_t)((uint32_t)(b >> 31) >> 1);
}

// (key, idx) lexicographic "less" used by every argmin reduction: lowest index among equal keys,
// which is what the reference's ascending scan with strict `<` yields (kernels.rs:79-89).
__device__ __forceinline__ bool better(float k_new, uint32_t i_new, float k_old, uint32_t i_old) {
  return k_new < k_old || (k_new == k_old && i_new < i_old);
}

}  // namespace lb2
