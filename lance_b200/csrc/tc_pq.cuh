// tc_pq.cuh -- internal interface of the tcgen05 PQ code-assignment path (tc_pq.cu)
#pragma once
#include <stdint.h>

#include "common.cuh"
namespace lb2 {
struct TcPqWorkspace {
  DevBuf<float> bm, cnh;  // codebook as a K-major 256 x d matrix; -|c|^2/2 per (m, c) + max|c|^2 per m
  DevBuf<uint32_t> fb_pairs, fb_count;
};
bool tc_pq_supported(uint64_t n, int d, int M, int ds, int Kc, int metric, const float* x);
// r_out (nullable) = x - cent[part] (cent nullable -> plain copy skipped); rn2[n][M] = |r_m|^2
void tc_pq_residual_norms(const float* x, const float* cent, const uint32_t* part, uint64_t n, int M,
                          float* r_out, float* rn2);
// codes != NULL: u8 [n][M] (encode); else ids/dists/valid [M][n] (training). Bit-identical to
// small_d_assign_f32 on the same inputs.
void tc_pq_assign(const float* r, const float* rn2, uint64_t n, int d, int M, const float* codebook,
                  const uint8_t* row_valid, uint8_t* codes, uint32_t* ids, float* dists,
                  uint8_t* valid, const uint8_t* active, TcPqWorkspace* ws);
// whole encode of n rows (residual fused when cent/part are given), chunked to bound temp memory
void pq_encode_dev(const float* x, uint64_t n, int d, int M, int ds, const float* codebook, int metric,
                   const float* cent, const uint32_t* part, const uint8_t* row_valid, uint8_t* codes);
}  // namespace lb2
