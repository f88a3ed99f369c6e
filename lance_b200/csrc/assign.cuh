// assign.cuh -- internal interface of the exact nearest-centroid kernels (assign.cu)
#pragma once
#include <stdint.h>
namespace lb2 {
// part/dist/valid are [n]; all_out (nullable) receives the full [n][K] distance matrix instead.
// bias (nullable, [K]) is added for the comparison only (kernels.rs:92-111).
void assign_f32(const float* x, uint64_t n, int d, const float* cent, int K, int metric,
                const float* bias, uint32_t* part, float* dist, uint8_t* valid, float* all_out);
// same, with a bias already padded to ceil(K/64)*64 floats and an optional device-side `active`
// flag (active[0] == 0 -> the kernel returns immediately; used by the Lloyd loop)
struct TcWorkspace;
void assign_f32_ex(const float* x, uint64_t n, int d, const float* cent, int K, int metric,
                   const float* bias, bool bias_padded, uint32_t* part, float* dist, uint8_t* valid,
                   float* all_out, const uint8_t* active, TcWorkspace* ws);
// exact tile kernel restricted to row_list[0 .. *row_count) (both on the device)
void assign_rows_f32(const float* x, uint64_t n_max, int d, const float* cent, int K, int metric,
                     const float* bias_padded, const uint32_t* row_list, const uint32_t* row_count,
                     uint32_t* part, float* dist, uint8_t* valid, const uint8_t* active,
                     TcWorkspace* ws, bool cT_ready = false);
// d < 16 path, batched over M sub-spaces; x row stride ldx, sub-space m reads columns [m*ds,(m+1)*ds).
// codes != NULL -> u8 [n][M] out (PQ encode), else ids/dists/valid [M][n] (PQ training).
bool small_d_supported(int ds);
void small_d_assign_f32(const float* x, uint64_t n, int ldx, int M, int ds, const float* codebook,
                        int Kc, int metric, const float* ivf_centroids, const uint32_t* part_ids,
                        const uint8_t* row_valid, uint8_t* codes, uint32_t* ids, float* dists,
                        uint8_t* valid, const uint8_t* active);
}  // namespace lb2
