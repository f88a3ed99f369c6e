// kmeans.cuh -- internal interface of kmeans.cu
#pragma once
#include <stdint.h>

#include <vector>

#include "common.cuh"
namespace lb2 {

// Stable counting sort of rows by cluster id, batched over B problems:
// members[b][offsets[b][k] .. offsets[b][k+1]) = rows of cluster k in ascending row order.
struct MemberSort {
  DevBuf<uint32_t> chunk_hist, counts, offsets, members;
  void run(const uint32_t* ids, const uint8_t* valid, uint64_t n, int K, int B,
           const uint8_t* active);
};

// B independent Lloyd problems over the columns [b*ds, (b+1)*ds) of x (row stride ldx).
// balance_factor is the post-division value (kmeans.rs:1344).  centroids: device [B][K][ds].
void lloyd_train(const float* x, uint64_t n, int ldx, int B, int ds, int K, int metric,
                 float balance_factor, int max_iters, double tolerance, uint64_t seed,
                 const float* init_dev, float* centroids, std::vector<double>* loss_out,
                 std::vector<uint32_t>* iters_out);
// k > 256: the reference's hierarchical scheme (kmeans.rs:746-1003); the loss is not meaningful (0)
void hierarchical_train(const float* x, uint64_t n, int d, int K, int metric, float balance_factor,
                        int max_iters, double tolerance, int hk, uint64_t seed, float* centroids_out);
}  // namespace lb2
