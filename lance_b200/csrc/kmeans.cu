// kmeans.cu -- Lloyd iterations on the device, batched over B independent problems
// (B = 1 for the IVF coarse quantiser, B = M for the PQ sub-space codebooks).
//
// Replaces  KMeans::train_kmeans            lance-index/src/vector/kmeans.rs:610-719
//           KMeansAlgoFloat::to_kmeans      kmeans.rs:371-446   (centroid update)
//           compute_membership_and_loss     kmeans.rs:250-281   (radius / f64 loss per cluster)
//           compute_cluster_sizes           kmeans.rs:210-232
//           split_clusters                  kmeans.rs:174-207
//
// Design: the reference sums each cluster's rows SEQUENTIALLY IN ROW ORDER in f32 and its losses in
// f64, so the result depends on the order.  Instead of atomics (fast but order-free) we build, per
// iteration, a stable counting sort of the rows by cluster (member lists in ascending row order)
// and add each cluster's members in that order: one warp per (cluster, 8-dimension chunk) gathers the
// member rows 128 at a time and lanes 0..7 run the sequential f32 chains (update_body_warp), one warp
// per cluster the f64 loss chain (stats_body).  Where no addition can round -- see "order-independent
// sums" below -- the chain is replaced by a parallel reduction that returns the same bits.  Given the
// same initial centroids the trained model is therefore BIT-IDENTICAL to the reference loop (checked
// against the oracle), at the cost of a sort of n 4-byte keys per iteration.  The scalar bookkeeping
// of an iteration runs in epilogue_kernel, and iterations 2.. replay one captured CUDA graph.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <limits>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <thread>
#include <unordered_map>

#include "../../include/lance_b200.h"

#include "assign.cuh"
#include "comm.cuh"
#include "common.cuh"
#include "exact.cuh"
#include "kmeans.cuh"
#include "tc_assign.cuh"
#include "tc_pq.cuh"

namespace lb2 {

// ------------------------------------------------------------------------------------------------
// stable counting sort of rows by cluster id
// ------------------------------------------------------------------------------------------------
__global__ void hist_kernel(const uint32_t* __restrict__ ids, const uint8_t* __restrict__ valid,
                            uint64_t n, int K, int chunk_rows, uint32_t* __restrict__ chunk_hist,
                            const uint8_t* __restrict__ active) {
  const int b = blockIdx.y;
  if (active && !active[b]) return;
  const uint64_t r0 = (uint64_t)blockIdx.x * chunk_rows;
  const uint64_t r1 = min(n, r0 + (uint64_t)chunk_rows);
  uint32_t* h = chunk_hist + ((size_t)b * gridDim.x + blockIdx.x) * K;
  for (uint64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x)
    if (!valid || valid[(size_t)b * n + r]) atomicAdd(&h[ids[(size_t)b * n + r]], 1u);
}

// per (b, k): exclusive scan over chunks (in place), total -> counts
__global__ void scan_chunks_kernel(uint32_t* __restrict__ chunk_hist, int nchunks, int K, int B,
                                   uint32_t* __restrict__ counts) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= B * K) return;
  const int b = g / K, k = g % K;
  uint32_t run = 0;
  for (int c = 0; c < nchunks; ++c) {
    uint32_t* p = chunk_hist + ((size_t)b * nchunks + c) * K + k;
    const uint32_t t = *p;
    *p = run;
    run += t;
  }
  counts[g] = run;
}

// per b: offsets[b][0..K] = exclusive scan of counts[b][:]
__global__ void offsets_kernel(const uint32_t* __restrict__ counts, int K,
                               uint32_t* __restrict__ offsets) {
  __shared__ uint32_t part[1024];
  const int b = blockIdx.x, t = threadIdx.x;
  const int seg = (K + 1023) / 1024;
  const int s = t * seg, e = min(K, s + seg);
  uint32_t sum = 0;
  for (int k = s; k < e; ++k) sum += counts[(size_t)b * K + k];
  part[t] = sum;
  __syncthreads();
  if (t == 0) {
    uint32_t run = 0;
    for (int i = 0; i < 1024; ++i) {
      uint32_t v = part[i];
      part[i] = run;
      run += v;
    }
    offsets[(size_t)b * (K + 1) + K] = run;
  }
  __syncthreads();
  uint32_t run = part[t];
  for (int k = s; k < e; ++k) {
    offsets[(size_t)b * (K + 1) + k] = run;
    run += counts[(size_t)b * K + k];
  }
}

// lanes of `act` that hold the same key as this lane.  (__match_any_sync gives the same mask but
// the MATCH unit is slow -- ~100 cycles per warp-wide call and not pipelined across warps, measured
// with ncu on the single-CTA sort -- while a ballot per key bit is a handful of cycles.)
__device__ __forceinline__ unsigned same_key_mask(unsigned act, uint32_t key, int nbits) {
  unsigned grp = act;
  for (int bit = 0; bit < nbits; ++bit) {
    const bool one = (key >> bit) & 1u;
    const unsigned bal = __ballot_sync(act, one);
    grp &= one ? bal : ~bal;
  }
  return grp;
}

// one warp per (chunk, b): rows in ascending order, rank inside a batch of 32 by match_any
__global__ void scatter_kernel(const uint32_t* __restrict__ ids, const uint8_t* __restrict__ valid,
                               uint64_t n, int K, int chunk_rows, uint32_t* __restrict__ chunk_hist,
                               const uint32_t* __restrict__ offsets, uint32_t* __restrict__ members,
                               const uint8_t* __restrict__ active) {
  const int b = blockIdx.y;
  if (active && !active[b]) return;
  const int lane = threadIdx.x;
  const int nbits = 32 - __clz(max(K - 1, 1));
  const uint64_t r0 = (uint64_t)blockIdx.x * chunk_rows;
  const uint64_t r1 = min(n, r0 + (uint64_t)chunk_rows);
  uint32_t* h = chunk_hist + ((size_t)b * gridDim.x + blockIdx.x) * K;
  const uint32_t* off = offsets + (size_t)b * (K + 1);
  for (uint64_t base = r0; base < r1; base += 32) {
    const uint64_t r = base + lane;
    const bool ok = r < r1 && (!valid || valid[(size_t)b * n + r]);
    const unsigned act = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const uint32_t key = ids[(size_t)b * n + r];
      const unsigned grp = same_key_mask(act, key, nbits);
      const int rank = __popc(grp & ((1u << lane) - 1));
      const uint32_t start = h[key];
      members[(size_t)b * n + off[key] + start + rank] = (uint32_t)r;
      __syncwarp(act);
      if (rank == 0) h[key] = start + __popc(grp);
    }
    __syncwarp();
  }
}

// Small problems (K <= 1024): the whole stable counting sort of one problem in ONE launch by a
// thread-block CLUSTER of 8 CTAs x 32 warps: every warp owns a contiguous chunk of rows, per-warp
// histograms and running counters live in shared memory, and the cross-CTA prefix is read through
// distributed shared memory between two cluster barriers (no global-memory round trips, no MATCH).
constexpr int SORT_CLUSTER = 8;
// the sort of ONE problem by the calling cluster (8 CTAs x 1024 threads; sm = 34 * K words of shared memory):
// also the member-list phase of the fused small-problem kernel below
template <int NT>
__device__ __forceinline__ void cluster_sort_body(const uint32_t* __restrict__ idb, const uint8_t* __restrict__ vb,
                                                  uint64_t n, int K, uint32_t* __restrict__ counts_b,
                                                  uint32_t* __restrict__ offsets_b, uint32_t* __restrict__ mem,
                                                  uint32_t* sm, uint32_t* wsum) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned crank = cluster.block_rank();
  constexpr int NW = NT / 32;
  uint32_t* wh = sm;             // [NW][K] per-warp histogram, then running counters
  uint32_t* tot = sm + NW * K;   // [K]     this CTA's per-key total (read by the other CTAs)
  uint32_t* off = tot + K;       // [K]     first output slot of this CTA's rows, per key
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  for (int i = tid; i < NW * K; i += NT) wh[i] = 0;
  if (tid < 32) wsum[tid] = 0;
  __syncthreads();
  constexpr uint32_t NONE = 0xffffffffu;
  const uint64_t nwarps = (uint64_t)NW * SORT_CLUSTER;
  const uint64_t chunk = ((n + nwarps - 1) / nwarps + 31) / 32 * 32;  // rows per warp, multiple of 32
  const uint64_t r0 = min(n, ((uint64_t)crank * NW + w) * chunk), r1 = min(n, r0 + chunk);
  for (uint64_t base = r0; base < r1; base += 32 * 8) {  // 8 independent loads in flight per lane
    uint32_t key[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint64_t r = base + u * 32 + lane;
      key[u] = (r < r1 && (!vb || vb[r])) ? idb[r] : NONE;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (key[u] != NONE) atomicAdd(&wh[w * K + key[u]], 1u);
  }
  __syncthreads();
  if (tid < K) {  // exclusive scan over this CTA's warps
    uint32_t run = 0;
    for (int ww = 0; ww < NW; ++ww) {
      const uint32_t t = wh[ww * K + tid];
      wh[ww * K + tid] = run;
      run += t;
    }
    tot[tid] = run;
  }
  cluster.sync();
  uint32_t total = 0, before = 0;  // over all CTAs / over the preceding CTAs, for key `tid`
  if (tid < K) {
    for (unsigned c = 0; c < SORT_CLUSTER; ++c) {
      const uint32_t t = cluster.map_shared_rank(tot, c)[tid];
      total += t;
      if (c < crank) before += t;
    }
  }
  uint32_t incl = total;  // inclusive scan of the per-key totals over the block (K <= 1024)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  if (w == 0) {
    uint32_t v = wsum[lane], inc2 = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc2, o);
      if (lane >= o) inc2 += t;
    }
    wsum[lane] = inc2 - v;  // exclusive
  }
  __syncthreads();
  const uint32_t excl = wsum[w] + incl - total;
  if (tid < K) {
    off[tid] = excl + before;
    if (crank == 0) {
      counts_b[tid] = total;
      offsets_b[tid] = excl;
      if (tid == K - 1) offsets_b[K] = excl + total;
    }
  }
  __syncthreads();
  const int nbits = 32 - __clz(max(K - 1, 1));
  for (uint64_t base0 = r0; base0 < r1; base0 += 32 * 8) {
    uint32_t keys[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint64_t r = base0 + u * 32 + lane;
      keys[u] = (r < r1 && (!vb || vb[r])) ? idb[r] : NONE;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint64_t r = base0 + u * 32 + lane;
      const uint32_t key = keys[u];
      const bool ok = key != NONE;
      const unsigned act = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const unsigned grp = same_key_mask(act, key, nbits);
        const int rank = __popc(grp & ((1u << lane) - 1));
        const uint32_t start = wh[w * K + key];
        mem[off[key] + start + rank] = (uint32_t)r;
        __syncwarp(act);
        if (rank == 0) wh[w * K + key] = start + __popc(grp);
      }
      __syncwarp();
    }
  }
  cluster.sync();  // nobody leaves while a neighbour may still read its `tot`
}

__global__ void __cluster_dims__(SORT_CLUSTER, 1, 1) __launch_bounds__(1024)
cluster_sort_kernel(const uint32_t* __restrict__ ids, const uint8_t* __restrict__ valid, uint64_t n,
                    int K, uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets,
                    uint32_t* __restrict__ members, const uint8_t* __restrict__ active) {
  const int b = blockIdx.y;
  if (active && !active[b]) return;  // uniform over the cluster
  extern __shared__ uint32_t sm[];
  __shared__ uint32_t wsum[32];
  cluster_sort_body<1024>(ids + (size_t)b * n, valid ? valid + (size_t)b * n : nullptr, n, K, counts + (size_t)b * K,
                          offsets + (size_t)b * (K + 1), members + (size_t)b * n, sm, wsum);
}

// ------------------------------------------------------------------------------------------------
// ordered centroid update (kmeans.rs:388-418): one thread per (b, cluster, t)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void update_body(size_t g, const float* __restrict__ x, int ldx, int ds, int K, int B, uint64_t n,
                              const uint32_t* __restrict__ members,
                              const uint32_t* __restrict__ offsets, float* __restrict__ centroids,
                              const uint8_t* __restrict__ active, int scale) {
  if (g >= (size_t)B * K * ds) return;
  const int b = g / ((size_t)K * ds);
  if (active && !active[b]) return;
  const int k = (g / ds) % K, t = g % ds;
  const uint32_t* off = offsets + (size_t)b * (K + 1);
  const uint32_t s = off[k], e = off[k + 1];
  const uint32_t* mem = members + (size_t)b * n;
  const float* col = x + (size_t)b * ds + t;
  float acc = 0.0f;
  uint32_t j = s;
  for (; j + 16 <= e; j += 16) {
    float v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = col[(size_t)mem[j + q] * ldx];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc = f_add(acc, v[q]);
  }
  for (; j < e; ++j) acc = f_add(acc, col[(size_t)mem[j] * ldx]);
  const uint32_t cnt = e - s;
  if (scale && cnt > 0) acc = __fmul_rn(acc, __fdiv_rn(1.0f, (float)cnt));  // kmeans.rs:414-416
  centroids[g] = acc;
}

// ---- order-independent sums -------------------------------------------------------------------
// The reference adds a cluster's members one after the other (f32 centroid sums, f64 loss), and a
// floating-point sum depends on that order -- unless no addition rounds.  If every term is an integer
// multiple of 2^g and sum|term| < 2^(g+p) (p = 24 for f32, 53 for f64), every partial sum of ANY
// association is such a multiple below 2^(g+p), hence exactly representable: all orders give the same
// bits.  g = min over the non-zero terms of (unbiased exponent - 23 + trailing zeros of the 24-bit
// significand).  The kernels below test this per cluster (with one bit of margin, the bound itself
// being computed in floating point) and then reduce in parallel; otherwise they run the sequential
// chain.  The f64 loss uses the general test (distances nearly always pass); the f32 centroid sums use
// the special case g = 0 (integer-valued columns: SIFT, u8, quantised data), which costs three adds per
// term to check.  A per-problem hint stops retrying once a problem's data has failed.
__device__ __forceinline__ int pow2_granule(float v, bool& bad) {
  const uint32_t bits = __float_as_uint(v) & 0x7fffffffu;
  if (bits == 0) return 0x7fffffff;  // zero is a multiple of everything
  uint32_t ex = bits >> 23;
  if (ex == 255) { bad = true; return 0x7fffffff; }
  uint32_t mant = bits & 0x7fffffu;
  if (ex) mant |= 0x800000u; else ex = 1;
  return (int)ex - 150 + (__ffs((int)mant) - 1);
}
__device__ __forceinline__ double pow2_f64(int e) {  // 2^e, -1022 <= e <= 1023
  return __longlong_as_double((long long)(e + 1023) << 52);
}

// The centroid update for ds % 8 == 0, one WARP per (b, cluster, 8-dim chunk).
//   fast path (see above): lanes stride the members, private f32 sums, one warp reduction;
//   sequential path: the 32 lanes fetch 128 member rows at a time (index load + 32-byte gather, all
//   independent -> one memory round trip per 128 members instead of one per 16), park them in shared
//   memory, then lanes 0..7 add their dimension in member order -- update_body's sums exactly.
constexpr int UPD_TILE = 128;
__device__ __forceinline__ void update_body_warp(size_t w, float* tile, const float* __restrict__ x, int ldx, int ds,
                                                 int K, int B, uint64_t n, const uint32_t* __restrict__ members,
                                                 const uint32_t* __restrict__ offsets,
                                                 float* __restrict__ centroids,
                                                 const uint8_t* __restrict__ active, int scale,
                                                 uint8_t* __restrict__ exact_hint) {
  const int lane = threadIdx.x & 31;
  const int nch = ds >> 3;
  if (w >= (size_t)B * K * nch) return;
  const int b = (int)(w / ((size_t)K * nch));
  if (active && !active[b]) return;
  const int k = (int)((w / nch) % K), c = (int)(w % nch);
  const uint32_t* off = offsets + (size_t)b * (K + 1);
  const uint32_t s = off[k], e = off[k + 1];
  const uint32_t* mem = members + (size_t)b * n;
  const float* col = x + (size_t)b * ds + c * 8;
  float* out = centroids + ((size_t)b * K + k) * ds + c * 8;
  const float inv = (scale && e > s) ? __fdiv_rn(1.0f, (float)(e - s)) : 1.0f;  // kmeans.rs:414-416
  const bool do_scale = scale && e > s;

  if (e - s >= 64 && exact_hint[b]) {  // ---- fast path: only worth it for long chains
    // Centroid sums: the cheap special case g = 0 -- every term an INTEGER (SIFT / u8 / quantised
    // columns) and sum|term| < 2^23.  (v + 1.5*2^23) - 1.5*2^23 == v  <=>  v is an integer, for |v| < 2^22.
    float sum[8], asum[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) { sum[t] = 0.0f; asum[t] = 0.0f; }
    float dev = 0.0f;  // max |v - round(v)|: 0 iff all terms are integers
    for (uint32_t j0 = s; j0 < e; j0 += 128) {
      float4 va[4], vb[4];
      bool have[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t j = j0 + u * 32 + lane;
        have[u] = j < e;
        if (have[u]) {
          const float4* src = reinterpret_cast<const float4*>(col + (size_t)mem[j] * ldx);
          va[u] = src[0];
          vb[u] = src[1];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (have[u]) {
          const float v[8] = {va[u].x, va[u].y, va[u].z, va[u].w, vb[u].x, vb[u].y, vb[u].z, vb[u].w};
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            sum[t] = __fadd_rn(sum[t], v[t]);
            asum[t] = __fadd_rn(asum[t], fabsf(v[t]));
            const float rt = __fadd_rn(__fadd_rn(v[t], 12582912.0f), -12582912.0f);
            dev = fmaxf(dev, fabsf(__fadd_rn(rt, -v[t])));
          }
        }
      }
    }
    // sum over the lanes of each lane's largest |.| sum bounds every dimension's sum|term| from above
    float amax = 0.0f;
#pragma unroll
    for (int t = 0; t < 8; ++t) amax = fmaxf(amax, asum[t]);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
#pragma unroll
      for (int t = 0; t < 8; ++t) sum[t] = __fadd_rn(sum[t], __shfl_xor_sync(0xffffffffu, sum[t], o));
      amax = __fadd_rn(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      dev = fmaxf(dev, __shfl_xor_sync(0xffffffffu, dev, o));
    }
    const bool exact = dev == 0.0f && amax < 8388608.0f;  // NaN / Inf terms fail the comparison
    if (exact) {
      if (lane < 8) {
        float r = 0.0f;
#pragma unroll
        for (int t = 0; t < 8; ++t)
          if (lane == t) r = sum[t];
        out[lane] = do_scale ? __fmul_rn(r, inv) : r;
      }
      return;
    }
    if (lane == 0) exact_hint[b] = 0;  // this problem's data is not of the exact kind: stop trying
  }

  float acc = 0.0f;
  constexpr int UU = UPD_TILE / 32;
  float4 pa[UU], pb[UU];  // the next tile's rows, fetched while the current tile is being summed
  auto fetch = [&](uint32_t base) {
    uint32_t r[UU];
#pragma unroll
    for (int u = 0; u < UU; ++u) {
      const uint32_t j = base + u * 32 + lane;
      r[u] = j < e ? mem[j] : 0xffffffffu;
    }
#pragma unroll
    for (int u = 0; u < UU; ++u) {
      if (r[u] != 0xffffffffu) {
        const float4* src = reinterpret_cast<const float4*>(col + (size_t)r[u] * ldx);
        pa[u] = src[0];
        pb[u] = src[1];
      }
    }
  };
  if (s < e) fetch(s);
  for (uint32_t base = s; base < e; base += UPD_TILE) {
    const uint32_t cnt = min((uint32_t)UPD_TILE, e - base);
#pragma unroll
    for (int u = 0; u < UU; ++u) {
      float4* dst = reinterpret_cast<float4*>(tile + (u * 32 + lane) * 8);
      dst[0] = pa[u];
      dst[1] = pb[u];
    }
    __syncwarp();
    if (base + UPD_TILE < e) fetch(base + UPD_TILE);
    if (lane < 8) {
      uint32_t q = 0;
      for (; q + 16 <= cnt; q += 16) {  // 16 loads ahead of a 4-cycle add chain
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = tile[(q + i) * 8 + lane];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc = f_add(acc, v[i]);
      }
      for (; q < cnt; ++q) acc = f_add(acc, tile[q * 8 + lane]);
    }
    __syncwarp();
  }
  if (lane < 8) out[lane] = do_scale ? __fmul_rn(acc, inv) : acc;
}

// per (b, cluster): f64 loss in row order, radius (max), last member row (kmeans.rs:266-280)
__device__ __forceinline__ void stats_body(int w, const float* __restrict__ dists, uint64_t n, int K, int B,
                             const uint32_t* __restrict__ members,
                             const uint32_t* __restrict__ offsets, double* __restrict__ losses,
                             float* __restrict__ radius, uint32_t* __restrict__ last_row,
                             const uint8_t* __restrict__ active, uint8_t* __restrict__ loss_hint) {
  const int lane = threadIdx.x & 31;
  if (w >= B * K) return;
  const int b = w / K, k = w % K;
  if (active && !active[b]) return;
  const uint32_t* off = offsets + (size_t)b * (K + 1);
  const uint32_t s = off[k], e = off[k + 1];
  const uint32_t* mem = members + (size_t)b * n;
  const float* dv = dists + (size_t)b * n;
  double loss = 0.0;
  float rad = 0.0f;
  if (e - s >= 64 && loss_hint[b]) {  // ---- order-independent f64 sum (see "order-independent sums")
    double sd = 0.0, ad = 0.0;
    float rm = 0.0f;
    int g = 0x7fffffff;
    bool bad = false;
    for (uint32_t j0 = s; j0 < e; j0 += 256) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t j = j0 + u * 32 + lane;
        v[u] = j < e ? dv[mem[j]] : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        sd += (double)v[u];
        ad += (double)fabsf(v[u]);
        rm = fmaxf(rm, v[u]);
        g = min(g, pow2_granule(v[u], bad));
      }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      sd += __shfl_xor_sync(0xffffffffu, sd, o);
      ad += __shfl_xor_sync(0xffffffffu, ad, o);
      rm = fmaxf(rm, __shfl_xor_sync(0xffffffffu, rm, o));
      g = min(g, __shfl_xor_sync(0xffffffffu, g, o));
    }
    bad = __any_sync(0xffffffffu, bad);
    if (!bad && (g == 0x7fffffff || ad < pow2_f64(g + 52))) {  // exact iff sum|.| < 2^(g+53); 1 bit margin
      if (lane == 0) {
        losses[w] = sd;
        radius[w] = rm;
        last_row[w] = mem[e - 1];
      }
      return;
    }
    if (lane == 0) loss_hint[b] = 0;
  }
  // 256 members per round, the NEXT round's (index, distance) gathers in flight while this round's
  // values are folded in member order: the only serial work left is the f64 add chain itself
  constexpr int SU = 8;
  float cur[SU], nxt[SU];
#pragma unroll
  for (int u = 0; u < SU; ++u) {
    const uint32_t j = s + u * 32 + lane;
    cur[u] = j < e ? dv[mem[j]] : 0.0f;
  }
  for (uint32_t base = s; base < e; base += SU * 32) {
#pragma unroll
    for (int u = 0; u < SU; ++u) {
      const uint32_t j = base + (SU + u) * 32 + lane;
      nxt[u] = j < e ? dv[mem[j]] : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < SU; ++u) {
      const uint32_t b0 = base + u * 32;
      if (b0 < e) {  // uniform
        const int cnt = min(32u, e - b0);
        if (cnt == 32) {
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const float v = __shfl_sync(0xffffffffu, cur[u], q);
            loss += (double)v;
            rad = fmaxf(rad, v);  // f32::max ignores NaN like fmaxf; dists of members are never NaN
          }
        } else {
          for (int q = 0; q < cnt; ++q) {
            const float v = __shfl_sync(0xffffffffu, cur[u], q);
            loss += (double)v;
            rad = fmaxf(rad, v);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < SU; ++u) cur[u] = nxt[u];
  }
  if (lane == 0) {
    losses[w] = loss;
    radius[w] = rad;
    last_row[w] = e > s ? mem[e - 1] : 0xffffffffu;
  }
}

// one launch for both: the first `update_blocks` blocks (128 threads each) run the ordered centroid
// update, the remaining blocks the per-cluster f64 loss / radius / last-member statistics
__global__ void __launch_bounds__(128)
update_stats_kernel(unsigned update_blocks, const float* __restrict__ x, int ldx, int ds, int K, int B,
                    uint64_t n, const uint32_t* __restrict__ members, const uint32_t* __restrict__ offsets,
                    float* __restrict__ centroids, const float* __restrict__ dists,
                    double* __restrict__ losses, float* __restrict__ radius,
                    uint32_t* __restrict__ last_row, const uint8_t* __restrict__ active, int scale,
                    int warp_update, uint8_t* __restrict__ hints /* [2][B]: exact-sum hints */) {
  __shared__ __align__(16) float tiles[4][UPD_TILE * 8];
  if (blockIdx.x < update_blocks) {
    if (warp_update)
      update_body_warp((size_t)blockIdx.x * 4 + (threadIdx.x >> 5), tiles[threadIdx.x >> 5], x, ldx, ds, K, B, n,
                       members, offsets, centroids, active, scale, hints);
    else
      update_body((size_t)blockIdx.x * 128 + threadIdx.x, x, ldx, ds, K, B, n, members, offsets, centroids, active, scale);
  } else {
    stats_body((int)(((size_t)(blockIdx.x - update_blocks) * 128 + threadIdx.x) >> 5), dists, n, K, B,
               members, offsets, losses, radius, last_row, active, hints + B);
  }
}

// ---- multi-GPU: ONE exchange per Lloyd iteration ------------------------------------------
// Every rank packs its partial results of the iteration into one blob
//     [ sums f32 BK*ds | counts u32 BK | radius f32 BK | last member row u32 BK (global, +1; 0 = none) | loss f64 BK ]
// (update_stats_kernel writes the sums straight into it), the blobs are all-gathered in ONE collective, and
// every rank reduces the gathered blobs in RANK ORDER -- so all ranks hold bit-identical models whatever
// algorithm the transport uses, and the whole iteration (collective included) replays from a CUDA graph.
struct ExchangeLayout {
  size_t off_counts, off_radius, off_last, off_loss, bytes;
};
static ExchangeLayout exchange_layout(size_t BK, int ds) {
  ExchangeLayout L;
  L.off_counts = (BK * ds * sizeof(float) + 15) / 16 * 16;
  L.off_radius = L.off_counts + BK * 4;
  L.off_last = L.off_radius + BK * 4;
  L.off_loss = (L.off_last + BK * 4 + 7) / 8 * 8;
  L.bytes = (L.off_loss + BK * 8 + 15) / 16 * 16;
  return L;
}
__global__ void pack_partials_kernel(uint8_t* __restrict__ blob, ExchangeLayout L, size_t BK,
                                     const uint32_t* __restrict__ counts, const float* __restrict__ radius,
                                     const uint32_t* __restrict__ last_row, const double* __restrict__ losses,
                                     uint32_t row_offset) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= BK) return;
  reinterpret_cast<uint32_t*>(blob + L.off_counts)[g] = counts[g];
  reinterpret_cast<float*>(blob + L.off_radius)[g] = radius[g];
  reinterpret_cast<uint32_t*>(blob + L.off_last)[g] = last_row[g] == 0xffffffffu ? 0u : last_row[g] + row_offset + 1u;
  reinterpret_cast<double*>(blob + L.off_loss)[g] = losses[g];
}
__global__ void reduce_partials_kernel(const uint8_t* __restrict__ gathered, int nranks, ExchangeLayout L, size_t BK,
                                       int ds, int K, float* __restrict__ centroids, uint32_t* __restrict__ counts,
                                       float* __restrict__ radius, uint32_t* __restrict__ last_row,
                                       double* __restrict__ losses, const uint8_t* __restrict__ active) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= BK * ds) return;
  const size_t ck = g / ds;
  if (active && !active[ck / K]) return;
  float sum = 0.0f;
  uint32_t cnt = 0;
  for (int r = 0; r < nranks; ++r) {
    const uint8_t* blob = gathered + (size_t)r * L.bytes;
    sum = __fadd_rn(sum, reinterpret_cast<const float*>(blob)[g]);
    cnt += reinterpret_cast<const uint32_t*>(blob + L.off_counts)[ck];
  }
  centroids[g] = cnt > 0 ? __fmul_rn(sum, __fdiv_rn(1.0f, (float)cnt)) : sum;  // kmeans.rs:414-416
  if (g % ds == 0) {
    float rad = 0.0f;
    uint32_t last = 0;
    double loss = 0.0;
    for (int r = 0; r < nranks; ++r) {
      const uint8_t* blob = gathered + (size_t)r * L.bytes;
      rad = fmaxf(rad, reinterpret_cast<const float*>(blob + L.off_radius)[ck]);
      last = max(last, reinterpret_cast<const uint32_t*>(blob + L.off_last)[ck]);
      loss += reinterpret_cast<const double*>(blob + L.off_loss)[ck];
    }
    counts[ck] = cnt;
    radius[ck] = rad;
    last_row[ck] = last == 0u ? 0xffffffffu : last - 1u;
    losses[ck] = loss;
  }
}

// sharded initialisation: the k picked rows are GLOBAL row numbers (rank-major order); every rank copies
// the rows it owns into a zeroed buffer and the buffers are summed (each row has exactly one owner, x + 0
// is exact) -- the picks, and with them the model, do not depend on how the sample is sharded
__global__ void gather_init_owned_kernel(const float* __restrict__ x, int ldx, int ds, int K, int B,
                                         const uint32_t* __restrict__ rows, uint32_t row_offset, uint32_t n_local,
                                         float* __restrict__ out) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (size_t)B * K * ds) return;
  const int b = g / ((size_t)K * ds), k = (g / ds) % K, t = g % ds;
  const uint32_t gr = rows[(size_t)b * K + k];
  const bool mine = gr >= row_offset && gr - row_offset < n_local;
  out[g] = mine ? x[(size_t)(gr - row_offset) * ldx + (size_t)b * ds + t] : 0.0f;
}

__global__ void split_kernel(float* __restrict__ c, int i, int j, int ds) {
  const float eps = 1.0f / 1024.0f;
  for (int t = threadIdx.x; t < ds; t += blockDim.x) {
    const float cj = c[(size_t)j * ds + t];
    if ((t & 1) == 0) {
      c[(size_t)i * ds + t] = __fmul_rn(cj, 1.0f + eps);
      c[(size_t)j * ds + t] = __fmul_rn(cj, 1.0f - eps);
    } else {
      c[(size_t)i * ds + t] = __fmul_rn(cj, 1.0f - eps);
      c[(size_t)j * ds + t] = __fmul_rn(cj, 1.0f + eps);
    }
  }
}

__global__ void gather_init_kernel(const float* __restrict__ x, int ldx, int ds, int K, int B,
                                   const uint32_t* __restrict__ rows, float* __restrict__ out) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (size_t)B * K * ds) return;
  const int b = g / ((size_t)K * ds), k = (g / ds) % K, t = g % ds;
  out[g] = x[(size_t)rows[(size_t)b * K + k] * ldx + (size_t)b * ds + t];
}

// ------------------------------------------------------------------------------------------------
// member lists (also used to group rows by partition when an index is loaded)
// ------------------------------------------------------------------------------------------------
void MemberSort::run(const uint32_t* ids, const uint8_t* valid, uint64_t n, int K, int B,
                     const uint8_t* active) {
  if (K <= 1024 && n <= (1ull << 21) && n >= 1) {
    if (counts.n < (size_t)B * K) counts.alloc((size_t)B * K);
    if (offsets.n < (size_t)B * (K + 1)) offsets.alloc((size_t)B * (K + 1));
    if (members.n < (size_t)B * n) members.alloc((size_t)B * n);
    const size_t smem = sizeof(uint32_t) * (34 * (size_t)K);
    set_smem(cluster_sort_kernel, smem);
    LB2_LAUNCH("member_sort_cluster", cluster_sort_kernel, dim3(SORT_CLUSTER, B), 1024, smem, ids,
               valid, n, K, counts.p, offsets.p, members.p, active);
    return;
  }
  // chunk size: keep the per-chunk histogram table below ~256 MB
  // one warp scatters one chunk sequentially (32 rows per step): keep chunks short so that the
  // grid is wide, but bound the per-chunk histogram table (B * nchunks * K counters) to ~64 MB
  int chunk_rows = 256;
  while ((uint64_t)chunk_rows * 1024 < n) chunk_rows *= 2;  // at most ~1024 chunks (scan is per chunk)
  while ((double)B * (double)cdiv(n, chunk_rows) * K * 4.0 > 64e6) chunk_rows *= 2;
  const int nchunks = std::max(1u, cdiv(n, chunk_rows));
  if (chunk_hist.n < (size_t)B * nchunks * K) chunk_hist.alloc((size_t)B * nchunks * K);
  if (counts.n < (size_t)B * K) counts.alloc((size_t)B * K);
  if (offsets.n < (size_t)B * (K + 1)) offsets.alloc((size_t)B * (K + 1));
  if (members.n < (size_t)B * n) members.alloc((size_t)B * std::max<uint64_t>(n, 1));
  LB2_CUDA(cudaMemsetAsync(chunk_hist.p, 0, sizeof(uint32_t) * (size_t)B * nchunks * K, ctx().stream));
  dim3 grid(nchunks, B);
  LB2_LAUNCH("member_sort", hist_kernel, grid, 256, 0, ids, valid, n, K, chunk_rows, chunk_hist.p, active);
  LB2_LAUNCH("member_sort", scan_chunks_kernel, cdiv((uint64_t)B * K, 128), 128, 0, chunk_hist.p,
             nchunks, K, B, counts.p);
  LB2_LAUNCH("member_sort", offsets_kernel, B, 1024, 0, counts.p, K, offsets.p);
  LB2_LAUNCH("member_sort", scatter_kernel, grid, 32, 0, ids, valid, n, K, chunk_rows, chunk_hist.p,
             offsets.p, members.p, active);
}

// ------------------------------------------------------------------------------------------------
// per-iteration scalar epilogue ON THE DEVICE: exactly the reference's bookkeeping
//   compute_cluster_sizes (kmeans.rs:210-232), compute_balance_loss (:234-237), loss sum (:693),
//   split_clusters (:174-207, our rng), tolerance test (:704), next iteration's bias (:341-345).
// One block per problem; thread 0 runs the order-dependent scalar parts.
// ------------------------------------------------------------------------------------------------
struct LloydState {  // one per problem, device resident
  double loss;        // previous iteration's loss (f64::MAX at start)
  double last_loss;   // this iteration's loss
  float adjusted;     // adjusted_balance_factor (f32::MAX at start)
  float bf_cur;       // balance factor used by the membership step that just ran
  uint64_t rng;       // splitmix64 state
  uint32_t iters;
  uint32_t pad;
};

__device__ __forceinline__ uint64_t sm64_next(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// (a device function so that the fused small-problem kernel can run it too: problem b, by the NT threads of
// the calling block)
template <int NT>
__device__ __forceinline__ void epilogue_body(int b, int K, int ds, uint64_t n, float bf_param, double tolerance,
                const uint32_t* __restrict__ counts, const double* __restrict__ losses,
                const float* __restrict__ radius, const uint32_t* __restrict__ last_row,
                uint64_t* __restrict__ cluster_sizes, float* __restrict__ bias, int bias_ld,
                float* __restrict__ centroids, LloydState* __restrict__ states,
                uint8_t* __restrict__ active, TcPqPrepArgs pq_prep) {
  const int tid = threadIdx.x;
  __shared__ int s_i, s_j;
  LloydState& st = states[b];
  uint64_t* cs = cluster_sizes + (size_t)b * K;
  const uint32_t* cnt = counts + (size_t)b * K;
  const double* ls = losses + (size_t)b * K;
  float* cb = centroids + (size_t)b * K * ds;
  // (1) order-independent parts in parallel: cluster sizes, sum of squares, and the
  //     "first cluster to reach the final maximum" = max count, then smallest last member row
  __shared__ unsigned long long s_red_a[NT];  // packed (count << 32 | ~last_row) -> max
  __shared__ unsigned long long s_red_sq[NT];
  __shared__ int s_red_id[NT];
  __shared__ double s_chunk[1024];
  __shared__ int s_any_empty;
  {
    unsigned long long best = 0, sq = 0;
    int best_id = 0;
    bool have = false;
    int empty = 0;
    for (int k = tid; k < K; k += blockDim.x) {
      const uint32_t c = cnt[k];
      cs[k] = c;
      sq += (unsigned long long)c * c;
      empty |= (c == 0);
      const uint32_t lr = c > 0 ? last_row[(size_t)b * K + k] : 0xffffffffu;
      const unsigned long long key = ((unsigned long long)c << 32) | (uint32_t)(~lr);
      // ties on (count, last_row) cannot happen for c > 0 (a row belongs to one cluster); for
      // c == 0 everywhere the reference keeps id 0 -> prefer the lowest k on equal keys
      if (!have || key > best) { best = key; best_id = k; have = true; }
    }
    s_red_a[tid] = have ? best : 0ull;
    s_red_id[tid] = have ? best_id : 0x7fffffff;
    s_red_sq[tid] = sq;
    const int any = __syncthreads_or(empty);
    if (tid == 0) s_any_empty = any;
    for (int off = NT / 2; off >= 1; off >>= 1) {
      if (tid < off) {
        const unsigned long long o = s_red_a[tid + off];
        const int oi = s_red_id[tid + off];
        if (o > s_red_a[tid] || (o == s_red_a[tid] && oi < s_red_id[tid])) {
          s_red_a[tid] = o;
          s_red_id[tid] = oi;
        }
        s_red_sq[tid] += s_red_sq[tid + off];
      }
      __syncthreads();
    }
  }
  // (2) the f64 loss sum is order dependent (kmeans.rs:693): staged through shared memory in
  //     chunks, added sequentially by thread 0
  double sum = 0.0;
  for (int k0 = 0; k0 < K; k0 += 1024) {
    for (int k = tid; k < 1024 && k0 + k < K; k += blockDim.x) s_chunk[k] = ls[k0 + k];
    __syncthreads();
    if (tid == 0) {
      const int m = min(1024, K - k0);
      for (int k = 0; k < m; ++k) sum += s_chunk[k];
    }
    __syncthreads();
  }
  if (tid == 0) {
    const int max_id = s_red_id[0] == 0x7fffffff ? 0 : s_red_id[0];
    const uint64_t size_sq = s_red_sq[0];
    st.adjusted = __fdiv_rn(__fsub_rn(radius[(size_t)b * K + max_id],
                                      __fdiv_rn((float)ls[max_id], (float)cs[max_id])),
                            (float)n);
    const float balance_loss =
        __fmul_rn(st.bf_cur, __fsub_rn((float)size_sq, __fdiv_rn((float)(n * n), (float)K)));
    st.last_loss = sum + (double)balance_loss;
    st.iters += 1;
  }
  __syncthreads();
  // split_clusters: sequential over empty clusters, vector part by the whole block
  int next = 0;
  for (; s_any_empty;) {
    if (tid == 0) {
      int i = next;
      while (i < K && cs[i] != 0) ++i;
      s_i = i < K ? i : -1;
      if (i < K) {
        uint64_t j = 0;
        for (uint64_t tries = 0;; ++tries) {
          const float p = __fdiv_rn(__fsub_rn((float)cs[j], 1.0f), (float)(n - (uint64_t)K));
          const float u = (float)(sm64_next(st.rng) >> 40) * (1.0f / 16777216.0f);
          if (u < p) break;
          j = (j + 1) % (uint64_t)K;
          if (tries >= 64ull * K) {
            j = 0;
            for (int c = 1; c < K; ++c)
              if (cs[c] > cs[j]) j = c;
            break;
          }
        }
        cs[i] = cs[j] / 2;
        cs[j] -= cs[i];
        s_j = (int)j;
      }
    }
    __syncthreads();
    const int i = s_i, j = s_j;
    if (i < 0) break;
    const float eps = 1.0f / 1024.0f;
    for (int t = tid; t < ds; t += blockDim.x) {
      const float cj = cb[(size_t)j * ds + t];
      if ((t & 1) == 0) {
        cb[(size_t)i * ds + t] = __fmul_rn(cj, 1.0f + eps);
        cb[(size_t)j * ds + t] = __fmul_rn(cj, 1.0f - eps);
      } else {
        cb[(size_t)i * ds + t] = __fmul_rn(cj, 1.0f - eps);
        cb[(size_t)j * ds + t] = __fmul_rn(cj, 1.0f + eps);
      }
    }
    next = i + 1;
    __syncthreads();
  }
  // convergence (kmeans.rs:704) and the next iteration's balance factor / bias
  __shared__ float s_bf;
  if (tid == 0) {
    if (fabs(st.loss - st.last_loss) < tolerance * st.last_loss) {
      active[b] = 0;
    } else {
      st.loss = st.last_loss;
    }
    st.bf_cur = fminf(st.adjusted, bf_param);  // f32::min: the non-NaN operand
    s_bf = st.bf_cur;
  }
  __syncthreads();
  if (bias)
    for (int k = tid; k < K; k += blockDim.x)
      bias[(size_t)b * bias_ld + k] = __fmul_rn(s_bf, (float)cs[k]);
  if (NT == 256 && pq_prep.bm) {  // PQ tensor path (K == 256 codewords x 8 dims, 256 threads): operands of the NEXT iteration
    __shared__ float s_n2[256];
    __syncthreads();  // the split above may have rewritten codewords
    tc_pq_prep_block(cb, b, pq_prep, s_n2);
  }
}

__global__ void __launch_bounds__(256)
epilogue_kernel(int K, int ds, uint64_t n, float bf_param, double tolerance,
                const uint32_t* __restrict__ counts, const double* __restrict__ losses,
                const float* __restrict__ radius, const uint32_t* __restrict__ last_row,
                uint64_t* __restrict__ cluster_sizes, float* __restrict__ bias, int bias_ld,
                float* __restrict__ centroids, LloydState* __restrict__ states,
                uint8_t* __restrict__ active, TcPqPrepArgs pq_prep, volatile uint32_t* host_words) {
  const int b = blockIdx.x;
  if (pq_prep.bm && threadIdx.x == 0) pq_prep.fb_count[b] = 0;  // next iteration's undecided-row list (active or not)
  if (active[b])
    epilogue_body<256>(b, K, ds, n, bf_param, tolerance, counts, losses, radius, last_row, cluster_sizes, bias, bias_ld,
                       centroids, states, active, pq_prep);
  // progress word of problem b in PINNED HOST memory (one posted 4-byte write over PCIe, no copy-engine operation
  // in the stream): (number of epilogues run so far) << 1 | still active.  The host reads it to stop enqueuing
  // iterations, without ever draining the stream (lloyd_train).
  if (host_words && threadIdx.x == 0) {  // thread 0 also wrote active[b] above
    const uint32_t tick = ++states[b].pad;
    host_words[b] = (tick << 1) | (active[b] ? 1u : 0u);
  }
}


// ------------------------------------------------------------------------------------------------
// Small problems (hierarchical k-means splits a cluster with k' <= 16, kmeans.rs:885-895: thousands of
// Lloyd runs over a few hundred .. a few thousand rows): the WHOLE run in one launch.  A thread-block
// cluster of 8 CTAs x 512 threads iterates  assign (exact, 16 lanes per row = the reference's lane
// accumulators) -> stable member sort (cluster_sort_body) -> ordered centroid sums + f64 loss (the same
// update_body_warp / stats_body as the general path) -> scalar epilogue (epilogue_body, CTA 0)  with cluster
// barriers between the phases; nothing returns to the host until the run has converged.  Same device functions,
// same order of every floating-point operation as the multi-kernel path -> bit-identical models; ~10 us per
// iteration instead of ~10 launches.
// ------------------------------------------------------------------------------------------------
constexpr int SMALL_UPD_WARPS = 8;  // warps per CTA that run update / stats tasks (one 4 KB tile each)
constexpr int SMALL_NT = 512;       // 512 threads: 128 registers each (the update / stats bodies need them)
template <int METRIC>
__global__ void __cluster_dims__(SORT_CLUSTER, 1, 1) __launch_bounds__(SMALL_NT)
lloyd_small_kernel(const float* __restrict__ x, uint32_t n, int d, int K, float* __restrict__ centroids,
                   float bf_param, double tolerance, int max_iters, uint32_t* __restrict__ ids,
                   float* __restrict__ dists, uint8_t* __restrict__ valid, uint32_t* __restrict__ counts,
                   uint32_t* __restrict__ offsets, uint32_t* __restrict__ members, double* __restrict__ losses,
                   float* __restrict__ radius, uint32_t* __restrict__ last_row,
                   uint64_t* __restrict__ cluster_sizes, float* __restrict__ bias, LloydState* __restrict__ state,
                   uint8_t* __restrict__ active, uint8_t* __restrict__ hints, int warp_update) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned crank = cluster.block_rank();
  extern __shared__ __align__(16) uint8_t small_smem[];
  float* cs = reinterpret_cast<float*>(small_smem);                    // [K][d] centroids of this iteration
  float* sb = cs + (size_t)K * d;                                       // [16] bias
  float* tiles = sb + 16;                                               // [SMALL_UPD_WARPS][UPD_TILE * 8]
  uint32_t* sort_sm = reinterpret_cast<uint32_t*>(tiles + SMALL_UPD_WARPS * UPD_TILE * 8);  // [(NT/32 + 2) * K]
  __shared__ uint32_t wsum[32];
  const int tid = threadIdx.x, warp = tid >> 5, l = tid & 15;
  const unsigned hmask = 0xffffu << (16 * ((tid >> 4) & 1));
  const int n16 = d & ~15;
  const int nch = d >> 3;
  const uint32_t hw_global = crank * (SMALL_NT / 16) + (tid >> 4), hw_total = SORT_CLUSTER * (SMALL_NT / 16);  // half-warps
  for (int it = 1; it <= max_iters; ++it) {
    // ---- centroids + bias of this iteration into shared memory
    for (int i = tid; i < K * d; i += SMALL_NT) cs[i] = centroids[i];
    if (tid < 16) sb[tid] = tid < K ? bias[tid] : 0.0f;
    __syncthreads();
    // ---- membership (kmeans.rs:317-369): lane l of a half-warp owns lane accumulator l (l2.rs:82-88)
    for (uint32_t r = hw_global; r < n; r += hw_total) {
      const float* xv = x + (size_t)r * d;
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
      for (int e0 = l; e0 < n16; e0 += 16 * 8) {  // eight row elements in flight per lane (the loads are what costs)
        float xr[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) xr[u] = e0 + 16 * u < n16 ? xv[e0 + 16 * u] : 0.0f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = e0 + 16 * u;
          if (e < n16) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < K) acc[j] = f_add(acc[j], term<METRIC>(xr[u], cs[j * d + e]));
          }
        }
      }
      float best_key = __int_as_float(0x7f800000), best_val = best_key;
      uint32_t best_idx = 0xffffffffu;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j < K) {  // uniform
          float sq = 0.0f;  // sequential tail (l2.rs:69-79), every lane redundantly
          for (int e = n16; e < d; ++e) sq = f_add(sq, term<METRIC>(xv[e], cs[j * d + e]));
          float t = 0.0f;
#pragma unroll
          for (int q = 0; q < 16; ++q) t = f_add(t, __shfl_sync(hmask, acc[j], q, 16));
          const float v = finish<METRIC>(f_add(sq, t));
          const float key = f_add(v, sb[j]);
          if (key < best_key) { best_key = key; best_val = v; best_idx = j; }
        }
      }
      if (l == 0) {
        const bool ok = best_idx != 0xffffffffu;
        ids[r] = ok ? best_idx : 0u;
        dists[r] = ok ? best_val : __int_as_float(0x7fc00000);
        valid[r] = ok ? 1 : 0;
      }
    }
    __threadfence();
    cluster.sync();
    // ---- member lists (stable counting sort of the rows by cluster)
    cluster_sort_body<SMALL_NT>(ids, valid, n, K, counts, offsets, members, sort_sm, wsum);
    __threadfence();
    cluster.sync();
    // ---- ordered centroid sums (kmeans.rs:388-418) and per-cluster f64 loss / radius (kmeans.rs:266-280)
    if (warp_update) {
      if (warp < SMALL_UPD_WARPS) {
        const int ntask = K * nch + K;
        for (int t = crank * SMALL_UPD_WARPS + warp; t < ntask; t += SORT_CLUSTER * SMALL_UPD_WARPS) {
          if (t < K * nch)
            update_body_warp((size_t)t, tiles + warp * UPD_TILE * 8, x, d, d, K, 1, n, members, offsets, centroids,
                             nullptr, 1, hints);
          else
            stats_body(t - K * nch, dists, n, K, 1, members, offsets, losses, radius, last_row, nullptr, hints + 1);
          __syncwarp();
        }
      }
    } else {
      for (size_t g = (size_t)crank * SMALL_NT + tid; g < (size_t)K * d; g += (size_t)SORT_CLUSTER * SMALL_NT)
        update_body(g, x, d, d, K, 1, n, members, offsets, centroids, nullptr, 1);
      for (int t = crank * (SMALL_NT / 32) + warp; t < K; t += SORT_CLUSTER * (SMALL_NT / 32))
        stats_body(t, dists, n, K, 1, members, offsets, losses, radius, last_row, nullptr, hints + 1);
    }
    __threadfence();
    cluster.sync();
    // ---- the iteration's scalar bookkeeping (cluster sizes, balance loss, split_clusters, tolerance test, bias)
    if (crank == 0)
      epilogue_body<SMALL_NT>(0, K, d, n, bf_param, tolerance, counts, losses, radius, last_row, cluster_sizes, bias,
                          16, centroids, state, active, TcPqPrepArgs());
    __threadfence();
    cluster.sync();
    if (!*reinterpret_cast<volatile uint8_t*>(active)) break;  // converged (kmeans.rs:704)
  }
}

static bool lloyd_small_ok(uint64_t n, int B, int ds, int K, bool dist) {
  static const bool off = getenv("LB2_NO_SMALL_KMEANS") && *getenv("LB2_NO_SMALL_KMEANS");
  // one cluster = 8 SMs of plain FP32 against ~10 launches: measured (tools/small_kmeans_timing.py) to pay off only
  // for the tiniest runs (n * k * d <= 2^20, e.g. 512 rows x 2 centroids x 128: 43 vs 66 us per iteration)
  return !off && !dist && B == 1 && K >= 1 && K <= 16 && n >= 1 && n <= 16384 && (uint64_t)K * ds <= 24576 &&
         n * (uint64_t)K * ds <= (1ull << 20) && !ctx().profiling;
}

// ------------------------------------------------------------------------------------------------
// the Lloyd loop: no host round trip per iteration; the host only polls the `active` flags
// ------------------------------------------------------------------------------------------------
// per-problem progress words in pinned (device-mapped) host memory, one block per (thread, device), kept for the
// thread's lifetime: written by epilogue_kernel, read by lloyd_train's host loop
struct PollWords {
  static constexpr int LAG = 1, MAX_B = 256;
  volatile uint32_t* host = nullptr;
};
static PollWords* poll_words() {
  static thread_local std::map<int, PollWords> words;
  PollWords& r = words[ctx().device];
  if (!r.host) {
    void* p = nullptr;
    LB2_CUDA(cudaHostAlloc(&p, sizeof(uint32_t) * PollWords::MAX_B, cudaHostAllocMapped | cudaHostAllocPortable));
    r.host = static_cast<volatile uint32_t*>(p);
  }
  return &r;
}

void lloyd_train(const float* x, uint64_t n_in, int ldx, int B, int ds, int K, int metric,
                 float balance_factor_param, int max_iters, double tolerance, uint64_t seed,
                 const float* init_dev, float* centroids, std::vector<double>* loss_out,
                 std::vector<uint32_t>* iters_out) {
  LB2_REQUIRE(current_comm() || n_in >= (uint64_t)K, "KMeans: can not train %d centroids with %llu vectors", K,
              (unsigned long long)n_in);
  // kmeans.rs:623-627: only the first 512*k rows are used
  Comm* cm = current_comm();
  const bool dist = cm && cm->nranks > 1;
  uint64_t n = n_in >= (uint64_t)K * 512 ? (uint64_t)K * 512 : n_in;
  uint64_t n_global = n, row_offset = 0;
  if (dist) {
    // every rank holds a row shard of the sample; rows are ordered rank-major
    n = std::min<uint64_t>(n_in, ((uint64_t)K * 512 + cm->nranks - 1) / cm->nranks);
    std::vector<uint32_t> all(cm->nranks, 0);
    all[cm->rank] = (uint32_t)n;
    DevBuf<uint32_t> all_d(cm->nranks);
    h2d(all_d.p, all.data(), cm->nranks);
    comm_allreduce_u32(all_d.p, cm->nranks, RedOp::Sum);
    d2h(all.data(), all_d.p, cm->nranks);
    sync_stream();
    n_global = 0;
    for (int r = 0; r < cm->nranks; ++r) {
      if (r < cm->rank) row_offset += all[r];
      n_global += all[r];
    }
    LB2_REQUIRE(n_global >= (uint64_t)K, "KMeans: can not train %d centroids with %llu vectors", K,
                (unsigned long long)n_global);
  }
  LB2_REQUIRE(n_global < 0xfffffffeull, "training sample too large");
  const size_t BK = (size_t)B * K;
  const bool small = B > 1;
  if (small && !small_d_supported(ds))
    fail(LB2_UNSUPPORTED, "PQ sub-vector width %d is not supported by the device trainer yet", ds);

  // ---- init (kmeans.rs:149-170; our rng): k distinct rows by a partial Fisher-Yates ------------
  std::vector<LloydState> h_states(B);
  for (int b = 0; b < B; ++b) {
    h_states[b].loss = std::numeric_limits<double>::max();
    h_states[b].last_loss = 0.0;
    h_states[b].adjusted = std::numeric_limits<float>::max();
    h_states[b].bf_cur = std::fmin(std::numeric_limits<float>::max(), balance_factor_param);
    h_states[b].iters = 0;
    h_states[b].pad = 0;
  }
  if (init_dev) {
    if (init_dev != centroids) d2d(centroids, init_dev, BK * ds);
    for (int b = 0; b < B; ++b) h_states[b].rng = seed + b;
  } else {
    // partial Fisher-Yates over the virtual array idx[i] = i, kept sparse (only touched slots are
    // stored): identical picks to the dense version, O(K) instead of O(n) host work per problem.
    // Sharded: the picks range over the GLOBAL rows (rank-major), see gather_init_owned_kernel.
    std::vector<uint32_t> rows(BK);
    std::unordered_map<uint64_t, uint32_t> moved;
    const uint64_t n_pick = dist ? n_global : n;
    for (int b = 0; b < B; ++b) {
      SplitMix64 rng(seed + b);
      moved.clear();
      auto at = [&](uint64_t i) {
        auto it = moved.find(i);
        return it == moved.end() ? (uint32_t)i : it->second;
      };
      for (int i = 0; i < K; ++i) {
        const uint64_t j = i + rng.next() % (n_pick - i);
        const uint32_t vi = at(i), vj = at(j);
        moved[i] = vj;
        moved[j] = vi;
        rows[(size_t)b * K + i] = vj;
      }
      h_states[b].rng = rng.s;  // split_clusters continues the same stream
    }
    DevBuf<uint32_t> rows_d(BK);
    h2d(rows_d.p, rows.data(), BK);
    if (!dist) {
      LB2_LAUNCH("kmeans_init", gather_init_kernel, cdiv(BK * ds, 256), 256, 0, x, ldx, ds, K, B,
                 rows_d.p, centroids);
    } else {
      LB2_LAUNCH("kmeans_init", gather_init_owned_kernel, cdiv(BK * ds, 256), 256, 0, x, ldx, ds, K, B,
                 rows_d.p, (uint32_t)row_offset, (uint32_t)n, centroids);
      comm_allreduce_f32(centroids, BK * ds, RedOp::Sum);
    }
    sync_stream();  // rows (host vector) must outlive the copy
  }

  const int Kp = (K + 63) / 64 * 64;
  DevBuf<uint32_t> ids((size_t)B * n), last_row(BK);
  DevBuf<float> dists((size_t)B * n), radius(BK), bias;
  DevBuf<uint8_t> valid((size_t)B * n), active_d(B);
  DevBuf<double> losses(BK);
  DevBuf<uint64_t> cluster_sizes(BK);
  DevBuf<LloydState> states(B);
  cluster_sizes.zero();
  h2d(states.p, h_states.data(), B);
  std::vector<uint8_t> active(B, 1);
  h2d(active_d.p, active.data(), B);
  if (!small) {
    bias.alloc(Kp);
    bias.zero();  // iteration 1: cluster sizes are all zero -> bias 0
  }
  // multi-GPU: this rank's packed partial results and the gathered blobs of all ranks (see "ONE exchange")
  const ExchangeLayout xl = exchange_layout(BK, ds);
  DevBuf<uint8_t> blob, gathered;
  if (dist) {
    blob.alloc(xl.bytes);
    gathered.alloc(xl.bytes * (size_t)cm->nranks);
  }
  float* sums_p = dist ? reinterpret_cast<float*>(blob.p) : nullptr;
  DevBuf<uint8_t> hints((size_t)2 * B);  // order-independent-sum hints (update, loss) per problem
  LB2_CUDA(cudaMemsetAsync(hints.p, 1, (size_t)2 * B, ctx().stream));
  MemberSort ms;
  if (lloyd_small_ok(n, B, ds, K, dist) && (metric == METRIC_L2 || metric == METRIC_DOT) && ldx == ds) {
    // the whole run in ONE launch (lloyd_small_kernel)
    ms.counts.alloc(K);
    ms.offsets.alloc(K + 1);
    ms.members.alloc(n);
    DevBuf<float> bias16(16);
    bias16.zero();
    const int warp_update = (ds % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) ? 1 : 0;
    const size_t smem = sizeof(float) * ((size_t)K * ds + 16 + SMALL_UPD_WARPS * UPD_TILE * 8) +
                        sizeof(uint32_t) * (SMALL_NT / 32 + 2) * (size_t)K;
#define LB2_SMALL(MET)                                                                                              \
    do {                                                                                                            \
      set_smem(lloyd_small_kernel<MET>, smem);                                                                      \
      LB2_LAUNCH("kmeans_small_fused", lloyd_small_kernel<MET>, SORT_CLUSTER, SMALL_NT, smem, x, (uint32_t)n, ds, K, \
                 centroids, balance_factor_param, tolerance, max_iters, ids.p, dists.p, valid.p, ms.counts.p,       \
                 ms.offsets.p, ms.members.p, losses.p, radius.p, last_row.p, cluster_sizes.p, bias16.p, states.p,   \
                 active_d.p, hints.p, warp_update);                                                                 \
    } while (0)
    if (metric == METRIC_DOT) LB2_SMALL(METRIC_DOT); else LB2_SMALL(METRIC_L2);
#undef LB2_SMALL
    d2h(h_states.data(), states.p, B);
    sync_stream();
    if (loss_out) loss_out->assign(1, h_states[0].last_loss);
    if (iters_out) iters_out->assign(1, h_states[0].iters);
    return;
  }
  TcWorkspace tcws;
  TcPqWorkspace pqws;
  DevBuf<float> rn2;
  const bool pq_tc = small && ldx == B * ds && tc_pq_supported(n, ldx, B, ds, K, metric, x);
  TcPqPrepArgs pq_prep;      // bm == nullptr unless the PQ tensor path is in use
  bool pq_prepared = false;  // true once an epilogue has written the next iteration's operands
  if (pq_tc) {  // per-sub-space norms of the (fixed) training rows, once
    rn2.alloc(n * B);
    tc_pq_residual_norms(x, nullptr, nullptr, n, B, nullptr, rn2.p);
    pq_prep = tc_pq_prep_args(B, ldx, &pqws);
  }
  sync_stream();

  // progress words in pinned host memory (see "Convergence is polled" below)
  static const bool blocking_poll = getenv("LB2_BLOCKING_POLL") && *getenv("LB2_BLOCKING_POLL");
  // Single-rank runs only.  The host may read a word late and see the active bit of a LATER iteration than the one
  // it waited for: harmless alone (it just stops a no-op earlier), but two ranks of a sharded run could then enqueue
  // different numbers of iterations -- and the exchange inside the extra one would wait forever.  Sharded runs keep
  // the blocking poll, which reads the flags of exactly iteration `it` on every rank.
  PollWords* words = B <= PollWords::MAX_B && !blocking_poll && !ctx().profiling && !dist ? poll_words() : nullptr;
  volatile uint32_t* words_dev = nullptr;
  if (words) {
    for (int b = 0; b < B; ++b) words->host[b] = 0;  // (the previous run of this thread ended with a synchronise)
    void* dp = nullptr;
    LB2_CUDA(cudaHostGetDevicePointer(&dp, const_cast<uint32_t*>(words->host), 0));
    words_dev = static_cast<volatile uint32_t*>(dp);
  }
  // one Lloyd iteration = ~18 short kernels: membership, member sort, stats, update, scalar epilogue
  auto iteration = [&]() {
    if (!small) {
      assign_f32_ex(x, n, ds, centroids, K, metric, bias.p, /*bias_padded=*/true, ids.p, dists.p,
                    valid.p, nullptr, active_d.p, &tcws);
    } else if (pq_tc) {
      // the first call prepares the operands itself; afterwards the epilogue of iteration i has
      // already written them for iteration i + 1
      tc_pq_assign(x, rn2.p, n, ldx, B, centroids, nullptr, nullptr, ids.p, dists.p, valid.p,
                   active_d.p, &pqws, /*prepared=*/pq_prepared);
      pq_prepared = true;
    } else {
      small_d_assign_f32(x, n, ldx, B, ds, centroids, K, metric, nullptr, nullptr, nullptr, nullptr,
                         ids.p, dists.p, valid.p, active_d.p);
    }
    ms.run(ids.p, valid.p, n, K, B, active_d.p);
    // ds % 8 == 0 (and 16-byte aligned rows): warp-cooperative update, one warp per (b, cluster, 8 dims)
    const int warp_update = (ds % 8 == 0 && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) ? 1 : 0;
    const unsigned ub = warp_update ? cdiv(BK * (uint64_t)(ds / 8), 4) : cdiv(BK * ds, 128);
    const unsigned sb = cdiv((uint64_t)BK * 32, 128);
    static const bool split_us = getenv("LB2_SPLIT_UPDATE_STATS") && *getenv("LB2_SPLIT_UPDATE_STATS");
    if (split_us) {  // diagnostics: time the two halves of the fused launch separately
      LB2_LAUNCH("kmeans_update_only", update_stats_kernel, ub, 128, 0, ub, x, ldx, ds, K, B, n,
                 ms.members.p, ms.offsets.p, dist ? sums_p : centroids, dists.p, losses.p, radius.p,
                 last_row.p, active_d.p, dist ? 0 : 1, warp_update, hints.p);
      LB2_LAUNCH("kmeans_stats_only", update_stats_kernel, sb, 128, 0, 0u, x, ldx, ds, K, B, n,
                 ms.members.p, ms.offsets.p, dist ? sums_p : centroids, dists.p, losses.p, radius.p,
                 last_row.p, active_d.p, dist ? 0 : 1, warp_update, hints.p);
    } else
    LB2_LAUNCH("kmeans_update_stats", update_stats_kernel, ub + sb, 128, 0, ub, x, ldx, ds, K, B, n,
               ms.members.p, ms.offsets.p, dist ? sums_p : centroids, dists.p, losses.p, radius.p,
               last_row.p, active_d.p, dist ? 0 : 1, warp_update, hints.p);
    if (dist) {  // one exchange step per iteration over NVLink
      LB2_LAUNCH("kmeans_pack_partials", pack_partials_kernel, cdiv(BK, 256), 256, 0, blob.p, xl, BK, ms.counts.p,
                 radius.p, last_row.p, losses.p, (uint32_t)row_offset);
      comm_allgather_bytes(blob.p, gathered.p, xl.bytes);
      LB2_LAUNCH("kmeans_reduce_partials", reduce_partials_kernel, cdiv(BK * ds, 256), 256, 0, gathered.p, cm->nranks,
                 xl, BK, ds, K, centroids, ms.counts.p, radius.p, last_row.p, losses.p, active_d.p);
    }
    LB2_LAUNCH("kmeans_epilogue", epilogue_kernel, B, 256, 0, K, ds, n_global, balance_factor_param,
               tolerance, ms.counts.p, losses.p, radius.p, last_row.p, cluster_sizes.p,
               small ? nullptr : bias.p, Kp, centroids, states.p, active_d.p, pq_prep, words_dev);
  };
  // The first iteration runs eagerly (allocates every workspace, sets kernel attributes); the
  // iteration is then captured ONCE into a CUDA graph and replayed, so that the loop is not bound
  // by ~18 host launches per iteration.  (Event profiling and LB2_TC_STATS need eager launches.)
  const bool stats_env = getenv("LB2_TC_STATS") && *getenv("LB2_TC_STATS");
  // (NCCL collectives are capturable; the sharded iteration is replayed from the graph like the local one)
  // (eager launches for the small splits of hierarchical training were tried and lose: 577 vs 481 ms for a
  // K = 8192 tree -- the loop is bound by host launch throughput, which is what the graph relieves;
  // LB2_GRAPH_MIN_ROWS=<rows> restores eager launches below that size)
  const char* graph_min = getenv("LB2_GRAPH_MIN_ROWS");
  const uint64_t graph_rows = graph_min && *graph_min ? strtoull(graph_min, nullptr, 10) : 0;
  const bool use_graph = max_iters > 1 && !ctx().profiling && !stats_env && (dist || n * (uint64_t)B >= graph_rows) &&
                         !(getenv("LB2_NO_GRAPH") && *getenv("LB2_NO_GRAPH"));
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  uint64_t graph_nodes = 0;
  auto poll_done = [&]() {
    d2h(active.data(), active_d.p, B);
    sync_stream();
    for (int b = 0; b < B; ++b)
      if (active[b]) return false;
    return true;
  };
  // Convergence is polled WITHOUT draining the stream and without any operation in it: the epilogue kernel of every
  // iteration posts one progress word per problem -- (epilogues run) << 1 | active -- into pinned host memory,
  // and after enqueuing iteration `it` the host waits (plain memory reads) until every problem has reported
  // iteration it - LAG, then looks at the active bits.  The device therefore always has the next iteration queued.
  // (A blocking copy + synchronise every 4 iterations left the GPU idle for a copy and a graph launch each time;
  // an asynchronous copy + event per iteration cost as much in the stream: measured, tools/iter_timing.py.)
  // Iterations enqueued past convergence are no-ops: every kernel of an iteration returns at once for a problem
  // whose `active` flag is 0.  (Event profiling keeps the blocking poll: launch counts are then those of real
  // iterations; so do sharded runs, see `words` above.)
  auto words_done = [&](int it) {
    const uint32_t want = (uint32_t)it;
    uint64_t spins = 0;
    bool any_active = false;
    for (int b = 0; b < B; ++b) {
      uint32_t w;
      while (((w = words->host[b]) >> 1) < want) {
        if ((++spins & 0xFFFFF) == 0) {  // every ~1M reads: has the stream died or drained without reporting?
          const cudaError_t q = cudaStreamQuery(ctx().stream);
          if (q != cudaErrorNotReady) {
            if (q != cudaSuccess) LB2_CUDA(q);
            if ((words->host[b] >> 1) < want) fail(LB2_CUDA_ERROR, "k-means progress word %d never arrived", b);
          }
        }
      }
      any_active |= (w & 1u) != 0;
    }
    return !any_active;
  };
  iteration();
  bool done = max_iters == 1;
  if (!done && use_graph) {
    const uint64_t l0 = ctx().launches;
    LB2_CUDA(cudaStreamBeginCapture(ctx().stream, cudaStreamCaptureModeThreadLocal));
    try {
      iteration();
    } catch (...) {
      cudaStreamEndCapture(ctx().stream, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    LB2_CUDA(cudaStreamEndCapture(ctx().stream, &graph));
    graph_nodes = ctx().launches - l0;
    ctx().launches = l0;
    LB2_CUDA(cudaGraphInstantiate(&exec, graph, 0));
  }
  for (int it = 2; it <= max_iters && !done; ++it) {
    if (exec) {
      LB2_CUDA(cudaGraphLaunch(exec, ctx().stream));
      ctx().launches += graph_nodes;
    } else {
      iteration();
    }
    if (words) {
      if (it - PollWords::LAG >= 1) done = words_done(it - PollWords::LAG);
    } else if ((it & 3) == 0 || it == max_iters) {
      done = poll_done();  // blocking poll every 4 iterations
    }
  }
  if (exec) cudaGraphExecDestroy(exec);
  if (graph) cudaGraphDestroy(graph);
  d2h(h_states.data(), states.p, B);
  sync_stream();
  if (loss_out) {
    loss_out->resize(B);
    for (int b = 0; b < B; ++b) (*loss_out)[b] = h_states[b].last_loss;
  }
  if (iters_out) {
    iters_out->resize(B);
    for (int b = 0; b < B; ++b) (*iters_out)[b] = h_states[b].iters;
  }
}

// ------------------------------------------------------------------------------------------------
// hierarchical k-means for k > 256 (kmeans.rs:746-1003): top level with k0 = min(16, k, n), then the
// largest cluster is repeatedly split by a small k-means (k' <= 16) on its rows until k clusters
// exist.  The heap is Rust's BinaryHeap restated (push = sift_up, pop = swap + sift_down_to_bottom +
// sift_up; library/alloc/src/collections/binary_heap) ordered by (not finalized, size); the j-th
// training call uses seed + j (the reference is unseeded).  Row index lists stay on the device.
// ------------------------------------------------------------------------------------------------
namespace {
struct HCluster {
  uint32_t id;
  uint32_t off, loc;  // this rank's segment of the device index array (loc rows)
  uint64_t len;       // rows of the cluster over ALL ranks: what the reference's heap orders by
  bool finalized;
};
inline bool hc_le(const HCluster& a, const HCluster& b) {  // a <= b in the reference's Ord
  if (a.finalized != b.finalized) return a.finalized;  // finalized < not finalized
  return a.len <= b.len;
}
struct RustHeap {
  std::vector<HCluster> data;
  void sift_up(size_t start, size_t pos) {
    HCluster elt = data[pos];
    while (pos > start) {
      size_t parent = (pos - 1) / 2;
      if (hc_le(elt, data[parent])) break;
      data[pos] = data[parent];
      pos = parent;
    }
    data[pos] = elt;
  }
  void push(const HCluster& c) {
    data.push_back(c);
    sift_up(0, data.size() - 1);
  }
  HCluster pop() {
    HCluster item = data.back();
    data.pop_back();
    if (!data.empty()) {
      std::swap(item, data[0]);
      size_t end = data.size(), pos = 0;
      HCluster elt = data[0];
      size_t child = 1;
      while (child + 1 < end) {
        if (hc_le(data[child], data[child + 1])) child += 1;
        data[pos] = data[child];
        pos = child;
        child = 2 * pos + 1;
      }
      if (child + 1 == end) {
        data[pos] = data[child];
        pos = child;
      }
      data[pos] = elt;
      sift_up(0, pos);
    }
    return item;
  }
};
__global__ void gather_rows_u32_kernel(const float* __restrict__ x, const uint32_t* __restrict__ rows,
                                       uint64_t s, int d, float* __restrict__ out) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= s * d) return;
  out[g] = x[(uint64_t)rows[g / d] * d + g % d];
}
__global__ void compose_index_kernel(const uint32_t* __restrict__ parent, const uint32_t* __restrict__ members,
                                     uint32_t cnt, uint32_t* __restrict__ out) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < cnt) out[g] = parent[members[g]];
}
__global__ void iota_kernel(uint32_t* p, uint32_t n) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) p[g] = g;
}
}  // namespace

namespace {
// ---- one split = gather the cluster's rows, Lloyd, membership, stable sort, children's row lists -----------------
struct SplitOut {
  int ck = 0;
  std::vector<uint32_t> counts, offs;
  DevBuf<float> subc;       // [ck][d]
  DevBuf<uint32_t> order;   // the parent's row list re-ordered by (child, position), `kept` entries
  uint32_t kept = 0;
  std::map<std::string, ProfEntry> prof;  // kernels of a worker thread, merged into the caller's profile
  uint64_t launches = 0;
  std::exception_ptr err;
};
void split_cluster(const float* x, int d, const uint32_t* idx_seg, uint32_t loc, int ck, int metric, float balance_factor,
                   int max_iters, double tolerance, uint64_t seed, SplitOut& o) {
  o.ck = ck;
  o.subc.alloc((size_t)ck * d);
  const uint64_t n1 = std::max<uint32_t>(loc, 1);
  DevBuf<float> sub(n1 * d);
  DevBuf<uint32_t> ids(n1);
  DevBuf<uint8_t> valid(n1);
  if (loc)
    LB2_LAUNCH("gather_rows", gather_rows_u32_kernel, cdiv((uint64_t)loc * d, 256), 256, 0, x, idx_seg, (uint64_t)loc, d,
               sub.p);
  lloyd_train(sub.p, loc, d, 1, d, ck, metric, balance_factor, max_iters, tolerance, seed, nullptr, o.subc.p, nullptr,
              nullptr);
  assign_f32(sub.p, loc, d, o.subc.p, ck, metric, nullptr, ids.p, nullptr, valid.p, nullptr);
  MemberSort ms;
  ms.run(ids.p, valid.p, loc, ck, 1, nullptr);
  o.counts.resize(ck);
  o.offs.resize(ck + 1);
  d2h(o.counts.data(), ms.counts.p, ck);
  d2h(o.offs.data(), ms.offsets.p, ck + 1);
  sync_stream();
  // children's row lists = parent's list re-ordered by (child, position): stable; rows dropped as None by the
  // membership step leave the lists
  o.kept = o.offs[ck];
  if (o.kept) {
    o.order.alloc(o.kept);
    LB2_LAUNCH("compose_index", compose_index_kernel, cdiv(o.kept, 256), 256, 0, idx_seg, ms.members.p, o.kept, o.order.p);
  }
  sync_stream();
}

// ---- worker threads: independent splits train concurrently (each thread has its own stream) ---------------------
// A split is a chain of small dependent kernels (a few thousand rows, k <= 16): one stream leaves the GPU almost
// idle.  The sequential algorithm is kept EXACTLY -- splits are committed in the heap's pop order -- but the
// clusters that will reach the top of the heap soon are trained ahead of time on worker threads; their results
// wait in a cache keyed by cluster id (a split depends only on the cluster's rows, ck and seed + 1 + id).
class SplitWorkers {
 public:
  static SplitWorkers& get() {
    static SplitWorkers* w = new SplitWorkers();  // lives (with its detached threads) until the process exits
    return *w;
  }
  int threads() const { return nthreads_; }
  void submit(std::function<void()> f) {
    {
      std::lock_guard<std::mutex> lk(m_);
      q_.push_back(std::move(f));
      ++pending_;
    }
    cv_.notify_one();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [&] { return pending_ == 0; });
  }

 private:
  SplitWorkers() {
    const char* e = getenv("LB2_SPLIT_THREADS");
    nthreads_ = e && *e ? std::max(0, atoi(e) - 1) : 2;  // measured: 482 / 379 / 378 / 484 ms at 1 / 2 / 4 / 8 threads (launch-throughput bound)
    for (int i = 0; i < nthreads_; ++i) std::thread([this] { loop(); }).detach();
  }
  void loop() {
    for (;;) {
      std::function<void()> f;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return !q_.empty(); });
        f = std::move(q_.front());
        q_.pop_front();
      }
      f();
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  int nthreads_ = 0;
  std::mutex m_;
  std::condition_variable cv_, done_;
  std::deque<std::function<void()>> q_;
  int pending_ = 0;
};
}  // namespace

// Sharded: every rank holds a row shard of the sample and runs the SAME split loop -- the heap is
// ordered by the clusters' global sizes (one small all-reduce of <= 16 counters per split), each Lloyd run
// exchanges its partial sums once per iteration (lloyd_train), the row index lists stay local to the rank.
void hierarchical_train(const float* x, uint64_t n, int d, int K, int metric, float balance_factor,
                        int max_iters, double tolerance, int hk, uint64_t seed, float* centroids_out) {
  Comm* cm = current_comm();
  const bool dist = cm && cm->nranks > 1;
  LB2_REQUIRE(n < 0xffffffffull, "KMeans: too many vectors");
  const int kmax = std::max(std::min(hk, K), hk);
  DevBuf<uint32_t> cnt_d(kmax);
  // local per-cluster counts -> global counts (identity on one GPU)
  auto global_counts = [&](const std::vector<uint32_t>& local, int k, std::vector<uint64_t>& out) {
    out.assign(k, 0);
    if (!dist) {
      for (int i = 0; i < k; ++i) out[i] = local[i];
      return;
    }
    std::vector<uint32_t> tmp(local.begin(), local.begin() + k);
    h2d(cnt_d.p, tmp.data(), k);
    comm_allreduce_u32(cnt_d.p, k, RedOp::Sum);
    d2h(tmp.data(), cnt_d.p, k);
    sync_stream();
    for (int i = 0; i < k; ++i) out[i] = tmp[i];
  };
  uint64_t n_global = n;
  {
    std::vector<uint32_t> one(1, (uint32_t)n);
    std::vector<uint64_t> g;
    global_counts(one, 1, g);
    n_global = g[0];
  }
  LB2_REQUIRE(n_global >= (uint64_t)K, "KMeans: can not train %d centroids with %llu vectors", K,
              (unsigned long long)n_global);
  const int k0 = (int)std::min<uint64_t>(std::min(hk, K), n_global);
  const uint64_t n1 = std::max<uint64_t>(n, 1);
  DevBuf<float> top((size_t)k0 * d), store((size_t)2 * K * d + (size_t)k0 * d);
  DevBuf<uint32_t> ids(n1), idx(n1);
  DevBuf<uint8_t> valid(n1);
  lloyd_train(x, n, d, 1, d, k0, metric, balance_factor, max_iters, tolerance, seed, nullptr, top.p, nullptr, nullptr);
  assign_f32(x, n, d, top.p, k0, metric, nullptr, ids.p, nullptr, valid.p, nullptr);
  std::vector<uint32_t> counts(std::max(k0, hk)), offs(std::max(k0, hk) + 1);
  std::vector<uint64_t> gcounts;
  {
    MemberSort ms;
    ms.run(ids.p, valid.p, n, k0, 1, nullptr);
    d2h(counts.data(), ms.counts.p, k0);
    d2h(offs.data(), ms.offsets.p, k0 + 1);
    if (n) d2d(idx.p, ms.members.p, n);
    sync_stream();
  }
  ids.release();
  valid.release();
  global_counts(counts, k0, gcounts);
  RustHeap heap;
  uint32_t next_id = 0;
  const size_t store_slots = (size_t)2 * K + k0;
  for (int i = 0; i < k0; ++i) {
    if (gcounts[i] == 0) continue;
    d2d(store.p + (size_t)next_id * d, top.p + (size_t)i * d, d);
    heap.push(HCluster{next_id++, offs[i], counts[i], gcounts[i], false});
  }
  auto children_of = [&](uint64_t len, int remaining) {
    if (len <= (uint64_t)hk) return std::min(std::min(2, remaining), (int)len);
    return std::max(2, std::min(std::min((int)std::min<uint64_t>(len / hk, 1u << 30), remaining), hk));
  };
  // concurrency only without a communicator (NCCL calls of one communicator must not interleave across threads;
  // sharded builds whose sample fits one GPU gather it and come here without one, api.cu:train_ivf)
  SplitWorkers* workers = dist ? nullptr : &SplitWorkers::get();
  const int width = workers ? workers->threads() + 1 : 1;
  std::unordered_map<uint32_t, std::unique_ptr<SplitOut>> cache;
  int device = 0;
  cudaGetDevice(&device);
  Ctx& me = ctx();
  while ((int)heap.data.size() < K) {
    LB2_REQUIRE(!heap.data.empty(), "No cluster can be further split");
    HCluster big = heap.pop();
    if (big.finalized || big.len <= 1) {  // kmeans.rs:868-881: stop splitting
      heap.push(big);
      break;
    }
    const int remaining = K - (int)heap.data.size();
    const int ck = children_of(big.len, remaining);
    std::unique_ptr<SplitOut> out;
    auto hit = cache.find(big.id);
    if (hit != cache.end()) {
      if (hit->second->ck == ck) out = std::move(hit->second);  // else: the end game changed ck -> train again
      cache.erase(hit);
    }
    if (!out) {
      // this cluster + the clusters that will be popped soon (largest first; the exact pop order among equals
      // does not matter here -- a result is only used when its cluster is popped, and checked against ck then)
      std::vector<HCluster> wave{big};
      std::vector<int> wck{ck};
      if (width > 1 && remaining - ck > 2 * hk) {
        auto lt = [&](size_t a, size_t b) { return !hc_le(heap.data[b], heap.data[a]); };  // max-first
        std::priority_queue<size_t, std::vector<size_t>, decltype(lt)> front(lt);
        if (!heap.data.empty()) front.push(0);
        uint64_t rows = big.loc;
        int budget = remaining - ck;
        while (!front.empty() && (int)wave.size() < width) {
          const size_t i = front.top();
          front.pop();
          const HCluster& c = heap.data[i];
          if (c.finalized || c.len <= 1) break;  // the sequential loop stops there
          if (2 * i + 1 < heap.data.size()) front.push(2 * i + 1);
          if (2 * i + 2 < heap.data.size()) front.push(2 * i + 2);
          if (cache.count(c.id)) continue;
          const int cck = children_of(c.len, K);  // `remaining` not binding ...
          budget -= cck;
          if (budget <= 2 * hk) break;            // ... which is only certain away from the end game
          rows += c.loc;
          if (rows * (uint64_t)d * 4 > (4ull << 30)) break;
          wave.push_back(c);
          wck.push_back(cck);
        }
      }
      std::vector<std::unique_ptr<SplitOut>> outs(wave.size());
      for (auto& o : outs) o.reset(new SplitOut());
      sync_stream();  // the row lists written by earlier commits are visible to the workers' streams
      const bool prof_on = me.profiling;
      const std::string tag = me.tag;
      for (size_t w = 1; w < wave.size(); ++w) {
        SplitOut* o = outs[w].get();
        const HCluster c = wave[w];
        const int cck = wck[w];
        workers->submit([=, &idx]() {
          try {
            lb2_set_device(device);
            Ctx& wc = ctx();
            wc.profiling = prof_on;
            wc.tag = tag;
            wc.prof.clear();
            wc.launches = 0;
            split_cluster(x, d, idx.p + c.off, c.loc, cck, metric, balance_factor, max_iters, tolerance, seed + 1 + c.id, *o);
            wc.flush_profile();
            o->prof.swap(wc.prof);
            o->launches = wc.launches;
            wc.profiling = false;
          } catch (...) {
            o->err = std::current_exception();
          }
        });
      }
      try {
        split_cluster(x, d, idx.p + big.off, big.loc, ck, metric, balance_factor, max_iters, tolerance, seed + 1 + big.id,
                      *outs[0]);
      } catch (...) {
        outs[0]->err = std::current_exception();
      }
      if (wave.size() > 1) workers->wait();
      for (size_t w = 0; w < wave.size(); ++w) {
        if (outs[w]->err) std::rethrow_exception(outs[w]->err);
        me.launches += outs[w]->launches;
        for (auto& kv : outs[w]->prof) {
          me.prof[kv.first].launches += kv.second.launches;
          me.prof[kv.first].total_ms += kv.second.total_ms;
        }
        if (w) cache[wave[w].id] = std::move(outs[w]);
      }
      out = std::move(outs[0]);
    }
    global_counts(out->counts, ck, gcounts);
    int nonzero = 0;
    for (int i = 0; i < ck; ++i) nonzero += gcounts[i] > 0;
    if (nonzero <= 1) {  // ineffective split: finalise the original cluster (kmeans.rs:957-962)
      big.finalized = true;
      heap.push(big);
      continue;
    }
    if (out->kept) d2d(idx.p + big.off, out->order.p, out->kept);
    for (int i = 0; i < ck; ++i) {
      if (gcounts[i] == 0) continue;
      LB2_REQUIRE(next_id < store_slots, "hierarchical k-means: centroid store exhausted");
      d2d(store.p + (size_t)next_id * d, out->subc.p + (size_t)i * d, d);
      heap.push(HCluster{next_id++, big.off + out->offs[i], out->counts[i], gcounts[i], false});
    }
    sync_stream();  // `out` (and its device buffers) goes away here
  }
  if ((int)heap.data.size() != K)
    fail(LB2_INVALID_ARG, "hierarchical k-means produced %zu of %d clusters (no cluster can be further split)",
         heap.data.size(), K);
  std::vector<HCluster> all = heap.data;
  std::sort(all.begin(), all.end(), [](const HCluster& a, const HCluster& b) { return a.id < b.id; });
  for (int i = 0; i < K; ++i) d2d(centroids_out + (size_t)i * d, store.p + (size_t)all[i].id * d, d);
  sync_stream();
}

}  // namespace lb2
