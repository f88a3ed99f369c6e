// search.cu -- query-time kernels.
//
// Replaces  kmeans_find_partitions               lance-index/src/vector/kmeans.rs:1134-1158
//           IVFIndex::preprocess_query           rust/lance/src/index/vector/ivf/v2.rs:316-332
//           build_distance_table_l2/_dot         lance-index/src/vector/pq/distance.rs:24-92
//           compute_pq_distance (+ Dot fix-up)   pq/distance.rs:109-144, pq/storage.rs:921-962
//           FlatIndex::search heap top-k         lance-index/src/vector/flat/index.rs:82-177
//           SortExec(_distance,_rowid).fetch(k)  rust/lance/src/dataset/scanner.rs:3450-3466
//
// One CTA per (query, probed partition): the residual query and its M x 256 f32 lookup table are
// built in shared memory (never written to HBM), the partition's codes are streamed once with
// 128-bit loads, each row's distance is the reference's m-ascending f32 sum (bit-exact), and
// warp-level sorting networks (k <= 16) or a block radix select produce the k smallest (distance,
// position) pairs.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "assign.cuh"
#include "comm.cuh"
#include "common.cuh"
#include "exact.cuh"
#include "search.cuh"

namespace lb2 {

// prefilter (PreFilter::mask, lance-index/src/prefilter.rs:27-51; FlatIndex::search :129-165): one bit
// per STORAGE position (partition-sorted order); a cleared bit removes the row from the scan.  Filtered
// rows get the maximal key, so they can only surface when fewer than k allowed rows exist, and the
// output stage drops them by re-testing the bit.
__device__ __forceinline__ bool row_allowed(const uint64_t* __restrict__ allow, uint64_t pos) {
  return allow == nullptr || ((allow[pos >> 6] >> (pos & 63)) & 1ull) != 0;
}
// range query (flat/index.rs:100-115): a row enters the heap iff lower <= dist < upper in f32::total_cmp
// order; an absent bound is f32::MIN / f32::MAX (NOT -inf / +inf), exactly as the reference unwraps them
__device__ __forceinline__ bool key_in_range(const ScanFilter& f, int32_t key) {
  return !f.range || (key >= f.lo_key && key < f.hi_key);
}

// ---- Rust std BinaryHeap<OrderedNode> restated (alloc::collections::binary_heap: push = sift_up,
// pop = swap with the last + sift_down_to_bottom + sift_up) on (unsigned order key, position) pairs.
// OrderedNode compares by distance only (graph.rs:117-121), so WHICH of several rows tied at the k-th
// distance survives FlatIndex::search's `if root.dist > dist { pop; push }` loop (flat/index.rs:116-126)
// depends on this exact sift order.  The parallel selections below return the k smallest (distance,
// position) pairs, which is the same SET unless more rows tie at the k-th distance than fit; exactly
// then (detected by selecting k + 1) the slot is replayed sequentially through this heap.
__device__ __forceinline__ void rheap_sift_up(uint32_t* hk, uint32_t* hp, uint32_t pos) {
  const uint32_t ek = hk[pos], ep = hp[pos];
  while (pos > 0) {
    const uint32_t parent = (pos - 1) >> 1;
    if (ek <= hk[parent]) break;
    hk[pos] = hk[parent];
    hp[pos] = hp[parent];
    pos = parent;
  }
  hk[pos] = ek;
  hp[pos] = ep;
}
__device__ __forceinline__ void rheap_push(uint32_t* hk, uint32_t* hp, uint32_t& len, uint32_t key, uint32_t pos) {
  hk[len] = key;
  hp[len] = pos;
  rheap_sift_up(hk, hp, len);
  ++len;
}
__device__ __forceinline__ void rheap_pop(uint32_t* hk, uint32_t* hp, uint32_t& len) {
  --len;
  if (len == 0) return;
  const uint32_t ek = hk[len], ep = hp[len];  // the last element moves to the root, then sinks to the bottom
  uint32_t pos = 0, child = 1;
  const uint32_t end = len;
  while (child + 1 < end) {
    if (hk[child] <= hk[child + 1]) child += 1;
    hk[pos] = hk[child];
    hp[pos] = hp[child];
    pos = child;
    child = 2 * pos + 1;
  }
  if (child + 1 == end) {
    hk[pos] = hk[child];
    hp[pos] = hp[child];
    pos = child;
  }
  hk[pos] = ek;
  hp[pos] = ep;
  rheap_sift_up(hk, hp, pos);
}
// FlatIndex::search's insertion rule for one row (flat/index.rs:116-126); keys are unsigned order keys
__device__ __forceinline__ void rheap_offer(uint32_t* hk, uint32_t* hp, uint32_t& len, uint32_t k, uint32_t key,
                                            uint32_t pos) {
  if (len < k) {
    rheap_push(hk, hp, len, key, pos);
  } else if (hk[0] > key) {
    rheap_pop(hk, hp, len);
    rheap_push(hk, hp, len, key, pos);
  }
}


// ------------------------------------------------------------------------------------------------
// block-level helpers
// ------------------------------------------------------------------------------------------------
struct KeyIdx {
  int32_t key;   // total-order key of the distance
  uint32_t idx;  // tie-breaker (position / id)
};
__device__ __forceinline__ bool ki_less(int32_t k1, uint64_t i1, int32_t k2, uint64_t i2) {
  return k1 < k2 || (k1 == k2 && i1 < i2);
}

// argmin over (key, tie) proposed by every thread of a 256/128-thread block; returns the winning
// thread id (all threads get it).  Threads with nothing to propose pass has=false.
template <int NT>
__device__ inline int block_argmin(bool has, int32_t key, uint64_t tie, int32_t* s_key,
                                   uint64_t* s_tie, int* s_tid) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int who = has ? tid : -1;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const int32_t ok = __shfl_xor_sync(0xffffffffu, key, off);
    const uint64_t ot = __shfl_xor_sync(0xffffffffu, tie, off);
    const int ow = __shfl_xor_sync(0xffffffffu, who, off);
    if (ow >= 0 && (who < 0 || ki_less(ok, ot, key, tie))) {
      key = ok; tie = ot; who = ow;
    }
  }
  if (lane == 0) { s_key[warp] = key; s_tie[warp] = tie; s_tid[warp] = who; }
  __syncthreads();
  if (tid == 0) {
    int bw = s_tid[0];
    int32_t bk = s_key[0];
    uint64_t bt = s_tie[0];
    for (int w = 1; w < NT / 32; ++w)
      if (s_tid[w] >= 0 && (bw < 0 || ki_less(s_key[w], s_tie[w], bk, bt))) {
        bw = s_tid[w]; bk = s_key[w]; bt = s_tie[w];
      }
    s_tid[NT / 32] = bw;
  }
  __syncthreads();
  const int winner = s_tid[NT / 32];
  __syncthreads();
  return winner;
}

// per-thread sorted (ascending) list of the k best (distance, position) seen by this thread
template <int KMAX>
struct ThreadTopK {
  float d[KMAX];
  uint32_t j[KMAX];
  int cnt = 0;
  __device__ __forceinline__ void push(float dist, uint32_t pos, int k) {
    const int32_t key = total_order_key(dist);
    if (cnt == k && !(key < total_order_key(d[cnt - 1]))) return;  // ties keep earlier rows
    int p = cnt < k ? cnt : k - 1;
    while (p > 0 && total_order_key(d[p - 1]) > key) {
      d[p] = d[p - 1];
      j[p] = j[p - 1];
      --p;
    }
    d[p] = dist;
    j[p] = pos;
    if (cnt < k) ++cnt;
  }
};

// ------------------------------------------------------------------------------------------------
// coarse probe selection: nprobes smallest (distance, id), ascending (kmeans.rs:1152-1157)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
select_probes_kernel(const float* __restrict__ all_dists, int K, int nprobes,
                     uint32_t* __restrict__ ids, float* __restrict__ dists) {
  __shared__ int32_t s_key[4];
  __shared__ uint64_t s_tie[4];
  __shared__ int s_tid[5];
  __shared__ int32_t prev_key;
  __shared__ uint32_t prev_id;
  const float* row = all_dists + (size_t)blockIdx.x * K;
  const int tid = threadIdx.x;
  bool first = true;
  for (int r = 0; r < nprobes; ++r) {
    int32_t bk = 0;
    uint32_t bi = 0;
    bool has = false;
    const int32_t pk = first ? 0 : prev_key;
    const uint32_t pi = first ? 0 : prev_id;
    for (int c = tid; c < K; c += 128) {
      const int32_t key = total_order_key(row[c]);
      if (!first && !ki_less(pk, pi, key, c)) continue;  // already emitted
      if (!has || ki_less(key, c, bk, bi)) { bk = key; bi = c; has = true; }
    }
    const int w = block_argmin<128>(has, bk, bi, s_key, s_tie, s_tid);
    if (w < 0) break;
    if (tid == w) {
      prev_key = bk;
      prev_id = bi;
      ids[(size_t)blockIdx.x * nprobes + r] = bi;
      dists[(size_t)blockIdx.x * nprobes + r] = row[bi];
    }
    __syncthreads();
    first = false;
  }
}

// LUT[m][c] = dist(q_m, cb[m][c])  (pq/distance.rs:38-56).  For the common sub-vector widths the
// codeword is fetched with 128-bit loads and the reference-order sum is fully unrolled.
template <int METRIC, int DS>
__device__ __forceinline__ float lut_entry_fixed(const float* __restrict__ qm, const float* __restrict__ cw) {
  float qv[DS], cv[DS];
#pragma unroll
  for (int t = 0; t < DS; t += 4) {
    const float4 a = *reinterpret_cast<const float4*>(qm + t);
    const float4 b = __ldg(reinterpret_cast<const float4*>(cw + t));
    qv[t] = a.x; qv[t + 1] = a.y; qv[t + 2] = a.z; qv[t + 3] = a.w;
    cv[t] = b.x; cv[t + 1] = b.y; cv[t + 2] = b.z; cv[t + 3] = b.w;
  }
  if (DS < 16) {  // tail-only path (l2.rs:69-79): plain left-to-right sum
    float s = 0.0f;
#pragma unroll
    for (int t = 0; t < DS; ++t) s = f_add(s, term<METRIC>(qv[t], cv[t]));
    return finish<METRIC>(f_add(s, 0.0f));
  } else {        // DS == 16: one chunk of 16 lanes, summed lane 0..15
    float t0 = 0.0f;
#pragma unroll
    for (int t = 0; t < 16; ++t) t0 = f_add(t0, f_add(0.0f, term<METRIC>(qv[t], cv[t])));
    return finish<METRIC>(f_add(0.0f, t0));
  }
}
template <int METRIC>
__device__ __forceinline__ void build_lut_smem(float* lut, const float* qr, const float* __restrict__ codebook,
                                               int M, int ds, int tid) {
  if (ds == 8) {
    for (int idx = tid; idx < M * 256; idx += 256)
      lut[idx] = lut_entry_fixed<METRIC, 8>(qr + (idx >> 8) * 8, codebook + (size_t)idx * 8);
  } else if (ds == 4) {
    for (int idx = tid; idx < M * 256; idx += 256)
      lut[idx] = lut_entry_fixed<METRIC, 4>(qr + (idx >> 8) * 4, codebook + (size_t)idx * 4);
  } else if (ds == 16) {
    for (int idx = tid; idx < M * 256; idx += 256)
      lut[idx] = lut_entry_fixed<METRIC, 16>(qr + (idx >> 8) * 16, codebook + (size_t)idx * 16);
  } else {
    for (int idx = tid; idx < M * 256; idx += 256)
      lut[idx] = dist_exact_thread<METRIC>(qr + (idx >> 8) * ds, codebook + (size_t)idx * ds, ds);
  }
}

// ------------------------------------------------------------------------------------------------
// the fused (residual query -> LUT -> code scan -> top-k) kernels: one CTA per (query, probed
// partition).  Partitions are processed in chunks of <= SCAN_CHUNK rows, the winners of a chunk
// joining the next chunk's candidate pool.
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_CHUNK = 4096;

__device__ __forceinline__ float key_to_float(int32_t key) {
  return __int_as_float(key ^ (int32_t)((uint32_t)(key >> 31) >> 1));
}

// ------------------------------------------------------------------------------------------------
// general k (16 < k <= 1024, e.g. k * refine_factor): per chunk the k-th smallest key is found by a
// 4-pass MSB radix select over shared-memory keys (256-bin histograms), everything below it is
// kept, ties AT the k-th key are resolved by position (earliest rows survive).
// ------------------------------------------------------------------------------------------------
// arguments shared by the fused scan kernels (one slot = one (query, probed partition) pair)
struct ScanArgs {
  const float* queries; int d; const float* centroids; const float* codebook; int M, ds;
  const uint32_t* probe_ids; int np; const uint64_t* part_offsets; const uint8_t* codes;
  const uint64_t* row_ids; int k; float* cand_d; uint64_t* cand_id; uint32_t* cand_cnt;
  ScanFilter flt;
};

template <int METRIC, int NBITS>
__device__ void radix_slot(const ScanArgs& a, size_t slot, bool replay) {
  extern __shared__ float smem[];
  constexpr int NCODE = 1 << NBITS;
  const int M = a.M, ds = a.ds, d = a.d, k = a.k, np = a.np;
  const int kk = k + 1;  // one more than asked for: exposes ties that overflow the k-th place
  const uint64_t* __restrict__ allow = a.flt.allow;
  const bool filtering = allow != nullptr || a.flt.range;
  float* lut = smem;                                                   // [M*NCODE] (8-bit: M*256)
  float* qr = lut + M * 256;                                           // [d]
  uint32_t* ukey = reinterpret_cast<uint32_t*>(qr + d);                // [SCAN_CHUNK + kk] order-preserving keys
  uint32_t* cpos = ukey + SCAN_CHUNK + kk;                             // [kk] positions of carried winners
  uint32_t* nkey = cpos + kk;                                          // [kk] next winners / the replay heap
  uint32_t* npos = nkey + kk;                                          // [kk]
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_prefix, s_need, s_eq, s_out, s_max, s_maxcnt;
  __shared__ int32_t s_key[8];
  __shared__ uint64_t s_tie[8];
  __shared__ int s_tid[9];
  __shared__ uint32_t prev_pos;
  const int tid = threadIdx.x;
  const int pi = (int)(slot % np);
  const size_t qi = slot / np;
  const uint32_t p = a.probe_ids[qi * np + pi];
  const uint64_t off = a.part_offsets[p];
  const uint32_t n_p = (uint32_t)(a.part_offsets[p + 1] - off);
  if (n_p == 0) {
    if (tid == 0) a.cand_cnt[slot] = 0;
    return;
  }
  const float* q = a.queries + qi * d;
  for (int t = tid; t < d; t += 256)
    qr[t] = METRIC == METRIC_DOT ? q[t] : __fsub_rn(q[t], a.centroids[(size_t)p * d + t]);  // v2.rs:316-332
  __syncthreads();
  if (NBITS == 8) {
    build_lut_smem<METRIC>(lut, qr, a.codebook, M, ds, tid);
  } else {
    for (int idx = tid; idx < M * NCODE; idx += 256)
      lut[idx] = dist_exact_thread<METRIC>(qr + (idx / NCODE) * ds, a.codebook + (size_t)idx * ds, ds);
  }
  __syncthreads();
  constexpr int CW_DIV = NBITS == 4 ? 2 : 1;
  const int cw = M / CW_DIV;  // code bytes per row
  const uint8_t* pc = a.codes + off * cw;
  const float dot_fix = (float)M - 1.0f;
  // ---- 4-bit (pq/distance.rs:147-242): rows [0, flat_num) and the last n_p % 16 rows are exact f32 sums;
  // the others go through the table quantised to u8 with qmin = min(table), qmax = max(flat rows).
  // With a prefilter the reference scores row by row with DistCalculator::distance (exact, pq/storage.rs:
  // 895-916), so every row is exact then.
  __shared__ uint8_t qt[NBITS == 4 ? 256 * 16 : 1];  // M <= 256 sub-vectors x 16 entries
  __shared__ float s_q[2];                                // qmin, (qmax - qmin) / 255
  const uint32_t flat_num = NBITS == 4 ? min((uint32_t)max(200, k), n_p) : 0;
  const uint32_t rem16 = NBITS == 4 ? n_p % 16 : 0;
  auto exact4 = [&](uint32_t j) -> float {  // two adds per byte, byte order
    const uint8_t* rp = pc + (size_t)j * cw;
    float dist = 0.0f;
    for (int i = 0; i < cw; ++i) {
      const uint8_t c = rp[i];
      dist = f_add(dist, lut[(2 * i) * 16 + (c & 0xF)]);
      dist = f_add(dist, lut[(2 * i + 1) * 16 + (c >> 4)]);
    }
    return dist;
  };
  if (NBITS == 4 && allow == nullptr) {
    int32_t mx = (int32_t)0x80000000;
    for (uint32_t j = tid; j < flat_num; j += 256) mx = max(mx, total_order_key(exact4(j)));
    float mn = __int_as_float(0x7f800000);
    for (int i = tid; i < M * 16; i += 256) mn = fminf(mn, lut[i]);
    // block reduce through the (still unused) selection scratch
    int32_t* r_mx = reinterpret_cast<int32_t*>(hist);
    float* r_mn = reinterpret_cast<float*>(ukey);
    r_mx[tid] = mx;
    r_mn[tid] = mn;
    __syncthreads();
    for (int o = 128; o >= 1; o >>= 1) {
      if (tid < o) {
        r_mx[tid] = max(r_mx[tid], r_mx[tid + o]);
        r_mn[tid] = fminf(r_mn[tid], r_mn[tid + o]);
      }
      __syncthreads();
    }
    const float qmax = key_to_float(r_mx[0]), qmin = r_mn[0];
    __syncthreads();
    const float factor = __fdiv_rn(255.0f, __fsub_rn(qmax, qmin));
    for (int i = tid; i < M * 16; i += 256) {
      const float v = roundf(__fmul_rn(__fsub_rn(lut[i], qmin), factor));
      qt[i] = (v != v) ? 0 : v <= 0.0f ? 0 : v >= 255.0f ? 255 : (uint8_t)v;
    }
    if (tid == 0) {
      s_q[0] = qmin;
      s_q[1] = __fdiv_rn(__fsub_rn(qmax, qmin), 255.0f);
    }
    __syncthreads();
  }
  // unsigned order key of row `row`'s distance (unsigned order == f32::total_cmp order)
  auto row_key = [&](uint32_t row) -> uint32_t {
    float dist = 0.0f;
    if (NBITS == 4) {
      if (allow != nullptr || row < flat_num || row >= n_p - rem16) {
        dist = exact4(row);
      } else {
        const uint8_t* rp = pc + (size_t)row * cw;
        uint32_t qs = 0;  // saturating u8 adds of non-negative terms == min(255, sum)
        for (int i2 = 0; i2 < cw; ++i2) {
          const uint8_t c = rp[i2];
          qs += qt[(2 * i2) * 16 + (c & 0xF)];
          qs += qt[(2 * i2 + 1) * 16 + (c >> 4)];
        }
        dist = __fadd_rn(__fmul_rn((float)min(qs, 255u), s_q[1]), s_q[0]);
      }
    } else if ((M & 15) == 0) {
      const uint4* rp = reinterpret_cast<const uint4*>(pc + (size_t)row * M);
      for (int c16 = 0; c16 < M / 16; ++c16) {
        const uint4 v = __ldg(rp + c16);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        const float* l0 = lut + c16 * 16 * 256;
#pragma unroll
        for (int aa = 0; aa < 4; ++aa)
#pragma unroll
          for (int bb = 0; bb < 4; ++bb)
            dist = f_add(dist, l0[(aa * 4 + bb) * 256 + ((w[aa] >> (8 * bb)) & 0xff)]);
      }
    } else {
      const uint8_t* rp = pc + (size_t)row * M;
      for (int m = 0; m < M; ++m) dist = f_add(dist, lut[m * 256 + rp[m]]);
    }
    if (METRIC == METRIC_DOT) dist = __fsub_rn(dist, dot_fix);
    return (uint32_t)total_order_key(dist) ^ 0x80000000u;
  };
  auto excluded = [&](uint32_t row, uint32_t key) -> bool {
    return !row_allowed(allow, off + row) || !key_in_range(a.flt, (int32_t)(key ^ 0x80000000u));
  };

  uint32_t nw = 0;
  if (!replay) {
    for (uint32_t c0 = 0; c0 < n_p; c0 += SCAN_CHUNK) {
      const uint32_t clen = min((uint32_t)SCAN_CHUNK, n_p - c0);
      for (uint32_t j = tid; j < clen; j += 256) {
        if (!row_allowed(allow, off + c0 + j)) {
          ukey[j] = 0xffffffffu;
          continue;
        }
        const uint32_t key = row_key(c0 + j);
        ukey[j] = key_in_range(a.flt, (int32_t)(key ^ 0x80000000u)) ? key : 0xffffffffu;
      }
      // carried winners sit at ukey[SCAN_CHUNK ..); pool element i: i < clen -> (ukey[i], c0+i), else carried
      __syncthreads();
      const uint32_t pool = clen + nw;
      auto key_at = [&](uint32_t i) { return i < clen ? ukey[i] : ukey[SCAN_CHUNK + (i - clen)]; };
      auto pos_at = [&](uint32_t i) { return i < clen ? c0 + i : cpos[i - clen]; };
      if (pool <= (uint32_t)kk) {
        for (uint32_t i = tid; i < pool; i += 256) { nkey[i] = key_at(i); npos[i] = pos_at(i); }
        __syncthreads();
        for (uint32_t i = tid; i < pool; i += 256) { ukey[SCAN_CHUNK + i] = nkey[i]; cpos[i] = npos[i]; }
        nw = pool;
        __syncthreads();
        continue;
      }
      if (tid == 0) { s_prefix = 0; s_need = (uint32_t)kk; }
      uint32_t mask = 0;
      for (int shift = 24; shift >= 0; shift -= 8) {
        hist[tid] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        for (uint32_t i = tid; i < pool; i += 256) {
          const uint32_t kv = key_at(i);
          if ((kv & mask) == prefix) atomicAdd(&hist[(kv >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
          uint32_t need = s_need, cum = 0;
          int b = 0;
          for (; b < 256; ++b) {
            if (cum + hist[b] >= need) break;
            cum += hist[b];
          }
          s_need = need - cum;
          s_prefix = prefix | ((uint32_t)b << shift);
          s_eq = hist[b];
        }
        mask |= 0xffu << shift;
        __syncthreads();
      }
      const uint32_t T = s_prefix, need = s_need, eq = s_eq;  // take all keys < T and `need` of the `eq` keys == T
      if (tid == 0) s_out = 0;
      __syncthreads();
      for (uint32_t i = tid; i < pool; i += 256) {
        const uint32_t kv = key_at(i);
        if (kv < T || (kv == T && eq == need)) {
          const uint32_t at = atomicAdd(&s_out, 1u);
          nkey[at] = kv;
          npos[at] = pos_at(i);
        }
      }
      __syncthreads();
      if (eq != need) {  // ties at the last key: the `need` smallest positions survive (rare)
        bool first = true;
        for (uint32_t r = 0; r < need; ++r) {
          uint32_t bp = 0xffffffffu;
          bool has = false;
          const uint32_t pp = first ? 0 : prev_pos;
          for (uint32_t i = tid; i < pool; i += 256) {
            if (key_at(i) != T) continue;
            const uint32_t ps = pos_at(i);
            if (!first && ps <= pp) continue;
            if (!has || ps < bp) { bp = ps; has = true; }
          }
          const int w = block_argmin<256>(has, 0, bp, s_key, s_tie, s_tid);
          if (tid == w) {
            prev_pos = bp;
            const uint32_t at = s_out;
            nkey[at] = T;
            npos[at] = bp;
            s_out = at + 1;
          }
          __syncthreads();
          first = false;
        }
      }
      const uint32_t got = s_out;  // == kk
      __syncthreads();
      for (uint32_t i = tid; i < got; i += 256) { ukey[SCAN_CHUNK + i] = nkey[i]; cpos[i] = npos[i]; }
      nw = got;
      __syncthreads();
    }
    // ---- the kk = k + 1 smallest (key, position) pairs are in hand.  Excluded rows (prefilter / range)
    // carry the maximal key and are dropped here; if the two largest survivors share a key, more rows tie
    // at the k-th distance than fit and the reference's heap decides -> replay
    if (tid == 0) { s_out = 0; s_max = 0; s_maxcnt = 0; }
    __syncthreads();
    for (uint32_t i = tid; i < nw; i += 256) {
      const uint32_t key = ukey[SCAN_CHUNK + i], pos = cpos[i];
      const bool keep = !filtering || (row_allowed(allow, off + pos) && !(a.flt.range && key == 0xffffffffu));
      if (keep) {
        const uint32_t at = atomicAdd(&s_out, 1u);
        nkey[at] = key;
        npos[at] = pos;
        atomicMax(&s_max, key);
      }
    }
    __syncthreads();
    nw = s_out;
    if (nw == (uint32_t)kk) {
      for (uint32_t i = tid; i < nw; i += 256)
        if (nkey[i] == s_max) atomicAdd(&s_maxcnt, 1u);
      __syncthreads();
      replay = s_maxcnt >= 2;  // block-uniform
    }
    if (!replay) {
      const bool drop_max = nw == (uint32_t)kk;
      const uint32_t mx = s_max;
      __syncthreads();
      if (tid == 0) s_out = 0;
      __syncthreads();
      for (uint32_t i = tid; i < nw; i += 256) {
        if (drop_max && nkey[i] == mx) continue;
        const uint32_t at = atomicAdd(&s_out, 1u);
        a.cand_d[slot * k + at] = key_to_float((int32_t)(nkey[i] ^ 0x80000000u));
        a.cand_id[slot * k + at] = a.row_ids[off + npos[i]];
      }
      __syncthreads();
      if (tid == 0) a.cand_cnt[slot] = s_out;
      return;
    }
    __syncthreads();
  }
  // ---- replay: the reference's own loop (flat/index.rs:116-165), rows in storage order, distances
  // computed in parallel one chunk ahead of the single thread that drives the heap
  uint32_t len = 0;
  for (uint32_t c0 = 0; c0 < n_p; c0 += SCAN_CHUNK) {
    const uint32_t clen = min((uint32_t)SCAN_CHUNK, n_p - c0);
    for (uint32_t j = tid; j < clen; j += 256) ukey[j] = row_key(c0 + j);
    __syncthreads();
    if (tid == 0) {
      for (uint32_t j = 0; j < clen; ++j) {
        const uint32_t key = ukey[j];
        if (filtering && excluded(c0 + j, key)) continue;
        rheap_offer(nkey, npos, len, (uint32_t)k, key, c0 + j);
      }
    }
    __syncthreads();
  }
  if (tid == 0) s_out = len;
  __syncthreads();
  len = s_out;
  for (uint32_t i = tid; i < len; i += 256) {
    a.cand_d[slot * k + i] = key_to_float((int32_t)(nkey[i] ^ 0x80000000u));
    a.cand_id[slot * k + i] = a.row_ids[off + npos[i]];
  }
  if (tid == 0) a.cand_cnt[slot] = len;
}

// grid (np, nq): one CTA per slot; or, with a replay list (slots the fast kernel could not settle because of
// ties at the k-th distance), a small persistent grid that replays the listed slots
template <int METRIC, int NBITS>
__global__ void __launch_bounds__(256)
ivfpq_scan_radix_kernel(const ScanArgs a, const uint32_t* __restrict__ rlist, const uint32_t* __restrict__ rcount) {
  if (rlist) {
    const uint32_t cnt = *rcount;
    for (uint32_t i = blockIdx.x; i < cnt; i += gridDim.x) {
      radix_slot<METRIC, NBITS>(a, rlist[i], true);
      __syncthreads();
    }
    return;
  }
  radix_slot<METRIC, NBITS>(a, (size_t)blockIdx.y * a.np + blockIdx.x, false);
}

// ---- warp-wide sorting network on packed (key, position) words -------------------------------------
// A candidate is one u64: (order-preserving u32 of the distance) << 32 | position inside the partition,
// so an unsigned compare IS the (distance, position) order every selection step needs ("ties keep the
// earlier row").  PACK_INF (no candidate) sorts last.
constexpr uint64_t PACK_INF = ~0ull;
__device__ __forceinline__ uint64_t pack_cand(int32_t key, uint32_t pos) {
  return ((uint64_t)((uint32_t)key ^ 0x80000000u) << 32) | pos;
}
__device__ __forceinline__ int32_t cand_key(uint64_t c) { return (int32_t)((uint32_t)(c >> 32) ^ 0x80000000u); }
__device__ __forceinline__ uint32_t cand_pos(uint64_t c) { return (uint32_t)c; }

// bitonic merge of a 32-lane bitonic sequence into ascending order (5 compare-exchange steps)
__device__ __forceinline__ uint64_t warp_bitonic_merge32(uint64_t v, int lane) {
#pragma unroll
  for (int j = 16; j >= 1; j >>= 1) {
    const uint64_t o = __shfl_xor_sync(0xffffffffu, v, j);
    const bool keep_min = (lane & j) == 0;
    v = (keep_min == (o < v)) ? o : v;
  }
  return v;
}
// full ascending sort of one value per lane (15 compare-exchange steps)
__device__ __forceinline__ uint64_t warp_sort32(uint64_t v, int lane) {
#pragma unroll
  for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
    for (int j = k2 >> 1; j >= 1; j >>= 1) {
      const uint64_t o = __shfl_xor_sync(0xffffffffu, v, j);
      const bool keep_min = ((lane & j) == 0) == ((lane & k2) == 0);
      v = (keep_min == (o < v)) ? o : v;
    }
  }
  return v;
}
// the 32 smallest of a shared-memory list, ascending, one per lane (lane r = r-th smallest)
__device__ __forceinline__ uint64_t warp_smallest32(const uint64_t* list, uint32_t cnt, int lane) {
  uint64_t best = warp_sort32(lane < (int)cnt ? list[lane] : PACK_INF, lane);
  for (uint32_t base = 32; base < cnt; base += 32) {
    uint64_t v = warp_sort32(base + lane < cnt ? list[base + lane] : PACK_INF, lane);
    v = __shfl_sync(0xffffffffu, v, 31 - lane);  // descending: min(best, v) is bitonic
    best = warp_bitonic_merge32(v < best ? v : best, lane);
  }
  return best;
}

// k <= 16.  Per chunk of 4096 rows every WARP works on its own 512 rows without block barriers:
//   Tw = k-th smallest of its 32 lane minima (one 32-lane sort; an upper bound of the warp's k-th
//   smallest element), the <= (k-1)*16+1 elements <= Tw are compacted into the warp's shared-memory
//   list and sorted 32 at a time; then warp 0 merges the 8 x k warp winners with the winners carried
//   from earlier chunks the same way.  All comparisons are on packed (key, position) words.
constexpr int SCAN_KFAST = 16;
constexpr int SCAN_WLIST = (SCAN_KFAST - 1) * 16 + 1;  // 241

template <int METRIC, bool FILTER>
__global__ void __launch_bounds__(256, 6)
ivfpq_scan_kernel(const ScanArgs a, uint32_t* __restrict__ rlist, uint32_t* __restrict__ rcount) {
  constexpr int RPT = SCAN_CHUNK / 256;  // rows per thread and chunk (16)
  extern __shared__ float smem[];
  const int M = a.M, ds = a.ds, d = a.d, k = a.k, np = a.np;
  const int kk = k + 1;  // <= SCAN_KFAST: one more than asked for, to expose ties that overflow the k-th place
  const uint64_t* __restrict__ allow = a.flt.allow;
  float* lut = smem;          // [M*256]
  float* qr = lut + M * 256;  // [d]
  __shared__ uint64_t wl[8][SCAN_WLIST];                 // per-warp compacted candidates
  __shared__ uint64_t fin[8 * SCAN_KFAST + SCAN_KFAST];  // 8 x kk warp winners, then the carried winners
  __shared__ uint64_t car[SCAN_KFAST];
  __shared__ uint32_t s_nw;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pi = blockIdx.x;
  const size_t qi = blockIdx.y;
  const uint32_t p = a.probe_ids[qi * np + pi];
  const uint64_t off = a.part_offsets[p];
  const uint32_t n_p = (uint32_t)(a.part_offsets[p + 1] - off);
  const size_t slot = qi * np + pi;
  if (n_p == 0) {
    if (tid == 0) a.cand_cnt[slot] = 0;
    return;
  }
  const float* q = a.queries + qi * d;
  for (int t = tid; t < d; t += 256)
    qr[t] = METRIC == METRIC_DOT ? q[t] : __fsub_rn(q[t], a.centroids[(size_t)p * d + t]);  // v2.rs:316-332
  if (tid == 0) s_nw = 0;
  __syncthreads();
  build_lut_smem<METRIC>(lut, qr, a.codebook, M, ds, tid);
  __syncthreads();

  const uint8_t* pc = a.codes + off * M;
  const float dot_fix = (float)M - 1.0f;
  for (uint32_t c0 = 0; c0 < n_p; c0 += SCAN_CHUNK) {
    const uint32_t clen = min((uint32_t)SCAN_CHUNK, n_p - c0);
    int32_t key[RPT];
    uint64_t mine = PACK_INF;  // this lane's smallest candidate
    uint32_t livemask = 0;     // FILTER: bit u = row u of this thread passed the prefilter and the range
    // rows of warp w in this chunk: c0 + w*512 + lane + 32*u  (a warp owns a contiguous 512-row slab)
    const uint32_t wbase = warp * (RPT * 32);
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
      const uint32_t j = wbase + lane + 32 * u;
      key[u] = 0x7fffffff;
      if (j < clen && (!FILTER || row_allowed(allow, off + c0 + j))) {
        float dist = 0.0f;
        if ((M & 15) == 0) {
          const uint4* rp = reinterpret_cast<const uint4*>(pc + (size_t)(c0 + j) * M);
          for (int c16 = 0; c16 < M / 16; ++c16) {
            const uint4 v = __ldg(rp + c16);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            const float* l0 = lut + c16 * 16 * 256;
#pragma unroll
            for (int aa = 0; aa < 4; ++aa)
#pragma unroll
              for (int bb = 0; bb < 4; ++bb)
                dist = f_add(dist, l0[(aa * 4 + bb) * 256 + ((w[aa] >> (8 * bb)) & 0xff)]);
          }
        } else {
          const uint8_t* rp = pc + (size_t)(c0 + j) * M;
          for (int m = 0; m < M; ++m) dist = f_add(dist, lut[m * 256 + rp[m]]);
        }
        if (METRIC == METRIC_DOT) dist = __fsub_rn(dist, dot_fix);  // pq/storage.rs:957-958
        const int32_t kv = total_order_key(dist);
        if (!FILTER || key_in_range(a.flt, kv)) {
          if (FILTER) livemask |= 1u << u;
          key[u] = kv;
          const uint64_t c = pack_cand(kv, c0 + j);
          mine = c < mine ? c : mine;
        }
      }
    }
    // ---- warp-local threshold: Tw = kk-th smallest lane minimum (PACK_INF if < kk lanes have rows)
    const uint64_t tw = __shfl_sync(0xffffffffu, warp_sort32(mine, lane), kk - 1);
    // ---- compact the warp's elements <= Tw (ballot-ranked: deterministic order, no atomics)
    uint32_t wcnt = 0;
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
      const uint32_t j = wbase + lane + 32 * u;
      const bool live = FILTER ? ((livemask >> u) & 1u) != 0 : j < clen;
      const uint64_t c = pack_cand(key[u], c0 + j);
      const bool take = live && c <= tw;
      const unsigned bal = __ballot_sync(0xffffffffu, take);
      if (take) wl[warp][wcnt + __popc(bal & ((1u << lane) - 1))] = c;
      wcnt += __popc(bal);
    }
    __syncwarp();
    // ---- the warp's kk smallest -> block list (lane r holds the r-th smallest; PACK_INF = none)
    {
      const uint64_t best = warp_smallest32(wl[warp], wcnt, lane);
      if (lane < kk) fin[warp * SCAN_KFAST + lane] = best;
    }
    __syncthreads();
    if (warp == 0) {  // merge: 8 x kk warp winners + carried winners -> kk block winners
      const uint32_t nw = s_nw;
      if (lane < kk) fin[8 * SCAN_KFAST + lane] = lane < (int)nw ? car[lane] : PACK_INF;
      __syncwarp();
      // the winners sit at fin[w * 16 + r], r < kk: visit them 32 at a time (2 warps' slots per pass)
      uint64_t best = PACK_INF;
      for (int base = 0; base < 9 * SCAN_KFAST; base += 32) {
        const int i = base + lane;
        uint64_t v = (i < 9 * SCAN_KFAST && (i % SCAN_KFAST) < kk) ? fin[i] : PACK_INF;
        v = warp_sort32(v, lane);
        if (base == 0) {
          best = v;
        } else {
          v = __shfl_sync(0xffffffffu, v, 31 - lane);
          best = warp_bitonic_merge32(v < best ? v : best, lane);
        }
      }
      if (lane < kk) car[lane] = best;
      const unsigned got = __ballot_sync(0xffffffffu, lane < kk && best != PACK_INF);
      if (lane == 0) s_nw = __popc(got);
    }
    __syncthreads();
  }
  // car[0..nw) ascending by (key, position).  If the k-th and the (k+1)-th share a key, more rows tie at the
  // k-th distance than fit: which of them the reference's BinaryHeap keeps depends on its sift order, so the
  // slot goes on the replay list (ivfpq_scan_radix_kernel in list mode restates that loop).
  uint32_t nw = s_nw;
  if (nw == (uint32_t)kk) {
    if (cand_key(car[k]) == cand_key(car[k - 1])) {
      if (tid == 0) {
        rlist[atomicAdd(rcount, 1u)] = (uint32_t)slot;
        a.cand_cnt[slot] = 0;
      }
      return;
    }
    nw = k;
  }
  for (uint32_t i = tid; i < nw; i += 256) {
    a.cand_d[slot * k + i] = key_to_float(cand_key(car[i]));
    a.cand_id[slot * k + i] = a.row_ids[off + cand_pos(car[i])];
  }
  if (tid == 0) a.cand_cnt[slot] = nw;
}

// ------------------------------------------------------------------------------------------------
// Conflict-free scan for the headline shape (8-bit codes, M = 16 sub-spaces of 8 dimensions; C1 / C3).
//
// What bounds the scan above is the shared-memory gather: 32 lanes look up LUT[m][code] for the SAME m and random
// codes, i.e. random banks -- 3.3 wavefronts per request (ncu: 454 M bank conflicts per 10 000 x 10 probes) -- and
// every (query, partition) CTA re-reads the 128 KB codebook through L2.  This kernel removes both:
//
//  * the LUT is stored as [code][team][copy][m] (two copies per team, 256 B per code for the CTA's two teams):
//    sub-space m lives in bank m (copy 0) and 16 + m (copy 1).  Lane l works on sub-space (t - l) mod 16 at step t, lanes 0-15 on copy 0 and lanes 16-31 on copy 1,
//    so the 32 lookups of a request always hit 32 different banks: ONE wavefront.
//  * the reference's sum is m-ascending and sequential in f32, so a lane cannot start its row at m != 0.  Instead
//    the lanes are SKEWED IN TIME: lane l starts each row l steps late.  The index keeps, next to the row-major
//    codes, a skewed copy (`build_skew_codes`): per 512-row slab and lane the 16 rows of that lane (rows l + 32 i)
//    form one byte stream that is preceded by l mod 16 pad bytes and cut into 17 units of 16 bytes, unit (r, lane)
//    at (r * 32 + lane) * 16 -- one coalesced 128-bit load per lane and round, and byte t of a unit is a
//    compile-time register/byte position.  Two accumulators take the steps before / after the lane's row boundary,
//    selected by per-lane 0/1 weights through FFMA: fma(v, 1, acc) is the reference's separately rounded add,
//    fma(v, 0, acc) leaves acc unchanged (all LUT entries finite, checked while the LUT is built; a slot whose
//    LUT is not goes to the replay list).  Per lookup: one PRMT (code byte -> address bits 8-15, the lane's
//    bank bits into the low byte), LDS, 2 FFMA.
//  * persistent CTAs (one per SM, two teams of 8 warps) keep the codebook in shared memory (padded so that the
//    16 sub-spaces a half-warp reads are in different banks) and build each slot's LUT from there; thread (m, c)
//    keeps its residual sub-vector in registers.  Team barriers are named barriers, so one team scans while the
//    other builds its LUT.
// Distances, candidate order and the tie / replay rule are those of ivfpq_scan_kernel (same bits).
// ------------------------------------------------------------------------------------------------
constexpr int SKEW_ROUNDS = 17;                          // 16 rows per lane and slab + one unit of skew
constexpr int SKEW_SLAB_ROWS = 512;
constexpr int SKEW_SLAB_BYTES = SKEW_ROUNDS * 32 * 16;   // 8704
constexpr int SKEW_CB_STRIDE = 256 * 8 + 4;              // floats per sub-space in shared memory (+16 B pad)
constexpr int SKEW_LUT_BYTES = 256 * 256;                // both teams' LUTs, interleaved per code
constexpr int SKEW_LIST = 1024;                          // capacity of a team's candidate list (two teams)
constexpr int SKEW_SMALL_BYTES = 2 * SCAN_KFAST * 8 + 8 * 32 * 4 + 8 * 4 + 16;   // per team (sized for 8 warps): winners, lane minima, ...
// the LUT at shared address 0x10000: [base, 0x10000) holds 7 codebook sub-spaces + the small scratch, above the LUT
// come 9 sub-spaces and the two candidate lists; the dynamic allocation covers the highest address for base = 0
constexpr uint32_t SKEW_MAX_BASE = 0x10000u - (7 * SKEW_CB_STRIDE * 4 + 4 * SKEW_SMALL_BYTES);
constexpr int SKEW_SMEM_BYTES = 0x10000 + SKEW_LUT_BYTES + 112 + 9 * SKEW_CB_STRIDE * 4 + 2 * SKEW_LIST * 8;

__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
template <int NT>
__device__ __forceinline__ void team_sync(int team) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(NT) : "memory");
}
template <int NT>
__device__ __forceinline__ bool team_or(int team, bool v) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %2, 0;\n\tbar.red.or.pred p, %1, %3, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r)
      : "r"(team + 1), "r"((uint32_t)v), "n"(NT)
      : "memory");
  return r != 0;
}

// slab_off[p] = number of 512-row slabs before partition p (exclusive scan of ceil(n_p / 512)); slab_off[K] = total
__global__ void __launch_bounds__(1024)
skew_offsets_kernel(const uint64_t* __restrict__ part_offsets, int K, uint64_t* __restrict__ slab_off) {
  __shared__ uint64_t part[1024];
  const int tid = threadIdx.x;
  const int per = (K + 1023) / 1024;
  const int b = tid * per, e = min(K, b + per);
  uint64_t s = 0;
  for (int p = b; p < e; ++p) s += (part_offsets[p + 1] - part_offsets[p] + SKEW_SLAB_ROWS - 1) / SKEW_SLAB_ROWS;
  part[tid] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const uint64_t v = tid >= o ? part[tid - o] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  uint64_t run = tid ? part[tid - 1] : 0;
  for (int p = b; p < e; ++p) {
    slab_off[p] = run;
    run += (part_offsets[p + 1] - part_offsets[p] + SKEW_SLAB_ROWS - 1) / SKEW_SLAB_ROWS;
  }
  if (tid == 1023) slab_off[K] = part[1023];
}

// one warp per slab: unit (r, lane) = bytes [16 r - l16, 16 r - l16 + 16) of the lane's row stream (rows lane + 32 i)
__global__ void __launch_bounds__(256)
skew_fill_kernel(const uint64_t* __restrict__ part_offsets, int K, const uint64_t* __restrict__ slab_off,
                 const uint8_t* __restrict__ codes, uint8_t* __restrict__ skew) {
  const uint64_t nslab = slab_off[K];
  const int lane = threadIdx.x & 31, l16 = lane & 15;
  for (uint64_t s = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5); s < nslab; s += (uint64_t)gridDim.x * 8) {
    int lo = 0, hi = K;  // last p with slab_off[p] <= s  (empty partitions share their successor's offset)
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (slab_off[mid] <= s) lo = mid; else hi = mid;
    }
    const int p = lo;
    const uint64_t off = part_offsets[p];
    const uint32_t n_p = (uint32_t)(part_offsets[p + 1] - off);
    const uint32_t base = (uint32_t)(s - slab_off[p]) * SKEW_SLAB_ROWS;
    const uint4* rows = reinterpret_cast<const uint4*>(codes) + off;
    uint4* out = reinterpret_cast<uint4*>(skew + s * SKEW_SLAB_BYTES) + lane;
    uint4 prev = make_uint4(0, 0, 0, 0);
    for (int r = 0; r < SKEW_ROUNDS; ++r) {
      const uint32_t j = base + lane + 32 * r;
      const uint4 cur = (r < 16 && j < n_p) ? __ldg(rows + j) : make_uint4(0, 0, 0, 0);
      uint4 u = cur;
      if (l16) {  // bytes [16 - l16, 32 - l16) of prev|cur
        const uint32_t w[8] = {prev.x, prev.y, prev.z, prev.w, cur.x, cur.y, cur.z, cur.w};
        const int b0 = 16 - l16, wq = b0 >> 2, sh = (b0 & 3) * 8;
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t lo32 = 0, hi32 = 0;
#pragma unroll
          for (int q = 0; q < 8; ++q) {  // static indexing of w[]
            if (q == wq + i) lo32 = w[q];
            if (q == wq + i + 1) hi32 = w[q];
          }
          o[i] = sh ? (lo32 >> sh) | (hi32 << (32 - sh)) : lo32;
        }
        u = make_uint4(o[0], o[1], o[2], o[3]);
      }
      out[r * 32] = u;
      prev = cur;
    }
  }
}

template <int METRIC, bool FILTER, int NTEAM>
__global__ void __launch_bounds__(512, 1)
ivfpq_scan_skew_kernel(const ScanArgs a, const uint64_t* __restrict__ slab_off, const uint8_t* __restrict__ skew,
                       uint32_t nslots, uint32_t* __restrict__ rlist, uint32_t* __restrict__ rcount) {
  extern __shared__ __align__(16) unsigned char sk_smem[];
  const int d = a.d, k = a.k, np = a.np;
  const int kk = k + 1;  // one more than asked for, to expose ties that overflow the k-th place
  const uint64_t* __restrict__ allow = a.flt.allow;
  // NTEAM = 2: teams of 8 warps, two LUT copies (no bank conflicts).  NTEAM = 4: teams of 4 warps, ONE copy each
  // (lanes l and l + 16 share a bank: two wavefronts per request) -- twice as many independent teams to fill the
  // issue slots a team leaves empty at its barriers and in its low-parallelism phases.
  constexpr int TT = 512 / NTEAM, TW = TT / 32, COPIES = NTEAM == 2 ? 2 : 1;
  constexpr int LIST = SKEW_LIST * 2 / NTEAM;          // candidate-list capacity per team
  constexpr uint32_t CHUNK = TW * SKEW_SLAB_ROWS;       // rows a team scans between two selections
  const int tid = threadIdx.x, team = tid / TT, ttid = tid % TT, lane = tid & 31, warp = ttid >> 5;
  const int l16 = lane & 15, half = lane >> 4;
  // Shared-memory map.  The LUT sits at SHARED ADDRESS 0x10000 exactly, so that a lookup address is
  // 0x10000 | code << 8 | bank bits -- all of it produced by the one byte permute.  The codebook is split around
  // it (sub-spaces 0-6 below, 7-15 above), the small per-team scratch goes below, the candidate lists above.
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sk_smem);
  if (sbase > SKEW_MAX_BASE) {  // never seen (the runtime reserves 1 KB: sbase = 0x400); the exact replay takes every slot
    for (uint32_t slot = blockIdx.x * 512 + tid; slot < nslots; slot += gridDim.x * 512) {
      rlist[atomicAdd(rcount, 1u)] = slot;
      a.cand_cnt[slot] = 0;
    }
    return;
  }
  unsigned char* lut_g = sk_smem + (0x10000u - sbase);                             // generic pointer to the LUT
  float* lut2 = reinterpret_cast<float*>(lut_g) + team * (16 * COPIES);            // [256 codes][64]: + copy * 16 + m
  float* cb_lo = reinterpret_cast<float*>(sk_smem);                                // sub-spaces 0..6
  // sub-spaces 7..15; the 112 bytes keep sub-space m in 16-byte bank group (m + 2 c + h) mod 8 on both sides of the LUT
  float* cb_hi = reinterpret_cast<float*>(lut_g + SKEW_LUT_BYTES + 112);
  unsigned char* tb = sk_smem + 7 * SKEW_CB_STRIDE * 4 + team * SKEW_SMALL_BYTES;  // small scratch (below the LUT)
  uint64_t* car = reinterpret_cast<uint64_t*>(tb);                                 // [2][SCAN_KFAST] winners so far
  int32_t* wmin = reinterpret_cast<int32_t*>(car + 2 * SCAN_KFAST);               // [TW][32] lane minima
  int32_t* s_tw = wmin + TW * 32;                                                  // [8] warp thresholds
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_tw + 8);                         // candidates in tl
  uint64_t* tl = reinterpret_cast<uint64_t*>(lut_g + SKEW_LUT_BYTES + 112 + 9 * SKEW_CB_STRIDE * 4) + team * LIST;

  // codebook -> shared memory, once per CTA (sub-space stride padded by 16 B)
  for (int i = tid; i < 16 * 256 * 2; i += 512) {
    const int e = i >> 1, m = e >> 8;
    const float4 v = __ldg(reinterpret_cast<const float4*>(a.codebook) + i);
    float* dstm = m < 7 ? cb_lo + m * SKEW_CB_STRIDE : cb_hi + (m - 7) * SKEW_CB_STRIDE;
    *reinterpret_cast<float4*>(dstm + (e & 255) * 8 + (i & 1) * 4) = v;
  }
  __syncthreads();

  // per-lane constants of the skewed schedule
  const int th = l16 ? l16 : 16;  // steps [0, th) of a round still belong to the row begun one round earlier
  float wA[16], wB[16];
  uint32_t lp[16];  // low address byte of LUT[.][team][this lane's copy][sub-space of step t]
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    wA[t] = t < th ? 1.0f : 0.0f;
    wB[t] = t < th ? 0.0f : 1.0f;
    lp[t] = 0x10000u | (uint32_t)((team * (16 * COPIES) + (COPIES == 2 ? half * 16 : 0) + ((t - l16) & 15)) << 2);
  }
  const int sh = l16 != 0;  // the row finished in round u is row u - sh of the lane
  const int lm = ttid & 15;                  // LUT build: this thread's sub-space
  const float* cbm = lm < 7 ? cb_lo + lm * SKEW_CB_STRIDE : cb_hi + (lm - 7) * SKEW_CB_STRIDE;
  const float dot_fix = 16.0f - 1.0f;
  constexpr int32_t MAXKEY = 0x7fffffff;     // no live row carries it: the LUT is finite, sums are at most +inf
  int par = 0;                               // which half of car[] holds the winners

  // slot metadata is a chain of dependent global loads (probe id -> partition offsets -> slab offset): it is
  // fetched one slot ahead, and the first code unit of a slot is requested before its LUT is built
  const uint32_t stride = gridDim.x * NTEAM;
  uint32_t slot = blockIdx.x * NTEAM + team;
  uint32_t p_n = slot < nslots ? a.probe_ids[slot] : 0u;
  uint64_t off_n = a.part_offsets[p_n], end_n = a.part_offsets[p_n + 1], so_n = slab_off[p_n];
  for (; slot < nslots; slot += stride) {
    const size_t qi = slot / np;
    const uint32_t p = p_n;
    const uint64_t off = off_n;
    const uint32_t n_p = (uint32_t)(end_n - off_n);
    const uint8_t* sp = skew + so_n * SKEW_SLAB_BYTES;
    p_n = slot + stride < nslots ? a.probe_ids[slot + stride] : 0u;
    if (n_p == 0) {
      off_n = a.part_offsets[p_n]; end_n = a.part_offsets[p_n + 1]; so_n = slab_off[p_n];
      if (ttid == 0) a.cand_cnt[slot] = 0;
      continue;
    }
    const uint4* up0 = reinterpret_cast<const uint4*>(sp + (size_t)warp * SKEW_SLAB_BYTES) + lane;
    uint4 first_unit = make_uint4(0, 0, 0, 0);
    if ((uint32_t)warp * SKEW_SLAB_ROWS < n_p) first_unit = __ldg(up0);
    // ---- residual query of this thread's sub-space (v2.rs:316-332) and the LUT (pq/distance.rs:38-56).
    // No barrier is needed before lut2 is overwritten: every warp of the team left its scan before the last
    // team barrier of the previous slot.
    float qm[8];
    {
      const float4* q4 = reinterpret_cast<const float4*>(a.queries + qi * d + lm * 8);
      const float4* c4 = reinterpret_cast<const float4*>(a.centroids + (size_t)p * d + lm * 8);
      const float4 x0 = __ldg(q4), x1 = __ldg(q4 + 1);
      qm[0] = x0.x; qm[1] = x0.y; qm[2] = x0.z; qm[3] = x0.w; qm[4] = x1.x; qm[5] = x1.y; qm[6] = x1.z; qm[7] = x1.w;
      if (METRIC != METRIC_DOT) {
        const float4 y0 = __ldg(c4), y1 = __ldg(c4 + 1);
        const float cv[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
        for (int t = 0; t < 8; ++t) qm[t] = __fsub_rn(qm[t], cv[t]);
      }
    }
    bool bad = false;
#pragma unroll 4
    for (int i = 0; i < 256 / (TT / 16); ++i) {
      const int c = (ttid >> 4) + (TT / 16) * i;
      const float4 b0 = *reinterpret_cast<const float4*>(cbm + c * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(cbm + c * 8 + 4);
      const float cv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float s = 0.0f;
#pragma unroll
      for (int t = 0; t < 8; ++t) s = f_add(s, term<METRIC>(qm[t], cv[t]));
      const float val = finish<METRIC>(f_add(s, 0.0f));
      bad |= !(fabsf(val) < 1.0e30f);
      if (COPIES == 2) {
        lut2[c * 64 + half * 16 + lm] = val;
        lut2[c * 64 + (half ^ 1) * 16 + lm] = val;
      } else {
        lut2[c * 64 + lm] = val;
      }
    }
    off_n = a.part_offsets[p_n]; end_n = a.part_offsets[p_n + 1]; so_n = slab_off[p_n];
    bool replay = team_or<TT>(team, bad);
    uint32_t nw = 0;  // winners carried from earlier chunks (uniform)
    for (uint32_t c0 = 0; c0 < n_p && !replay; c0 += CHUNK) {
      const uint32_t clen = min(CHUNK, n_p - c0);
      const uint32_t wbase = warp * SKEW_SLAB_ROWS;
      if (ttid == 0) *s_cnt = 0;  // read last before the previous chunk's / slot's final barrier
      int32_t key[SKEW_ROUNDS];
      int32_t mk = MAXKEY;  // this lane's smallest live key
      if (wbase < clen) {   // warp-uniform: this warp's slab exists
        // rows of this lane: wbase + lane + 32 i < clen, i < 16; they finish in rounds sh .. sh + cnt - 1
        uint32_t livemask = 0;
        if (!FILTER) {
          const uint32_t first = wbase + lane;
          const uint32_t cnt = first < clen ? min(16u, (clen - first + 31u) >> 5) : 0u;
          livemask = ((1u << cnt) - 1u) << sh;
        }
        const uint4* up = reinterpret_cast<const uint4*>(sp + (size_t)((c0 >> 9) + warp) * SKEW_SLAB_BYTES) + lane;
        float A = 0.0f, B = 0.0f;
        uint4 cur = c0 == 0 ? first_unit : __ldg(up);
#pragma unroll
        for (int r = 0; r < SKEW_ROUNDS; ++r) {
          uint4 nxt = cur;
          if (r + 1 < SKEW_ROUNDS) nxt = __ldg(up + (r + 1) * 32);
          const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            // shared address = 0x10000 | code << 8 | bank bits: one byte permute (bytes 0, 2, 3 <- lp, byte 1 <- code)
            const float v = lds_f32(__byte_perm(w[t >> 2], lp[t], 0x7604u | ((uint32_t)(t & 3) << 4)));
            A = __fmaf_rn(v, wA[t], A);
            B = __fmaf_rn(v, wB[t], B);
          }
          float dist = A;
          A = B;
          B = 0.0f;
          if (METRIC == METRIC_DOT) dist = __fsub_rn(dist, dot_fix);  // pq/storage.rs:957-958
          const int32_t kv = total_order_key(dist);
          bool live;
          if (FILTER) {
            const int ri = r - sh;
            const uint32_t j = wbase + lane + 32 * ri;
            live = ri >= 0 && ri < 16 && j < clen && row_allowed(allow, off + c0 + j) && key_in_range(a.flt, kv);
          } else {
            live = ((livemask >> r) & 1u) != 0;
          }
          key[r] = live ? kv : MAXKEY;
          mk = min(mk, key[r]);
          cur = nxt;
        }
      } else {
#pragma unroll
        for (int r = 0; r < SKEW_ROUNDS; ++r) key[r] = MAXKEY;
      }
      // ---- selection.  With two teams per SM nothing hides the dependent shuffle steps of sorting networks, and
      // instruction issue is what bounds the kernel, so: (1) per warp, Tw = kk-th smallest lane minimum = the
      // largest lane minimum with fewer than kk smaller ones (32 broadcast reads + one warp reduction);
      wmin[warp * 32 + lane] = mk;
      __syncwarp();
      {
        int lt = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) lt += wmin[warp * 32 + j] < mk ? 1 : 0;
        const int32_t twv = __reduce_max_sync(0xffffffffu, lt < kk ? mk : (int32_t)0x80000000);
        if (lane == 0) s_tw[warp] = twv;
      }
      team_sync<TT>(team);
      // (2) T = the smallest warp threshold: at least kk rows of the team have key <= T; every row with key <= T
      // goes to the team list (typically kk + a few rows; ballots that come back empty cost three instructions);
      int32_t T = s_tw[0];
#pragma unroll
      for (int w = 1; w < TW; ++w) T = min(T, s_tw[w]);
      T = min(T, MAXKEY - 1);
      if (__any_sync(0xffffffffu, mk <= T)) {
#pragma unroll
        for (int u = 0; u < SKEW_ROUNDS; ++u) {
          const bool take = key[u] <= T;
          const unsigned bal = __ballot_sync(0xffffffffu, take);
          if (bal) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(s_cnt, (uint32_t)__popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0) + __popc(bal & ((1u << lane) - 1));
            if (take && base < (uint32_t)LIST) tl[base] = pack_cand(key[u], c0 + wbase + lane + 32 * (u - sh));
          }
        }
      }
      team_sync<TT>(team);
      // (3) the kk smallest of list + carried winners by RANK (packed (key, position) words are unique): thread i
      // counts the entries smaller than its own and stores it at that rank.
      const uint32_t cnt = *s_cnt;
      if (cnt > (uint32_t)LIST) {  // a flood of equal keys: the exact replay takes the slot
        replay = true;
      } else {
        const uint32_t tot = cnt + nw;
        const uint64_t* cold = car + par * SCAN_KFAST;
        uint64_t* cnew = car + (par ^ 1) * SCAN_KFAST;
        for (uint32_t i = ttid; i < tot; i += TT) {
          const uint64_t v = i < cnt ? tl[i] : cold[i - cnt];
          int rank = 0;
          for (uint32_t j = 0; j < cnt; ++j) rank += tl[j] < v ? 1 : 0;
          for (uint32_t j = 0; j < nw; ++j) rank += cold[j] < v ? 1 : 0;
          if (rank < kk) cnew[rank] = v;
        }
        nw = min((uint32_t)kk, tot);
        par ^= 1;
      }
      team_sync<TT>(team);
    }
    const uint64_t* win = car + par * SCAN_KFAST;  // ascending by (key, position)
    // If the k-th and the (k+1)-th share a key, more rows tie at the k-th distance than fit: which of them the
    // reference's BinaryHeap keeps depends on its sift order, so the slot goes on the replay list.
    if (!replay && nw == (uint32_t)kk) {
      if (cand_key(win[k]) == cand_key(win[k - 1])) replay = true;
      nw = k;
    }
    if (replay) {
      if (ttid == 0) {
        rlist[atomicAdd(rcount, 1u)] = slot;
        a.cand_cnt[slot] = 0;
      }
      continue;
    }
    for (uint32_t i = ttid; i < nw; i += TT) {
      a.cand_d[(size_t)slot * k + i] = key_to_float(cand_key(win[i]));
      a.cand_id[(size_t)slot * k + i] = a.row_ids[off + cand_pos(win[i])];
    }
    if (ttid == 0) a.cand_cnt[slot] = nw;
  }
}

// ------------------------------------------------------------------------------------------------
// IVF_FLAT: exact distances of the query to every row of a probed partition
// (FlatDistanceCal::distance_all, lance-index/src/vector/flat/storage.rs:397-403) + top-k.
// 16 lanes per row: lane l owns the reference's lane-accumulator l (elements 16c + l), so the L2 /
// dot results are bit-identical to l2.rs:57-91 / dot.rs:30-58; cosine follows cosine.rs:143-174 in
// structure (f32 FMA lanes) and is checked to the reference's own tolerance.
// ------------------------------------------------------------------------------------------------
// element of a stored / raw vector as f32 (l2.rs:100-106,156: f16 / bf16 elements are converted one by one)
template <class T> __device__ __forceinline__ float ldf(const T* p, int e);
template <> __device__ __forceinline__ float ldf<float>(const float* p, int e) { return p[e]; }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p, int e) { return __half2float(p[e]); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p, int e) { return __bfloat162float(p[e]); }
template <> __device__ __forceinline__ float ldf<uint8_t>(const uint8_t* p, int e) { return (float)p[e]; }

template <int METRIC, class T = float>
__device__ __forceinline__ float flat_row_distance(const float* __restrict__ q, const T* __restrict__ v,
                                                   int d, int l, unsigned mask, float q_norm) {
  const int n16 = d & ~15;
  if (METRIC == METRIC_COSINE) {
    float xy = 0.0f, yy = 0.0f;
    for (int e = l; e < d; e += 16) {
      const float y = ldf<T>(v, e);
      xy = fmaf(q[e], y, xy);
      yy = fmaf(y, y, yy);
    }
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
      xy += __shfl_xor_sync(mask, xy, off, 16);
      yy += __shfl_xor_sync(mask, yy, off, 16);
    }
    return 1.0f - xy / q_norm / sqrtf(yy);
  }
  float acc = 0.0f;
  for (int e = l; e < n16; e += 16) acc = f_add(acc, term<METRIC>(q[e], ldf<T>(v, e)));
  float s = 0.0f;  // sequential tail, every lane redundantly (l2.rs:69-79)
  for (int e = n16; e < d; ++e) s = f_add(s, term<METRIC>(q[e], ldf<T>(v, e)));
  float t = 0.0f;
#pragma unroll
  for (int qq = 0; qq < 16; ++qq) t = f_add(t, __shfl_sync(mask, acc, qq, 16));
  return finish<METRIC>(f_add(s, t));
}

template <int METRIC, class T>
__global__ void __launch_bounds__(256)
ivfflat_scan_kernel(const float* __restrict__ queries, int d, const uint32_t* __restrict__ probe_ids,
                    int np, const uint64_t* __restrict__ part_offsets,
                    const T* __restrict__ vectors, const uint64_t* __restrict__ row_ids, int k,
                    float* __restrict__ cand_d, uint64_t* __restrict__ cand_id,
                    uint32_t* __restrict__ cand_cnt, const ScanFilter flt) {
  extern __shared__ float smem[];
  __shared__ uint32_t s_outc;
  const int kk = k + 1;  // see radix_slot: exposes ties that overflow the k-th place
  const uint64_t* __restrict__ allow = flt.allow;
  const bool filtering = allow != nullptr || flt.range;
  float* qs = smem;                          // [d]
  float* cd = qs + d;                        // [SCAN_CHUNK + kk]
  uint32_t* cp = reinterpret_cast<uint32_t*>(cd + SCAN_CHUNK + kk);
  float* wd = reinterpret_cast<float*>(cp + kk);
  uint32_t* wp = reinterpret_cast<uint32_t*>(wd + kk);
  __shared__ int32_t s_key[8];
  __shared__ uint64_t s_tie[8];
  __shared__ int s_tid[9];
  __shared__ int32_t prev_key;
  __shared__ uint32_t prev_pos;
  __shared__ float s_qnorm;
  const int tid = threadIdx.x, l = tid & 15;
  const unsigned hmask = 0xffffu << (16 * ((tid >> 4) & 1));
  const int pi = blockIdx.x;
  const size_t qi = blockIdx.y;
  const uint32_t p = probe_ids[qi * np + pi];
  const uint64_t off = part_offsets[p];
  const uint32_t n_p = (uint32_t)(part_offsets[p + 1] - off);
  const size_t slot = qi * np + pi;
  if (n_p == 0) {
    if (tid == 0) cand_cnt[slot] = 0;
    return;
  }
  for (int t = tid; t < d; t += 256) qs[t] = queries[qi * d + t];
  __syncthreads();
  if (METRIC == METRIC_COSINE && tid < 32) {  // norm_l2(query): 16 lanes + sqrt (norm_l2.rs:106-130)
    float a = 0.0f;
    for (int e = (tid & 15); e < d; e += 16) a = fmaf(qs[e], qs[e], a);
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o, 16);
    if (tid == 0) s_qnorm = sqrtf(a);
  }
  __syncthreads();
  const float qn = METRIC == METRIC_COSINE ? s_qnorm : 0.0f;
  const float excluded_key = __int_as_float(0x7fffffff);  // maximal key of the total order
  uint32_t nw = 0;
  for (uint32_t c0 = 0; c0 < n_p; c0 += SCAN_CHUNK) {
    const uint32_t clen = min((uint32_t)SCAN_CHUNK, n_p - c0);
    for (uint32_t j = tid >> 4; j < clen; j += 16) {  // 16 rows per pass, 16 lanes each
      if (!row_allowed(allow, off + c0 + j)) {  // uniform per half-warp
        if (l == 0) cd[j] = excluded_key;
        continue;
      }
      const float dist = flat_row_distance<METRIC, T>(qs, vectors + (off + c0 + j) * (uint64_t)d, d, l, hmask, qn);
      if (l == 0) cd[j] = key_in_range(flt, total_order_key(dist)) ? dist : excluded_key;
    }
    __syncthreads();
    const uint32_t pool = clen + nw;
    const uint32_t rounds = pool < (uint32_t)kk ? pool : (uint32_t)kk;
    bool first = true;
    for (uint32_t r = 0; r < rounds; ++r) {
      int32_t bk = 0;
      uint32_t bpos = 0, bslot = 0;
      bool has = false;
      const int32_t pk = first ? 0 : prev_key;
      const uint32_t pp = first ? 0 : prev_pos;
      for (uint32_t i = tid; i < pool; i += 256) {
        const int32_t key = total_order_key(cd[i < clen ? i : SCAN_CHUNK + (i - clen)]);
        const uint32_t pos = i < clen ? c0 + i : cp[i - clen];
        if (!first && !ki_less(pk, pp, key, pos)) continue;
        if (!has || ki_less(key, pos, bk, bpos)) { bk = key; bpos = pos; bslot = i; has = true; }
      }
      const int w = block_argmin<256>(has, bk, bpos, s_key, s_tie, s_tid);
      if (tid == w) {
        prev_key = bk;
        prev_pos = bpos;
        wd[r] = cd[bslot < clen ? bslot : SCAN_CHUNK + (bslot - clen)];
        wp[r] = bpos;
      }
      __syncthreads();
      first = false;
    }
    for (uint32_t i = tid; i < rounds; i += 256) {
      cd[SCAN_CHUNK + i] = wd[i];
      cp[i] = wp[i];
    }
    nw = rounds;
    __syncthreads();
  }
  // winners ascending by (key, position) in cd[SCAN_CHUNK ..), cp[]; excluded rows (maximal key) sort last
  auto dropped = [&](uint32_t i) -> bool {
    return !row_allowed(allow, off + cp[i]) || (flt.range && __float_as_int(cd[SCAN_CHUNK + i]) == 0x7fffffff);
  };
  if (filtering) {
    uint32_t keep = nw;
    while (keep > 0 && dropped(keep - 1)) --keep;
    nw = keep;  // every thread computes the same value
  }
  bool replay = false;
  if (nw == (uint32_t)kk) {
    replay = total_order_key(cd[SCAN_CHUNK + k]) == total_order_key(cd[SCAN_CHUNK + k - 1]);
    nw = k;
  }
  if (!replay) {
    if (filtering) {  // (a NaN distance can leave a dropped row in front of the tail: re-test every entry)
      if (tid == 0) s_outc = 0;
      __syncthreads();
      for (uint32_t i = tid; i < nw; i += 256)
        if (!dropped(i)) {
          const uint32_t at = atomicAdd(&s_outc, 1u);
          cand_d[slot * k + at] = cd[SCAN_CHUNK + i];
          cand_id[slot * k + at] = row_ids[off + cp[i]];
        }
      __syncthreads();
      if (tid == 0) cand_cnt[slot] = s_outc;
      return;
    }
    for (uint32_t i = tid; i < nw; i += 256) {
      cand_d[slot * k + i] = cd[SCAN_CHUNK + i];
      cand_id[slot * k + i] = row_ids[off + cp[i]];
    }
    if (tid == 0) cand_cnt[slot] = nw;
    return;
  }
  // ---- ties overflow the k-th place: the reference's heap loop (flat/index.rs:116-165), see radix_slot
  __syncthreads();
  uint32_t* hk = reinterpret_cast<uint32_t*>(wd);
  uint32_t* hp = wp;
  uint32_t len = 0;
  for (uint32_t c0 = 0; c0 < n_p; c0 += SCAN_CHUNK) {
    const uint32_t clen = min((uint32_t)SCAN_CHUNK, n_p - c0);
    for (uint32_t j = tid >> 4; j < clen; j += 16) {
      const float dist = flat_row_distance<METRIC, T>(qs, vectors + (off + c0 + j) * (uint64_t)d, d, l, hmask, qn);
      if (l == 0) cd[j] = dist;
    }
    __syncthreads();
    if (tid == 0) {
      for (uint32_t j = 0; j < clen; ++j) {
        const int32_t key = total_order_key(cd[j]);
        if (filtering && (!row_allowed(allow, off + c0 + j) || !key_in_range(flt, key))) continue;
        rheap_offer(hk, hp, len, (uint32_t)k, (uint32_t)key ^ 0x80000000u, c0 + j);
      }
    }
    __syncthreads();
  }
  if (tid == 0) s_outc = len;
  __syncthreads();
  len = s_outc;
  for (uint32_t i = tid; i < len; i += 256) {
    cand_d[slot * k + i] = key_to_float((int32_t)(hk[i] ^ 0x80000000u));
    cand_id[slot * k + i] = row_ids[off + hp[i]];
  }
  if (tid == 0) cand_cnt[slot] = len;
}

// global merge per query: ascending (distance, row id), first k.  Candidate e of list pi of query qi sits
// at cand[pi * stride_p + qi * stride_q + e] (per-partition lists of one GPU: stride_p = k, stride_q = np * k;
// per-rank results gathered from a sharded index: stride_p = the rank stride, stride_q = k).
// Lists of up to MERGE_RANK_MAX candidates in total are merged by RANK COUNTING in shared memory: every candidate
// counts the candidates that precede it in (distance, row id) order -- the pairs are unique -- and the ones with rank
// < k are the output, already in place.  (The k-round argmin below re-reads all candidates from global memory per
// round: 3.6 ms for 10 000 queries x 10 lists x k = 100; it remains for larger totals.)
constexpr int MERGE_RANK_MAX = 2048;
__global__ void __launch_bounds__(256)
merge_rank_kernel(const float* __restrict__ cand_d, const uint64_t* __restrict__ cand_id,
                  const uint32_t* __restrict__ cand_cnt, int np, int k, size_t stride_p_d, size_t stride_p_id,
                  size_t stride_q, size_t cnt_stride_p, size_t cnt_stride_q, uint64_t* __restrict__ out_id,
                  float* __restrict__ out_d, uint32_t* __restrict__ out_cnt) {
  extern __shared__ __align__(16) unsigned char mr_smem[];
  const int total = np * k;
  uint64_t* s_id = reinterpret_cast<uint64_t*>(mr_smem);             // [total]
  int32_t* s_key = reinterpret_cast<int32_t*>(s_id + total);         // [total]; invalid entries: key = INT_MAX, id = ~0
  __shared__ uint32_t s_valid;
  const size_t qi = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid == 0) s_valid = 0;
  __syncthreads();
  uint32_t myvalid = 0;
  for (int c = tid; c < total; c += 256) {
    const int pi = c / k, e = c % k;
    const bool ok = (uint32_t)e < cand_cnt[pi * cnt_stride_p + qi * cnt_stride_q];
    s_key[c] = ok ? total_order_key(cand_d[pi * stride_p_d + qi * stride_q + e]) : 0x7fffffff;
    s_id[c] = ok ? cand_id[pi * stride_p_id + qi * stride_q + e] : ~0ull;
    myvalid += ok ? 1u : 0u;
  }
  if (myvalid) atomicAdd(&s_valid, myvalid);
  __syncthreads();
  for (int c = tid; c < total; c += 256) {
    const uint64_t id = s_id[c];
    if (id == ~0ull && s_key[c] == 0x7fffffff) continue;
    const int32_t key = s_key[c];
    int rank = 0;
    for (int j = 0; j < total; ++j) {
      const int32_t kj = s_key[j];
      rank += (kj < key || (kj == key && s_id[j] < id)) ? 1 : 0;
    }
    if (rank < k) {
      out_id[qi * k + rank] = id;
      out_d[qi * k + rank] = key_to_float(key);
    }
  }
  const int r = min((uint32_t)k, s_valid);
  for (int e = r + tid; e < k; e += 256) {
    out_id[qi * k + e] = ~0ull;
    out_d[qi * k + e] = __int_as_float(0x7f800000);
  }
  if (tid == 0 && out_cnt) out_cnt[qi] = r;
}

__global__ void __launch_bounds__(128)
merge_kernel(const float* __restrict__ cand_d, const uint64_t* __restrict__ cand_id,
             const uint32_t* __restrict__ cand_cnt, int np, int k, size_t stride_p_d, size_t stride_p_id,
             size_t stride_q, size_t cnt_stride_p, size_t cnt_stride_q, uint64_t* __restrict__ out_id,
             float* __restrict__ out_d, uint32_t* __restrict__ out_cnt) {
  __shared__ int32_t s_key[4];
  __shared__ uint64_t s_tie[4];
  __shared__ int s_tid[5];
  __shared__ int32_t prev_key;
  __shared__ uint64_t prev_id;
  const size_t qi = blockIdx.x;
  const int tid = threadIdx.x;
  const int total = np * k;
  bool first = true;
  int r = 0;
  for (; r < k; ++r) {
    int32_t bk = 0;
    uint64_t bi = 0;
    int bslot = -1;
    const int32_t pk = first ? 0 : prev_key;
    const uint64_t pid = first ? 0 : prev_id;
    for (int c = tid; c < total; c += 128) {
      const int pi = c / k, e = c % k;
      if ((uint32_t)e >= cand_cnt[pi * cnt_stride_p + qi * cnt_stride_q]) continue;
      const int32_t key = total_order_key(cand_d[pi * stride_p_d + qi * stride_q + e]);
      const uint64_t id = cand_id[pi * stride_p_id + qi * stride_q + e];
      if (!first && !ki_less(pk, pid, key, id)) continue;
      if (bslot < 0 || ki_less(key, id, bk, bi)) { bk = key; bi = id; bslot = c; }
    }
    const int w = block_argmin<128>(bslot >= 0, bk, bi, s_key, s_tie, s_tid);
    if (w < 0) break;
    if (tid == w) {
      prev_key = bk;
      prev_id = bi;
      out_id[qi * k + r] = bi;
      out_d[qi * k + r] = cand_d[(bslot / k) * stride_p_d + qi * stride_q + (bslot % k)];
    }
    __syncthreads();
    first = false;
  }
  for (int e = r + tid; e < k; e += 128) {
    out_id[qi * k + e] = ~0ull;
    out_d[qi * k + e] = __int_as_float(0x7f800000);
  }
  if (tid == 0 && out_cnt) out_cnt[qi] = r;
}

// ------------------------------------------------------------------------------------------------
// primitives exported one-to-one (used by the trait-level shim and by the parity tests)
// ------------------------------------------------------------------------------------------------
template <int METRIC>
__global__ void build_lut_kernel(const float* __restrict__ codebook, int M, int ncode, int ds,
                                 const float* __restrict__ query, float* __restrict__ lut) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * ncode) return;
  const int m = idx / ncode;
  lut[idx] = dist_exact_thread<METRIC>(query + m * ds, codebook + (size_t)idx * ds, ds);
}

__global__ void pq_scan_transposed_kernel(const float* __restrict__ lut, int M,
                                          const uint8_t* __restrict__ codes_t, uint64_t n,
                                          int is_dot, float* __restrict__ out) {
  extern __shared__ float s_lut[];
  for (int i = threadIdx.x; i < M * 256; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float dist = 0.0f;
  for (int m = 0; m < M; ++m) dist = f_add(dist, s_lut[m * 256 + codes_t[(size_t)m * n + j]]);
  if (is_dot) dist = __fsub_rn(dist, (float)M - 1.0f);
  out[j] = dist;
}

// FlatIndex::search over a distance array (flat/index.rs:97-127): the heap's final content, written
// ascending by (distance, row id).  Selection of k + 1 by per-thread sorted lists; when rows tie at the
// k-th distance beyond what fits, thread 0 replays the reference's loop through the Rust heap.
template <int KMAX>
__global__ void __launch_bounds__(256)
flat_topk_kernel(const float* __restrict__ dists, const uint64_t* __restrict__ row_ids, uint64_t n,
                 int k, const ScanFilter flt, uint64_t* __restrict__ out_id, float* __restrict__ out_d,
                 uint32_t* __restrict__ out_cnt) {
  __shared__ int32_t s_key[8];
  __shared__ uint64_t s_tie[8];
  __shared__ int s_tid[9];
  __shared__ uint32_t wk[KMAX], wpos[KMAX];  // winners: unsigned order key, position
  __shared__ uint32_t s_len;
  const int tid = threadIdx.x;
  const int kk = k + 1;
  ThreadTopK<KMAX> top;
  for (uint64_t j = tid; j < n; j += 256) {
    const float dv = dists[j];
    if (key_in_range(flt, total_order_key(dv))) top.push(dv, (uint32_t)j, kk);
  }
  int head = 0;
  uint32_t cnt = 0;
  for (int r = 0; r < kk; ++r) {
    const bool has = head < top.cnt;
    const int32_t key = has ? total_order_key(top.d[head]) : 0;
    const uint64_t tie = has ? top.j[head] : 0;
    const int w = block_argmin<256>(has, key, tie, s_key, s_tie, s_tid);
    if (w < 0) break;
    if (tid == w) {
      wk[r] = (uint32_t)key ^ 0x80000000u;
      wpos[r] = top.j[head];
      ++head;
    }
    ++cnt;
  }
  __syncthreads();
  if (cnt == (uint32_t)kk) {
    if (wk[k] == wk[k - 1]) {  // block-uniform: replay (see radix_slot)
      __syncthreads();
      if (tid == 0) {
        uint32_t len = 0;
        for (uint64_t j = 0; j < n; ++j) {
          const int32_t key = total_order_key(dists[j]);
          if (!key_in_range(flt, key)) continue;
          rheap_offer(wk, wpos, len, (uint32_t)k, (uint32_t)key ^ 0x80000000u, (uint32_t)j);
        }
        s_len = len;
      }
      __syncthreads();
      cnt = s_len;
    } else {
      cnt = k;
    }
  }
  // ascending (distance, row id) over the <= k survivors
  __shared__ int32_t prev_key;
  __shared__ uint64_t prev_id;
  bool first = true;
  for (uint32_t r = 0; r < cnt; ++r) {
    int32_t bk = 0;
    uint64_t bi = 0;
    bool has = false;
    const int32_t pk = first ? 0 : prev_key;
    const uint64_t pid = first ? 0 : prev_id;
    for (uint32_t i = tid; i < cnt; i += 256) {
      const int32_t key = (int32_t)(wk[i] ^ 0x80000000u);
      const uint64_t id = row_ids ? row_ids[wpos[i]] : (uint64_t)wpos[i];
      if (!first && !ki_less(pk, pid, key, id)) continue;
      if (!has || ki_less(key, id, bk, bi)) { bk = key; bi = id; has = true; }
    }
    const int w = block_argmin<256>(has, bk, bi, s_key, s_tie, s_tid);
    if (tid == w) {
      prev_key = bk;
      prev_id = bi;
      out_d[r] = key_to_float(bk);
      out_id[r] = bi;
    }
    __syncthreads();
    first = false;
  }
  if (tid == 0) *out_cnt = cnt;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
void find_partitions_f32(const float* centroids, int K, int d, int metric, const float* queries,
                         uint64_t nq, int nprobes, uint32_t* ids, float* dists) {
  if (nq == 0) return;
  DevBuf<float> all((size_t)nq * K);
  assign_f32(queries, nq, d, centroids, K, metric, nullptr, nullptr, nullptr, nullptr, all.p);
  LB2_LAUNCH("select_probes", select_probes_kernel, (unsigned)nq, 128, 0, all.p, K, nprobes, ids, dists);
}

// which fast kernel serves an 8-bit scan: LB2_SCAN=classic|skew overrides the size rule (tests run both)
static int scan_mode_env() {
  const char* e = getenv("LB2_SCAN");
  return !e ? 0 : (!strcmp(e, "classic") ? 1 : (!strcmp(e, "skew") ? 2 : 0));
}

template <int METRIC>
static void scan_launch(int nbits, dim3 grid, size_t smem, const ScanArgs& a, uint32_t* rlist, uint32_t* rcount,
                        const uint64_t* slab_off, const uint8_t* skew) {
  const bool filtering = a.flt.allow != nullptr || a.flt.range;
  if (nbits == 8 && a.k + 1 <= SCAN_KFAST) {
    const size_t smem_fast = sizeof(float) * ((size_t)a.M * 256 + a.d);
    LB2_CUDA(cudaMemsetAsync(rcount, 0, sizeof(uint32_t), ctx().stream));
    const uint64_t nslots = (uint64_t)grid.x * grid.y;
    const bool skew_ok = skew && a.M == 16 && a.ds == 8 && (reinterpret_cast<uintptr_t>(a.queries) & 15) == 0 &&
                         (size_t)SKEW_SMEM_BYTES <= ctx().smem_optin;
    // The persistent kernel wins at every batch size measured (profiles/scan_variants_r02.json: 19 vs 33 us for one
    // query, 3.9 vs 5.9 ms for 10 000 x 10 probes); LB2_SCAN=classic|skew and LB2_SCAN_TEAMS=2|4 override for tests.
    const bool use_skew = skew_ok && scan_mode_env() != 1;
    if (use_skew) {
      const char* te = getenv("LB2_SCAN_TEAMS");
      // four single-copy teams per SM once every SM has several slots per team; two double-copy teams below that
      const int nteam = te && atoi(te) == 2 ? 2 : (te && atoi(te) == 4 ? 4 : (nslots >= 16ull * ctx().num_sms ? 4 : 2));
      const unsigned g = (unsigned)std::min<uint64_t>((nslots + nteam - 1) / nteam, (uint64_t)ctx().num_sms);
      auto go = [&](auto kern) {
        set_smem(kern, SKEW_SMEM_BYTES);
        LB2_LAUNCH("pq_scan_skew", kern, g, 512, SKEW_SMEM_BYTES, a, slab_off, skew, (uint32_t)nslots, rlist, rcount);
      };
      if (filtering) {
        if (nteam == 2) go(ivfpq_scan_skew_kernel<METRIC, true, 2>); else go(ivfpq_scan_skew_kernel<METRIC, true, 4>);
      } else {
        if (nteam == 2) go(ivfpq_scan_skew_kernel<METRIC, false, 2>); else go(ivfpq_scan_skew_kernel<METRIC, false, 4>);
      }
    } else if (filtering) {  // filtered rows never enter the candidate lists
      set_smem(ivfpq_scan_kernel<METRIC, true>, smem_fast);
      LB2_LAUNCH("pq_scan", (ivfpq_scan_kernel<METRIC, true>), grid, 256, smem_fast, a, rlist, rcount);
    } else {
      set_smem(ivfpq_scan_kernel<METRIC, false>, smem_fast);
      LB2_LAUNCH("pq_scan", (ivfpq_scan_kernel<METRIC, false>), grid, 256, smem_fast, a, rlist, rcount);
    }
    // slots with ties beyond the k-th place (rare): the reference's heap loop, restated
    set_smem((ivfpq_scan_radix_kernel<METRIC, 8>), smem);
    const unsigned rgrid = (unsigned)std::min<uint64_t>((uint64_t)grid.x * grid.y, 4 * (uint64_t)ctx().num_sms);
    LB2_LAUNCH("pq_scan_tie_replay", (ivfpq_scan_radix_kernel<METRIC, 8>), rgrid, 256, smem, a,
               (const uint32_t*)rlist, (const uint32_t*)rcount);
    return;
  }
  if (nbits == 4) {
    set_smem((ivfpq_scan_radix_kernel<METRIC, 4>), smem);
    LB2_LAUNCH("pq_scan", (ivfpq_scan_radix_kernel<METRIC, 4>), grid, 256, smem, a, (const uint32_t*)nullptr,
               (const uint32_t*)nullptr);
    return;
  }
  set_smem((ivfpq_scan_radix_kernel<METRIC, 8>), smem);
  LB2_LAUNCH("pq_scan", (ivfpq_scan_radix_kernel<METRIC, 8>), grid, 256, smem, a, (const uint32_t*)nullptr,
             (const uint32_t*)nullptr);
}

// np lists of <= k candidates per query -> the k smallest by (distance, row id)
static void merge_lists(const char* name, uint64_t nq, const float* cand_d, const uint64_t* cand_id, const uint32_t* cand_cnt,
                        int np, int k, size_t stride_p_d, size_t stride_p_id, size_t stride_q, size_t cnt_stride_p,
                        size_t cnt_stride_q, uint64_t* out_ids, float* out_dists, uint32_t* out_counts) {
  if (nq == 0) return;
  const size_t total = (size_t)np * k;
  if (total <= (size_t)MERGE_RANK_MAX) {
    LB2_LAUNCH(name, merge_rank_kernel, (unsigned)nq, 256, total * 12, cand_d, cand_id, cand_cnt, np, k, stride_p_d, stride_p_id,
               stride_q, cnt_stride_p, cnt_stride_q, out_ids, out_dists, out_counts);
  } else {
    LB2_LAUNCH(name, merge_kernel, (unsigned)nq, 128, 0, cand_d, cand_id, cand_cnt, np, k, stride_p_d, stride_p_id, stride_q,
               cnt_stride_p, cnt_stride_q, out_ids, out_dists, out_counts);
  }
}

// the skewed copy of an index's codes (see ivfpq_scan_skew_kernel); sizes: slab_off u64[K + 1],
// skew (n / 512 + K) slabs of 8704 bytes at most
bool skew_layout_applies(int M, int d, int nbits) { return nbits == 8 && M == 16 && d == 128; }
size_t skew_bytes_bound(uint64_t n, int K) { return (size_t)(n / SKEW_SLAB_ROWS + (uint64_t)K) * SKEW_SLAB_BYTES; }
void build_skew_codes(const uint64_t* part_offsets, int K, const uint8_t* codes, uint64_t n, uint64_t* slab_off,
                      uint8_t* skew) {
  LB2_LAUNCH("skew_offsets", skew_offsets_kernel, 1, 1024, 0, part_offsets, K, slab_off);
  const unsigned g = (unsigned)std::min<uint64_t>(cdiv(n / SKEW_SLAB_ROWS + (uint64_t)K, 8), 8ull * ctx().num_sms);
  if (n) LB2_LAUNCH("skew_fill", skew_fill_kernel, std::max(1u, g), 256, 0, part_offsets, K, (const uint64_t*)slab_off, codes, skew);
}

void ivfpq_search_f32(const float* centroids, int K, int d, int metric, const float* codebook, int M,
                      int nbits, const uint64_t* part_offsets, const uint8_t* codes,
                      const uint64_t* row_ids, const float* queries, uint64_t nq, int k, int nprobes,
                      uint64_t* out_ids, float* out_dists, uint32_t* out_counts, const ScanFilter& flt,
                      const uint64_t* slab_off, const uint8_t* skew) {
  if (nq == 0 || k == 0) return;
  if (nbits != 8 && nbits != 4) fail(LB2_INVALID_ARG, "PQ: num_bits must be 4 or 8, got %d", nbits);
  if (nbits == 4 && (M % 2 != 0 || M > 256)) fail(LB2_UNSUPPORTED, "4-bit PQ needs an even num_sub_vectors <= 256");
  if (k > 1024) fail(LB2_UNSUPPORTED, "k (incl. refine factor) > 1024 is not implemented");
  const int np = nprobes < K ? nprobes : K;
  const int ds = d / M;
  const int cmetric = metric == METRIC_DOT ? METRIC_DOT : METRIC_L2;
  DevBuf<uint32_t> pids((size_t)nq * np), cand_cnt((size_t)nq * np);
  DevBuf<float> pd((size_t)nq * np), cand_d((size_t)nq * np * k);
  DevBuf<uint64_t> cand_id((size_t)nq * np * k);
  find_partitions_f32(centroids, K, d, cmetric, queries, nq, np, pids.p, pd.p);
  const size_t smem = sizeof(float) * ((size_t)M * 256 + d + SCAN_CHUNK + 4 * (size_t)(k + 1));
  if (smem > ctx().smem_optin) fail(LB2_UNSUPPORTED, "LUT of %zu bytes exceeds shared memory", smem);
  const uint64_t slab = 32768;  // grid.y limit: queries are processed in slabs
  DevBuf<uint32_t> rlist((size_t)std::min<uint64_t>(nq, slab) * np), rcount(1);
  for (uint64_t q0 = 0; q0 < nq; q0 += slab) {
    const uint64_t qn = std::min<uint64_t>(slab, nq - q0);
    dim3 g(np, (unsigned)qn);
    ScanArgs a{queries + q0 * d, d, centroids, codebook, M, ds, pids.p + q0 * np, np, part_offsets, codes, row_ids, k,
               cand_d.p + q0 * np * k, cand_id.p + q0 * np * k, cand_cnt.p + q0 * np, flt};
    if (cmetric == METRIC_DOT)
      scan_launch<METRIC_DOT>(nbits, g, smem, a, rlist.p, rcount.p, slab_off, skew);
    else
      scan_launch<METRIC_L2>(nbits, g, smem, a, rlist.p, rcount.p, slab_off, skew);
  }
  merge_lists("merge_topk", nq, cand_d.p, cand_id.p, cand_cnt.p, np, k, (size_t)k, (size_t)k, (size_t)np * k, (size_t)1,
              (size_t)np, out_ids, out_dists, out_counts);
}

// Row-sharded index (SURVEY 8e search (ii)): every rank has searched its own shard; the per-rank top-k lists
// are exchanged in ONE collective and merged on every rank by (_distance, _rowid), the order of the
// reference's final SortExec (rust/lance/src/dataset/scanner.rs:3450-3466).  ids / dists / counts: this
// rank's [nq][k] / [nq] results on the device; outputs likewise.
void merge_sharded_topk(const uint64_t* ids, const float* dists, const uint32_t* counts, uint64_t nq, int k,
                        uint64_t* out_ids, float* out_dists, uint32_t* out_counts) {
  Comm* c = current_comm();
  const int nr = c ? c->nranks : 1;
  const size_t id_bytes = (size_t)nq * k * 8, d_bytes = ((size_t)nq * k * 4 + 7) / 8 * 8, c_bytes = ((size_t)nq * 4 + 7) / 8 * 8;
  const size_t S = id_bytes + d_bytes + c_bytes;
  DevBuf<uint8_t> blob(S), gathered(S * nr);
  LB2_CUDA(cudaMemcpyAsync(blob.p, ids, (size_t)nq * k * 8, cudaMemcpyDeviceToDevice, ctx().stream));
  LB2_CUDA(cudaMemcpyAsync(blob.p + id_bytes, dists, (size_t)nq * k * 4, cudaMemcpyDeviceToDevice, ctx().stream));
  LB2_CUDA(cudaMemcpyAsync(blob.p + id_bytes + d_bytes, counts, (size_t)nq * 4, cudaMemcpyDeviceToDevice, ctx().stream));
  comm_allgather_bytes(blob.p, gathered.p, S);
  merge_lists("merge_sharded_topk", nq, reinterpret_cast<const float*>(gathered.p + id_bytes),
              reinterpret_cast<const uint64_t*>(gathered.p), reinterpret_cast<const uint32_t*>(gathered.p + id_bytes + d_bytes),
              nr, k, S / 4, S / 8, (size_t)k, S / 4, (size_t)1, out_ids, out_dists, out_counts);
  sync_stream();  // the exchange buffers are freed on return
}

__device__ __forceinline__ bool sorted_contains(const uint64_t* __restrict__ a, uint64_t n, uint64_t v) {
  uint64_t lo = 0, hi = n;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo < n && a[lo] == v;
}
__global__ void row_mask_kernel(const uint64_t* __restrict__ row_ids, uint64_t n,
                                const uint64_t* __restrict__ allow, uint64_t n_allow, int has_allow,
                                const uint64_t* __restrict__ block, uint64_t n_block, int has_block,
                                uint32_t* __restrict__ bitmap32) {
  const uint64_t pos = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;  // n padded to 64 by the grid
  bool sel = false;
  if (pos < n) {
    const uint64_t id = row_ids[pos];
    sel = (!has_allow || sorted_contains(allow, n_allow, id)) && !(has_block && sorted_contains(block, n_block, id));
  }
  const unsigned bal = __ballot_sync(0xffffffffu, sel);
  // the bitmap holds ceil(n / 64) u64 words; the last CTA may reach beyond it
  if ((threadIdx.x & 31) == 0 && (pos >> 5) < ((n + 63) / 64) * 2) bitmap32[pos >> 5] = bal;
}
void row_mask_f32(const uint64_t* row_ids, uint64_t n, const uint64_t* allow, uint64_t n_allow, bool has_allow,
                  const uint64_t* block, uint64_t n_block, bool has_block, uint64_t* bitmap) {
  const uint64_t padded = (n + 63) / 64 * 64;
  if (padded == 0) return;
  LB2_LAUNCH("row_mask", row_mask_kernel, cdiv(padded, 256), 256, 0, row_ids, n, allow, n_allow,
             has_allow ? 1 : 0, block, n_block, has_block ? 1 : 0, reinterpret_cast<uint32_t*>(bitmap));
}

void ivfflat_search_f32(const float* centroids, int K, int d, int metric, const uint64_t* part_offsets,
                        const void* vectors, int vdt, const uint64_t* row_ids, const float* queries, uint64_t nq,
                        int k, int nprobes, uint64_t* out_ids, float* out_dists, uint32_t* out_counts,
                        const ScanFilter& flt) {
  if (nq == 0 || k == 0) return;
  if (k > 1024) fail(LB2_UNSUPPORTED, "k (incl. refine factor) > 1024 is not implemented");
  const int np = nprobes < K ? nprobes : K;
  // partitions are found with L2 on the (normalised) vectors for cosine (ivf.rs:149-185)
  const int cmetric = metric == METRIC_DOT ? METRIC_DOT : METRIC_L2;
  DevBuf<uint32_t> pids((size_t)nq * np), cand_cnt((size_t)nq * np);
  DevBuf<float> pd((size_t)nq * np), cand_d((size_t)nq * np * k);
  DevBuf<uint64_t> cand_id((size_t)nq * np * k);
  find_partitions_f32(centroids, K, d, cmetric, queries, nq, np, pids.p, pd.p);
  const size_t smem = sizeof(float) * ((size_t)d + SCAN_CHUNK + 4 * (size_t)(k + 1));
  if (smem > ctx().smem_optin) fail(LB2_UNSUPPORTED, "dimension %d too large for the flat scan", d);
  for (uint64_t q0 = 0; q0 < nq; q0 += 32768) {
    const uint64_t qn = std::min<uint64_t>(32768, nq - q0);
    dim3 g(np, (unsigned)qn);
#define LB2_FLAT_T(MET, TT)                                                                             \
    {                                                                                                   \
      set_smem((ivfflat_scan_kernel<MET, TT>), smem);                                                   \
      LB2_LAUNCH("flat_scan", (ivfflat_scan_kernel<MET, TT>), g, 256, smem, queries + q0 * d, d,         \
                 pids.p + q0 * np, np, part_offsets, reinterpret_cast<const TT*>(vectors), row_ids, k,   \
                 cand_d.p + q0 * np * k, cand_id.p + q0 * np * k, cand_cnt.p + q0 * np, flt);            \
    }
#define LB2_FLAT(MET)                                                                                   \
    {                                                                                                   \
      if (vdt == LB2_F16) LB2_FLAT_T(MET, __half)                                                       \
      else if (vdt == LB2_BF16) LB2_FLAT_T(MET, __nv_bfloat16)                                          \
      else LB2_FLAT_T(MET, float)                                                                       \
    }
    if (metric == METRIC_DOT) LB2_FLAT(METRIC_DOT)
    else if (metric == METRIC_COSINE) LB2_FLAT(METRIC_COSINE)
    else LB2_FLAT(METRIC_L2)
#undef LB2_FLAT_T
#undef LB2_FLAT
  }
  merge_lists("merge_topk", nq, cand_d.p, cand_id.p, cand_cnt.p, np, k, (size_t)k, (size_t)k, (size_t)np * k, (size_t)1,
              (size_t)np, out_ids, out_dists, out_counts);
}

// ------------------------------------------------------------------------------------------------
// refine: exact distances of k' = k * refine_factor candidates from the raw vectors, then the k
// best by (distance, row id)  (scanner.rs:2884-2905, flat.rs:95-148)
// ------------------------------------------------------------------------------------------------
template <int METRIC, class T>
__global__ void __launch_bounds__(256)
refine_kernel(const float* __restrict__ queries, int d, const T* __restrict__ vectors,
              uint64_t num_vectors, const uint64_t* __restrict__ cand_id, const uint32_t* __restrict__ cand_cnt,
              int kc, int k, uint64_t* __restrict__ out_id, float* __restrict__ out_d,
              uint32_t* __restrict__ out_cnt, int has_lower, float lower, int has_upper, float upper) {
  extern __shared__ float smem[];
  float* qs = smem;       // [d]
  float* cd = qs + d;     // [kc]
  __shared__ int32_t s_key[8];
  __shared__ uint64_t s_tie[8];
  __shared__ int s_tid[9];
  __shared__ int32_t prev_key;
  __shared__ uint64_t prev_id;
  __shared__ float s_qnorm;
  const size_t qi = blockIdx.x;
  const int tid = threadIdx.x, l = tid & 15;
  const unsigned hmask = 0xffffu << (16 * ((tid >> 4) & 1));
  const uint32_t cnt = min(cand_cnt[qi], (uint32_t)kc);
  for (int t = tid; t < d; t += 256) qs[t] = queries[qi * d + t];
  __syncthreads();
  if (METRIC == METRIC_COSINE && tid < 32) {
    float a = 0.0f;
    for (int e = (tid & 15); e < d; e += 16) a = fmaf(qs[e], qs[e], a);
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o, 16);
    if (tid == 0) s_qnorm = sqrtf(a);
  }
  __syncthreads();
  const float qn = METRIC == METRIC_COSINE ? s_qnorm : 0.0f;
  const uint64_t* ids = cand_id + qi * kc;
  for (uint32_t c = tid >> 4; c < cnt; c += 16) {
    const uint64_t id = ids[c];
    float dist = __int_as_float(0x7fc00000);
    if (id < num_vectors) dist = flat_row_distance<METRIC, T>(qs, vectors + id * (uint64_t)d, d, l, hmask, qn);
    if (l == 0) cd[c] = dist;
  }
  __syncthreads();
  bool first = true;
  uint32_t r = 0;
  const uint32_t rounds = cnt < (uint32_t)k ? cnt : (uint32_t)k;
  auto passes = [&](float dv) {  // LanceFilterExec(_distance >= lower AND _distance < upper): SQL compares
    return (!has_lower || dv >= lower) && (!has_upper || dv < upper);
  };
  for (; r < rounds; ++r) {
    int32_t bk = 0;
    uint64_t bi = 0;
    uint32_t bslot = 0;
    bool has = false;
    const int32_t pk = first ? 0 : prev_key;
    const uint64_t pid = first ? 0 : prev_id;
    for (uint32_t c = tid; c < cnt; c += 256) {
      if (!passes(cd[c])) continue;
      const int32_t key = total_order_key(cd[c]);
      const uint64_t id = ids[c];
      if (!first && !ki_less(pk, pid, key, id)) continue;
      if (!has || ki_less(key, id, bk, bi)) { bk = key; bi = id; bslot = c; has = true; }
    }
    const int w = block_argmin<256>(has, bk, bi, s_key, s_tie, s_tid);
    if (w < 0) break;
    if (tid == w) {
      prev_key = bk;
      prev_id = bi;
      out_id[qi * k + r] = bi;
      out_d[qi * k + r] = cd[bslot];
    }
    __syncthreads();
    first = false;
  }
  for (uint32_t e = r + tid; e < (uint32_t)k; e += 256) {
    out_id[qi * k + e] = ~0ull;
    out_d[qi * k + e] = __int_as_float(0x7f800000);
  }
  if (tid == 0 && out_cnt) out_cnt[qi] = r;
}

void refine_f32(const float* queries, uint64_t nq, int d, int metric, const void* vectors, int vdt,
                uint64_t num_vectors, const uint64_t* cand_id, const uint32_t* cand_cnt, int kc, int k,
                uint64_t* out_id, float* out_d, uint32_t* out_cnt, int has_lower, float lower, int has_upper,
                float upper) {
  if (nq == 0) return;
  const size_t smem = sizeof(float) * ((size_t)d + kc);
#define LB2_REF_T(MET, TT)                                                                           \
  {                                                                                                   \
    set_smem((refine_kernel<MET, TT>), smem);                                                         \
    LB2_LAUNCH("refine", (refine_kernel<MET, TT>), (unsigned)nq, 256, smem, queries, d,                \
               reinterpret_cast<const TT*>(vectors), num_vectors, cand_id, cand_cnt, kc, k, out_id,    \
               out_d, out_cnt, has_lower, lower, has_upper, upper);                                    \
  }
#define LB2_REF(MET)                                                                                  \
  {                                                                                                   \
    if (vdt == LB2_F16) LB2_REF_T(MET, __half)                                                        \
    else if (vdt == LB2_BF16) LB2_REF_T(MET, __nv_bfloat16)                                           \
    else if (vdt == LB2_U8) LB2_REF_T(MET, uint8_t)                                                   \
    else LB2_REF_T(MET, float)                                                                        \
  }
  if (metric == METRIC_DOT) LB2_REF(METRIC_DOT)
  else if (metric == METRIC_COSINE) LB2_REF(METRIC_COSINE)
  else LB2_REF(METRIC_L2)
#undef LB2_REF_T
#undef LB2_REF
}

void build_lut_f32(const float* codebook, int M, int nbits, int d, int metric, const float* query,
                   float* lut) {
  const int ncode = 1 << nbits, ds = d / M;
  if (metric == METRIC_DOT)
    LB2_LAUNCH("build_lut", build_lut_kernel<METRIC_DOT>, cdiv((uint64_t)M * ncode, 128), 128, 0,
               codebook, M, ncode, ds, query, lut);
  else
    LB2_LAUNCH("build_lut", build_lut_kernel<METRIC_L2>, cdiv((uint64_t)M * ncode, 128), 128, 0,
               codebook, M, ncode, ds, query, lut);
}

void pq_scan_transposed_f32(const float* lut, int M, int metric, const uint8_t* codes_t, uint64_t n,
                            float* out) {
  if (n == 0) return;
  const size_t smem = sizeof(float) * (size_t)M * 256;
  if (smem > ctx().smem_optin) fail(LB2_UNSUPPORTED, "LUT of %zu bytes exceeds shared memory", smem);
  set_smem(pq_scan_transposed_kernel, smem);
  LB2_LAUNCH("pq_scan_transposed", pq_scan_transposed_kernel, cdiv(n, 256), 256, smem, lut, M,
             codes_t, n, metric == METRIC_DOT ? 1 : 0, out);
}

// ------------------------------------------------------------------------------------------------
// a19  4-bit PQ scan: compute_pq_distance_4bit (pq/distance.rs:147-242).  lut = M x 16 f32, codes_t =
// transposed packed codes [M/2][n] (low nibble = sub-vector 2i, high nibble = 2i+1).
//   rows [0, flat_num) and the last n % 16 rows: exact f32, two adds per byte in byte order;
//   the others: saturating u8 sum of the table quantised with qmin = min(table), qmax = max of the
//   flat rows (total order), then q * ((qmax - qmin) / 255) + qmin.
// ------------------------------------------------------------------------------------------------
__global__ void pq4_flat_kernel(const float* __restrict__ lut, int nb, const uint8_t* __restrict__ codes_t,
                                uint64_t n, uint64_t off, uint64_t len, float* __restrict__ out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= len) return;
  const uint64_t j = off + t;
  float dist = 0.0f;
  for (int i = 0; i < nb; ++i) {
    const uint8_t c = codes_t[(size_t)i * n + j];
    dist = f_add(dist, lut[(2 * i) * 16 + (c & 0xF)]);
    dist = f_add(dist, lut[(2 * i + 1) * 16 + (c >> 4)]);
  }
  out[j] = dist;
}
// one block: qmax over the flat rows (total order), qmin over the table (f32::min ignores NaN), the u8 table
__global__ void pq4_quantize_kernel(const float* __restrict__ lut, int M, const float* __restrict__ flat,
                                    uint64_t flat_num, uint8_t* __restrict__ qt, float* __restrict__ params) {
  __shared__ int32_t s_max[256];
  __shared__ float s_min[256];
  const int tid = threadIdx.x;
  int32_t mx = (int32_t)0x80000000;
  for (uint64_t j = tid; j < flat_num; j += 256) mx = max(mx, total_order_key(flat[j]));
  float mn = __int_as_float(0x7f800000);
  for (int i = tid; i < M * 16; i += 256) mn = fminf(mn, lut[i]);
  s_max[tid] = mx;
  s_min[tid] = mn;
  __syncthreads();
  for (int o = 128; o >= 1; o >>= 1) {
    if (tid < o) {
      s_max[tid] = max(s_max[tid], s_max[tid + o]);
      s_min[tid] = fminf(s_min[tid], s_min[tid + o]);
    }
    __syncthreads();
  }
  const float qmax = key_to_float(s_max[0]), qmin = s_min[0];
  const float factor = __fdiv_rn(255.0f, __fsub_rn(qmax, qmin));
  for (int i = tid; i < M * 16; i += 256) {
    const float v = roundf(__fmul_rn(__fsub_rn(lut[i], qmin), factor));  // f32::round: half away from zero
    qt[i] = (v != v) ? 0 : v <= 0.0f ? 0 : v >= 255.0f ? 255 : (uint8_t)v;  // `as u8`: saturating, NaN -> 0
  }
  if (tid == 0) {
    params[0] = qmin;
    params[1] = __fdiv_rn(__fsub_rn(qmax, qmin), 255.0f);
  }
}
__global__ void pq4_quant_scan_kernel(const uint8_t* __restrict__ qt, int nb, const uint8_t* __restrict__ codes_t,
                                      uint64_t n, uint64_t begin, uint64_t end,
                                      const float* __restrict__ params, float* __restrict__ out) {
  extern __shared__ uint8_t s_qt[];
  for (int i = threadIdx.x; i < nb * 32; i += blockDim.x) s_qt[i] = qt[i];
  __syncthreads();
  const uint64_t j = begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= end) return;
  uint32_t q = 0;  // saturating u8 adds of non-negative terms == min(255, sum)
  for (int i = 0; i < nb; ++i) {
    const uint8_t c = codes_t[(size_t)i * n + j];
    q += s_qt[(2 * i) * 16 + (c & 0xF)];
    q += s_qt[(2 * i + 1) * 16 + (c >> 4)];
  }
  q = min(q, 255u);
  out[j] = __fadd_rn(__fmul_rn((float)q, params[1]), params[0]);
}
__global__ void sub_scalar_kernel(float* __restrict__ v, uint64_t n, float s) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) v[j] = __fsub_rn(v[j], s);
}
void pq_scan_4bit_f32(const float* lut, int M, int metric, const uint8_t* codes_t, uint64_t n, uint64_t k_hint,
                      float* out) {
  if (n == 0) return;
  const int nb = M / 2;
  k_hint = std::min<uint64_t>(k_hint, n);
  const uint64_t flat_num = std::min<uint64_t>(std::max<uint64_t>(200, k_hint), n);  // FLAT_NUM_4BIT_PQ = 200
  const uint64_t rem = n % 16;
  LB2_LAUNCH("pq4_flat", pq4_flat_kernel, cdiv(flat_num, 256), 256, 0, lut, nb, codes_t, n, (uint64_t)0, flat_num, out);
  DevBuf<uint8_t> qt((size_t)M * 16);
  DevBuf<float> params(2);
  LB2_LAUNCH("pq4_quantize", pq4_quantize_kernel, 1, 256, 0, lut, M, (const float*)out, flat_num, qt.p, params.p);
  if (n - rem > flat_num)
    LB2_LAUNCH("pq4_scan", pq4_quant_scan_kernel, cdiv(n - rem - flat_num, 256), 256, (size_t)M * 16, qt.p, nb,
               codes_t, n, flat_num, n - rem, params.p, out);
  if (rem > 0) {
    const uint64_t off = std::max(n - rem, flat_num);
    if (n > off) LB2_LAUNCH("pq4_flat", pq4_flat_kernel, cdiv(n - off, 256), 256, 0, lut, nb, codes_t, n, off, n - off, out);
  }
  if (metric == METRIC_DOT)
    LB2_LAUNCH("pq4_dot_fix", sub_scalar_kernel, cdiv(n, 256), 256, 0, out, n, (float)M - 1.0f);
  sync_stream();  // qt / params are freed on return
}

// two 4-bit codes per byte: (v[1] << 4) | v[0]  (pq.rs:168-173)
__global__ void pack_nibbles_kernel(const uint8_t* __restrict__ codes, uint64_t total_bytes, uint8_t* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total_bytes) out[i] = (uint8_t)((codes[2 * i + 1] << 4) | (codes[2 * i] & 0xF));
}
void pack_nibbles(const uint8_t* codes, uint64_t n, int M, uint8_t* out) {
  const uint64_t total = n * (uint64_t)(M / 2);
  if (total) LB2_LAUNCH("pack_nibbles", pack_nibbles_kernel, cdiv(total, 256), 256, 0, codes, total, out);
}

void flat_topk_f32(const float* dists, const uint64_t* row_ids, uint64_t n, int k, const ScanFilter& flt,
                   uint64_t* out_id, float* out_d, uint32_t* out_cnt) {
  if (k > 1024) fail(LB2_UNSUPPORTED, "k > 1024 is not implemented");
  if (k < 16)
    LB2_LAUNCH("flat_topk", flat_topk_kernel<16>, 1, 256, 0, dists, row_ids, n, k, flt, out_id, out_d, out_cnt);
  else if (k < 128)
    LB2_LAUNCH("flat_topk", flat_topk_kernel<128>, 1, 256, 0, dists, row_ids, n, k, flt, out_id, out_d, out_cnt);
  else
    LB2_LAUNCH("flat_topk", flat_topk_kernel<1025>, 1, 256, 0, dists, row_ids, n, k, flt, out_id, out_d, out_cnt);
}

}  // namespace lb2
