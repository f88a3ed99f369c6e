// tc_common.cuh -- PTX wrappers shared by the tcgen05 kernels (tc_assign.cu, tc_pq.cu):
// mbarrier, TMA (cp.async.bulk.tensor), UMMA descriptors, tcgen05.mma/commit/ld.
// Bit layouts follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor).
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdio.h>

#include "common.cuh"

namespace lb2 {
namespace tc {

constexpr int TM = 128;                  // rows per tile (UMMA M)
constexpr int TN = 256;                  // centroids per tile (UMMA N)
constexpr int KC = 32;                   // f32 per 128-byte swizzle row
constexpr int A_STAGE_BYTES = TM * 128;  // 16 KB
constexpr int B_CHUNK_BYTES = TN * 128;  // 32 KB
constexpr int MAX_STAGES = 5;
constexpr int NUM_THREADS = 384;         // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warps 4-7 / 8-11: two epilogue groups
                                         // (group g drains TMEM buffer g, so a tcgen05.ld stall of one group is hidden by the other)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Watchdog: a wait that spins for more than ~4 s of SM clocks is a protocol bug, never a slow kernel.
// It reports which barrier starved and traps (the launch fails with an error instead of hanging the
// GPU).  With -DLB2_TC_WATCHDOG_SOFT the first starved wait only raises a flag that makes every
// later wait fall through, so the kernel ends and the printf buffer reaches the host.
#ifdef LB2_TC_WATCHDOG_SOFT
static __device__ int g_wd_abort = 0;
#endif
static __device__ __noinline__ void mbar_watchdog_report(uint32_t bar, uint32_t parity) {
  printf("[lb2 watchdog] block %d thread %d: mbarrier smem+0x%x parity %u starved\n", (int)blockIdx.x,
         (int)threadIdx.x, bar, parity);
#ifdef LB2_TC_WATCHDOG_SOFT
  g_wd_abort = 1;
  __threadfence();
#else
  __trap();
#endif
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// slow path of a wait (kept out of line so the hot loops stay small): called every few thousand failed
// polls; the first call records the start time, later calls compare against it
static __device__ __noinline__ bool mbar_watchdog_tick(uint32_t bar, uint32_t parity, long long* t0) {
#ifdef LB2_TC_WATCHDOG_SOFT
  if (*(volatile int*)&g_wd_abort) return true;
#endif
  const long long now = clock64();
  if (*t0 == 0) { *t0 = now; return false; }
  if (now - *t0 > 8000000000ll) { mbar_watchdog_report(bar, parity); return true; }
  return false;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  long long t0 = 0;
  for (uint32_t spins = 1;; ++spins) {
    if (mbar_try(bar, parity)) return;
    if ((spins & 0x3FFFu) == 0 && mbar_watchdog_tick(bar, parity, &t0)) return;
  }
}
// same, for the single-thread producer / issuer roles: back off between polls so that the spin
// loop does not steal issue slots from the epilogue warps sharing the scheduler
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  long long t0 = 0;
  for (uint32_t spins = 1;; ++spins) {
    if (mbar_try(bar, parity)) return;
    __nanosleep(40);
    if ((spins & 0x3FFu) == 0 && mbar_watchdog_tick(bar, parity, &t0)) return;
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp:SmemDescriptor):
// start>>4 [0,14), LBO>>4 [16,30) (unused for swizzled K-major, 1), SBO>>4 [32,46) = 1024 B (8 rows),
// version=1 [46,48), layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::tf32 (InstrDescriptor): c_format F32=1 [4,6), a/b format TF32=2 [7,10)/[10,13),
// K-major A and B (bits 15,16 = 0), N>>3 [17,23), M>>4 [24,29)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(IDESC), "r"(accumulate)
      : "memory");
}
// 16-bit operands (kind::f16): a/b format F16 = 0, BF16 = 1; one instruction covers K = 16 elements (32 B of a
// 128-byte swizzled row, like K = 8 for TF32), f32 accumulation, twice the TF32 rate
constexpr uint32_t IDESC_F16 = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
constexpr uint32_t IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
// OPK: 0 = f32 operands as TF32, 1 = f16, 2 = bf16
template <int OPK>
__device__ __forceinline__ void umma_op(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  if (OPK == 0) {
    umma_tf32(d_tmem, a_desc, b_desc, accumulate);
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(OPK == 1 ? IDESC_F16 : IDESC_BF16), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}


// tcgen05.wait::ld that also names the destination registers of the load it completes, so the
// compiler cannot schedule their consumers above the wait
__device__ __forceinline__ void tmem_wait_ld(uint32_t* v) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

// insert g into the sorted triple (m1 >= m2 >= m3)
__device__ __forceinline__ void top3_insert(float g, float& m1, float& m2, float& m3) {
  const float t1 = fminf(m1, g);
  m1 = fmaxf(m1, g);
  const float t2 = fminf(m2, t1);
  m2 = fmaxf(m2, t1);
  m3 = fmaxf(m3, t2);
}
// ---- top-3 of a 256-column accumulator row by TOURNAMENT ------------------------------------------
// The epilogue is bound by the half-rate ALU pipe (FMNMX / FMNMX3 / PRMT), so what counts is min/max
// instructions per column.  Pair the values: the winners (max) go on, and of the losers (min) only the
// LARGEST can be among the row's top 3 -- a loser is beaten by its own partner, so two losers in the top
// 3 would need four distinct values ahead of the smaller one.  (All values are distinct: each carries
// its column index in the low mantissa byte.)  Applying this at every level,
//     top3(row) = top3( top3(last-level winners)  U  { largest loser of each level } ),
// i.e. per level one running FMNMX3 maximum over the losers plus one exact top-3 tracker fed by a single
// value per 32-column chunk: 83 min/max instructions per chunk instead of 128.
constexpr int TOUR_LEVELS = 5;

// one 32-column chunk.  C0 is a compile-time constant so that the index OR-ed into the mantissa is an
// immediate.  T = exact top-3 of the chunk winners so far, L[j] = largest level-j loser so far.
template <int C0>
__device__ __forceinline__ void tour_chunk(const uint32_t* v, const float* cn, float* T, float* L) {
  float g[32];
#pragma unroll
  for (int u = 0; u < 32; ++u) {
    const float f = __uint_as_float(v[u]) + cn[C0 + u];
    // replace the low mantissa byte by the column index: one PRMT (byte permute) with an immediate
    g[u] = __uint_as_float(__byte_perm(__float_as_uint(f), (uint32_t)(C0 + u), 0x3214));
  }
#pragma unroll
  for (int lvl = 0, cnt = 32; lvl < TOUR_LEVELS; ++lvl, cnt >>= 1) {
    // cnt values in g[0..cnt) -> cnt/2 winners in g[0..cnt/2), losers folded into L[lvl]
    float lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < cnt / 2) {
        const float x = g[2 * i], y = g[2 * i + 1];
        g[i] = fmaxf(x, y);
        lo[i] = fminf(x, y);
      }
    }
    if (cnt / 2 >= 2) {
#pragma unroll
      for (int i = 0; i < 16; i += 2)
        if (i < cnt / 2) L[lvl] = fmaxf(fmaxf(L[lvl], lo[i]), lo[i + 1]);  // -> FMNMX3
    } else {
      L[lvl] = fmaxf(L[lvl], lo[0]);
    }
  }
  top3_insert(g[0], T[0], T[1], T[2]);
}

// top-3 of a 256-column accumulator row; the TMEM loads are software pipelined (the load of chunk
// c+1 is in flight while chunk c is reduced); fully unrolled over the 8 chunks
__device__ __forceinline__ void top3_row256(uint32_t taddr, const float* cn, float& m1, float& m2, float& m3) {
  uint32_t va[32], vb[32];
  const float ninf = __int_as_float(0xff800000);
  float T[3] = {ninf, ninf, ninf};
  float L[TOUR_LEVELS];
#pragma unroll
  for (int j = 0; j < TOUR_LEVELS; ++j) L[j] = ninf;
  tmem_ld32(taddr, va);
#define LB2_TOP3_STEP(C)                                   \
  tmem_wait_ld(va);                                        \
  tmem_ld32(taddr + (C) + 32, vb);                         \
  tour_chunk<(C)>(va, cn, T, L);                           \
  tmem_wait_ld(vb);                                        \
  if ((C) + 64 < TN) tmem_ld32(taddr + (C) + 64, va);      \
  tour_chunk<(C) + 32>(vb, cn, T, L);
  LB2_TOP3_STEP(0)
  LB2_TOP3_STEP(64)
  LB2_TOP3_STEP(128)
  LB2_TOP3_STEP(192)
#undef LB2_TOP3_STEP
  m1 = T[0]; m2 = T[1]; m3 = T[2];
#pragma unroll
  for (int j = 0; j < TOUR_LEVELS; ++j) top3_insert(L[j], m1, m2, m3);
}

// ---- candidate pass: every column of a 256-column accumulator row whose score reaches thr -------------------
// (rows that the top-3 passes could not settle; hits are rare, so the common path is one add + half a 3-input max
// per column and a single compare per 32-column chunk)
constexpr int CAND_SLOTS = 16;
template <int C0>
__device__ __forceinline__ void cand_chunk(const uint32_t* v, const float* cn, float thr, uint32_t col_base,
                                           uint32_t* cnt, uint32_t* cand) {
  float s[32];
#pragma unroll
  for (int u = 0; u < 32; ++u) s[u] = __uint_as_float(v[u]) + cn[C0 + u];
  float m = s[0];
#pragma unroll
  for (int u = 1; u < 31; u += 2) m = fmaxf(fmaxf(m, s[u]), s[u + 1]);
  m = fmaxf(m, s[31]);
  if (m >= thr) {  // thr = +inf for rows beyond the list; NaN never compares true
#pragma unroll
    for (int u = 0; u < 32; ++u) {
      if (s[u] >= thr) {
        const uint32_t slot = atomicAdd(cnt, 1u);
        if (slot < (uint32_t)CAND_SLOTS) cand[slot] = col_base + (uint32_t)(C0 + u);
      }
    }
  }
}
__device__ __forceinline__ void cand_row256(uint32_t taddr, const float* cn, float thr, uint32_t col_base,
                                            uint32_t* cnt, uint32_t* cand) {
  uint32_t va[32], vb[32];
  tmem_ld32(taddr, va);
#define LB2_CAND_STEP(C)                                   \
  tmem_wait_ld(va);                                        \
  tmem_ld32(taddr + (C) + 32, vb);                         \
  cand_chunk<(C)>(va, cn, thr, col_base, cnt, cand);       \
  tmem_wait_ld(vb);                                        \
  if ((C) + 64 < TN) tmem_ld32(taddr + (C) + 64, va);      \
  cand_chunk<(C) + 32>(vb, cn, thr, col_base, cnt, cand);
  LB2_CAND_STEP(0)
  LB2_CAND_STEP(64)
  LB2_CAND_STEP(128)
  LB2_CAND_STEP(192)
#undef LB2_CAND_STEP
}

}  // namespace tc

// host: 2-D f32 tensor map, box = [32 floats (128 B, SWIZZLE_128B)] x box_rows
CUtensorMap make_map_2d(const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows);
// same for f16 / bf16 rows: box = [64 elements (128 B)] x box_rows
CUtensorMap make_map_2d_16(const void* base, bool bf16, uint64_t rows, uint64_t cols, uint32_t box_rows);

}  // namespace lb2
