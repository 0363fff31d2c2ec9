// Pieces shared by the tcgen05 scoring kernels (score_tc.cu: one CTA per tile, score_tc2.cu: CTA pairs).
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace b2k { namespace tc {

constexpr int kBlockM = 128;           // queries per CTA (TMEM lanes)
constexpr int kBlockN = 256;           // DB rows per accumulator (TMEM columns)
constexpr int kBlockK = 64;            // bf16 elements per K-block = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kTmemCols = 512;         // two 128 x 256 fp32 accumulators
constexpr int kListBytes = kList * kBlockM * 8;  // per-thread top-32 lists: scores + rows
constexpr int kThreads = 192;          // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue

// Shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);      // start address  [0,14)
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}

// Instruction descriptor: D=f32, A=B=bf16, both K-major.
constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// Per-thread top-32 lists live in shared memory as [entry][128 threads] (conflict-free across a
// warp); they are addressed with 32-bit shared-window addresses so the accesses are LDS/STS.
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int32_t lds_s32(uint32_t a) {
  int32_t v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_s32(uint32_t a, int32_t v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

constexpr uint32_t kEntryStride = kBlockM * 4;   // bytes between consecutive entries of one thread

// Replace the worst entry (slot min_e) of a per-thread list and rescan for the new worst.
// Returns (new worst slot << 32) | bits(new worst score).
static __device__ __noinline__ uint64_t list_insert(uint32_t s_addr, uint32_t r_addr, int min_e, float sc, int32_t row) {
  sts_f32(s_addr + (uint32_t)min_e * kEntryStride, sc);
  sts_s32(r_addr + (uint32_t)min_e * kEntryStride, row);
  float mn = INFINITY;
  int me = 0;
#pragma unroll
  for (int e = 0; e < kList; ++e) {
    const float se = lds_f32(s_addr + (uint32_t)e * kEntryStride);
    if (se < mn) { mn = se; me = e; }
  }
  return ((uint64_t)(uint32_t)me << 32) | (uint64_t)__float_as_uint(mn);
}

// Drain one 128 x 256 accumulator: this thread owns TMEM lane (= query row) `taddr`'s lane field and
// filters the 256 scores of DB rows row0 .. row0+255 (only the first `valid` are real rows)
// against its admission threshold thr = max(worst listed score if the list is full, floor).
// `floor` is the seeded per-query bound (a row below it cannot be a re-rank candidate).
__device__ __forceinline__ void drain_accumulator(uint32_t taddr, int64_t row0, int valid, uint32_t s_addr,
                                                  uint32_t r_addr, float floor, float& thr, int& min_e) {
#pragma unroll 1
  for (int c = 0; c < kBlockN / 32; ++c) {
    uint32_t v[32];
    ptx::tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
    ptx::tmem_ld_wait();
    float mx = -INFINITY;
    const int nvalid = valid - c * 32;           // columns of this chunk that are real rows
    if (nvalid >= 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j >= nvalid) v[j] = 0xff800000u;     // -inf: never inserted
        mx = fmaxf(mx, __uint_as_float(v[j]));
      }
    }
    if (mx > thr) {
      // rare after warm-up; statically indexed so v[] stays in registers
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float sc = __uint_as_float(v[j]);
        if (sc > thr) {
          const uint64_t r = list_insert(s_addr, r_addr, min_e, sc, (int32_t)(row0 + c * 32 + j));
          thr = fmaxf(__uint_as_float((uint32_t)r), floor);
          min_e = (int)(r >> 32);
        }
      }
    }
  }
}

// First accumulator of a CTA when no admission floor is known yet (unseeded launch, in-kernel seeding):
// filling an empty top-32 list element by element costs ~100 insertions per query, serialised over the
// lanes of the warp (60-100 us at 128 live queries, which stalls the stream).  Instead the 256 columns are
// read twice from TMEM (the accumulator stays there until it is released):
//   pass 1 writes the maximum of each 8-column group straight into list entry (column / 8) — 32 distinct
//          rows, no rescans — and one rescan yields the threshold;
//   pass 2 inserts only the columns above the threshold that are not their group's maximum (a dozen without
//          a floor; next to none when the in-kernel seeding has published the floor between the passes).
// `live` = this thread owns a real query (padded rows keep thr = +inf and write nothing).  Warp-collective:
// every lane of the warp must call both (tcgen05.ld is .sync.aligned).
struct FirstTileCodes {
  uint64_t lo, hi;               // 3-bit arg-max of group (c, g) at bit 12*(c%4) + 3*g; lo: chunks 0..3, hi: 4..7
};

__device__ __forceinline__ FirstTileCodes first_tile_pass1(uint32_t taddr, int64_t row0, int valid, uint32_t s_addr,
                                                          uint32_t r_addr, bool live, float floor, float& thr, int& min_e) {
  FirstTileCodes codes{0ull, 0ull};
#pragma unroll 1
  for (int c = 0; c < kBlockN / 32; ++c) {
    uint32_t v[32];
    ptx::tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
    ptx::tmem_ld_wait();
    const int nvalid = valid - c * 32;
    uint32_t code = 0u;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float m = -INFINITY;
      int a = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x = (g * 8 + i) < nvalid ? __uint_as_float(v[g * 8 + i]) : -INFINITY;
        if (x > m) { m = x; a = i; }
      }
      code |= (uint32_t)a << (3 * g);
      if (live) {
        const uint32_t e = (uint32_t)(c * 4 + g);
        sts_f32(s_addr + e * kEntryStride, m);
        sts_s32(r_addr + e * kEntryStride, m > -INFINITY ? (int32_t)(row0 + c * 32 + g * 8 + a) : -1);
      }
    }
    if (c < 4) codes.lo |= (uint64_t)code << (12 * c);
    else codes.hi |= (uint64_t)code << (12 * (c - 4));
  }
  if (live) {
    float mn = INFINITY;
    int me = 0;
#pragma unroll
    for (int e = 0; e < kList; ++e) {
      const float se = lds_f32(s_addr + (uint32_t)e * kEntryStride);
      if (se < mn) { mn = se; me = e; }
    }
    thr = fmaxf(mn, floor);
    min_e = me;
  }
  return codes;
}

__device__ __forceinline__ void first_tile_pass2(uint32_t taddr, int64_t row0, int valid, uint32_t s_addr, uint32_t r_addr,
                                                 FirstTileCodes codes, float floor, float& thr, int& min_e) {
#pragma unroll 1
  for (int c = 0; c < kBlockN / 32; ++c) {
    uint32_t v[32];
    ptx::tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
    ptx::tmem_ld_wait();
    const int nvalid = valid - c * 32;
    const uint32_t code = (uint32_t)((c < 4 ? codes.lo >> (12 * c) : codes.hi >> (12 * (c - 4))) & 0xfffull);
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const bool is_group_max = ((code >> (3 * (j >> 3))) & 7u) == (uint32_t)(j & 7);
      if (j >= nvalid || is_group_max) v[j] = 0xff800000u;       // -inf: listed already, or not a row
      mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    if (mx > thr) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float sc = __uint_as_float(v[j]);
        if (sc > thr) {
          const uint64_t r = list_insert(s_addr, r_addr, min_e, sc, (int32_t)(row0 + c * 32 + j));
          thr = fmaxf(__uint_as_float((uint32_t)r), floor);
          min_e = (int)(r >> 32);
        }
      }
    }
  }
}

// Epilogue prologue / epilogue shared by both kernels.
__device__ __forceinline__ void list_init(uint32_t s_addr, uint32_t r_addr) {
#pragma unroll
  for (int e = 0; e < kList; ++e) {
    sts_f32(s_addr + (uint32_t)e * kEntryStride, -INFINITY);
    sts_s32(r_addr + (uint32_t)e * kEntryStride, -1);
  }
}
__device__ __forceinline__ void list_store(uint32_t s_addr, uint32_t r_addr, Cand* out) {
  // 32 (score, row) records = 256 contiguous, 256-byte aligned bytes per thread: sixteen 16-byte stores
  uint4* o4 = reinterpret_cast<uint4*>(out);
#pragma unroll 8
  for (int e = 0; e < kList; e += 2) {
    uint4 v;
    v.x = __float_as_uint(lds_f32(s_addr + (uint32_t)e * kEntryStride));
    v.y = (uint32_t)lds_s32(r_addr + (uint32_t)e * kEntryStride);
    v.z = __float_as_uint(lds_f32(s_addr + (uint32_t)(e + 1) * kEntryStride));
    v.w = (uint32_t)lds_s32(r_addr + (uint32_t)(e + 1) * kEntryStride);
    o4[e >> 1] = v;
  }
}

// ---- in-kernel threshold seeding ---------------------------------------------------------------------------
// What the sampling pass + seed_kernel do in two extra launches, without re-reading the sampled tile, for
// grids whose CTAs are all resident at once (one CTA per SM).  Called by the 128 epilogue threads of a CTA
// between the two passes over its first tile: the lists hold the 32 group maxima of that tile; every CTA
// publishes them and arrives at a counter barrier (the spins are bounded and a CTA that gives up just keeps
// its own threshold); CTA c computes the floors of queries c, c + n_ctas, ... from the exchanged lists and
// publishes them; a second barrier, and the second pass over this tile and every later tile admit only rows
// above the floor.  (A non-blocking variant — publish, keep draining, pick the floor up when it shows —
// measured slower: the tiles scored meanwhile insert against an unseeded threshold.)
constexpr int kSeedKeysPerThread = 40;        // 128 epilogue threads x 40 >= 160 lists x 32 entries

// Counter barrier across the CTAs of the grid: 1 = everyone arrived, 0 = gave up.  The launcher only enables
// the in-kernel seeding for grids that fit the device in one wave (occupancy query, api.cu), but co-residency
// can still fail at run time (another stream or process holding SMs, a profiler serialising CTAs): the wait is
// therefore bounded by WALL TIME (kGridBarrierTimeoutNs on %globaltimer), after which the CTA gives up and just
// keeps its own threshold — correctness never depends on the barrier, and the worst case costs a fraction of a
// millisecond per search instead of stalling it (all CTAs arrive within ~10 us in the normal case).
constexpr unsigned long long kGridBarrierTimeoutNs = 200000ull;
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// The first CTA that times out POISONS the barrier (top bit of the counter): CTAs that arrive later — the rest of
// a grid that is trickling onto the SMs another kernel leaves free — see the bit and leave at once instead of
// burning the timeout one after the other (26 ms for a 148-CTA grid before: tests/test_gpu_round2.py).
constexpr unsigned int kGridBarrierPoison = 0x80000000u;
__device__ __forceinline__ void grid_barrier_poison(unsigned int* ctr) { atomicOr(ctr, kGridBarrierPoison); }
__device__ __forceinline__ uint32_t grid_barrier_arrive_wait(unsigned int* ctr, unsigned int n) {
  __threadfence();
  if (atomicAdd(ctr, 1u) & kGridBarrierPoison) return 0u;
  const unsigned long long t0 = global_timer_ns();
  for (;;) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if (v & kGridBarrierPoison) return 0u;
    if (v >= n) return 1u;
    if (global_timer_ns() - t0 > kGridBarrierTimeoutNs) { grid_barrier_poison(ctr); return 0u; }
    __nanosleep(40);
  }
}

// scratch: 12 u32; topk: 4 x 32 u64 of shared memory private to the epilogue warps (named barrier 2).
__device__ __forceinline__ void seed_exchange(int seed_k, const float* __restrict__ seed_eps, float* seed_floor,
                                              unsigned int* grid_bar, Cand* partial, int n_lists, int n_splits, int nq,
                                              int qi, int split, uint32_t s_addr, uint32_t r_addr, int cta_id, int n_ctas,
                                              uint32_t* scratch, uint64_t* topk, int et, int quarter, int lane,
                                              float& floor, float& thr) {
  if (qi < nq) list_store(s_addr, r_addr, partial + ((int64_t)qi * n_lists + split) * kList);
  __threadfence();
  ptx::named_bar_sync(2, 128);
  if (et == 0) {
    scratch[8] = grid_barrier_arrive_wait(grid_bar + 0, (unsigned)n_ctas);
    if (scratch[8] == 0u) grid_barrier_poison(grid_bar + 1);      // nobody waits for this CTA at the second barrier
  }
  ptx::named_bar_sync(2, 128);
  const bool lists_ready = scratch[8] != 0u;
  if (lists_ready) {
    for (int hq = cta_id; hq < nq; hq += n_ctas) {
      // Any k distinct listed rows bound b_k from below, so the selection may drop rows as long as it never
      // invents one: every lane keeps only the BEST of its ~37 entries (no passes over registers), the k
      // best of the 128 lane maxima are then found with shuffles.  Two of the true top-k share a lane about
      // once in three searches; the floor is then the (k+1)-th best instead of the k-th: still a lower
      // bound, imperceptibly weaker.
      const Cand* ql = partial + (int64_t)hq * n_lists * kList;
      const int E = n_splits * kList;
      uint64_t best = 0ull;                                       // (score key << 32 | entry): unique
      long long raw[kSeedKeysPerThread];                          // all loads in flight at once: the lists sit in
#pragma unroll                                                    // other SMs' L2 slices, ~2 us away under load
      for (int u = 0; u < kSeedKeysPerThread; ++u) {
        const int e = (u * 4 + quarter) * 32 + lane;
        raw[u] = e < E ? __ldcg(reinterpret_cast<const long long*>(ql + e)) : (long long)0xffffffff00000000ull;
      }
#pragma unroll
      for (int u = 0; u < kSeedKeysPerThread; ++u) {
        const int e = (u * 4 + quarter) * 32 + lane;
        const int32_t row = (int32_t)(raw[u] >> 32);
        const uint32_t fk = row < 0 ? 0u : float_key(__int_as_float((int)(raw[u] & 0xffffffffll)));
        const uint64_t kk = fk ? (((uint64_t)fk << 32) | (uint32_t)e) : 0ull;
        best = kk > best ? kk : best;
      }
      uint64_t prev = ~0ull;
      for (int j = 0; j < seed_k; ++j) {
        uint64_t m = best < prev ? best : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const uint64_t x = __shfl_xor_sync(0xffffffffu, m, o); m = x > m ? x : m; }
        if (lane == 0) topk[quarter * kList + j] = m;
        prev = m;
      }
      ptx::named_bar_sync(2, 128);
      if (quarter == 0) {
        uint64_t c4[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) c4[w] = lane < seed_k ? topk[w * kList + lane] : 0ull;
        uint64_t prev2 = ~0ull, bkey = 0ull;
        for (int j = 0; j < seed_k; ++j) {
          uint64_t m = 0ull;
#pragma unroll
          for (int w = 0; w < 4; ++w) if (c4[w] < prev2 && c4[w] > m) m = c4[w];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) { const uint64_t x = __shfl_xor_sync(0xffffffffu, m, o); m = x > m ? x : m; }
          prev2 = m;
          bkey = m;
        }
        const uint32_t bk = (uint32_t)(bkey >> 32);               // 0: fewer than k rows listed, no floor
        if (lane == 0 && bk != 0u) seed_floor[hq] = nextafterf(key_minus_2eps(bk, seed_eps[hq]), -INFINITY);
      }
      ptx::named_bar_sync(2, 128);                                // topk[] is rewritten for the next query
    }
  }
  __threadfence();
  ptx::named_bar_sync(2, 128);
  if (et == 0) scratch[9] = lists_ready ? grid_barrier_arrive_wait(grid_bar + 1, (unsigned)n_ctas) : 0u;
  ptx::named_bar_sync(2, 128);
  if (scratch[9] != 0u && qi < nq) {
    floor = fmaxf(floor, __ldcg(seed_floor + qi));
    thr = fmaxf(thr, floor);
  }
}

int encode_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);

}}  // namespace b2k::tc
