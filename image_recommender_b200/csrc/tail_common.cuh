// Device-side pieces shared by the tail kernels: K-select / K-rerank / K-finalize as separate launches
// (select.cu: k > 32 and the A/B path), the fused per-query tail kernel (select.cu: tail_kernel) and the
// deferred paths that finish a query inside K-collect / K-exact (scan.cu).
#pragma once
#include "common.cuh"
#include "kernels.h"

#ifndef B2K_PHASE
#define B2K_PHASE(i) do { } while (0)      // debug build of select.cu only: phase stamps (scripts/exp_phase.py)
#endif

namespace b2k {

namespace {

constexpr int kSelThreads = 256;

__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint64_t u = __shfl_xor_sync(0xffffffffu, v, o);
    v = u > v ? u : v;
  }
  return v;
}

// Block-cooperative top-k of n distinct non-zero u64 keys in shared memory (0 = empty slot):
// every warp extracts the k best of its interleaved share with warp shuffles only (no block
// barrier inside the loop), then warp 0 merges the nw*k survivors.  out[0..k) = the k largest keys,
// descending, zero padded.  wtop: nw*32 slots of scratch.  All threads of the block must call.
__device__ __forceinline__ void block_topk_u64_rounds(const uint64_t* keys, int n, int k, uint64_t* wtop, uint64_t* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint64_t prev = ~0ull;
  for (int j = 0; j < k; ++j) {
    uint64_t m = 0ull;
    for (int e = warp * 32 + lane; e < n; e += nw * 32) {
      const uint64_t v = keys[e];
      if (v < prev && v > m) m = v;
    }
    m = warp_max_u64(m);
    if (lane == 0) wtop[warp * 32 + j] = m;
    prev = m;
  }
  __syncthreads();
  if (warp == 0) {
    uint64_t prev2 = ~0ull;
    for (int j = 0; j < k; ++j) {
      uint64_t m = 0ull;
      for (int e = lane; e < nw * k; e += 32) {
        const uint64_t v = wtop[(e / k) * 32 + (e % k)];
        if (v < prev2 && v > m) m = v;
      }
      m = warp_max_u64(m);
      if (lane == 0) out[j] = m;
      prev2 = m;
    }
  }
  __syncthreads();
}

// The same result by rank counting (blockDim.x == 256 = the size of wtop): every thread's best key — the maximum of
// its strided share, 256 DISJOINT sets — has a k-th best L that bounds the k-th best key overall from below; the few
// keys >= L are collected and ranked among themselves.  Two passes of <= 256 broadcast shared-memory reads instead
// of 2 k rounds of warp reductions over all n keys (batch-1 selection: 8.4 -> 4.5 us).  Falls back to the rounds
// when more than 256 keys sit at or above L (fewer than k non-empty shares and many keys).  Keys must be distinct.
__device__ __forceinline__ void block_topk_u64(const uint64_t* keys, int n, int k, uint64_t* wtop, uint64_t* out) {
  __shared__ uint64_t s_l;
  __shared__ unsigned int s_m;
  const int tid = threadIdx.x, T = blockDim.x;
  if (n <= T) {                                   // one key per thread: rank it directly
    const uint64_t mine = tid < n ? keys[tid] : 0ull;
    if (tid < k) out[tid] = 0ull;
    __syncthreads();
    int rank = 0;
    for (int e = 0; e < n; ++e) rank += keys[e] > mine ? 1 : 0;
    if (mine != 0ull && rank < k) out[rank] = mine;
    __syncthreads();
    return;
  }
  uint64_t cmax = 0ull;
  for (int e = tid; e < n; e += T) { const uint64_t v = keys[e]; cmax = v > cmax ? v : cmax; }
  wtop[tid] = cmax;
  if (tid == 0) { s_l = 0ull; s_m = 0u; }
  if (tid < k) out[tid] = 0ull;
  __syncthreads();
  {
    int rank = 0;
    for (int e = 0; e < T; ++e) { const uint64_t v = wtop[e]; rank += (v > cmax || (v == cmax && e < tid)) ? 1 : 0; }
    if (rank == k - 1) s_l = cmax;
  }
  __syncthreads();
  const uint64_t L = s_l;
  for (int e = tid; e < n; e += T) {
    const uint64_t v = keys[e];
    if (v != 0ull && v >= L) {
      const unsigned pos = atomicAdd(&s_m, 1u);
      if (pos < (unsigned)T) wtop[pos] = v;
    }
  }
  __syncthreads();
  const int m = (int)s_m;
  if (m > T) { __syncthreads(); block_topk_u64_rounds(keys, n, k, wtop, out); return; }
  if (tid < m) {
    const uint64_t mine = wtop[tid];
    int rank = 0;
    for (int e = 0; e < m; ++e) rank += wtop[e] > mine ? 1 : 0;
    if (rank < k) out[rank] = mine;
  }
  __syncthreads();
}

// Partial-list entries as distinct u64 keys: (order-preserving score key << 32) | reversed slot.
__device__ __forceinline__ void load_list_keys(const Cand* lst, int E, uint64_t* keys) {
  // four independent 8-byte loads in flight per thread: the lists were just written by the scoring
  // kernel and come from L2 (a dependent one-at-a-time loop costs 6 us per 4736 entries at batch 1)
  const int T = blockDim.x;
  int e = threadIdx.x;
  for (; e + 3 * T < E; e += 4 * T) {
    Cand c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) c[u] = lst[e + u * T];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t fk = c[u].row < 0 ? 0u : float_key(c[u].score);     // NaN scores -> 0: dropped
      keys[e + u * T] = fk ? (((uint64_t)fk << 32) | (uint32_t)(E - 1 - (e + u * T))) : 0ull;
    }
  }
  for (; e < E; e += T) {
    const Cand c = lst[e];
    const uint32_t fk = c.row < 0 ? 0u : float_key(c.score);
    keys[e] = fk ? (((uint64_t)fk << 32) | (uint32_t)(E - 1 - e)) : 0ull;
  }
  __syncthreads();
}

}  // namespace

// ---------------------------------------------------------------------------------------
// Tightening (both branches of select_kernel): the k rows with the best approximate scores are k
// distinct rows, so the smallest of their EXACT scores s' is a lower bound of the exact k-th best
// score, and every row of the exact top-k has b >= s' - eps.  s' >= b_k - eps, so this threshold is
// never looser than b_k - 2 eps and typically one eps tighter: several times fewer rows to re-rank.
// topk[j] = j-th best key (low word = reversed slot); s_exact: k floats of scratch.  Block-wide call.
static __device__ __forceinline__ void tighten_threshold(const SelectArgs& a, int q, const Cand* lst, int E,
                                                  const uint64_t* topk, float* s_exact, float& thr, float& lb) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* qv = a.q + (int64_t)q * a.D;
  // (gathering two rows per warp at once — lane_dot64_x2 — was measured: tighten 21 -> 16 us per CTA at batch 4096,
  // but the extra live registers spill in the register-resident selection around it and give it all back)
  for (int j = warp; j < a.k; j += (kSelThreads >> 5)) {
    const int e = E - 1 - (int)(uint32_t)(topk[j] & 0xffffffffull);
    const float* x = a.db_f32 + (int64_t)lst[e].row * a.D;
    const double p = lane_dot64(qv, x, a.D, lane);
    const float sj = (float)warp_sum_f64(p);
    if (lane == 0) s_exact[j] = sj;
  }
  __syncthreads();
  float smin = INFINITY;
  for (int j = 0; j < a.k; ++j) smin = fminf(smin, s_exact[j]);
  const float t2 = __fsub_rd(smin, a.eps[q]);
  if (t2 > thr) { thr = t2; lb = smin; }      // NaN-safe: keeps the looser bound
}


// ---------------------------------------------------------------------------------------
// K-select for k <= 32, one CTA (kSelThreads) per query q.  skey: n_lists*32 u64 of shared memory; wtop: 8*32 u64;
// top: 32 u64; s_exact32: 32 floats; s_ints: [0] candidate count, [1] saturation overflow flag, [2] pairs handed
// to K-collect (all three zeroed by the caller before the first barrier).  Writes the query's candidate rows,
// cand_count, flags, thr, lb; returns the flags word (0 = certified so far) to every thread.
static __device__ __forceinline__ int select_small_k(const SelectArgs& a, int q, uint64_t* skey, uint64_t* wtop, uint64_t* top,
                                                     float* s_exact32, int* s_ints) {
  const int E = a.n_lists * kList;
  const Cand* lst = a.partial + (int64_t)q * a.list_stride * kList;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  load_list_keys(lst, E, skey);
  int32_t* out_rows = a.cand_rows + (int64_t)q * a.cand_cap;
  // b_k = k-th best approximate score over every list; 0: fewer than k rows listed -> everything
  // listed is a candidate.
  block_topk_u64(skey, E, a.k, wtop, top);
  const uint32_t bk = (uint32_t)(top[a.k - 1] >> 32);
  float thr = bk != 0u ? key_minus_2eps(bk, a.eps[q]) : -INFINITY;
  // lower bound of the exact k-th best score: the k best approximate rows have exact >= b_k - eps
  float lb = bk != 0u ? __fadd_rd(thr, a.eps[q]) : -INFINITY;
  if (bk != 0u && a.db_f32 != nullptr) tighten_threshold(a, q, lst, E, top, s_exact32, thr, lb);

  // candidates + saturation, from the shared-memory keys (rows are fetched for hits only).
  // Warp w owns lists w, w+8, ...; lane j = entry j of the list.
  const uint32_t thr_key = float_key(thr);              // score >= thr  <=>  key >= thr_key
  for (int l = warp; l < a.n_lists; l += (kSelThreads >> 5)) {
    const uint64_t key = skey[l * kList + lane];
    const uint32_t sk = (uint32_t)(key >> 32);
    const bool hit = sk != 0u && sk >= thr_key;
    const unsigned hm = __ballot_sync(0xffffffffu, hit);
    // a list whose 32 slots are all at-risk rows may hide a 33rd: K-collect re-scans that DB split
    // for this query and lists EVERY row at or above the threshold (so nothing is emitted here);
    // only when the pair table is full does the query fall back to the exhaustive scan
    if (hm == 0xffffffffu) {
      if (lane == 0) {
        const int slot = a.sat_pairs ? atomicAdd(a.sat_count, 1) : a.sat_cap;
        if (slot < a.sat_cap) { a.sat_pairs[slot] = make_int2(q, l); atomicAdd(&s_ints[2], 1); }
        else s_ints[1] = 1;
      }
      if (a.sat_pairs) continue;     // on pair-table overflow the query is flagged: its candidates are unused
    }
    if (hm) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_ints[0], __popc(hm));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (hit) {
        const int pos = base + __popc(hm & ((1u << lane) - 1u));
        if (pos < a.cand_cap) out_rows[pos] = lst[l * kList + lane].row;
      }
    }
  }
  __syncthreads();
  int cnt = s_ints[0];
  int flag = 0;
  if (s_ints[1]) flag |= 1;                       // a list may hide at-risk rows
  if (cnt > a.cand_cap) { flag |= 2; cnt = a.cand_cap; }
  if (a.force_exact) flag |= 4;
  if (tid == 0) {
    a.cand_count[q] = cnt;
    a.flags[q] = flag;
    a.thr[q] = thr;
    a.lb[q] = lb;
  }
  return flag;
}

// Register-resident form of select_small_k for n_lists <= kSelRegLists (the usual 148 or fewer): thread t holds
// entries t, t + 256, ... — exactly the (warp = list mod 8, lane = entry) layout the candidate emission walks —
// so the lists are fetched from L2 once, all loads in flight together, and neither the k-th-best search nor the
// emission touches shared memory for them (batch 1, r1 phase timers: 6 us list load + 11 us top-k + 3 us
// emission with the shared-memory form).  Same results bit for bit.
constexpr int kSelRegPer = 20;
constexpr int kSelRegLists = kSelRegPer * kSelThreads / kList;          // 160

static __device__ __forceinline__ int select_small_k_reg(const SelectArgs& a, int q, uint64_t* wtop, uint64_t* top,
                                                         float* s_exact32, int* s_ints) {
  const int E = a.n_lists * kList;
  const Cand* lst = a.partial + (int64_t)q * a.list_stride * kList;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int32_t* out_rows = a.cand_rows + (int64_t)q * a.cand_cap;
  Cand cs[kSelRegPer];
#pragma unroll
  for (int i = 0; i < kSelRegPer; ++i) {
    const int e = tid + i * kSelThreads;
    Cand c; c.score = 0.f; c.row = -1;
    if (e < E) {      // 8-byte L2 load: the lists were just written by the scoring kernel's other SMs
      const long long raw = __ldcg(reinterpret_cast<const long long*>(lst + e));
      c.score = __int_as_float((int)(raw & 0xffffffffll));
      c.row = (int32_t)(raw >> 32);
    }
    cs[i] = c;
  }
  uint64_t key[kSelRegPer];
#pragma unroll
  for (int i = 0; i < kSelRegPer; ++i) {
    const int e = tid + i * kSelThreads;
    const uint32_t fk = cs[i].row < 0 ? 0u : float_key(cs[i].score);     // NaN scores -> 0: dropped
    key[i] = fk ? (((uint64_t)fk << 32) | (uint32_t)(E - 1 - e)) : 0ull;
  }
  B2K_PHASE(6);
  // The k best of the E keys, sorted, into top[0..k).  Two rank-counting passes over at most 256 keys each
  // instead of k rounds of warp reductions over all of them (8.4 us of the batch-1 tail):
  //   1. every thread's best key (the maximum of its <= 20 entries): 256 keys of DISJOINT entry sets, so their k-th
  //      best L bounds the k-th best key overall from below;
  //   2. every key >= L is collected (the k column maxima and the few other entries that beat L) and ranked.
  // More than 256 keys at or above L (only with fewer than k non-empty columns, i.e. L = 0, and many entries): the
  // k-round form below.
  uint64_t* s_u64 = reinterpret_cast<uint64_t*>(s_exact32);       // [0] L, [1] collected count (free until tighten)
  uint64_t cmax = 0ull;
#pragma unroll
  for (int i = 0; i < kSelRegPer; ++i) cmax = key[i] > cmax ? key[i] : cmax;
  wtop[tid] = cmax;
  if (tid == 0) { s_u64[0] = 0ull; s_u64[1] = 0ull; }
  if (tid < kList) top[tid] = 0ull;
  __syncthreads();
  {
    int rank = 0;
    for (int e = 0; e < kSelThreads; ++e) { const uint64_t v = wtop[e]; rank += (v > cmax || (v == cmax && e < tid)) ? 1 : 0; }
    if (rank == a.k - 1) s_u64[0] = cmax;                          // exactly one thread has this rank
  }
  __syncthreads();
  const uint64_t L = s_u64[0];
  __syncthreads();                                                 // wtop is reused for the collected keys
#pragma unroll
  for (int i = 0; i < kSelRegPer; ++i)
    if (key[i] != 0ull && key[i] >= L) {
      const unsigned pos = atomicAdd(reinterpret_cast<unsigned int*>(s_u64 + 1), 1u);
      if (pos < (unsigned)kSelThreads) wtop[pos] = key[i];
    }
  __syncthreads();
  const int m = (int)(unsigned int)s_u64[1];
  if (m <= kSelThreads) {
    if (tid < m) {
      const uint64_t mine = wtop[tid];
      int rank = 0;
      for (int e = 0; e < m; ++e) rank += wtop[e] > mine ? 1 : 0;  // keys are distinct (the slot is part of the key)
      if (rank < a.k) top[rank] = mine;
    }
    __syncthreads();
  } else {
    __syncthreads();
    // k best of this warp's share, then warp 0 merges the 8 x k survivors (as block_topk_u64)
    uint64_t prev = ~0ull;
    for (int j = 0; j < a.k; ++j) {
      uint64_t mx = 0ull;
#pragma unroll
      for (int i = 0; i < kSelRegPer; ++i) if (key[i] < prev && key[i] > mx) mx = key[i];
      mx = warp_max_u64(mx);
      if (lane == 0) wtop[warp * 32 + j] = mx;
      prev = mx;
    }
    __syncthreads();
    if (warp == 0) {
      const int nw = kSelThreads >> 5;
      uint64_t prev2 = ~0ull;
      for (int j = 0; j < a.k; ++j) {
        uint64_t mx = 0ull;
        for (int e = lane; e < nw * a.k; e += 32) {
          const uint64_t v = wtop[(e / a.k) * 32 + (e % a.k)];
          if (v < prev2 && v > mx) mx = v;
        }
        mx = warp_max_u64(mx);
        if (lane == 0) top[j] = mx;
        prev2 = mx;
      }
    }
    __syncthreads();
  }
  const uint32_t bk = (uint32_t)(top[a.k - 1] >> 32);
  float thr = bk != 0u ? key_minus_2eps(bk, a.eps[q]) : -INFINITY;
  float lb = bk != 0u ? __fadd_rd(thr, a.eps[q]) : -INFINITY;
  B2K_PHASE(7);
  if (bk != 0u && a.db_f32 != nullptr) tighten_threshold(a, q, lst, E, top, s_exact32, thr, lb);
  B2K_PHASE(8);
  const uint32_t thr_key = float_key(thr);              // score >= thr  <=>  key >= thr_key
#pragma unroll
  for (int i = 0; i < kSelRegPer; ++i) {
    const int l = warp + i * (kSelThreads >> 5);          // this warp's i-th list; lane = entry
    if (l < a.n_lists) {
      const uint32_t sk = (uint32_t)(key[i] >> 32);
      const bool hit = sk != 0u && sk >= thr_key;
      const unsigned hm = __ballot_sync(0xffffffffu, hit);
      bool emit = hm != 0u;
      if (hm == 0xffffffffu) {                            // saturated list: K-collect's (select_small_k explains)
        if (lane == 0) {
          const int slot = a.sat_pairs ? atomicAdd(a.sat_count, 1) : a.sat_cap;
          if (slot < a.sat_cap) { a.sat_pairs[slot] = make_int2(q, l); atomicAdd(&s_ints[2], 1); }
          else s_ints[1] = 1;
        }
        if (a.sat_pairs) emit = false;
      }
      if (emit) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_ints[0], __popc(hm));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (hit) {
          const int pos = base + __popc(hm & ((1u << lane) - 1u));
          if (pos < a.cand_cap) out_rows[pos] = cs[i].row;
        }
      }
    }
  }
  __syncthreads();
  int cnt = s_ints[0];
  int flag = 0;
  if (s_ints[1]) flag |= 1;
  if (cnt > a.cand_cap) { flag |= 2; cnt = a.cand_cap; }
  if (a.force_exact) flag |= 4;
  if (tid == 0) {
    a.cand_count[q] = cnt;
    a.flags[q] = flag;
    a.thr[q] = thr;
    a.lb[q] = lb;
  }
  return flag;
}

// K-rerank of one query by `n_warps` warps (this warp is number `warp_id` of them): exact score (Spec R) of
// candidates warp_id, warp_id + n_warps, ...  qd: the query widened to fp64 in shared memory, or null.
static __device__ __forceinline__ void rerank_query(const RerankArgs& a, int q, int cnt, int warp_id, int n_warps, int lane,
                                                    const double* qd) {
  const float* __restrict__ qv = a.q + (int64_t)q * a.D;
  const int32_t* rows = a.cand_rows + (int64_t)q * a.cand_cap;       // written by another CTA just before: L2 reads
  float* out = a.cand_ip + (int64_t)q * a.cand_cap;
  if (qd) {
    for (int c = warp_id; c < cnt; c += n_warps) {
      const float* x = a.db_f32 + (int64_t)__ldcg(rows + c) * a.D;
      const float ip = (float)warp_sum_f64(lane_dot64_qd(qd, x, a.D, lane));
      if (lane == 0) out[c] = ip;
    }
    return;
  }
  // latency-bound form: two candidates per warp step (candidates 2w, 2w+1, then + 2 n_warps, ...)
  for (int c = 2 * warp_id; c < cnt; c += 2 * n_warps) {
    const bool two = c + 1 < cnt;
    const float* x0 = a.db_f32 + (int64_t)__ldcg(rows + c) * a.D;
    const float* x1 = two ? a.db_f32 + (int64_t)__ldcg(rows + c + 1) * a.D : x0;
    double p0, p1;
    lane_dot64_x2(qv, x0, x1, a.D, lane, p0, p1);
    const float i0 = (float)warp_sum_f64(p0), i1 = (float)warp_sum_f64(p1);
    if (lane == 0) { out[c] = i0; if (two) out[c + 1] = i1; }
  }
}

// K-finalize for k <= 32, one CTA per query: top-k of the re-ranked candidates by (score desc, row asc), global
// offsets, squared L2.  fkeys: cnt u64 of shared memory.  cand_ip / cand_rows are read through L2 (they were
// written by other CTAs of the cluster, or by this CTA's warps, just before).
static __device__ __forceinline__ void finalize_small_k(const FinalizeArgs& a, int q, int cnt, uint64_t* fkeys, uint64_t* wtop,
                                                        uint64_t* top) {
  const int tid = threadIdx.x;
  for (int c = tid; c < cnt; c += kSelThreads)
    fkeys[c] = cand_key(__ldcg(a.cand_ip + (int64_t)q * a.cand_cap + c), __ldcg(a.cand_rows + (int64_t)q * a.cand_cap + c));
  __syncthreads();
  if (cnt <= kSelThreads) {
    // few candidates (the usual case): every thread ranks its own key by counting the larger ones (keys are
    // distinct: the row is part of the key) — one pass of broadcast shared-memory reads instead of 2 k rounds of
    // warp reductions; the same k keys in the same order as block_topk_u64
    const uint64_t mine = tid < cnt ? fkeys[tid] : 0ull;
    int rank = 0;
    for (int e = 0; e < cnt; ++e) { const uint64_t v = fkeys[e]; rank += (v > mine || (v == mine && e < tid)) ? 1 : 0; }
    if (tid < kList) top[tid] = 0ull;
    __syncthreads();
    if (tid < cnt && mine != 0ull && rank < a.k) top[rank] = mine;
    __syncthreads();
  } else {
    block_topk_u64(fkeys, cnt, a.k, wtop, top);
  }
  for (int j = tid; j < a.k; j += kSelThreads) {
    const uint64_t best = top[j];
    float ip = -3.402823466e38f, dist = 3.402823466e38f;
    int64_t lab = -1;
    if (best != 0ull) {
      const int32_t row = key_row(best);
      ip = key_score(best);
      lab = a.base_offset + row;
      dist = fmaxf(__fmaf_rn(-2.0f, ip, __fadd_rn(a.qn2[q], a.norm2[row])), 0.f);
    }
    if (a.out_ip) a.out_ip[(int64_t)q * a.k + j] = ip;
    a.out_dist[(int64_t)q * a.k + j] = dist;
    a.out_labels[(int64_t)q * a.k + j] = lab;
  }
}

}  // namespace b2k
