// K-select / K-rerank / K-finalize / K-merge: everything between the approximate (bf16)
// partial lists and the exact fp32 result of index.search (main/search_from_image.py:247).
//
//   select   : per query, b_k = k-th best bf16 score over all partial lists; every listed row
//              with bf16 score >= b_k - 2*eps becomes a re-rank candidate.  Since
//              |bf16 score - exact score| <= eps for every row of the shard, the exact top-k is
//              a subset of those rows PROVIDED no partial list is saturated (all 32 entries
//              above the threshold): that is the certificate.  Uncertified queries are served
//              by the exhaustive fp32 scan (scan.cu: K-exact).
//   rerank   : exact score (Spec R: fp64 accumulation of fp32 products, one rounding) of every
//              candidate, one warp per (query, candidate); gathers rows from the fp32 copy.
//   finalize : per query top-k of the candidates by (score desc, row asc); converts to the
//              squared-L2 distance the reference's METRIC_L2 index returns
//              (main/create_index.py:219,230) and to global offsets.
//   merge    : cross-shard merge of per-GPU top-k lists.
#include "common.cuh"
#include "kernels.h"

namespace b2k {

#ifdef B2K_PHASE_TIMERS
// debug build only (scripts/exp_phase.py): %globaltimer at the phase boundaries of select_kernel, CTA 0
__device__ unsigned long long g_phase_t[16];
#define B2K_PHASE(i) do { __syncthreads(); if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t_; \
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_phase_t[i] = t_; } } while (0)
extern "C" int b2k_debug_phase_times(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_phase_t, sizeof(g_phase_t));
}
#else
#define B2K_PHASE(i) do { } while (0)
#endif

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
__device__ __forceinline__ void block_topk_u64(const uint64_t* keys, int n, int k, uint64_t* wtop, uint64_t* out) {
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
__device__ __forceinline__ void tighten_threshold(const SelectArgs& a, int q, const Cand* lst, int E,
                                                  const uint64_t* topk, float* s_exact, float& thr, float& lb) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* qv = a.q + (int64_t)q * a.D;
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

// One CTA per query.  Dynamic smem: k <= 32: n_lists*32 u64 keys; k > 32: P u64 keys (P = the power of
// two >= n_lists*32, sorted in place) + n_lists ints + k floats.
__global__ void __launch_bounds__(kSelThreads)
select_kernel(SelectArgs a, int P) {
  extern __shared__ uint64_t skey[];
  __shared__ uint64_t wtop[(kSelThreads / 32) * 32];
  __shared__ uint64_t top[kList];
  __shared__ float s_exact32[kList];
  __shared__ int s_count, s_sat;
  const int q = blockIdx.x;
  const int E = a.n_lists * kList;
  const Cand* lst = a.partial + (int64_t)q * a.list_stride * kList;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { s_count = 0; s_sat = 0; }
  B2K_PHASE(0);
  load_list_keys(lst, E, skey);
  B2K_PHASE(1);
  int32_t* out_rows = a.cand_rows + (int64_t)q * a.cand_cap;

  if (a.k <= kList) {
    // b_k = k-th best approximate score over every list; 0: fewer than k rows listed -> everything
    // listed is a candidate.
    block_topk_u64(skey, E, a.k, wtop, top);
    B2K_PHASE(2);
    const uint32_t bk = (uint32_t)(top[a.k - 1] >> 32);
    float thr = bk != 0u ? key_minus_2eps(bk, a.eps[q]) : -INFINITY;
    // lower bound of the exact k-th best score: the k best approximate rows have exact >= b_k - eps
    float lb = bk != 0u ? __fadd_rd(thr, a.eps[q]) : -INFINITY;
    if (bk != 0u && a.db_f32 != nullptr) tighten_threshold(a, q, lst, E, top, s_exact32, thr, lb);

    B2K_PHASE(3);
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
          if (slot < a.sat_cap) a.sat_pairs[slot] = make_int2(q, l);
          else s_sat = 1;
        }
        if (a.sat_pairs) continue;     // on pair-table overflow the query is flagged: its candidates are unused
      }
      if (hm) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_count, __popc(hm));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (hit) {
          const int pos = base + __popc(hm & ((1u << lane) - 1u));
          if (pos < a.cand_cap) out_rows[pos] = lst[l * kList + lane].row;
        }
      }
    }
    __syncthreads();
    B2K_PHASE(4);
    if (tid == 0) {
      int cnt = s_count;
      int flag = 0;
      if (s_sat) flag |= 1;                       // a list may hide at-risk rows
      if (cnt > a.cand_cap) { flag |= 2; cnt = a.cand_cap; }
      if (a.force_exact) flag |= 4;
      a.cand_count[q] = cnt;
      a.flags[q] = flag;
      a.thr[q] = thr;
      a.lb[q] = lb;
    }
    return;
  }

  // ---- k > 32: sort every listed entry; the candidates are a prefix of the sorted keys
  int* l_cnt = reinterpret_cast<int*>(skey + P);                 // [n_lists] entries at or above the threshold
  float* s_exact = reinterpret_cast<float*>(l_cnt + a.n_lists);  // [k]
  for (int i = E + tid; i < P; i += kSelThreads) skey[i] = 0ull;
  for (int l = tid; l < a.n_lists; l += kSelThreads) l_cnt[l] = 0;
  __syncthreads();
  block_bitonic_desc(skey, P);
  const uint32_t bk = a.k <= E ? (uint32_t)(skey[a.k - 1] >> 32) : 0u;
  float thr = bk != 0u ? key_minus_2eps(bk, a.eps[q]) : -INFINITY;
  float lb = bk != 0u ? __fadd_rd(thr, a.eps[q]) : -INFINITY;
  if (bk != 0u && a.db_f32 != nullptr) tighten_threshold(a, q, lst, E, skey, s_exact, thr, lb);
  const uint32_t thr_key = float_key(thr);
  // pass 1: how many entries of each list reach the threshold (32 = saturated -> K-collect)
  for (int i = tid; i < E; i += kSelThreads) {
    const uint64_t key = skey[i];
    const uint32_t sk = (uint32_t)(key >> 32);
    if (sk != 0u && sk >= thr_key) atomicAdd(&l_cnt[(E - 1 - (int)(uint32_t)(key & 0xffffffffull)) / kList], 1);
  }
  __syncthreads();
  for (int l = tid; l < a.n_lists; l += kSelThreads) {
    if (l_cnt[l] == kList) {
      const int slot = a.sat_pairs ? atomicAdd(a.sat_count, 1) : a.sat_cap;
      if (slot < a.sat_cap) a.sat_pairs[slot] = make_int2(q, l);
      else s_sat = 1;
    }
  }
  // pass 2: emit the entries of the unsaturated lists
  for (int i = tid; i < E; i += kSelThreads) {
    const uint64_t key = skey[i];
    const uint32_t sk = (uint32_t)(key >> 32);
    if (sk == 0u || sk < thr_key) continue;
    const int e = E - 1 - (int)(uint32_t)(key & 0xffffffffull);
    if (l_cnt[e / kList] == kList && a.sat_pairs) continue;
    const int pos = atomicAdd(&s_count, 1);
    if (pos < a.cand_cap) out_rows[pos] = lst[e].row;
  }
  __syncthreads();
  if (tid == 0) {
    int cnt = s_count;
    int flag = 0;
    if (s_sat) flag |= 1;
    if (cnt > a.cand_cap) { flag |= 2; cnt = a.cand_cap; }
    if (a.force_exact) flag |= 4;
    a.cand_count[q] = cnt;
    a.flags[q] = flag;
    a.thr[q] = thr;
    a.lb[q] = lb;
  }
}

// One CTA per query: admission floor for the full pass from the lists of the sampling pass.
__global__ void __launch_bounds__(kSelThreads)
seed_kernel(SeedArgs a, int P) {
  extern __shared__ uint64_t skey[];
  __shared__ uint64_t wtop[(kSelThreads / 32) * 32];
  __shared__ uint64_t top[kList];
  const int q = blockIdx.x;
  const int E = a.n_lists * kList;
  load_list_keys(a.partial + (int64_t)q * a.list_stride * kList, E, skey);
  uint32_t bk;
  if (a.k <= kList) {
    block_topk_u64(skey, E, a.k, wtop, top);
    bk = (uint32_t)(top[a.k - 1] >> 32);
  } else {
    for (int i = E + threadIdx.x; i < P; i += kSelThreads) skey[i] = 0ull;
    __syncthreads();
    block_bitonic_desc(skey, P);
    bk = a.k <= E ? (uint32_t)(skey[a.k - 1] >> 32) : 0u;
  }
  if (threadIdx.x == 0) {
    // strictly below b_k(sample) - 2 eps <= b_k(shard) - 2 eps: rows at or under the floor are
    // never candidates, and rows above it are admitted (strict compare in the scoring epilogue)
    a.thr_floor[q] = bk != 0u ? nextafterf(key_minus_2eps(bk, a.eps[q]), -INFINITY) : -INFINITY;
  }
}

// ---------------------------------------------------------------------------------------
// grid = (blocks_per_query, nq); warp-per-candidate, Spec R.
__global__ void __launch_bounds__(256)
rerank_kernel(RerankArgs a, int q_smem) {
  extern __shared__ double qd[];          // [D] the query widened to fp64, when q_smem
  const int q = blockIdx.y;
  const int cnt = min(a.cand_count[q], a.cand_cap);    // K-collect may have counted past the capacity
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int w0 = blockIdx.x * wpb + (threadIdx.x >> 5);
  const int wstride = gridDim.x * wpb;
  if (blockIdx.x * wpb >= cnt) return;                 // nothing for this block (uniform)
  const float* __restrict__ qv = a.q + (int64_t)q * a.D;
  if (q_smem) {
    for (int i = threadIdx.x; i < a.D; i += blockDim.x) qd[i] = (double)__ldg(qv + i);
    __syncthreads();
  }
  for (int c = w0; c < cnt; c += wstride) {
    const int32_t row = a.cand_rows[(int64_t)q * a.cand_cap + c];
    const float* x = a.db_f32 + (int64_t)row * a.D;
    const double p = q_smem ? lane_dot64_qd(qd, x, a.D, lane) : lane_dot64(qv, x, a.D, lane);
    const float ip = (float)warp_sum_f64(p);
    if (lane == 0) a.cand_ip[(int64_t)q * a.cand_cap + c] = ip;
  }
}

// ---------------------------------------------------------------------------------------
// One CTA per query: top-k of the re-ranked candidates.  Dynamic smem: cand_cap u64 keys (k <= 32) or
// P >= cand_cap keys sorted in place (k > 32).
__global__ void __launch_bounds__(kSelThreads)
finalize_kernel(FinalizeArgs a, int P) {
  extern __shared__ uint64_t fkeys[];
  __shared__ uint64_t wtop[(kSelThreads / 32) * 32];
  __shared__ uint64_t top[kList];
  const int q = blockIdx.x;
  const int tid = threadIdx.x;
  const int flag = a.flags[q];
  if (flag != 0) {
    if (tid == 0) {
      const int slot = atomicAdd(a.fail_count, 1);
      a.fail_list[slot] = q;
    }
    return;   // K-exact writes this query's outputs
  }
  const int cnt = min(a.cand_count[q], a.cand_cap);
  for (int c = tid; c < cnt; c += kSelThreads)
    fkeys[c] = cand_key(a.cand_ip[(int64_t)q * a.cand_cap + c], a.cand_rows[(int64_t)q * a.cand_cap + c]);
  const uint64_t* best_keys = top;
  if (a.k <= kList) {
    __syncthreads();
    block_topk_u64(fkeys, cnt, a.k, wtop, top);
  } else {
    for (int c = cnt + tid; c < P; c += kSelThreads) fkeys[c] = 0ull;
    __syncthreads();
    block_bitonic_desc(fkeys, P);
    best_keys = fkeys;
  }
  for (int j = tid; j < a.k; j += kSelThreads) {
    const uint64_t best = j < P || a.k <= kList ? best_keys[j] : 0ull;
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

// ---------------------------------------------------------------------------------------
// One warp per query, n_lists <= 32 lists in the global order (what K-finalize / K-exact emit).
__global__ void __launch_bounds__(128)
merge_sorted_kernel(MergeArgs a) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= a.nq) return;
  const int64_t nq = a.nq, k = a.k;
  warp_merge_sorted(a.ip, a.dist, a.labels, [=](int g, int j) { return ((int64_t)g * nq + q) * k + j; }, a.n_lists, a.k, lane,
                    a.out_ip ? a.out_ip + (int64_t)q * a.k : nullptr, a.out_dist + (int64_t)q * a.k,
                    a.out_labels + (int64_t)q * a.k);
}

// One warp per query; any order inside the lists, n_lists*k <= 32*32 entries (more than 32 lists).
__global__ void __launch_bounds__(128)
merge_kernel(MergeArgs a) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= a.nq) return;
  const int E = a.n_lists * a.k;
  uint32_t last_key = 0xffffffffu;
  int64_t last_off = -1;            // entries strictly after (last_key, last_off) remain
  for (int j = 0; j < a.k; ++j) {
    // lane-local best among remaining
    uint32_t bk = 0u; int64_t bo = INT64_MAX; int bsrc = -1;
    for (int e = lane; e < E; e += 32) {
      const int l = e / a.k, jj = e % a.k;
      const int64_t src = ((int64_t)l * a.nq + q) * a.k + jj;
      const int64_t off = a.labels[src];
      if (off < 0) continue;
      const uint32_t key = float_key(a.ip[src]) | 0u;
      // remaining iff (key, off) is worse than the last emitted
      const bool remaining = (j == 0) || key < last_key || (key == last_key && off > last_off);
      if (!remaining) continue;
      if (key > bk || (key == bk && off < bo) || bsrc < 0) { bk = key; bo = off; bsrc = (int)src; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint32_t ok = __shfl_xor_sync(0xffffffffu, bk, o);
      const int64_t oo = __shfl_xor_sync(0xffffffffu, bo, o);
      const int os = __shfl_xor_sync(0xffffffffu, bsrc, o);
      const bool take = os >= 0 && (bsrc < 0 || ok > bk || (ok == bk && oo < bo));
      if (take) { bk = ok; bo = oo; bsrc = os; }
    }
    if (lane == 0) {
      float ip = -3.402823466e38f, dist = 3.402823466e38f;
      int64_t lab = -1;
      if (bsrc >= 0) { ip = a.ip[bsrc]; dist = a.dist[bsrc]; lab = bo; }
      if (a.out_ip) a.out_ip[(int64_t)q * a.k + j] = ip;
      a.out_dist[(int64_t)q * a.k + j] = dist;
      a.out_labels[(int64_t)q * a.k + j] = lab;
    }
    if (bsrc < 0) { last_key = 0u; last_off = INT64_MAX; }
    else { last_key = bk; last_off = bo; }
  }
}

// ---------------------------------------------------------------------------------------
static int pow2_at_least(int n) { int p = 1; while (p < n) p <<= 1; return p; }

int launch_select(const SelectArgs& a, int nq, cudaStream_t st) {
  const int E = a.n_lists * kList;
  const int P = a.k <= kList ? 0 : pow2_at_least(E);
  const size_t smem = a.k <= kList ? (size_t)E * sizeof(uint64_t)
                                   : (size_t)P * sizeof(uint64_t) + (size_t)a.n_lists * sizeof(int) + (size_t)a.k * sizeof(float);
  if (smem > 200 * 1024) { set_error("select: too many partial lists (%d) for k=%d", a.n_lists, a.k); return B2K_E_INVALID; }
  if (smem > 48 * 1024)
    B2K_CUDA(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  select_kernel<<<nq, kSelThreads, smem, st>>>(a, P);
  B2K_CHECK_LAUNCH();
  return 0;
}

int launch_seed(const SeedArgs& a, int nq, cudaStream_t st) {
  const int E = a.n_lists * kList;
  const int P = a.k <= kList ? 0 : pow2_at_least(E);
  const size_t smem = (size_t)(a.k <= kList ? E : P) * sizeof(uint64_t);
  if (smem > 200 * 1024) { set_error("seed: too many partial lists (%d)", a.n_lists); return B2K_E_INVALID; }
  if (smem > 48 * 1024)
    B2K_CUDA(cudaFuncSetAttribute(seed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  seed_kernel<<<nq, kSelThreads, smem, st>>>(a, P);
  B2K_CHECK_LAUNCH();
  return 0;
}

int launch_rerank(const RerankArgs& a, int n_sm, cudaStream_t st) {
  int bpq = (8 * n_sm + a.nq - 1) / a.nq;
  if (bpq < 1) bpq = 1;
  if (bpq > 128) bpq = 128;
  dim3 grid(bpq, a.nq);
  // one block serves many candidates of one query only when blocks are scarce (large batches): then the
  // widened query in shared memory halves the conversions; small batches keep the latency-lean form
  const size_t q_bytes = (size_t)a.D * sizeof(double);
  const int q_smem = bpq == 1 && (a.D & 3) == 0 && q_bytes <= 96 * 1024 &&
                     ((reinterpret_cast<uintptr_t>(a.db_f32) | reinterpret_cast<uintptr_t>(a.q)) & 15) == 0;
  if (q_smem && q_bytes > 48 * 1024)
    B2K_CUDA(cudaFuncSetAttribute(rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q_bytes));
  rerank_kernel<<<grid, 256, q_smem ? q_bytes : 0, st>>>(a, q_smem);
  B2K_CHECK_LAUNCH();
  return 0;
}

int launch_finalize(const FinalizeArgs& a, cudaStream_t st) {
  const int P = a.k <= kList ? 0 : pow2_at_least(a.cand_cap);
  const size_t smem = (size_t)(a.k <= kList ? a.cand_cap : P) * sizeof(uint64_t);
  if (smem > 200 * 1024) { set_error("finalize: %d candidate slots do not fit shared memory", a.cand_cap); return B2K_E_INVALID; }
  if (smem > 48 * 1024)
    B2K_CUDA(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  finalize_kernel<<<a.nq, kSelThreads, smem, st>>>(a, P);
  B2K_CHECK_LAUNCH();
  return 0;
}

int launch_merge(const MergeArgs& a, cudaStream_t st) {
  if (a.nq <= 0) return 0;
  if (a.n_lists <= 32) merge_sorted_kernel<<<(a.nq + 3) / 4, 128, 0, st>>>(a);
  else merge_kernel<<<(a.nq + 3) / 4, 128, 0, st>>>(a);
  B2K_CHECK_LAUNCH();
  return 0;
}

}  // namespace b2k
