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
#include <algorithm>

#ifdef B2K_PHASE_TIMERS
// debug build only (scripts/exp_phase.py): %globaltimer at the phase boundaries of select_kernel, CTA 0
namespace b2k { __device__ unsigned long long g_phase_t[16]; }
#define B2K_PHASE(i) do { __syncthreads(); if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t_; \
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); b2k::g_phase_t[i] = t_; } } while (0)
extern "C" int b2k_debug_phase_times(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, b2k::g_phase_t, sizeof(b2k::g_phase_t));
}
#else
#define B2K_PHASE(i) do { } while (0)
#endif
#include "tail_common.cuh"

namespace b2k {


// One CTA per query.  Dynamic smem: k <= 32: n_lists*32 u64 keys; k > 32: P u64 keys (P = the power of
// two >= n_lists*32, sorted in place) + n_lists ints + k floats.
__global__ void __launch_bounds__(kSelThreads)
select_kernel(SelectArgs a, int P) {
  extern __shared__ uint64_t skey[];
  __shared__ uint64_t wtop[(kSelThreads / 32) * 32];
  __shared__ uint64_t top[kList];
  __shared__ __align__(8) float s_exact32[kList];
  __shared__ int s_ints[4];
  int& s_count = s_ints[0];
  int& s_sat = s_ints[1];
  const int q = blockIdx.x;
  const int E = a.n_lists * kList;
  const Cand* lst = a.partial + (int64_t)q * a.list_stride * kList;
  const int tid = threadIdx.x;
  if (tid < 4) s_ints[tid] = 0;
  if (a.k <= kList) {
    if (a.n_lists <= kSelRegLists) select_small_k_reg(a, q, wtop, top, s_exact32, s_ints);
    else select_small_k(a, q, skey, wtop, top, s_exact32, s_ints);     // first barrier inside: after the key load
    return;
  }
  load_list_keys(lst, E, skey);
  int32_t* out_rows = a.cand_rows + (int64_t)q * a.cand_cap;

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

// ---------------------------------------------------------------------------------------
// Fused tail, k <= 32 (kernels.h: TailArgs).  Cluster of C CTAs per query (gridDim.x = nq * C):
//   rank 0   : K-select -> candidates, flags, state
//   all ranks: K-rerank, warp (rank, w) takes candidates rank * 8 + w, + 8 C, ...
//   rank 0   : K-finalize
// Dynamic shared memory: max(n_lists * 32, cand_cap) u64 keys, then (q_smem) D doubles.
// kMinBlocks: 2 = the cluster form of small batches (126 registers: the lists of the selection live in registers);
//         3, 4 = one CTA per query at large batches, where the kernel is a throughput problem — gathers in
//                 flight per SM — and 2 resident CTAs of 126 registers left the SMs 25 % occupied
//                 (profiles/r02_prof_tail_b4096_*: long-scoreboard 47 %, DRAM 35 % of peak); 4 when the shared
//                 memory of four CTAs fits (lists of <= 74 splits).
template <int kMinBlocks>
__global__ void __launch_bounds__(kSelThreads, kMinBlocks)
tail_kernel(TailArgs t, int key_slots, int q_smem) {
  extern __shared__ uint64_t skey[];
  __shared__ uint64_t wtop[(kSelThreads / 32) * 32];
  __shared__ uint64_t top[kList];
  __shared__ __align__(8) float s_exact32[kList];
  __shared__ int s_ints[4];
  unsigned int rank, csize;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  const int q = blockIdx.x / (int)csize;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int state = 0;
  B2K_PHASE(0);
  if (rank == 0) {
    if (tid < 4) s_ints[tid] = 0;
    // (the dense instantiation keeps the lists in shared memory: with 80 registers the register-resident form spills)
    const int flag = (kMinBlocks < 3 && t.se.n_lists <= kSelRegLists)
                         ? select_small_k_reg(t.se, q, wtop, top, s_exact32, s_ints)
                         : select_small_k(t.se, q, skey, wtop, top, s_exact32, s_ints);
    state = flag != 0 ? 2 : (s_ints[2] > 0 ? 1 : 0);
    if (tid == 0) {
      t.state[q] = state;
      t.sat_n[q] = s_ints[2];
      if (state == 2) {                       // K-exact serves it: register it right away
        const int slot = atomicAdd(t.fa.fail_count, 1);
        t.fa.fail_list[slot] = q;
      }
    }
  }
  B2K_PHASE(1);
  if (csize > 1) {
    // rank 0's candidate list / state -> the other CTAs of the cluster (release / acquire at cluster scope)
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (rank != 0) state = __ldcg(t.state + q);
  }
  B2K_PHASE(2);
  if (state != 0) return;                     // uniform over the cluster
  // rank 0 knows the count from its own shared memory (its global copy is written by ONE thread and is only
  // ordered for the other CTAs, by the cluster barrier: reading it back here raced with that store)
  const int cnt = rank == 0 ? min(s_ints[0], t.rr.cand_cap) : min(__ldcg(t.rr.cand_count + q), t.rr.cand_cap);
  double* qd = nullptr;
  if (q_smem) {
    qd = reinterpret_cast<double*>(skey + key_slots);
    const float* qv = t.rr.q + (int64_t)q * t.rr.D;
    for (int i = tid; i < t.rr.D; i += kSelThreads) qd[i] = (double)__ldg(qv + i);
    __syncthreads();
  }
  rerank_query(t.rr, q, cnt, (int)rank * (kSelThreads >> 5) + warp, (int)csize * (kSelThreads >> 5), lane, qd);
  B2K_PHASE(3);
  if (csize > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (rank != 0) return;
  } else {
    __syncthreads();
  }
  B2K_PHASE(4);
  finalize_small_k(t.fa, q, cnt, skey, wtop, top);
  B2K_PHASE(5);
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

int launch_tail(const TailArgs& a, int nq, int n_sm, cudaStream_t st, bool dense) {
  const int E = a.se.n_lists * kList;
  // the tail kernel finalises only queries whose candidates all come from the lists (state 0: no K-collect rows):
  // at most E of them, whatever the candidate capacity
  const int key_slots = E;
  // small batches: up to 8 CTAs share one query's row gathers; large ones: one CTA per query, the query widened
  // to fp64 in shared memory (halves the conversions of a throughput-bound re-rank, as in rerank_kernel)
  int csize = 1;
  while (csize < 16 && nq * csize * 2 <= 4 * n_sm) csize *= 2;       // 16: non-portable cluster size, opted in below
  const size_t q_bytes = (size_t)a.rr.D * sizeof(double);
  const int q_smem = csize == 1 && nq >= 8 * n_sm && (a.rr.D & 3) == 0 && q_bytes <= 64 * 1024 &&
                     ((reinterpret_cast<uintptr_t>(a.rr.db_f32) | reinterpret_cast<uintptr_t>(a.rr.q)) & 15) == 0;
  const size_t smem = (size_t)key_slots * sizeof(uint64_t) + (q_smem ? q_bytes : 0);
  if (smem > 200 * 1024) { set_error("tail: %d key slots do not fit shared memory", key_slots); return B2K_E_INVALID; }
  // large batches: as many resident CTAs per SM as shared memory allows (4 on shards with <= 74 lists, else 3)
  const bool four = 4 * (smem + 4608) <= 227 * 1024;
  void (*kern)(TailArgs, int, int) = !(dense && csize == 1 && nq >= 4 * n_sm) ? tail_kernel<2>
                                     : four ? tail_kernel<4> : tail_kernel<3>;
  if (smem > 48 * 1024)
    B2K_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (csize > 8) {
    // a function attribute belongs to the CURRENT device's context: set it on every launch (a process-wide
    // "done" flag left the second GPU of a single-process group without it: invalid cluster size there)
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      cudaGetLastError();
      csize = 8;
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(nq * csize), 1, 1);
  cfg.blockDim = dim3(kSelThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = (unsigned)csize; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  B2K_CUDA(cudaLaunchKernelEx(&cfg, kern, a, key_slots, q_smem));
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
