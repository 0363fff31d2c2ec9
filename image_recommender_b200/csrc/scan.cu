// K-scan: small-batch (nq <= 4 per pass) approximate scoring of the whole bf16 shard on the
// CUDA cores, HBM-bound by construction.  Replaces the inner loop of
// index.search(query_vec, top_k) (main/search_from_image.py:247) for tiny query batches
// (the reference is always batch-1).
//
// Layout: one persistent CTA per SM owns a contiguous row range (split).  A producer thread
// streams 8-row slabs (contiguous bytes) into a shared-memory ring with 1-D bulk async
// copies (UBLKCP) completing on mbarriers; 8 consumer warps read the slab with conflict-free
// 16-byte LDS, multiply against the fp32 query held in registers, and reduce 32 (row,query)
// sums per step with a transposed warp butterfly (31 shuffles for 32 values).  The running
// top-32 per (query, split) lives in shared memory, one entry per lane of warp 0, guarded by a
// threshold so inserts become rare.  Only 32 records per (query, split) ever reach HBM.
//
// Also here: K-exact, the fp32/fp64 exhaustive scan serving queries whose certificate failed.
#include <algorithm>

#include "tail_common.cuh"
#include "ptx.cuh"

namespace b2k {

namespace {
constexpr int kScanRowsPerStage = 8;
constexpr int kScanConsumers = 256;
constexpr int kScanThreads = kScanConsumers + 32;
constexpr int kScanSmemBudget = 200 * 1024;

__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint64_t u = __shfl_xor_sync(0xffffffffu, v, o);
    v = u < v ? u : v;
  }
  return v;
}

// Insert (s,row) into the lane-per-entry list `lst` (shared, 32 Cands), replacing the worst
// entry (lowest key; empty slots first).  Returns the new threshold: the worst score when the
// list is full, else -inf.  Callers only insert entries that beat the current threshold.
__device__ __forceinline__ float warp_list_insert(Cand* lst, int lane, float s, int32_t row) {
  Cand e = lst[lane];
  const uint64_t k = cand_key(e.score, e.row);
  const uint64_t kmin = warp_min_u64(k);
  const unsigned vm = __ballot_sync(0xffffffffu, k == kmin);   // ties only among empty slots
  if (lane == __ffs(vm) - 1) { e.score = s; e.row = row; lst[lane] = e; }
  __syncwarp();
  float sc = e.row < 0 ? -INFINITY : e.score;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sc = fminf(sc, __shfl_xor_sync(0xffffffffu, sc, o));
  return sc;
}
}  // namespace

template <int NQ, int CH>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_bf16_kernel(ScanArgs a, int n_stages, int stage_bytes) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* ring = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)n_stages * stage_bytes);
  uint64_t* empty = full + n_stages;
  float* red = reinterpret_cast<float*>(empty + n_stages);      // [2][8][32]
  Cand* lists = reinterpret_cast<Cand*>(red + 2 * 8 * 32);      // [NQ][32]

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x;
  const int64_t r_begin = a.n_rows * (int64_t)split / a.n_splits;
  const int64_t r_end = a.n_rows * (int64_t)(split + 1) / a.n_splits;
  const int64_t n_local = r_end - r_begin;
  const int n_iter = (int)((n_local + kScanRowsPerStage - 1) / kScanRowsPerStage);
  const int row_bytes = a.Dp * 2;

  if (tid == 0) {
    for (int s = 0; s < n_stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 8); }
    ptx::fence_mbar_init();
  }
  if (tid < NQ * 32) { lists[tid].score = -INFINITY; lists[tid].row = -1; }
  __syncthreads();

  if (warp == 8) {
    // ------------------------------ producer ------------------------------
    if (lane == 0) {
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % n_stages;
        if (it >= n_stages) ptx::mbar_wait(&empty[s], ((it / n_stages) - 1) & 1);
        const int64_t r0 = r_begin + (int64_t)it * kScanRowsPerStage;
        const int nr = (int)min((int64_t)kScanRowsPerStage, r_end - r0);
        const uint32_t bytes = (uint32_t)nr * row_bytes;
        ptx::mbar_arrive_expect_tx(&full[s], bytes);
        ptx::bulk_g2s(ring + (size_t)s * stage_bytes, a.db + r0 * a.Dp, bytes, &full[s]);
      }
    }
    return;
  }

  // ------------------------------ consumers ------------------------------
  const int n_chunks = a.Dp >> 3;
  // query registers: chunk tid (+256 for CH==2), 8 floats per query
  float qr[CH][NQ][8];
#pragma unroll
  for (int m = 0; m < CH; ++m) {
    const int c = tid + m * kScanConsumers;
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = c * 8 + j;
        const int qi = a.q0 + q;
        qr[m][q][j] = (c < n_chunks && col < a.D && qi < a.nq) ? a.q[(int64_t)qi * a.D + col] : 0.f;
      }
  }

  constexpr int kRowsPerGroup = 32 / NQ;                       // rows per reduction
  constexpr int kSubs = kRowsPerGroup / kScanRowsPerStage;     // stages per reduction
  static_assert(kSubs >= 1, "NQ too large");
  float tau[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) tau[q] = -INFINITY;

  int it = 0;
  int buf = 0;
  for (int64_t g0 = 0; g0 < n_local; g0 += kRowsPerGroup) {
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
#pragma unroll
    for (int sub = 0; sub < kSubs; ++sub) {
      if (it < n_iter) {
        const int s = it % n_stages;
        ptx::mbar_wait(&full[s], (it / n_stages) & 1);
        const unsigned char* slab = ring + (size_t)s * stage_bytes;
#pragma unroll
        for (int m = 0; m < CH; ++m) {
          const int c = tid + m * kScanConsumers;
          if (c < n_chunks) {
            uint4 v[kScanRowsPerStage];
#pragma unroll
            for (int r = 0; r < kScanRowsPerStage; ++r)
              v[r] = *reinterpret_cast<const uint4*>(slab + (size_t)r * row_bytes + (size_t)c * 16);
#pragma unroll
            for (int r = 0; r < kScanRowsPerStage; ++r) {
              float x[8];
              x[0] = __uint_as_float(v[r].x << 16); x[1] = __uint_as_float(v[r].x & 0xffff0000u);
              x[2] = __uint_as_float(v[r].y << 16); x[3] = __uint_as_float(v[r].y & 0xffff0000u);
              x[4] = __uint_as_float(v[r].z << 16); x[5] = __uint_as_float(v[r].z & 0xffff0000u);
              x[6] = __uint_as_float(v[r].w << 16); x[7] = __uint_as_float(v[r].w & 0xffff0000u);
#pragma unroll
              for (int q = 0; q < NQ; ++q) {
                float s2 = acc[(sub * kScanRowsPerStage + r) * NQ + q];
#pragma unroll
                for (int j = 0; j < 8; ++j) s2 = fmaf(x[j], qr[m][q][j], s2);
                acc[(sub * kScanRowsPerStage + r) * NQ + q] = s2;
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[s]);
        ++it;
      }
    }
    // transposed butterfly: afterwards lane L holds Σ_lanes acc[L]
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const bool upper = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < o; ++i) {
        const float send = upper ? acc[i] : acc[i + o];
        const float keep = upper ? acc[i + o] : acc[i];
        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    red[(buf * 8 + warp) * 32 + lane] = acc[0];
    ptx::named_bar_sync(1, kScanConsumers);
    if (warp == 0) {
      float sc = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sc += red[(buf * 8 + w) * 32 + lane];
      const int rr = lane / NQ, qq = lane % NQ;
      const int64_t lrow = g0 + rr;                      // row within the split
      const int32_t row = (int32_t)(r_begin + lrow);
      float my_tau = tau[0];
#pragma unroll
      for (int q = 1; q < NQ; ++q) my_tau = (qq == q) ? tau[q] : my_tau;
      bool want = (lrow < n_local) && (a.q0 + qq < a.nq) && (sc > my_tau);
      unsigned m = __ballot_sync(0xffffffffu, want);
      while (m) {
        const int src = __ffs(m) - 1;
        const float s_in = __shfl_sync(0xffffffffu, sc, src);
        const int32_t r_in = __shfl_sync(0xffffffffu, row, src);
        const int q_in = src % NQ;
        const float nt = warp_list_insert(lists + q_in * 32, lane, s_in, r_in);
#pragma unroll
        for (int q = 0; q < NQ; ++q) if (q == q_in) tau[q] = nt;
        if (lane == src) want = false;
        if (qq == q_in) want = want && (sc > nt);
        m = __ballot_sync(0xffffffffu, want);
      }
    }
    buf ^= 1;
  }

  // all consumers must be past their last red[] read before exit is irrelevant; only warp 0
  // touches the lists, so it can write them out directly.
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int qi = a.q0 + q;
      if (qi < a.nq) a.partial[((int64_t)qi * a.n_lists + split) * kList + lane] = lists[q * 32 + lane];
    }
  }
}

int scan_num_splits(int n_sm) { return n_sm; }
bool scan_supports(int Dp) { return Dp >= 8 && (Dp % 8) == 0 && Dp <= 2 * kScanConsumers * 8; }

template <int NQ, int CH>
static int launch_scan_t(const ScanArgs& a, cudaStream_t st) {
  const int stage_bytes = kScanRowsPerStage * a.Dp * 2;
  const int fixed = 2 * 8 * 32 * 4 + NQ * 32 * (int)sizeof(Cand) + 256;
  int n_stages = (kScanSmemBudget - fixed) / (stage_bytes + 16);
  if (n_stages > 8) n_stages = 8;
  if (n_stages < 2) { set_error("scan: Dp=%d does not fit the shared-memory ring", a.Dp); return B2K_E_INVALID; }
  const size_t smem = (size_t)n_stages * stage_bytes + (size_t)n_stages * 16 + fixed;
  auto kern = scan_bf16_kernel<NQ, CH>;
  B2K_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<a.n_splits, kScanThreads, smem, st>>>(a, n_stages, stage_bytes);
  B2K_CHECK_LAUNCH();
  return 0;
}

int launch_scan(const ScanArgs& a, int nq_pass, cudaStream_t st) {
  if (!scan_supports(a.Dp)) { set_error("scan: unsupported Dp=%d", a.Dp); return B2K_E_INVALID; }
  const bool ch2 = (a.Dp >> 3) > kScanConsumers;
  if (nq_pass <= 1) return ch2 ? launch_scan_t<1, 2>(a, st) : launch_scan_t<1, 1>(a, st);
  if (nq_pass <= 2) return ch2 ? launch_scan_t<2, 2>(a, st) : launch_scan_t<2, 1>(a, st);
  if (nq_pass <= 4) return ch2 ? launch_scan_t<4, 2>(a, st) : launch_scan_t<4, 1>(a, st);
  set_error("scan: at most 4 queries per pass (got %d)", nq_pass);
  return B2K_E_INVALID;
}

// =========================================================================================
// K-collect: second look at the DB splits whose partial list was saturated for a query (K-select
// found all 32 entries at or above the candidate threshold, so the list may hide more such rows —
// the usual cause is a run of near-duplicate images stored next to each other).  The split is
// re-scored for that query alone (fp32 query x bf16 rows, fp32 accumulation: the K-scan arithmetic
// class, bound eps) and EVERY row with score >= lb - eps is appended to the query's re-rank
// candidates, where lb <= exact k-th best score.  A row of the exact top-k has exact score >= lb,
// hence approximate score >= lb - eps: none is missed, and the query stays certified.  Only a
// candidate-buffer overflow (thousands of near-ties) still needs the exhaustive K-exact scan.
// Work items = (pair, 1/kCollectSub of its split), one warp per row; launched unconditionally and
// exits at once when no list was saturated.
namespace {
constexpr int kCollectSub = 64;
constexpr int kCollectWarps = 8;
}  // namespace

// One (saturated pair, sub-split) work item: every row of the sub-split at or above the threshold is appended.
// Returns the query of the item.  Block-wide.
static __device__ __forceinline__ int collect_item(const CollectArgs& a, int64_t w, int64_t tiles_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_chunks = a.Dp >> 3;
  const int2 pr = a.sat_pairs[w / kCollectSub];
  const int sub = (int)(w % kCollectSub);
  const int q = pr.x, split = pr.y;
  int64_t r_begin, r_end;
  if (a.tile_rows > 0) {
    r_begin = (tiles_total * split / a.n_splits) * a.tile_rows;
    r_end = min(a.n_rows, (tiles_total * (split + 1) / a.n_splits) * a.tile_rows);
  } else {
    r_begin = a.n_rows * (int64_t)split / a.n_splits;
    r_end = a.n_rows * (int64_t)(split + 1) / a.n_splits;
  }
  const int64_t len = r_end - r_begin;
  const int64_t s0 = r_begin + len * sub / kCollectSub, s1 = r_begin + len * (sub + 1) / kCollectSub;
  const float thr = __fsub_rd(a.lb[q], a.eps[q]);
  if (!(thr == thr)) {     // poisoned bound (NaN rows in the shard): only the exhaustive scan is safe
    if (threadIdx.x == 0) atomicOr(a.flags + q, 1);
    return q;
  }
  const float* __restrict__ qv = a.q + (int64_t)q * a.D;
  const bool qvec = ((a.D & 7) == 0) && ((reinterpret_cast<uintptr_t>(qv) & 15) == 0);
  for (int64_t row = s0 + warp; row < s1; row += kCollectWarps) {
    const uint4* __restrict__ x = reinterpret_cast<const uint4*>(a.db + row * a.Dp);
    float acc = 0.f;
    for (int c = lane; c < n_chunks; c += 32) {
      const uint4 v = __ldcs(x + c);
      float qf[8];
      if (qvec && c * 8 < a.D) {
        const float4 q0 = __ldg(reinterpret_cast<const float4*>(qv) + 2 * c);
        const float4 q1 = __ldg(reinterpret_cast<const float4*>(qv) + 2 * c + 1);
        qf[0] = q0.x; qf[1] = q0.y; qf[2] = q0.z; qf[3] = q0.w;
        qf[4] = q1.x; qf[5] = q1.y; qf[6] = q1.z; qf[7] = q1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int i = c * 8 + j; qf[j] = i < a.D ? __ldg(qv + i) : 0.f; }
      }
      acc = fmaf(__uint_as_float(v.x << 16), qf[0], acc); acc = fmaf(__uint_as_float(v.x & 0xffff0000u), qf[1], acc);
      acc = fmaf(__uint_as_float(v.y << 16), qf[2], acc); acc = fmaf(__uint_as_float(v.y & 0xffff0000u), qf[3], acc);
      acc = fmaf(__uint_as_float(v.z << 16), qf[4], acc); acc = fmaf(__uint_as_float(v.z & 0xffff0000u), qf[5], acc);
      acc = fmaf(__uint_as_float(v.w << 16), qf[6], acc); acc = fmaf(__uint_as_float(v.w & 0xffff0000u), qf[7], acc);
    }
    acc = warp_sum_f32(acc);
    if (lane == 0 && acc >= thr) {
      const int pos = atomicAdd(a.cand_count + q, 1);
      if (pos < a.cand_cap) a.cand_rows[(int64_t)q * a.cand_cap + pos] = (int32_t)row;
      else atomicOr(a.flags + q, 2);
    }
  }
  return q;
}

__global__ void __launch_bounds__(kCollectWarps * 32)
collect_kernel(CollectArgs a) {
  const int n_pairs = min(*a.sat_count, a.sat_cap);
  if (n_pairs == 0) return;
  const int64_t tiles_total = a.tile_rows > 0 ? (a.n_rows + a.tile_rows - 1) / a.tile_rows : 0;
  for (int64_t w = blockIdx.x; w < (int64_t)n_pairs * kCollectSub; w += gridDim.x) collect_item(a, w, tiles_total);
}

// K-collect + deferred finish (k <= 32; kernels.h: DeferredArgs): the CTA that completes the LAST work item of a
// query (a per-query counter, no grid-wide barrier, no co-residency assumption) re-ranks the query's candidates
// and finalises it.  Launched unconditionally behind the fused tail kernel; exits at once when no list was
// saturated.  Dynamic shared memory: cand_cap u64 keys.
__global__ void __launch_bounds__(kCollectWarps * 32)
collect_finish_kernel(CollectArgs a, DeferredArgs d) {
  extern __shared__ uint64_t fkeys[];
  __shared__ uint64_t wtop[kCollectWarps * 32];
  __shared__ uint64_t top[kList];
  __shared__ int s_last;
  const int n_pairs = min(*a.sat_count, a.sat_cap);
  if (n_pairs == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t tiles_total = a.tile_rows > 0 ? (a.n_rows + a.tile_rows - 1) / a.tile_rows : 0;
  for (int64_t w = blockIdx.x; w < (int64_t)n_pairs * kCollectSub; w += gridDim.x) {
    const int q = collect_item(a, w, tiles_total);
    __syncthreads();                                   // every warp's appends are issued
    if (threadIdx.x == 0) {
      __threadfence();                                 // ... and visible before the item counts as done
      const int done = atomicAdd(d.done + q, 1) + 1;
      s_last = done == d.sat_n[q] * kCollectSub;
    }
    __syncthreads();
    if (!s_last) continue;
    __threadfence();                                   // the other CTAs' appends of this query
    if (d.state[q] != 1) continue;                     // already handed to K-exact by the tail kernel
    const int flag = __ldcg(a.flags + q);
    if (flag != 0) {                                   // overflow / poisoned bound while collecting: K-exact
      if (threadIdx.x == 0) {
        const int slot = atomicAdd(d.fa.fail_count, 1);
        d.fa.fail_list[slot] = q;
      }
      continue;
    }
    const int cnt = min(__ldcg(a.cand_count + q), a.cand_cap);
    rerank_query(d.rr, q, cnt, warp, kCollectWarps, lane, nullptr);
    __syncthreads();
    finalize_small_k(d.fa, q, cnt, fkeys, wtop, top);
    __syncthreads();                                   // fkeys / top are reused by the next query this CTA finishes
  }
}

int launch_collect(const CollectArgs& a, int n_sm, cudaStream_t st) {
  collect_kernel<<<4 * n_sm, kCollectWarps * 32, 0, st>>>(a);
  B2K_CHECK_LAUNCH();
  return 0;
}

int launch_collect_finish(const CollectArgs& a, const DeferredArgs& d, int n_sm, cudaStream_t st) {
  static_assert(kCollectWarps * 32 == kSelThreads, "finalize_small_k strides by kSelThreads");
  const size_t smem = (size_t)a.cand_cap * sizeof(uint64_t);
  if (smem > 200 * 1024) { set_error("collect: %d candidate slots do not fit shared memory", a.cand_cap); return B2K_E_INVALID; }
  if (smem > 48 * 1024)
    B2K_CUDA(cudaFuncSetAttribute(collect_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  collect_finish_kernel<<<4 * n_sm, kCollectWarps * 32, smem, st>>>(a, d);
  B2K_CHECK_LAUNCH();
  return 0;
}

// =========================================================================================
// K-exact: exhaustive fp32 scan with fp64 accumulation (Spec R) for uncertified queries.
// One warp per row, up to 4 failed queries per sweep, warp-private top-32 in registers
// (one entry per lane), merged per CTA at the end.  Launched unconditionally; exits at once
// when fail_count == 0, so no host synchronisation is needed to decide.
namespace {
constexpr int kExactFQ = 4;
constexpr int kExactWarps = 8;
constexpr int kExactRows = 4;       // rows per warp step of the staged-query form

}  // namespace

// Offer (ip, row) to the warp's register-resident top-32 of failed query f (one entry per lane).
__device__ __forceinline__ void exact_offer(float ip, int64_t row, uint64_t ceil_key, float& e_s, int32_t& e_r,
                                            float& tau, int lane) {
  if (ip > tau && cand_key(ip, (int32_t)row) < ceil_key) {   // equal scores: the earlier (lower) row already listed wins
    const uint64_t k = cand_key(e_s, e_r);
    const uint64_t kmin = warp_min_u64(k);
    const unsigned vm = __ballot_sync(0xffffffffu, k == kmin);
    if (lane == __ffs(vm) - 1) { e_s = ip; e_r = (int32_t)row; }
    float sc = e_r < 0 ? -INFINITY : e_s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sc = fminf(sc, __shfl_xor_sync(0xffffffffu, sc, o));
    tau = sc;
  }
}

// q_smem: the (<= 4) failed queries are staged as fp64 in dynamic shared memory and every warp step
// takes kExactRows rows, so each row element costs ONE fp32->fp64 conversion (the 16/clk/SM
// conversion rate, not HBM, bounded the one-row-at-a-time form) and each staged query chunk serves
// four rows.  The FMA order per lane is Spec R's in both forms.
// fused (k <= 32, one page): the CTA that completes a group's scan LAST (per-group counter, no grid barrier)
// also finalises the group's queries — K-exact is then ONE launch.  Dynamic shared memory then also holds
// n_splits * 32 u64 keys (it is re-staged with the next group's queries afterwards).
__global__ void __launch_bounds__(kExactWarps * 32)
exact_scan_kernel(ExactArgs a, int q_smem, int fused) {
  extern __shared__ double qd[];                                    // [kExactFQ][D] when q_smem
  __shared__ uint64_t keys[kExactFQ][kExactWarps * 32];
  __shared__ uint64_t ftop[kList];
  __shared__ int s_last;
  const int n_fail = *a.fail_count;
  if (n_fail == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t gw = (int64_t)blockIdx.x * kExactWarps + warp;
  const int64_t nw = (int64_t)gridDim.x * kExactWarps;
  const bool vec = ((a.D & 3) == 0);
  for (int f0 = 0; f0 < n_fail; f0 += kExactFQ) {
    const float* qp[kExactFQ];
    float e_s[kExactFQ]; int32_t e_r[kExactFQ]; float tau[kExactFQ];
    uint64_t ceil_key[kExactFQ];      // paging (k > 32): only rows strictly after the last emitted result
#pragma unroll
    for (int f = 0; f < kExactFQ; ++f) {
      const int fi = f0 + f < n_fail ? f0 + f : n_fail - 1;   // duplicates are harmless
      qp[f] = a.q + (int64_t)a.fail_list[fi] * a.D;
      e_s[f] = -INFINITY; e_r[f] = -1; tau[f] = -INFINITY;
      ceil_key[f] = a.page == 0 ? ~0ull : a.ceil_keys[fi];
    }
    if (q_smem) {
      __syncthreads();                                              // previous group's readers are done
#pragma unroll
      for (int f = 0; f < kExactFQ; ++f)
        for (int i = threadIdx.x; i < a.D; i += blockDim.x) qd[(size_t)f * a.D + i] = (double)__ldg(qp[f] + i);
      __syncthreads();
      const int n4 = a.D >> 2;
      for (int64_t row0 = gw * kExactRows; row0 < a.n_rows; row0 += nw * kExactRows) {
        const float4* x4[kExactRows];
#pragma unroll
        for (int r = 0; r < kExactRows; ++r)                        // rows past the end re-read the last row, unused
          x4[r] = reinterpret_cast<const float4*>(a.db_f32 + min(row0 + r, a.n_rows - 1) * (int64_t)a.D);
        double p[kExactRows][kExactFQ];
#pragma unroll
        for (int r = 0; r < kExactRows; ++r)
#pragma unroll
          for (int f = 0; f < kExactFQ; ++f) p[r][f] = 0.0;
        float4 nxt[kExactRows];
#pragma unroll
        for (int r = 0; r < kExactRows; ++r) nxt[r] = lane < n4 ? __ldcs(x4[r] + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = lane; c < n4; c += 32) {
          double xv[kExactRows][4];
#pragma unroll
          for (int r = 0; r < kExactRows; ++r) {
            xv[r][0] = (double)nxt[r].x; xv[r][1] = (double)nxt[r].y; xv[r][2] = (double)nxt[r].z; xv[r][3] = (double)nxt[r].w;
          }
          if (c + 32 < n4) {                                        // next chunk's loads fly during this chunk's FMAs
#pragma unroll
            for (int r = 0; r < kExactRows; ++r) nxt[r] = __ldcs(x4[r] + c + 32);
          }
#pragma unroll
          for (int f = 0; f < kExactFQ; ++f) {
            const double2 w01 = *reinterpret_cast<const double2*>(qd + (size_t)f * a.D + 4 * c);
            const double2 w23 = *reinterpret_cast<const double2*>(qd + (size_t)f * a.D + 4 * c + 2);
#pragma unroll
            for (int r = 0; r < kExactRows; ++r) {
              p[r][f] = __fma_rn(w01.x, xv[r][0], p[r][f]); p[r][f] = __fma_rn(w01.y, xv[r][1], p[r][f]);
              p[r][f] = __fma_rn(w23.x, xv[r][2], p[r][f]); p[r][f] = __fma_rn(w23.y, xv[r][3], p[r][f]);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < kExactRows; ++r) {
          if (row0 + r >= a.n_rows) break;                          // warp-uniform
#pragma unroll
          for (int f = 0; f < kExactFQ; ++f)
            exact_offer((float)warp_sum_f64(p[r][f]), row0 + r, ceil_key[f], e_s[f], e_r[f], tau[f], lane);
        }
      }
    } else {
      for (int64_t row = gw; row < a.n_rows; row += nw) {
        const float* x = a.db_f32 + row * (int64_t)a.D;
        double p[kExactFQ];
#pragma unroll
        for (int f = 0; f < kExactFQ; ++f) p[f] = 0.0;
        if (vec) {
          const float4* x4 = reinterpret_cast<const float4*>(x);
          for (int c = lane; c < (a.D >> 2); c += 32) {
            const float4 v = x4[c];
#pragma unroll
            for (int f = 0; f < kExactFQ; ++f) {
              const float4 w = __ldg(reinterpret_cast<const float4*>(qp[f]) + c);
              p[f] = __fma_rn((double)w.x, (double)v.x, p[f]);
              p[f] = __fma_rn((double)w.y, (double)v.y, p[f]);
              p[f] = __fma_rn((double)w.z, (double)v.z, p[f]);
              p[f] = __fma_rn((double)w.w, (double)v.w, p[f]);
            }
          }
        } else {
          for (int c = lane; c * 4 < a.D; c += 32)
            for (int j = 0; j < 4; ++j) {
              const int i = c * 4 + j;
              if (i < a.D) {
                const double v = (double)x[i];
#pragma unroll
                for (int f = 0; f < kExactFQ; ++f) p[f] = __fma_rn((double)__ldg(qp[f] + i), v, p[f]);
              }
            }
        }
#pragma unroll
        for (int f = 0; f < kExactFQ; ++f)
          exact_offer((float)warp_sum_f64(p[f]), row, ceil_key[f], e_s[f], e_r[f], tau[f], lane);
      }
    }
    // CTA merge of the 8 warp lists -> top-32 per failed query
#pragma unroll
    for (int f = 0; f < kExactFQ; ++f) keys[f][threadIdx.x] = cand_key(e_s[f], e_r[f]);
    __syncthreads();
    for (int f = 0; f < kExactFQ; ++f) {
      block_bitonic_desc(keys[f], kExactWarps * 32);
      if (f0 + f < n_fail && threadIdx.x < kList) {
        Cand c; const uint64_t k = keys[f][threadIdx.x];
        c.score = key_score(k); c.row = key_row(k);
        if (c.row < 0) c.score = -INFINITY;
        a.partial[((int64_t)(f0 + f) * a.n_splits + blockIdx.x) * kList + threadIdx.x] = c;
      }
    }
    __syncthreads();
    if (!fused) continue;
    if (threadIdx.x == 0) {
      __threadfence();
      s_last = atomicAdd(a.group_done + f0 / kExactFQ, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) continue;
    __threadfence();
    uint64_t* skeys = reinterpret_cast<uint64_t*>(qd);              // the staged queries are dead for this group
    const int M = a.n_splits * kList;
    for (int f = 0; f < kExactFQ && f0 + f < n_fail; ++f) {
      const int q = a.fail_list[f0 + f];
      const Cand* src = a.partial + (int64_t)(f0 + f) * M;
      for (int i = threadIdx.x; i < M; i += blockDim.x) {
        const float sc = __ldcg(&src[i].score);
        const int32_t rw = __ldcg(&src[i].row);
        skeys[i] = cand_key(sc, rw);
      }
      __syncthreads();
      block_topk_u64(skeys, M, a.k, &keys[0][0], ftop);            // keys[][]: 4 * 256 u64 >= 8 warps * 32 of scratch
      if ((int)threadIdx.x < a.k) {
        const uint64_t kk = ftop[threadIdx.x];
        const int32_t row = key_row(kk);
        float ip = -3.402823466e38f, dist = 3.402823466e38f; int64_t lab = -1;
        if (row >= 0) {
          ip = key_score(kk);
          lab = a.base_offset + row;
          dist = fmaxf(__fmaf_rn(-2.0f, ip, __fadd_rn(a.qn2[q], a.norm2[row])), 0.f);
        }
        if (a.out_ip) a.out_ip[(int64_t)q * a.k + threadIdx.x] = ip;
        a.out_dist[(int64_t)q * a.k + threadIdx.x] = dist;
        a.out_labels[(int64_t)q * a.k + threadIdx.x] = lab;
      }
      __syncthreads();
    }
  }
}

// One CTA per failed query: sort n_splits*32 exact records, emit the top-k.
__global__ void __launch_bounds__(512)
exact_finalize_kernel(ExactArgs a, int P) {
  extern __shared__ uint64_t skeys[];
  const int n_fail = *a.fail_count;
  const int f = blockIdx.x;
  if (f >= n_fail) return;
  const int q = a.fail_list[f];
  const int M = a.n_splits * kList;
  const Cand* src = a.partial + (int64_t)f * M;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    uint64_t k = 0;
    if (i < M) { const Cand c = src[i]; k = cand_key(c.score, c.row); }
    skeys[i] = k;
  }
  __syncthreads();
  block_bitonic_desc(skeys, P);
  const int j0 = a.page * kList;                      // this page emits results [j0, j0 + 32)
  if (threadIdx.x < kList && j0 + (int)threadIdx.x < a.k) {
    const int j = j0 + threadIdx.x;
    const uint64_t k = skeys[threadIdx.x];
    const int32_t row = key_row(k);
    float ip = -3.402823466e38f, dist = 3.402823466e38f; int64_t lab = -1;
    if (row >= 0) {
      ip = key_score(k);
      lab = a.base_offset + row;
      dist = fmaxf(__fmaf_rn(-2.0f, ip, __fadd_rn(a.qn2[q], a.norm2[row])), 0.f);
    }
    if (a.out_ip) a.out_ip[(int64_t)q * a.k + j] = ip;
    a.out_dist[(int64_t)q * a.k + j] = dist;
    a.out_labels[(int64_t)q * a.k + j] = lab;
  }
  // next page continues strictly after this page's last result (0: nothing is left)
  if (threadIdx.x == 0) a.ceil_keys[f] = skeys[kList - 1];
}

int exact_num_splits(int n_sm) { return 2 * n_sm; }

int launch_exact(const ExactArgs& a, cudaStream_t st) {
  // staged fp64 queries need 4 x D x 8 bytes of shared memory, 16-byte aligned fp32 rows and D % 4 == 0
  const size_t q_bytes = (size_t)kExactFQ * a.D * sizeof(double);
  const int q_smem = (a.D & 3) == 0 && q_bytes <= 160 * 1024 && (reinterpret_cast<uintptr_t>(a.db_f32) & 15) == 0;
  if (q_smem && q_bytes > 32 * 1024)
    B2K_CUDA(cudaFuncSetAttribute(exact_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q_bytes));
  exact_scan_kernel<<<a.n_splits, kExactWarps * 32, q_smem ? q_bytes : 0, st>>>(a, q_smem, 0);
  B2K_CHECK_LAUNCH();
  int P = 1; while (P < a.n_splits * kList) P <<= 1;
  const size_t smem = (size_t)P * 8;
  B2K_CUDA(cudaFuncSetAttribute(exact_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  exact_finalize_kernel<<<a.nq, 512, smem, st>>>(a, P);
  B2K_CHECK_LAUNCH();
  return 0;
}

int launch_exact_fused(const ExactArgs& a, cudaStream_t st) {
  static_assert(kExactWarps * 32 == kSelThreads, "block_topk_u64 scratch is sized for kSelThreads");
  const size_t q_bytes = (size_t)kExactFQ * a.D * sizeof(double);
  const int q_smem = (a.D & 3) == 0 && q_bytes <= 160 * 1024 && (reinterpret_cast<uintptr_t>(a.db_f32) & 15) == 0;
  const size_t key_bytes = (size_t)a.n_splits * kList * sizeof(uint64_t);
  const size_t smem = std::max(q_smem ? q_bytes : (size_t)0, key_bytes);
  if (smem > 200 * 1024 || a.k > kList) { set_error("exact: fused form needs k <= %d", (int)kList); return B2K_E_INVALID; }
  if (smem > 32 * 1024)
    B2K_CUDA(cudaFuncSetAttribute(exact_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  exact_scan_kernel<<<a.n_splits, kExactWarps * 32, smem, st>>>(a, q_smem, 1);
  B2K_CHECK_LAUNCH();
  return 0;
}

}  // namespace b2k
