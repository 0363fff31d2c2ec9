// C ABI of the b2k engine (include/b2k.h): index object, build (add), search pipeline,
// persistence and the synthetic-data helpers.  Host side only; kernels live in pack.cu,
// scan.cu, score_tc.cu and select.cu.
#define _FILE_OFFSET_BITS 64      // fseeko/off_t are 64-bit whatever the ABI's `long` is
#include <cuda.h>
#include <stdarg.h>
#include <stdio.h>
#include <sys/types.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace b2k {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {

constexpr int64_t kMaxStageRows = 65536;  // rows per staging chunk of add() / load() / fill_synthetic() (fewer for very wide rows)
constexpr int kMaxNqPerPass = 16384;      // queries per pipeline pass (workspace sizing)
constexpr int kDefaultCandCap = 0;      // 0: every listed entry can be a candidate (no overflow)
constexpr int kMaxStageSlots = 32;     // pinned ingest staging slots (b2k_stage_open_n)
constexpr int kCounterHead = 4;          // ints in front of the per-query counters of Workspace::fail_count
constexpr int kEvRing = 64;             // pipeline passes whose device times b2k_get_stats can average

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename T>
int dev_alloc(T** p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  B2K_CUDA(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
  return 0;
}
template <typename T>
void dev_free(T*& p) { if (p) cudaFree(p); p = nullptr; }

struct Workspace {
  int nq_cap = 0, n_lists = 0, cand_cap = 0, exact_splits = 0, k_cap = 0;
  unsigned long long* exact_ceil = nullptr;   // [nq] K-exact paging: key of the last emitted result per failed slot
  float* q = nullptr;             // [nq, D] staging for the host API
  float* parts = nullptr;         // [parts_cap, D] image vectors of b2k_search_groups
  int32_t* goffs = nullptr;       // [nq + 1] group offsets
  int64_t parts_cap = 0;
  uint16_t* q_bf16 = nullptr;     // [nq_pad, Dp]
  float *qn2 = nullptr, *eps_scan = nullptr, *eps_tc = nullptr, *thr = nullptr, *thr_floor = nullptr, *lb = nullptr;
  int2* sat_pairs = nullptr;      // [sat_cap] saturated (query, list) pairs of the last pass (K-collect's work list)
  int sat_cap = 0;
  Cand* partial = nullptr;        // [nq, n_lists, 32]
  int32_t *cand_rows = nullptr, *cand_count = nullptr, *flags = nullptr;
  float* cand_ip = nullptr;
  int32_t *fail_count = nullptr, *fail_list = nullptr, *state = nullptr, *sat_n = nullptr;
  Cand* exact_partial = nullptr;  // [nq, exact_splits, 32]
  float *out_ip = nullptr, *out_dist = nullptr;
  int64_t* out_labels = nullptr;
  void release() {
    dev_free(q); dev_free(parts); dev_free(goffs); parts_cap = 0; dev_free(q_bf16); dev_free(qn2); dev_free(eps_scan); dev_free(eps_tc); dev_free(thr); dev_free(thr_floor);
    dev_free(lb); dev_free(sat_pairs); dev_free(exact_ceil);
    dev_free(partial); dev_free(cand_rows); dev_free(cand_count); dev_free(flags); dev_free(cand_ip);
    dev_free(fail_count); dev_free(fail_list); dev_free(state); dev_free(sat_n); dev_free(exact_partial);
    dev_free(out_ip); dev_free(out_dist); dev_free(out_labels);
    nq_cap = 0;
  }
};

}  // namespace
}  // namespace b2k

using namespace b2k;

struct b2k_index {
  int device = 0, n_sm = 0;
  int n_tables = 0;
  int32_t dims[B2K_MAX_TABLES] = {0}, col_off[B2K_MAX_TABLES] = {0};
  int32_t D = 0, Dp = 0;
  int64_t cap = 0, ntotal = 0, base = 0;
  float* f32 = nullptr;
  uint16_t* bf16 = nullptr;
  float* norm2 = nullptr;
  unsigned int* stat_bits = nullptr;
  cudaStream_t stream = nullptr;
  float* stage[B2K_MAX_TABLES] = {nullptr};   // device staging of raw per-table rows
  float* stage_rows_f32 = nullptr;            // device staging of packed rows (load)
  int64_t stage_rows = kMaxStageRows;         // rows per staging chunk: <= 512 MB of fp32 per chunk
  // pinned two-slot ingest staging (b2k_stage_*): host rows -> async H2D -> K-pack, overlapped with
  // the caller filling the other slot
  float* pin[kMaxStageSlots][B2K_MAX_TABLES] = {{nullptr}};
  unsigned char* pin_base = nullptr;          // the one pinned allocation behind pin[][]
  int64_t pin_rows = 0;
  int pin_slots = 0;                          // 2 (one decoding thread) .. kMaxStageSlots (two per ingest thread)
  cudaEvent_t pin_ev[kMaxStageSlots] = {nullptr};
  bool pin_busy[kMaxStageSlots] = {false};
  Workspace ws;
  // TMA descriptors (host copies; passed by value at launch)
  alignas(64) CUtensorMap tmap_q, tmap_db, tmap_db2;   // tmap_db2: 128-row boxes for the CTA-pair kernels
  alignas(64) CUtensorMap tmap_qn;                     // queries as the N operand of the transposed kernel
  const void* tmap_qn_ptr = nullptr; int tmap_qn_rows = 0, tmap_qn_n16 = 0;
  int opt_tn = -1;                                     // transposed kernel: -1 auto, 0 never, 1 whenever it applies
  int opt_sample_wave = 1;                             // sampling pass on one wave of long strided CTAs (0: first tiles of every split)
  const void* tmap_q_ptr = nullptr; int tmap_q_rows = 0;
  const void* tmap_db_ptr = nullptr; int64_t tmap_db_rows = -1;
  int coresident[2] = {-1, -1};               // CTAs of score_tc / score_tc2 resident at once (occupancy query, lazily)
  int opt_pair = -1;                          // -1 auto (nq > 128), 0 never, 1 always
  int opt_seed = 1;                           // threshold seeding for the tcgen05 paths
  int opt_tighten = 1;                        // exact-score tightening of the candidate threshold
  int opt_collect = 1;                        // K-collect serves saturated lists (else: exhaustive scan)
  int opt_inline_seed = 1;                    // single-CTA kernel: seeding inside the main launch when eligible
  int opt_fused_tail = 1;                     // k <= 32: select + re-rank + finalize in one launch, deferred paths self-finishing
  // options
  int opt_path = 0, opt_cand_cap = kDefaultCandCap, opt_force_exact = 0, opt_scan_max_b = 0, opt_splits = 0;
  b2k_stats stats;
  int32_t* h_fail = nullptr;                  // pinned
  // timing: a ring of (before scoring, after scoring, after the tail) event triples, one per pipeline pass, so
  // that back-to-back searches can be timed kernel by kernel without a host sync in between (b2k_get_stats
  // averages the passes recorded since the previous b2k_get_stats)
  cudaEvent_t ev[kEvRing][3] = {{nullptr}};
  int64_t ev_head = 0, ev_read = 0;           // passes recorded / passes already reported
};

namespace {

int make_col_offsets(b2k_index* ix) {
  int off = 0;
  for (int t = 0; t < ix->n_tables; ++t) { ix->col_off[t] = off; off += ix->dims[t]; }
  ix->D = off;
  ix->Dp = (off + 63) / 64 * 64;
  ix->stage_rows = std::max<int64_t>(1024, std::min<int64_t>(kMaxStageRows, ((int64_t)1 << 27) / off));
  return 0;
}

int ensure_stage(b2k_index* ix) {
  for (int t = 0; t < ix->n_tables; ++t)
    if (!ix->stage[t]) { int rc = dev_alloc(&ix->stage[t], (size_t)ix->stage_rows * ix->dims[t]); if (rc) return rc; }
  return 0;
}

void fill_pack_args(const b2k_index* ix, PackArgs& a, int64_t n, int64_t row0, int normalize) {
  memset(&a, 0, sizeof(a));
  a.n_tables = ix->n_tables;
  for (int t = 0; t < ix->n_tables; ++t) {
    a.dims[t] = ix->dims[t]; a.strides[t] = ix->dims[t]; a.col_off[t] = ix->col_off[t];
  }
  a.D = ix->D; a.Dp = ix->Dp; a.normalize = normalize; a.n = n; a.row0 = row0;
  a.out_f32 = ix->f32; a.out_bf16 = ix->bf16; a.out_norm2 = ix->norm2; a.stat_bits = ix->stat_bits;
}

// Partial lists hold 32 entries per (query, DB split): the split count grows with k so that the lists
// offer ~8 slots per wanted neighbour (k <= 592: one split per SM as ever; up to 1024: two per SM).
int min_splits_for_k(int k) { return k <= kList ? 0 : (k + 3) / 4; }

int ensure_workspace(b2k_index* ix, int nq, int k) {
  Workspace& w = ix->ws;
  const int scan_lists = scan_num_splits(ix->n_sm);
  const int tc_lists = ix->n_sm;
  int n_lists = std::max(scan_lists, tc_lists);
  while (n_lists < min_splits_for_k(k)) n_lists += ix->n_sm;
  const int exact_splits = exact_num_splits(ix->n_sm);
  const int cand_cap = ix->opt_cand_cap > 0 ? ix->opt_cand_cap : n_lists * kList;
  const int k_cap = std::max(k, (int)kList);
  if (nq <= w.nq_cap && w.n_lists == n_lists && w.cand_cap == cand_cap && k_cap <= w.k_cap) return 0;
  B2K_CUDA(cudaStreamSynchronize(ix->stream));
  w.release();
  const int cap = std::max(nq, 8);
  const int nq_pad = (cap + 255) / 256 * 256;
  int rc = 0;
  if ((rc = dev_alloc(&w.q, (size_t)cap * ix->D))) return rc;
  if ((rc = dev_alloc(&w.goffs, (size_t)cap + 1))) return rc;
  if ((rc = dev_alloc(&w.q_bf16, (size_t)nq_pad * ix->Dp))) return rc;
  if ((rc = dev_alloc(&w.qn2, cap))) return rc;
  if ((rc = dev_alloc(&w.eps_scan, cap))) return rc;
  if ((rc = dev_alloc(&w.eps_tc, cap))) return rc;
  if ((rc = dev_alloc(&w.thr, cap))) return rc;
  if ((rc = dev_alloc(&w.thr_floor, cap))) return rc;
  if ((rc = dev_alloc(&w.lb, cap))) return rc;
  w.sat_cap = 1024 + 2 * cap;
  if ((rc = dev_alloc(&w.sat_pairs, (size_t)w.sat_cap))) return rc;
  if ((rc = dev_alloc(&w.partial, (size_t)cap * n_lists * kList))) return rc;
  if ((rc = dev_alloc(&w.cand_rows, (size_t)cap * cand_cap))) return rc;
  if ((rc = dev_alloc(&w.cand_ip, (size_t)cap * cand_cap))) return rc;
  if ((rc = dev_alloc(&w.cand_count, cap))) return rc;
  if ((rc = dev_alloc(&w.flags, cap))) return rc;
  // [0] failed queries, [1] saturated pairs, [2..3] grid barrier, [4, 4+cap) K-collect items done per query,
  // then ceil(cap/4)+1 K-exact CTAs done per group of failed queries
  if ((rc = dev_alloc(&w.fail_count, (size_t)kCounterHead + cap + cap / 4 + 1))) return rc;
  if ((rc = dev_alloc(&w.state, cap))) return rc;
  if ((rc = dev_alloc(&w.sat_n, cap))) return rc;
  if ((rc = dev_alloc(&w.fail_list, cap))) return rc;
  if ((rc = dev_alloc(&w.exact_partial, (size_t)cap * exact_splits * kList))) return rc;
  if ((rc = dev_alloc(&w.exact_ceil, cap))) return rc;
  if ((rc = dev_alloc(&w.out_ip, (size_t)cap * k_cap))) return rc;
  if ((rc = dev_alloc(&w.out_dist, (size_t)cap * k_cap))) return rc;
  if ((rc = dev_alloc(&w.out_labels, (size_t)cap * k_cap))) return rc;
  w.nq_cap = cap; w.n_lists = n_lists; w.cand_cap = cand_cap; w.exact_splits = exact_splits; w.k_cap = k_cap;
  ix->tmap_q_ptr = nullptr;   // q_bf16 moved
  return 0;
}

// One pass of the search pipeline over nq <= ws.nq_cap device-resident queries.
int search_pass(b2k_index* ix, const float* q_dev, int nq, int k, float* dist_dev, int64_t* labels_dev,
                float* ip_dev, cudaStream_t st) {
  Workspace& w = ix->ws;
  int launches = 0;
  const int nq_pad = (nq + 255) / 256 * 256;

  // ---- path selection
  int path = ix->opt_path;
  if (path == 0) path = (nq <= ix->opt_scan_max_b && scan_supports(ix->Dp)) ? 1 : 2;
  if (path == 2 && (!score_tc_supports(ix->Dp) || ix->ntotal == 0)) path = 1;
  if (path == 1 && !scan_supports(ix->Dp)) { set_error("no scoring path supports Dp=%d", ix->Dp); return B2K_E_INVALID; }

  QueryPrepArgs qp;
  qp.q = q_dev; qp.q_bf16 = path == 1 ? nullptr : w.q_bf16; qp.qn2 = w.qn2; qp.eps_scan = w.eps_scan;
  qp.eps_tc = w.eps_tc; qp.stat_bits = ix->stat_bits; qp.floor_init = w.thr_floor;
  qp.nq = nq; qp.nq_pad = path == 1 ? nq : nq_pad; qp.D = ix->D; qp.Dp = ix->Dp;
  int rc = launch_query_prep(qp, st);
  if (rc) return rc;
  ++launches;
  // [0] failed queries, [1] saturated pairs, [2..3] grid barrier, then per-query / per-group completion counters
  B2K_CUDA(cudaMemsetAsync(w.fail_count, 0, (size_t)(kCounterHead + w.nq_cap + w.nq_cap / 4 + 1) * sizeof(int32_t), st));

  int n_lists_used = 0, split_tile_rows = 0;
  const float* eps = nullptr;
  cudaEvent_t* ev = ix->ev[ix->ev_head % kEvRing];
  B2K_CUDA(cudaEventRecord(ev[0], st));
  if (path == 1) {
    ScanArgs sa;
    sa.db = ix->bf16; sa.n_rows = ix->ntotal; sa.D = ix->D; sa.Dp = ix->Dp; sa.q = q_dev; sa.nq = nq;
    sa.n_splits = scan_num_splits(ix->n_sm); sa.partial = w.partial; sa.n_lists = w.n_lists;
    for (int q0 = 0; q0 < nq; q0 += 4) {
      sa.q0 = q0;
      rc = launch_scan(sa, std::min(4, nq - q0), st);
      if (rc) return rc;
      ++launches;
    }
    n_lists_used = sa.n_splits;
    eps = w.eps_scan;
  } else {
    if (ix->tmap_db_ptr != ix->bf16 || ix->tmap_db_rows != ix->ntotal) {
      rc = score_tc_encode_maps(nullptr, &ix->tmap_db, nullptr, 0, ix->bf16, ix->ntotal, ix->Dp);
      if (!rc) rc = score_tc2_encode_db_map(&ix->tmap_db2, ix->bf16, ix->ntotal, ix->Dp);
      if (rc) return rc;
      ix->tmap_db_ptr = ix->bf16; ix->tmap_db_rows = ix->ntotal;
    }
    if (ix->tmap_q_ptr != w.q_bf16 || ix->tmap_q_rows != nq_pad) {
      rc = score_tc_encode_maps(&ix->tmap_q, nullptr, w.q_bf16, nq_pad, nullptr, 0, ix->Dp);
      if (rc) return rc;
      ix->tmap_q_ptr = w.q_bf16; ix->tmap_q_rows = nq_pad;
    }
    // CTA pairs work on 256-query tiles: when the last one would be at most half full (3, 5 or 7
    // tiles of 128) the single-CTA kernel wastes no tensor work and measures 6-40 % faster
    // (profiles/experiments/r01_exp_mid.log); from 1024 queries on the pair kernel's lower operand traffic wins.
    const int n_qt128 = (nq + 127) / 128;
    const bool pair = ix->opt_pair < 0 ? (nq > 128 && !((n_qt128 & 1) && n_qt128 <= 7)) : ix->opt_pair != 0;
    const int forced = std::min(ix->opt_splits, w.n_lists);
    const int min_splits = std::min(min_splits_for_k(k), w.n_lists);
    ScoreTcArgs ta;
    ta.tmap_q = &ix->tmap_q; ta.tmap_db = pair ? &ix->tmap_db2 : &ix->tmap_db;
    ta.n_rows = ix->ntotal; ta.Dp = ix->Dp; ta.nq = nq;
    ta.plan = pair ? score_tc2_plan(nq, ix->ntotal, ix->n_sm, forced, min_splits)
                   : score_tc_plan(nq, ix->ntotal, ix->n_sm, forced, min_splits);
    ta.partial = w.partial; ta.n_lists = w.n_lists; ta.max_tiles = 0; ta.thr_floor = nullptr;
    ta.seed_k = 0; ta.seed_eps = w.eps_tc; ta.seed_floor = w.thr_floor;
    ta.grid_bar = reinterpret_cast<unsigned int*>(w.fail_count + 2);
    // Threshold seeding: a sampling pass over the first tiles of every split bounds each query's
    // k-th best score from below, so the full pass admits only rows that can still matter and
    // the fused selection stops being the epilogue's bottleneck (DESIGN.md "seeding").
    const int64_t tiles_total = (ix->ntotal + 255) / 256;
    const int64_t tiles_per_split = tiles_total / ta.plan.n_splits;
    // tiny batches on short splits: one launch without a floor beats sampling + seeding + main pass
    // (a single live query per warp inserts without divergence; profiles/experiments/r01_exp_path.log)
    bool seed = ix->opt_seed && tiles_per_split >= 16 && (ix->opt_seed > 1 || nq > 4 || tiles_per_split >= 128);
    // Sample size.  A sample of f x N rows costs f x N rows of scoring and leaves ~k / f rows per query above the
    // floor (each a potential list insertion in the main pass), whatever N: the optimum is f ~ 1 / sqrt(N).
    // 1.5 sqrt(tiles) tiles in total: 0.76 % of a 10 M-row shard (the measured optimum there), 2.1 % of a 1.25 M-row
    // shard (batch 4096: 15.05 vs 16.05 ms with 0.76 %; profiles/experiments/r02_exp_sample.log).
    const int64_t sample_total = std::max<int64_t>(1, std::min<int64_t>(
        ix->opt_seed > 1 ? (int64_t)ix->opt_seed * ta.plan.n_splits : (int64_t)(1.5 * std::sqrt((double)tiles_total) + 0.5),
        tiles_total / 8));
    int sample_tiles = (int)((sample_total + ta.plan.n_splits - 1) / ta.plan.n_splits);
    sample_tiles = (int)std::max<int64_t>(1, std::min<int64_t>(sample_tiles, tiles_per_split / 8));
    // One query tile on the single-CTA kernel = one resident CTA per split: the sampling pass, the seed
    // kernel and the re-read of the sampled tiles fold into the main launch (in-kernel seeding).
    // (CTA pairs: up to 256 queries when the pairs of one query tile fill at most one wave.)
    // Transposed kernel (score_tn.cu): tensor work proportional to the live queries and an epilogue thread per DB
    // row.  Auto (only while the kernel choice itself is on auto): (a) 129..208 queries, where the M = queries
    // kernels pay for 256 (steady state on 10 M rows: 130 queries 6.9 vs 7.9 ms, 160: 7.0 vs 8.0, 192: 7.8 vs 8.3,
    // 224: equal; beyond, the pair kernel's M = 256 tile is full and its 4-deep ring of 32 KB stages wins) — on
    // shards long enough that its two extra launches (sampling pass + seed kernel; the pair kernel seeds inside
    // its launch) are noise: from 2.5 M rows up to 160 queries, from 4 M rows up to 208 (1.25 M rows: pair 1.01 vs
    // 1.06 ms at 160 queries; 5 M rows: 4.27 vs 3.68; profiles/experiments/r02_exp_tn_short.log);
    // (b) narrow rows on long shards, where the 256-column drain per tile bounds the M = queries kernel: up to 48
    // queries at <= 64 columns (D = 48, 10 M rows: 0.66 vs 0.92 ms at 32 queries, 0.30 vs 0.32 at 1), 16..48 queries
    // at <= 128 columns (0.44 vs 0.53 ms).  Behind a seeded floor only.  profiles/experiments/r02_exp_tn.log.
    const bool tn_ok = k <= kList && score_tn_supports(ix->Dp, nq) && (ix->n_sm & ~1) <= w.n_lists && forced == 0;
    const bool tn_auto = seed && ix->opt_pair < 0 && ix->opt_path == 0 &&
                         ((nq > 128 && nq <= 208 && ix->ntotal >= (nq <= 160 ? 2500000 : 4000000)) ||
                          (tiles_per_split >= 64 && nq <= 48 && (ix->Dp <= 64 || (ix->Dp <= 128 && nq >= 16))));
    const bool use_tn = tn_ok && (ix->opt_tn > 0 || (ix->opt_tn < 0 && tn_auto));
    // The grid barrier needs every CTA resident: the grid must fit what the occupancy query says this device
    // holds at once (the barrier itself is bounded in time should that still fail at run time).
    const int n_ctas = pair ? 2 * ta.plan.n_splits : ta.plan.n_splits;
    int& resident = ix->coresident[pair ? 1 : 0];
    if (resident < 0) resident = pair ? score_tc2_max_coresident(ix->n_sm) : score_tc_max_coresident(ix->n_sm);
    if (!use_tn && seed && ix->opt_inline_seed && ix->opt_seed == 1 && k <= kList && ta.plan.n_qtiles == 1 &&
        n_ctas <= resident && ta.plan.n_splits <= 160 && nq <= (pair ? 2 : 1) * n_ctas) {
      ta.seed_k = k;
      seed = false;
    }
    if (seed) {
      // The sampling pass scores the same NUMBER of tiles whatever its grid.  With the main pass's grid (every
      // split its first sample_tiles tiles) a 4096-query batch on 10 M rows is 4736 CTAs of 2 tiles each: 32 waves
      // of CTA set-up (TMEM allocation, barrier init, cluster sync, pipeline fill) around 16 us of work — 4.1-6.6 ms
      // per step in the launch list (profiles/r02_launches_raw.csv), 4 % of the step.  Instead: ONE wave of CTAs,
      // n_sm / (query tiles x CTAs per tile) splits per query tile, each scoring its share of the sample as tiles
      // spread evenly over its row range (tile_stride; contiguous runs of near-duplicate images do not dominate
      // the sample).  Fewer, longer lists also mean fewer list insertions (32 ln(rows/32) per list).
      ScoreTcArgs ts = ta;
      ts.max_tiles = sample_tiles;
      const int cpp = pair ? 2 : 1;
      const int one_wave = std::max(1, ix->n_sm / (ta.plan.n_qtiles * cpp));
      if (ix->opt_sample_wave && k <= kList && one_wave < ta.plan.n_splits) {
        const int64_t total = sample_total;
        const int64_t per_split = tiles_total / one_wave;              // the shortest split's tiles
        const int64_t mine = std::min<int64_t>((total + one_wave - 1) / one_wave, per_split);
        if (mine >= 1) {
          ts.plan = pair ? score_tc2_plan(nq, ix->ntotal, ix->n_sm, one_wave, 0) : score_tc_plan(nq, ix->ntotal, ix->n_sm, one_wave, 0);
          ts.max_tiles = (int32_t)mine;
          ts.tile_stride = (int32_t)std::max<int64_t>(1, per_split / mine);
        }
      }
      rc = pair ? launch_score_tc2(ts, st) : launch_score_tc(ts, st);
      if (rc) return rc;
      SeedArgs sd;
      sd.partial = w.partial; sd.n_lists = ts.plan.n_splits; sd.list_stride = w.n_lists; sd.k = k;
      sd.eps = w.eps_tc; sd.thr_floor = w.thr_floor;
      rc = launch_seed(sd, nq, st);
      if (rc) return rc;
      launches += 2;
      ta.max_tiles = 0; ta.thr_floor = w.thr_floor;
    }
    if (use_tn) {
      // main pass on the transposed kernel (queries on N); the floor comes from the sampling pass above
      if (ix->tmap_qn_ptr != w.q_bf16 || ix->tmap_qn_rows != nq_pad || ix->tmap_qn_n16 != score_tn_n16(nq)) {
        rc = score_tn_encode_q_map(&ix->tmap_qn, w.q_bf16, nq_pad, ix->Dp, nq);
        if (rc) return rc;
        ix->tmap_qn_ptr = w.q_bf16; ix->tmap_qn_rows = nq_pad; ix->tmap_qn_n16 = score_tn_n16(nq);
      }
      ta.tmap_q = &ix->tmap_qn; ta.tmap_db = &ix->tmap_db2;
      ta.plan = score_tn_plan(nq, ix->ntotal, ix->n_sm);
      rc = launch_score_tn(ta, st);
      if (rc) return rc;
      ++launches;
      path = 4;
      n_lists_used = ta.plan.n_splits;
      split_tile_rows = score_tn_tile_rows();
    } else {
      rc = pair ? launch_score_tc2(ta, st) : launch_score_tc(ta, st);
      if (rc) return rc;
      ++launches;
      path = pair ? 3 : 2;
      n_lists_used = ta.plan.n_splits;
      split_tile_rows = score_tc_tile_rows();
    }
    eps = w.eps_tc;
  }
  B2K_CUDA(cudaEventRecord(ev[1], st));
  // the select step reads lists [0, n_lists_used) of the stride-n_lists layout
  SelectArgs se;
  se.partial = w.partial; se.n_lists = n_lists_used; se.list_stride = w.n_lists; se.k = k; se.eps = eps;
  se.cand_cap = w.cand_cap; se.force_exact = ix->opt_force_exact;
  se.cand_rows = w.cand_rows; se.cand_count = w.cand_count; se.flags = w.flags; se.thr = w.thr;
  // tightening trades 2 x k latency-bound row gathers per query (25 us at batch 1) for ~3x fewer rows
  // to re-rank: a win once the re-rank is throughput-bound, a loss for the latency of small batches
  // (batch 128 at 10 M rows: tail 0.36 ms without, 0.15 ms with; batch 16: 0.11 ms either way)
  // (with the fused cluster tail the cross-over moved up on short shards, where few rows sit within eps of the k-th
  // score: 1.25 M rows, batch 32: tail 0.077 ms without, 0.103 with; batch 128: 0.105 vs 0.096; 10 M rows, batch 32:
  // 0.220 vs 0.205; profiles/experiments/r02_exp_tighten.log)
  const bool tighten = ix->opt_tighten > 1 ||
                       (ix->opt_tighten == 1 && (nq >= 128 || (nq >= 32 && ix->ntotal >= 4000000)));
  se.db_f32 = tighten ? ix->f32 : nullptr; se.q = q_dev; se.D = ix->D;
  se.lb = w.lb; se.sat_count = w.fail_count + 1; se.sat_pairs = ix->opt_collect ? w.sat_pairs : nullptr;
  se.sat_cap = ix->opt_collect ? w.sat_cap : 0;

  CollectArgs ca;
  ca.db = ix->bf16; ca.n_rows = ix->ntotal; ca.D = ix->D; ca.Dp = ix->Dp; ca.q = q_dev; ca.lb = w.lb;
  ca.eps = w.eps_scan; ca.sat_count = w.fail_count + 1; ca.sat_pairs = w.sat_pairs; ca.sat_cap = w.sat_cap;
  ca.n_splits = n_lists_used; ca.tile_rows = split_tile_rows;
  ca.cand_rows = w.cand_rows; ca.cand_count = w.cand_count; ca.flags = w.flags; ca.cand_cap = w.cand_cap;

  RerankArgs rr;
  rr.db_f32 = ix->f32; rr.q = q_dev; rr.cand_rows = w.cand_rows; rr.cand_count = w.cand_count;
  rr.cand_ip = w.cand_ip; rr.nq = nq; rr.cand_cap = w.cand_cap; rr.D = ix->D;

  FinalizeArgs fa;
  fa.cand_rows = w.cand_rows; fa.cand_count = w.cand_count; fa.cand_ip = w.cand_ip; fa.flags = w.flags;
  fa.qn2 = w.qn2; fa.norm2 = ix->norm2; fa.nq = nq; fa.cand_cap = w.cand_cap; fa.k = k;
  fa.base_offset = ix->base; fa.out_ip = ip_dev; fa.out_dist = dist_dev; fa.out_labels = labels_dev;
  fa.fail_count = w.fail_count; fa.fail_list = w.fail_list;

  ExactArgs ea;
  ea.db_f32 = ix->f32; ea.norm2 = ix->norm2; ea.n_rows = ix->ntotal; ea.D = ix->D; ea.q = q_dev;
  ea.qn2 = w.qn2; ea.nq = nq; ea.k = k; ea.base_offset = ix->base; ea.fail_count = w.fail_count;
  ea.fail_list = w.fail_list; ea.partial = w.exact_partial; ea.n_splits = w.exact_splits;
  ea.out_ip = ip_dev; ea.out_dist = dist_dev; ea.out_labels = labels_dev; ea.ceil_keys = w.exact_ceil;
  ea.page = 0; ea.group_done = w.fail_count + kCounterHead + w.nq_cap;

  if (ix->opt_fused_tail && k <= kList) {
    // k <= 32: ONE launch per query stage instead of three (select -> re-rank -> finalize by a cluster of CTAs per
    // query), and the two rarely needed paths finish their queries themselves (last-CTA patterns), so a search
    // is prep, score, tail, K-collect (no-op), K-exact (no-op): 5 launches, 7 with a separate sampling pass
    TailArgs tl;
    tl.se = se; tl.rr = rr; tl.fa = fa; tl.state = w.state; tl.sat_n = w.sat_n;
    rc = launch_tail(tl, nq, ix->n_sm, st, ix->opt_fused_tail != 2);
    if (rc) return rc;
    ++launches;
    if (ix->opt_collect) {
      DeferredArgs df;
      df.rr = rr; df.fa = fa; df.state = w.state; df.sat_n = w.sat_n; df.done = w.fail_count + kCounterHead;
      rc = launch_collect_finish(ca, df, ix->n_sm, st);
      if (rc) return rc;
      ++launches;
    }
    rc = launch_exact_fused(ea, st);
    if (rc) return rc;
    ++launches;
  } else {
    rc = launch_select(se, nq, st);
    if (rc) return rc;
    ++launches;
    if (ix->opt_collect) {
      rc = launch_collect(ca, ix->n_sm, st);
      if (rc) return rc;
      ++launches;
    }
    rc = launch_rerank(rr, ix->n_sm, st);
    if (rc) return rc;
    ++launches;
    rc = launch_finalize(fa, st);
    if (rc) return rc;
    ++launches;
    // the exhaustive scan keeps 32 results per pass: k > 32 is served in pages, each page scanning for
    // the rows strictly after the last result of the page before
    for (int page = 0; page * kList < k; ++page) {
      ea.page = page;
      rc = launch_exact(ea, st);
      if (rc) return rc;
      launches += 2;
    }
  }
  B2K_CUDA(cudaEventRecord(ev[2], st));
  ix->ev_head += 1;

  ix->stats.path = path;
  ix->stats.n_queries = nq;
  ix->stats.n_uncertified = -1;      // on the device until read back
  ix->stats.n_splits = n_lists_used;
  ix->stats.cand_slots = w.cand_cap;
  ix->stats.launches = launches;
  return 0;
}

int check_search_args(const b2k_index* ix, const void* q, int nq, int k, const void* dist, const void* labels) {
  if (!ix || !q || !dist || !labels || nq <= 0 || k <= 0) { set_error("search: bad argument"); return B2K_E_INVALID; }
  if (k > B2K_MAX_K) { set_error("search: k=%d exceeds B2K_MAX_K=%d", k, B2K_MAX_K); return B2K_E_INVALID; }
  if (k > kList && ix->opt_cand_cap > 0) { set_error("search: B2K_OPT_RERANK budgets apply to k <= %d only", (int)kList); return B2K_E_INVALID; }
  return 0;
}

// ---- file format -------------------------------------------------------------------------
struct FileHeader {
  char magic[8];            // "B2KIDX01"
  int32_t version, n_tables;
  int32_t dims[B2K_MAX_TABLES];
  int32_t D, has_ids;
  int64_t n_rows;
  int64_t rows_offset, ids_offset;
  char pad[48];
};
static_assert(sizeof(FileHeader) == 128, "header layout");

int read_header(FILE* f, FileHeader& h, const char* path) {
  if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "B2KIDX01", 8) != 0 || h.version != 1 ||
      h.n_tables < 1 || h.n_tables > B2K_MAX_TABLES) {
    set_error("%s: not a b2k index file", path);
    return B2K_E_IO;
  }
  return 0;
}

}  // namespace

// =========================================================================================
extern "C" {

const char* b2k_last_error(void) { return g_err; }
int32_t b2k_abi_version(void) { return B2K_ABI_VERSION; }

int b2k_device_count(int32_t* n) {
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) { *n = 0; set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e)); return B2K_E_NODEVICE; }
  *n = c;
  return 0;
}

int b2k_create(const int32_t* table_dims, int32_t n_tables, int64_t capacity_rows, int32_t device,
               int64_t base_offset, b2k_index** out) {
  if (!out) { set_error("create: out is null"); return B2K_E_INVALID; }
  *out = nullptr;
  if (!table_dims || n_tables < 1 || n_tables > B2K_MAX_TABLES || capacity_rows < 0 ||
      capacity_rows > 0x7fffff00ll) {
    set_error("create: bad argument (n_tables=%d capacity=%lld)", n_tables, (long long)capacity_rows);
    return B2K_E_INVALID;
  }
  for (int t = 0; t < n_tables; ++t)
    if (table_dims[t] < 1) { set_error("create: table %d has dim %d", t, table_dims[t]); return B2K_E_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    set_error("create: no CUDA device (this engine has no CPU path)");
    return B2K_E_NODEVICE;
  }
  if (device < 0 || device >= ndev) { set_error("create: device %d of %d", device, ndev); return B2K_E_INVALID; }
  cudaDeviceProp prop;
  B2K_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    return B2K_E_NODEVICE;
  }
  DeviceGuard g(device);
  b2k_index* ix = new (std::nothrow) b2k_index();
  if (!ix) { set_error("create: out of host memory"); return B2K_E_NOMEM; }
  ix->device = device; ix->n_sm = prop.multiProcessorCount; ix->n_tables = n_tables;
  for (int t = 0; t < n_tables; ++t) ix->dims[t] = table_dims[t];
  make_col_offsets(ix);
  ix->cap = capacity_rows; ix->base = base_offset;
  memset(&ix->stats, 0, sizeof(ix->stats));
  int rc = 0;
  cudaError_t e;
  if ((e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking)) != cudaSuccess) rc = (int)e;
  if (!rc) rc = dev_alloc(&ix->f32, (size_t)ix->cap * ix->D);
  if (!rc) rc = dev_alloc(&ix->bf16, (size_t)ix->cap * ix->Dp);
  if (!rc) rc = dev_alloc(&ix->norm2, (size_t)ix->cap);
  if (!rc) rc = dev_alloc(&ix->stat_bits, 2);
  if (!rc && (e = cudaMemset(ix->stat_bits, 0, 2 * sizeof(unsigned int))) != cudaSuccess) rc = (int)e;
  if (!rc && (e = cudaMallocHost(reinterpret_cast<void**>(&ix->h_fail), 64)) != cudaSuccess) rc = (int)e;
  for (int i = 0; i < kEvRing * 3 && !rc; ++i)
    if ((e = cudaEventCreate(&ix->ev[i / 3][i % 3])) != cudaSuccess) rc = (int)e;
  if (rc) {
    if (rc > 0) set_error("create: %s (capacity %lld rows x %d dims)", cudaGetErrorString((cudaError_t)rc),
                          (long long)capacity_rows, ix->D);
    cudaGetLastError();
    b2k_destroy(ix);
    return rc;
  }
  *out = ix;
  return 0;
}

void b2k_destroy(b2k_index* ix) {
  if (!ix) return;
  DeviceGuard g(ix->device);
  if (ix->stream) cudaStreamSynchronize(ix->stream);
  ix->ws.release();
  dev_free(ix->f32); dev_free(ix->bf16); dev_free(ix->norm2); dev_free(ix->stat_bits);
  for (int t = 0; t < B2K_MAX_TABLES; ++t) dev_free(ix->stage[t]);
  dev_free(ix->stage_rows_f32);
  b2k_stage_close(ix);
  if (ix->h_fail) cudaFreeHost(ix->h_fail);
  for (int i = 0; i < kEvRing * 3; ++i) if (ix->ev[i / 3][i % 3]) cudaEventDestroy(ix->ev[i / 3][i % 3]);
  if (ix->stream) cudaStreamDestroy(ix->stream);
  delete ix;
}

int64_t b2k_capacity(const b2k_index* ix) { return ix ? ix->cap : 0; }

int b2k_reset(b2k_index* ix) {
  if (!ix) { set_error("reset: null index"); return B2K_E_INVALID; }
  DeviceGuard g(ix->device);
  B2K_CUDA(cudaDeviceSynchronize());
  B2K_CUDA(cudaMemset(ix->stat_bits, 0, 2 * sizeof(unsigned int)));
  ix->ntotal = 0;
  ix->tmap_db_ptr = nullptr;
  return 0;
}

int b2k_reserve(b2k_index* ix, int64_t capacity_rows) {
  if (!ix || capacity_rows < ix->ntotal || capacity_rows > 0x7fffff00ll) { set_error("reserve: bad capacity"); return B2K_E_INVALID; }
  if (capacity_rows == ix->cap) return 0;
  DeviceGuard g(ix->device);
  B2K_CUDA(cudaDeviceSynchronize());
  float* nf = nullptr; uint16_t* nb = nullptr; float* nn = nullptr;
  int rc = dev_alloc(&nf, (size_t)capacity_rows * ix->D);
  if (!rc) rc = dev_alloc(&nb, (size_t)capacity_rows * ix->Dp);
  if (!rc) rc = dev_alloc(&nn, (size_t)capacity_rows);
  cudaError_t e = cudaSuccess;
  if (!rc && ix->ntotal > 0) {
    e = cudaMemcpy(nf, ix->f32, (size_t)ix->ntotal * ix->D * 4, cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(nb, ix->bf16, (size_t)ix->ntotal * ix->Dp * 2, cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(nn, ix->norm2, (size_t)ix->ntotal * 4, cudaMemcpyDeviceToDevice);
    if (e != cudaSuccess) { set_error("reserve: %s", cudaGetErrorString(e)); rc = (int)e; }
  }
  if (rc) { dev_free(nf); dev_free(nb); dev_free(nn); cudaGetLastError(); return rc; }
  dev_free(ix->f32); dev_free(ix->bf16); dev_free(ix->norm2);
  ix->f32 = nf; ix->bf16 = nb; ix->norm2 = nn; ix->cap = capacity_rows;
  ix->tmap_db_ptr = nullptr;
  return 0;
}

int b2k_add_device(b2k_index* ix, const float* const* dev_tables, int64_t n, void* stream) {
  if (!ix || !dev_tables || n < 0) { set_error("add: bad argument"); return B2K_E_INVALID; }
  if (ix->ntotal + n > ix->cap) {
    set_error("add: %lld + %lld rows exceed the capacity %lld", (long long)ix->ntotal, (long long)n, (long long)ix->cap);
    return B2K_E_CAPACITY;
  }
  if (n == 0) return 0;
  DeviceGuard g(ix->device);
  PackArgs a;
  fill_pack_args(ix, a, n, ix->ntotal, 1);
  for (int t = 0; t < ix->n_tables; ++t) a.tables[t] = dev_tables[t];
  int rc = launch_pack(a, (cudaStream_t)stream);
  if (rc) return rc;
  ix->ntotal += n;
  return 0;
}

int b2k_add(b2k_index* ix, const float* const* host_tables, int64_t n) {
  if (!ix || !host_tables || n < 0) { set_error("add: bad argument"); return B2K_E_INVALID; }
  if (ix->ntotal + n > ix->cap) {
    set_error("add: %lld + %lld rows exceed the capacity %lld", (long long)ix->ntotal, (long long)n, (long long)ix->cap);
    return B2K_E_CAPACITY;
  }
  DeviceGuard g(ix->device);
  int rc = ensure_stage(ix);
  if (rc) return rc;
  for (int64_t r0 = 0; r0 < n; r0 += ix->stage_rows) {
    const int64_t m = std::min(ix->stage_rows, n - r0);
    const float* devp[B2K_MAX_TABLES];
    for (int t = 0; t < ix->n_tables; ++t) {
      B2K_CUDA(cudaMemcpyAsync(ix->stage[t], host_tables[t] + r0 * ix->dims[t], (size_t)m * ix->dims[t] * sizeof(float),
                               cudaMemcpyHostToDevice, ix->stream));
      devp[t] = ix->stage[t];
    }
    rc = b2k_add_device(ix, devp, m, ix->stream);
    if (rc) return rc;
    // the staging buffers are reused by the next chunk and the host arrays are borrowed
    B2K_CUDA(cudaStreamSynchronize(ix->stream));
  }
  return 0;
}

// ---- pinned two-slot ingest staging --------------------------------------------------------
int b2k_stage_open(b2k_index* ix, int64_t rows_per_slot) { return b2k_stage_open_n(ix, rows_per_slot, 2); }

int b2k_stage_open_n(b2k_index* ix, int64_t rows_per_slot, int32_t n_slots) {
  if (!ix || rows_per_slot < 1 || n_slots < 2 || n_slots > kMaxStageSlots) { set_error("stage_open: bad argument (2..%d slots)", kMaxStageSlots); return B2K_E_INVALID; }
  DeviceGuard g(ix->device);
  b2k_stage_close(ix);
  int rc = ensure_stage(ix);
  if (rc) return rc;
  rows_per_slot = std::min(rows_per_slot, ix->stage_rows);
  ix->pin_slots = n_slots;
  // ONE pinned allocation for every slot and table (pinning and un-pinning cost per call and per page: 48 separate
  // 32 MB buffers took 0.3 s to allocate and up to 1.4 s to free); each table part starts 256-byte aligned
  size_t off = 0, part_off[B2K_MAX_TABLES];
  for (int t = 0; t < ix->n_tables; ++t) {
    part_off[t] = off;
    off += ((size_t)rows_per_slot * ix->dims[t] * sizeof(float) + 255) & ~(size_t)255;
  }
  const size_t slot_bytes = off;
  B2K_CUDA(cudaMallocHost(reinterpret_cast<void**>(&ix->pin_base), slot_bytes * (size_t)n_slots));
  for (int s = 0; s < n_slots; ++s) {
    for (int t = 0; t < ix->n_tables; ++t)
      ix->pin[s][t] = reinterpret_cast<float*>(ix->pin_base + (size_t)s * slot_bytes + part_off[t]);
    B2K_CUDA(cudaEventCreateWithFlags(&ix->pin_ev[s], cudaEventDisableTiming));
    ix->pin_busy[s] = false;
  }
  ix->pin_rows = rows_per_slot;
  return 0;
}

int64_t b2k_stage_rows(const b2k_index* ix) { return ix ? ix->pin_rows : 0; }

int b2k_stage_ptr(b2k_index* ix, int32_t slot, int32_t table, float** host_ptr) {
  if (!ix || !host_ptr || slot < 0 || slot >= ix->pin_slots || table < 0 || table >= ix->n_tables || ix->pin_rows == 0) {
    set_error("stage_ptr: bad argument (stage open?)");
    return B2K_E_INVALID;
  }
  *host_ptr = ix->pin[slot][table];
  return 0;
}

int b2k_stage_wait(b2k_index* ix, int32_t slot) {
  if (!ix || slot < 0 || slot >= ix->pin_slots) { set_error("stage_wait: bad argument"); return B2K_E_INVALID; }
  if (ix->pin_busy[slot]) {
    DeviceGuard g(ix->device);
    B2K_CUDA(cudaEventSynchronize(ix->pin_ev[slot]));
    ix->pin_busy[slot] = false;
  }
  return 0;
}

int b2k_stage_commit(b2k_index* ix, int32_t slot, int64_t n) {
  if (!ix || slot < 0 || slot >= ix->pin_slots || n < 0 || n > ix->pin_rows) { set_error("stage_commit: bad argument"); return B2K_E_INVALID; }
  if (ix->ntotal + n > ix->cap) {
    set_error("stage_commit: %lld + %lld rows exceed the capacity %lld", (long long)ix->ntotal, (long long)n, (long long)ix->cap);
    return B2K_E_CAPACITY;
  }
  if (n == 0) return 0;
  DeviceGuard g(ix->device);
  // one device staging set: copy(i+1) is ordered after pack(i) on the same stream; the overlap that
  // matters is the caller's decode of the other slot against this slot's DMA + pack
  const float* devp[B2K_MAX_TABLES];
  for (int t = 0; t < ix->n_tables; ++t) {
    B2K_CUDA(cudaMemcpyAsync(ix->stage[t], ix->pin[slot][t], (size_t)n * ix->dims[t] * sizeof(float),
                             cudaMemcpyHostToDevice, ix->stream));
    devp[t] = ix->stage[t];
  }
  int rc = b2k_add_device(ix, devp, n, ix->stream);
  if (rc) return rc;
  B2K_CUDA(cudaEventRecord(ix->pin_ev[slot], ix->stream));
  ix->pin_busy[slot] = true;
  return 0;
}

int b2k_stage_close(b2k_index* ix) {
  if (!ix) return 0;
  DeviceGuard g(ix->device);
  if (ix->stream) cudaStreamSynchronize(ix->stream);
  for (int s = 0; s < kMaxStageSlots; ++s) {
    for (int t = 0; t < B2K_MAX_TABLES; ++t) ix->pin[s][t] = nullptr;
    if (ix->pin_ev[s]) cudaEventDestroy(ix->pin_ev[s]);
    ix->pin_ev[s] = nullptr; ix->pin_busy[s] = false;
  }
  if (ix->pin_base) cudaFreeHost(ix->pin_base);
  ix->pin_base = nullptr;
  ix->pin_rows = 0;
  ix->pin_slots = 0;
  return 0;
}

int64_t b2k_ntotal(const b2k_index* ix) { return ix ? ix->ntotal : 0; }
int32_t b2k_dim(const b2k_index* ix) { return ix ? ix->D : 0; }
int32_t b2k_table_dims(const b2k_index* ix, int32_t* dims_out) {
  if (!ix) return 0;
  if (dims_out) for (int t = 0; t < ix->n_tables; ++t) dims_out[t] = ix->dims[t];
  return ix->n_tables;
}
int32_t b2k_dim_padded(const b2k_index* ix) { return ix ? ix->Dp : 0; }
int64_t b2k_base_offset(const b2k_index* ix) { return ix ? ix->base : 0; }

int b2k_search_device(b2k_index* ix, const float* q_dev, int32_t nq, int32_t k, float* dist_dev,
                      int64_t* labels_dev, float* ip_dev, void* stream) {
  int rc = check_search_args(ix, q_dev, nq, k, dist_dev, labels_dev);
  if (rc) return rc;
  DeviceGuard g(ix->device);
  cudaStream_t st = (cudaStream_t)stream;      // NULL = the legacy default stream, as given
  rc = ensure_workspace(ix, std::min(nq, kMaxNqPerPass), k);
  if (rc) return rc;
  for (int q0 = 0; q0 < nq; q0 += kMaxNqPerPass) {
    const int m = std::min(kMaxNqPerPass, nq - q0);
    rc = search_pass(ix, q_dev + (int64_t)q0 * ix->D, m, k, dist_dev + (int64_t)q0 * k,
                     labels_dev + (int64_t)q0 * k, ip_dev ? ip_dev + (int64_t)q0 * k : nullptr, st);
    if (rc) return rc;
  }
  return 0;
}

int b2k_search(b2k_index* ix, const float* q_host, int32_t nq, int32_t k, float* dist_host,
               int64_t* labels_host, float* ip_host) {
  int rc = check_search_args(ix, q_host, nq, k, dist_host, labels_host);
  if (rc) return rc;
  DeviceGuard g(ix->device);
  rc = ensure_workspace(ix, std::min(nq, kMaxNqPerPass), k);
  if (rc) return rc;
  Workspace& w = ix->ws;
  cudaStream_t st = ix->stream;
  int n_fail = 0;
  for (int q0 = 0; q0 < nq; q0 += kMaxNqPerPass) {
    const int m = std::min(kMaxNqPerPass, nq - q0);
    B2K_CUDA(cudaMemcpyAsync(w.q, q_host + (int64_t)q0 * ix->D, (size_t)m * ix->D * sizeof(float),
                             cudaMemcpyHostToDevice, st));
    rc = search_pass(ix, w.q, m, k, w.out_dist, w.out_labels, w.out_ip, st);
    if (rc) return rc;
    B2K_CUDA(cudaMemcpyAsync(dist_host + (int64_t)q0 * k, w.out_dist, (size_t)m * k * sizeof(float),
                             cudaMemcpyDeviceToHost, st));
    B2K_CUDA(cudaMemcpyAsync(labels_host + (int64_t)q0 * k, w.out_labels, (size_t)m * k * sizeof(int64_t),
                             cudaMemcpyDeviceToHost, st));
    if (ip_host)
      B2K_CUDA(cudaMemcpyAsync(ip_host + (int64_t)q0 * k, w.out_ip, (size_t)m * k * sizeof(float),
                               cudaMemcpyDeviceToHost, st));
    B2K_CUDA(cudaMemcpyAsync(ix->h_fail, w.fail_count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    B2K_CUDA(cudaStreamSynchronize(st));
    n_fail += ix->h_fail[0];
  }
  ix->stats.n_uncertified = n_fail;
  return 0;
}

int b2k_prep_groups_device(const float* parts_dev, const int32_t* group_offsets_dev, int32_t n_groups, int32_t d,
                           float* q_dev, int32_t device, void* stream) {
  if (!parts_dev || !group_offsets_dev || !q_dev || n_groups < 0 || d < 1) { set_error("prep_groups: bad argument"); return B2K_E_INVALID; }
  DeviceGuard g(device);
  return launch_group_prep(parts_dev, group_offsets_dev, n_groups, d, q_dev, (cudaStream_t)stream);
}

int b2k_search_groups(b2k_index* ix, const float* parts_host, int64_t n_images, const int32_t* group_offsets,
                      int32_t n_groups, int32_t k, float* dist_host, int64_t* labels_host, float* ip_host) {
  int rc = check_search_args(ix, parts_host, n_groups, k, dist_host, labels_host);
  if (rc) return rc;
  if (!group_offsets || n_images < n_groups || group_offsets[0] != 0 || group_offsets[n_groups] != n_images) {
    set_error("search_groups: offsets must run from 0 to n_images (%lld) with no empty group", (long long)n_images);
    return B2K_E_INVALID;
  }
  for (int g = 0; g < n_groups; ++g)
    if (group_offsets[g + 1] <= group_offsets[g]) { set_error("search_groups: group %d is empty", g); return B2K_E_INVALID; }
  DeviceGuard g(ix->device);
  Workspace& w = ix->ws;
  cudaStream_t st = ix->stream;
  int n_fail = 0;
  for (int g0 = 0; g0 < n_groups; g0 += kMaxNqPerPass) {
    const int m = std::min(kMaxNqPerPass, n_groups - g0);
    rc = ensure_workspace(ix, m, k);
    if (rc) return rc;
    const int64_t r0 = group_offsets[g0], r1 = group_offsets[g0 + m];
    if (r1 - r0 > w.parts_cap) {
      B2K_CUDA(cudaStreamSynchronize(st));
      dev_free(w.parts);
      w.parts_cap = 0;
      if ((rc = dev_alloc(&w.parts, (size_t)(r1 - r0) * ix->D))) return rc;
      w.parts_cap = r1 - r0;
    }
    // offsets of this pass, rebased to its first image (the host array is borrowed: staged synchronously)
    std::vector<int32_t> offs((size_t)m + 1);
    for (int i = 0; i <= m; ++i) offs[i] = (int32_t)(group_offsets[g0 + i] - r0);
    B2K_CUDA(cudaMemcpyAsync(w.parts, parts_host + r0 * ix->D, (size_t)(r1 - r0) * ix->D * sizeof(float), cudaMemcpyHostToDevice, st));
    B2K_CUDA(cudaMemcpyAsync(w.goffs, offs.data(), offs.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    B2K_CUDA(cudaStreamSynchronize(st));                 // offs goes out of scope; pageable copies are staged anyway
    // mean over the group's images + whole-vector normalise, straight into the search workspace
    rc = launch_group_prep(w.parts, w.goffs, m, ix->D, w.q, st);
    if (rc) return rc;
    rc = search_pass(ix, w.q, m, k, w.out_dist, w.out_labels, w.out_ip, st);
    if (rc) return rc;
    B2K_CUDA(cudaMemcpyAsync(dist_host + (int64_t)g0 * k, w.out_dist, (size_t)m * k * sizeof(float), cudaMemcpyDeviceToHost, st));
    B2K_CUDA(cudaMemcpyAsync(labels_host + (int64_t)g0 * k, w.out_labels, (size_t)m * k * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    if (ip_host)
      B2K_CUDA(cudaMemcpyAsync(ip_host + (int64_t)g0 * k, w.out_ip, (size_t)m * k * sizeof(float), cudaMemcpyDeviceToHost, st));
    B2K_CUDA(cudaMemcpyAsync(ix->h_fail, w.fail_count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    B2K_CUDA(cudaStreamSynchronize(st));
    n_fail += ix->h_fail[0];
  }
  ix->stats.n_uncertified = n_fail;
  return 0;
}

int b2k_get_stats(b2k_index* ix, b2k_stats* out) {
  if (!ix || !out) { set_error("get_stats: bad argument"); return B2K_E_INVALID; }
  DeviceGuard g(ix->device);
  B2K_CUDA(cudaDeviceSynchronize());
  unsigned int bits[2] = {0, 0};
  B2K_CUDA(cudaMemcpy(bits, ix->stat_bits, sizeof(bits), cudaMemcpyDeviceToHost));
  float e2, n2;
  memcpy(&e2, &bits[0], 4); memcpy(&n2, &bits[1], 4);
  ix->stats.err_max = sqrtf(e2);
  ix->stats.norm_max = sqrtf(n2);
  if (ix->ws.fail_count && ix->stats.path != 0 && ix->stats.n_uncertified < 0) {
    // device API: count of the last pass (the host API sums over passes itself)
    int32_t nf = 0;
    B2K_CUDA(cudaMemcpy(&nf, ix->ws.fail_count, sizeof(nf), cudaMemcpyDeviceToHost));
    ix->stats.n_uncertified = nf;
  }
  if (ix->ws.eps_tc && ix->stats.path != 0) {
    float eps0 = 0.f;
    B2K_CUDA(cudaMemcpy(&eps0, ix->stats.path == 1 ? ix->ws.eps_scan : ix->ws.eps_tc, sizeof(float),
                        cudaMemcpyDeviceToHost));
    ix->stats.eps_max = eps0;   // slack of query 0 of the last pass (representative; per-query on device)
  }
  if (ix->ws.fail_count && ix->stats.path != 0) {
    int32_t ns = 0;
    B2K_CUDA(cudaMemcpy(&ns, ix->ws.fail_count + 1, sizeof(ns), cudaMemcpyDeviceToHost));
    ix->stats.n_saturated = ns;
  }
  if (ix->ws.cand_count && ix->stats.path != 0 && ix->stats.n_queries > 0) {
    std::vector<int32_t> cnt((size_t)ix->stats.n_queries);
    B2K_CUDA(cudaMemcpy(cnt.data(), ix->ws.cand_count, cnt.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    int64_t tot = 0;
    for (int32_t c : cnt) tot += std::min(c, ix->ws.cand_cap);
    ix->stats.n_candidates = (int32_t)std::min<int64_t>(tot, 0x7fffffff);
  }
  if (ix->ev_head > ix->ev_read) {
    // mean over the passes recorded since the last call (at most the ring's depth, the most recent ones)
    const int64_t first = std::max(ix->ev_read, ix->ev_head - kEvRing);
    double sa = 0.0, sb = 0.0;
    int n = 0;
    for (int64_t i = first; i < ix->ev_head; ++i) {
      float a = 0.f, b = 0.f;
      cudaEvent_t* ev = ix->ev[i % kEvRing];
      if (cudaEventElapsedTime(&a, ev[0], ev[1]) == cudaSuccess && cudaEventElapsedTime(&b, ev[1], ev[2]) == cudaSuccess) {
        sa += a; sb += b; ++n;
      }
    }
    cudaGetLastError();
    if (n > 0) { ix->stats.score_ms = (float)(sa / n); ix->stats.tail_ms = (float)(sb / n); }
    ix->stats.n_timed = n;
    ix->ev_read = ix->ev_head;
  }
  *out = ix->stats;
  return 0;
}

int b2k_set_option(b2k_index* ix, int32_t key, int64_t value) {
  if (!ix) { set_error("set_option: null index"); return B2K_E_INVALID; }
  switch (key) {
    case B2K_OPT_PATH:
      if (value < 0 || value > 2) break;
      ix->opt_path = (int)value; return 0;
    case B2K_OPT_RERANK:
      if (value != 0 && (value < 32 || value > 8192)) break;
      ix->opt_cand_cap = (int)value; return 0;
    case B2K_OPT_FORCE_EXACT:
      ix->opt_force_exact = value != 0; return 0;
    case B2K_OPT_SCAN_MAX_B:
      if (value < 0) break;
      ix->opt_scan_max_b = (int)value; return 0;
    case B2K_OPT_SPLITS:
      if (value < 0 || value > 4096) break;
      ix->opt_splits = (int)value; return 0;
    case B2K_OPT_TIGHTEN:
      if (value < 0 || value > 2) break;
      ix->opt_tighten = (int)value; return 0;
    case B2K_OPT_COLLECT:
      ix->opt_collect = value != 0; return 0;
    case B2K_OPT_INLINE_SEED:
      ix->opt_inline_seed = value != 0; return 0;
    case B2K_OPT_FUSED_TAIL:
      if (value < 0 || value > 2) break;
      ix->opt_fused_tail = (int)value; return 0;
    case B2K_OPT_TN:
      if (value < -1 || value > 1) break;
      ix->opt_tn = (int)value; return 0;
    case B2K_OPT_SAMPLE_WAVE:
      ix->opt_sample_wave = value != 0; return 0;
    case B2K_OPT_SEED:
      if (value < 0 || value > 4096) break;
      ix->opt_seed = (int)value; return 0;
    case B2K_OPT_TC_PAIR:
      if (value < -1 || value > 1) break;
      ix->opt_pair = (int)value; return 0;
    default: break;
  }
  set_error("set_option: key %d value %lld rejected", key, (long long)value);
  return B2K_E_INVALID;
}

int b2k_merge_topk_device(const float* ip, const float* dist, const int64_t* labels, int32_t n_lists,
                          int32_t nq, int32_t k, float* out_ip, float* out_dist, int64_t* out_labels,
                          int32_t device, void* stream) {
  if (!ip || !dist || !labels || !out_dist || !out_labels || n_lists < 1 || nq < 0 || k < 1 || k > B2K_MAX_K ||
      (n_lists > 32 && n_lists * k > 1024)) {
    set_error("merge: bad argument");
    return B2K_E_INVALID;
  }
  DeviceGuard g(device);
  MergeArgs a;
  a.ip = ip; a.dist = dist; a.labels = labels; a.n_lists = n_lists; a.nq = nq; a.k = k;
  a.out_ip = out_ip; a.out_dist = out_dist; a.out_labels = out_labels;
  return launch_merge(a, (cudaStream_t)stream);
}

int b2k_normalize_l2(float* x_host, int64_t n, int32_t d, int32_t device) {
  if (!x_host || n < 0 || d < 1) { set_error("normalize_l2: bad argument"); return B2K_E_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    set_error("normalize_l2: no CUDA device (this engine has no CPU path)");
    return B2K_E_NODEVICE;
  }
  if (n == 0) return 0;
  DeviceGuard g(device);
  float* d_x = nullptr;
  int rc = dev_alloc(&d_x, (size_t)n * d);
  if (rc) return rc;
  cudaError_t e = cudaMemcpy(d_x, x_host, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) { rc = launch_normalize(d_x, n, d, 0); if (!rc) e = cudaDeviceSynchronize(); }
  if (e == cudaSuccess && !rc) e = cudaMemcpy(x_host, d_x, (size_t)n * d * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(d_x);
  if (e != cudaSuccess) { set_error("normalize_l2: %s", cudaGetErrorString(e)); return (int)e; }
  return rc;
}

// ---- persistence -------------------------------------------------------------------------
namespace {
// Rows [0, ntotal) of the shard -> file rows [file_row_begin, ...) of a file laid out for file_total_rows rows.
// device -> pinned slot (async D2H) -> fwrite; the copy of chunk i+1 overlaps the write of chunk i.
int write_shard_rows(b2k_index* ix, FILE* f, const char* path, const int64_t* ids, int64_t file_row_begin,
                     int64_t file_total_rows) {
  const int64_t rows_offset = sizeof(FileHeader);
  const int64_t ids_offset = rows_offset + file_total_rows * (int64_t)ix->D * 4;
  bool ok = fseeko(f, (off_t)(rows_offset + file_row_begin * (int64_t)ix->D * 4), SEEK_SET) == 0;
  const int64_t chunk = std::max<int64_t>(256, std::min<int64_t>(ix->stage_rows, ((int64_t)1 << 26) / ((int64_t)ix->D * 4)));
  float* pin[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaError_t e = cudaSuccess;
  for (int s = 0; s < 2 && e == cudaSuccess; ++s) {
    e = cudaMallocHost(reinterpret_cast<void**>(&pin[s]), (size_t)chunk * ix->D * 4);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[s], cudaEventDisableTiming);
  }
  auto copy_chunk = [&](int64_t r0, int s) {
    const int64_t m = std::min(chunk, ix->ntotal - r0);
    cudaError_t c = cudaMemcpyAsync(pin[s], ix->f32 + r0 * ix->D, (size_t)m * ix->D * 4, cudaMemcpyDeviceToHost, ix->stream);
    return c == cudaSuccess ? cudaEventRecord(ev[s], ix->stream) : c;
  };
  if (e == cudaSuccess && ix->ntotal > 0) e = copy_chunk(0, 0);
  int slot = 0;
  for (int64_t r0 = 0; ok && e == cudaSuccess && r0 < ix->ntotal; r0 += chunk, slot ^= 1) {
    const int64_t m = std::min(chunk, ix->ntotal - r0);
    if (r0 + chunk < ix->ntotal) e = copy_chunk(r0 + chunk, slot ^ 1);
    if (e == cudaSuccess) e = cudaEventSynchronize(ev[slot]);
    if (e == cudaSuccess) ok = fwrite(pin[slot], (size_t)ix->D * 4, (size_t)m, f) == (size_t)m;
  }
  cudaStreamSynchronize(ix->stream);
  for (int s = 0; s < 2; ++s) { if (pin[s]) cudaFreeHost(pin[s]); if (ev[s]) cudaEventDestroy(ev[s]); }
  if (e != cudaSuccess) { set_error("save: %s", cudaGetErrorString(e)); return (int)e; }
  if (ok && ids && ix->ntotal > 0)
    ok = fseeko(f, (off_t)(ids_offset + file_row_begin * 8), SEEK_SET) == 0 &&
         fwrite(ids, 8, (size_t)ix->ntotal, f) == (size_t)ix->ntotal;
  if (!ok) { set_error("save: write to %s failed", path); return B2K_E_IO; }
  return 0;
}

void fill_header(const b2k_index* ix, FileHeader& h, int64_t total_rows, bool has_ids) {
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, "B2KIDX01", 8);
  h.version = 1; h.n_tables = ix->n_tables;
  for (int t = 0; t < ix->n_tables; ++t) h.dims[t] = ix->dims[t];
  h.D = ix->D; h.has_ids = has_ids ? 1 : 0; h.n_rows = total_rows;
  h.rows_offset = sizeof(FileHeader);
  h.ids_offset = h.rows_offset + total_rows * (int64_t)ix->D * 4;
}
}  // namespace

int b2k_save(b2k_index* ix, const char* path, const int64_t* ids, int64_t n_ids) {
  if (!ix || !path) { set_error("save: bad argument"); return B2K_E_INVALID; }
  if (ids && n_ids != ix->ntotal) { set_error("save: %lld ids for %lld rows", (long long)n_ids, (long long)ix->ntotal); return B2K_E_INVALID; }
  DeviceGuard g(ix->device);
  B2K_CUDA(cudaStreamSynchronize(ix->stream));
  FILE* f = fopen(path, "wb");
  if (!f) { set_error("save: cannot open %s", path); return B2K_E_IO; }
  FileHeader h;
  fill_header(ix, h, ix->ntotal, ids != nullptr);
  int rc = fwrite(&h, sizeof(h), 1, f) == 1 ? 0 : B2K_E_IO;
  if (rc) set_error("save: write to %s failed", path);
  if (!rc) rc = write_shard_rows(ix, f, path, ids, 0, ix->ntotal);
  if (fclose(f) != 0 && !rc) { set_error("save: write to %s failed", path); rc = B2K_E_IO; }
  return rc;
}

int b2k_save_shard(b2k_index* ix, const char* path, const int64_t* ids, int64_t file_row_begin, int64_t file_total_rows,
                   int32_t create) {
  if (!ix || !path || file_row_begin < 0 || file_row_begin + ix->ntotal > file_total_rows) {
    set_error("save_shard: bad argument (rows [%lld, +%lld) of %lld)", (long long)file_row_begin,
              ix ? (long long)ix->ntotal : 0ll, (long long)file_total_rows);
    return B2K_E_INVALID;
  }
  DeviceGuard g(ix->device);
  B2K_CUDA(cudaStreamSynchronize(ix->stream));
  FILE* f = fopen(path, create ? "wb" : "r+b");
  if (!f) { set_error("save_shard: cannot open %s", path); return B2K_E_IO; }
  int rc = 0;
  if (create) {
    // the creating rank lays the whole file out (header + sparse extent); the others write into it afterwards
    FileHeader h;
    fill_header(ix, h, file_total_rows, ids != nullptr);
    const int64_t size = h.ids_offset + (ids ? file_total_rows * 8 : 0);
    if (fwrite(&h, sizeof(h), 1, f) != 1 || (size > (int64_t)sizeof(h) && (fseeko(f, (off_t)(size - 1), SEEK_SET) != 0 || fputc(0, f) == EOF))) {
      set_error("save_shard: cannot lay out %s", path);
      rc = B2K_E_IO;
    }
  } else {
    FileHeader h;
    rc = read_header(f, h, path);
    if (!rc && (h.n_rows != file_total_rows || h.D != ix->D || (h.has_ids != 0) != (ids != nullptr))) {
      set_error("save_shard: %s was laid out for another index (%lld rows, D=%d)", path, (long long)h.n_rows, h.D);
      rc = B2K_E_INVALID;
    }
  }
  if (!rc) rc = write_shard_rows(ix, f, path, ids, file_row_begin, file_total_rows);
  if (fclose(f) != 0 && !rc) { set_error("save_shard: write to %s failed", path); rc = B2K_E_IO; }
  return rc;
}

int b2k_file_info(const char* path, int64_t* n_rows, int32_t* n_tables, int32_t* table_dims, int32_t* has_ids) {
  if (!path) { set_error("file_info: null path"); return B2K_E_INVALID; }
  FILE* f = fopen(path, "rb");
  if (!f) { set_error("file_info: cannot open %s", path); return B2K_E_IO; }
  FileHeader h;
  int rc = read_header(f, h, path);
  fclose(f);
  if (rc) return rc;
  if (n_rows) *n_rows = h.n_rows;
  if (n_tables) *n_tables = h.n_tables;
  if (table_dims) for (int t = 0; t < h.n_tables; ++t) table_dims[t] = h.dims[t];
  if (has_ids) *has_ids = h.has_ids;
  return 0;
}

int b2k_load_ids(const char* path, int64_t row_begin, int64_t n, int64_t* ids_out) {
  if (!path || !ids_out || row_begin < 0 || n < 0) { set_error("load_ids: bad argument"); return B2K_E_INVALID; }
  FILE* f = fopen(path, "rb");
  if (!f) { set_error("load_ids: cannot open %s", path); return B2K_E_IO; }
  FileHeader h;
  int rc = read_header(f, h, path);
  if (!rc && (!h.has_ids || row_begin + n > h.n_rows)) { set_error("load_ids: %s holds no ids for that range", path); rc = B2K_E_IO; }
  if (!rc && (fseeko(f, (off_t)(h.ids_offset + row_begin * 8), SEEK_SET) != 0 ||
              fread(ids_out, 8, (size_t)n, f) != (size_t)n)) { set_error("load_ids: short read"); rc = B2K_E_IO; }
  fclose(f);
  return rc;
}

int b2k_load(const char* path, int32_t device, int64_t row_begin, int64_t row_end, int64_t capacity_rows,
             b2k_index** out) {
  if (!path || !out) { set_error("load: bad argument"); return B2K_E_INVALID; }
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) { set_error("load: cannot open %s", path); return B2K_E_IO; }
  FileHeader h;
  int rc = read_header(f, h, path);
  if (rc) { fclose(f); return rc; }
  if (row_end < 0 || row_end > h.n_rows) row_end = h.n_rows;
  if (row_begin < 0 || row_begin > row_end) { fclose(f); set_error("load: bad row range"); return B2K_E_INVALID; }
  const int64_t n = row_end - row_begin;
  b2k_index* ix = nullptr;
  rc = b2k_create(h.dims, h.n_tables, std::max(n, capacity_rows), device, row_begin, &ix);
  if (rc) { fclose(f); return rc; }
  if (ix->D != h.D) { fclose(f); b2k_destroy(ix); set_error("load: corrupt header"); return B2K_E_IO; }
  DeviceGuard g(device);
  // file -> pinned slot (fread) -> async H2D -> re-pack; the read of chunk i+1 overlaps the DMA and
  // the pack of chunk i (one device staging buffer: copies and packs are ordered on the stream)
  const int64_t chunk = std::max<int64_t>(256, std::min<int64_t>(ix->stage_rows, ((int64_t)1 << 26) / ((int64_t)ix->D * 4)));
  rc = dev_alloc(&ix->stage_rows_f32, (size_t)chunk * ix->D);
  float* pin[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool busy[2] = {false, false};
  for (int s = 0; s < 2 && !rc; ++s) {
    cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&pin[s]), (size_t)chunk * ix->D * 4);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[s], cudaEventDisableTiming);
    if (e != cudaSuccess) { set_error("load: %s", cudaGetErrorString(e)); rc = (int)e; }
  }
  if (!rc && fseeko(f, (off_t)(h.rows_offset + row_begin * (int64_t)ix->D * 4), SEEK_SET) != 0) { set_error("load: seek failed"); rc = B2K_E_IO; }
  int slot = 0;
  for (int64_t r0 = 0; !rc && r0 < n; r0 += chunk, slot ^= 1) {
    const int64_t m = std::min(chunk, n - r0);
    cudaError_t e = cudaSuccess;
    if (busy[slot]) { e = cudaEventSynchronize(ev[slot]); busy[slot] = false; }
    if (e == cudaSuccess && fread(pin[slot], (size_t)ix->D * 4, (size_t)m, f) != (size_t)m) { set_error("load: short read in %s", path); rc = B2K_E_IO; break; }
    if (e == cudaSuccess) e = cudaMemcpyAsync(ix->stage_rows_f32, pin[slot], (size_t)m * ix->D * 4, cudaMemcpyHostToDevice, ix->stream);
    if (e != cudaSuccess) { set_error("load: %s", cudaGetErrorString(e)); rc = (int)e; break; }
    // stored rows are already normalised: re-pack without scaling (same per-table sums -> same norm2 bits)
    PackArgs a;
    fill_pack_args(ix, a, m, r0, 0);
    for (int t = 0; t < ix->n_tables; ++t) { a.tables[t] = ix->stage_rows_f32 + ix->col_off[t]; a.strides[t] = ix->D; }
    rc = launch_pack(a, ix->stream);
    if (!rc) { e = cudaEventRecord(ev[slot], ix->stream); busy[slot] = true; if (e != cudaSuccess) { set_error("load: %s", cudaGetErrorString(e)); rc = (int)e; } }
  }
  {
    cudaError_t e = cudaStreamSynchronize(ix->stream);
    if (!rc && e != cudaSuccess) { set_error("load: %s", cudaGetErrorString(e)); rc = (int)e; }
  }
  for (int s = 0; s < 2; ++s) { if (pin[s]) cudaFreeHost(pin[s]); if (ev[s]) cudaEventDestroy(ev[s]); }
  dev_free(ix->stage_rows_f32);
  fclose(f);
  if (rc) { b2k_destroy(ix); return rc; }
  ix->ntotal = n;
  *out = ix;
  return 0;
}

int b2k_get_rows(b2k_index* ix, int64_t row0, int64_t n, float* f32_host, uint16_t* bf16_host, float* norm2_host) {
  if (!ix || row0 < 0 || n < 0 || row0 + n > ix->ntotal) { set_error("get_rows: bad range"); return B2K_E_INVALID; }
  DeviceGuard g(ix->device);
  B2K_CUDA(cudaStreamSynchronize(ix->stream));
  if (f32_host) B2K_CUDA(cudaMemcpy(f32_host, ix->f32 + row0 * ix->D, (size_t)n * ix->D * 4, cudaMemcpyDeviceToHost));
  if (bf16_host) B2K_CUDA(cudaMemcpy(bf16_host, ix->bf16 + row0 * ix->Dp, (size_t)n * ix->Dp * 2, cudaMemcpyDeviceToHost));
  if (norm2_host) B2K_CUDA(cudaMemcpy(norm2_host, ix->norm2 + row0, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return 0;
}

// ---- synthetic data ------------------------------------------------------------------------
static void fill_synth_args(const b2k_index* ix, SynthArgs& s, const b2k_synth* p) {
  memset(&s, 0, sizeof(s));
  s.n_tables = ix->n_tables;
  for (int t = 0; t < ix->n_tables; ++t) { s.dims[t] = ix->dims[t]; s.tables[t] = ix->stage[t]; }
  s.seed = p->seed; s.n_clusters = p->n_clusters > 0 ? p->n_clusters : 4096;
  s.sigma = p->sigma; s.abs_mask = p->abs_mask; s.total_rows = p->total_rows;
}

int b2k_fill_synthetic(b2k_index* ix, int64_t n, const b2k_synth* p) {
  if (!ix || !p || n < 0) { set_error("fill_synthetic: bad argument"); return B2K_E_INVALID; }
  if (ix->ntotal + n > ix->cap) { set_error("fill_synthetic: beyond capacity"); return B2K_E_CAPACITY; }
  DeviceGuard g(ix->device);
  int rc = ensure_stage(ix);
  if (rc) return rc;
  for (int64_t r0 = 0; r0 < n; r0 += ix->stage_rows) {
    const int64_t m = std::min(ix->stage_rows, n - r0);
    SynthArgs s;
    fill_synth_args(ix, s, p);
    s.query_mode = 0; s.n = m; s.first = ix->base + ix->ntotal;
    rc = launch_synth(s, ix->stream);
    if (rc) return rc;
    const float* devp[B2K_MAX_TABLES];
    for (int t = 0; t < ix->n_tables; ++t) devp[t] = ix->stage[t];
    rc = b2k_add_device(ix, devp, m, ix->stream);   // same stream: ordered after the generator
    if (rc) return rc;
  }
  B2K_CUDA(cudaStreamSynchronize(ix->stream));
  return 0;
}

int b2k_synth_queries_device(b2k_index* ix, int32_t nq, const b2k_synth* p, uint64_t qseed, float sigma_q,
                             float* q_dev, void* stream) {
  if (!ix || !p || !q_dev || nq < 0 || nq > ix->stage_rows || p->total_rows < 1) {
    set_error("synth_queries: bad argument (nq <= %lld)", (long long)(ix ? ix->stage_rows : 0));
    return B2K_E_INVALID;
  }
  DeviceGuard g(ix->device);
  int rc = ensure_stage(ix);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  SynthArgs s;
  fill_synth_args(ix, s, p);
  s.query_mode = 1; s.n = nq; s.first = 0; s.qseed = qseed; s.sigma_q = sigma_q;
  rc = launch_synth(s, st);
  if (rc) return rc;
  // parts -> unit norm (extractor output), concatenated; then whole-vector normalise
  PackArgs a;
  fill_pack_args(ix, a, nq, 0, 1);
  for (int t = 0; t < ix->n_tables; ++t) a.tables[t] = ix->stage[t];
  a.out_f32 = q_dev; a.out_bf16 = nullptr; a.out_norm2 = nullptr; a.stat_bits = nullptr;
  rc = launch_pack(a, st);
  if (rc) return rc;
  return launch_normalize(q_dev, nq, ix->D, st);
}

}  // extern "C"
