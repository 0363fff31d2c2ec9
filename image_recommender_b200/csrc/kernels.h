// Host-visible launch wrappers and argument blocks of the b2k kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

namespace b2k {

struct PackArgs {
  const float* tables[B2K_MAX_TABLES];  // device [n, d_t]
  int32_t dims[B2K_MAX_TABLES];
  int32_t strides[B2K_MAX_TABLES];  // input row stride of table t in floats (d_t when dense)
  int32_t col_off[B2K_MAX_TABLES];
  int32_t n_tables;
  int32_t D, Dp;
  int32_t normalize;      // 1: per-table L2 normalise (build); 0: copy/convert only (load)
  int64_t n;              // rows in this call
  int64_t row0;           // first output row
  float* out_f32;         // [cap, D]   (may be null)
  uint16_t* out_bf16;     // [cap, Dp]  (may be null)
  float* out_norm2;       // [cap]      (may be null)
  unsigned int* stat_bits;  // [0] max ||bf16(y)-y||² bits, [1] max ||y||² bits (may be null)
};

struct QueryPrepArgs {
  const float* q;       // [nq, D]
  uint16_t* q_bf16;     // [nq_pad, Dp]
  float* qn2;           // [nq]
  float* eps_scan;      // [nq]
  float* eps_tc;        // [nq]
  const unsigned int* stat_bits;
  int32_t nq, nq_pad, D, Dp;
  float* floor_init;    // [nq] set to -inf (the in-kernel seeding publishes floors into it); may be null
};

struct SynthArgs {
  float* tables[B2K_MAX_TABLES];
  int32_t dims[B2K_MAX_TABLES];
  int32_t n_tables;
  int32_t query_mode;
  int64_t n, first, total_rows;
  uint64_t seed, qseed;
  int32_t n_clusters;
  uint32_t abs_mask;
  float sigma, sigma_q;
};

struct ScanArgs {
  const uint16_t* db;   // bf16 [n_rows, Dp]
  int64_t n_rows;
  int32_t D, Dp;
  const float* q;       // fp32 [nq, D]
  int32_t nq;           // total queries
  int32_t q0;           // first query of this pass
  int32_t n_splits;     // == gridDim.x
  Cand* partial;        // [nq, n_lists, kList]; this kernel fills list `split`
  int32_t n_lists;
};

struct SelectArgs {
  const Cand* partial;  // [nq, list_stride, kList] approximate (bf16) scores
  int32_t n_lists;      // lists [0, n_lists) of each query are in use
  int32_t list_stride;
  int32_t k;
  const float* eps;     // [nq] bound on |approximate - exact| for the path that filled `partial`
  int32_t cand_cap;     // candidate slots per query
  int32_t force_exact;
  int32_t* cand_rows;   // [nq, cand_cap]
  int32_t* cand_count;  // [nq]
  int32_t* flags;       // [nq] 0 = certified; bit0 saturated list not handed to K-collect, bit1 overflow, bit2 forced
  float* thr;           // [nq] candidate threshold actually used
  float* lb;            // [nq] lower bound of the exact k-th best score (K-collect derives its own threshold)
  // saturated (query, list) pairs are handed to K-collect instead of failing the query (null = off)
  int32_t* sat_count;   // [1]
  int2* sat_pairs;      // [sat_cap] (query, list)
  int32_t sat_cap;
  // optional tightening: exact scores of the k rows with the best approximate scores (both null = off)
  const float* db_f32;  // [n_rows, D]
  const float* q;       // [nq, D]
  int32_t D;
};

// K-collect: re-scan of the DB splits whose partial list was saturated (every one of its 32 entries
// at or above the candidate threshold, so it may hide more): EVERY row of such a split whose
// approximate score reaches the threshold becomes a re-rank candidate.
struct CollectArgs {
  const uint16_t* db;   // bf16 [n_rows, Dp]
  int64_t n_rows;
  int32_t D, Dp;
  const float* q;       // fp32 [nq, D]
  const float* lb;      // [nq] lower bound of the exact k-th best score (from K-select)
  const float* eps;     // [nq] |fp32 query x bf16 row - exact| bound (the K-scan bound)
  const int32_t* sat_count;
  const int2* sat_pairs;
  int32_t sat_cap;
  int32_t n_splits;     // splits of the scoring pass that filled the lists
  int32_t tile_rows;    // 0: split s = rows [n*s/S, n*(s+1)/S) (K-scan); else tiles of that many rows (K-score)
  int32_t* cand_rows;   // [nq, cand_cap]
  int32_t* cand_count;  // [nq] appended to
  int32_t* flags;       // [nq] bit1 set on overflow
  int32_t cand_cap;
};

struct RerankArgs {
  const float* db_f32;  // [n_rows, D]
  const float* q;       // [nq, D]
  const int32_t* cand_rows;   // [nq, cand_cap]
  const int32_t* cand_count;  // [nq]
  float* cand_ip;       // [nq, cand_cap] exact scores (Spec R)
  int32_t nq, cand_cap, D;
};

struct FinalizeArgs {
  const int32_t* cand_rows;
  const int32_t* cand_count;
  const float* cand_ip;
  const int32_t* flags;
  const float* qn2;
  const float* norm2;     // per DB row
  int32_t nq, cand_cap, k;
  int64_t base_offset;
  float* out_ip;          // [nq, k] (may be null)
  float* out_dist;        // [nq, k]
  int64_t* out_labels;    // [nq, k]
  int32_t* fail_count;    // [1]
  int32_t* fail_list;     // [nq]
};

// Fused per-query tail (k <= 32): K-select, K-rerank and K-finalize of one query in ONE launch, by a cluster
// of 1..8 CTAs (more CTAs per query for small batches: the re-rank is a latency-bound row gather).  Queries
// that need more than that are left to the two deferred kernels, which exit at once when there are none:
//   state 1: some partial list was saturated -> K-collect re-scans those splits, and the CTA that finishes
//            the query's last work item re-ranks and finalises it on the spot (no grid-wide barrier);
//   state 2: not certifiable (overflow, NaN bound, forced) -> fail list -> K-exact, whose last CTA per group
//            of failed queries finalises them.
struct TailArgs {
  SelectArgs se;
  RerankArgs rr;
  FinalizeArgs fa;
  int32_t* state;       // [nq] 0 = finished by the tail kernel, 1 = deferred to K-collect, 2 = to K-exact
  int32_t* sat_n;       // [nq] saturated (query, list) pairs handed to K-collect
};
int launch_tail(const TailArgs& a, int nq, int n_sm, cudaStream_t st, bool dense = true);   // dense: 3 CTAs / SM at large batches

// K-collect's deferred finish: after the last (pair, sub-split) work item of a query the same CTA runs
// K-rerank and K-finalize for it.
struct DeferredArgs {
  RerankArgs rr;
  FinalizeArgs fa;
  const int32_t* state;   // [nq]
  const int32_t* sat_n;   // [nq]
  int32_t* done;          // [nq] finished work items per query (zeroed per search)
};

struct ExactArgs {
  const float* db_f32;
  const float* norm2;
  int64_t n_rows;
  int32_t D;
  const float* q;
  const float* qn2;
  int32_t nq, k;
  int64_t base_offset;
  const int32_t* fail_count;
  const int32_t* fail_list;
  Cand* partial;          // [max_fail_slots, n_splits, kList] exact scores
  int32_t n_splits;
  float* out_ip; float* out_dist; int64_t* out_labels;
  unsigned long long* ceil_keys;   // [max_fail_slots] paging state: key of the last result emitted per failed slot
  int32_t page;                    // results [32*page, 32*page + 32) of every failed query
  int32_t* group_done;             // fused form (k <= 32): [ceil(nq/4)] CTAs that finished a group's scan (zeroed per search)
};

struct MergeArgs {
  const float* ip; const float* dist; const int64_t* labels;  // [n_lists, nq, k]
  int32_t n_lists, nq, k;
  float* out_ip; float* out_dist; int64_t* out_labels;        // [nq, k]
};

int launch_pack(const PackArgs& a, cudaStream_t st);
int launch_normalize(float* x, int64_t n, int d, cudaStream_t st);
int launch_query_prep(const QueryPrepArgs& a, cudaStream_t st);
// groups of image vectors -> mean -> whole-vector L2 normalise (main/search_from_image.py:305-322)
int launch_group_prep(const float* parts, const int32_t* offs, int n_groups, int d, float* q_out, cudaStream_t st);
int launch_synth(const SynthArgs& a, cudaStream_t st);

// K-scan: returns the number of splits it will use for a device with n_sm SMs.
int scan_num_splits(int n_sm);
bool scan_supports(int Dp);
int launch_scan(const ScanArgs& a, int n_queries_this_pass, cudaStream_t st);

int launch_select(const SelectArgs& a, int nq, cudaStream_t st);
int launch_collect(const CollectArgs& a, int n_sm, cudaStream_t st);
int launch_collect_finish(const CollectArgs& a, const DeferredArgs& d, int n_sm, cudaStream_t st);   // k <= 32
int launch_rerank(const RerankArgs& a, int n_sm, cudaStream_t st);
int launch_finalize(const FinalizeArgs& a, cudaStream_t st);
int exact_num_splits(int n_sm);
int launch_exact(const ExactArgs& a, cudaStream_t st);   // scan + finalize of failed queries
int launch_exact_fused(const ExactArgs& a, cudaStream_t st);   // k <= 32: one launch, last CTA per group finalises
int launch_merge(const MergeArgs& a, cudaStream_t st);

// K-score (tcgen05): see score_tc.cu
struct ScoreTcPlan {
  int32_t n_qtiles;     // query tiles of 128
  int32_t n_splits;     // DB splits per query tile
  int32_t grid;         // CTAs
};
struct ScoreTcArgs {
  const void* tmap_q;   // CUtensorMap* (host copy passed by value at launch)
  const void* tmap_db;
  int64_t n_rows;
  int32_t Dp;
  int32_t nq;
  ScoreTcPlan plan;
  Cand* partial;        // [nq, n_lists, kList]; CTA (qtile, split) fills list `split`
  int32_t n_lists;
  int32_t max_tiles;    // > 0: sampling pass, every split scores only max_tiles of its tiles:
  int32_t tile_stride = 1;  //      tiles begin, begin + stride, ... (spread over the split; main pass: 1)
  const float* thr_floor;   // [nq] seeded admission floor (may be null)
  // in-kernel seeding (single-CTA kernel, one query tile, every CTA resident): inside their first tile the
  // CTAs exchange their lists through `partial`, CTA s computes query s's floor into seed_floor, and all
  // continue with it.  seed_k = 0: off.  grid_bar: two zeroed counters.
  int32_t seed_k;
  const float* seed_eps;    // [nq]
  float* seed_floor;        // [nq], pre-set to -inf
  unsigned int* grid_bar;   // [2]
};

// Threshold seeding: from the partial lists of a sampling pass, floor[q] = (k-th best sampled
// score) - 2 eps[q], one ulp lower.  No row of the shard below it can be a re-rank candidate.
struct SeedArgs {
  const Cand* partial; int32_t n_lists, list_stride, k;
  const float* eps; float* thr_floor;
};
int launch_seed(const SeedArgs& a, int nq, cudaStream_t st);
bool score_tc_supports(int Dp);
int score_tc_tile_rows();      // DB rows per accumulator tile: split boundaries are multiples of it
ScoreTcPlan score_tc_plan(int nq, int64_t n_rows, int n_sm, int forced_splits, int min_splits);
int score_tc_encode_maps(void* tmap_q_out, void* tmap_db_out, const uint16_t* q_bf16, int nq_pad,
                         const uint16_t* db_bf16, int64_t n_rows, int Dp);
int launch_score_tc(const ScoreTcArgs& a, cudaStream_t st);
int score_tc_max_coresident(int n_sm);    // CTAs resident at once (occupancy query): bound of the in-kernel seeding
int score_tc2_max_coresident(int n_sm);
// CTA-pair variant (score_tc2.cu): 256 queries per pair, DB tile halves of 128 rows per CTA
ScoreTcPlan score_tc2_plan(int nq, int64_t n_rows, int n_sm, int forced_splits, int min_splits);
int score_tc2_encode_db_map(void* tmap_db_out, const uint16_t* db_bf16, int64_t n_rows, int Dp);
int launch_score_tc2(const ScoreTcArgs& a, cudaStream_t st);

// Transposed CTA-pair variant (score_tn.cu): DB rows on M, up to 256 queries on N (tensor work proportional to
// ceil(nq/16)*16; thread = DB row in the epilogue).  plan.n_splits = sub-splits (= partial lists), tiles of
// score_tn_tile_rows() rows.  Needs tmap_db with 128-row boxes (score_tc2_encode_db_map) and its own query map.
bool score_tn_supports(int Dp, int nq);
int score_tn_tile_rows();
int score_tn_n16(int nq);
ScoreTcPlan score_tn_plan(int nq, int64_t n_rows, int n_sm);
int score_tn_encode_q_map(void* tmap_q_out, const uint16_t* q_bf16, int nq_pad, int Dp, int nq);
int launch_score_tn(const ScoreTcArgs& a, cudaStream_t st);

}  // namespace b2k
