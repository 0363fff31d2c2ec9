/*
 * b2k.h — C ABI of the B200-native exact top-k engine behind the image_recommender
 * retrieval hot path.
 *
 * The reference (AAPPHH/image_recommender) has no FFI layer of its own; its boundary to
 * native arithmetic is the faiss Python object protocol.  Every entry point below names
 * the reference call site (path:line under /root/reference) it replaces.  All functions
 * are called from ONE host thread per index; the library owns all device memory; host
 * pointers are borrowed only for the duration of a call.  No torch types appear here.
 *
 * Status convention: 0 = ok, >0 = cudaError_t, <0 = B2K_E_* below.  After any non-zero
 * status b2k_last_error() returns a thread-local human-readable message.
 *
 * There is deliberately NO CPU implementation behind this ABI: without a CUDA device
 * every compute entry point returns an error.
 */
#ifndef B2K_H_
#define B2K_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2K_ABI_VERSION 2
#define B2K_MAX_TABLES 8
#define B2K_MAX_K 1024        /* top-k limit.  k <= 32 is the tuned case; above it the engine uses more DB  */
                              /* splits and sorts the per-query lists instead of extracting k items         */
#define B2K_LIST 32           /* entries kept per (query, DB split) by the scoring kernels   */

#define B2K_E_INVALID  (-1)   /* bad argument                                   */
#define B2K_E_CAPACITY (-2)   /* add() beyond the capacity given to create()    */
#define B2K_E_IO       (-3)   /* save/load failure                              */
#define B2K_E_NODEVICE (-4)   /* no CUDA device / wrong architecture            */
#define B2K_E_NOMEM    (-5)   /* host allocation failure                        */
#define B2K_E_UNSUPPORTED (-6) /* input format this fast path does not handle    */
#define B2K_E_PEER     (-7)   /* a peer rank did not take part in the exchange  */

typedef struct b2k_index b2k_index;

/* b2k_set_option keys */
#define B2K_OPT_PATH          1  /* 0 auto, 1 = K-scan (CUDA-core stream), 2 = K-score (tcgen05 GEMM) */
#define B2K_OPT_RERANK        2  /* re-rank candidate slots per query: 0 = every listed entry (default),  */
                                 /* else 32..8192 (overflowing queries take the exact fp32 scan)          */
#define B2K_OPT_FORCE_EXACT   3  /* 1 = treat every query as uncertified (exercise the exact fp32 scan) */
#define B2K_OPT_SCAN_MAX_B    4  /* auto path: nq <= this uses K-scan, above uses K-score (default 0: the  */
                                 /* tcgen05 path measured faster at every batch size and row width)      */
#define B2K_OPT_SPLITS        5  /* 0 auto; else number of DB splits per query tile                      */
#define B2K_OPT_TIGHTEN       8  /* candidate threshold from the exact scores of the k best rows: 1 auto (default: */
                                 /* from 128 queries; from 32 on shards of 4 M rows and more), 0 off, 2 always */
#define B2K_OPT_COLLECT       9  /* saturated partial lists are re-scanned by K-collect: 1 on (default), 0 = exhaustive scan */
#define B2K_OPT_INLINE_SEED  10  /* <= 128 queries: seeding folded into the scoring launch (grid barrier): 1 on (default) */
#define B2K_OPT_FUSED_TAIL   11  /* k <= 32: select + re-rank + finalize as ONE launch (cluster of CTAs per query), K-collect and  */
                                 /* K-exact finish their queries themselves: 1 on (default), 0 = one launch per stage (as k > 32), */
                                 /* 2 = on, without the 3-CTAs-per-SM instantiation for large batches (A/B measurements)          */
#define B2K_OPT_TN           12  /* transposed K-score kernel (queries on the MMA's N, <= 256 queries): -1 auto (129..240 queries; */
                                 /* narrow rows at small batches on long shards), 0 never, 1 whenever it applies              */
#define B2K_OPT_SAMPLE_WAVE  13  /* sampling pass of the seeding on ONE wave of long CTAs with strided tiles: 1 on (default), */
                                 /* 0 = the main pass's grid, every split its first tiles                                      */
#define B2K_OPT_SEED          7  /* K-score threshold seeding: 1 auto (default), 0 off, N > 1 = a sampling pass of  */
                                 /* N tiles per split (forces the three-launch form, no in-kernel seeding)        */
#define B2K_OPT_TC_PAIR       6  /* K-score kernel: -1 auto (CTA pairs above 128 queries unless the last 256-query */
                                 /* tile would be half empty, up to 896 queries), 0 single CTA, 1 pairs          */

typedef struct b2k_stats {
  int32_t path;            /* last search: 1 = K-scan, 2 = K-score (1 CTA), 3 = K-score (CTA pairs), 4 = K-score transposed (CTA pairs, queries on N) */
  int32_t n_splits;        /* DB splits (partial lists) per query                            */
  int32_t cand_slots;      /* candidate slots per query (capacity; rows actually re-ranked: n_candidates) */
  int32_t n_uncertified;   /* queries whose certificate failed -> served by exact fp32 scan  */
  float   eps_max;         /* largest certificate slack used (bound on |bf16 score - exact|) */
  float   err_max;         /* max_row ||bf16(x) - x||_2 over the shard                       */
  float   norm_max;        /* max_row ||x||_2 over the shard                                 */
  int32_t launches;        /* kernels launched by the last search                            */
  float   score_ms;        /* device time of the scoring kernel(s) (CUDA events on the         */
                           /* launching stream; the roofline numerator's clock): mean over the */
                           /* passes since the previous b2k_get_stats (the last 64 at most)    */
  float   tail_ms;         /* same for select + rerank + finalize (+ collect / exact)          */
  int32_t n_queries;       /* queries of the last pass                                        */
  int32_t n_candidates;    /* rows re-ranked in fp32 over all queries of the last pass        */
  int32_t n_saturated;     /* (query, DB split) pairs whose partial list was full of candidates */
                           /* in the last pass: re-scanned by K-collect                        */
  int32_t n_timed;         /* passes averaged into score_ms / tail_ms                          */
} b2k_stats;

typedef struct b2k_synth {
  uint64_t seed;
  int32_t  n_clusters;     /* image-level clusters shared by all tables (default 4096) */
  float    sigma;          /* intra-cluster noise scale (default 0.3)                  */
  uint32_t abs_mask;       /* bit t set: table t is non-negative (colour histograms)   */
  int64_t  total_rows;     /* N of the whole (all-shard) DB; queries draw rows from it  */
} b2k_synth;

const char* b2k_last_error(void);
int32_t     b2k_abi_version(void);
int         b2k_device_count(int32_t* n);

/* Replaces faiss.IndexHNSWFlat(dim, M) / IndexIVFPQ(...) construction
 * (main/create_index.py:207-234).  One flat exact store per GPU shard.
 * table_dims[t] = d_t in concat order; rows get global offsets base_offset + local row. */
int b2k_create(const int32_t* table_dims, int32_t n_tables, int64_t capacity_rows,
               int32_t device, int64_t base_offset, b2k_index** out);
void b2k_destroy(b2k_index* idx);
/* faiss indexes grow on add(); this store is sized explicitly.  Re-allocates the shard for
 * capacity_rows >= ntotal and moves the rows device-to-device (bits unchanged). */
int b2k_reserve(b2k_index* idx, int64_t capacity_rows);
int64_t b2k_capacity(const b2k_index* idx);
/* index.reset(): forget every row (capacity and allocations are kept). */
int b2k_reset(b2k_index* idx);

/* Replaces index.add(arr) (main/create_index.py:311) fused with the per-row
 * np.concatenate of _process_batch (main/create_index.py:176-188): host_tables[t] is a
 * C-contiguous fp32 [n, d_t] host array; rows are per-table L2-normalised, concatenated,
 * stored as fp32 and bf16.  Row i gets offset base_offset + ntotal_before + i. */
int b2k_add(b2k_index* idx, const float* const* host_tables, int64_t n);
/* Same with device-resident inputs on `stream` (cudaStream_t as void*). */
int b2k_add_device(b2k_index* idx, const float* const* dev_tables, int64_t n, void* stream);

/* Pinned two-slot ingest staging: the zero-copy form of add() for a host-side decoder
 * (replaces the np.stack(...).astype("float32") batch materialisation of
 * main/create_index.py:288,310 together with index.add, :311).  The caller writes decoded rows of
 * table t straight into slot s's pinned buffer ([rows_per_slot, d_t] fp32, b2k_stage_ptr), commits
 * the slot (asynchronous H2D + K-pack, appends n rows) and fills the other slot meanwhile;
 * b2k_stage_wait(s) blocks until slot s may be overwritten again.  close() drains and frees. */
int     b2k_stage_open(b2k_index* idx, int64_t rows_per_slot);
int     b2k_stage_open_n(b2k_index* idx, int64_t rows_per_slot, int32_t n_slots);   /* 2..32 slots (several decoders) */
int64_t b2k_stage_rows(const b2k_index* idx);      /* rows per slot actually allocated (0 = closed) */
int     b2k_stage_ptr(b2k_index* idx, int32_t slot, int32_t table, float** host_ptr);
int     b2k_stage_wait(b2k_index* idx, int32_t slot);
int     b2k_stage_commit(b2k_index* idx, int32_t slot, int64_t n_rows);
int     b2k_stage_close(b2k_index* idx);

/* Native ingest: replaces the Python row loop of FAISSIndexBuilderDB._batch_records +
 * _process_batch (main/create_index.py:144-189) for databases whose vector blobs are what the
 * reference's extractors write (vector_scripts/create_vector_base.py:142-145:
 * pickle.dumps(1-D float32 ndarray, HIGHEST_PROTOCOL) under numpy >= 2).  Runs `sql` (first column
 * = images.id, then one blob column per table, in the index's table order) through libsqlite3,
 * views each blob's float32 payload in place, writes it into the pinned staging slots and commits
 * them as they fill.  ids_out[i] = image id of the i-th appended row.  Any blob in another format
 * (or of another length) stops the call with B2K_E_UNSUPPORTED *before* that row is staged:
 * the caller resets the index and falls back to its own decoder (the reference's pickle.loads).
 * b2k_parse_f32_blob is the strict recogniser it uses (host only; 0 = recognised). */
int b2k_ingest_sqlite(b2k_index* idx, const char* db_path, const char* sql, int64_t* ids_out,
                      int64_t ids_cap, int64_t* n_added);
/* The same with n_threads reader threads (the single-threaded loop is bound by SQLite's page reads and the row
 * copies: ~0.35 M rows/s).  `sql_range` is the same query with two parameters bounding images.id,
 * "... WHERE i.id >= ?1 AND i.id < ?2 ORDER BY i.id"; id_bounds[0..n_chunks] cuts the id space into chunks of at most
 * b2k_stage_rows() images each (the caller takes every rows-per-slot-th id of `SELECT id FROM images ORDER BY id`;
 * id_bounds[n_chunks] = last id + 1).  Thread p owns its own read-only connection, decodes chunks p, p + n_threads, ...
 * into its own two staging slots (b2k_stage_open_n(idx, rows, 2 * n_threads) first) and commits them IN CHUNK ORDER, so
 * rows, offsets and ids_out are exactly those of the single-threaded call.  Same error behaviour. */
int b2k_ingest_sqlite_mt(b2k_index* idx, const char* db_path, const char* sql_range, const int64_t* id_bounds,
                         int64_t n_chunks, int32_t n_threads, int64_t* ids_out, int64_t ids_cap, int64_t* n_added);
int b2k_parse_f32_blob(const void* blob, int64_t n_bytes, const float** payload, int64_t* dim);

/* index.ntotal (main/create_index.py:321, main/search_from_image.py:340) */
int64_t b2k_ntotal(const b2k_index* idx);
int32_t b2k_dim(const b2k_index* idx);
int32_t b2k_table_dims(const b2k_index* idx, int32_t* dims_out);   /* returns n_tables; dims_out may be NULL */
int32_t b2k_dim_padded(const b2k_index* idx);
int64_t b2k_base_offset(const b2k_index* idx);

/* Replaces index.search(query_vec, top_k) (main/search_from_image.py:247).
 * q: fp32 [nq, D] host.  Outputs (host): dist fp32 [nq,k] squared-L2 (ascending when all
 * rows share one norm, as per-table normalisation guarantees), labels int64 [nq,k] global
 * offsets, -1 padded when k > ntotal (faiss convention).  ip (optional, may be NULL)
 * receives the exact fp32 inner products, descending. */
int b2k_search(b2k_index* idx, const float* q_host, int32_t nq, int32_t k,
               float* dist_host, int64_t* labels_host, float* ip_host);
/* Device-resident variant: q_dev/outputs are device pointers, work is enqueued on
 * `stream` (cudaStream_t as void*; NULL = the legacy default stream) and NOT synchronised.
 * The index owns one search workspace: use one stream at a time per index. */
int b2k_search_device(b2k_index* idx, const float* q_dev, int32_t nq, int32_t k,
                      float* dist_dev, int64_t* labels_dev, float* ip_dev, void* stream);
/* Replaces the query-vector assembly of ImageRecommender._extract_query_vector together with the search
 * (main/search_from_image.py:305-322 + :247) for MANY query groups at once: parts_host is the [n_images, D]
 * matrix of concatenated per-image vectors (row = np.concatenate(parts, axis=1), :309), group g owns rows
 * [group_offsets[g], group_offsets[g+1]) (offsets[0] = 0, offsets[n_groups] = n_images, no empty group).  On
 * the device: np.mean over the group's images (:317; fp32, images added in order, one division), then
 * faiss.normalize_L2 of the mean (:322), written straight into the search workspace — the normalised
 * queries never travel back to the host — then the search.  Outputs as b2k_search, one row per group. */
int b2k_search_groups(b2k_index* idx, const float* parts_host, int64_t n_images, const int32_t* group_offsets,
                      int32_t n_groups, int32_t k, float* dist_host, int64_t* labels_host, float* ip_host);
/* The prep step alone on device buffers (row-sharded deployments run it on every rank before the local
 * search): q_dev [n_groups, d] = normalize_L2(mean of each group's rows). */
int b2k_prep_groups_device(const float* parts_dev, const int32_t* group_offsets_dev, int32_t n_groups, int32_t d,
                           float* q_dev, int32_t device, void* stream);
int b2k_get_stats(b2k_index* idx, b2k_stats* out);   /* synchronises the last search */
int b2k_set_option(b2k_index* idx, int32_t key, int64_t value);

/* Cross-shard merge (new): n_lists per-shard results [n_lists, nq, k] ->
 * [nq, k], order = higher ip first, then lower offset.  All pointers on `device`.  Up to 32 lists
 * (any k): every list must already be in that order with its -1 padding last, which is what
 * b2k_search* returns; more lists: any order, n_lists * k <= 1024. */
int b2k_merge_topk_device(const float* ip, const float* dist, const int64_t* labels,
                          int32_t n_lists, int32_t nq, int32_t k,
                          float* out_ip, float* out_dist, int64_t* out_labels,
                          int32_t device, void* stream);

/* Peer-memory exchange for the row-sharded deployment (one process per GPU of one box): each rank's
 * final [nq, k] records are stored straight into every rank's receive buffer over NVLink (CUDA IPC
 * mappings) and merged there — the low-latency alternative to an NCCL all-gather + K-merge.
 *   create  -> handle (64-byte cudaIpcMemHandle_t; exchange the handles of all ranks out of band,
 *   e.g. torch.distributed.all_gather) -> connect -> per search: push, then merge (same stream).
 * connect(): `handles` = world x 64 bytes, or for ranks living in ONE process `raw_ptrs` = the
 * device pointers returned through handle()'s raw_ptr.  max_entries bounds nq*k of a search. */
typedef struct b2k_xchg b2k_xchg;
int  b2k_xchg_create(int32_t device, int32_t rank, int32_t world, int64_t max_entries, b2k_xchg** out);
void b2k_xchg_destroy(b2k_xchg* x);
int  b2k_xchg_handle(b2k_xchg* x, void* handle64, void** raw_ptr);
int  b2k_xchg_connect(b2k_xchg* x, const void* handles, const void* const* raw_ptrs);
int  b2k_xchg_push(b2k_xchg* x, const float* ip, const float* dist, const int64_t* labels, int32_t nq,
                   int32_t k, void* stream);
int  b2k_xchg_merge(b2k_xchg* x, int32_t nq, int32_t k, float* out_ip, float* out_dist,
                    int64_t* out_labels, void* stream);
/* Synchronises the device and reports merges that gave up waiting for a peer (a dead or stalled rank: the
 * merge kernel waits 10 s, pads its outputs with label -1 and returns instead of hanging or trapping, so the
 * resident index survives).  *timed_out_ranks: bit g = rank g was missing; returns B2K_E_PEER when non-zero
 * (the bits are cleared by the call). */
int  b2k_xchg_status(b2k_xchg* x, uint32_t* timed_out_ranks);
/* A rank whose local search FAILED still owes its peers an epoch: skip() publishes the next epoch without records
 * (the peers' merges of this search return at once with stale entries for this rank; the caller is about to report
 * the failure anyway) — otherwise every peer would wait out the 10 s timeout. */
int  b2k_xchg_skip(b2k_xchg* x, void* stream);

/* ---- several GPUs, ONE process ---------------------------------------------------------------------------
 * The reference's CLI / ImageRecommender is a single process (main/search_from_image.py:430-441): a b2k_group
 * drives the row shards on n_devices GPUs of the box from the one calling thread (internally one worker thread
 * per device; the devices' results travel to the first device over NVLink peer memory and are merged there) —
 * no process group, no NCCL, no IPC handles.  devices == NULL: 0 .. n_devices-1.
 *   load()         every device loads ITS contiguous row range of the index file (rows ceil(n/G) * r ...)
 *   set_shard()    or: adopt shards created by the caller (one per rank, base offsets = their row ranges)
 *   search()       = put_queries (H2D to every device) + run (local search, exchange, merge; synchronous)
 *                    + get_results (D2H from the first device); the three steps are exposed so that a caller
 *                    can keep the queries resident, as bench.py does
 *   search_groups()  b2k_search_groups over the group. */
typedef struct b2k_group b2k_group;
int  b2k_group_create(const int32_t* devices, int32_t n_devices, b2k_group** out);
void b2k_group_destroy(b2k_group* g);
int32_t b2k_group_size(const b2k_group* g);
int  b2k_group_load(b2k_group* g, const char* path);
int  b2k_group_set_shard(b2k_group* g, int32_t rank, b2k_index* shard);     /* not owned by the group */
b2k_index* b2k_group_shard(b2k_group* g, int32_t rank);
int64_t b2k_group_ntotal(const b2k_group* g);
int32_t b2k_group_dim(const b2k_group* g);
int  b2k_group_search(b2k_group* g, const float* q_host, int32_t nq, int32_t k, float* dist_host,
                      int64_t* labels_host, float* ip_host);
int  b2k_group_put_queries(b2k_group* g, const float* q_host, int32_t nq, int32_t k);
int  b2k_group_run(b2k_group* g, int32_t nq, int32_t k);
int  b2k_group_get_results(b2k_group* g, float* dist_host, int64_t* labels_host, float* ip_host);
int  b2k_group_search_groups(b2k_group* g, const float* parts_host, int64_t n_images, const int32_t* group_offsets,
                             int32_t n_groups, int32_t k, float* dist_host, int64_t* labels_host, float* ip_host);
int  b2k_group_last_run_ms(const b2k_group* g, float* max_ms);   /* device time of the last run(), max over devices */

/* Replaces faiss.normalize_L2(x) (main/search_from_image.py:322): in place on a host
 * array, rows with zero norm untouched; computed on `device`. */
int b2k_normalize_l2(float* x_host, int64_t n, int32_t d, int32_t device);

/* Replace faiss.write_index / faiss.read_index (main/create_index.py:320,
 * main/search_from_image.py:339).  ids (optional) = image_id per row, carried in the file.
 * load() reads rows [row_begin, row_end) of the file (row_end < 0: to the end) so that
 * each GPU of a row-sharded deployment loads only its shard. */
int b2k_save(b2k_index* idx, const char* path, const int64_t* ids, int64_t n_ids);
/* Row-sharded build: every rank of a box writes ITS rows into one index file laid out for
 * file_total_rows rows.  The rank that owns row 0 calls first with create = 1 (header + extent),
 * the others after it (a barrier between) with create = 0; ids = this shard's image ids or NULL
 * on every rank alike. */
int b2k_save_shard(b2k_index* idx, const char* path, const int64_t* ids, int64_t file_row_begin,
                   int64_t file_total_rows, int32_t create);
/* capacity_rows: rows the loaded shard is allocated for (<= 0 or smaller than the range: exactly the
 * range).  An index that is about to grow (--update) is loaded straight into its final capacity, so
 * that no re-allocation (old + new arrays resident at once) ever happens. */
int b2k_load(const char* path, int32_t device, int64_t row_begin, int64_t row_end,
             int64_t capacity_rows, b2k_index** out);
int b2k_file_info(const char* path, int64_t* n_rows, int32_t* n_tables, int32_t* table_dims,
                  int32_t* has_ids);
int b2k_load_ids(const char* path, int64_t row_begin, int64_t n, int64_t* ids_out);

/* Read back packed rows (parity tests): any of the outputs may be NULL. */
int b2k_get_rows(b2k_index* idx, int64_t row0, int64_t n, float* f32_host,
                 uint16_t* bf16_host, float* norm2_host);

/* Device-side synthetic data for benches: appends n clustered rows whose
 * global offsets start at base_offset + ntotal; bit-reproducible on the CPU
 * (oracle/b2k_oracle.c: orc_synth_rows).  Queries are noisy copies of seeded DB rows, whole-vector normalised. */
int b2k_fill_synthetic(b2k_index* idx, int64_t n, const b2k_synth* p);
int b2k_synth_queries_device(b2k_index* idx, int32_t nq, const b2k_synth* p, uint64_t qseed,
                             float sigma_q, float* q_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2K_H_ */
