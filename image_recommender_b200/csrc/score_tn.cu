// K-score, transposed ("swap-AB") CTA-pair version: DB rows are the M operand, QUERIES the N operand.
//
// Same contract as score_tc.cu / score_tc2.cu (bf16 scores S = Q · Xᵀ on the tensor cores, a running top-32 per
// (query, DB split) fused into the epilogue; replaces the contraction of index.search,
// main/search_from_image.py:247), for batches of at most 256 queries.  tcgen05.mma fixes M at 128 rows per CTA
// whatever is live in it, but takes any N that is a multiple of 16: with the queries on N
//   * the tensor work is proportional to ceil(nq / 16) * 16 queries — 160 queries cost 62.5 % of what 256 cost
//     (the M = queries kernels pay for 256; profiles/r02_prof_mid_*: that regime is tensor/power bound), and
//   * an epilogue thread owns a DB ROW and reads only the N live columns: one query costs a 16-column
//     tcgen05.ld per 128 rows instead of draining 256 columns per 256 rows with 1 of 128 lanes useful (narrow
//     rows at tiny batches were epilogue bound: 0.43 of the HBM peak at D = 48).
// Two CTAs of one TPC form a pair (tcgen05 cta_group::2, M = 256): each stages 128 DB rows (A half, 16 KB per
// K-block) and half of the queries (B half, N/2 rows); CTA (pair p, rank r) scores its OWN contiguous range of
// 128-row tiles — "sub-split" 2p + r — so that a partial list maps to one row range (K-collect re-scans it).
//   warp 0    TMA producer (both CTAs), bytes accounted on the leader's mbarrier
//   warp 1    MMA issuer (leader only), tcgen05.mma.cta_group::2.kind::f16 M=256 N=nq16 K=16
//   warps 2-5 epilogue (both CTAs): thread = DB row; per 16 queries one tcgen05.ld.32x32b.x16, compare against
//             the per-query admission thresholds kept in shared memory; a (rare) hit is inserted into the query's
//             32-entry list in shared memory by the whole warp under a per-query lock (4 epilogue warps share it)
// The kernel always runs with a seeded admission floor (sampling pass of the M = queries kernel + seed_kernel,
// api.cu): without one the first tile would insert 128 rows x N queries one by one.
#include "tc_common.cuh"

namespace b2k {

using namespace tc;

namespace {
constexpr int kTnRows = 128;                       // DB rows per CTA and tile (TMEM lanes)
constexpr int kTnABytes = kTnRows * kBlockK * 2;   // 16 KB per K-block
constexpr int kTnMaxQ = 256;
constexpr int kTnMaxAcc = 8;                       // accumulators in flight: 512 TMEM columns / N, at most 8 (narrow
                                                   // rows are bound by the MMA -> epilogue -> MMA round trip per tile)
__host__ __device__ inline int tn_n_acc(int n16) { int a = 2; while (a < kTnMaxAcc && 2 * a * n16 <= kTmemCols) a *= 2; return a; }

struct TnSmem {                                    // byte offsets from the 1024-aligned base
  int stage_bytes, n_stages, lists, thr, floor, lock, bars, total;
};
__host__ __device__ inline TnSmem tn_layout(int n16) {
  TnSmem s;
  const int b_bytes = (n16 / 2) * kBlockK * 2;     // this CTA's half of the queries per K-block (multiple of 1024)
  s.stage_bytes = kTnABytes + b_bytes;
  const int fixed = n16 * kList * 8 + 3 * n16 * 4 + 1024;
  int st = (220 * 1024 - 1024 - fixed) / s.stage_bytes;
  s.n_stages = st > 8 ? 8 : st;
  s.lists = s.n_stages * s.stage_bytes;
  s.thr = s.lists + n16 * kList * 8;
  s.floor = s.thr + n16 * 4;
  s.lock = s.floor + n16 * 4;
  s.bars = (s.lock + n16 * 4 + 15) & ~15;
  s.total = s.bars + 1024 + 1024;
  return s;
}

// TMEM -> registers: 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// Whole-warp insertion of (sc, row) into query q's 32-entry list (lane e owns entry e), under the query's lock:
// the entry with the lowest score is replaced when sc beats it; the admission threshold follows the new minimum.
__device__ __forceinline__ void tn_insert(Cand* list, float* thr, const float* floor, int* lock, int q, float sc,
                                          int32_t row, int lane) {
  if (lane == 0) {
    while (atomicCAS(lock + q, 0, 1) != 0) { }
    __threadfence_block();
  }
  __syncwarp();
  // (the lock's fences and __syncwarp are compiler barriers: plain shared-memory accesses are re-issued here)
  Cand* mine = list + q * kList + lane;
  const Cand e = *mine;
  float ms = e.row < 0 ? -INFINITY : e.score;
  // arg-min over the lanes (ties: the higher lane, any is fine)
  float m = ms;
  int who = lane;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int ow = __shfl_xor_sync(0xffffffffu, who, o);
    if (om < m || (om == m && ow > who)) { m = om; who = ow; }
  }
  if (sc > m) {
    if (lane == who) { Cand n; n.score = sc; n.row = row; *mine = n; ms = sc; }
    float nm = ms;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nm = fminf(nm, __shfl_xor_sync(0xffffffffu, nm, o));
    if (lane == 0) thr[q] = fmaxf(nm, floor[q]);
  }
  __syncwarp();
  if (lane == 0) {
    __threadfence_block();
    atomicExch(lock + q, 0);
  }
}
}  // namespace

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
score_tn_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_db,
                int64_t n_rows, int32_t n_kblocks, int32_t nq, int32_t n16, uint32_t idesc, int32_t n_sub,
                int32_t n_lists, const float* __restrict__ thr_floor, Cand* __restrict__ partial) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const TnSmem L = tn_layout(n16);
  unsigned char* ring = smem;
  Cand* lists = reinterpret_cast<Cand*>(smem + L.lists);            // [n16][32]
  float* s_thr = reinterpret_cast<float*>(smem + L.thr);            // [n16]
  float* s_floor = reinterpret_cast<float*>(smem + L.floor);        // [n16]
  int* s_lock = reinterpret_cast<int*>(smem + L.lock);              // [n16]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bars);      // [8] used in the leader only
  uint64_t* empty = full + 8;                                       // [8]
  uint64_t* tfull = empty + 8;                                      // [8]
  uint64_t* tempty = tfull + 8;                                     // [8] used in the leader only
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tempty + 8);
  const int n_stages = L.n_stages;
  const int n_acc = tn_n_acc(n16);
  const int acc_stride = kTmemCols / n_acc;                         // TMEM columns between accumulators
  const int b_bytes = L.stage_bytes - kTnABytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();          // 0 = leader
  const int sub = (int)blockIdx.x;                        // sub-split = 2 * pair + rank: its own contiguous tile range
  const int64_t tiles_total = (n_rows + kTnRows - 1) / kTnRows;
  const int64_t my_begin = tiles_total * sub / n_sub, my_end = tiles_total * (sub + 1) / n_sub;
  const int peer = sub ^ 1;
  const int64_t pr_begin = tiles_total * peer / n_sub, pr_end = tiles_total * (peer + 1) / n_sub;
  const int my_tiles = (int)(my_end - my_begin);
  const int n_tiles = max(my_tiles, (int)(pr_end - pr_begin));      // the pair runs in lockstep

  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int s = 0; s < kTnMaxAcc; ++s) { ptx::mbar_init(&tfull[s], 1); ptx::mbar_init(&tempty[s], 8); }
    ptx::fence_mbar_init();
  }
  // lists / thresholds / locks (epilogue state) are initialised by all threads before the cluster barrier
  for (int i = threadIdx.x; i < n16 * kList; i += kThreads) { Cand c; c.score = -INFINITY; c.row = -1; lists[i] = c; }
  for (int i = threadIdx.x; i < n16; i += kThreads) {
    const float f = i < nq ? (thr_floor ? thr_floor[i] : -INFINITY) : INFINITY;   // padded queries never insert
    s_floor[i] = f; s_thr[i] = f; s_lock[i] = 0;
  }
  if (warp == 0 && lane == 0) { ptx::tma_prefetch_desc(&tmap_q); ptx::tma_prefetch_desc(&tmap_db); }
  if (warp == 1) { ptx::tmem_alloc_pair(tmem_base_slot, kTmemCols); ptx::tmem_relinquish_pair(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();               // barriers of both CTAs initialised, TMEM allocated
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    if (ptx::elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < n_tiles; ++t) {
        // a CTA whose own range is exhausted keeps feeding the pair's MMA: any in-bounds tile will do (its rows are
        // masked in the epilogue); clamp to the last tile of the database
        const int64_t tile = t < my_tiles ? my_begin + t : tiles_total - 1;
        const int32_t row0 = (int32_t)(tile * kTnRows);
        for (int kb = 0; kb < n_kblocks; ++kb) {
          ptx::mbar_wait(&empty[s], ph ^ 1);
          unsigned char* a_dst = ring + (size_t)s * L.stage_bytes;
          unsigned char* b_dst = a_dst + kTnABytes;
          if (rank == 0) ptx::mbar_arrive_expect_tx(&full[s], 2u * (uint32_t)L.stage_bytes);   // bytes of both CTAs
          ptx::tma_load_2d_pair(a_dst, &tmap_db, kb * kBlockK, row0, &full[s], ptx::kEvictFirst);
          ptx::tma_load_2d_pair(b_dst, &tmap_q, kb * kBlockK, (int32_t)rank * (n16 / 2), &full[s], ptx::kEvictLast);
          if (++s == n_stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader only) ------------------------------
    if (rank == 0) {
      const uint32_t ring_u = __shfl_sync(0xffffffffu, ptx::smem_u32(ring), 0);
      const uint32_t full_u = __shfl_sync(0xffffffffu, ptx::smem_u32(full), 0);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t empty_u = full_u + 64, tfull_u = empty_u + 64, tempty_u = tfull_u + 64;
      const uint64_t desc0 = make_sw128_desc(ring_u);
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int acc = t % n_acc;
        ptx::mbar_wait_a(tempty_u + acc * 8, ((t / n_acc) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_u + (uint32_t)(acc * acc_stride);
        for (int kb = 0; kb < n_kblocks; ++kb) {
          ptx::mbar_wait_a(full_u + s * 8, ph);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint64_t a_desc = desc0 + (uint64_t)(s * (L.stage_bytes >> 4));     // DB rows: the M operand
            const uint64_t b_desc = a_desc + (uint64_t)(kTnABytes >> 4);              // queries: the N operand
            static_assert(kBlockK / kUmmaK == 4, "umma_bf16_pair_kblock issues four K=16 MMAs");
            ptx::umma_bf16_pair_kblock(d_tmem, a_desc, b_desc, idesc, kb != 0 ? 1u : 0u, empty_u + s * 8, 3);
            if (kb == n_kblocks - 1) ptx::umma_commit_pair_a(tfull_u + acc * 8, 3);
          }
          __syncwarp();
          if (++s == n_stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    // ------------------------------ epilogue (both CTAs) ------------------------------
    const int quarter = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    for (int t = 0; t < n_tiles; ++t) {
      const int acc = t % n_acc;
      const int64_t row = (my_begin + t) * kTnRows + quarter * 32 + lane;
      const bool valid = t < my_tiles && row < n_rows;
      ptx::mbar_wait(&tfull[acc], (t / n_acc) & 1);
      ptx::tc_fence_after();
      if (__any_sync(0xffffffffu, valid)) {
        for (int c = 0; c < n16; c += 16) {
          uint32_t v[16];
          tmem_ld_32x16(lane_base + (uint32_t)(acc * acc_stride + c), v);
          ptx::tmem_ld_wait();
          // any score above its query's threshold?  thr - v is negative exactly then: OR the sign bits
          uint32_t any = 0u;
          const float4* t4 = reinterpret_cast<const float4*>(s_thr + c);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 th = t4[j4];                                      // same address in every lane: broadcast
            any |= __float_as_uint(th.x - __uint_as_float(v[4 * j4 + 0]));
            any |= __float_as_uint(th.y - __uint_as_float(v[4 * j4 + 1]));
            any |= __float_as_uint(th.z - __uint_as_float(v[4 * j4 + 2]));
            any |= __float_as_uint(th.w - __uint_as_float(v[4 * j4 + 3]));
          }
          if (__any_sync(0xffffffffu, valid && (any >> 31))) {
            // rare after the seeded floor: per query, the lanes (rows) that beat the threshold insert one by one
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float sc = __uint_as_float(v[j]);
              const float th = *reinterpret_cast<volatile float*>(s_thr + c + j);
              unsigned m = __ballot_sync(0xffffffffu, valid && sc > th);
              while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const float bs = __shfl_sync(0xffffffffu, sc, b);
                tn_insert(lists, s_thr, s_floor, s_lock, c + j, bs, (int32_t)(row - lane + b), lane);
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_remote(&tempty[acc], 0);          // the leader's barrier
    }
    // the four epilogue warps share the lists: all done before they are written out, one warp per query
    ptx::named_bar_sync(2, 128);
    for (int q = quarter; q < nq; q += 4) {
      const Cand e = lists[q * kList + lane];
      partial[((int64_t)q * n_lists + sub) * kList + lane] = e;
    }
  }

  // neither CTA may exit (or free TMEM) while its peer can still touch its shared memory
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc_pair(tmem_base, kTmemCols); }
}

// ---------------------------------------------------------------------------------------
int score_tn_tile_rows() { return kTnRows; }
bool score_tn_supports(int Dp, int nq) { return Dp >= kBlockK && (Dp % kBlockK) == 0 && nq >= 1 && nq <= kTnMaxQ; }
int score_tn_n16(int nq) { return (nq + 15) / 16 * 16; }

ScoreTcPlan score_tn_plan(int nq, int64_t n_rows, int n_sm) {
  ScoreTcPlan p;
  p.n_qtiles = 1;
  const int64_t tiles_total = (n_rows + kTnRows - 1) / kTnRows;
  int sub = n_sm & ~1;                                   // one CTA per SM, in pairs
  while (sub > 2 && (int64_t)sub > tiles_total) sub -= 2;
  if (sub < 2) sub = 2;
  p.n_splits = sub;                                      // partial lists = sub-splits
  p.grid = sub;
  return p;
}

int score_tn_encode_q_map(void* tmap_q_out, const uint16_t* q_bf16, int nq_pad, int Dp, int nq) {
  const int n16 = score_tn_n16(nq);
  return encode_2d(reinterpret_cast<CUtensorMap*>(tmap_q_out), q_bf16, (uint64_t)nq_pad, (uint64_t)Dp, (uint32_t)(n16 / 2));
}

int launch_score_tn(const ScoreTcArgs& a, cudaStream_t st) {
  const int n16 = score_tn_n16(a.nq);
  const TnSmem L = tn_layout(n16);
  if (L.n_stages < 2) { set_error("score_tn: shared memory layout failed for %d queries", a.nq); return B2K_E_INVALID; }
  B2K_CUDA(cudaFuncSetAttribute(score_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
  const CUtensorMap* mq = reinterpret_cast<const CUtensorMap*>(a.tmap_q);
  const CUtensorMap* md = reinterpret_cast<const CUtensorMap*>(a.tmap_db);
  const uint32_t idesc = make_idesc(2 * kTnRows, n16);
  score_tn_kernel<<<a.plan.grid, kThreads, L.total, st>>>(*mq, *md, a.n_rows, a.Dp / kBlockK, a.nq, n16, idesc,
                                                         a.plan.n_splits, a.n_lists, a.thr_floor, a.partial);
  B2K_CHECK_LAUNCH();
  return 0;
}

}  // namespace b2k
