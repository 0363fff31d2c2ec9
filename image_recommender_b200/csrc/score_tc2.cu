// K-score, CTA-pair version (tcgen05 cta_group::2): the large-batch scoring kernel.
//
// Same contract as score_tc.cu (S = Q · Xᵀ in bf16 on the tensor cores, running top-32 per
// (query, DB split) fused into the epilogue; replaces the contraction of index.search,
// main/search_from_image.py:247), but two CTAs of one TPC cooperate on a 256-query x 256-row tile:
//   * each CTA stages ITS 128 queries (A, 16 KB) and HALF of the DB tile (B, 128 rows, 16 KB) per
//     K-block, so a stage is 32 KB instead of 48 KB: 6 stages fit and the L2 -> SM traffic per
//     flop drops by a third (the single-CTA kernel is L2-bandwidth bound, profiles/r01_*);
//   * the leader CTA (cluster rank 0) issues tcgen05.mma.cta_group::2 M=256 N=256 K=16; the
//     hardware reads A/B from both CTAs' shared memory and writes each CTA's 128 accumulator
//     rows into that CTA's TMEM;
//   * TMA loads of both CTAs complete on the LEADER's full barrier; tcgen05.commit multicasts the
//     "stage free" / "accumulator ready" arrivals to both CTAs; the epilogue warps of both CTAs
//     arrive remotely on the leader's "accumulator drained" barrier.
#include "tc_common.cuh"

namespace b2k {

using namespace tc;

namespace {
constexpr int kStages2 = 6;
constexpr int kABytes2 = kBlockM * kBlockK * 2;          // 16 KB: this CTA's 128 queries
constexpr int kBHalfRows = kBlockN / 2;                  // 128 DB rows staged per CTA
constexpr int kBBytes2 = kBHalfRows * kBlockK * 2;       // 16 KB
constexpr int kStageBytes2 = kABytes2 + kBBytes2;        // 32 KB per CTA
constexpr size_t kSmemBytes2 = 1024 + (size_t)kStages2 * kStageBytes2 + kListBytes + 256 + 1024 /*seed_topk*/;
constexpr uint32_t kIdesc2 = make_idesc(2 * kBlockM, kBlockN);   // M=256 across the pair
}  // namespace

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
score_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_db,
                 int64_t n_rows, int32_t n_kblocks, int32_t nq, int32_t n_qtiles, int32_t n_splits,
                 int32_t n_lists, int32_t max_tiles, int32_t tile_stride, const float* __restrict__ thr_floor,
                Cand* __restrict__ partial,
                 int32_t seed_k, const float* __restrict__ seed_eps, float* seed_floor, unsigned int* grid_bar) {
  extern __shared__ unsigned char smem_raw[];
  // identical carve-up in both CTAs: the MMA and the multicast commits address the peer by offset
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* ring = smem;
  float* lscore = reinterpret_cast<float*>(smem + (size_t)kStages2 * kStageBytes2);   // [32][128]
  int32_t* lrow = reinterpret_cast<int32_t*>(lscore + kList * kBlockM);
  uint64_t* full = reinterpret_cast<uint64_t*>(lrow + kList * kBlockM);   // used in the leader only
  uint64_t* empty = full + kStages2;
  uint64_t* tfull = empty + kStages2;     // [2]
  uint64_t* tempty = tfull + 2;           // [2] used in the leader only
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint32_t* seed_scratch = tmem_base_slot + 4;      // [12] in-kernel seeding: barrier verdicts
  uint64_t* seed_topk = reinterpret_cast<uint64_t*>(seed_scratch + 12);   // [4][32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();          // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int qtile = pair % n_qtiles;                     // 256 queries per pair
  const int split = pair / n_qtiles;
  const int64_t tiles_total = (n_rows + kBlockN - 1) / kBlockN;
  const int64_t tile_begin = tiles_total * split / n_splits;
  const int64_t tile_end = tiles_total * (split + 1) / n_splits;
  int n_tiles = (int)(tile_end - tile_begin);
  if (max_tiles > 0 && n_tiles > max_tiles) n_tiles = max_tiles;   // sampling pass (threshold seeding)
  const int q_row0 = qtile * 2 * kBlockM + (int)rank * kBlockM;   // first query of this CTA

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages2; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tfull[s], 1); ptx::mbar_init(&tempty[s], 8); }
    ptx::fence_mbar_init();
  }
  if (warp == 0 && lane == 0) { ptx::tma_prefetch_desc(&tmap_q); ptx::tma_prefetch_desc(&tmap_db); }
  if (warp == 1) { ptx::tmem_alloc_pair(tmem_base_slot, kTmemCols); ptx::tmem_relinquish_pair(); }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();               // barriers of both CTAs initialised, TMEM allocated
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    if (ptx::elect_one()) {
      int it = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int32_t row0 = (int32_t)((tile_begin + (int64_t)t * tile_stride) * kBlockN) + (int32_t)rank * kBHalfRows;
        for (int kb = 0; kb < n_kblocks; ++kb, ++it) {
          const int s = it % kStages2;
          ptx::mbar_wait(&empty[s], ((it / kStages2) & 1) ^ 1);
          unsigned char* a_dst = ring + (size_t)s * kStageBytes2;
          unsigned char* b_dst = a_dst + kABytes2;
          if (rank == 0) ptx::mbar_arrive_expect_tx(&full[s], 2 * kStageBytes2);   // bytes of both CTAs
          ptx::tma_load_2d_pair(a_dst, &tmap_q, kb * kBlockK, q_row0, &full[s], ptx::kEvictLast);
          ptx::tma_load_2d_pair(b_dst, &tmap_db, kb * kBlockK, row0, &full[s], ptx::kEvictNormal);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader only) ------------------------------
    if (rank == 0) {
      // every address below derives from warp-uniform bases, so ptxas keeps descriptors and barrier
      // addresses in uniform registers (no R2UR/ELECT waterfall around each UTCHMMA / UTCBAR)
      const uint32_t ring_u = __shfl_sync(0xffffffffu, ptx::smem_u32(ring), 0);
      const uint32_t full_u = __shfl_sync(0xffffffffu, ptx::smem_u32(full), 0);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t empty_u = full_u + kStages2 * 8, tfull_u = empty_u + kStages2 * 8, tempty_u = tfull_u + 16;
      const uint64_t desc0 = make_sw128_desc(ring_u);                   // stage s: + s * (kStageBytes2 >> 4)
      int s = 0;
      uint32_t ph = 0;                                                  // parity of the ring pass
      for (int t = 0; t < n_tiles; ++t) {
        const int acc = t & 1;
        ptx::mbar_wait_a(tempty_u + acc * 8, ((t >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_u + (uint32_t)(acc * kBlockN);
        for (int kb = 0; kb < n_kblocks; ++kb) {
          ptx::mbar_wait_a(full_u + s * 8, ph);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {       // elect.sync: ptxas then issues the tcgen05 instructions without a per-instruction waterfall
            const uint64_t a_desc = desc0 + (uint64_t)(s * (kStageBytes2 >> 4));
            const uint64_t b_desc = a_desc + (uint64_t)(kABytes2 >> 4);
            static_assert(kBlockK / kUmmaK == 4, "umma_bf16_pair_kblock issues four K=16 MMAs");
            // four MMAs + the commit that frees the stage in both CTAs, one asm statement (ptx.cuh)
            ptx::umma_bf16_pair_kblock(d_tmem, a_desc, b_desc, kIdesc2, kb != 0 ? 1u : 0u, empty_u + s * 8, 3);
            if (kb == n_kblocks - 1) ptx::umma_commit_pair_a(tfull_u + acc * 8, 3);   // accumulators ready in both
          }
          __syncwarp();
          if (++s == kStages2) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    // ------------------------------ epilogue (both CTAs) ------------------------------
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const int qi = q_row0 + m;
    const uint32_t s_addr = ptx::smem_u32(lscore + m);            // entry e at + e * 512 B
    const uint32_t r_addr = ptx::smem_u32(lrow + m);
    list_init(s_addr, r_addr);
    // padded query rows (qi >= nq) never insert; seeded floor = b_k(sample) - 2 eps (api.cu)
    float floor = qi < nq ? (thr_floor ? thr_floor[qi] : -INFINITY) : INFINITY;
    float thr = floor;
    int min_e = 0;
    // no floor known for any live query of this warp: build the list of the first tile in bulk
    const bool bulk_first = __all_sync(0xffffffffu, qi >= nq || floor == -INFINITY);
    for (int t = 0; t < n_tiles; ++t) {
      const int acc = t & 1;
      const int64_t row0 = (tile_begin + (int64_t)t * tile_stride) * kBlockN;
      const int valid = (int)min((int64_t)kBlockN, n_rows - row0);
      ptx::mbar_wait(&tfull[acc], (t >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kBlockN);
      if (t == 0 && bulk_first) {
        const FirstTileCodes codes = first_tile_pass1(taddr, row0, valid, s_addr, r_addr, qi < nq, floor, thr, min_e);
        // in-kernel seeding (tc_common.cuh) when every CTA of the grid is resident: one query tile, at most
        // n_sm / 2 pairs; CTA c computes the floors of queries c and c + gridDim.x
        if (seed_k > 0 && n_tiles > 1)
          seed_exchange(seed_k, seed_eps, seed_floor, grid_bar, partial, n_lists, n_splits, nq, qi, split, s_addr, r_addr,
                        /*cta_id=*/(int)blockIdx.x, /*n_ctas=*/(int)gridDim.x, seed_scratch, seed_topk,
                        (int)threadIdx.x - 64, quarter, lane, floor, thr);
        first_tile_pass2(taddr, row0, valid, s_addr, r_addr, codes, floor, thr, min_e);
      } else {
        drain_accumulator(taddr, row0, valid, s_addr, r_addr, floor, thr, min_e);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_remote(&tempty[acc], 0);          // the leader's barrier
    }
    if (qi < nq) {
      list_store(s_addr, r_addr, partial + ((int64_t)qi * n_lists + split) * kList);
    }
  }

  // neither CTA may exit (or free TMEM) while its peer can still touch its shared memory
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc_pair(tmem_base, kTmemCols); }
}

// ---------------------------------------------------------------------------------------
ScoreTcPlan score_tc2_plan(int nq, int64_t n_rows, int n_sm, int forced_splits, int min_splits) {
  ScoreTcPlan p;
  p.n_qtiles = (nq + 2 * kBlockM - 1) / (2 * kBlockM);
  const int64_t tiles_total = (n_rows + kBlockN - 1) / kBlockN;
  // One CTA pair per two SMs is resident at a time, so the grid runs in waves of n_sm/2 pairs and every
  // CTA pays one pipeline fill + one un-overlapped last epilogue (about 1.5 tile times).  With >= 128
  // tiles per CTA that is noise and more, shorter CTAs balance better (10 M rows: 148 splits best);
  // on small shards fewer splits win (1.25 M rows: 37 instead of 148 splits is +8 % at batch 4096,
  // +10 % at batch 512; profiles/experiments/r01_exp_splits.log).  Only split counts that keep
  // n_qtiles * n_splits a whole number of waves are considered.
  int splits = n_sm;
  if (forced_splits > 0) {
    splits = forced_splits;
  } else {
    const int wave = n_sm / 2 > 0 ? n_sm / 2 : 1;
    // One query tile (129..256 queries): ONE wave of pairs, whatever the split length — the grid then seeds inside the
    // launch (grid barrier among resident CTAs; no sampling pass, no seed kernel): 9.59 -> 9.31 ms at 256 queries on
    // 10 M rows, 4.97 -> 4.59 ms on 5 M (profiles/experiments/r02_exp_splits74.log).
    for (int div = p.n_qtiles == 1 ? 2 : 1; div <= 4; div <<= 1) {
      const int s = n_sm / div;
      if (s < 1 || s < min_splits || n_sm % div != 0 || ((int64_t)p.n_qtiles * s) % wave != 0) continue;
      splits = s;
      if (tiles_total / s >= 128) break;
    }
    if (splits < min_splits) splits = (min_splits + n_sm - 1) / n_sm * n_sm;   // k > 32: more lists
  }
  if ((int64_t)splits > tiles_total) splits = (int)(tiles_total > 0 ? tiles_total : 1);
  p.n_splits = splits;
  p.grid = 2 * p.n_qtiles * p.n_splits;
  return p;
}

int score_tc2_encode_db_map(void* tmap_db_out, const uint16_t* db_bf16, int64_t n_rows, int Dp) {
  return encode_2d(reinterpret_cast<CUtensorMap*>(tmap_db_out), db_bf16, (uint64_t)n_rows, (uint64_t)Dp, kBHalfRows);
}

// CTAs (2 per cluster) of this kernel that can be resident at once on the current device.
int score_tc2_max_coresident(int n_sm) {
  int clusters = 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (unsigned)n_sm, 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = kSmemBytes2;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  if (cudaFuncSetAttribute(score_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes2) != cudaSuccess ||
      cudaOccupancyMaxActiveClusters(&clusters, score_tc2_kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return 2 * clusters;
}

int launch_score_tc2(const ScoreTcArgs& a, cudaStream_t st) {
  B2K_CUDA(cudaFuncSetAttribute(score_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes2));
  const CUtensorMap* mq = reinterpret_cast<const CUtensorMap*>(a.tmap_q);
  const CUtensorMap* md = reinterpret_cast<const CUtensorMap*>(a.tmap_db);
  score_tc2_kernel<<<a.plan.grid, kThreads, kSmemBytes2, st>>>(*mq, *md, a.n_rows, a.Dp / kBlockK, a.nq,
                                                              a.plan.n_qtiles, a.plan.n_splits, a.n_lists, a.max_tiles, a.max_tiles > 0 && a.tile_stride > 1 ? a.tile_stride : 1, a.thr_floor,
                                                              a.partial, a.seed_k, a.seed_eps, a.seed_floor, a.grid_bar);
  B2K_CHECK_LAUNCH();
  return 0;
}

}  // namespace b2k
