// K-score: batched approximate scoring S = Q · Xᵀ on the 5th-gen tensor cores with the
// running top-32 selection fused into the epilogue, so the [nq, n_rows] score matrix never
// reaches HBM.  Replaces the contraction inside index.search(query_vec, top_k)
// (main/search_from_image.py:247) for query batches (a new-build extension: the reference
// is batch-1).
//
// Work decomposition: CTA = (query tile of 128, DB split).  blockIdx = split * n_qtiles + qtile
// so that CTAs resident at the same time share DB rows through L2.  Per CTA:
//   warp 0   TMA producer: per K-block of 64 one box of the query tile [128 x 64] and one box of
//            the DB tile [256 x 64] (bf16, 128-byte swizzle) into a 4-stage shared-memory ring
//   warp 1   allocates TMEM (512 columns = two 128x256 fp32 accumulators) and issues
//            tcgen05.mma.cta_group::1.kind::f16  M=128 N=256 K=16, 4 per K-block
//   warps 2-5 epilogue: tcgen05.ld 32x32b (thread = query row, 32 DB columns per load),
//            per-thread threshold filter, per-thread top-32 list in shared memory
// Pipelines: smem full/empty (TMA <-> MMA) and TMEM full/empty (MMA <-> epilogue), all mbarrier.
#include "tc_common.cuh"

namespace b2k {

using namespace tc;

namespace {
constexpr int kStages = 4;
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr int kBBytes = kBlockN * kBlockK * 2;   // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr size_t kSmemBytes = 1024 /*align slack*/ + (size_t)kStages * kStageBytes + kListBytes + 256 + 1024 /*seed_topk*/;
constexpr uint32_t kIdesc = make_idesc(kBlockM, kBlockN);
}  // namespace

#ifdef B2K_PHASE_TIMERS
__device__ unsigned long long g_tc_phase_t[2][8];
#define B2K_TC_PHASE(i) do { if ((blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && threadIdx.x == 64) { unsigned long long t_; \
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_tc_phase_t[blockIdx.x == 0 ? 0 : 1][i] = t_; } } while (0)
extern "C" int b2k_debug_tc_phase_times(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_tc_phase_t, sizeof(g_tc_phase_t));
}
#else
#define B2K_TC_PHASE(i) do { } while (0)
#endif

__global__ void __launch_bounds__(kThreads, 1)
score_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_db,
                int64_t n_rows, int32_t n_kblocks, int32_t nq, int32_t n_qtiles, int32_t n_splits,
                int32_t n_lists, int32_t max_tiles, int32_t tile_stride, const float* __restrict__ thr_floor,
                Cand* __restrict__ partial,
                int32_t seed_k, const float* __restrict__ seed_eps, float* seed_floor,
                unsigned int* grid_bar) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* ring = smem;
  float* lscore = reinterpret_cast<float*>(smem + (size_t)kStages * kStageBytes);   // [32][128]
  int32_t* lrow = reinterpret_cast<int32_t*>(lscore + kList * kBlockM);             // [32][128]
  uint64_t* full = reinterpret_cast<uint64_t*>(lrow + kList * kBlockM);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;      // [2] accumulator ready
  uint64_t* tempty = tfull + 2;           // [2] accumulator drained
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint32_t* seed_scratch = tmem_base_slot + 4;      // [12] in-kernel seeding: barrier verdicts
  uint64_t* seed_topk = reinterpret_cast<uint64_t*>(seed_scratch + 12);   // [4][32] per-warp top-k of the handler

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qtile = blockIdx.x % n_qtiles;
  const int split = blockIdx.x / n_qtiles;
  const int64_t tiles_total = (n_rows + kBlockN - 1) / kBlockN;
  const int64_t tile_begin = tiles_total * split / n_splits;
  const int64_t tile_end = tiles_total * (split + 1) / n_splits;
  int n_tiles = (int)(tile_end - tile_begin);
  if (max_tiles > 0 && n_tiles > max_tiles) n_tiles = max_tiles;   // sampling pass (threshold seeding)

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tfull[s], 1); ptx::mbar_init(&tempty[s], 4); }
    ptx::fence_mbar_init();
  }
  if (warp == 0 && lane == 0) { ptx::tma_prefetch_desc(&tmap_q); ptx::tma_prefetch_desc(&tmap_db); }
  if (warp == 1) { ptx::tmem_alloc(tmem_base_slot, kTmemCols); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (ptx::elect_one()) {
      int it = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int32_t row0 = (int32_t)((tile_begin + (int64_t)t * tile_stride) * kBlockN);
        for (int kb = 0; kb < n_kblocks; ++kb, ++it) {
          const int s = it % kStages;
          ptx::mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
          unsigned char* a_dst = ring + (size_t)s * kStageBytes;
          unsigned char* b_dst = a_dst + kABytes;
          ptx::mbar_arrive_expect_tx(&full[s], kStageBytes);
          ptx::tma_load_2d(a_dst, &tmap_q, kb * kBlockK, qtile * kBlockM, &full[s], ptx::kEvictLast);
          ptx::tma_load_2d(b_dst, &tmap_db, kb * kBlockK, row0, &full[s], ptx::kEvictNormal);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    // every address below derives from warp-uniform bases, so ptxas keeps descriptors and barrier
    // addresses in uniform registers (no R2UR/ELECT waterfall around each UTCHMMA / UTCBAR)
    const uint32_t ring_u = __shfl_sync(0xffffffffu, ptx::smem_u32(ring), 0);
    const uint32_t full_u = __shfl_sync(0xffffffffu, ptx::smem_u32(full), 0);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t empty_u = full_u + kStages * 8, tfull_u = empty_u + kStages * 8, tempty_u = tfull_u + 16;
    const uint64_t desc0 = make_sw128_desc(ring_u);                     // stage s: + s * (kStageBytes >> 4)
    int s = 0;
    uint32_t ph = 0;                                                    // parity of the ring pass
    for (int t = 0; t < n_tiles; ++t) {
      const int acc = t & 1;
      ptx::mbar_wait_a(tempty_u + acc * 8, ((t >> 1) & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_u + (uint32_t)(acc * kBlockN);
      for (int kb = 0; kb < n_kblocks; ++kb) {
        ptx::mbar_wait_a(full_u + s * 8, ph);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {       // elect.sync: ptxas then issues the tcgen05 instructions without a per-instruction waterfall
          const uint64_t a_desc = desc0 + (uint64_t)(s * (kStageBytes >> 4));
          const uint64_t b_desc = a_desc + (uint64_t)(kABytes >> 4);
          static_assert(kBlockK / kUmmaK == 4, "umma_bf16_kblock issues four K=16 MMAs");
          // four MMAs (descriptors advance 32 bytes = +2 inside the swizzled row) + the commit that frees the stage,
          // one asm statement (ptx.cuh)
          ptx::umma_bf16_kblock(d_tmem, a_desc, b_desc, kIdesc, kb != 0 ? 1u : 0u, empty_u + s * 8);
          if (kb == n_kblocks - 1) ptx::umma_commit_a(tfull_u + acc * 8);   // accumulator complete
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ------------------------------ epilogue ------------------------------
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may read
    const int m = quarter * 32 + lane;               // query row inside the tile
    const int qi = qtile * kBlockM + m;
    const uint32_t s_addr = ptx::smem_u32(lscore + m);            // entry e at + e * 512 B
    const uint32_t r_addr = ptx::smem_u32(lrow + m);
    list_init(s_addr, r_addr);
    // padded query rows (qi >= nq) never insert; seeded floor = b_k(sample) - 2 eps (api.cu)
    float floor = qi < nq ? (thr_floor ? thr_floor[qi] : -INFINITY) : INFINITY;
    float thr = floor;
    int min_e = 0;
    // no floor known for any live query of this warp: build the list of the first tile in bulk
    const bool bulk_first = __all_sync(0xffffffffu, qi >= nq || floor == -INFINITY);
    const int et = threadIdx.x - 64;                               // 0..127 among the epilogue threads
    B2K_TC_PHASE(0);
    for (int t = 0; t < n_tiles; ++t) {
      const int acc = t & 1;
      const int64_t row0 = (tile_begin + (int64_t)t * tile_stride) * kBlockN;
      const int valid = (int)min((int64_t)kBlockN, n_rows - row0);
      ptx::mbar_wait(&tfull[acc], (t >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kBlockN);
      if (t == 0 && bulk_first) {
        const FirstTileCodes codes = first_tile_pass1(taddr, row0, valid, s_addr, r_addr, qi < nq, floor, thr, min_e);
        if (seed_k > 0 && n_tiles > 1) {
          B2K_TC_PHASE(1);
          seed_exchange(seed_k, seed_eps, seed_floor, grid_bar, partial, n_lists, n_splits, nq, qi, split, s_addr, r_addr,
                        /*cta_id=*/split, /*n_ctas=*/n_splits, seed_scratch, seed_topk, et, quarter, lane, floor, thr);
          B2K_TC_PHASE(5);
        }
        first_tile_pass2(taddr, row0, valid, s_addr, r_addr, codes, floor, thr, min_e);
      } else {
        drain_accumulator(taddr, row0, valid, s_addr, r_addr, floor, thr, min_e);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
    }
    B2K_TC_PHASE(6);
    if (qi < nq) {
      list_store(s_addr, r_addr, partial + ((int64_t)qi * n_lists + split) * kList);
    }
    B2K_TC_PHASE(7);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, kTmemCols); }
}

// ---------------------------------------------------------------------------------------
int score_tc_tile_rows() { return kBlockN; }
bool score_tc_supports(int Dp) { return Dp >= kBlockK && (Dp % kBlockK) == 0; }

ScoreTcPlan score_tc_plan(int nq, int64_t n_rows, int n_sm, int forced_splits, int min_splits) {
  ScoreTcPlan p;
  p.n_qtiles = (nq + kBlockM - 1) / kBlockM;
  const int64_t tiles_total = (n_rows + kBlockN - 1) / kBlockN;
  // one CTA per SM is resident.  Long splits (>= 128 tiles): one split per SM, the query tiles of a
  // split run back to back.  Short splits: a single wave of n_qtiles * n_splits <= n_sm CTAs, so the
  // per-CTA pipeline fill / last epilogue is paid once (the query tiles of a split then run
  // concurrently and share the DB tiles through L2).
  int splits = n_sm;
  if (forced_splits > 0) splits = forced_splits;
  else if (p.n_qtiles > 1 && tiles_total / n_sm < 128) splits = n_sm / p.n_qtiles > 0 ? n_sm / p.n_qtiles : 1;
  if (forced_splits <= 0 && splits < min_splits) splits = (min_splits + n_sm - 1) / n_sm * n_sm;   // k > 32: more lists
  if ((int64_t)splits > tiles_total) splits = (int)(tiles_total > 0 ? tiles_total : 1);
  p.n_splits = splits;
  p.grid = p.n_qtiles * p.n_splits;
  return p;
}

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

int tc::encode_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return B2K_E_NODEVICE; }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return B2K_E_INVALID; }
  return 0;
}

int score_tc_encode_maps(void* tmap_q_out, void* tmap_db_out, const uint16_t* q_bf16, int nq_pad,
                         const uint16_t* db_bf16, int64_t n_rows, int Dp) {
  if (tmap_q_out) {
    int rc = encode_2d(reinterpret_cast<CUtensorMap*>(tmap_q_out), q_bf16, (uint64_t)nq_pad, (uint64_t)Dp, kBlockM);
    if (rc) return rc;
  }
  if (tmap_db_out) {
    int rc = encode_2d(reinterpret_cast<CUtensorMap*>(tmap_db_out), db_bf16, (uint64_t)n_rows, (uint64_t)Dp, kBlockN);
    if (rc) return rc;
  }
  return 0;
}

// CTAs of this kernel that can be resident at once on the current device (occupancy query): the in-kernel
// seeding's grid barrier is only enabled for grids within it.
int score_tc_max_coresident(int n_sm) {
  int per_sm = 0;
  if (cudaFuncSetAttribute(score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, score_tc_kernel, kThreads, kSmemBytes) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return per_sm * n_sm;
}

int launch_score_tc(const ScoreTcArgs& a, cudaStream_t st) {
  B2K_CUDA(cudaFuncSetAttribute(score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  const CUtensorMap* mq = reinterpret_cast<const CUtensorMap*>(a.tmap_q);
  const CUtensorMap* md = reinterpret_cast<const CUtensorMap*>(a.tmap_db);
  score_tc_kernel<<<a.plan.grid, kThreads, kSmemBytes, st>>>(*mq, *md, a.n_rows, a.Dp / kBlockK, a.nq,
                                                            a.plan.n_qtiles, a.plan.n_splits, a.n_lists, a.max_tiles, a.max_tiles > 0 && a.tile_stride > 1 ? a.tile_stride : 1, a.thr_floor,
                                                            a.partial, a.seed_k, a.seed_eps, a.seed_floor, a.grid_bar);
  B2K_CHECK_LAUNCH();
  return 0;
}

}  // namespace b2k
