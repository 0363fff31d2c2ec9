/*
 * cpu_flat.c — the result-handler half of faiss_cpu 1.10.0's IndexFlatIP.search, restated
 * (BENCH/TEST INFRASTRUCTURE ONLY; never linked or imported by the product).
 *
 * faiss (faiss/utils/distances.cpp, not vendored in /root/reference; call site
 * main/search_from_image.py:247 via IndexFlat*) answers nq >= 20 queries with
 *   exhaustive_inner_product_blas: for query blocks of 4096 and database blocks of 1024 rows
 *   one sgemm, then `res.add_results(j0, j1, ip_block)`: an OpenMP-parallel loop over the
 *   queries of the block, each pushing the block's scores into its own k-entry min-heap
 *   (HeapBlockResultHandler<CMin<float, int64_t>>), and a final per-query heap reorder;
 * and nq < 20 queries with
 *   exhaustive_inner_product_seq: an OpenMP loop over the queries, each scanning every row
 *   with a SIMD dot product and pushing into its heap.
 * The sgemm itself is called from Python (numpy -> multi-threaded OpenBLAS, the same library
 * class faiss links); this file is everything else, so that the port scales with the host's
 * cores the way faiss does (round 1 used single-threaded numpy argpartition here).
 *
 * Ordering of exact fp32 near-ties follows the BLAS summation order and the heap's
 * replacement order, as in faiss; it is NOT the bit-exact Spec R oracle (b2k_oracle.c) and is
 * never used for parity.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <omp.h>

/* min-heap on (val, id): root = the worst kept entry (faiss CMin: smaller val is worse; on equal
 * values faiss's heap_replace_top compares ids too, larger id = worse for CMin) */
static inline int worse(float va, int64_t ia, float vb, int64_t ib) {
  return va < vb || (va == vb && ia > ib);
}

static inline void heap_replace_top(int k, float* hv, int64_t* hi, float v, int64_t id) {
  int i = 0;
  for (;;) {
    int l = 2 * i + 1, r = l + 1, m;
    if (l >= k) break;
    m = (r < k && worse(hv[r], hi[r], hv[l], hi[l])) ? r : l;
    if (!worse(hv[m], hi[m], v, id)) break;
    hv[i] = hv[m]; hi[i] = hi[m];
    i = m;
  }
  hv[i] = v; hi[i] = id;
}

/* heaps start as k x (-inf, -1): a valid min-heap */
void flat_heap_init(float* hv, int64_t* hi, int64_t n) {
  for (int64_t i = 0; i < n; ++i) { hv[i] = -INFINITY; hi[i] = -1; }
}

/* res.add_results(j0, j0 + nb, ip): ip is [nq, nb] row-major with leading dimension ld */
void flat_heap_addn(const float* ip, int32_t nq, int32_t nb, int64_t ld, int64_t j0, int32_t k,
                    float* hv, int64_t* hi) {
#pragma omp parallel for schedule(static)
  for (int32_t q = 0; q < nq; ++q) {
    float* v = hv + (int64_t)q * k;
    int64_t* id = hi + (int64_t)q * k;
    const float* row = ip + (int64_t)q * ld;
    float thr = v[0];
    for (int32_t j = 0; j < nb; ++j) {
      const float s = row[j];
      if (s > thr) {                      /* CMin::cmp(heap top, s) */
        heap_replace_top(k, v, id, s, j0 + j);
        thr = v[0];
      }
    }
  }
}

/* exhaustive_inner_product_seq: one thread per query, every row */
void flat_scan_seq(const float* db, int64_t n, int32_t d, const float* q, int32_t nq, int32_t k,
                   float* hv, int64_t* hi) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int32_t qi = 0; qi < nq; ++qi) {
    const float* x = q + (int64_t)qi * d;
    float* v = hv + (int64_t)qi * k;
    int64_t* id = hi + (int64_t)qi * k;
    float thr = v[0];
    for (int64_t r = 0; r < n; ++r) {
      const float* y = db + r * d;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f, a6 = 0.f, a7 = 0.f;
      int32_t i = 0;
      for (; i + 8 <= d; i += 8) {       /* 8 independent lanes: vectorises to one AVX2 FMA stream */
        a0 += x[i] * y[i]; a1 += x[i + 1] * y[i + 1]; a2 += x[i + 2] * y[i + 2]; a3 += x[i + 3] * y[i + 3];
        a4 += x[i + 4] * y[i + 4]; a5 += x[i + 5] * y[i + 5]; a6 += x[i + 6] * y[i + 6]; a7 += x[i + 7] * y[i + 7];
      }
      float s = ((a0 + a4) + (a1 + a5)) + ((a2 + a6) + (a3 + a7));
      for (; i < d; ++i) s += x[i] * y[i];
      if (s > thr) { heap_replace_top(k, v, id, s, r); thr = v[0]; }
    }
  }
}

/* heap -> descending (score, then ascending id); unfilled slots (-1) last, as faiss's reorder */
static int cmp_desc(const void* a, const void* b) {
  const float va = ((const float*)a)[0], vb = ((const float*)b)[0];
  int64_t ia, ib;
  __builtin_memcpy(&ia, (const char*)a + 8, 8);
  __builtin_memcpy(&ib, (const char*)b + 8, 8);
  if ((ia < 0) != (ib < 0)) return ia < 0 ? 1 : -1;
  if (va != vb) return va > vb ? -1 : 1;
  return ia < ib ? -1 : (ia > ib ? 1 : 0);
}

void flat_heap_reorder(int32_t nq, int32_t k, float* hv, int64_t* hi) {
#pragma omp parallel
  {
    char* tmp = (char*)malloc((size_t)k * 16);
#pragma omp for schedule(static)
    for (int32_t q = 0; q < nq; ++q) {
      float* v = hv + (int64_t)q * k;
      int64_t* id = hi + (int64_t)q * k;
      for (int32_t j = 0; j < k; ++j) {
        __builtin_memcpy(tmp + (size_t)j * 16, &v[j], 4);
        __builtin_memcpy(tmp + (size_t)j * 16 + 8, &id[j], 8);
      }
      qsort(tmp, (size_t)k, 16, cmp_desc);
      for (int32_t j = 0; j < k; ++j) {
        __builtin_memcpy(&v[j], tmp + (size_t)j * 16, 4);
        __builtin_memcpy(&id[j], tmp + (size_t)j * 16 + 8, 8);
      }
    }
    free(tmp);
  }
}

int32_t flat_num_threads(void) { return omp_get_max_threads(); }
void flat_set_num_threads(int32_t n) { if (n > 0) omp_set_num_threads(n); }
