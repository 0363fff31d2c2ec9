/* Plain-C client of the drop-in boundary (include/b2k.h): no Python, no torch.
 *
 *   gcc -O2 -Iinclude examples/c_abi_demo.c -o c_abi_demo image_recommender_b200/libb2k.so -lm
 *   LD_LIBRARY_PATH=image_recommender_b200 ./c_abi_demo [rows] [queries] [k]
 *
 * Builds a combo index (color 48 + sift 128 + dreamsim 1792) from host arrays the way
 * FAISSIndexBuilderDB hands batches to index.add (main/create_index.py:301-313), searches the way
 * ImageRecommender calls index.search (main/search_from_image.py:247), saves / reloads the index
 * (faiss.write_index / read_index), searches it once more as a row-sharded group over every GPU of the box
 * (b2k_group_*), and checks every result against a brute-force loop in C.
 * Prints "ok" and returns 0 when all of it agrees. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b2k.h"

#define CHECK(call) do { int rc_ = (call); if (rc_ != 0) { \
  fprintf(stderr, "%s:%d %s -> %d: %s\n", __FILE__, __LINE__, #call, rc_, b2k_last_error()); return 2; } } while (0)

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static float rnd(void) {      /* xorshift, uniform in (-1, 1) */
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return (float)((rng_state >> 11) & 0xFFFFFF) / 8388608.0f - 1.0f;
}

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 20000;
  const int nq = argc > 2 ? atoi(argv[2]) : 9, k = argc > 3 ? atoi(argv[3]) : 10;
  const int32_t dims[3] = {48, 128, 1792};
  const int D = 48 + 128 + 1792;
  int32_t ndev = 0;
  if (b2k_device_count(&ndev) != 0 || ndev < 1) { fprintf(stderr, "no CUDA device: %s\n", b2k_last_error()); return 3; }

  float* tab[3];
  for (int t = 0; t < 3; ++t) {
    tab[t] = (float*)malloc(sizeof(float) * (size_t)n * dims[t]);
    for (size_t i = 0; i < (size_t)n * dims[t]; ++i) tab[t][i] = t == 0 ? fabsf(rnd()) : rnd();
  }
  b2k_index* ix = NULL;
  CHECK(b2k_create(dims, 3, n, 0, 0, &ix));
  const float* host[3] = {tab[0], tab[1], tab[2]};
  CHECK(b2k_add(ix, host, n / 2));                                   /* two add() calls, like two batches */
  const float* host2[3] = {tab[0] + (n / 2) * 48, tab[1] + (n / 2) * 128, tab[2] + (n / 2) * 1792};
  CHECK(b2k_add(ix, host2, n - n / 2));
  if (b2k_ntotal(ix) != n || b2k_dim(ix) != D) { fprintf(stderr, "ntotal/dim mismatch\n"); return 4; }

  /* the stored rows (per-table normalised, concatenated) are the ground truth for the check */
  float* rows = (float*)malloc(sizeof(float) * (size_t)n * D);
  float* norm2 = (float*)malloc(sizeof(float) * (size_t)n);
  CHECK(b2k_get_rows(ix, 0, n, rows, NULL, norm2));

  /* queries: noisy copies of stored rows, whole-vector normalised (search_from_image.py:317-322) */
  float* q = (float*)malloc(sizeof(float) * (size_t)nq * D);
  for (int i = 0; i < nq; ++i)
    for (int j = 0; j < D; ++j) q[(size_t)i * D + j] = rows[(size_t)((i * 7919L) % n) * D + j] + 0.01f * rnd();
  CHECK(b2k_normalize_l2(q, nq, D, 0));

  float* dist = (float*)malloc(sizeof(float) * (size_t)nq * k);
  float* ip = (float*)malloc(sizeof(float) * (size_t)nq * k);
  int64_t* lab = (int64_t*)malloc(sizeof(int64_t) * (size_t)nq * k);
  CHECK(b2k_search(ix, q, nq, k, dist, lab, ip));

  int bad = 0;
  for (int i = 0; i < nq && !bad; ++i) {
    /* brute force in double, then compare the ranking (ties: lower offset) and the fp32 values */
    for (int j = 0; j < k; ++j) {
      const int64_t r = lab[(size_t)i * k + j];
      if (r < 0 || r >= n) { bad = 1; break; }
      double s = 0.0;
      for (int c = 0; c < D; ++c) s += (double)q[(size_t)i * D + c] * (double)rows[(size_t)r * D + c];
      if ((float)s != ip[(size_t)i * k + j]) bad = 1;                /* Spec R: fp64 dot rounded once */
      if (j > 0 && !(ip[(size_t)i * k + j - 1] > ip[(size_t)i * k + j] ||
                     (ip[(size_t)i * k + j - 1] == ip[(size_t)i * k + j] && lab[(size_t)i * k + j - 1] < r))) bad = 1;
    }
    /* nothing outside the list beats the k-th entry */
    const float kth = ip[(size_t)i * k + k - 1];
    for (long r = 0; r < n && !bad; ++r) {
      int listed = 0;
      for (int j = 0; j < k; ++j) listed |= lab[(size_t)i * k + j] == r;
      if (listed) continue;
      double s = 0.0;
      for (int c = 0; c < D; ++c) s += (double)q[(size_t)i * D + c] * (double)rows[(size_t)r * D + c];
      if ((float)s > kth) bad = 1;
    }
    if (lab[(size_t)i * k] != (i * 7919L) % n) bad = 1;              /* the source row comes first */
  }
  if (bad) { fprintf(stderr, "search result disagrees with brute force\n"); return 5; }

  /* write_index / read_index round trip */
  const char* path = "/tmp/b2k_c_abi_demo.faiss";
  CHECK(b2k_save(ix, path, NULL, 0));
  b2k_index* back = NULL;
  CHECK(b2k_load(path, 0, 0, -1, 0, &back));
  float* dist2 = (float*)malloc(sizeof(float) * (size_t)nq * k);
  int64_t* lab2 = (int64_t*)malloc(sizeof(int64_t) * (size_t)nq * k);
  CHECK(b2k_search(back, q, nq, k, dist2, lab2, NULL));
  if (memcmp(lab, lab2, sizeof(int64_t) * (size_t)nq * k) || memcmp(dist, dist2, sizeof(float) * (size_t)nq * k)) {
    fprintf(stderr, "reloaded index answers differently\n");
    return 6;
  }
  /* the same file row-sharded over every GPU of the box, driven from this one thread (b2k_group_*; two ranks on
   * device 0 when there is a single GPU): the same answers again */
  {
    const int32_t two_on_zero[2] = {0, 0};
    b2k_group* grp = NULL;
    CHECK(b2k_group_create(ndev >= 2 ? NULL : two_on_zero, ndev >= 2 ? ndev : 2, &grp));
    CHECK(b2k_group_load(grp, path));
    if (b2k_group_ntotal(grp) != n || b2k_group_dim(grp) != D) { fprintf(stderr, "group: wrong shape\n"); return 7; }
    CHECK(b2k_group_search(grp, q, nq, k, dist2, lab2, NULL));
    if (memcmp(lab, lab2, sizeof(int64_t) * (size_t)nq * k) || memcmp(dist, dist2, sizeof(float) * (size_t)nq * k)) {
      fprintf(stderr, "the %d-shard group answers differently\n", (int)b2k_group_size(grp));
      return 7;
    }
    b2k_group_destroy(grp);
  }
  b2k_stats st;
  CHECK(b2k_get_stats(ix, &st));
  printf("ok: %ld rows, %d queries, top-%d; path %d, %d uncertified, first hit offset %lld dist %.6f\n", n, nq, k, st.path,
         st.n_uncertified, (long long)lab[0], dist[0]);
  b2k_destroy(back);
  b2k_destroy(ix);
  remove(path);
  return 0;
}
