/*
 * hnsw_ref.c — CPU restatement of the reference's approximate index, for the "recall@k of the
 * reference HNSW index" column of the bench.  TEST/BENCH INFRASTRUCTURE ONLY: never
 * linked into the product.
 *
 * The reference builds faiss.IndexHNSWFlat(dim, M=32) with hnsw.efConstruction = 200 and
 * hnsw.efSearch = 64 (ctor) / 50 (__main__), METRIC_L2 (main/create_index.py:20-22, 229-234,
 * 336-339) and queries it with index.search (main/search_from_image.py:247).  faiss_cpu==1.10.0 is
 * an un-vendored, absent dependency, so this file restates the published algorithm faiss
 * implements (Malkov & Yashunin, "Efficient and robust approximate nearest neighbor search using
 * Hierarchical Navigable Small World graphs"; faiss/impl/HNSW.cpp): exponential level assignment
 * with mult = 1/ln(M), 2M links on level 0 and M above, construction search with efConstruction,
 * the diversity heuristic for neighbour selection and for shrinking overfull lists, greedy descent
 * through the upper levels and a best-first level-0 search bounded by efSearch.  PARITY UNPINNED:
 * results are statistically, not bitwise, those of faiss (random levels, insertion interleaving).
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int32_t n, d, M, M0, efc;
  const float* x;            /* [n, d] borrowed */
  int32_t* level;            /* [n] */
  int32_t* links0;           /* [n, M0 + 1]: count, ids */
  int32_t** linksu;          /* [n] -> [level, M + 1] or NULL */
  int32_t entry, maxlevel;
  omp_lock_t* locks;         /* [n] */
  omp_lock_t glock;
} hnsw_t;

static inline float l2sq(const float* a, const float* b, int d) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int i = 0;
  for (; i + 32 <= d; i += 32) {
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    for (int j = 0; j < 8; ++j) {
      float e0 = a[i + j] - b[i + j], e1 = a[i + 8 + j] - b[i + 8 + j];
      float e2 = a[i + 16 + j] - b[i + 16 + j], e3 = a[i + 24 + j] - b[i + 24 + j];
      t0 += e0 * e0; t1 += e1 * e1; t2 += e2 * e2; t3 += e3 * e3;
    }
    s0 += t0; s1 += t1; s2 += t2; s3 += t3;
  }
  for (; i < d; ++i) { float e = a[i] - b[i]; s0 += e * e; }
  return (s0 + s1) + (s2 + s3);
}

/* ---- binary heaps of (dist, id) ---------------------------------------------------------- */
typedef struct { float d; int32_t id; } item_t;
typedef struct { item_t* a; int n, cap; } heap_t;

static void heap_init(heap_t* h, int cap) { h->a = (item_t*)malloc(sizeof(item_t) * (size_t)cap); h->n = 0; h->cap = cap; }
static void heap_free(heap_t* h) { free(h->a); }
static void heap_reserve(heap_t* h) {
  if (h->n == h->cap) { h->cap *= 2; h->a = (item_t*)realloc(h->a, sizeof(item_t) * (size_t)h->cap); }
}
/* max-heap if sign = +1 (largest dist on top), min-heap if sign = -1 */
static void heap_push(heap_t* h, float d, int32_t id, float sign) {
  heap_reserve(h);
  int i = h->n++;
  while (i > 0) {
    int p = (i - 1) >> 1;
    if (sign * h->a[p].d >= sign * d) break;
    h->a[i] = h->a[p]; i = p;
  }
  h->a[i].d = d; h->a[i].id = id;
}
static item_t heap_pop(heap_t* h, float sign) {
  item_t top = h->a[0], last = h->a[--h->n];
  int i = 0;
  for (;;) {
    int c = 2 * i + 1;
    if (c >= h->n) break;
    if (c + 1 < h->n && sign * h->a[c + 1].d > sign * h->a[c].d) ++c;
    if (sign * last.d >= sign * h->a[c].d) break;
    h->a[i] = h->a[c]; i = c;
  }
  if (h->n > 0) h->a[i] = last;
  return top;
}

static inline int32_t* links_of(const hnsw_t* g, int32_t v, int lev) {
  return lev == 0 ? g->links0 + (size_t)v * (g->M0 + 1) : g->linksu[v] + (size_t)(lev - 1) * (g->M + 1);
}
static inline int cap_of(const hnsw_t* g, int lev) { return lev == 0 ? g->M0 : g->M; }

/* best-first search on one level; results (<= ef closest) left in `res` (max-heap) */
static void search_layer(const hnsw_t* g, const float* q, int32_t ep, float ep_d, int ef, int lev,
                         uint32_t* visited, uint32_t tag, heap_t* cand, heap_t* res, int locked) {
  cand->n = 0; res->n = 0;
  heap_push(cand, ep_d, ep, -1.f);
  heap_push(res, ep_d, ep, +1.f);
  visited[ep] = tag;
  int32_t nb[512];
  while (cand->n > 0) {
    item_t c = heap_pop(cand, -1.f);
    if (res->n >= ef && c.d > res->a[0].d) break;
    int32_t* l = links_of(g, c.id, lev);
    int cnt;
    if (locked) omp_set_lock(&g->locks[c.id]);
    cnt = l[0];
    memcpy(nb, l + 1, sizeof(int32_t) * (size_t)cnt);
    if (locked) omp_unset_lock(&g->locks[c.id]);
    for (int j = 0; j < cnt; ++j) {
      const int32_t v = nb[j];
      if (visited[v] == tag) continue;
      visited[v] = tag;
      const float dv = l2sq(q, g->x + (size_t)v * g->d, g->d);
      if (res->n < ef || dv < res->a[0].d) {
        heap_push(cand, dv, v, -1.f);
        heap_push(res, dv, v, +1.f);
        if (res->n > ef) heap_pop(res, +1.f);
      }
    }
  }
}

/* diversity heuristic: from candidates sorted by distance to the base point, keep c iff it is closer
 * to the base than to every already kept neighbour; at most m are kept */
static int select_heuristic(const hnsw_t* g, item_t* sorted, int n, int m, int32_t* out) {
  int k = 0;
  for (int i = 0; i < n && k < m; ++i) {
    const float* xc = g->x + (size_t)sorted[i].id * g->d;
    int good = 1;
    for (int j = 0; j < k; ++j) {
      if (l2sq(xc, g->x + (size_t)out[j] * g->d, g->d) < sorted[i].d) { good = 0; break; }
    }
    if (good) out[k++] = sorted[i].id;
  }
  return k;
}

static int cmp_item(const void* a, const void* b) {
  const item_t* x = (const item_t*)a; const item_t* y = (const item_t*)b;
  return (x->d > y->d) - (x->d < y->d);
}

static void add_link(hnsw_t* g, int32_t src, int32_t dst, int lev) {
  omp_set_lock(&g->locks[src]);
  int32_t* l = links_of(g, src, lev);
  const int cap = cap_of(g, lev);
  if (l[0] < cap) {
    l[1 + l[0]++] = dst;
  } else {
    item_t tmp[513];
    const float* xs = g->x + (size_t)src * g->d;
    for (int j = 0; j < cap; ++j) { tmp[j].id = l[1 + j]; tmp[j].d = l2sq(xs, g->x + (size_t)l[1 + j] * g->d, g->d); }
    tmp[cap].id = dst; tmp[cap].d = l2sq(xs, g->x + (size_t)dst * g->d, g->d);
    qsort(tmp, (size_t)cap + 1, sizeof(item_t), cmp_item);
    l[0] = select_heuristic(g, tmp, cap + 1, cap, l + 1);
  }
  omp_unset_lock(&g->locks[src]);
}

hnsw_t* hnsw_build(const float* x, int32_t n, int32_t d, int32_t M, int32_t efc, uint64_t seed) {
  hnsw_t* g = (hnsw_t*)calloc(1, sizeof(hnsw_t));
  g->n = n; g->d = d; g->M = M; g->M0 = 2 * M; g->efc = efc; g->x = x;
  g->level = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  g->links0 = (int32_t*)calloc((size_t)n * (g->M0 + 1), sizeof(int32_t));
  g->linksu = (int32_t**)calloc((size_t)n, sizeof(int32_t*));
  g->locks = (omp_lock_t*)malloc(sizeof(omp_lock_t) * (size_t)n);
  omp_init_lock(&g->glock);
  const double mult = 1.0 / log((double)M);
  uint64_t s = seed ? seed : 0x9E3779B97F4A7C15ull;
  for (int32_t i = 0; i < n; ++i) {
    omp_init_lock(&g->locks[i]);
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;                     /* xorshift64 */
    const double u = ((double)(s >> 11) + 1.0) / 9007199254740993.0;
    int lev = (int)(-log(u) * mult);
    if (lev > 16) lev = 16;
    g->level[i] = lev;
    if (lev > 0) g->linksu[i] = (int32_t*)calloc((size_t)lev * (M + 1), sizeof(int32_t));
  }
  g->entry = 0; g->maxlevel = g->level[0];
  if (n <= 1) return g;
#pragma omp parallel
  {
    uint32_t* visited = (uint32_t*)calloc((size_t)n, sizeof(uint32_t));
    uint32_t tag = 0;
    heap_t cand, res;
    heap_init(&cand, 1024); heap_init(&res, 1024);
    item_t* sorted = (item_t*)malloc(sizeof(item_t) * (size_t)(efc + 2));
    int32_t sel[512];
#pragma omp for schedule(dynamic, 64)
    for (int32_t i = 1; i < n; ++i) {
      const float* q = x + (size_t)i * d;
      const int lev = g->level[i];
      omp_set_lock(&g->glock);
      int32_t ep = g->entry; int maxl = g->maxlevel;
      const int promote = lev > maxl;
      if (!promote) omp_unset_lock(&g->glock);                  /* a new top node inserts under the lock */
      float ep_d = l2sq(q, x + (size_t)ep * d, d);
      for (int l = maxl; l > lev; --l) {                         /* greedy descent */
        int changed = 1;
        while (changed) {
          changed = 0;
          int32_t nb[512];
          omp_set_lock(&g->locks[ep]);
          int32_t* ll = links_of(g, ep, l);
          const int cnt = ll[0];
          memcpy(nb, ll + 1, sizeof(int32_t) * (size_t)cnt);
          omp_unset_lock(&g->locks[ep]);
          for (int j = 0; j < cnt; ++j) {
            const float dv = l2sq(q, x + (size_t)nb[j] * d, d);
            if (dv < ep_d) { ep_d = dv; ep = nb[j]; changed = 1; }
          }
        }
      }
      for (int l = lev < maxl ? lev : maxl; l >= 0; --l) {
        ++tag;
        search_layer(g, q, ep, ep_d, efc, l, visited, tag, &cand, &res, 1);
        int m = res.n;
        for (int j = 0; j < m; ++j) sorted[j] = res.a[j];
        qsort(sorted, (size_t)m, sizeof(item_t), cmp_item);
        const int k = select_heuristic(g, sorted, m, cap_of(g, l), sel);   /* faiss: nb_neighbors(level) = 2M at level 0 */
        omp_set_lock(&g->locks[i]);
        int32_t* li = links_of(g, i, l);
        li[0] = k;
        memcpy(li + 1, sel, sizeof(int32_t) * (size_t)k);
        omp_unset_lock(&g->locks[i]);
        for (int j = 0; j < k; ++j) add_link(g, sel[j], i, l);
        ep = sorted[0].id; ep_d = sorted[0].d;
      }
      if (promote) { g->entry = i; g->maxlevel = lev; omp_unset_lock(&g->glock); }
    }
    heap_free(&cand); heap_free(&res); free(sorted); free(visited);
  }
  return g;
}

/* squared-L2 k-NN for nq queries (OpenMP over queries, as faiss does); labels -1 padded */
void hnsw_search(const hnsw_t* g, const float* q, int32_t nq, int32_t k, int32_t efs, float* out_d, int64_t* out_i) {
#pragma omp parallel
  {
    uint32_t* visited = (uint32_t*)calloc((size_t)g->n, sizeof(uint32_t));
    uint32_t tag = 0;
    heap_t cand, res;
    heap_init(&cand, 1024); heap_init(&res, 1024);
    const int ef = efs > k ? efs : k;
#pragma omp for schedule(dynamic, 4)
    for (int32_t qi = 0; qi < nq; ++qi) {
      const float* qv = q + (size_t)qi * g->d;
      int32_t ep = g->entry;
      float ep_d = l2sq(qv, g->x + (size_t)ep * g->d, g->d);
      for (int l = g->maxlevel; l > 0; --l) {
        int changed = 1;
        while (changed) {
          changed = 0;
          const int32_t* ll = links_of(g, ep, l);
          for (int j = 0; j < ll[0]; ++j) {
            const float dv = l2sq(qv, g->x + (size_t)ll[1 + j] * g->d, g->d);
            if (dv < ep_d) { ep_d = dv; ep = ll[1 + j]; changed = 1; }
          }
        }
      }
      ++tag;
      search_layer(g, qv, ep, ep_d, ef, 0, visited, tag, &cand, &res, 0);
      while (res.n > k) heap_pop(&res, +1.f);
      const int m = res.n;
      for (int j = m - 1; j >= 0; --j) {
        item_t it = heap_pop(&res, +1.f);
        out_d[(size_t)qi * k + j] = it.d; out_i[(size_t)qi * k + j] = it.id;
      }
      for (int j = m; j < k; ++j) { out_d[(size_t)qi * k + j] = 3.402823466e38f; out_i[(size_t)qi * k + j] = -1; }
    }
    heap_free(&cand); heap_free(&res); free(visited);
  }
}

void hnsw_free(hnsw_t* g) {
  if (!g) return;
  for (int32_t i = 0; i < g->n; ++i) { free(g->linksu[i]); omp_destroy_lock(&g->locks[i]); }
  omp_destroy_lock(&g->glock);
  free(g->level); free(g->links0); free(g->linksu); free(g->locks); free(g);
}
