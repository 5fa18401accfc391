/*
 * ivf_oracle.c — CPU restatement of the faiss algorithms on the abstracts-search hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under abstracts-search_b200/ may link, import or call this
 * file; it is the checker used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.
 *
 * PARITY UNPINNED: the algorithm lives in faiss (un-vendored, unpinned, transitive dependency of
 * sidecar-search @0.3.0 — /root/reference/requirements.txt:1); the reference tree holds no test,
 * golden vector or fixture for it and faiss cannot be installed offline.  The
 * functions below restate the published faiss behaviour at the reference's call sites:
 *   Index.search  -> /root/reference/Makefile:31-32 (tune), README.md:16,28 (app.py)
 *   Index.add     -> /root/reference/Makefile:24-25 (fill)
 *   Index.train   -> /root/reference/Makefile:38-39 (train; nlist 65536 README.md:60)
 *
 * Restated faiss pieces (names are faiss's, code is ours):
 *   fvec_inner_product           scalar/SIMD dot product
 *   IndexFlatIP::search          all-pairs IP + top-k
 *   IVFFlatScanner::scan_codes   for every vector of every probed list: ip; keep if better than heap min
 *   heap_replace_top/heap_reorder  k-best container, emitted best first, padded with id -1 / -FLT_MAX
 *   rand_perm + RandomGenerator  Fisher-Yates driven by std::mt19937: i2 = i + mt() % (n - i)
 *   Clustering::{compute_centroids, split_clusters}
 *
 * Tie rule (documented deviation): results are ordered by (score desc, id asc)
 * and the k-th place is decided by the same total order, which makes the answer independent of
 * scan order.  faiss keeps the first-seen candidate on exact score ties; the two agree whenever
 * ties happen inside one list with ids in insertion order.
 */
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ mt19937 ---------------- */
typedef struct {
  uint32_t mt[624];
  int idx;
} orc_mt_t;

static void mt_seed(orc_mt_t* s, uint32_t seed) {
  s->mt[0] = seed;
  for (int i = 1; i < 624; i++)
    s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
  s->idx = 624;
}

static uint32_t mt_next(orc_mt_t* s) {
  if (s->idx >= 624) {
    for (int i = 0; i < 624; i++) {
      uint32_t y = (s->mt[i] & 0x80000000u) | (s->mt[(i + 1) % 624] & 0x7fffffffu);
      uint32_t v = s->mt[(i + 397) % 624] ^ (y >> 1);
      if (y & 1u) v ^= 0x9908b0dfu;
      s->mt[i] = v;
    }
    s->idx = 0;
  }
  uint32_t y = s->mt[s->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

/* raw outputs, for pinning against numpy's RandomState */
void orc_mt_raw(uint32_t seed, int64_t n, uint32_t* out) {
  orc_mt_t s;
  mt_seed(&s, seed);
  for (int64_t i = 0; i < n; i++) out[i] = mt_next(&s);
}

/* faiss rand_perm(perm, n, seed): perm of 0..n-1 (faiss/utils/random.cpp) */
void orc_rand_perm(int64_t n, int64_t seed, int32_t* perm) {
  orc_mt_t s;
  mt_seed(&s, (uint32_t)seed);
  for (int64_t i = 0; i < n; i++) perm[i] = (int32_t)i;
  for (int64_t i = 0; i + 1 < n; i++) {
    /* RandomGenerator::rand_int(max) = mt() % max, with max an int */
    int64_t i2 = i + (int64_t)(mt_next(&s) % (uint32_t)(n - i));
    int32_t t = perm[i];
    perm[i] = perm[i2];
    perm[i2] = t;
  }
}

/* ------------------------------------------------------------------ k-best container ------- */
/* "worse" in the total order (score desc, id asc): a is worse than b */
static inline int worse(float sa, int64_t ia, float sb, int64_t ib) {
  return (sa < sb) || (sa == sb && ia > ib);
}

/* binary heap whose root is the WORST kept candidate */
static void heap_sift_down(int k, float* hs, int64_t* hi, int i) {
  for (;;) {
    int l = 2 * i + 1, r = l + 1, w = i;
    if (l < k && worse(hs[l], hi[l], hs[w], hi[w])) w = l;
    if (r < k && worse(hs[r], hi[r], hs[w], hi[w])) w = r;
    if (w == i) return;
    float ts = hs[i]; hs[i] = hs[w]; hs[w] = ts;
    int64_t ti = hi[i]; hi[i] = hi[w]; hi[w] = ti;
    i = w;
  }
}

static inline void heap_init(int k, float* hs, int64_t* hi) {
  for (int i = 0; i < k; i++) { hs[i] = -FLT_MAX; hi[i] = INT64_MAX; }
}

static inline void heap_offer(int k, float* hs, int64_t* hi, float s, int64_t id) {
  /* empty slots are (-FLT_MAX, INT64_MAX): any real candidate with s > -FLT_MAX beats them */
  if (worse(hs[0], hi[0], s, id)) {
    hs[0] = s; hi[0] = id;
    heap_sift_down(k, hs, hi, 0);
  }
}

/* emit best first; unfilled -> (-FLT_MAX, -1) as faiss does for inner product */
static void heap_emit(int k, float* hs, int64_t* hi, float* D, int64_t* I) {
  for (int n = k; n > 0; n--) {
    D[n - 1] = hs[0];
    I[n - 1] = (hi[0] == INT64_MAX) ? -1 : hi[0];
    hs[0] = hs[n - 1]; hi[0] = hi[n - 1];
    heap_sift_down(n - 1, hs, hi, 0);
  }
}

/* ------------------------------------------------------------------ distances -------------- */
static inline float fvec_ip(const float* a, const float* b, int d) {
  float r = 0.f;
#pragma omp simd reduction(+ : r)
  for (int i = 0; i < d; i++) r += a[i] * b[i];
  return r;
}

/* IndexFlatIP::search */
void orc_flat_search(int64_t nq, const float* q, int64_t nb, const float* xb, int d, int k,
                     float* D, int64_t* I) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t i = 0; i < nq; i++) {
    float* hs = (float*)malloc(sizeof(float) * k);
    int64_t* hi = (int64_t*)malloc(sizeof(int64_t) * k);
    heap_init(k, hs, hi);
    for (int64_t j = 0; j < nb; j++) heap_offer(k, hs, hi, fvec_ip(q + i * d, xb + j * d, d), j);
    heap_emit(k, hs, hi, D + i * k, I + i * k);
    free(hs); free(hi);
  }
}

/* IndexIVFFlat::search_preassigned: lists as CSR (offsets[nlist+1] into codes/ids) */
void orc_ivf_scan(int64_t nq, const float* q, int d, int k, int nprobe, const int64_t* coarse,
                  const int64_t* offsets, const float* codes, const int64_t* ids, float* D,
                  int64_t* I, int64_t* nscanned) {
  int64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
  for (int64_t i = 0; i < nq; i++) {
    float* hs = (float*)malloc(sizeof(float) * k);
    int64_t* hi = (int64_t*)malloc(sizeof(int64_t) * k);
    heap_init(k, hs, hi);
    for (int p = 0; p < nprobe; p++) {
      int64_t l = coarse[i * nprobe + p];
      if (l < 0) continue;
      for (int64_t j = offsets[l]; j < offsets[l + 1]; j++)
        heap_offer(k, hs, hi, fvec_ip(q + i * d, codes + j * d, d), ids[j]);
      total += offsets[l + 1] - offsets[l];
    }
    heap_emit(k, hs, hi, D + i * k, I + i * k);
    free(hs); free(hi);
  }
  if (nscanned) *nscanned = total;
}

/* quantizer.search(x, nprobe) as a plain loop (the blocked-sgemm variant lives in ivf.py) */
void orc_coarse(int64_t nq, const float* q, int64_t nlist, const float* cent, int d, int nprobe,
                float* Dc, int64_t* Ic) {
  orc_flat_search(nq, q, nlist, cent, d, nprobe, Dc, Ic);
}

/* ------------------------------------------------------------------ k-means pieces --------- */
/* Clustering::compute_centroids: fp32 running sums in row order, then c *= 1/count */
void orc_compute_centroids(int64_t n, const float* x, int d, int64_t k, const int64_t* assign,
                           float* centroids, float* hassign) {
  memset(centroids, 0, sizeof(float) * k * d);
  memset(hassign, 0, sizeof(float) * k);
  for (int64_t i = 0; i < n; i++) {
    int64_t c = assign[i];
    float* cc = centroids + c * d;
    const float* xi = x + i * d;
    hassign[c] += 1.f;
    for (int j = 0; j < d; j++) cc[j] += xi[j];
  }
  for (int64_t c = 0; c < k; c++) {
    if (hassign[c] == 0.f) continue;
    float norm = 1.f / hassign[c];
    for (int j = 0; j < d; j++) centroids[c * d + j] *= norm;
  }
}

/* Clustering::split_clusters: every empty cluster steals half of a cluster drawn with
 * probability (size-1)/(n-k), both copies perturbed by (1 +- 1/1024) on alternating dims. */
int64_t orc_split_clusters(int d, int64_t k, int64_t n, float* hassign, float* centroids) {
  const float EPS = 1.f / 1024.f;
  int64_t nsplit = 0;
  orc_mt_t s;
  mt_seed(&s, 1234u);
  for (int64_t ci = 0; ci < k; ci++) {
    if (hassign[ci] != 0.f) continue;
    int64_t cj;
    for (cj = 0;; cj = (cj + 1) % k) {
      float p = (float)((hassign[cj] - 1.0) / (float)(n - k));
      float r = mt_next(&s) / (float)4294967295u;
      if (r < p) break;
    }
    memcpy(centroids + ci * d, centroids + cj * d, sizeof(float) * d);
    for (int j = 0; j < d; j++) {
      if (j % 2 == 0) {
        centroids[ci * d + j] *= 1 + EPS;
        centroids[cj * d + j] *= 1 - EPS;
      } else {
        centroids[ci * d + j] *= 1 - EPS;
        centroids[cj * d + j] *= 1 + EPS;
      }
    }
    hassign[ci] = hassign[cj] / 2;
    hassign[cj] -= hassign[ci];
    nsplit++;
  }
  return nsplit;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
