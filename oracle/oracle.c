/*
 * oracle.c — CPU restatement (plain C99 + OpenMP) of the libpointmatcher /
 * libnabo ICP chain that pgslam drives.  See oracle.h for the scope, the
 * "parity unpinned" statement and the numeric contract.
 *
 * TEST INFRASTRUCTURE ONLY — never linked into the product library.
 *
 * Build: gcc -O3 -fopenmp -ffp-contract=off -fno-fast-math (see Makefile);
 * -ffp-contract=off is load-bearing: fp32 distances / transforms must not be
 * fused.
 *
 * Citations: "A.n" = SURVEY.md Appendix A section n (the [UPSTREAM-RECALLED]
 * behavioural spec); file:line = /root/reference/src/pgslam/.
 */
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BUCKET_SIZE 8

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ======================================================================== */
/* DataPoints                                                               */
/* ======================================================================== */
orc_cloud *orc_cloud_new(int64_t n) {
  orc_cloud *c = (orc_cloud *)calloc(1, sizeof(orc_cloud));
  c->n = n;
  c->feat = (float *)calloc((size_t)(n > 0 ? n : 1) * 4, sizeof(float));
  for (int64_t i = 0; i < n; ++i) c->feat[4 * i + 3] = 1.0f;
  return c;
}

static float *dup_f(const float *p, size_t cnt) {
  if (!p) return NULL;
  float *q = (float *)malloc((cnt ? cnt : 1) * sizeof(float));
  memcpy(q, p, cnt * sizeof(float));
  return q;
}

orc_cloud *orc_cloud_copy(const orc_cloud *c) {
  orc_cloud *o = (orc_cloud *)calloc(1, sizeof(orc_cloud));
  size_t n = (size_t)c->n;
  o->n = c->n;
  o->feat = dup_f(c->feat, 4 * n);
  o->normals = dup_f(c->normals, 3 * n);
  o->obsdir = dup_f(c->obsdir, 3 * n);
  o->noise = dup_f(c->noise, n);
  o->dens = dup_f(c->dens, n);
  o->eigval = dup_f(c->eigval, 3 * n);
  o->eigvec = dup_f(c->eigvec, 9 * n);
  o->meandist = dup_f(c->meandist, n);
  o->matched = dup_f(c->matched, (size_t)c->matched_span * n);
  o->matched_span = c->matched_span;
  return o;
}

void orc_cloud_free(orc_cloud *c) {
  if (!c) return;
  free(c->feat); free(c->normals); free(c->obsdir); free(c->noise);
  free(c->dens); free(c->eigval); free(c->eigvec); free(c->meandist); free(c->matched);
  free(c);
}

static void cat_desc(float **a, int64_t na, const float *b, int64_t nb, int span) {
  /* concatenate keeps a descriptor only if both clouds carry it (A.9) */
  if (*a && b) {
    *a = (float *)realloc(*a, (size_t)(na + nb) * span * sizeof(float) + 4);
    memcpy(*a + (size_t)na * span, b, (size_t)nb * span * sizeof(float));
  } else {
    free(*a);
    *a = NULL;
  }
}

void orc_cloud_concatenate(orc_cloud *a, const orc_cloud *b) {
  int64_t na = a->n, nb = b->n;
  a->feat = (float *)realloc(a->feat, (size_t)(na + nb) * 4 * sizeof(float) + 4);
  memcpy(a->feat + 4 * na, b->feat, (size_t)nb * 4 * sizeof(float));
  cat_desc(&a->normals, na, b->normals, nb, 3);
  cat_desc(&a->obsdir, na, b->obsdir, nb, 3);
  cat_desc(&a->noise, na, b->noise, nb, 1);
  cat_desc(&a->dens, na, b->dens, nb, 1);
  cat_desc(&a->eigval, na, b->eigval, nb, 3);
  cat_desc(&a->eigvec, na, b->eigvec, nb, 9);
  cat_desc(&a->meandist, na, b->meandist, nb, 1);
  /* same name but different span: not the same descriptor -> dropped */
  cat_desc(&a->matched, na, a->matched_span == b->matched_span ? b->matched : NULL, nb, a->matched_span);
  if (!a->matched) a->matched_span = 0;
  a->n = na + nb;
}

/* keep columns listed in idx[0..m) (ascending), compacting every field */
static void cloud_select(orc_cloud *c, const int64_t *idx, int64_t m) {
#define SEL(field, span)                                                   \
  if (c->field) {                                                          \
    for (int64_t j = 0; j < m; ++j)                                        \
      if (idx[j] != j)                                                     \
        memmove(c->field + (size_t)j * (span), c->field + (size_t)idx[j] * (span), \
                (span) * sizeof(float));                                   \
  }
  SEL(feat, 4) SEL(normals, 3) SEL(obsdir, 3) SEL(noise, 1) SEL(dens, 1)
  SEL(eigval, 3) SEL(eigvec, 9) SEL(meandist, 1) SEL(matched, c->matched_span)
#undef SEL
  c->n = m;
}

/* ======================================================================== */
/* fp32 primitives of the numeric contract                                  */
/* ======================================================================== */
static inline float dist2_f32(const float *a, const float *b) {
  /* A.2: dist = 0; for d: diff = q[d]-p[d]; dist += diff*diff   (type T)  */
  float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  float d = dx * dx;
  d = d + dy * dy;
  d = d + dz * dz;
  return d;
}

static inline void xform_point_f32(const float *Tf, const float *p, float *o) {
  /* Eigen 4x4 * 4xN product, sequential-k accumulation, w == 1 */
  float x = p[0], y = p[1], z = p[2];
  for (int r = 0; r < 3; ++r) {
    float a = Tf[r] * x;
    a = a + Tf[4 + r] * y;
    a = a + Tf[8 + r] * z;
    a = a + Tf[12 + r];
    o[r] = a;
  }
  o[3] = 1.0f;
}

static inline void rot_vec_f32(const float *Tf, const float *v, float *o) {
  float x = v[0], y = v[1], z = v[2];
  for (int r = 0; r < 3; ++r) {
    float a = Tf[r] * x;
    a = a + Tf[4 + r] * y;
    a = a + Tf[8 + r] * z;
    o[r] = a;
  }
}

/* ======================================================================== */
/* 4x4 double helpers (col-major)                                           */
/* ======================================================================== */
static void m4_identity(double *M) {
  memset(M, 0, 16 * sizeof(double));
  M[0] = M[5] = M[10] = M[15] = 1.0;
}
static void m4_mul(const double *A, const double *B, double *C) {
  double R[16];
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += A[k * 4 + r] * B[c * 4 + k];
      R[c * 4 + r] = s;
    }
  memcpy(C, R, sizeof(R));
}
/* inverse of a rigid transform: [R t]^-1 = [R^T  -R^T t] */
static void m4_rigid_inv(const double *M, double *O) {
  double R[16];
  m4_identity(R);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[c * 4 + r] = M[r * 4 + c];
  for (int r = 0; r < 3; ++r) {
    double s = 0.0;
    for (int k = 0; k < 3; ++k) s += R[k * 4 + r] * M[12 + k];
    R[12 + r] = -s;
  }
  memcpy(O, R, sizeof(R));
}
static double m3_det_of4(const double *M) {
  return M[0] * (M[5] * M[10] - M[9] * M[6]) - M[4] * (M[1] * M[10] - M[9] * M[2]) +
         M[8] * (M[1] * M[6] - M[5] * M[2]);
}

/* ======================================================================== */
/* libnabo kd-tree restatement (A.1, A.2)                                   */
/* ======================================================================== */
typedef struct {
  /* inner: dim in {0,1,2}, cut value, right child index (left = self+1)
   * leaf : dim == 3, bucket start, bucket size                             */
  uint32_t dim;
  uint32_t right_or_start;
  float cut;
  uint32_t size;
} kd_node;

struct orc_kdtree {
  const float *pts; /* borrowed: 4*n (libnabo keeps a reference, A8)        */
  int64_t n;
  kd_node *nodes;
  int64_t n_nodes, cap_nodes;
  int32_t *bucket; /* point indices in leaf order                           */
  int64_t n_bucket;
};

static int64_t kd_push_node(orc_kdtree *t) {
  if (t->n_nodes == t->cap_nodes) {
    t->cap_nodes = t->cap_nodes ? t->cap_nodes * 2 : 1024;
    t->nodes = (kd_node *)realloc(t->nodes, (size_t)t->cap_nodes * sizeof(kd_node));
  }
  return t->n_nodes++;
}

static int64_t kd_build(orc_kdtree *t, int32_t *idx, int64_t first, int64_t last,
                        float minV[3], float maxV[3]) {
  int64_t count = last - first;
  int64_t pos = kd_push_node(t);
  if (count <= BUCKET_SIZE) {
    t->nodes[pos].dim = 3;
    t->nodes[pos].right_or_start = (uint32_t)t->n_bucket;
    t->nodes[pos].size = (uint32_t)count;
    t->nodes[pos].cut = 0.f;
    for (int64_t i = first; i < last; ++i) t->bucket[t->n_bucket++] = idx[i];
    return pos;
  }
  /* cut dimension: first max extent of the INHERITED box */
  int cd = 0;
  float ext = maxV[0] - minV[0];
  for (int d = 1; d < 3; ++d)
    if (maxV[d] - minV[d] > ext) { ext = maxV[d] - minV[d]; cd = d; }
  float ideal = (maxV[cd] + minV[cd]) / 2;
  float lo = t->pts[4 * (int64_t)idx[first] + cd], hi = lo;
  for (int64_t i = first + 1; i < last; ++i) {
    float v = t->pts[4 * (int64_t)idx[i] + cd];
    if (v < lo) lo = v;
    if (v > hi) hi = v;
  }
  float cut = ideal < lo ? lo : (ideal > hi ? hi : ideal);
  /* three-way partition: [<cut | ==cut | >cut] */
  int64_t l = first, r = last - 1;
  while (1) {
    while (l <= r && t->pts[4 * (int64_t)idx[l] + cd] < cut) ++l;
    while (l <= r && t->pts[4 * (int64_t)idx[r] + cd] >= cut) --r;
    if (l >= r) break;
    int32_t tmp = idx[l]; idx[l] = idx[r]; idx[r] = tmp;
    ++l; --r;
  }
  int64_t br1 = l - first;
  r = last - 1;
  while (1) {
    while (l <= r && t->pts[4 * (int64_t)idx[l] + cd] <= cut) ++l;
    while (l <= r && t->pts[4 * (int64_t)idx[r] + cd] > cut) --r;
    if (l >= r) break;
    int32_t tmp = idx[l]; idx[l] = idx[r]; idx[r] = tmp;
    ++l; --r;
  }
  int64_t br2 = l - first;
  int64_t left;
  if (ideal < lo) left = 1;
  else if (ideal > hi) left = count - 1;
  else if (br1 > count / 2) left = br1;
  else if (br2 < count / 2) left = br2;
  else left = count / 2;
  t->nodes[pos].dim = (uint32_t)cd;
  t->nodes[pos].cut = cut;
  t->nodes[pos].size = 0;
  float save = maxV[cd];
  maxV[cd] = cut;
  kd_build(t, idx, first, first + left, minV, maxV);
  maxV[cd] = save;
  save = minV[cd];
  minV[cd] = cut;
  int64_t right = kd_build(t, idx, first + left, last, minV, maxV);
  minV[cd] = save;
  t->nodes[pos].right_or_start = (uint32_t)right;
  return pos;
}

orc_kdtree *orc_kdtree_build(const float *feat4, int64_t n) {
  orc_kdtree *t = (orc_kdtree *)calloc(1, sizeof(orc_kdtree));
  t->pts = feat4;
  t->n = n;
  t->bucket = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
  if (n <= 0) return t;
  int32_t *idx = (int32_t *)malloc((size_t)n * sizeof(int32_t));
  float minV[3], maxV[3];
  for (int d = 0; d < 3; ++d) minV[d] = maxV[d] = feat4[d];
  for (int64_t i = 0; i < n; ++i) {
    idx[i] = (int32_t)i;
    for (int d = 0; d < 3; ++d) {
      float v = feat4[4 * i + d];
      if (v < minV[d]) minV[d] = v;
      if (v > maxV[d]) maxV[d] = v;
    }
  }
  kd_build(t, idx, 0, n, minV, maxV);
  free(idx);
  return t;
}

void orc_kdtree_free(orc_kdtree *t) {
  if (!t) return;
  free(t->nodes);
  free(t->bucket);
  free(t);
}

/* sorted-vector heap (A.2 IndexHeapBruteForceVector) with the repo's tie
 * rule: entries ordered by (dist, index) lexicographically.     */
typedef struct {
  int k;
  float *d;
  int32_t *id;
} kd_heap;

static inline int lex_less(float d, int32_t id, float hd, int32_t hid) {
  /* index -1 (empty slot) has dist +inf; a real point at +inf never enters */
  return d < hd || (d == hd && id < hid);
}

static inline void heap_insert(kd_heap *h, int32_t id, float d) {
  int j = h->k - 1;
  while (j > 0 && lex_less(d, id, h->d[j - 1], h->id[j - 1])) {
    h->d[j] = h->d[j - 1];
    h->id[j] = h->id[j - 1];
    --j;
  }
  h->d[j] = d;
  h->id[j] = id;
}

typedef struct {
  const orc_kdtree *t;
  const float *q;
  float off[3];
  float maxr2;
  int allow_self;
  kd_heap h;
  uint64_t visits;
} kd_search;

static void kd_recurse(kd_search *s, int64_t n) {
  const kd_node *nd = &s->t->nodes[n];
  if (nd->dim == 3) {
    const int32_t *b = s->t->bucket + nd->right_or_start;
    for (uint32_t i = 0; i < nd->size; ++i) {
      int32_t pi = b[i];
      float d = dist2_f32(s->q, s->t->pts + 4 * (int64_t)pi);
      int last = s->h.k - 1;
      if (d <= s->maxr2 && lex_less(d, pi, s->h.d[last], s->h.id[last]) &&
          (s->allow_self || d > FLT_EPSILON)) {
        /* an empty slot has id -1: lex_less with hd=+inf handles it        */
        heap_insert(&s->h, pi, d);
      }
    }
    s->visits++;
    return;
  }
  int cd = (int)nd->dim;
  float old = s->off[cd];
  float nw = s->q[cd] - nd->cut;
  int64_t nearc, farc;
  if (nw > 0) { nearc = nd->right_or_start; farc = n + 1; }
  else        { nearc = n + 1; farc = nd->right_or_start; }
  kd_recurse(s, nearc);
  /* Exact, conservative lower bound: recomputed from the offset vector with
   * the same fp32 operation order as dist2_f32 (monotone => never prunes a
   * true (dist,index) minimum).  libnabo's incremental rd update (A.2) can
   * drift by an ulp; this is the documented fix.               */
  s->off[cd] = nw;
  float rd = s->off[0] * s->off[0];
  rd = rd + s->off[1] * s->off[1];
  rd = rd + s->off[2] * s->off[2];
  if (rd <= s->maxr2 && rd <= s->h.d[s->h.k - 1]) kd_recurse(s, farc);
  s->off[cd] = old;
}

/* ---- libnabo-faithful search (A.2 verbatim) -------------------------------
 * What the contract mode above changes on purpose, undone here so that the
 * two can be compared (tests/test_cpu.py, oracle/README.md):
 *   - candidates enter on a STRICT distance compare (dist < headValue), and an
 *     equal value stays behind the earlier-visited one (no index tie-break);
 *   - the far side's lower bound is the INCREMENTAL rd += -old*old + new*new;
 *   - pruning uses rd * (1+eps)^2 < headValue, i.e. epsilon is honoured.      */
static inline void heap_insert_nabo(kd_heap *h, int32_t id, float d) {
  int j = h->k - 1;
  while (j > 0 && h->d[j - 1] > d) {
    h->d[j] = h->d[j - 1];
    h->id[j] = h->id[j - 1];
    --j;
  }
  h->d[j] = d;
  h->id[j] = id;
}

static void kd_recurse_nabo(kd_search *s, int64_t n, float rd, float max_error2) {
  const kd_node *nd = &s->t->nodes[n];
  if (nd->dim == 3) {
    const int32_t *b = s->t->bucket + nd->right_or_start;
    for (uint32_t i = 0; i < nd->size; ++i) {
      int32_t pi = b[i];
      float d = dist2_f32(s->q, s->t->pts + 4 * (int64_t)pi);
      if (d <= s->maxr2 && d < s->h.d[s->h.k - 1] && (s->allow_self || d > FLT_EPSILON))
        heap_insert_nabo(&s->h, pi, d);
    }
    s->visits++;
    return;
  }
  int cd = (int)nd->dim;
  float old = s->off[cd];
  float nw = s->q[cd] - nd->cut;
  int64_t nearc, farc;
  if (nw > 0) { nearc = nd->right_or_start; farc = n + 1; }
  else        { nearc = n + 1; farc = nd->right_or_start; }
  kd_recurse_nabo(s, nearc, rd, max_error2);
  rd += -old * old + nw * nw;
  if (rd <= s->maxr2 && rd * max_error2 < s->h.d[s->h.k - 1]) {
    s->off[cd] = nw;
    kd_recurse_nabo(s, farc, rd, max_error2);
    s->off[cd] = old;
  }
}

static int g_search_mode = ORC_SEARCH_CONTRACT;
void orc_set_search_mode(int mode) { g_search_mode = mode; }
int orc_search_mode(void) { return g_search_mode; }

uint64_t orc_kdtree_knn_ex(const orc_kdtree *t, const float *query4, int64_t nq, int k, float max_dist,
                           int allow_self, float epsilon, int mode, int32_t *ids, float *d2) {
  uint64_t total = 0;
  float maxr2 = isinf(max_dist) ? INFINITY : max_dist * max_dist;
  const float max_error2 = (1.f + epsilon) * (1.f + epsilon);
#pragma omp parallel reduction(+ : total)
  {
    kd_search s;
    s.t = t;
    s.maxr2 = maxr2;
    s.allow_self = allow_self;
    s.h.k = k;
#pragma omp for schedule(guided, 32)
    for (int64_t i = 0; i < nq; ++i) {
      s.q = query4 + 4 * i;
      s.off[0] = s.off[1] = s.off[2] = 0.f;
      s.h.d = d2 + (size_t)i * k;
      s.h.id = ids + (size_t)i * k;
      for (int j = 0; j < k; ++j) { s.h.d[j] = INFINITY; s.h.id[j] = -1; }
      s.visits = 0;
      if (t->n > 0) {
        if (mode == ORC_SEARCH_NABO) kd_recurse_nabo(&s, 0, 0.f, max_error2);
        else kd_recurse(&s, 0);
      }
      total += s.visits;
    }
  }
  return total;
}

/* the contract: exact (eps = 0), ties -> lower index */
uint64_t orc_kdtree_knn(const orc_kdtree *t, const float *query4, int64_t nq, int k,
                        float max_dist, int allow_self, int32_t *ids, float *d2) {
  return orc_kdtree_knn_ex(t, query4, nq, k, max_dist, allow_self, 0.f, ORC_SEARCH_CONTRACT, ids, d2);
}

void orc_knn_brute(const float *ref4, int64_t n, const float *query4, int64_t nq,
                   int k, float max_dist, int32_t *ids, float *d2) {
  float maxr2 = isinf(max_dist) ? INFINITY : max_dist * max_dist;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nq; ++i) {
    kd_heap h;
    h.k = k;
    h.d = d2 + (size_t)i * k;
    h.id = ids + (size_t)i * k;
    for (int j = 0; j < k; ++j) { h.d[j] = INFINITY; h.id[j] = -1; }
    const float *q = query4 + 4 * i;
    for (int64_t p = 0; p < n; ++p) {
      float d = dist2_f32(q, ref4 + 4 * p);
      if (d <= maxr2 && lex_less(d, (int32_t)p, h.d[k - 1], h.id[k - 1]))
        heap_insert(&h, (int32_t)p, d);
    }
  }
}

/* ======================================================================== */
/* small dense algebra (fp64; +,-,*,/,sqrt only => reproducible bit-for-bit */
/* on any IEEE machine without contraction)                                 */
/* ======================================================================== */

/* cyclic Jacobi on a symmetric n x n (n <= 6) matrix, col-major.           */
static void jacobi_sym(int n, double *A, double *w, double *V) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[j * n + i] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) off += fabs(A[q * n + p]);
    if (off == 0.0) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        double apq = A[q * n + p];
        if (apq == 0.0) continue;
        double app = A[p * n + p], aqq = A[q * n + q];
        double g = 100.0 * fabs(apq);
        if (sweep > 3 && fabs(app) + g == fabs(app) && fabs(aqq) + g == fabs(aqq)) {
          A[q * n + p] = 0.0;
          A[p * n + q] = 0.0;
          continue;
        }
        double h = aqq - app, t;
        if (fabs(h) + g == fabs(h)) {
          t = apq / h;
        } else {
          double theta = 0.5 * h / apq;
          t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
          if (theta < 0.0) t = -t;
        }
        double c = 1.0 / sqrt(t * t + 1.0);
        double s = t * c;
        double tau = s / (1.0 + c);
        double hh = t * apq;
        A[p * n + p] = app - hh;
        A[q * n + q] = aqq + hh;
        A[q * n + p] = 0.0;
        A[p * n + q] = 0.0;
        for (int r = 0; r < n; ++r) {
          if (r != p && r != q) {
            double arp = A[p * n + r], arq = A[q * n + r];
            double nrp = arp - s * (arq + arp * tau);
            double nrq = arq + s * (arp - arq * tau);
            A[p * n + r] = nrp; A[r * n + p] = nrp;
            A[q * n + r] = nrq; A[r * n + q] = nrq;
          }
        }
        for (int r = 0; r < n; ++r) {
          double vrp = V[p * n + r], vrq = V[q * n + r];
          V[p * n + r] = vrp - s * (vrq + vrp * tau);
          V[q * n + r] = vrq + s * (vrp - vrq * tau);
        }
      }
  }
  for (int i = 0; i < n; ++i) w[i] = A[i * n + i];
}

void orc_eig3_sym(const double *A, double *w, double *V) {
  double B[9];
  memcpy(B, A, sizeof(B));
  jacobi_sym(3, B, w, V);
}

/* A x = b, A symmetric PSD 6x6.  Cholesky when well conditioned, else the
 * minimum-norm solution through the eigen-decomposition (A.5's
 * solvePossiblyUnderdeterminedLinearSystem restated).  Returns rank.        */
static int solve_n(int n, const double *A, const double *b, double *x) {
  double L[36];
  double maxd = 0.0;
  for (int i = 0; i < n; ++i)
    if (A[i * n + i] > maxd) maxd = A[i * n + i];
  int ok = maxd > 0.0;
  memset(L, 0, sizeof(L));
  for (int j = 0; j < n && ok; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= L[k * n + j] * L[k * n + j];
    if (!(d > 1e-12 * maxd)) { ok = 0; break; }
    double ljj = sqrt(d);
    L[j * n + j] = ljj;
    for (int i = j + 1; i < n; ++i) {
      double s = A[j * n + i];
      for (int k = 0; k < j; ++k) s -= L[k * n + i] * L[k * n + j];
      L[j * n + i] = s / ljj;
    }
  }
  if (ok) {
    double y[6];
    for (int i = 0; i < n; ++i) {
      double s = b[i];
      for (int k = 0; k < i; ++k) s -= L[k * n + i] * y[k];
      y[i] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = y[i];
      for (int k = i + 1; k < n; ++k) s -= L[i * n + k] * x[k];
      x[i] = s / L[i * n + i];
    }
    return n;
  }
  double B[36], w[6], V[36];
  memcpy(B, A, sizeof(double) * n * n);
  jacobi_sym(n, B, w, V);
  double wmax = 0.0;
  for (int i = 0; i < n; ++i)
    if (fabs(w[i]) > wmax) wmax = fabs(w[i]);
  for (int i = 0; i < n; ++i) x[i] = 0.0;
  int rank = 0;
  for (int e = 0; e < n; ++e) {
    if (!(w[e] > 1e-12 * wmax)) continue;
    ++rank;
    double vb = 0.0;
    for (int i = 0; i < n; ++i) vb += V[e * n + i] * b[i];
    vb = vb / w[e];
    for (int i = 0; i < n; ++i) x[i] += vb * V[e * n + i];
  }
  return rank;
}

int orc_solve6(const double *A, const double *b, double *x) { return solve_n(6, A, b, x); }

/* 3x3 SVD M = U diag(S) V^T through the eigen-decomposition of M^T M with a
 * Gram–Schmidt completion of U; singular values sorted descending.          */
void orc_svd3(const double *M, double *U, double *S, double *V) {
  double MtM[9], w[3], Vv[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += M[i * 3 + k] * M[j * 3 + k];
      MtM[j * 3 + i] = s;
    }
  jacobi_sym(3, MtM, w, Vv);
  int ord[3] = {0, 1, 2};
  for (int a = 0; a < 2; ++a)
    for (int b2 = a + 1; b2 < 3; ++b2)
      if (w[ord[b2]] > w[ord[a]]) { int t = ord[a]; ord[a] = ord[b2]; ord[b2] = t; }
  for (int c = 0; c < 3; ++c) {
    for (int r = 0; r < 3; ++r) V[c * 3 + r] = Vv[ord[c] * 3 + r];
    S[c] = w[ord[c]] > 0.0 ? sqrt(w[ord[c]]) : 0.0;
  }
  /* make V a proper rotation-or-reflection consistently: keep as is */
  double smax = S[0];
  int good = 0;
  for (int c = 0; c < 3; ++c) {
    double u[3];
    for (int r = 0; r < 3; ++r) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += M[k * 3 + r] * V[c * 3 + k];
      u[r] = s;
    }
    /* orthogonalise against previous columns */
    for (int p = 0; p < good; ++p) {
      double dp = u[0] * U[p * 3] + u[1] * U[p * 3 + 1] + u[2] * U[p * 3 + 2];
      for (int r = 0; r < 3; ++r) u[r] -= dp * U[p * 3 + r];
    }
    double nrm = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    if (S[c] > 1e-13 * smax && nrm > 0.0) {
      for (int r = 0; r < 3; ++r) U[c * 3 + r] = u[r] / nrm;
    } else {
      /* null direction: complete an orthonormal basis */
      double best[3] = {0, 0, 0};
      double bestn = -1.0;
      for (int e = 0; e < 3; ++e) {
        double v[3] = {0, 0, 0};
        v[e] = 1.0;
        for (int p = 0; p < good; ++p) {
          double dp = v[0] * U[p * 3] + v[1] * U[p * 3 + 1] + v[2] * U[p * 3 + 2];
          for (int r = 0; r < 3; ++r) v[r] -= dp * U[p * 3 + r];
        }
        double vn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        if (vn > bestn) { bestn = vn; for (int r = 0; r < 3; ++r) best[r] = v[r] / vn; }
      }
      for (int r = 0; r < 3; ++r) U[c * 3 + r] = best[r];
    }
    ++good;
  }
}

/* ======================================================================== */
/* filters (A.9)                                                            */
/* ======================================================================== */
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
/* counter-based replacement for the sequential rand() of A.9    */
static inline float hash_uniform(uint64_t seed, uint64_t i) {
  uint64_t h = splitmix64(splitmix64(seed) ^ (i * 0xD1B54A32D192ED03ull));
  return (float)(h >> 40) * (1.0f / 16777216.0f);
}

static int filter_random_sampling(const orc_filter *f, orc_cloud *c) {
  float prob = (float)f->p0;
  int64_t *keep = (int64_t *)malloc((size_t)(c->n + 1) * sizeof(int64_t));
  int64_t m = 0;
  for (int64_t i = 0; i < c->n; ++i)
    if (hash_uniform((uint64_t)f->i0, (uint64_t)i) < prob) keep[m++] = i;
  cloud_select(c, keep, m);
  free(keep);
  return ORC_OK;
}

typedef struct { uint64_t vox; int64_t p; } vox_pair;
static int cmp_vox(const void *a, const void *b) {
  const vox_pair *x = (const vox_pair *)a, *y = (const vox_pair *)b;
  if (x->vox != y->vox) return x->vox < y->vox ? -1 : 1;
  return x->p < y->p ? -1 : (x->p > y->p ? 1 : 0);
}
static int cmp_i64(const void *a, const void *b) {
  int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

static int filter_voxel_grid(const orc_filter *f, orc_cloud *c) {
  int64_t n = c->n;
  if (n == 0) return ORC_OK;
  float vs[3] = {(float)f->p0, (float)f->p1, (float)f->p2};
  int use_centroid = (int)f->i0, avg_desc = (int)f->i1;
  float minV[3], maxV[3], minB[3], maxB[3];
  uint64_t nd[3];
  for (int d = 0; d < 3; ++d) minV[d] = maxV[d] = c->feat[d];
  for (int64_t i = 0; i < n; ++i)
    for (int d = 0; d < 3; ++d) {
      float v = c->feat[4 * i + d];
      if (v < minV[d]) minV[d] = v;
      if (v > maxV[d]) maxV[d] = v;
    }
  for (int d = 0; d < 3; ++d) {
    minB[d] = minV[d] / vs[d];
    maxB[d] = maxV[d] / vs[d];
    float nf = 1.0f + maxB[d];
    nf = nf - minB[d];
    nd[d] = (uint64_t)nf;
  }
  vox_pair *vp = (vox_pair *)malloc((size_t)n * sizeof(vox_pair));
  for (int64_t p = 0; p < n; ++p) {
    uint64_t ijk[3];
    for (int d = 0; d < 3; ++d) {
      float q = c->feat[4 * p + d] / vs[d];
      q = q - minB[d];
      ijk[d] = (uint64_t)floorf(q);
    }
    vp[p].vox = ijk[0] + ijk[1] * nd[0] + ijk[2] * nd[0] * nd[1];
    vp[p].p = p;
  }
  /* (voxel, input index) order == the sequential three-pass procedure of A4 */
  qsort(vp, (size_t)n, sizeof(vox_pair), cmp_vox);
  int64_t *firsts = (int64_t *)malloc((size_t)n * sizeof(int64_t));
  int64_t m = 0;
  for (int64_t s = 0; s < n;) {
    int64_t e = s;
    while (e < n && vp[e].vox == vp[s].vox) ++e;
    int64_t first = vp[s].p;
    float cnt = (float)(e - s);
#define AVG(field, span)                                                   \
    if (c->field) {                                                        \
      for (int64_t j = s + 1; j < e; ++j)                                  \
        for (int d = 0; d < (span); ++d)                                   \
          c->field[(size_t)first * (span) + d] =                           \
              c->field[(size_t)first * (span) + d] + c->field[(size_t)vp[j].p * (span) + d]; \
      for (int d = 0; d < (span); ++d)                                     \
        c->field[(size_t)first * (span) + d] = c->field[(size_t)first * (span) + d] / cnt; \
    }
    if (use_centroid) {
      for (int64_t j = s + 1; j < e; ++j)
        for (int d = 0; d < 3; ++d)
          c->feat[4 * first + d] = c->feat[4 * first + d] + c->feat[4 * vp[j].p + d];
      for (int d = 0; d < 3; ++d) c->feat[4 * first + d] = c->feat[4 * first + d] / cnt;
    } else {
      uint64_t v = vp[s].vox;
      uint64_t ijk[3] = {v % nd[0], (v / nd[0]) % nd[1], v / (nd[0] * nd[1])};
      for (int d = 0; d < 3; ++d) {
        float ctr = (float)ijk[d] + 0.5f;
        ctr = ctr + minB[d];
        c->feat[4 * first + d] = vs[d] * ctr;
      }
    }
    if (avg_desc) {
      AVG(normals, 3) AVG(obsdir, 3) AVG(noise, 1) AVG(dens, 1) AVG(eigval, 3) AVG(eigvec, 9)
      AVG(meandist, 1) AVG(matched, c->matched_span)
    }
#undef AVG
    firsts[m++] = first;
    s = e;
  }
  qsort(firsts, (size_t)m, sizeof(int64_t), cmp_i64);
  cloud_select(c, firsts, m);
  free(firsts);
  free(vp);
  return ORC_OK;
}

static int filter_surface_normal(const orc_filter *f, orc_cloud *c) {
  int64_t n = c->n;
  int k = (int)f->i0;
  int flags = (int)f->i1;
  float max_dist = (float)f->p0;
  if (k < 3) return ORC_INVALID_PARAMETER;
  int32_t *ids = (int32_t *)malloc((size_t)(n + 1) * k * sizeof(int32_t));
  float *d2 = (float *)malloc((size_t)(n + 1) * k * sizeof(float));
  orc_kdtree *t = orc_kdtree_build(c->feat, n);
  orc_kdtree_knn(t, c->feat, n, k, max_dist, 1, ids, d2);
  orc_kdtree_free(t);
  if (flags & 1) { free(c->normals); c->normals = (float *)calloc((size_t)(n + 1) * 3, sizeof(float)); }
  if (flags & 2) { free(c->dens); c->dens = (float *)calloc((size_t)(n + 1), sizeof(float)); }
  if (flags & 4) { free(c->eigval); c->eigval = (float *)calloc((size_t)(n + 1) * 3, sizeof(float)); }
  if (flags & 8) { free(c->eigvec); c->eigvec = (float *)calloc((size_t)(n + 1) * 9, sizeof(float)); }
  if (flags & 16) {
    /* keepMatchedIds: the k x N ids cast to T (unfound -> -1)                */
    free(c->matched);
    c->matched = (float *)calloc((size_t)(n + 1) * k, sizeof(float));
    c->matched_span = k;
    for (int64_t m = 0; m < n * k; ++m) c->matched[m] = (float)ids[m];
  }
  if (flags & 32) { free(c->meandist); c->meandist = (float *)calloc((size_t)(n + 1), sizeof(float)); }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double mean[3] = {0, 0, 0};
    int real = 0;
    for (int j = 0; j < k; ++j) {
      int32_t id = ids[(size_t)i * k + j];
      if (id < 0) continue;
      for (int d = 0; d < 3; ++d) mean[d] += (double)c->feat[4 * (int64_t)id + d];
      ++real;
    }
    if (real == 0) continue;
    for (int d = 0; d < 3; ++d) mean[d] = mean[d] / (double)real;
    double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double maxr2 = 0.0;
    for (int j = 0; j < k; ++j) {
      int32_t id = ids[(size_t)i * k + j];
      if (id < 0) continue;
      double dx = (double)c->feat[4 * (int64_t)id + 0] - mean[0];
      double dy = (double)c->feat[4 * (int64_t)id + 1] - mean[1];
      double dz = (double)c->feat[4 * (int64_t)id + 2] - mean[2];
      C[0] += dx * dx; C[1] += dx * dy; C[2] += dx * dz;
      C[4] += dy * dy; C[5] += dy * dz; C[8] += dz * dz;
      double r2 = dx * dx + dy * dy + dz * dz;
      if (r2 > maxr2) maxr2 = r2;
    }
    C[0] = C[0] / (double)real; C[1] = C[1] / (double)real; C[2] = C[2] / (double)real;
    C[4] = C[4] / (double)real; C[5] = C[5] / (double)real; C[8] = C[8] / (double)real;
    C[3] = C[1]; C[6] = C[2]; C[7] = C[5];
    if (flags & 32) {
      /* keepMeanDist: distance from the point to the mean of its neighbours
       * [UPSTREAM-RECALLED, SurfaceNormal.cpp: (point - mean).norm()]         */
      double ex = (double)c->feat[4 * i + 0] - mean[0];
      double ey = (double)c->feat[4 * i + 1] - mean[1];
      double ez = (double)c->feat[4 * i + 2] - mean[2];
      c->meandist[i] = (float)sqrt(ex * ex + ey * ey + ez * ez);
    }
    double w[3], V[9];
    orc_eig3_sym(C, w, V);
    if (flags & 64) {
      /* sortEigen: eigenvalues ascending, eigenvectors permuted with them
       * (stable: equal values keep their order)                              */
      int ord[3] = {0, 1, 2};
      for (int a = 1; a < 3; ++a)
        for (int b2 = a; b2 > 0 && w[ord[b2]] < w[ord[b2 - 1]]; --b2) {
          int tmp = ord[b2]; ord[b2] = ord[b2 - 1]; ord[b2 - 1] = tmp;
        }
      double w2[3], V2[9];
      for (int e = 0; e < 3; ++e) {
        w2[e] = w[ord[e]];
        for (int d = 0; d < 3; ++d) V2[e * 3 + d] = V[ord[e] * 3 + d];
      }
      memcpy(w, w2, sizeof(w));
      memcpy(V, V2, sizeof(V));
    }
    /* rank test (A.9): need rank(C) >= 2; threshold 3*eps_T relative       */
    double wmax = w[0] > w[1] ? w[0] : w[1];
    if (w[2] > wmax) wmax = w[2];
    int rank = 0;
    for (int e = 0; e < 3; ++e)
      if (w[e] > 3.0 * (double)FLT_EPSILON * wmax) ++rank;
    if (flags & 2) {
      double r = sqrt(maxr2);
      double vol = (4.0 / 3.0) * 3.14159265358979323846 * (r * r * r);
      c->dens[i] = (float)((double)real / vol);
    }
    if (rank >= 2) {
      int e0 = 0;
      for (int e = 1; e < 3; ++e)
        if (w[e] < w[e0]) e0 = e;
      if (flags & 1)
        for (int d = 0; d < 3; ++d) {
          float v = (float)V[e0 * 3 + d];
          c->normals[3 * i + d] = v < -1.f ? -1.f : (v > 1.f ? 1.f : v);
        }
      if (flags & 4)
        for (int e = 0; e < 3; ++e) c->eigval[3 * i + e] = (float)w[e];
      if (flags & 8)
        for (int e = 0; e < 9; ++e) c->eigvec[9 * i + e] = (float)V[e];
    }
  }
  free(ids);
  free(d2);
  return ORC_OK;
}

static int filter_observation_direction(const orc_filter *f, orc_cloud *c) {
  float s[3] = {(float)f->p0, (float)f->p1, (float)f->p2};
  free(c->obsdir);
  c->obsdir = (float *)calloc((size_t)(c->n + 1) * 3, sizeof(float));
  for (int64_t i = 0; i < c->n; ++i)
    for (int d = 0; d < 3; ++d) c->obsdir[3 * i + d] = s[d] - c->feat[4 * i + d];
  return ORC_OK;
}

static int filter_orient_normals(const orc_filter *f, orc_cloud *c) {
  if (!c->normals || !c->obsdir) return ORC_INVALID_FIELD;
  int toward = (int)f->i0;
  for (int64_t i = 0; i < c->n; ++i) {
    const float *o = c->obsdir + 3 * i;
    float *nn = c->normals + 3 * i;
    float s = o[0] * nn[0];
    s = s + o[1] * nn[1];
    s = s + o[2] * nn[2];
    int flip = toward ? (s < 0.f) : (s > 0.f);
    if (flip) { nn[0] = -nn[0]; nn[1] = -nn[1]; nn[2] = -nn[2]; }
  }
  return ORC_OK;
}

static int filter_simple_sensor_noise(const orc_filter *f, orc_cloud *c) {
  /* A.9 / SimpleSensorNoise.cpp [UPSTREAM-RECALLED, constants (verify)]    */
  static const float tab[5][3] = {
      {0.012f, 0.0068f, 0.0008f},  /* 0 Sick LMS-1xx   */
      {0.028f, 0.0013f, 0.0001f},  /* 1 Hokuyo URG-04LX */
      {0.018f, 0.0006f, 0.0015f},  /* 2 Hokuyo UTM-30LX */
      {0.f, 0.f, 0.f},             /* 3 Kinect / Xtion  */
      {0.004f, 0.0053f, -0.0092f}, /* 4 Sick Tim3xx     */
  };
  int st = (int)f->i0;
  float gain = (float)f->p0;
  if (st < 0 || st > 4) return ORC_INVALID_PARAMETER;
  free(c->noise);
  c->noise = (float *)calloc((size_t)(c->n + 1), sizeof(float));
  for (int64_t i = 0; i < c->n; ++i) {
    const float *p = c->feat + 4 * i;
    float r2 = p[0] * p[0];
    r2 = r2 + p[1] * p[1];
    r2 = r2 + p[2] * p[2];
    float v;
    if (st == 3) {
      v = 0.5f * 0.00285f;
      v = v * r2;
    } else {
      float r = sqrtf(r2);
      v = tab[st][1] * r;
      v = v + tab[st][2];
      if (v < tab[st][0]) v = tab[st][0];
    }
    c->noise[i] = gain * v;
  }
  return ORC_OK;
}

static int filter_dist(const orc_filter *f, orc_cloud *c, int is_max) {
  int dim = (int)f->i0;
  float lim = (float)f->p0;
  int64_t *keep = (int64_t *)malloc((size_t)(c->n + 1) * sizeof(int64_t));
  int64_t m = 0;
  for (int64_t i = 0; i < c->n; ++i) {
    const float *p = c->feat + 4 * i;
    float v, l;
    if (dim < 0) {
      v = p[0] * p[0];
      v = v + p[1] * p[1];
      v = v + p[2] * p[2];
      l = lim * lim;
    } else {
      v = p[dim];
      l = lim;
    }
    if (is_max ? (v < l) : (v > l)) keep[m++] = i;
  }
  cloud_select(c, keep, m);
  free(keep);
  return ORC_OK;
}

static int filter_bounding_box(const orc_filter *f, orc_cloud *c) {
  /* BoundingBoxDataPointsFilter [UPSTREAM-RECALLED]: strict inequalities */
  float lo[3] = {(float)f->box[0], (float)f->box[2], (float)f->box[4]};
  float hi[3] = {(float)f->box[1], (float)f->box[3], (float)f->box[5]};
  int remove_inside = (int)f->i0;
  int64_t *keep = (int64_t *)malloc((size_t)(c->n + 1) * sizeof(int64_t));
  int64_t m = 0;
  for (int64_t i = 0; i < c->n; ++i) {
    const float *p = c->feat + 4 * i;
    int in = p[0] > lo[0] && p[0] < hi[0] && p[1] > lo[1] && p[1] < hi[1] && p[2] > lo[2] && p[2] < hi[2];
    if (remove_inside ? !in : in) keep[m++] = i;
  }
  cloud_select(c, keep, m);
  free(keep);
  return ORC_OK;
}

static int filter_remove_nan(orc_cloud *c) {
  /* RemoveNaNDataPointsFilter [UPSTREAM-RECALLED, DataPointsFilters/RemoveNaN.cpp]:
   * a point survives iff none of its coordinates is NaN; order preserved.     */
  int64_t *keep = (int64_t *)malloc((size_t)(c->n + 1) * sizeof(int64_t));
  int64_t m = 0;
  for (int64_t i = 0; i < c->n; ++i) {
    const float *p = c->feat + 4 * i;
    if (!(isnan(p[0]) || isnan(p[1]) || isnan(p[2]))) keep[m++] = i;
  }
  cloud_select(c, keep, m);
  free(keep);
  return ORC_OK;
}

static int filter_fix_step_sampling(const orc_filter *f, orc_cloud *c) {
  /* FixStepSamplingDataPointsFilter [UPSTREAM-RECALLED, DataPointsFilters/FixStepSampling.cpp]
   * with stepMult = 1 (constant step): keeps points phase, phase+step, ...; upstream draws
   * the phase with rand() % step, here it is a hash of the seed.            */
  int64_t step = f->i0 < 1 ? 1 : f->i0;
  int64_t phase = (int64_t)(splitmix64(splitmix64((uint64_t)f->i1)) % (uint64_t)step);
  int64_t *keep = (int64_t *)malloc((size_t)(c->n + 1) * sizeof(int64_t));
  int64_t m = 0;
  for (int64_t i = phase; i < c->n; i += step) keep[m++] = i;
  cloud_select(c, keep, m);
  free(keep);
  return ORC_OK;
}

static int filter_shadow(const orc_filter *f, orc_cloud *c) {
  /* ShadowDataPointsFilter [UPSTREAM-RECALLED, DataPointsFilters/Shadow.cpp]: a point whose
   * normal is (nearly) perpendicular to the ray from the origin lies on a grazing surface or
   * a depth discontinuity; keep iff |normalized(normal) . normalized(point)| > eps.
   * fp32, sums in x,y,z order, v / sqrt(v.v) (a zero vector stays zero).               */
  if (!c->normals) return ORC_INVALID_FIELD;
  float eps = (float)f->p0;
  int64_t *keep = (int64_t *)malloc((size_t)(c->n + 1) * sizeof(int64_t));
  int64_t m = 0;
  for (int64_t i = 0; i < c->n; ++i) {
    const float *p = c->feat + 4 * i, *nr = c->normals + 3 * i;
    float v[2][3] = {{nr[0], nr[1], nr[2]}, {p[0], p[1], p[2]}};
    for (int k = 0; k < 2; ++k) {
      float s = v[k][0] * v[k][0];
      s = s + v[k][1] * v[k][1];
      s = s + v[k][2] * v[k][2];
      if (s > 0.0f) {
        float len = sqrtf(s);
        v[k][0] = v[k][0] / len; v[k][1] = v[k][1] / len; v[k][2] = v[k][2] / len;
      }
    }
    float d = v[0][0] * v[1][0];
    d = d + v[0][1] * v[1][1];
    d = d + v[0][2] * v[1][2];
    if (fabsf(d) > eps) keep[m++] = i;
  }
  cloud_select(c, keep, m);
  free(keep);
  return ORC_OK;
}

static int filter_max_density(const orc_filter *f, orc_cloud *c) {
  /* MaxDensityDataPointsFilter [UPSTREAM-RECALLED, DataPointsFilters/MaxDensity.cpp]:
   * a point denser than maxDensity survives with probability maxDensity/density;
   * points at the saturation value (the cloud's maximum density) have that
   * probability multiplied by (1 - nbSaturated/nbPoints) in INTEGER arithmetic,
   * i.e. by 1 unless every point is saturated (then 0).  rand() -> counter hash (H7). */
  if (!c->dens) return ORC_INVALID_FIELD;
  float max_density = (float)f->p0;
  int64_t n = c->n;
  float last = -INFINITY;
  for (int64_t i = 0; i < n; ++i)
    if (c->dens[i] > last) last = c->dens[i];
  int64_t saturated = 0;
  for (int64_t i = 0; i < n; ++i)
    if (c->dens[i] == last) ++saturated;
  float sat_factor = (float)(1 - (n > 0 ? saturated / n : 0));
  int64_t *keep = (int64_t *)malloc((size_t)(n + 1) * sizeof(int64_t));
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i) {
    float density = c->dens[i];
    if (density > max_density) {
      float r = hash_uniform((uint64_t)f->i0, (uint64_t)i);
      float accept = max_density / density;
      if (density == last) accept = accept * sat_factor;
      if (r < accept) keep[m++] = i;
    } else {
      keep[m++] = i;
    }
  }
  cloud_select(c, keep, m);
  free(keep);
  return ORC_OK;
}

/* SamplingSurfaceNormalDataPointsFilter [UPSTREAM-RECALLED,
 * DataPointsFilters/SamplingSurfaceNormal.cpp buildNew / fuseRange]:
 * the cloud is split recursively -- cut dimension = widest side of the INHERITED
 * box, left half = count - count/2 smallest points along it, the inherited box
 * is clipped at the first point of the right half -- until a cell holds at most
 * knn points; every cell gets ONE normal (smallest eigenvector of its scatter
 * matrix NN NN^T, not divided by the count) and is then subsampled: method 0
 * keeps each point with probability `ratio`, method 1 keeps one point placed at
 * the cell mean.  Cells wider than maxBoxDim or of rank < 2 are dropped.
 * Pinned here where upstream leaves it to std::nth_element / rand(): the
 * partition is the STABLE order by (coordinate, previous order) with floats
 * compared through their total order (-0 < +0); the uniform variate of point k
 * is the counter hash (H7); mean and scatter are fp64 (H4).                  */
typedef struct { uint32_t key; int32_t ord; int32_t id; } ssn_item;
static int cmp_ssn(const void *a, const void *b) {
  const ssn_item *x = (const ssn_item *)a, *y = (const ssn_item *)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->ord < y->ord ? -1 : (x->ord > y->ord ? 1 : 0);
}
static inline uint32_t ord_u32(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

typedef struct {
  const orc_filter *f;
  orc_cloud *c;
  int32_t *idx;
  ssn_item *scratch;
  uint8_t *keep;
} ssn_ctx;

static void ssn_fuse(ssn_ctx *s, int64_t first, int64_t last) {
  orc_cloud *c = s->c;
  const int flags = (int)s->f->i1;
  const int cnt = (int)(last - first);
  float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  double mean[3] = {0, 0, 0};
  for (int j = 0; j < cnt; ++j) {
    const float *p = c->feat + 4 * (int64_t)s->idx[first + j];
    for (int d = 0; d < 3; ++d) {
      if (j == 0 || p[d] < lo[d]) lo[d] = p[d];
      if (j == 0 || p[d] > hi[d]) hi[d] = p[d];
      mean[d] += (double)p[d];
    }
  }
  float box_dim = hi[0] - lo[0];
  if (hi[1] - lo[1] > box_dim) box_dim = hi[1] - lo[1];
  if (hi[2] - lo[2] > box_dim) box_dim = hi[2] - lo[2];
  if (box_dim > (float)s->f->p1) return; /* drop box if it is too large */
  for (int d = 0; d < 3; ++d) mean[d] = mean[d] / (double)cnt;
  double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, maxr2 = 0.0;
  for (int j = 0; j < cnt; ++j) {
    const float *p = c->feat + 4 * (int64_t)s->idx[first + j];
    double dx = (double)p[0] - mean[0], dy = (double)p[1] - mean[1], dz = (double)p[2] - mean[2];
    C[0] += dx * dx; C[1] += dx * dy; C[2] += dx * dz;
    C[4] += dy * dy; C[5] += dy * dz; C[8] += dz * dz;
    double r2 = dx * dx + dy * dy + dz * dz;
    if (r2 > maxr2) maxr2 = r2;
  }
  C[3] = C[1]; C[6] = C[2]; C[7] = C[5];
  double w[3] = {1.0, 0.0, 0.0}, V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (flags & (1 | 4 | 8)) {
    orc_eig3_sym(C, w, V);
    double wmax = w[0] > w[1] ? w[0] : w[1];
    if (w[2] > wmax) wmax = w[2];
    int rank = 0;
    for (int e = 0; e < 3; ++e)
      if (w[e] > 3.0 * (double)FLT_EPSILON * wmax) ++rank;
    if (rank < 2) return; /* degenerate cell */
  }
  float normal[3] = {0, 0, 0};
  if (flags & 1) {
    int e0 = 0;
    for (int e = 1; e < 3; ++e)
      if (w[e] < w[e0]) e0 = e;
    for (int d = 0; d < 3; ++d) {
      float v = (float)V[e0 * 3 + d];
      normal[d] = v < -1.f ? -1.f : (v > 1.f ? 1.f : v);
    }
  }
  float density = 0.f;
  if (flags & 2) {
    double r = sqrt(maxr2);
    density = (float)((double)cnt / ((4.0 / 3.0) * 3.14159265358979323846 * (r * r * r)));
  }
#define SSN_WRITE(k)                                                            \
  do {                                                                          \
    if (flags & 1) for (int d = 0; d < 3; ++d) c->normals[3 * (k) + d] = normal[d]; \
    if (flags & 2) c->dens[(k)] = density;                                       \
    if (flags & 4) for (int e = 0; e < 3; ++e) c->eigval[3 * (k) + e] = (float)w[e]; \
    if (flags & 8) for (int e = 0; e < 9; ++e) c->eigvec[9 * (k) + e] = (float)V[e]; \
  } while (0)
  if (!(flags & 16)) {
    for (int j = 0; j < cnt; ++j) {
      int64_t k = s->idx[first + j];
      if (hash_uniform((uint64_t)s->f->p2, (uint64_t)k) < (float)s->f->p0) {
        s->keep[k] = 1;
        SSN_WRITE(k);
      }
    }
  } else {
    int64_t k = s->idx[first];
    s->keep[k] = 1;
    if (flags & 32) {
      /* average the existing descriptors over the cell, in cell order, in T */
#define SSN_AVG(field, span)                                                    \
      if (c->field) {                                                            \
        for (int d = 0; d < (span); ++d) {                                       \
          float acc = 0.f;                                                       \
          for (int j = 0; j < cnt; ++j) acc = acc + c->field[(size_t)s->idx[first + j] * (span) + d]; \
          c->field[(size_t)k * (span) + d] = acc / (float)cnt;                   \
        }                                                                        \
      }
      SSN_AVG(normals, 3) SSN_AVG(obsdir, 3) SSN_AVG(noise, 1) SSN_AVG(dens, 1) SSN_AVG(eigval, 3)
      SSN_AVG(eigvec, 9) SSN_AVG(meandist, 1) SSN_AVG(matched, c->matched_span)
#undef SSN_AVG
    }
    for (int d = 0; d < 3; ++d) c->feat[4 * k + d] = (float)mean[d];
    c->feat[4 * k + 3] = 1.f;
    SSN_WRITE(k);
  }
#undef SSN_WRITE
}

static void ssn_build(ssn_ctx *s, int64_t first, int64_t last, const float *minv, const float *maxv) {
  const int64_t count = last - first;
  if (count <= s->f->i0) {
    if (count > 0) ssn_fuse(s, first, last);
    return;
  }
  int cut = 0;
  for (int d = 1; d < 3; ++d)
    if (maxv[d] - minv[d] > maxv[cut] - minv[cut]) cut = d;
  const int64_t right = count / 2, left = count - right;
  for (int64_t j = 0; j < count; ++j) {
    s->scratch[j].key = ord_u32(s->c->feat[4 * (int64_t)s->idx[first + j] + cut]);
    s->scratch[j].ord = (int32_t)j;
    s->scratch[j].id = s->idx[first + j];
  }
  qsort(s->scratch, (size_t)count, sizeof(ssn_item), cmp_ssn);
  for (int64_t j = 0; j < count; ++j) s->idx[first + j] = s->scratch[j].id;
  const float cut_val = s->c->feat[4 * (int64_t)s->idx[first + left] + cut];
  float lmax[3] = {maxv[0], maxv[1], maxv[2]}, rmin[3] = {minv[0], minv[1], minv[2]};
  lmax[cut] = cut_val;
  rmin[cut] = cut_val;
  ssn_build(s, first, first + left, minv, lmax);
  ssn_build(s, first + left, last, rmin, maxv);
}

static int filter_sampling_surface_normal(const orc_filter *f, orc_cloud *c) {
  const int64_t n = c->n;
  const int flags = (int)f->i1;
  if (f->i0 < 3) return ORC_INVALID_PARAMETER;
  if ((flags & 1) && !c->normals) c->normals = (float *)calloc((size_t)(n + 1) * 3, sizeof(float));
  if ((flags & 2) && !c->dens) c->dens = (float *)calloc((size_t)(n + 1), sizeof(float));
  if ((flags & 4) && !c->eigval) c->eigval = (float *)calloc((size_t)(n + 1) * 3, sizeof(float));
  if ((flags & 8) && !c->eigvec) c->eigvec = (float *)calloc((size_t)(n + 1) * 9, sizeof(float));
  ssn_ctx s;
  s.f = f;
  s.c = c;
  s.idx = (int32_t *)malloc((size_t)(n + 1) * sizeof(int32_t));
  s.scratch = (ssn_item *)malloc((size_t)(n + 1) * sizeof(ssn_item));
  s.keep = (uint8_t *)calloc((size_t)(n + 1), 1);
  float minv[3] = {0, 0, 0}, maxv[3] = {0, 0, 0};
  for (int64_t i = 0; i < n; ++i) {
    s.idx[i] = (int32_t)i;
    for (int d = 0; d < 3; ++d) {
      float v = c->feat[4 * i + d];
      if (i == 0 || v < minv[d]) minv[d] = v;
      if (i == 0 || v > maxv[d]) maxv[d] = v;
    }
  }
  ssn_build(&s, 0, n, minv, maxv);
  int64_t *sel = (int64_t *)malloc((size_t)(n + 1) * sizeof(int64_t));
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i)
    if (s.keep[i]) sel[m++] = i;
  cloud_select(c, sel, m);
  free(sel); free(s.idx); free(s.scratch); free(s.keep);
  return ORC_OK;
}

int orc_filter_apply(const orc_filter *f, orc_cloud *c) {
  switch (f->type) {
    case ORC_F_SAMPLING_SURFACE_NORMAL: return filter_sampling_surface_normal(f, c);
    case ORC_F_MAX_DENSITY: return filter_max_density(f, c);
    case ORC_F_BOUNDING_BOX: return filter_bounding_box(f, c);
    case ORC_F_RANDOM_SAMPLING: return filter_random_sampling(f, c);
    case ORC_F_VOXEL_GRID: return filter_voxel_grid(f, c);
    case ORC_F_SURFACE_NORMAL: return filter_surface_normal(f, c);
    case ORC_F_OBSERVATION_DIRECTION: return filter_observation_direction(f, c);
    case ORC_F_ORIENT_NORMALS: return filter_orient_normals(f, c);
    case ORC_F_SIMPLE_SENSOR_NOISE: return filter_simple_sensor_noise(f, c);
    case ORC_F_IDENTITY: return ORC_OK;
    case ORC_F_REMOVE_NAN: return filter_remove_nan(c);
    case ORC_F_FIX_STEP_SAMPLING: return filter_fix_step_sampling(f, c);
    case ORC_F_SHADOW: return filter_shadow(f, c);
    case ORC_F_MAX_DIST: return filter_dist(f, c, 1);
    case ORC_F_MIN_DIST: return filter_dist(f, c, 0);
    default: return ORC_INVALID_PARAMETER;
  }
}

int orc_filters_apply(const orc_filter *f, int nf, orc_cloud *c) {
  for (int i = 0; i < nf; ++i) {
    int st = orc_filter_apply(&f[i], c);
    if (st != ORC_OK) return st;
  }
  return ORC_OK;
}

int orc_rigid_transform(orc_cloud *c, const double *T) {
  /* A.9 RigidTransformation::compute; rigidity eps 0.001 (verify)          */
  if (fabs(1.0 - m3_det_of4(T)) > 0.001) return ORC_TRANSFORMATION_ERROR;
  float Tf[16];
  for (int i = 0; i < 16; ++i) Tf[i] = (float)T[i];
  for (int64_t i = 0; i < c->n; ++i) {
    float o[4];
    xform_point_f32(Tf, c->feat + 4 * i, o);
    memcpy(c->feat + 4 * i, o, 3 * sizeof(float));
    if (c->normals) { rot_vec_f32(Tf, c->normals + 3 * i, o); memcpy(c->normals + 3 * i, o, 12); }
    if (c->obsdir) { rot_vec_f32(Tf, c->obsdir + 3 * i, o); memcpy(c->obsdir + 3 * i, o, 12); }
  }
  return ORC_OK;
}

/* ======================================================================== */
/* outlier filters (A.3)                                                    */
/* ======================================================================== */
static int cmp_f32(const void *a, const void *b) {
  float x = *(const float *)a, y = *(const float *)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

int orc_dists_quantile(const float *d2, int64_t nk, double q, float *out) {
  float *vals = (float *)malloc((size_t)(nk + 1) * sizeof(float));
  int64_t m = 0;
  for (int64_t i = 0; i < nk; ++i)
    if (!isinf(d2[i]) && d2[i] > 0.f) vals[m++] = d2[i];
  if (m == 0) { free(vals); return ORC_CONVERGENCE_ERROR; }
  if (q < 0.0 || q > 1.0) { free(vals); return ORC_CONVERGENCE_ERROR; }
  qsort(vals, (size_t)m, sizeof(float), cmp_f32); /* exact order statistic  */
  if (q == 1.0) *out = vals[m - 1];
  else {
    /* upstream: values.size() * quantile with quantile of type T = float (Matches.cpp), so the
     * product is a float; (float)0.85 > 0.85, which can move the truncated rank by one */
    size_t j = (size_t)((float)m * (float)q);
    if (j >= (size_t)m) j = (size_t)m - 1;
    *out = vals[j];
  }
  free(vals);
  return ORC_OK;
}

/* VarTrimmedDistOutlierFilter::optimizeInlierRatio [UPSTREAM-RECALLED,
 * OutlierFiltersImpl.cpp; Phillips et al., "Outlier robust ICP for minimizing
 * fractional RMSD"]: over the sorted valid squared distances s_1 <= ... <= s_M,
 *     FRMS(j) = (1 / (j/N)^lambda)^2 * (1/j) * (s_1 + ... + s_j),   N = k x N_r,
 * minimised over j in (floor(minRatio N), floor(maxRatio N)]; the optimal ratio is
 * (float)(j* - 1) / (float)N and the weight test is dist <= quantile(ratio).
 * Stated deviations: the running sum and FRMS are fp64 (upstream: T, in a
 * sequential order no parallel scan reproduces -- same rule as H4/H9), and the
 * range is clamped to the M valid distances (upstream reads past them when some
 * matches are invalid).                                                     */
int orc_var_trimmed_ratio(const float *d2, int64_t nk, double min_ratio, double max_ratio,
                          double lambda, float *ratio) {
  float *vals = (float *)malloc((size_t)(nk + 1) * sizeof(float));
  int64_t m = 0;
  for (int64_t i = 0; i < nk; ++i)
    if (!isinf(d2[i]) && d2[i] > 0.f) vals[m++] = d2[i];
  if (m == 0) { free(vals); return ORC_CONVERGENCE_ERROR; }
  qsort(vals, (size_t)m, sizeof(float), cmp_f32);
  int64_t min_el = (int64_t)floorf((float)min_ratio * (float)nk);
  int64_t max_el = (int64_t)floorf((float)max_ratio * (float)nk);
  if (max_el > m) max_el = m;
  if (min_el > max_el - 1) min_el = max_el - 1;
  if (min_el < 0) min_el = 0;
  double cum = 0.0, best = INFINITY;
  int64_t best_j = min_el;
  for (int64_t j = 0; j < max_el; ++j) {
    cum += (double)vals[j];
    if (j < min_el) continue;
    double id = (double)(j + 1);
    double deno = pow(id / (double)nk, lambda);
    double inv = 1.0 / deno;
    double frms = inv * inv * (1.0 / id) * cum;
    if (frms < best) { best = frms; best_j = j; }
  }
  *ratio = (float)best_j / (float)nk;
  free(vals);
  return ORC_OK;
}

int orc_outlier_weights(const orc_outlier *o, int no, const float *d2, int64_t nk, float *w) {
  if (no == 0) {
    for (int64_t i = 0; i < nk; ++i) w[i] = isinf(d2[i]) ? 0.f : 1.f;
    return ORC_OK;
  }
  for (int64_t i = 0; i < nk; ++i) w[i] = 1.f;
  for (int f = 0; f < no; ++f) {
    float limit = 0.f;
    int st;
    switch (o[f].type) {
      case ORC_O_TRIMMED_DIST:
        st = orc_dists_quantile(d2, nk, o[f].p0, &limit);
        if (st) return st;
        for (int64_t i = 0; i < nk; ++i) w[i] *= (d2[i] <= limit) ? 1.f : 0.f;
        break;
      case ORC_O_MAX_DIST:
        limit = (float)o[f].p0 * (float)o[f].p0;
        for (int64_t i = 0; i < nk; ++i) w[i] *= (d2[i] <= limit) ? 1.f : 0.f;
        break;
      case ORC_O_MIN_DIST:
        limit = (float)o[f].p0 * (float)o[f].p0;
        for (int64_t i = 0; i < nk; ++i) w[i] *= (d2[i] >= limit) ? 1.f : 0.f;
        break;
      case ORC_O_VAR_TRIMMED_DIST: {
        float ratio = 0.f;
        st = orc_var_trimmed_ratio(d2, nk, o[f].p0, o[f].p1, o[f].p2, &ratio);
        if (st) return st;
        st = orc_dists_quantile(d2, nk, (double)ratio, &limit);
        if (st) return st;
        for (int64_t i = 0; i < nk; ++i) w[i] *= (d2[i] <= limit) ? 1.f : 0.f;
        break;
      }
      case ORC_O_MEDIAN_DIST:
        st = orc_dists_quantile(d2, nk, 0.5, &limit);
        if (st) return st;
        limit = (float)o[f].p0 * limit;
        for (int64_t i = 0; i < nk; ++i) w[i] *= (d2[i] <= limit) ? 1.f : 0.f;
        break;
      default: return ORC_INVALID_PARAMETER;
    }
  }
  return ORC_OK;
}

int orc_outlier_weights_full(const orc_outlier *o, int no, const orc_cloud *reading,
                             const orc_cloud *reference, const int32_t *ids, const float *d2,
                             int k, float *w) {
  int64_t nk = reading->n * k;
  orc_outlier dist_only[ORC_MAX_MODS];
  int nd = 0, has_sn = 0;
  for (int f = 0; f < no; ++f) {
    if (o[f].type == ORC_O_SURFACE_NORMAL) has_sn = 1;
    else dist_only[nd++] = o[f];
  }
  int st;
  if (nd == 0 && no > 0) {
    for (int64_t i = 0; i < nk; ++i) w[i] = 1.f;
    st = ORC_OK;
  } else {
    st = orc_outlier_weights(dist_only, nd, d2, nk, w);
  }
  if (st || !has_sn) return st;
  for (int f = 0; f < no; ++f) {
    if (o[f].type != ORC_O_SURFACE_NORMAL) continue;
    /* SurfaceNormalOutlierFilter [UPSTREAM-RECALLED]: skipped (weights 1) unless
     * both clouds carry normals; w = 0 where the unit normals' dot < cos(maxAngle) */
    if (!reading->normals || !reference->normals) continue;
    float eps = (float)cos(o[f].p0);
    for (int64_t i = 0; i < reading->n; ++i) {
      const float *a = reading->normals + 3 * i;
      float na = a[0] * a[0];
      na = na + a[1] * a[1];
      na = na + a[2] * a[2];
      na = sqrtf(na);
      float ax = a[0] / na, ay = a[1] / na, az = a[2] / na;
      for (int kk = 0; kk < k; ++kk) {
        size_t m = (size_t)i * k + kk;
        if (ids[m] < 0) { w[m] = 0.f; continue; }
        const float *b = reference->normals + 3 * (int64_t)ids[m];
        float nb = b[0] * b[0];
        nb = nb + b[1] * b[1];
        nb = nb + b[2] * b[2];
        nb = sqrtf(nb);
        float bx = b[0] / nb, by = b[1] / nb, bz = b[2] / nb;
        float dot = ax * bx;
        dot = dot + ay * by;
        dot = dot + az * bz;
        w[m] *= (dot < eps) ? 0.f : 1.f;
      }
    }
  }
  return ORC_OK;
}

/* ======================================================================== */
/* error minimizers (A.4 - A.7)                                             */
/* ======================================================================== */
static void angle_axis_to_T(const double *x, double *T) {
  /* A.5: AngleAxis(|x0..2|, x0..2/|.|), translation x3..5; NaN -> R = I   */
  m4_identity(T);
  double th = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  if (th > 0.0 && isfinite(th)) {
    double ax = x[0] / th, ay = x[1] / th, az = x[2] / th;
    double c = cos(th), s = sin(th), v = 1.0 - c;
    T[0] = c + ax * ax * v;       T[4] = ax * ay * v - az * s;  T[8] = ax * az * v + ay * s;
    T[1] = ay * ax * v + az * s;  T[5] = c + ay * ay * v;       T[9] = ay * az * v - ax * s;
    T[2] = az * ax * v - ay * s;  T[6] = az * ay * v + ax * s;  T[10] = c + az * az * v;
  }
  T[12] = x[3]; T[13] = x[4]; T[14] = x[5];
}

static void inv6_sym(const double *H, double *Hi) {
  for (int c = 0; c < 6; ++c) {
    double e[6] = {0, 0, 0, 0, 0, 0}, x[6];
    e[c] = 1.0;
    orc_solve6(H, e, x);
    for (int r = 0; r < 6; ++r) Hi[c * 6 + r] = x[r];
  }
}

int orc_minimize(int type, double sensor_std_dev, const orc_cloud *reading,
                 const orc_cloud *reference, const int32_t *ids, const float *d2,
                 const float *w, int k, orc_min_out *out) {
  return orc_minimize_ex(type, ORC_FORCE_NONE, sensor_std_dev, reading, reference, ids, d2, w, k, out);
}

/* PointToPlaneErrorMinimizer force2D / force4DOF (A.5, upstream
 * ErrorMinimizers/PointToPlane.cpp compute_in_place [UPSTREAM-RECALLED]):
 *   force2D   : features and normals cut to x,y; F = [x*ny - y*nx; nx; ny];
 *               e = dx*nx + dy*ny; unknowns [theta, tx, ty]; Rotation2D(theta)
 *               written into the top-left of a 4x4 identity.
 *   force4DOF : F = [(Gamma p).n; nx; ny; nz] with Gamma p = (-y, x, 0);
 *               e = d.n in 3-D; unknowns [yaw, tx, ty, tz]; AngleAxis(yaw, Z). */
static void yaw_to_T(double th, double tx, double ty, double tz, double *T) {
  m4_identity(T);
  double c = cos(th), s = sin(th);
  T[0] = c; T[4] = -s;
  T[1] = s; T[5] = c;
  T[12] = tx; T[13] = ty; T[14] = tz;
}

/* A.6 estimateCovariance (Censi) from the error elements and the transform just computed; order
 * x,y,z,rx,ry,rz.  The estimate is the 6-DOF one whatever the solve was restricted to (upstream's
 * PointToPlaneWithCov calls it on the result of compute_in_place).                              */
static void censi_covariance(orc_min_out *out, double sensor_std_dev, const orc_cloud *reading,
                             const orc_cloud *reference, const int32_t *ids, const float *d2,
                             const float *w, int k) {
  const int64_t nr = reading->n;
    /* A.6 estimateCovariance (Censi), order x,y,z,rx,ry,rz              */
    const double *T = out->T;
    double beta = -asin(T[2]);
    double alpha = atan2(T[6], T[10]);
    double cb = cos(beta);
    double gamma = atan2(T[1] / cb, T[0] / cb);
    double t[3] = {T[12], T[13], T[14]};
    double H[36], DD[36];
    memset(H, 0, sizeof(H));
    memset(DD, 0, sizeof(DD));
    for (int64_t i = 0; i < nr; ++i)
      for (int kk = 0; kk < k; ++kk) {
        size_t m = (size_t)i * k + kk;
        if (isinf(d2[m]) || w[m] == 0.f) continue;
        const float *pf = reading->feat + 4 * i;
        const float *qf = reference->feat + 4 * (int64_t)ids[m];
        const float *nf = reference->normals + 3 * (int64_t)ids[m];
        double p[3] = {pf[0], pf[1], pf[2]}, q[3] = {qf[0], qf[1], qf[2]};
        double n[3] = {nf[0], nf[1], nf[2]};
        double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        double rp = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
        double rq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
        if (!(nn > 0.0) || !(rp > 0.0) || !(rq > 0.0)) continue; /* degenerate pair */
        for (int d = 0; d < 3; ++d) n[d] = n[d] / nn;
        double dp[3] = {p[0] / rp, p[1] / rp, p[2] / rp};
        double dq[3] = {q[0] / rq, q[1] / rq, q[2] / rq};
        double na = n[2] * dp[1] - n[1] * dp[2];
        double nb = n[0] * dp[2] - n[2] * dp[0];
        double ng = n[1] * dp[0] - n[0] * dp[1];
        double E = n[0] * (p[0] - gamma * p[1] + beta * p[2] + t[0] - q[0]);
        E += n[1] * (gamma * p[0] + p[1] - alpha * p[2] + t[1] - q[1]);
        E += n[2] * (-beta * p[0] + alpha * p[1] + p[2] + t[2] - q[2]);
        double Np = n[0] * (dp[0] - gamma * dp[1] + beta * dp[2]);
        Np += n[1] * (gamma * dp[0] + dp[1] - alpha * dp[2]);
        Np += n[2] * (-beta * dp[0] + alpha * dp[1] + dp[2]);
        double Nq = -(n[0] * dq[0] + n[1] * dq[1] + n[2] * dq[2]);
        double g[6] = {n[0], n[1], n[2], rp * na, rp * nb, rp * ng};
        double en = E + rp * Np;
        double u[6] = {n[0] * Np, n[1] * Np, n[2] * Np, na * en, nb * en, ng * en};
        double v[6] = {n[0] * Nq, n[1] * Nq, n[2] * Nq, rq * na * Nq, rq * nb * Nq, rq * ng * Nq};
        for (int c = 0; c < 6; ++c)
          for (int r = 0; r <= c; ++r) {
            H[c * 6 + r] += g[c] * g[r];
            DD[c * 6 + r] += u[c] * u[r] + v[c] * v[r];
          }
      }
    for (int c = 0; c < 6; ++c)
      for (int r = 0; r < c; ++r) { H[r * 6 + c] = H[c * 6 + r]; DD[r * 6 + c] = DD[c * 6 + r]; }
    double Hi[36], tmp[36];
    inv6_sym(H, Hi);
    for (int c = 0; c < 6; ++c)
      for (int r = 0; r < 6; ++r) {
        double s = 0.0;
        for (int kx = 0; kx < 6; ++kx) s += Hi[kx * 6 + r] * DD[c * 6 + kx];
        tmp[c * 6 + r] = s;
      }
    double s2 = sensor_std_dev * sensor_std_dev;
    for (int c = 0; c < 6; ++c)
      for (int r = 0; r < 6; ++r) {
        double s = 0.0;
        for (int kx = 0; kx < 6; ++kx) s += tmp[kx * 6 + r] * Hi[c * 6 + kx];
        out->cov[c * 6 + r] = s2 * s;
      }
  }

int orc_minimize_ex(int type, int force_mode, double sensor_std_dev, const orc_cloud *reading,
                    const orc_cloud *reference, const int32_t *ids, const float *d2,
                    const float *w, int k, orc_min_out *out) {
  memset(out, 0, sizeof(*out));
  m4_identity(out->T);
  int64_t nr = reading->n;
  int p2plane = (type == ORC_E_POINT_TO_PLANE || type == ORC_E_POINT_TO_PLANE_WITH_COV);
  if (p2plane && !reference->normals) return ORC_INVALID_FIELD;
  /* force2D / force4DOF: PointToPlane; force4DOF also with the covariance estimate (force2D cuts the
   * error elements to 2-D upstream, which the 6-DOF covariance formula cannot take)                */
  if (force_mode != ORC_FORCE_NONE && !(type == ORC_E_POINT_TO_PLANE ||
                                        (type == ORC_E_POINT_TO_PLANE_WITH_COV && force_mode == ORC_FORCE_4DOF)))
    return ORC_INVALID_PARAMETER;
  /* ErrorElements (A.4) */
  int64_t kept = 0;
  double wsum = 0.0;
  for (int64_t i = 0; i < nr; ++i)
    for (int kk = 0; kk < k; ++kk) {
      size_t m = (size_t)i * k + kk;
      if (isinf(d2[m])) continue;
      if (w[m] != 0.f) { ++kept; wsum += (double)w[m]; }
    }
  if (kept == 0) return ORC_CONVERGENCE_ERROR; /* "no point to minimize" */
  out->kept = kept;
  out->point_used_ratio = (double)kept / (double)(k * nr);
  out->weighted_point_used_ratio = wsum / (double)(k * nr);

  if (p2plane && force_mode != ORC_FORCE_NONE) {
    const int nd = force_mode == ORC_FORCE_2D ? 3 : 4;
    double A[16], b[4];
    memset(A, 0, sizeof(A));
    memset(b, 0, sizeof(b));
    double resid = 0.0;
    for (int64_t i = 0; i < nr; ++i)
      for (int kk = 0; kk < k; ++kk) {
        size_t m = (size_t)i * k + kk;
        if (isinf(d2[m]) || w[m] == 0.f) continue;
        const float *pf = reading->feat + 4 * i;
        const float *qf = reference->feat + 4 * (int64_t)ids[m];
        const float *nf = reference->normals + 3 * (int64_t)ids[m];
        double p[3] = {pf[0], pf[1], pf[2]}, n[3] = {nf[0], nf[1], nf[2]};
        double wt = (double)w[m];
        double F[4];
        F[0] = p[0] * n[1] - p[1] * n[0];
        F[1] = n[0]; F[2] = n[1]; F[3] = n[2];
        double e = (p[0] - (double)qf[0]) * n[0] + (p[1] - (double)qf[1]) * n[1];
        if (force_mode == ORC_FORCE_4DOF) e += (p[2] - (double)qf[2]) * n[2];
        for (int c = 0; c < nd; ++c) {
          double wf = wt * F[c];
          for (int r = 0; r <= c; ++r) A[c * nd + r] += wf * F[r];
          b[c] -= wf * e;
        }
        resid += wt * (e * e);
      }
    for (int c = 0; c < nd; ++c)
      for (int r = 0; r < c; ++r) A[r * nd + c] = A[c * nd + r];
    out->residual = resid;
    double x[4] = {0, 0, 0, 0};
    solve_n(nd, A, b, x);
    yaw_to_T(x[0], x[1], x[2], force_mode == ORC_FORCE_4DOF ? x[3] : 0.0, out->T);
    if (type == ORC_E_POINT_TO_PLANE_WITH_COV) censi_covariance(out, sensor_std_dev, reading, reference, ids, d2, w, k);
    return ORC_OK;
  }

  if (p2plane) {
    double A[36], b[6];
    memset(A, 0, sizeof(A));
    memset(b, 0, sizeof(b));
    double resid = 0.0;
    for (int64_t i = 0; i < nr; ++i)
      for (int kk = 0; kk < k; ++kk) {
        size_t m = (size_t)i * k + kk;
        if (isinf(d2[m]) || w[m] == 0.f) continue;
        const float *pf = reading->feat + 4 * i;
        const float *qf = reference->feat + 4 * (int64_t)ids[m];
        const float *nf = reference->normals + 3 * (int64_t)ids[m];
        double p[3] = {pf[0], pf[1], pf[2]}, n[3] = {nf[0], nf[1], nf[2]};
        double wt = (double)w[m];
        double F[6];
        F[0] = p[1] * n[2] - p[2] * n[1];
        F[1] = p[2] * n[0] - p[0] * n[2];
        F[2] = p[0] * n[1] - p[1] * n[0];
        F[3] = n[0]; F[4] = n[1]; F[5] = n[2];
        double e = (p[0] - (double)qf[0]) * n[0] + (p[1] - (double)qf[1]) * n[1] +
                   (p[2] - (double)qf[2]) * n[2];
        for (int c = 0; c < 6; ++c) {
          double wf = wt * F[c];
          for (int r = 0; r <= c; ++r) A[c * 6 + r] += wf * F[r];
          b[c] -= wf * e;
        }
        resid += wt * (e * e);
      }
    for (int c = 0; c < 6; ++c)
      for (int r = 0; r < c; ++r) A[r * 6 + c] = A[c * 6 + r];
    memcpy(out->A, A, sizeof(A));
    memcpy(out->b, b, sizeof(b));
    out->residual = resid;
    double x[6];
    orc_solve6(A, b, x);
    angle_axis_to_T(x, out->T);

    if (type == ORC_E_POINT_TO_PLANE_WITH_COV) censi_covariance(out, sensor_std_dev, reading, reference, ids, d2, w, k);
    return ORC_OK;
  }

  /* point-to-point (A.7): weighted centroids, 3x3 cross-covariance, SVD    */
  double W = 0.0, mp[3] = {0, 0, 0}, mq[3] = {0, 0, 0};
  for (int64_t i = 0; i < nr; ++i)
    for (int kk = 0; kk < k; ++kk) {
      size_t m = (size_t)i * k + kk;
      if (isinf(d2[m]) || w[m] == 0.f) continue;
      const float *pf = reading->feat + 4 * i;
      const float *qf = reference->feat + 4 * (int64_t)ids[m];
      double wt = (double)w[m];
      W += wt;
      for (int d = 0; d < 3; ++d) { mp[d] += wt * (double)pf[d]; mq[d] += wt * (double)qf[d]; }
    }
  for (int d = 0; d < 3; ++d) { mp[d] = mp[d] / W; mq[d] = mq[d] / W; }
  double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double resid = 0.0;
  for (int64_t i = 0; i < nr; ++i)
    for (int kk = 0; kk < k; ++kk) {
      size_t m = (size_t)i * k + kk;
      if (isinf(d2[m]) || w[m] == 0.f) continue;
      const float *pf = reading->feat + 4 * i;
      const float *qf = reference->feat + 4 * (int64_t)ids[m];
      double wt = (double)w[m];
      double pc[3], qc[3], dl[3];
      for (int d = 0; d < 3; ++d) {
        pc[d] = (double)pf[d] - mp[d];
        qc[d] = (double)qf[d] - mq[d];
        dl[d] = (double)pf[d] - (double)qf[d];
      }
      for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) M[c * 3 + r] += wt * qc[r] * pc[c];
      resid += sqrt(dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2]);
    }
  out->residual = resid;
  double U[9], S[3], V[9], R[9];
  orc_svd3(M, U, S, V);
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) {
      double s = 0.0;
      for (int kx = 0; kx < 3; ++kx) s += U[kx * 3 + r] * V[kx * 3 + c];
      R[c * 3 + r] = s;
    }
  double det = R[0] * (R[4] * R[8] - R[7] * R[5]) - R[3] * (R[1] * R[8] - R[7] * R[2]) +
               R[6] * (R[1] * R[5] - R[4] * R[2]);
  if (det < 0.0) {
    /* negate the last row of V^T == last column of V */
    for (int r = 0; r < 3; ++r) V[6 + r] = -V[6 + r];
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) {
        double s = 0.0;
        for (int kx = 0; kx < 3; ++kx) s += U[kx * 3 + r] * V[kx * 3 + c];
        R[c * 3 + r] = s;
      }
  }
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) out->T[c * 4 + r] = R[c * 3 + r];
  for (int r = 0; r < 3; ++r) {
    double s = 0.0;
    for (int kx = 0; kx < 3; ++kx) s += R[kx * 3 + r] * mp[kx];
    out->T[12 + r] = mq[r] - s;
  }
  return ORC_OK;
}

double orc_overlap(int type, const orc_cloud *reading, const orc_cloud *reference,
                   const int32_t *ids, const float *d2, const float *w, int k) {
  /* A.5 getOverlap: falls back to weightedPointUsedRatio unless the reading
   * carries both simpleSensorNoise and normals                             */
  int64_t nr = reading->n, kept = 0, good = 0;
  double wsum = 0.0;
  int have = reading->noise && reading->normals && type != ORC_E_POINT_TO_POINT;
  for (int64_t i = 0; i < nr; ++i)
    for (int kk = 0; kk < k; ++kk) {
      size_t m = (size_t)i * k + kk;
      if (isinf(d2[m]) || w[m] == 0.f) continue;
      ++kept;
      wsum += (double)w[m];
      if (have) {
        const float *pf = reading->feat + 4 * i;
        const float *qf = reference->feat + 4 * (int64_t)ids[m];
        const float *nf = reading->normals + 3 * i;
        double n[3] = {nf[0], nf[1], nf[2]};
        double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        double e = 0.0;
        for (int d = 0; d < 3; ++d) e += ((double)pf[d] - (double)qf[d]) * (n[d] / nn);
        if (fabs(e) < (double)reading->noise[i]) ++good;
      }
    }
  if (!have) return wsum / (double)(k * nr);
  return kept ? (double)good / (double)kept : 0.0;
}

/* ======================================================================== */
/* transformation checkers (A.8)                                            */
/* ======================================================================== */
static void quat_from_T(const double *T, double *q /* w,x,y,z */) {
  double m00 = T[0], m11 = T[5], m22 = T[10];
  double tr = m00 + m11 + m22;
  if (tr > 0.0) {
    double s = sqrt(tr + 1.0);
    q[0] = 0.5 * s;
    s = 0.5 / s;
    q[1] = (T[6] - T[9]) * s;
    q[2] = (T[8] - T[2]) * s;
    q[3] = (T[1] - T[4]) * s;
  } else {
    int i = 0;
    if (m11 > m00) i = 1;
    if (m22 > (i == 0 ? m00 : m11)) i = 2;
    int j = (i + 1) % 3, kx = (j + 1) % 3;
#define MM(r, c) T[(c) * 4 + (r)]
    double s = sqrt(MM(i, i) - MM(j, j) - MM(kx, kx) + 1.0);
    q[1 + i] = 0.5 * s;
    s = 0.5 / s;
    q[0] = (MM(kx, j) - MM(j, kx)) * s;
    q[1 + j] = (MM(j, i) + MM(i, j)) * s;
    q[1 + kx] = (MM(kx, i) + MM(i, kx)) * s;
#undef MM
  }
}

static double quat_angular_distance(const double *a, const double *b) {
  /* d = a * conj(b); 2*atan2(|vec d|, |w d|)                               */
  double w = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
  double x = -a[0] * b[1] + a[1] * b[0] - a[2] * b[3] + a[3] * b[2];
  double y = -a[0] * b[2] + a[1] * b[3] + a[2] * b[0] - a[3] * b[1];
  double z = -a[0] * b[3] - a[1] * b[2] + a[2] * b[1] + a[3] * b[0];
  return 2.0 * atan2(sqrt(x * x + y * y + z * z), fabs(w));
}

#define CHK_HIST 64
typedef struct {
  const orc_icp_config *cfg;
  int counter;
  int nhist;
  double q[CHK_HIST][4], t[CHK_HIST][3];
  double q0[4], t0[3];
} checkers;

static void chk_push(checkers *c, const double *T) {
  if (c->nhist == CHK_HIST) {
    memmove(c->q[0], c->q[1], sizeof(double) * 4 * (CHK_HIST - 1));
    memmove(c->t[0], c->t[1], sizeof(double) * 3 * (CHK_HIST - 1));
    c->nhist--;
  }
  quat_from_T(T, c->q[c->nhist]);
  c->t[c->nhist][0] = T[12]; c->t[c->nhist][1] = T[13]; c->t[c->nhist][2] = T[14];
  c->nhist++;
}

static void chk_init(checkers *c, const orc_icp_config *cfg, const double *T) {
  memset(c, 0, sizeof(*c));
  c->cfg = cfg;
  chk_push(c, T);
  memcpy(c->q0, c->q[0], sizeof(c->q0));
  memcpy(c->t0, c->t[0], sizeof(c->t0));
}

/* returns status; clears *iterate / sets *max_reached as the checkers say  */
static int chk_check(checkers *c, const double *T, int *iterate, int *max_reached) {
  const orc_icp_config *cfg = c->cfg;
  int status = ORC_OK;
  if (cfg->max_iterations > 0) {
    c->counter++;
    if (c->counter >= cfg->max_iterations) { *iterate = 0; *max_reached = 1; }
  }
  chk_push(c, T);
  if (cfg->has_differential) {
    int sl = cfg->smooth_length;
    if (c->nhist > sl) {
      double c0 = 0.0, c1 = 0.0;
      for (int i = c->nhist - 1; i >= c->nhist - sl; --i) {
        c0 += fabs(quat_angular_distance(c->q[i], c->q[i - 1]));
        double dx = c->t[i][0] - c->t[i - 1][0], dy = c->t[i][1] - c->t[i - 1][1],
               dz = c->t[i][2] - c->t[i - 1][2];
        c1 += sqrt(dx * dx + dy * dy + dz * dz);
      }
      c0 = c0 / (double)sl;
      c1 = c1 / (double)sl;
      if (isnan(c0) || isnan(c1)) status = ORC_CONVERGENCE_ERROR;
      else if (c0 < cfg->min_diff_rot && c1 < cfg->min_diff_trans) *iterate = 0;
    }
  }
  if (cfg->has_bound) {
    double qn[4];
    quat_from_T(T, qn);
    double dx = T[12] - c->t0[0], dy = T[13] - c->t0[1], dz = T[14] - c->t0[2];
    if (quat_angular_distance(qn, c->q0) > cfg->max_rot_norm ||
        sqrt(dx * dx + dy * dy + dz * dz) > cfg->max_trans_norm)
      status = ORC_CONVERGENCE_ERROR;
  }
  return status;
}

/* ======================================================================== */
/* ICP chain (§3.3, A.8)                                                    */
/* ======================================================================== */
void orc_icp_config_default(orc_icp_config *cfg) {
  /* ICPChainBase::setDefault (A17): RandomSampling(0.75) reading filter,
   * SamplingSurfaceNormal reference filter, TrimmedDist(0.85), k = 1,
   * PointToPlane, Counter(40) + Differential(0.001, 0.001, 3)              */
  memset(cfg, 0, sizeof(*cfg));
  cfg->reading_filters[0].type = ORC_F_RANDOM_SAMPLING;
  cfg->reading_filters[0].p0 = 0.75;
  cfg->n_reading_filters = 1;
  cfg->reference_filters[0].type = ORC_F_SAMPLING_SURFACE_NORMAL;
  cfg->reference_filters[0].p0 = 0.5;
  cfg->reference_filters[0].p1 = INFINITY;
  cfg->reference_filters[0].i0 = 7;
  cfg->reference_filters[0].i1 = 1 | 32;
  cfg->n_reference_filters = 1;
  cfg->knn = 1;
  cfg->epsilon = 0.0;
  cfg->max_dist = INFINITY;
  cfg->outliers[0].type = ORC_O_TRIMMED_DIST;
  cfg->outliers[0].p0 = 0.85;
  cfg->n_outliers = 1;
  cfg->minimizer = ORC_E_POINT_TO_PLANE;
  cfg->sensor_std_dev = 0.01;
  cfg->max_iterations = 40;
  cfg->has_differential = 1;
  cfg->min_diff_rot = 0.001;
  cfg->min_diff_trans = 0.001;
  cfg->smooth_length = 3;
  cfg->has_bound = 0;
  cfg->max_rot_norm = 1.0;
  cfg->max_trans_norm = 1.0;
}

static void mean_centre(orc_cloud *ref, double *T_refIn_refMean) {
  double m[3] = {0, 0, 0};
  for (int64_t i = 0; i < ref->n; ++i)
    for (int d = 0; d < 3; ++d) m[d] += (double)ref->feat[4 * i + d];
  m4_identity(T_refIn_refMean);
  float mf[3];
  for (int d = 0; d < 3; ++d) {
    mf[d] = (float)(m[d] / (double)ref->n);
    T_refIn_refMean[12 + d] = (double)mf[d];
  }
  for (int64_t i = 0; i < ref->n; ++i)
    for (int d = 0; d < 3; ++d) ref->feat[4 * i + d] = ref->feat[4 * i + d] - mf[d];
}

static int icp_loop(const orc_icp_config *cfg, const orc_cloud *readingIn,
                    const orc_cloud *reference, const orc_kdtree *tree,
                    const double *T_refIn_refMean, const double *T_refIn_dataIn,
                    orc_icp_result *res) {
  double t0 = now_s();
  orc_cloud *reading = orc_cloud_copy(readingIn);
  int st = orc_filters_apply(cfg->reading_filters, cfg->n_reading_filters, reading);
  if (st) { orc_cloud_free(reading); return res->status = st; }
  double inv[16], T_refMean_dataIn[16];
  m4_rigid_inv(T_refIn_refMean, inv);
  m4_mul(inv, T_refIn_dataIn, T_refMean_dataIn);
  st = orc_rigid_transform(reading, T_refMean_dataIn);
  if (st) { orc_cloud_free(reading); return res->status = st; }
  res->time_filters_s += now_s() - t0;

  double tl = now_s();
  double T_iter[16];
  m4_identity(T_iter);
  int iterate = 1, max_reached = 0, iters = 0;
  checkers chk;
  chk_init(&chk, cfg, T_iter);
  int k = cfg->knn;
  int64_t cap = reading->n > 0 ? reading->n : 1;
  int32_t *ids = (int32_t *)malloc((size_t)cap * k * sizeof(int32_t));
  float *d2 = (float *)malloc((size_t)cap * k * sizeof(float));
  float *w = (float *)malloc((size_t)cap * k * sizeof(float));
  orc_min_out mo;
  memset(&mo, 0, sizeof(mo));
  orc_cloud *step = NULL;
  while (iterate) {
    orc_cloud_free(step);
    step = orc_cloud_copy(reading);
    st = orc_filters_apply(cfg->reading_step_filters, cfg->n_reading_step_filters, step);
    if (st) break;
    st = orc_rigid_transform(step, T_iter);
    if (st) break;
    double tk = now_s();
    res->visits += orc_kdtree_knn_ex(tree, step->feat, step->n, k, (float)cfg->max_dist, 1, (float)cfg->epsilon,
                                      g_search_mode, ids, d2);
    res->time_knn_s += now_s() - tk;
    st = orc_outlier_weights_full(cfg->outliers, cfg->n_outliers, step, reference, ids, d2, k, w);
    if (st) break;
    st = orc_minimize_ex(cfg->minimizer, cfg->force_mode, cfg->sensor_std_dev, step, reference, ids, d2, w, k, &mo);
    if (st) break;
    m4_mul(mo.T, T_iter, T_iter);
    ++iters;
    st = chk_check(&chk, T_iter, &iterate, &max_reached);
    if (st) break;
  }
  res->status = st;
  res->iterations = iters;
  res->max_iter_reached = max_reached;
  memcpy(res->last_T_iter, T_iter, sizeof(T_iter));
  if (st == ORC_OK && step) {
    res->weighted_ratio = mo.weighted_point_used_ratio;
    res->point_used_ratio = mo.point_used_ratio;
    res->residual = mo.residual;
    res->overlap = orc_overlap(cfg->minimizer, step, reference, ids, d2, w, k);
    memcpy(res->cov, mo.cov, sizeof(mo.cov));
  }
  double tmp[16];
  m4_mul(T_iter, T_refMean_dataIn, tmp);
  m4_mul(T_refIn_refMean, tmp, res->T);
  res->time_loop_s += now_s() - tl;
  orc_cloud_free(step);
  orc_cloud_free(reading);
  free(ids); free(d2); free(w);
  return st;
}

int orc_icp_run(const orc_icp_config *cfg, const orc_cloud *readingIn,
                const orc_cloud *referenceIn, const double *T_init, orc_icp_result *res) {
  memset(res, 0, sizeof(*res));
  double t0 = now_s();
  orc_cloud *ref = orc_cloud_copy(referenceIn);
  int st = orc_filters_apply(cfg->reference_filters, cfg->n_reference_filters, ref);
  if (st) { orc_cloud_free(ref); return res->status = st; }
  double T_refIn_refMean[16];
  if (ref->n == 0) { orc_cloud_free(ref); return res->status = ORC_CONVERGENCE_ERROR; }
  mean_centre(ref, T_refIn_refMean);
  res->time_filters_s += now_s() - t0;
  t0 = now_s();
  orc_kdtree *tree = orc_kdtree_build(ref->feat, ref->n);
  res->time_index_s += now_s() - t0;
  st = icp_loop(cfg, readingIn, ref, tree, T_refIn_refMean, T_init, res);
  orc_kdtree_free(tree);
  orc_cloud_free(ref);
  return st;
}

struct orc_icp_seq {
  orc_icp_config cfg;
  orc_cloud *map;
  orc_kdtree *tree;
  double T_refIn_refMean[16];
};

orc_icp_seq *orc_icp_seq_new(const orc_icp_config *cfg) {
  orc_icp_seq *s = (orc_icp_seq *)calloc(1, sizeof(orc_icp_seq));
  s->cfg = *cfg;
  m4_identity(s->T_refIn_refMean);
  return s;
}
void orc_icp_seq_free(orc_icp_seq *s) {
  if (!s) return;
  orc_kdtree_free(s->tree);
  orc_cloud_free(s->map);
  free(s);
}
int orc_icp_seq_set_map(orc_icp_seq *s, const orc_cloud *map) {
  /* ICPSequence::setMap: mean-centre FIRST, then reference filters (A16)   */
  orc_kdtree_free(s->tree); s->tree = NULL;
  orc_cloud_free(s->map);
  s->map = orc_cloud_copy(map);
  if (s->map->n == 0) return ORC_CONVERGENCE_ERROR;
  mean_centre(s->map, s->T_refIn_refMean);
  int st = orc_filters_apply(s->cfg.reference_filters, s->cfg.n_reference_filters, s->map);
  if (st) return st;
  s->tree = orc_kdtree_build(s->map->feat, s->map->n);
  return ORC_OK;
}
int orc_icp_seq_run(orc_icp_seq *s, const orc_cloud *reading, const double *T_init,
                    orc_icp_result *res) {
  memset(res, 0, sizeof(*res));
  if (!s->tree) return res->status = ORC_INVALID_FIELD;
  return icp_loop(&s->cfg, reading, s->map, s->tree, s->T_refIn_refMean, T_init, res);
}
const orc_cloud *orc_icp_seq_map(const orc_icp_seq *s) { return s->map; }

int orc_probe_overlap(const orc_icp_config *cfg, const orc_cloud *readingIn,
                      const orc_cloud *referenceIn, const double *T_world_robot,
                      double *weighted_ratio) {
  /* Localizer.hpp:309-347, module by module */
  orc_cloud *ref = orc_cloud_copy(referenceIn);
  int st = orc_filters_apply(cfg->reference_filters, cfg->n_reference_filters, ref);
  orc_cloud *rd = orc_cloud_copy(readingIn);
  if (!st) st = orc_filters_apply(cfg->reading_filters, cfg->n_reading_filters, rd);
  if (!st) st = orc_rigid_transform(rd, T_world_robot);
  if (!st) st = orc_filters_apply(cfg->reading_step_filters, cfg->n_reading_step_filters, rd);
  if (!st) {
    int k = cfg->knn;
    orc_kdtree *t = orc_kdtree_build(ref->feat, ref->n);
    int32_t *ids = (int32_t *)malloc((size_t)(rd->n + 1) * k * sizeof(int32_t));
    float *d2 = (float *)malloc((size_t)(rd->n + 1) * k * sizeof(float));
    float *w = (float *)malloc((size_t)(rd->n + 1) * k * sizeof(float));
    orc_kdtree_knn_ex(t, rd->feat, rd->n, k, (float)cfg->max_dist, 1, (float)cfg->epsilon, g_search_mode, ids, d2);
    st = orc_outlier_weights_full(cfg->outliers, cfg->n_outliers, rd, ref, ids, d2, k, w);
    if (!st) {
      int64_t kept = 0;
      double wsum = 0.0;
      for (int64_t m = 0; m < rd->n * k; ++m)
        if (!isinf(d2[m]) && w[m] != 0.f) { ++kept; wsum += (double)w[m]; }
      if (kept == 0) st = ORC_CONVERGENCE_ERROR;
      else *weighted_ratio = wsum / (double)(k * rd->n);
    }
    orc_kdtree_free(t);
    free(ids); free(d2); free(w);
  }
  orc_cloud_free(ref);
  orc_cloud_free(rd);
  return st;
}

int orc_probe_residual(const orc_icp_config *cfg, const orc_cloud *readingIn,
                       const orc_cloud *reference, const double *T, double *residual) {
  /* LoopCloser.hpp:346-362: raw (un-centred, unfiltered) candidate cloud   */
  orc_cloud *rd = orc_cloud_copy(readingIn);
  int st = orc_rigid_transform(rd, T);
  if (!st) {
    int k = cfg->knn;
    orc_kdtree *t = orc_kdtree_build(reference->feat, reference->n);
    int32_t *ids = (int32_t *)malloc((size_t)(rd->n + 1) * k * sizeof(int32_t));
    float *d2 = (float *)malloc((size_t)(rd->n + 1) * k * sizeof(float));
    float *w = (float *)malloc((size_t)(rd->n + 1) * k * sizeof(float));
    orc_kdtree_knn_ex(t, rd->feat, rd->n, k, (float)cfg->max_dist, 1, (float)cfg->epsilon, g_search_mode, ids, d2);
    st = orc_outlier_weights_full(cfg->outliers, cfg->n_outliers, rd, reference, ids, d2, k, w);
    if (!st) {
      orc_min_out mo;
      st = orc_minimize_ex(cfg->minimizer, cfg->force_mode, cfg->sensor_std_dev, rd, reference, ids, d2, w, k, &mo);
      if (!st) *residual = mo.residual;
    }
    orc_kdtree_free(t);
    free(ids); free(d2); free(w);
  }
  orc_cloud_free(rd);
  return st;
}
