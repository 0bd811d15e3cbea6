/*
 * kmerlr_oracle.c -- CPU ORACLE (test infrastructure, never shipped, never on the product path).
 * See kmerlr_oracle.h for the parity status.  Compile with -ffp-contract=off so that the
 * floating point follows Go on amd64 (no fused multiply-add).
 *
 * All file:line citations are into the reference tree (pbenner/kmerLr).  The k-mer semantics
 * of the un-vendored dependency github.com/pbenner/gonetics @40fc6f7ffc3c are restated from the
 * behavioural spec in SURVEY.md section 8c, which is pinned by kmerLr_test.go:40-43,55-66.
 */
#include "kmerlr_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ============================================================================================ */
/* k-mer classes                                                                                 */
/* ============================================================================================ */

#define KO_MAXK 24

static inline int alphabet_size(const ko_config *cfg) { return cfg->alphabet == 1 ? 5 : 4; }

/* letter codes a=0 c=1 g=2 t=3 (n=4 only as a gap wildcard); case-folded; anything else = 255 */
static inline uint8_t letter_code(uint8_t ch) {
  switch (ch) {
    case 'a': case 'A': return 0;
    case 'c': case 'C': return 1;
    case 'g': case 'G': return 2;
    case 't': case 'T': return 3;
    default: return 255;
  }
}
static inline uint8_t comp_code(uint8_t c) { return c < 4 ? (uint8_t)(3 - c) : c; }

static inline uint64_t code_of(const uint8_t *l, int k, int A) {
  uint64_t r = 0;
  for (int i = 0; i < k; i++) r = r * (uint64_t)A + l[i];
  return r;
}

/* class id = (k, min index over {kmer} U {complement} U {reverse} U {revcomp}) per enabled flag
 * (pinned for revcomp by kmerLr_test.go:40-43) */
uint64_t ko_class_code(const ko_config *cfg, const uint8_t *l, int32_t k) {
  int A = alphabet_size(cfg);
  uint8_t t[KO_MAXK];
  uint64_t best = code_of(l, k, A), c;
  if (cfg->complement) {
    for (int i = 0; i < k; i++) t[i] = comp_code(l[i]);
    c = code_of(t, k, A); if (c < best) best = c;
  }
  if (cfg->reverse) {
    for (int i = 0; i < k; i++) t[i] = l[k - 1 - i];
    c = code_of(t, k, A); if (c < best) best = c;
  }
  if (cfg->revcomp) {
    for (int i = 0; i < k; i++) t[i] = comp_code(l[k - 1 - i]);
    c = code_of(t, k, A); if (c < best) best = c;
  }
  return best;
}

static void decode(uint64_t code, int k, int A, uint8_t *l) {
  for (int i = k - 1; i >= 0; i--) { l[i] = (uint8_t)(code % (uint64_t)A); code /= (uint64_t)A; }
}

/* printed name: members joined by '|', smaller index first, duplicates kept ("gntanc|gntanc") */
int ko_class_name(const ko_config *cfg, int32_t k, uint64_t code, char *buf, int buflen) {
  static const char L[] = "acgtn";
  int A = alphabet_size(cfg);
  uint8_t l[KO_MAXK] = {0}, t[4][KO_MAXK];
  uint64_t c[4];
  int nm = 0;
  decode(code, k, A, l);
  memcpy(t[nm], l, (size_t)k); c[nm] = code; nm++;
  if (cfg->complement) { for (int i = 0; i < k; i++) t[nm][i] = comp_code(l[i]);         c[nm] = code_of(t[nm], k, A); nm++; }
  if (cfg->reverse)    { for (int i = 0; i < k; i++) t[nm][i] = l[k - 1 - i];            c[nm] = code_of(t[nm], k, A); nm++; }
  if (cfg->revcomp)    { for (int i = 0; i < k; i++) t[nm][i] = comp_code(l[k - 1 - i]); c[nm] = code_of(t[nm], k, A); nm++; }
  /* insertion sort by index */
  int ord[4] = {0, 1, 2, 3};
  for (int i = 1; i < nm; i++)
    for (int j = i; j > 0 && c[ord[j]] < c[ord[j - 1]]; j--) { int s = ord[j]; ord[j] = ord[j - 1]; ord[j - 1] = s; }
  int p = 0;
  for (int a = 0; a < nm; a++) {
    if (a > 0 && p < buflen - 1) buf[p++] = '|';
    for (int i = 0; i < k && p < buflen - 1; i++) buf[p++] = L[t[ord[a]][i]];
  }
  buf[p] = 0;
  return p;
}

/* ============================================================================================ */
/* per-sequence hash map  (stands in for gonetics KmerCounts.Counts map[KmerClassId]int)         */
/* ============================================================================================ */

typedef struct { uint64_t key; int32_t cnt; } kc_entry;
typedef struct { kc_entry *e; uint64_t cap, used; } kc_map;

#define KEY_EMPTY 0xFFFFFFFFFFFFFFFFull
static inline uint64_t mk_key(int k, uint64_t code) { return ((uint64_t)k << 56) | code; }
static inline uint64_t hash64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x;
}
static void kc_init(kc_map *m, uint64_t expect) {
  uint64_t cap = 64; while (cap < 2 * expect) cap <<= 1;
  m->cap = cap; m->used = 0; m->e = (kc_entry *)malloc(cap * sizeof(kc_entry));
  for (uint64_t i = 0; i < cap; i++) m->e[i].key = KEY_EMPTY;
}
static void kc_grow(kc_map *m);
static inline void kc_add(kc_map *m, uint64_t key, int binarize) {
  if (2 * (m->used + 1) > m->cap) kc_grow(m);
  uint64_t i = hash64(key) & (m->cap - 1);
  for (;;) {
    if (m->e[i].key == key) { if (!binarize) m->e[i].cnt++; return; }
    if (m->e[i].key == KEY_EMPTY) { m->e[i].key = key; m->e[i].cnt = 1; m->used++; return; }
    i = (i + 1) & (m->cap - 1);
  }
}
static void kc_grow(kc_map *m) {
  kc_map o = *m;
  m->cap = o.cap * 2; m->used = 0; m->e = (kc_entry *)malloc(m->cap * sizeof(kc_entry));
  for (uint64_t i = 0; i < m->cap; i++) m->e[i].key = KEY_EMPTY;
  for (uint64_t i = 0; i < o.cap; i++) if (o.e[i].key != KEY_EMPTY) {
    uint64_t j = hash64(o.e[i].key) & (m->cap - 1);
    while (m->e[j].key != KEY_EMPTY) j = (j + 1) & (m->cap - 1);
    m->e[j] = o.e[i]; m->used++;
  }
  free(o.e);
}
static inline int32_t kc_get(const kc_map *m, uint64_t key) {
  uint64_t i = hash64(key) & (m->cap - 1);
  for (;;) {
    if (m->e[i].key == key) return m->e[i].cnt;
    if (m->e[i].key == KEY_EMPTY) return 0;
    i = (i + 1) & (m->cap - 1);
  }
}
static int cmp_u64(const void *a, const void *b) {
  uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

/* CountKmers / IdentifyKmers of one sequence (scan_sequence, kmerLr_data.go:257-263; SURVEY 8c
 * items 1-6).  Every k in [M,N], every start with i+k <= len; gapped alphabet: every subset of the
 * interior positions replaced by n.  A k-mer touching a non-ACGT byte is skipped (unpinned). */
static void count_sequence(const ko_config *cfg, const uint8_t *s, int64_t len, kc_map *map) {
  int A = alphabet_size(cfg);
  uint8_t l[KO_MAXK], v[KO_MAXK];
  if (cfg->alphabet == 0) {
    /* incremental: extend the k-mer starting at i one base at a time */
    for (int64_t i = 0; i < len; i++) {
      uint64_t fw = 0, cp = 0, rv = 0, rc = 0, pw = 1;
      for (int k = 1; k <= cfg->N && i + k <= len; k++) {
        uint8_t b = letter_code(s[i + k - 1]);
        if (b > 3) break;
        fw = fw * 4 + b;            cp = cp * 4 + (uint64_t)(3 - b);
        rv = rv + (uint64_t)b * pw; rc = rc + (uint64_t)(3 - b) * pw;
        pw *= 4;
        if (k < cfg->M) continue;
        uint64_t best = fw;
        if (cfg->complement && cp < best) best = cp;
        if (cfg->reverse    && rv < best) best = rv;
        if (cfg->revcomp    && rc < best) best = rc;
        kc_add(map, mk_key(k, best), cfg->binarize);
      }
    }
    return;
  }
  for (int64_t i = 0; i < len; i++) {
    for (int k = 1; k <= cfg->N && i + k <= len; k++) {
      uint8_t b = letter_code(s[i + k - 1]);
      if (b > 3) break;
      l[k - 1] = b;
      if (k < cfg->M) continue;
      int nin = k >= 2 ? k - 2 : 0;
      for (uint32_t mask = 0; mask < (1u << nin); mask++) {
        if (cfg->max_ambiguous >= 0 && __builtin_popcount(mask) > cfg->max_ambiguous) continue;
        for (int j = 0; j < k; j++) v[j] = l[j];
        for (int j = 0; j < nin; j++) if (mask & (1u << j)) v[j + 1] = 4;
        kc_add(map, mk_key(k, ko_class_code(cfg, v, k)), cfg->binarize);
      }
    }
  }
  (void)A;
}

/* ============================================================================================ */
/* matrix                                                                                        */
/* ============================================================================================ */

struct ko_matrix {
  int64_t n, ncol, nnz, n_classes;
  int64_t *rowptr; int32_t *col; double *val;
  int32_t *class_k; uint64_t *class_code;
};

void ko_matrix_free(ko_matrix *m) {
  if (!m) return;
  free(m->rowptr); free(m->col); free(m->val); free(m->class_k); free(m->class_code); free(m);
}
void ko_matrix_info(const ko_matrix *m, int64_t *n, int64_t *ncol, int64_t *nnz, int64_t *n_classes) {
  if (n) *n = m->n; if (ncol) *ncol = m->ncol; if (nnz) *nnz = m->nnz; if (n_classes) *n_classes = m->n_classes;
}
void ko_matrix_classes(const ko_matrix *m, int32_t *k_out, uint64_t *code_out) {
  memcpy(k_out, m->class_k, (size_t)m->n_classes * sizeof(int32_t));
  memcpy(code_out, m->class_code, (size_t)m->n_classes * sizeof(uint64_t));
}
void ko_matrix_rows(const ko_matrix *m, int64_t *rowptr, int32_t *col, double *val) {
  memcpy(rowptr, m->rowptr, (size_t)(m->n + 1) * sizeof(int64_t));
  memcpy(col, m->col, (size_t)m->nnz * sizeof(int32_t));
  memcpy(val, m->val, (size_t)m->nnz * sizeof(double));
}
ko_matrix *ko_matrix_from_csr(int64_t n, int64_t ncol, const int64_t *rowptr, const int32_t *col, const double *val) {
  ko_matrix *m = (ko_matrix *)calloc(1, sizeof(ko_matrix));
  m->n = n; m->ncol = ncol; m->nnz = rowptr[n];
  m->rowptr = (int64_t *)malloc((size_t)(n + 1) * sizeof(int64_t));
  m->col = (int32_t *)malloc((size_t)(m->nnz ? m->nnz : 1) * sizeof(int32_t));
  m->val = (double *)malloc((size_t)(m->nnz ? m->nnz : 1) * sizeof(double));
  memcpy(m->rowptr, rowptr, (size_t)(n + 1) * sizeof(int64_t));
  memcpy(m->col, col, (size_t)m->nnz * sizeof(int32_t));
  memcpy(m->val, val, (size_t)m->nnz * sizeof(double));
  return m;
}

typedef struct { int32_t col; double val; } cv_pair;
static int cmp_cv(const void *a, const void *b) {
  int32_t x = ((const cv_pair *)a)->col, y = ((const cv_pair *)b)->col;
  return x < y ? -1 : (x > y ? 1 : 0);
}

/* compile_training_data (kmerLr_data.go:306-325): scan_sequences -> NewKmerCountsList (union of
 * observed classes sorted by (k, index)) -> SetKmers when frozen ->
 * convert_counts_list (kmerLr_data.go:197-253). */
ko_matrix *ko_extract(const ko_config *cfg, const uint8_t *seq, const int64_t *off, int64_t n,
                      const int32_t *frozen_k, const uint64_t *frozen_code, int64_t n_frozen,
                      const int32_t *features, int64_t n_features, int threads, int faithful) {
  if (cfg->N > KO_MAXK || cfg->M < 1 || cfg->M > cfg->N) return NULL;
  kc_map *maps = (kc_map *)calloc((size_t)(n ? n : 1), sizeof(kc_map));
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#else
  (void)threads;
#endif
  /* scan_sequences: one job per sequence (kmerLr_data.go:265-284) */
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < n; i++) {
    int64_t len = off[i + 1] - off[i];
    kc_init(&maps[i], (uint64_t)(len > 0 ? len : 1));
    count_sequence(cfg, seq + off[i], len, &maps[i]);
  }
  ko_matrix *m = (ko_matrix *)calloc(1, sizeof(ko_matrix));
  m->n = n;
  /* class list */
  uint64_t *ckeys = NULL; int64_t nc = 0;
  if (n_frozen > 0) {
    nc = n_frozen;
    ckeys = (uint64_t *)malloc((size_t)nc * sizeof(uint64_t));
    for (int64_t j = 0; j < nc; j++) ckeys[j] = mk_key(frozen_k[j], frozen_code[j]);
  } else {
    int64_t tot = 0;
    for (int64_t i = 0; i < n; i++) tot += (int64_t)maps[i].used;
    ckeys = (uint64_t *)malloc((size_t)(tot ? tot : 1) * sizeof(uint64_t));
    int64_t p = 0;
    for (int64_t i = 0; i < n; i++)
      for (uint64_t j = 0; j < maps[i].cap; j++) if (maps[i].e[j].key != KEY_EMPTY) ckeys[p++] = maps[i].e[j].key;
    qsort(ckeys, (size_t)tot, sizeof(uint64_t), cmp_u64);
    for (int64_t i = 0; i < tot; i++) if (i == 0 || ckeys[i] != ckeys[i - 1]) ckeys[nc++] = ckeys[i];
  }
  m->n_classes = nc;
  m->class_k = (int32_t *)malloc((size_t)(nc ? nc : 1) * sizeof(int32_t));
  m->class_code = (uint64_t *)malloc((size_t)(nc ? nc : 1) * sizeof(uint64_t));
  for (int64_t j = 0; j < nc; j++) { m->class_k[j] = (int32_t)(ckeys[j] >> 56); m->class_code[j] = ckeys[j] & ((1ull << 56) - 1); }
  m->ncol = n_features > 0 ? n_features : nc;
  /* class key -> column map for the non-faithful path */
  kc_map cmap; kc_init(&cmap, (uint64_t)(nc ? nc : 1));
  for (int64_t j = 0; j < nc; j++) {
    if (2 * (cmap.used + 1) > cmap.cap) kc_grow(&cmap);
    uint64_t i = hash64(ckeys[j]) & (cmap.cap - 1);
    while (cmap.e[i].key != KEY_EMPTY && cmap.e[i].key != ckeys[j]) i = (i + 1) & (cmap.cap - 1);
    if (cmap.e[i].key == KEY_EMPTY) { cmap.e[i].key = ckeys[j]; cmap.e[i].cnt = (int32_t)j; cmap.used++; }
  }
  /* convert_counts per sample */
  cv_pair **rows = (cv_pair **)calloc((size_t)(n ? n : 1), sizeof(cv_pair *));
  int64_t *rlen = (int64_t *)calloc((size_t)(n ? n : 1), sizeof(int64_t));
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < n; i++) {
    cv_pair *r; int64_t q = 0;
    if (n_features == 0) {
      if (faithful) {
        /* kmerLr_data.go:204-209: walk every class of the union list, look its count up */
        r = (cv_pair *)malloc((size_t)(maps[i].used ? maps[i].used : 1) * sizeof(cv_pair));
        for (int64_t j = 0; j < nc; j++) {
          int32_t c = kc_get(&maps[i], ckeys[j]);
          if (c != 0) { r[q].col = (int32_t)j; r[q].val = (double)c; q++; }
        }
      } else {
        r = (cv_pair *)malloc((size_t)(maps[i].used ? maps[i].used : 1) * sizeof(cv_pair));
        for (uint64_t j = 0; j < maps[i].cap; j++) if (maps[i].e[j].key != KEY_EMPTY) {
          uint64_t key = maps[i].e[j].key, h = hash64(key) & (cmap.cap - 1);
          while (cmap.e[h].key != KEY_EMPTY && cmap.e[h].key != key) h = (h + 1) & (cmap.cap - 1);
          if (cmap.e[h].key == key) { r[q].col = cmap.e[h].cnt; r[q].val = (double)maps[i].e[j].cnt; q++; }
        }
        qsort(r, (size_t)q, sizeof(cv_pair), cmp_cv);
      }
    } else {
      /* kmerLr_data.go:210-229: explicit feature list, singles and pair products */
      r = (cv_pair *)malloc((size_t)n_features * sizeof(cv_pair));
      for (int64_t j = 0; j < n_features; j++) {
        int32_t i1 = features[2 * j], i2 = features[2 * j + 1];
        if (i1 == i2) {
          int32_t c = kc_get(&maps[i], ckeys[i1]);
          if (c != 0) { r[q].col = (int32_t)j; r[q].val = (double)c; q++; }
        } else {
          int32_t c1 = kc_get(&maps[i], ckeys[i1]), c2 = kc_get(&maps[i], ckeys[i2]);
          if (c1 != 0 && c2 != 0) { r[q].col = (int32_t)j; r[q].val = (double)(c1 * c2); q++; }
        }
      }
    }
    rows[i] = r; rlen[i] = q;
    free(maps[i].e); maps[i].e = NULL;
  }
  m->rowptr = (int64_t *)malloc((size_t)(n + 1) * sizeof(int64_t));
  m->rowptr[0] = 0;
  for (int64_t i = 0; i < n; i++) m->rowptr[i + 1] = m->rowptr[i] + rlen[i];
  m->nnz = m->rowptr[n];
  m->col = (int32_t *)malloc((size_t)(m->nnz ? m->nnz : 1) * sizeof(int32_t));
  m->val = (double *)malloc((size_t)(m->nnz ? m->nnz : 1) * sizeof(double));
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    int64_t p = m->rowptr[i];
    for (int64_t j = 0; j < rlen[i]; j++) { m->col[p + j] = rows[i][j].col; m->val[p + j] = rows[i][j].val; }
    free(rows[i]);
  }
  free(rows); free(rlen); free(maps); free(ckeys); free(cmap.e);
  return m;
}

/* ============================================================================================ */
/* CoeffIndex  (kmerLr_coefficients_index.go:26-54)                                              */
/* ============================================================================================ */

int64_t ko_coeff_dim(int64_t n) { return (n + 1) * n / 2 + 1; }
int64_t ko_coeff_ind2sub(int64_t n, int64_t k1, int64_t k2) {
  if (k1 == k2) return k1 + 1;
  return n + (n * (n - 1) / 2) - (n - k1) * ((n - k1) - 1) / 2 + k2 - k1;
}
void ko_coeff_sub2ind(int64_t n, int64_t i, int64_t *k1, int64_t *k2) {
  if (i < n) { *k1 = i; *k2 = i; return; }
  i = i - n;
  int64_t a = n - 2 - (int64_t)floor(sqrt((double)(-8 * i + 4 * n * (n - 1) - 7)) / 2.0 - 0.5);
  int64_t b = i + a + 1 - n * (n - 1) / 2 + (n - a) * ((n - a) - 1) / 2;
  *k1 = a; *k2 = b;
}

/* ============================================================================================ */
/* logistic regression  (kmerLr_logistic_regression.go)                                          */
/* ============================================================================================ */

/* autodiff logarithmetic.LogAdd(a, b) = log(exp(a)+exp(b)); for a = 0 this is
 * max(0,x) + log1p(exp(-|x|))  (validated against kmerLr_test.go:224,228,248) */
static inline double log_add0(double x) {
  return x > 0.0 ? x + log1p(exp(-x)) : log1p(exp(x));
}

/* LinearPdf for one row (kmerLr_logistic_regression.go:47-85, Transform.Nil() branch).
 * Row arrays are without the bias; theta[0] is the bias, feature col j lives at theta[j+1]. */
static double linear_row(const ko_matrix *m, int64_t i, const double *theta, int cooc) {
  int64_t a = m->rowptr[i], b = m->rowptr[i + 1], q = b - a;
  const int32_t *c = m->col + a; const double *v = m->val + a;
  double r = theta[0];
  for (int64_t j = 0; j < q; j++) r += v[j] * theta[c[j] + 1];
  if (cooc) {
    /* :69-84  s[j1] += v[j1]*v[j2]*theta[Ind2Sub(i1,i2)]; then r += s[...] in order */
    int64_t n = m->ncol;
    for (int64_t j1 = 0; j1 < q; j1++) {
      double s = 0.0;
      for (int64_t j2 = j1 + 1; j2 < q; j2++) {
        int64_t j = ko_coeff_ind2sub(n, c[j1], c[j2]);
        s += v[j1] * v[j2] * theta[j];
      }
      r += s;
    }
  }
  return r;
}
void ko_linear_pdf(const ko_matrix *m, const double *theta, int cooc, double *out) {
  for (int64_t i = 0; i < m->n; i++) out[i] = linear_row(m, i, theta, cooc);
}
/* LogPdf = ClassLogPdf(x, true) = -LogAdd(0, -r)  (:138-149) */
void ko_log_pdf(const ko_matrix *m, const double *theta, int cooc, double *out) {
  for (int64_t i = 0; i < m->n; i++) out[i] = -log_add0(-linear_row(m, i, theta, cooc));
}

/* Gradient (:151-248): serial over samples, w = 1/n * cw[y] * (exp(logpdf) - y) */
void ko_gradient(const ko_matrix *m, const uint8_t *labels, const double *theta, int64_t ntheta,
                 const double cw[2], double lambda, int cooc, double *g) {
  for (int64_t j = 0; j < ntheta; j++) g[j] = 0.0;
  if (m->n == 0) return;
  int64_t nn = m->ncol;
  for (int64_t i = 0; i < m->n; i++) {
    double r = -log_add0(-linear_row(m, i, theta, cooc));
    double w;
    if (labels[i]) w = 1.0 / (double)m->n * cw[1] * (exp(r) - 1.0);
    else           w = 1.0 / (double)m->n * cw[0] * (exp(r));
    int64_t a = m->rowptr[i], b = m->rowptr[i + 1], q = b - a;
    const int32_t *c = m->col + a; const double *v = m->val + a;
    g[0] += w * 1.0;
    for (int64_t j = 0; j < q; j++) g[c[j] + 1] += w * v[j];
    if (cooc) {
      for (int64_t j1 = 0; j1 < q; j1++)
        for (int64_t j2 = j1 + 1; j2 < q; j2++) {
          int64_t j = ko_coeff_ind2sub(nn, c[j1], c[j2]);
          g[j] += w * v[j1] * v[j2];
        }
    }
  }
  if (!isnan(lambda) && lambda != 0.0) {
    for (int64_t j = 1; j < ntheta; j++) {
      if (theta[j] < 0) g[j] -= lambda; else if (theta[j] > 0) g[j] += lambda;
    }
  }
}

/* Loss (:250-272); the L1 loop bound is m = data[0].Dim() = ncol+1, also in pair mode */
double ko_loss(const ko_matrix *m, const uint8_t *labels, const double *theta, int64_t ntheta,
               const double cw[2], double lambda, int cooc) {
  (void)ntheta;
  if (m->n == 0) return 0.0;
  double r = 0.0;
  for (int64_t i = 0; i < m->n; i++) {
    double z = linear_row(m, i, theta, cooc);
    if (labels[i]) r -= cw[1] * (-log_add0(-z));
    else           r -= cw[0] * (-log_add0(z));
  }
  r = r / (double)m->n;
  if (!isnan(lambda) && lambda != 0.0)
    for (int64_t j = 1; j < m->ncol + 1; j++) r += lambda * fabs(theta[j]);
  return r;
}

/* compute_class_weights (kmerLr_data.go:178-193) */
void ko_class_weights(const uint8_t *labels, int64_t n, double cw[2]) {
  int64_t n1 = 0, n0 = 0;
  for (int64_t i = 0; i < n; i++) { if (labels[i]) n1++; else n0++; }
  cw[0] = (double)(n0 + n1) / (double)(2 * n0);
  cw[1] = (double)(n0 + n1) / (double)(2 * n1);
}

/* ============================================================================================ */
/* NLargestAbsFloat64 (kmerLr_sort.go:120-131) with Go <= 1.18 sort.Sort     */
/* ============================================================================================ */

typedef struct { double *a; int64_t *b; } afi;
/* Less under sort.Reverse(AbsFloatInt): |a[j]| < |a[i]| */
static inline int go_less(const afi *s, int64_t i, int64_t j) { return fabs(s->a[j]) < fabs(s->a[i]); }
static inline void go_swap(afi *s, int64_t i, int64_t j) {
  double t = s->a[i]; s->a[i] = s->a[j]; s->a[j] = t;
  int64_t u = s->b[i]; s->b[i] = s->b[j]; s->b[j] = u;
}
static void go_insertion(afi *s, int64_t a, int64_t b) {
  for (int64_t i = a + 1; i < b; i++)
    for (int64_t j = i; j > a && go_less(s, j, j - 1); j--) go_swap(s, j, j - 1);
}
static void go_sift(afi *s, int64_t lo, int64_t hi, int64_t first) {
  int64_t root = lo;
  for (;;) {
    int64_t child = 2 * root + 1;
    if (child >= hi) return;
    if (child + 1 < hi && go_less(s, first + child, first + child + 1)) child++;
    if (!go_less(s, first + root, first + child)) return;
    go_swap(s, first + root, first + child);
    root = child;
  }
}
static void go_heapsort(afi *s, int64_t a, int64_t b) {
  int64_t first = a, lo = 0, hi = b - a;
  for (int64_t i = (hi - 1) / 2; i >= 0; i--) go_sift(s, i, hi, first);
  for (int64_t i = hi - 1; i >= 0; i--) { go_swap(s, first, first + i); go_sift(s, lo, i, first); }
}
static void go_med3(afi *s, int64_t m1, int64_t m0, int64_t m2) {
  if (go_less(s, m1, m0)) go_swap(s, m1, m0);
  if (go_less(s, m2, m1)) { go_swap(s, m2, m1); if (go_less(s, m1, m0)) go_swap(s, m1, m0); }
}
static void go_pivot(afi *s, int64_t lo, int64_t hi, int64_t *midlo, int64_t *midhi) {
  int64_t m = (int64_t)((uint64_t)(lo + hi) >> 1);
  if (hi - lo > 40) {
    int64_t t = (hi - lo) / 8;
    go_med3(s, lo, lo + t, lo + 2 * t);
    go_med3(s, m, m - t, m + t);
    go_med3(s, hi - 1, hi - 1 - t, hi - 1 - 2 * t);
  }
  go_med3(s, lo, m, hi - 1);
  int64_t pivot = lo, a = lo + 1, c = hi - 1;
  for (; a < c && go_less(s, a, pivot); a++) {}
  int64_t b = a;
  for (;;) {
    for (; b < c && !go_less(s, pivot, b); b++) {}
    for (; b < c && go_less(s, pivot, c - 1); c--) {}
    if (b >= c) break;
    go_swap(s, b, c - 1); b++; c--;
  }
  int protect = hi - c < 5;
  if (!protect && hi - c < (hi - lo) / 4) {
    int dups = 0;
    if (!go_less(s, pivot, hi - 1)) { go_swap(s, c, hi - 1); c++; dups++; }
    if (!go_less(s, b - 1, pivot)) { b--; dups++; }
    if (!go_less(s, m, pivot)) { go_swap(s, m, b - 1); b--; dups++; }
    protect = dups > 1;
  }
  if (protect) {
    for (;;) {
      for (; a < b && !go_less(s, b - 1, pivot); b--) {}
      for (; a < b && go_less(s, a, pivot); a++) {}
      if (a >= b) break;
      go_swap(s, a, b - 1); a++; b--;
    }
  }
  go_swap(s, pivot, b - 1);
  *midlo = b - 1; *midhi = c;
}
static void go_quicksort(afi *s, int64_t a, int64_t b, int maxDepth) {
  while (b - a > 12) {
    if (maxDepth == 0) { go_heapsort(s, a, b); return; }
    maxDepth--;
    int64_t mlo, mhi;
    go_pivot(s, a, b, &mlo, &mhi);
    if (mlo - a < b - mhi) { go_quicksort(s, a, mlo, maxDepth); a = mhi; }
    else                   { go_quicksort(s, mhi, b, maxDepth); b = mlo; }
  }
  if (b - a > 1) {
    for (int64_t i = a + 6; i < b; i++) if (go_less(s, i, i - 6)) go_swap(s, i, i - 6);
    go_insertion(s, a, b);
  }
}

typedef struct { double v; int64_t i; } vi_pair;
static int cmp_vi(const void *a, const void *b) {
  const vi_pair *x = (const vi_pair *)a, *y = (const vi_pair *)b;
  double ax = fabs(x->v), ay = fabs(y->v);
  if (ax > ay) return -1; if (ax < ay) return 1;
  return x->i < y->i ? -1 : (x->i > y->i ? 1 : 0);
}

void ko_nlargest_abs(double *x, int64_t *idx, int64_t len, int tie) {
  for (int64_t j = 0; j < len; j++) idx[j] = j;
  if (tie == KO_TIE_GO118) {
    afi s = {x, idx};
    int depth = 0;
    for (int64_t i = len; i > 0; i >>= 1) depth++;
    go_quicksort(&s, 0, len, 2 * depth);
  } else {
    vi_pair *p = (vi_pair *)malloc((size_t)(len ? len : 1) * sizeof(vi_pair));
    for (int64_t j = 0; j < len; j++) { p[j].v = x[j]; p[j].i = j; }
    qsort(p, (size_t)len, sizeof(vi_pair), cmp_vi);
    for (int64_t j = 0; j < len; j++) { x[j] = p[j].v; idx[j] = p[j].i; }
    free(p);
  }
}

/* ============================================================================================ */
/* featureSelector.Select  (kmerLr_feature_selection.go:78-134, 179-219)                         */
/* ============================================================================================ */

int ko_select(const ko_matrix *m, const uint8_t *labels, const double cw[2], int cooc,
              int64_t N, double theta0, const int64_t *active_idx, const double *active_theta,
              int64_t n_active, int tie, double epsilon_lambda, double prev_lambda,
              uint8_t *b, int64_t ntheta, double *lambda_out, int64_t *c_out, double *g_out) {
  /* alloc + restoreNonzero (:164-219): only features with theta != 0 are carried over */
  double *t = (double *)calloc((size_t)ntheta, sizeof(double));
  memset(b, 0, (size_t)ntheta);
  b[0] = 1; t[0] = theta0;
  int64_t c = 0;
  for (int64_t i = 0; i < n_active; i++) {
    if (active_theta[i] != 0.0) { t[active_idx[i]] = active_theta[i]; b[active_idx[i]] = 1; c++; }
  }
  /* gradient(data, t)[1:]  (:221-229; lr.Lambda zero value -> no penalty term) */
  double *gfull = (double *)malloc((size_t)ntheta * sizeof(double));
  ko_gradient(m, labels, t, ntheta, cw, 0.0, cooc, gfull);
  if (g_out) memcpy(g_out, gfull, (size_t)ntheta * sizeof(double));
  int64_t len = ntheta - 1;
  double *gs = (double *)malloc((size_t)(len ? len : 1) * sizeof(double));
  int64_t *ix = (int64_t *)malloc((size_t)(len ? len : 1) * sizeof(int64_t));
  memcpy(gs, gfull + 1, (size_t)len * sizeof(double));
  ko_nlargest_abs(gs, ix, len, tie);
  int64_t top = len <= 2 * N ? len : 2 * N;
  int ok = 0;
  for (int64_t k = 0; k < top; k++) {         /* :95-105 add new features */
    if (c >= N) break;
    if (!b[ix[k] + 1] && gs[k] != 0.0) { ok = 1; b[ix[k] + 1] = 1; c++; }
  }
  for (int64_t k = 0; k < top; k++) {         /* :107-116 add old features */
    if (c >= N) break;
    if (!b[ix[k] + 1]) { b[ix[k] + 1] = 1; c++; }
  }
  if (c > N) ok = 1;
  /* computeLambda (:179-192) */
  double l = 0.0;
  if (N <= top) {
    double v = fabs(gs[N - 1]), w = 0.0;
    for (int64_t k = 0; k < len; k++) { double a = fabs(gs[k]); if (a > w && a < v) w = a; }
    l = (v + w) / 2.0;
  }
  *lambda_out = l; *c_out = c;
  free(t); free(gfull); free(gs); free(ix);
  return ok || (epsilon_lambda > 0.0 && fabs(prev_lambda - l) >= epsilon_lambda);
}

/* Float64At on a CSR row: binary search (indices are sorted) */
static double row_at(const ko_matrix *m, int64_t i, int64_t col) {
  int64_t lo = m->rowptr[i], hi = m->rowptr[i + 1];
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (m->col[mid] < col) lo = mid + 1; else hi = mid;
  }
  return (lo < m->rowptr[i + 1] && m->col[lo] == col) ? m->val[lo] : 0.0;
}

/* featureSelection.Data (kmerLr_feature_selection.go:309-343) */
ko_matrix *ko_reduce(const ko_matrix *m, const int64_t *sel, int64_t nsel) {
  ko_matrix *r = (ko_matrix *)calloc(1, sizeof(ko_matrix));
  r->n = m->n; r->ncol = nsel - 1;
  r->rowptr = (int64_t *)malloc((size_t)(m->n + 1) * sizeof(int64_t));
  int64_t cap = 1024, p = 0;
  r->col = (int32_t *)malloc((size_t)cap * sizeof(int32_t));
  r->val = (double *)malloc((size_t)cap * sizeof(double));
  r->rowptr[0] = 0;
  for (int64_t i = 0; i < m->n; i++) {
    for (int64_t j1 = 1; j1 < nsel; j1++) {
      int64_t j2 = sel[j1];
      double value;
      if (j2 >= m->ncol + 1) {
        int64_t i1, i2; ko_coeff_sub2ind(m->ncol, j2 - 1, &i1, &i2);
        value = row_at(m, i, i1) * row_at(m, i, i2);
      } else {
        value = row_at(m, i, j2 - 1);
      }
      if (value != 0.0) {
        if (p == cap) { cap *= 2; r->col = (int32_t *)realloc(r->col, (size_t)cap * sizeof(int32_t)); r->val = (double *)realloc(r->val, (size_t)cap * sizeof(double)); }
        r->col[p] = (int32_t)(j1 - 1); r->val[p] = value; p++;
      }
    }
    r->rowptr[i + 1] = p;
  }
  r->nnz = p;
  return r;
}

/* ============================================================================================ */
/* proximal gradient  (kmerLr_estimator_proximal.go:30-120)                                      */
/* ============================================================================================ */

/* estimate_step_size (:54-76); max_weight = 1 */
double ko_step_size(const ko_matrix *m, double l2, double step_factor) {
  double mx = 0.0;
  for (int64_t i = 0; i < m->n; i++) {
    double r = 0.0;
    for (int64_t p = m->rowptr[i]; p < m->rowptr[i + 1]; p++) r += m->val[p] * m->val[p];
    if (r > mx) mx = r;
  }
  double L = (0.25 * (mx + 1.0) + l2 / (double)m->n);
  L *= 1.0;
  double s = 1.0 / (2.0 * L + fmin(2.0 * l2, L));
  return s * step_factor;
}

/* eval_stopping (:30-52) */
static int eval_stopping(const double *xs, const double *x1, int64_t len, double eps, double *delta_out) {
  double max_x = 0.0, max_delta = 0.0, delta;
  for (int64_t i = 0; i < len; i++) {
    if (isnan(x1[i])) { *delta_out = NAN; return 1; }
    max_x = fmax(max_x, fabs(x1[i]));
    max_delta = fmax(max_delta, fabs(x1[i] - xs[i]));
  }
  delta = max_x != 0.0 ? max_delta / max_x : max_delta;
  *delta_out = delta;
  if ((max_x != 0.0 && max_delta / max_x <= eps) || (max_x == 0.0 && max_delta == 0.0)) return 1;
  return 0;
}

/* estimate_proximal (:78-120) with theta0/theta1 de-aliased (SURVEY section 0 finding 1) and
 * the soft threshold at s*lambda on the mean-loss scale (SURVEY section 7 hard part 1).
 * Hook = kmerLr_estimator_hook.go:46-99 with EvalLoss on when epsilon_loss != 0. */
int64_t ko_proxgrad(const ko_matrix *rm, const uint8_t *labels, double *theta, const double cw[2],
                    double lambda, double l2, double step_factor, double epsilon, double epsilon_loss,
                    int64_t max_iter, ko_hook_state *hook, double *delta_out) {
  int64_t len = rm->ncol + 1, it = 0;
  double s = ko_step_size(rm, l2, step_factor);
  double *theta0 = (double *)malloc((size_t)len * sizeof(double));
  double *g = (double *)malloc((size_t)len * sizeof(double));
  double delta = 0.0;
  ko_hook_state local = {NAN, NAN};
  if (!hook) hook = &local;
  for (int64_t i = 0; i < max_iter; i++) {
    ko_gradient(rm, labels, theta, len, cw, 0.0, 0, g);
    for (int64_t k = 0; k < len; k++) {
      theta0[k] = theta[k];
      theta[k] = theta[k] - s * g[k];
      if (k > 0) {
        if (theta[k] >= 0.0) theta[k] =  fmax(fabs(theta[k]) - s * lambda, 0.0);
        else                 theta[k] = -fmax(fabs(theta[k]) - s * lambda, 0.0);
      }
    }
    it = i + 1;
    if (eval_stopping(theta0, theta, len, epsilon, &delta)) break;
    /* hook */
    double t = hook->loss_old; hook->loss_old = hook->loss_new; hook->loss_new = t;
    if (epsilon_loss != 0.0) {
      hook->loss_new = ko_loss(rm, labels, theta, len, cw, lambda, 0);
      if (fabs(hook->loss_old - hook->loss_new) < epsilon_loss) break;
    }
  }
  if (delta_out) *delta_out = delta;
  free(theta0); free(g);
  return it;
}

/* estimate_loop (kmerLr_estimator.go:209-255) */
int64_t ko_estimate_loop(const ko_matrix *m, const uint8_t *labels, const double cw[2], int cooc,
                         int64_t N, int tie, double epsilon_lambda, double l2, double step_factor,
                         double epsilon, double epsilon_loss, int64_t max_iter, int64_t max_epochs,
                         ko_estimator *est, double *path_lambda, int64_t *path_iters, int64_t path_cap) {
  int64_t ntheta = cooc ? ko_coeff_dim(m->ncol) : m->ncol + 1;
  uint8_t *mask = (uint8_t *)malloc((size_t)ntheta);
  int have_r = 0;
  int64_t epoch = 0;
  for (; max_epochs == 0 || epoch < max_epochs; epoch++) {
    double lambda; int64_t c;
    /* Select(..., obj.L1Reg, ...): the reference passes L1Reg = lambda*n here (:234) */
    int ok = ko_select(m, labels, cw, cooc, N, est->theta[0], est->active_idx, est->theta + 1, est->n_active,
                       tie, epsilon_lambda, est->l1reg_over_n * (double)m->n, mask, ntheta, &lambda, &c, NULL);
    if (!ok && have_r) break;
    est->l1reg_over_n = lambda;
    /* selection.Features()/Theta(): all coefficient indices with b set, ascending */
    int64_t nsel = 0;
    for (int64_t j = 0; j < ntheta; j++) if (mask[j]) nsel++;
    if (nsel - 1 > est->state_cap) { free(mask); return -1; }
    int64_t *sel = (int64_t *)malloc((size_t)nsel * sizeof(int64_t));
    double *th = (double *)calloc((size_t)nsel, sizeof(double));
    int64_t p = 0;
    for (int64_t j = 0; j < ntheta; j++) if (mask[j]) sel[p++] = j;
    th[0] = est->theta[0];
    for (int64_t i = 0; i < est->n_active; i++) if (est->theta[1 + i] != 0.0) {
      /* position of active_idx[i] in sel */
      int64_t lo = 0, hi = nsel;
      while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (sel[mid] < est->active_idx[i]) lo = mid + 1; else hi = mid; }
      th[lo] = est->theta[1 + i];
    }
    ko_matrix *rm = ko_reduce(m, sel, nsel);
    int64_t iters = ko_proxgrad(rm, labels, th, cw, lambda, l2, step_factor, epsilon, epsilon_loss, max_iter, &est->hook, NULL);
    ko_matrix_free(rm);
    est->n_active = nsel - 1;
    for (int64_t j = 1; j < nsel; j++) est->active_idx[j - 1] = sel[j];
    memcpy(est->theta, th, (size_t)nsel * sizeof(double));
    if (epoch < path_cap) { if (path_lambda) path_lambda[epoch] = lambda; if (path_iters) path_iters[epoch] = iters; }
    free(sel); free(th);
    have_r = 1;
  }
  free(mask);
  return epoch;
}

/* ============================================================================================ */
/* genomic sliding-window scoring  (kmerLr_predict_genomic.go:134-171)                           */
/* ============================================================================================ */

int64_t ko_window_slots(int64_t len, int64_t W, int64_t step) {
  int64_t n = len - W;
  return n > 0 ? n / step + 1 : 0;
}

/* KmerLrEnsemble.Summarize (kmerLr_classifier_ensemble.go:64-105) */
static double summarize(int summary, const double *x, int64_t n) {
  double r;
  if (n == 0) return NAN;
  switch (summary) {
    case KO_SUMMARY_MEAN:    r = 0.0; for (int64_t j = 0; j < n; j++) r += x[j]; return r / (double)n;
    case KO_SUMMARY_PRODUCT: r = 1.0; for (int64_t j = 0; j < n; j++) r *= x[j]; return r;
    case KO_SUMMARY_MIN:     r = x[0]; for (int64_t j = 1; j < n; j++) if (r > x[j]) r = x[j]; return r;
    case KO_SUMMARY_MAX:     r = x[0]; for (int64_t j = 1; j < n; j++) if (r < x[j]) r = x[j]; return r;
    default:                 return n == 1 ? x[0] : NAN;
  }
}

/* genomicKmerLr.Predict (:134-143): scan_sequence -> SetKmers(model) -> convert_counts(features,
 * false) -> KmerLrEnsemble.Predict (kmerLr_classifier_ensemble.go:125-139), summed over models */
static double predict_window(const ko_model *models, int n_models, const uint8_t *s, int64_t W, double *tmp) {
  double r = 0.0;
  for (int mi = 0; mi < n_models; mi++) {
    const ko_model *md = &models[mi];
    kc_map map; kc_init(&map, (uint64_t)W);
    count_sequence(&md->cfg, s, W, &map);
    double *t = tmp;
    for (int64_t e = 0; e < md->n_members; e++) {
      const double *theta = md->theta + e * (md->n_features + 1);
      double z = theta[0];
      for (int64_t j = 0; j < md->n_features; j++) {
        int32_t i1 = md->features[2 * j], i2 = md->features[2 * j + 1];
        if (i1 == i2) {
          int32_t c = kc_get(&map, mk_key(md->class_k[i1], md->class_code[i1]));
          if (c != 0) z += (double)c * theta[j + 1];
        } else {
          int32_t c1 = kc_get(&map, mk_key(md->class_k[i1], md->class_code[i1]));
          int32_t c2 = kc_get(&map, mk_key(md->class_k[i2], md->class_code[i2]));
          if (c1 != 0 && c2 != 0) z += (double)(c1 * c2) * theta[j + 1];
        }
      }
      t[e] = -log_add0(-z);
    }
    r += summarize(md->summary, t, md->n_members);
    free(map.e);
  }
  return r;
}

/* predict_window_genomic (:147-171): loop j < len-W (strict), slots n/step+1 */
void ko_score_windows(const ko_model *models, int n_models, const uint8_t *seq, const int64_t *region_off,
                      int64_t n_regions, int64_t W, int64_t step, double *out, int threads) {
  int64_t *slot_off = (int64_t *)malloc((size_t)(n_regions + 1) * sizeof(int64_t));
  slot_off[0] = 0;
  int64_t maxmem = 1;
  for (int mi = 0; mi < n_models; mi++) if (models[mi].n_members > maxmem) maxmem = models[mi].n_members;
  for (int64_t i = 0; i < n_regions; i++)
    slot_off[i + 1] = slot_off[i] + ko_window_slots(region_off[i + 1] - region_off[i], W, step);
  for (int64_t i = 0; i < slot_off[n_regions]; i++) out[i] = 0.0;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#else
  (void)threads;
#endif
  for (int64_t i = 0; i < n_regions; i++) {
    int64_t len = region_off[i + 1] - region_off[i];
    int64_t nw = len - W > 0 ? (len - W + step - 1) / step : 0; /* j = 0, step, ... < len-W */
#pragma omp parallel
    {
      double *tmp = (double *)malloc((size_t)maxmem * sizeof(double));
#pragma omp for schedule(dynamic, 64)
      for (int64_t w = 0; w < nw; w++) {
        int64_t j = w * step;
        out[slot_off[i] + j / step] = predict_window(models, n_models, seq + region_off[i] + j, W, tmp);
      }
      free(tmp);
    }
  }
  free(slot_off);
}
