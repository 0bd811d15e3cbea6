// select.cu -- featureSelector.Select (kmerLr_feature_selection.go:78-134,179-219) on top of the
// device gradient.  The full-space gradient is computed on the GPU (logistic.cu); the top-2N
// choice, the tie handling and lambda run on the host over the copied-back gradient
// (0.35 - 30 MB, once per leapfrog epoch).
//
// Two tie rules:
//   KMERLR_TIE_GO118  the order sort.Sort(sort.Reverse(AbsFloatInt)) produced with Go <= 1.18
//                     (kmerLr_sort.go:85-131): quickSort + ninther doPivot + shell pass +
//                     insertion sort + heapSort fallback, restated from the behavioural spec in
//                     SURVEY.md appendix A.  This is what the reference's goldens encode.
//   KMERLR_TIE_INDEX  |g| descending, coefficient index ascending.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace kl {

namespace {

// ---- Go <= 1.18 sort.Sort on (|value| descending) with the index riding along ---------------------
struct LegacySorter {
  double *a;
  int64_t *b;
  bool less(int64_t i, int64_t j) const { return std::fabs(a[j]) < std::fabs(a[i]); }  // sort.Reverse
  void swap(int64_t i, int64_t j) { std::swap(a[i], a[j]); std::swap(b[i], b[j]); }

  void insertion(int64_t lo, int64_t hi) {
    for (int64_t i = lo + 1; i < hi; i++)
      for (int64_t j = i; j > lo && less(j, j - 1); j--) swap(j, j - 1);
  }
  void sift_down(int64_t lo, int64_t hi, int64_t first) {
    int64_t root = lo;
    while (true) {
      int64_t child = 2 * root + 1;
      if (child >= hi) return;
      if (child + 1 < hi && less(first + child, first + child + 1)) child++;
      if (!less(first + root, first + child)) return;
      swap(first + root, first + child);
      root = child;
    }
  }
  void heap_sort(int64_t a0, int64_t b0) {
    int64_t first = a0, lo = 0, hi = b0 - a0;
    for (int64_t i = (hi - 1) / 2; i >= 0; i--) sift_down(i, hi, first);
    for (int64_t i = hi - 1; i >= 0; i--) { swap(first, first + i); sift_down(lo, i, first); }
  }
  void median3(int64_t m1, int64_t m0, int64_t m2) {
    if (less(m1, m0)) swap(m1, m0);
    if (less(m2, m1)) {
      swap(m2, m1);
      if (less(m1, m0)) swap(m1, m0);
    }
  }
  void pivot(int64_t lo, int64_t hi, int64_t &midlo, int64_t &midhi) {
    int64_t m = (int64_t)((uint64_t)(lo + hi) >> 1);
    if (hi - lo > 40) {
      int64_t s = (hi - lo) / 8;
      median3(lo, lo + s, lo + 2 * s);
      median3(m, m - s, m + s);
      median3(hi - 1, hi - 1 - s, hi - 1 - 2 * s);
    }
    median3(lo, m, hi - 1);
    int64_t pv = lo, x = lo + 1, c = hi - 1;
    while (x < c && less(x, pv)) x++;
    int64_t y = x;
    while (true) {
      while (y < c && !less(pv, y)) y++;
      while (y < c && less(pv, c - 1)) c--;
      if (y >= c) break;
      swap(y, c - 1);
      y++; c--;
    }
    bool protect = hi - c < 5;
    if (!protect && hi - c < (hi - lo) / 4) {
      int dups = 0;
      if (!less(pv, hi - 1)) { swap(c, hi - 1); c++; dups++; }
      if (!less(y - 1, pv)) { y--; dups++; }
      if (!less(m, pv)) { swap(m, y - 1); y--; dups++; }
      protect = dups > 1;
    }
    if (protect) {
      while (true) {
        while (x < y && !less(y - 1, pv)) y--;
        while (x < y && less(x, pv)) x++;
        if (x >= y) break;
        swap(x, y - 1);
        x++; y--;
      }
    }
    swap(pv, y - 1);
    midlo = y - 1; midhi = c;
  }
  void quick(int64_t lo, int64_t hi, int depth) {
    while (hi - lo > 12) {
      if (depth == 0) { heap_sort(lo, hi); return; }
      depth--;
      int64_t mlo, mhi;
      pivot(lo, hi, mlo, mhi);
      if (mlo - lo < hi - mhi) { quick(lo, mlo, depth); lo = mhi; }
      else { quick(mhi, hi, depth); hi = mlo; }
    }
    if (hi - lo > 1) {
      for (int64_t i = lo + 6; i < hi; i++)
        if (less(i, i - 6)) swap(i, i - 6);
      insertion(lo, hi);
    }
  }
  void sort(int64_t n) {
    int depth = 0;
    for (int64_t i = n; i > 0; i >>= 1) depth++;
    quick(0, n, 2 * depth);
  }
};

}  // namespace

// featureSelector.Select after its gradient call (kmerLr_feature_selection.go:88-134,179-192): b = mask,
// top 2N by |g| under the tie rule, lambda.  g is the full-space gradient at the embedded theta.
void select_from_gradient(const double *gin, int64_t ntheta, int64_t N, const int64_t *active_idx,
                          const double *active_theta, int64_t n_active, int tie, double eps_lambda, double prev_lambda,
                          uint8_t *b, double *lambda_out, int64_t *c_out, int *ok_out) {
  KL_REQUIRE(N >= 1, "select: N must be positive");
  KL_REQUIRE(tie == KMERLR_TIE_GO118 || tie == KMERLR_TIE_INDEX, "select: unknown tie rule");
  std::vector<double> g(gin, gin + ntheta);
  // alloc + restoreNonzero (:164-219): only coefficients with theta != 0 survive
  std::fill(b, b + ntheta, (uint8_t)0);
  b[0] = 1;
  int64_t c = 0;
  for (int64_t i = 0; i < n_active; i++) {
    KL_REQUIRE(active_idx[i] >= 1 && active_idx[i] < ntheta, "select: active coefficient index out of range");
    if (active_theta[i] != 0.0) { b[active_idx[i]] = 1; c++; }
  }
  const int64_t len = ntheta - 1;
  const int64_t top = len <= 2 * N ? len : 2 * N;
  std::vector<double> gs(g.begin() + 1, g.end());
  std::vector<int64_t> ix((size_t)len);
  std::iota(ix.begin(), ix.end(), (int64_t)0);
  if (tie == KMERLR_TIE_GO118) {
    LegacySorter s{gs.data(), ix.data()};
    s.sort(len);
  } else {
    // only the first 2N places are ever looked at (kmerLr_sort.go:126-130)
    auto cmp = [&](int64_t x, int64_t y) {
      double ax = std::fabs(g[(size_t)x + 1]), ay = std::fabs(g[(size_t)y + 1]);
      return ax > ay || (ax == ay && x < y);
    };
    std::partial_sort(ix.begin(), ix.begin() + top, ix.end(), cmp);
    for (int64_t k = 0; k < top; k++) gs[(size_t)k] = g[(size_t)ix[(size_t)k] + 1];
  }
  int ok = 0;
  for (int64_t k = 0; k < top; k++) {   // :95-105 new features
    if (c >= N) break;
    if (!b[ix[k] + 1] && gs[k] != 0.0) { ok = 1; b[ix[k] + 1] = 1; c++; }
  }
  for (int64_t k = 0; k < top; k++) {   // :107-116 old features
    if (c >= N) break;
    if (!b[ix[k] + 1]) { b[ix[k] + 1] = 1; c++; }
  }
  if (c > N) ok = 1;
  // computeLambda (:179-192): v = N-th largest |g|, w = largest |g| strictly below v
  double l = 0.0;
  if (N <= top) {
    double v = std::fabs(gs[(size_t)N - 1]), w = 0.0;
    for (int64_t k = 1; k < ntheta; k++) {
      double a = std::fabs(g[(size_t)k]);
      if (a > w && a < v) w = a;
    }
    l = (v + w) / 2.0;
  }
  *lambda_out = l; *c_out = c;
  *ok_out = ok || (eps_lambda > 0.0 && std::fabs(prev_lambda - l) >= eps_lambda);
}

void select(Matrix &M, const double cw[2], int cooc, int64_t N, double theta0, const int64_t *active_idx,
            const double *active_theta, int64_t n_active, int tie, double eps_lambda, double prev_lambda,
            uint8_t *b, int64_t ntheta, double *lambda_out, int64_t *c_out, int *ok_out, double *g_out) {
  require_ready();
  const int64_t dim = cooc ? kmerlr_coeff_dim(M.m) : M.m + 1;
  KL_INVARIANT(ntheta == dim);
  KL_REQUIRE(N >= 1, "select: N must be positive");
  // theta embedded in the full space (:164-219)
  std::vector<double> t((size_t)ntheta, 0.0), g((size_t)ntheta);
  t[0] = theta0;
  for (int64_t i = 0; i < n_active; i++) {
    KL_REQUIRE(active_idx[i] >= 1 && active_idx[i] < ntheta, "select: active coefficient index out of range");
    if (active_theta[i] != 0.0) t[active_idx[i]] = active_theta[i];
  }
  // gradient(data, t)[1:]  (:221-229): no penalty term
  gradient(M, t.data(), ntheta, cw, 0.0, cooc, g.data());
  if (g_out) std::copy(g.begin(), g.end(), g_out);
  select_from_gradient(g.data(), ntheta, N, active_idx, active_theta, n_active, tie, eps_lambda, prev_lambda, b,
                       lambda_out, c_out, ok_out);
}

}  // namespace kl
