// select.cu -- featureSelector.Select (kmerLr_feature_selection.go:78-134,179-219) on top of the
// device gradient.  The full-space gradient is computed on the GPU (logistic.cu).  KMERLR_TIE_INDEX:
// the 2N largest |g| are found on the device (radix select + ordered compaction of the ties), only
// those 2N pairs go to the host.  KMERLR_TIE_GO118: the whole gradient is copied back (0.35 - 30 MB
// per leapfrog epoch) and the legacy sort runs on the host -- its result depends on the entire array.
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

// ---- KMERLR_TIE_INDEX on the device: the 2N largest |g| (ties: lower index first) ------------------------
// Radix select over the bit patterns of |g_k| (monotone for non-negative doubles), most significant digit
// first, 11 bits per pass: a pass is one histogram over the elements that still match the digits found so
// far and a one-block step that picks the bin holding rank K.  Then one ordered compaction: everything above
// the threshold, and of the elements equal to it the first ones in index order.  Only the 2N (index, value)
// pairs cross PCIe (the host keeps the rest of Select: it needs them sorted, the mask and lambda).
constexpr int SEL_BITS = 11, SEL_BINS = 1 << SEL_BITS, SEL_PASSES = 6;       // 6 x 11 >= 64
struct SelState {
  unsigned long long prefix;     // digits found so far (high bits of the threshold key)
  unsigned long long remaining;  // rank still to go inside the matching elements
  unsigned long long greater;    // elements strictly above the threshold
  unsigned long long cursor;     // output cursor of the "greater" elements
};
__device__ __forceinline__ unsigned long long abs_key(double g) { return (unsigned long long)__double_as_longlong(fabs(g)); }
__device__ __forceinline__ int sel_shift(int pass) { const int s = 64 - SEL_BITS * (pass + 1); return s < 0 ? 0 : s; }
__device__ __forceinline__ int sel_width(int pass) { return pass == SEL_PASSES - 1 ? 64 - SEL_BITS * (SEL_PASSES - 1) : SEL_BITS; }

__global__ void sel_hist(const double *__restrict__ g, int64_t len, int pass, const SelState *st, uint32_t *__restrict__ hist) {
  __shared__ uint32_t sh[SEL_BINS];
  for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const int shift = sel_shift(pass), width = sel_width(pass);
  const unsigned long long prefix = st->prefix;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = abs_key(g[i + 1]);
    // (shift + width == 64 in the first pass: no digits to match yet)
    if (pass == 0 || (key >> (shift + width)) == prefix) atomicAdd(&sh[(key >> shift) & ((1u << width) - 1u)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// one block: walk the bins from the top until the rank is reached; clears the histogram for the next pass
__global__ void sel_pick(int pass, SelState *st, uint32_t *__restrict__ hist) {
  __shared__ uint32_t sh[SEL_BINS];
  for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) { sh[i] = hist[i]; hist[i] = 0; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int width = sel_width(pass);
    unsigned long long rem = st->remaining, above = 0;
    int b = (1 << width) - 1;
    for (; b > 0; b--) {
      if (above + sh[b] >= rem) break;
      above += sh[b];
    }
    st->prefix = (pass == 0 ? 0ull : st->prefix << width) | (unsigned long long)b;
    st->remaining = rem - above;
    st->greater += above;
  }
}
// ordered compaction, three steps: per-block counts of the ties, their exclusive scan, the writes
__global__ void sel_count(const double *__restrict__ g, int64_t len, const SelState *st, uint32_t *__restrict__ blockties) {
  __shared__ uint32_t sh[256];
  const unsigned long long T = st->prefix;
  const int64_t per = (len + gridDim.x - 1) / gridDim.x, lo = (int64_t)blockIdx.x * per, hi = lo + per < len ? lo + per : len;
  uint32_t c = 0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) c += abs_key(g[i + 1]) == T ? 1u : 0u;
  sh[threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) blockties[blockIdx.x] = sh[0];
}
__global__ void sel_scan(uint32_t *__restrict__ blockties, int nblocks) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    uint32_t run = 0;
    for (int b = 0; b < nblocks; b++) { const uint32_t c = blockties[b]; blockties[b] = run; run += c; }
  }
}
__global__ void sel_write(const double *__restrict__ g, int64_t len, SelState *st, const uint32_t *__restrict__ blockties,
                          int64_t K, int64_t *__restrict__ out_idx, double *__restrict__ out_val) {
  __shared__ uint32_t sh[256];
  __shared__ uint32_t s_base;
  const unsigned long long T = st->prefix;
  const unsigned long long greater = st->greater;          // < K
  const int64_t per = (len + gridDim.x - 1) / gridDim.x, lo = (int64_t)blockIdx.x * per, hi = lo + per < len ? lo + per : len;
  if (threadIdx.x == 0) s_base = blockties[blockIdx.x];
  __syncthreads();
  for (int64_t i0 = lo; i0 < hi; i0 += blockDim.x) {
    const int64_t i = i0 + threadIdx.x;
    unsigned long long key = 0;
    double v = 0.0;
    bool tie = false;
    if (i < hi) {
      v = g[i + 1]; key = abs_key(v);
      if (key > T) {
        const unsigned long long p = atomicAdd(&st->cursor, 1ull);
        out_idx[p] = i; out_val[p] = v;
      }
      tie = key == T;
    }
    // ties in index order: block scan of the flags (Hillis-Steele over 256 threads)
    sh[threadIdx.x] = tie ? 1u : 0u;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
      const uint32_t y = (int)threadIdx.x >= o ? sh[threadIdx.x - o] : 0u;
      __syncthreads();
      sh[threadIdx.x] += y;
      __syncthreads();
    }
    const uint32_t rank = s_base + sh[threadIdx.x] - (tie ? 1u : 0u);
    if (tie && greater + rank < (unsigned long long)K) { out_idx[greater + rank] = i; out_val[greater + rank] = v; }
    __syncthreads();
    if (threadIdx.x == 255) s_base += sh[255];
    __syncthreads();
  }
}
// largest |g_k| strictly below v (computeLambda, kmerLr_feature_selection.go:186-190), as a key
__global__ void sel_max_below(const double *__restrict__ g, int64_t len, unsigned long long vkey, unsigned long long *__restrict__ out) {
  unsigned long long m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = abs_key(g[i + 1]);
    if (key < vkey && key > m) m = key;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long y = __shfl_xor_sync(0xffffffffu, m, o); m = y > m ? y : m; }
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

}  // namespace

// the `top` largest |g_k|, k >= 1, of a gradient resident on the device, ties by lower index: (k - 1, g_k) pairs in
// no particular order
static void device_top(const double *g, int64_t ntheta, int64_t top, std::vector<int64_t> &idx, std::vector<double> &val) {
  const int64_t len = ntheta - 1;
  idx.assign((size_t)top, 0); val.assign((size_t)top, 0.0);
  if (top <= 0) return;
  DevBuf<SelState> st(1);
  DevBuf<uint32_t> hist(SEL_BINS);
  SelState h{};
  h.remaining = (unsigned long long)top;
  st.upload(&h, 1);
  hist.zero();
  int blocks = ctx().sm_count * 4;
  if ((int64_t)blocks * 256 > len) blocks = (int)((len + 255) / 256);
  for (int pass = 0; pass < SEL_PASSES; pass++) {
    KL_LAUNCH(sel_hist, (unsigned)blocks, 256, 0, g, len, pass, st.p, hist.p);
    KL_LAUNCH(sel_pick, 1, 256, 0, pass, st.p, hist.p);
  }
  DevBuf<uint32_t> blockties((size_t)blocks);
  DevBuf<int64_t> oi((size_t)top);
  DevBuf<double> ov((size_t)top);
  KL_LAUNCH(sel_count, (unsigned)blocks, 256, 0, g, len, st.p, blockties.p);
  KL_LAUNCH(sel_scan, 1, 32, 0, blockties.p, blocks);
  KL_LAUNCH(sel_write, (unsigned)blocks, 256, 0, g, len, st.p, blockties.p, top, oi.p, ov.p);
  oi.download(idx.data(), (size_t)top);
  ov.download(val.data(), (size_t)top);
  sync_stream();
}

static double device_max_below(const double *g, int64_t ntheta, double v) {
  DevBuf<unsigned long long> out(1);
  out.zero();
  unsigned long long vkey;
  const double av = std::fabs(v);
  memcpy(&vkey, &av, sizeof(vkey));
  KL_LAUNCH(sel_max_below, (unsigned)(ctx().sm_count * 4), 256, 0, g, ntheta - 1, vkey, out.p);
  unsigned long long m = 0;
  out.download(&m, 1);
  sync_stream();
  double w;
  memcpy(&w, &m, sizeof(w));
  return w;
}

// featureSelector.Select after its gradient call (kmerLr_feature_selection.go:88-134,179-192): b = mask,
// top 2N by |g| under the tie rule, lambda.  g is the full-space gradient at the embedded theta.
// the part of Select that only looks at the first `top` places of the sorted gradient: ix[k] = coefficient index - 1,
// gs[k] = its gradient, in sorted order; w_below(v) = the largest |g| strictly below v over the whole gradient
template <typename WBelow>
static void select_core(int64_t ntheta, int64_t N, const int64_t *active_idx, const double *active_theta, int64_t n_active,
                        double eps_lambda, double prev_lambda, uint8_t *b, const std::vector<int64_t> &ix,
                        const std::vector<double> &gs, int64_t top, WBelow &&w_below, double *lambda_out, int64_t *c_out,
                        int *ok_out) {
  // alloc + restoreNonzero (:164-219): only coefficients with theta != 0 survive
  std::fill(b, b + ntheta, (uint8_t)0);
  b[0] = 1;
  int64_t c = 0;
  for (int64_t i = 0; i < n_active; i++) {
    KL_REQUIRE(active_idx[i] >= 1 && active_idx[i] < ntheta, "select: active coefficient index out of range");
    if (active_theta[i] != 0.0) { b[active_idx[i]] = 1; c++; }
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
    const double v = std::fabs(gs[(size_t)N - 1]);
    l = (v + w_below(v)) / 2.0;
  }
  *lambda_out = l; *c_out = c;
  *ok_out = ok || (eps_lambda > 0.0 && std::fabs(prev_lambda - l) >= eps_lambda);
}

void select_from_gradient(const double *gin, int64_t ntheta, int64_t N, const int64_t *active_idx,
                          const double *active_theta, int64_t n_active, int tie, double eps_lambda, double prev_lambda,
                          uint8_t *b, double *lambda_out, int64_t *c_out, int *ok_out) {
  KL_REQUIRE(N >= 1, "select: N must be positive");
  KL_REQUIRE(tie == KMERLR_TIE_GO118 || tie == KMERLR_TIE_INDEX, "select: unknown tie rule");
  std::vector<double> g(gin, gin + ntheta);
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
  auto w_below = [&](double v) {
    double w = 0.0;
    for (int64_t k = 1; k < ntheta; k++) {
      double a = std::fabs(g[(size_t)k]);
      if (a > w && a < v) w = a;
    }
    return w;
  };
  select_core(ntheta, N, active_idx, active_theta, n_active, eps_lambda, prev_lambda, b, ix, gs, top, w_below, lambda_out,
              c_out, ok_out);
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
  if (tie == KMERLR_TIE_INDEX && !g_out) {
    // the gradient stays on the device: radix select of the 2N largest |g| (ties: lower index first)
    DevBuf<double> gd;
    gradient(M, t.data(), ntheta, cw, 0.0, cooc, nullptr, &gd);
    const int64_t len = ntheta - 1, top = len <= 2 * N ? len : 2 * N;
    std::vector<int64_t> ix;
    std::vector<double> gs;
    device_top(gd.p, ntheta, top, ix, gs);
    std::vector<int64_t> order((size_t)top);
    std::iota(order.begin(), order.end(), (int64_t)0);
    std::sort(order.begin(), order.end(), [&](int64_t x, int64_t y) {
      const double ax = std::fabs(gs[(size_t)x]), ay = std::fabs(gs[(size_t)y]);
      return ax > ay || (ax == ay && ix[(size_t)x] < ix[(size_t)y]);
    });
    std::vector<int64_t> six((size_t)top);
    std::vector<double> sgs((size_t)top);
    for (int64_t k = 0; k < top; k++) { six[(size_t)k] = ix[(size_t)order[(size_t)k]]; sgs[(size_t)k] = gs[(size_t)order[(size_t)k]]; }
    auto w_below = [&](double v) {
      // the largest |g| below v is among the top places unless the ties at v fill them all
      for (int64_t k = 0; k < top; k++)
        if (std::fabs(sgs[(size_t)k]) < v) return std::fabs(sgs[(size_t)k]);
      return top < len ? device_max_below(gd.p, ntheta, v) : 0.0;
    };
    select_core(ntheta, N, active_idx, active_theta, n_active, eps_lambda, prev_lambda, b, six, sgs, top, w_below,
                lambda_out, c_out, ok_out);
    return;
  }
  gradient(M, t.data(), ntheta, cw, 0.0, cooc, g.data());
  if (g_out) std::copy(g.begin(), g.end(), g_out);
  select_from_gradient(g.data(), ntheta, N, active_idx, active_theta, n_active, tie, eps_lambda, prev_lambda, b,
                       lambda_out, c_out, ok_out);
}

}  // namespace kl
