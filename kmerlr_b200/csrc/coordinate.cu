// coordinate.cu -- the IRLS + coordinate-descent estimator of the reference
// (kmerLr_estimator_coordinate.go:31-139), for reduced matrices (<= 1023 columns).
// PARITY UNPINNED IN THE REFERENCE ITSELF: the function has no caller, no test and no golden there; the numpy
// restatement the tests compare this kernel with is pinned by optimality properties only.
//
//   outer iteration: r_i = x_i theta, p_i = sigma(r_i), w_i = p_i (1 - p_i), z_i = r_i + (y_i - p_i) / w_i,
//                    w_i *= class weight of y_i                                          (:97-111)
//   inner loop:      xy_j = sum_i w_i z_i x_ij,  xx_jk = sum_i w_i x_ij x_ik  (norm_j = xx_jj)   (:32-52)
//                    cyclic sweeps  t = xy_j + norm_j theta_j - sum_k xx_jk theta_k,
//                    soft threshold by L1Reg (j > 0), t / (norm_j + L2Reg)               (:55-72)
//   eval_stopping after every sweep and after every outer iteration, then the hook (:74-82, :117-125).
//
// The reference never calls this function and, as written, all of its theta slices alias one array
// (:89-91) so that it would stop after the first sweep; like estimate_proximal it is built here with
// the slices de-aliased (theta0 = start of the outer iteration, theta0_ = before the sweep).
//
// Device work: the row pass (one thread per row, the Gram matrix and xy as exact 64-bit fixed-point
// sums: deterministic, any order) and the sweeps (one block: the dot product of a Gram row with theta
// is a fixed tree over 1024 threads).  The host only sequences the launches and reads the stopping state.
#include <cmath>

#include "common.cuh"

namespace kl {

namespace {

constexpr int CD_MAX_THETA = 1024;
constexpr int CD_THREADS = 1024;

template <typename VT>
__device__ __forceinline__ double cd_val(const VT *val, int64_t p) { return val ? (double)val[p] : 1.0; }

// row pass: Gram and xy of the weighted least-squares problem of this outer iteration
template <typename VT>
__global__ void __launch_bounds__(256) cd_rows_kernel(const Rows R, const uint32_t *__restrict__ col,
                                                      const VT *__restrict__ val, int64_t n, int64_t ntheta,
                                                      const double *__restrict__ theta,
                                                      const uint8_t *__restrict__ labels, double cw0, double cw1,
                                                      double scale_xy, double scale_xx, long long *__restrict__ XY,
                                                      long long *__restrict__ XX, int *__restrict__ bad) {
  __shared__ double sth[CD_MAX_THETA];
  for (int i = threadIdx.x; i < ntheta; i += blockDim.x) sth[i] = theta[i];
  __syncthreads();
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  int64_t a, b;
  R.range(row, a, b);
  double s = 0.0;
  for (int64_t p = a; p < b; p++) s += cd_val(val, p) * sth[col[p] + 1];
  const double r = sth[0] + s;                                   // LinearPdf (:98)
  const double pr = exp(-log_add0(-r));                          // (:99)
  double w = pr * (1.0 - pr);                                    // (:100)
  const bool y = labels[row] != 0;
  const double z = y ? r + (1.0 - pr) / w : r + (0.0 - pr) / w;  // (:101-105)
  w *= y ? cw1 : cw0;                                            // (:106-110)
  const double wz = w * z;
  if (!isfinite(wz) || !isfinite(w)) { atomicExch(bad, 1); return; }   // the reference's sums turn NaN here
  // entry -1 = the bias column (index 0, value 1)
  for (int64_t p1 = a - 1; p1 < b; p1++) {
    const int64_t j1 = p1 < a ? 0 : (int64_t)col[p1] + 1;
    const double v1 = p1 < a ? 1.0 : cd_val(val, p1);
    atomicAdd((unsigned long long *)&XY[j1], (unsigned long long)__double2ll_rn(wz * v1 * scale_xy));
    const double wv1 = w * v1;
    for (int64_t p2 = a - 1; p2 < b; p2++) {
      const int64_t j2 = p2 < a ? 0 : (int64_t)col[p2] + 1;
      const double v2 = p2 < a ? 1.0 : cd_val(val, p2);
      atomicAdd((unsigned long long *)&XX[j1 * ntheta + j2], (unsigned long long)__double2ll_rn(wv1 * v2 * scale_xx));
    }
  }
}

__global__ void cd_convert_kernel(const long long *__restrict__ XY, const long long *__restrict__ XX, int64_t ntheta,
                                  double inv_xy, double inv_xx, double *__restrict__ xy, double *__restrict__ xx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ntheta) xy[i] = (double)XY[i] * inv_xy;
  if (i < ntheta * ntheta) xx[i] = (double)XX[i] * inv_xx;
}

struct CdState {
  double delta;     // eval_stopping's delta of the last sweep
  int stop;         // eval_stopping said stop
};

// one cyclic sweep over the coordinates (:57-72) + eval_stopping(theta0_, theta1) (:74), one block
__global__ void __launch_bounds__(CD_THREADS) cd_sweep_kernel(const double *__restrict__ xy, const double *__restrict__ xx,
                                                               int64_t ntheta, double l1, double l2, double eps,
                                                               double *theta, CdState *st) {
  __shared__ double sth[CD_MAX_THETA];
  __shared__ double red[CD_THREADS / 32];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  for (int i = t; i < ntheta; i += CD_THREADS) sth[i] = theta[i];
  __syncthreads();
  double max_x = 0.0, max_d = 0.0;
  bool isnan_seen = false;
  for (int64_t j = 0; j < ntheta; j++) {
    // sum_k xx[j][k] theta1[k]: one product per thread, fixed tree
    double v = t < ntheta ? xx[j * ntheta + t] * sth[t] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (t == 0) {
      double dot = red[0];
      for (int w = 1; w < CD_THREADS / 32; w++) dot += red[w];
      const double norm = xx[j * ntheta + j];
      double tj = xy[j] + norm * sth[j];
      tj -= dot;
      if (j > 0) {
        if (tj >= 0.0) tj =  fmax(fabs(tj) - l1, 0.0);
        else           tj = -fmax(fabs(tj) - l1, 0.0);
      }
      tj /= (norm + l2);
      const double old = sth[j];
      sth[j] = tj;
      if (isnan(tj)) isnan_seen = true;
      max_x = fmax(max_x, fabs(tj));
      max_d = fmax(max_d, fabs(tj - old));
    }
    __syncthreads();
  }
  for (int i = t; i < ntheta; i += CD_THREADS) theta[i] = sth[i];
  if (t == 0) {
    // eval_stopping (kmerLr_estimator_proximal.go:30-52)
    if (isnan_seen) { st->delta = nan(""); st->stop = 1; }
    else {
      st->delta = max_x != 0.0 ? max_d / max_x : max_d;
      st->stop = ((max_x != 0.0 && max_d / max_x <= eps) || (max_x == 0.0 && max_d == 0.0)) ? 1 : 0;
    }
  }
}

// eval_stopping on the host for the outer iteration (two short vectors)
bool eval_stopping_host(const std::vector<double> &x0, const std::vector<double> &x1, double eps, double *delta) {
  double max_x = 0.0, max_d = 0.0;
  for (size_t i = 0; i < x1.size(); i++) {
    if (std::isnan(x1[i])) { *delta = NAN; return true; }
    max_x = std::fmax(max_x, std::fabs(x1[i]));
    max_d = std::fmax(max_d, std::fabs(x1[i] - x0[i]));
  }
  *delta = max_x != 0.0 ? max_d / max_x : max_d;
  return (max_x != 0.0 && max_d / max_x <= eps) || (max_x == 0.0 && max_d == 0.0);
}

}  // namespace

void coordinate(Matrix &M, double *theta, int64_t ntheta, const double cw_hook[2], double l1reg, double l2reg,
                double epsilon, double epsilon_loss, int64_t max_iter, double hook[2], int64_t *sweeps_out,
                double *delta_out) {
  require_ready();
  KL_REQUIRE(ntheta == M.m + 1, "coordinate: theta must have one entry per column plus the bias");
  KL_REQUIRE(ntheta <= CD_MAX_THETA, "coordinate: the coordinate-descent estimator keeps a dense Gram matrix; use it on a "
                                     "reduced matrix (at most 1023 columns)");
  KL_REQUIRE(M.has_labels, "coordinate: the matrix has no labels (kmerlr_matrix_set_labels)");
  KL_REQUIRE(!M.sharded, "coordinate: single GPU only");
  KL_REQUIRE(M.n > 0, "coordinate: empty data set");
  // compute_class_weights(data_train.Labels) (:88; kmerLr_data.go:178-193)
  matrix_label_counts(M);
  KL_REQUIRE(M.n_pos > 0 && M.n_neg > 0, "coordinate: both classes must be present");
  const double ntot = (double)(M.n_pos + M.n_neg);
  const double cw[2] = {ntot / (2.0 * (double)M.n_neg), ntot / (2.0 * (double)M.n_pos)};
  const double cwmax = std::fmax(cw[0], cw[1]);
  const double vmax = std::fmax(matrix_vmax(M), 1.0), maxsq = matrix_maxsq(M);

  DevBuf<double> dtheta((size_t)ntheta), xy((size_t)ntheta), xx((size_t)(ntheta * ntheta));
  DevBuf<long long> XY((size_t)ntheta), XX((size_t)(ntheta * ntheta));
  DevBuf<int> bad(1);
  DevBuf<CdState> dst(1);
  std::vector<double> th1(theta, theta + ntheta), th0((size_t)ntheta);
  double hk_old = hook ? hook[0] : NAN, hk_new = hook ? hook[1] : NAN;
  // the hook (kmerLr_estimator_hook.go:46-99): loss at theta (the ESTIMATOR's class weights, lambda = L1Reg / n),
  // stop when it moved by less than epsilon_loss
  auto run_hook = [&](const std::vector<double> &th) -> bool {
    double t = hk_old; hk_old = hk_new; hk_new = t;
    if (epsilon_loss != 0.0) {
      hk_new = loss(M, th.data(), ntheta, cw_hook, l1reg / (double)M.n, 0);
      if (std::fabs(hk_old - hk_new) < epsilon_loss) return true;
    }
    return false;
  };
  int64_t sweeps = 0;
  double delta = 0.0;
  const unsigned row_grid = (unsigned)((M.n + 255) / 256);
  const unsigned cv_grid = (unsigned)((ntheta * ntheta + 255) / 256);
  bool nan_abort = false;
  for (int64_t iter = 0; iter < max_iter; iter++) {
    // scales of the fixed-point sums: |w x x| <= cw vmax^2 / 4, |w z x| <= cw vmax (|r| / 4 + 1)
    double th2 = 0.0;
    for (int64_t k = 1; k < ntheta; k++) th2 += th1[k] * th1[k];
    const double rmax = std::fabs(th1[0]) + std::sqrt(th2 * maxsq);
    double bxx = (double)M.n * 0.25 * cwmax * vmax * vmax, bxy = (double)M.n * cwmax * vmax * (0.25 * rmax + 1.0);
    if (!(bxx > 0.0) || !std::isfinite(bxx)) bxx = 1.0;
    if (!(bxy > 0.0) || !std::isfinite(bxy)) { nan_abort = true; break; }
    const int exx = 60 - (int)std::ceil(std::log2(bxx)), exy = 60 - (int)std::ceil(std::log2(bxy));
    dtheta.upload(th1.data(), (size_t)ntheta);
    XY.zero(); XX.zero(); bad.zero();
    if (M.vt == VAL_F64)
      KL_LAUNCH((cd_rows_kernel<double>), row_grid, 256, 0, M.rows(), M.col.p, M.val_f64.p, M.n, ntheta, dtheta.p, M.labels.p,
                cw[0], cw[1], std::ldexp(1.0, exy), std::ldexp(1.0, exx), XY.p, XX.p, bad.p);
    else
      KL_LAUNCH((cd_rows_kernel<uint32_t>), row_grid, 256, 0, M.rows(), M.col.p,
                M.vt == VAL_U32 ? M.val_u32.p : (const uint32_t *)nullptr, M.n, ntheta, dtheta.p, M.labels.p, cw[0], cw[1],
                std::ldexp(1.0, exy), std::ldexp(1.0, exx), XY.p, XX.p, bad.p);
    KL_LAUNCH(cd_convert_kernel, cv_grid, 256, 0, XY.p, XX.p, ntheta, std::ldexp(1.0, -exy), std::ldexp(1.0, -exx), xy.p,
              xx.p);
    int hbad = 0;
    bad.download(&hbad, 1);
    sync_stream();
    if (hbad) { nan_abort = true; break; }
    th0 = th1;                                                   // (:112-115)
    // estimate_coordinate_loop (:55-83): sweeps until eval_stopping or the hook says stop
    for (; iter < max_iter; iter++) {
      KL_LAUNCH(cd_sweep_kernel, 1, CD_THREADS, 0, xy.p, xx.p, ntheta, l1reg, l2reg, epsilon, dtheta.p, dst.p);
      CdState hs{};
      dst.download(&hs, 1);
      dtheta.download(th1.data(), (size_t)ntheta);
      sync_stream();
      sweeps++;
      delta = hs.delta;
      if (hs.stop) break;
      if (run_hook(th1)) break;
    }
    // (:117-125)
    if (eval_stopping_host(th0, th1, epsilon, &delta)) break;
    if (run_hook(th1)) break;
  }
  if (nan_abort) {
    for (auto &v : th1) v = NAN;
    delta = NAN;
  }
  for (int64_t k = 0; k < ntheta; k++) theta[k] = th1[k];
  if (hook) { hook[0] = hk_old; hook[1] = hk_new; }
  if (sweeps_out) *sweeps_out = sweeps;
  if (delta_out) *delta_out = delta;
}

}  // namespace kl
