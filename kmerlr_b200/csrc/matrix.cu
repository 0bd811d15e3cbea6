// matrix.cu -- KmerDataSet in HBM: CSR import/export, labels, the stable CSR -> CSC transpose that
// makes X^T w deterministic, featureSelection.Data (reduced matrix) and the step-size row norm.
#include "common.cuh"

#include <cmath>

namespace kl {

namespace {

// ---- stable transpose ---------------------------------------------------------------------------
// Rows are split into B contiguous blocks.  cnt[b][c] = entries of column c in block b (integer
// atomics: order independent).  A scan over b gives every block its slice of every column, and each
// block then fills its slices walking its rows in ascending order, so the entries of a column end
// up in ascending row order -- the order the reference accumulates a gradient entry in
// (kmerLr_logistic_regression.go:166-181) -- independent of scheduling.
__global__ void csc_count(const int64_t *__restrict__ rowptr, const uint32_t *__restrict__ col, int64_t n,
                          int64_t rows_per_block, int64_t m, uint32_t *__restrict__ cnt) {
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  unsigned lane = lane_id();
  uint32_t *c = cnt + (row / rows_per_block) * m;
  for (int64_t p = rowptr[row] + lane; p < rowptr[row + 1]; p += 32) atomicAdd(c + col[p], 1u);
}

__global__ void csc_block_scan(uint32_t *__restrict__ cnt, int64_t B, int64_t m, uint32_t *__restrict__ colcnt) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  uint32_t run = 0;
  for (int64_t b = 0; b < B; b++) {
    uint32_t t = cnt[b * m + c];
    cnt[b * m + c] = run;
    run += t;
  }
  colcnt[c] = run;
}

template <typename VT>
__global__ void csc_fill(const int64_t *__restrict__ rowptr, const uint32_t *__restrict__ col,
                         const VT *__restrict__ val, int64_t n, int64_t rows_per_block, int64_t m,
                         uint32_t *__restrict__ cnt, const int64_t *__restrict__ colptr,
                         uint32_t *__restrict__ crow, VT *__restrict__ cval) {
  int64_t b = blockIdx.x;
  int64_t r0 = b * rows_per_block, r1 = r0 + rows_per_block;
  if (r1 > n) r1 = n;
  uint32_t *cur = cnt + b * m;
  for (int64_t row = r0; row < r1; row++) {
    // columns are distinct inside a row: no two threads touch the same cursor between barriers
    for (int64_t p = rowptr[row] + threadIdx.x; p < rowptr[row + 1]; p += blockDim.x) {
      uint32_t c = col[p];
      int64_t pos = colptr[c] + cur[c];
      cur[c] = cur[c] + 1;
      crow[pos] = (uint32_t)row;
      if (val) cval[pos] = val[p];
    }
    __syncthreads();
  }
}

constexpr int TASK_CHUNK = 1024;   // CSC entries per warp task of the X^T w reduction

__global__ void task_counts(const int64_t *__restrict__ colptr, int64_t m, uint32_t *__restrict__ nt) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < m) nt[c] = (uint32_t)((colptr[c + 1] - colptr[c] + TASK_CHUNK - 1) / TASK_CHUNK);
}
__global__ void task_fill(const int64_t *__restrict__ taskptr, int64_t m, uint32_t *__restrict__ taskcol) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  for (int64_t t = taskptr[c]; t < taskptr[c + 1]; t++) taskcol[t] = (uint32_t)c;
}

// ---- featureSelection.Data (kmerLr_feature_selection.go:309-343) ----------------------------------
template <typename VT>
__device__ __forceinline__ double row_value(const uint32_t *col, const VT *val, int64_t a, int64_t b, uint32_t c) {
  int64_t lo = a, hi = b;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (col[mid] < c) lo = mid + 1; else hi = mid;
  }
  if (lo < b && col[lo] == c) return val ? (double)val[lo] : 1.0;
  return 0.0;
}

template <typename VT, bool WRITE>
__global__ void reduce_rows(const int64_t *__restrict__ rowptr, const uint32_t *__restrict__ col,
                            const VT *__restrict__ val, int64_t n, const int64_t *__restrict__ selA,
                            const int64_t *__restrict__ selB, int64_t nsel, uint32_t *__restrict__ cnt_out,
                            const int64_t *__restrict__ orowptr, uint32_t *__restrict__ ocol,
                            double *__restrict__ oval) {
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  unsigned lane = lane_id();
  int64_t a = rowptr[row], b = rowptr[row + 1];
  int64_t outp = WRITE ? orowptr[row] : 0;
  uint32_t tot = 0;
  for (int64_t j0 = 0; j0 < nsel; j0 += 32) {
    int64_t j = j0 + lane;
    double v = 0.0;
    if (j < nsel) {
      v = row_value(col, val, a, b, (uint32_t)selA[j]);
      if (selB[j] >= 0 && v != 0.0) v = v * row_value(col, val, a, b, (uint32_t)selB[j]);
    }
    unsigned km = __ballot_sync(0xffffffffu, v != 0.0);
    if (WRITE && v != 0.0) {
      int64_t pos = outp + __popc(km & lanemask_lt());
      ocol[pos] = (uint32_t)j;
      oval[pos] = v;
    }
    outp += __popc(km);
    tot += __popc(km);
  }
  if (!WRITE && lane == 0) cnt_out[row] = tot;
}

template <typename VT>
__global__ void row_sqnorm_max(const int64_t *__restrict__ rowptr, const VT *__restrict__ val, int64_t n,
                               unsigned long long *__restrict__ out) {
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  unsigned lane = lane_id();
  double s = 0.0;
  if (val) {
    for (int64_t p = rowptr[row] + lane; p < rowptr[row + 1]; p += 32) { double v = (double)val[p]; s += v * v; }
    s = warp_sum(s);
  } else {
    s = (double)(rowptr[row + 1] - rowptr[row]);
  }
  // non-negative doubles compare like their bit patterns
  if (lane == 0) atomicMax(out, (unsigned long long)__double_as_longlong(s));
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
std::shared_ptr<Matrix> matrix_from_csr(int64_t n, int64_t m, const int64_t *rowptr, const int32_t *col,
                                        const double *val, int flags) {
  require_ready();
  KL_REQUIRE(n >= 0 && m >= 0 && rowptr && rowptr[0] == 0, "from_csr: bad arguments");
  int64_t nnz = rowptr[n];
  for (int64_t i = 0; i < n; i++) {
    KL_REQUIRE(rowptr[i + 1] >= rowptr[i], "from_csr: rowptr must be non-decreasing");
    for (int64_t p = rowptr[i]; p < rowptr[i + 1]; p++) {
      KL_REQUIRE(col[p] >= 0 && col[p] < m, "from_csr: column out of range");
      KL_REQUIRE(p == rowptr[i] || col[p] > col[p - 1], "from_csr: columns must be strictly increasing in a row");
    }
  }
  auto M = std::make_shared<Matrix>();
  M->n = n; M->m = m; M->nnz = nnz; M->vt = VAL_F64;
  M->sharded = (flags & KMERLR_FLAG_SHARDED) != 0 && ctx().world > 1;
  M->n_global = n;
  M->rowptr.alloc((size_t)n + 1);
  M->col.alloc((size_t)(nnz ? nnz : 1));
  M->val_f64.alloc((size_t)(nnz ? nnz : 1));
  M->rowptr.upload(rowptr, (size_t)n + 1);
  std::vector<uint32_t> c32((size_t)nnz);
  for (int64_t p = 0; p < nnz; p++) c32[p] = (uint32_t)col[p];
  M->col.upload(c32.data(), (size_t)nnz);
  M->val_f64.upload(val, (size_t)nnz);
  if (M->sharded) {
    DevBuf<int64_t> tmp(1);
    int64_t nn = n;
    tmp.upload(&nn, 1);
    comm_allreduce_sum_i64(tmp.p, 1);
    tmp.download(&nn, 1);
    sync_stream();
    M->n_global = nn;
  }
  sync_stream();
  return M;
}

void matrix_rows(const Matrix &M, int64_t *rowptr, int32_t *col, double *val) {
  require_ready();
  M.rowptr.download(rowptr, (size_t)M.n + 1);
  std::vector<uint32_t> c32((size_t)M.nnz);
  M.col.download(c32.data(), (size_t)M.nnz);
  std::vector<uint32_t> v32;
  if (M.vt == VAL_U32) { v32.resize((size_t)M.nnz); M.val_u32.download(v32.data(), (size_t)M.nnz); }
  else if (M.vt == VAL_F64) M.val_f64.download(val, (size_t)M.nnz);
  sync_stream();
  for (int64_t p = 0; p < M.nnz; p++) {
    col[p] = (int32_t)c32[p];
    if (M.vt == VAL_U32) val[p] = (double)v32[p];
    else if (M.vt == VAL_ONE) val[p] = 1.0;
  }
}

void matrix_set_labels(Matrix &M, const uint8_t *labels, int64_t n) {
  require_ready();
  KL_REQUIRE(n == M.n, "labels: length does not match the number of rows");
  M.labels.alloc((size_t)(n ? n : 1));
  std::vector<uint8_t> l((size_t)n);
  int64_t cnt[2] = {0, 0};
  for (int64_t i = 0; i < n; i++) { l[i] = labels[i] ? 1 : 0; cnt[l[i]]++; }
  M.labels.upload(l.data(), (size_t)n);
  if (M.sharded) {
    DevBuf<int64_t> tmp(2);
    tmp.upload(cnt, 2);
    comm_allreduce_sum_i64(tmp.p, 2);
    tmp.download(cnt, 2);
  }
  sync_stream();
  M.n_neg = cnt[0]; M.n_pos = cnt[1];
  M.has_labels = true;
}

template <typename VT>
static void build_csc(Matrix &M, const VT *val, VT *cval) {
  // number of row blocks: bounded by the 1 GiB budget of the count table
  int64_t B = (int64_t)ctx().sm_count * 4;
  int64_t cap = ((int64_t)1 << 28) / (M.m > 0 ? M.m : 1);
  if (B > cap) B = cap;
  if (B > M.n) B = M.n;
  if (B < 1) B = 1;
  int64_t rpb = (M.n + B - 1) / B;
  B = (M.n + rpb - 1) / rpb;
  DevBuf<uint32_t> cnt((size_t)(B * M.m)), colcnt((size_t)M.m);
  cnt.zero();
  unsigned wgrid = (unsigned)((M.n * 32 + 255) / 256);
  KL_LAUNCH(csc_count, wgrid, 256, 0, M.rowptr.p, M.col.p, M.n, rpb, M.m, cnt.p);
  KL_LAUNCH(csc_block_scan, (unsigned)((M.m + 255) / 256), 256, 0, cnt.p, B, M.m, colcnt.p);
  exclusive_scan_u32_to_i64(colcnt.p, M.colptr.p, M.m);
  KL_LAUNCH((csc_fill<VT>), (unsigned)B, 256, 0, M.rowptr.p, M.col.p, val, M.n, rpb, M.m, cnt.p, M.colptr.p,
            M.crow.p, cval);
  sync_stream();
}

void ensure_csc(Matrix &M) {
  if (M.has_csc) return;
  M.colptr.alloc((size_t)M.m + 1);
  M.crow.alloc((size_t)(M.nnz ? M.nnz : 1));
  if (M.m == 0 || M.n == 0) {
    M.colptr.zero();
  } else if (M.vt == VAL_U32) {
    M.cval_u32.alloc((size_t)(M.nnz ? M.nnz : 1));
    build_csc<uint32_t>(M, M.val_u32.p, M.cval_u32.p);
  } else if (M.vt == VAL_F64) {
    M.cval_f64.alloc((size_t)(M.nnz ? M.nnz : 1));
    build_csc<double>(M, M.val_f64.p, M.cval_f64.p);
  } else {
    build_csc<uint32_t>(M, nullptr, nullptr);
  }
  // task list
  M.taskptr.alloc((size_t)M.m + 1);
  if (M.m > 0) {
    DevBuf<uint32_t> nt((size_t)M.m);
    KL_LAUNCH(task_counts, (unsigned)((M.m + 255) / 256), 256, 0, M.colptr.p, M.m, nt.p);
    exclusive_scan_u32_to_i64(nt.p, M.taskptr.p, M.m);
    KL_CUDA(cudaMemcpyAsync(&M.n_tasks, M.taskptr.p + M.m, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx().stream));
    sync_stream();
    M.taskcol.alloc((size_t)(M.n_tasks ? M.n_tasks : 1));
    KL_LAUNCH(task_fill, (unsigned)((M.m + 255) / 256), 256, 0, M.taskptr.p, M.m, M.taskcol.p);
  } else {
    M.taskptr.zero(); M.n_tasks = 0;
    M.taskcol.alloc(1);
  }
  sync_stream();
  M.has_csc = true;
}

std::shared_ptr<Matrix> matrix_reduce(Matrix &M, const int64_t *sel, int64_t nsel) {
  require_ready();
  KL_REQUIRE(nsel >= 1 && sel[0] == 0, "reduce: sel[0] must be the bias (0)");
  const int64_t dim = kmerlr_coeff_dim(M.m);
  std::vector<int64_t> A((size_t)(nsel - 1)), B((size_t)(nsel - 1));
  for (int64_t j = 1; j < nsel; j++) {
    KL_REQUIRE(sel[j] > sel[j - 1] && sel[j] < dim, "reduce: coefficient indices must be ascending and in range");
    if (sel[j] >= M.m + 1) {
      int64_t i1, i2;
      kmerlr_coeff_sub2ind(M.m, sel[j] - 1, &i1, &i2);
      A[j - 1] = i1; B[j - 1] = i2;
    } else {
      A[j - 1] = sel[j] - 1; B[j - 1] = -1;
    }
  }
  auto R = std::make_shared<Matrix>();
  R->n = M.n; R->m = nsel - 1; R->vt = VAL_F64;
  R->sharded = M.sharded; R->n_global = M.n_global;
  const int64_t ns = nsel - 1;
  DevBuf<int64_t> dA((size_t)(ns ? ns : 1)), dB((size_t)(ns ? ns : 1));
  dA.upload(A.data(), (size_t)ns); dB.upload(B.data(), (size_t)ns);
  DevBuf<uint32_t> cnt((size_t)(M.n ? M.n : 1));
  R->rowptr.alloc((size_t)M.n + 1);
  unsigned wgrid = (unsigned)((M.n * 32 + 127) / 128);
  auto run = [&](auto *valp) {
    using VT = typename std::remove_const<typename std::remove_pointer<decltype(valp)>::type>::type;
    if (M.n > 0) {
      KL_LAUNCH((reduce_rows<VT, false>), wgrid, 128, 0, M.rowptr.p, M.col.p, valp, M.n, dA.p, dB.p, ns, cnt.p,
                nullptr, nullptr, nullptr);
      exclusive_scan_u32_to_i64(cnt.p, R->rowptr.p, M.n);
      KL_CUDA(cudaMemcpyAsync(&R->nnz, R->rowptr.p + M.n, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx().stream));
      sync_stream();
    } else {
      R->rowptr.zero(); R->nnz = 0;
    }
    R->col.alloc((size_t)(R->nnz ? R->nnz : 1));
    R->val_f64.alloc((size_t)(R->nnz ? R->nnz : 1));
    if (M.n > 0)
      KL_LAUNCH((reduce_rows<VT, true>), wgrid, 128, 0, M.rowptr.p, M.col.p, valp, M.n, dA.p, dB.p, ns, nullptr,
                R->rowptr.p, R->col.p, R->val_f64.p);
    sync_stream();
  };
  if (M.vt == VAL_U32) run((const uint32_t *)M.val_u32.p);
  else if (M.vt == VAL_F64) run((const double *)M.val_f64.p);
  else run((const uint32_t *)nullptr);
  if (M.has_labels) {
    R->labels.alloc((size_t)(M.n ? M.n : 1));
    KL_CUDA(cudaMemcpyAsync(R->labels.p, M.labels.p, (size_t)M.n, cudaMemcpyDeviceToDevice, ctx().stream));
    R->n_pos = M.n_pos; R->n_neg = M.n_neg; R->has_labels = true;
    sync_stream();
  }
  return R;
}

// max_i ||x_i||^2 without the bias (estimate_step_size, kmerLr_estimator_proximal.go:54-69)
double matrix_maxsq(Matrix &M) {
  if (M.has_maxsq) return M.maxsq;
  DevBuf<unsigned long long> d(1);
  d.zero();
  if (M.n > 0) {
    unsigned wgrid = (unsigned)((M.n * 32 + 127) / 128);
    if (M.vt == VAL_U32) KL_LAUNCH((row_sqnorm_max<uint32_t>), wgrid, 128, 0, M.rowptr.p, M.val_u32.p, M.n, d.p);
    else if (M.vt == VAL_F64) KL_LAUNCH((row_sqnorm_max<double>), wgrid, 128, 0, M.rowptr.p, M.val_f64.p, M.n, d.p);
    else KL_LAUNCH((row_sqnorm_max<uint32_t>), wgrid, 128, 0, M.rowptr.p, (const uint32_t *)nullptr, M.n, d.p);
  }
  if (M.sharded) comm_allreduce_max_f64((double *)d.p, 1);
  unsigned long long bits = 0;
  d.download(&bits, 1);
  sync_stream();
  double v;
  memcpy(&v, &bits, sizeof(v));
  M.maxsq = v; M.has_maxsq = true;
  return v;
}

}  // namespace kl
