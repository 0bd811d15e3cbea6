// formats.cu -- the on-disk formats either side of the hot path, byte for byte as the reference's
// fmt verbs print them:
//   * wiggle tracks of the sliding-window scores (saveWindowPredictionsWiggle, kmerLr_predict_genomic.go:37-60):
//     one "%0.15f\n" of exp(prediction) per window -- 3e8 records for a 3 Gbp genome at step 10.  The records are
//     produced ON THE DEVICE (wiggle_kernel): a prediction is a log-probability, exp of it lies in [0, 1], and
//     "%0.15f" of such a value is always 18 bytes, so record j sits at byte 18 j and the kernel is one thread per
//     record: exp, the exact decimal expansion by 128-bit integer arithmetic (round half to even on the exact binary
//     value, what Go's strconv and C's printf both do), digits staged in shared memory, coalesced 16-byte stores.
//   * export_kmers (kmerLr_data.go:127-174), KmerRegularizationPath.Export (kmerLr_estimator_path.go:41-73) and
//     Trace.Export (kmerLr_estimator_trace.go:49-80): small tables, written by the host side of the library.
// exp: Go's portable math.Exp (src/math/exp.go, the FreeBSD e_exp.c algorithm: k = round(x / ln 2), r = x - k ln2 in
// two parts, a degree-5 rational correction, Ldexp) evaluated without fused multiply-adds.  Go's amd64 / arm64 / s390x
// builds use assembly kernels for Exp that can differ from it in the last bit; a value that differs by one ulp changes
// the 15th decimal of a record in roughly one case out of ten near 1.0.  Given the same double the bytes are identical.
#include <cerrno>
#include <cmath>

#include "common.cuh"

namespace kl {

namespace {

constexpr int WIG_REC = 18;          // "d.ddddddddddddddd\n"
constexpr int WIG_THREADS = 256;     // 256 records = 4 608 bytes = 288 16-byte words per block
constexpr int64_t WIG_CHUNK = (int64_t)1 << 21;   // records per chunk of the pipeline (36 MB of text)

__host__ __device__ inline double go_ldexp(double y, int k) {
#ifdef __CUDA_ARCH__
  return ldexp(y, k);
#else
  return std::ldexp(y, k);
#endif
}

// math.Exp of the Go standard library, portable version (exp.go: exp + expmulti)
__host__ __device__ inline double go_exp(double x) {
  const double Ln2Hi = 6.93147180369123816490e-01, Ln2Lo = 1.90821492927058770002e-10, Log2e = 1.44269504088896338700e+00;
  const double Overflow = 7.09782712893383973096e+02, Underflow = -7.45133219101941108420e+02;
  const double NearZero = 1.0 / (double)(1 << 28);
  if (x != x) return x;
  if (x > Overflow) return x > 1.7976931348623157e308 ? x : (double)INFINITY;
  if (x < Underflow) return 0.0;
  if (-NearZero < x && x < NearZero) return 1.0 + x;
  int k = 0;
  if (x < 0.0) k = (int)(Log2e * x - 0.5);
  else if (x > 0.0) k = (int)(Log2e * x + 0.5);
  const double hi = x - (double)k * Ln2Hi, lo = (double)k * Ln2Lo;
  const double P1 = 1.66666666666666657415e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
               P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
  const double r = hi - lo, t = r * r;
  const double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  const double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
  return go_ldexp(y, k);
}

// x in [0, 10): N = x 10^15 rounded half to even on the exact binary value; false when the record is not 18 bytes
// (negative, not finite, or the rounded value reaches 10)
__host__ __device__ inline bool fixed15(double x, unsigned long long &N) {
  if (!(x >= 0.0 && x < 10.0)) return false;
  unsigned long long bits;
  memcpy(&bits, &x, 8);
  const int be = (int)((bits >> 52) & 0x7FFull);
  unsigned long long m = bits & ((1ull << 52) - 1ull);
  int s;                                              // x = m 2^-s
  if (be == 0) s = 1074; else { m |= 1ull << 52; s = 1075 - be; }
  const unsigned __int128 P = (unsigned __int128)m * 1000000000000000ull;     // < 2^53 2^50
  if (s >= 128) { N = 0ull; return true; }            // P 2^-s < 2^-25: rounds to zero
  N = (unsigned long long)(P >> s);
  const unsigned __int128 rem = P & ((((unsigned __int128)1) << s) - 1), half = ((unsigned __int128)1) << (s - 1);
  if (rem > half || (rem == half && (N & 1ull))) N++;
  return N < 10000000000000000ull;
}

__host__ __device__ inline void put_record(unsigned long long N, char *rec) {
  unsigned int lo8 = (unsigned int)(N % 100000000ull), hi8 = (unsigned int)(N / 100000000ull);
  for (int i = 16; i >= 9; i--) { rec[i] = (char)('0' + lo8 % 10u); lo8 /= 10u; }
  for (int i = 8; i >= 2; i--) { rec[i] = (char)('0' + hi8 % 10u); hi8 /= 10u; }
  rec[0] = (char)('0' + hi8); rec[1] = '.'; rec[17] = '\n';
}

// one thread per prediction; an irregular record (see fixed15) starts with a 0 byte and is counted
__global__ void __launch_bounds__(WIG_THREADS) wiggle_kernel(const double *__restrict__ pred, int64_t n, char *__restrict__ out,
                                                             unsigned long long *__restrict__ irregular) {
  __shared__ __align__(16) char rec[WIG_THREADS * WIG_REC];
  const int64_t j0 = (int64_t)blockIdx.x * WIG_THREADS, j = j0 + threadIdx.x;
  if (j < n) {
    unsigned long long N = 0ull;
    char *r = rec + threadIdx.x * WIG_REC;
    if (fixed15(go_exp(pred[j]), N)) put_record(N, r);
    else {
      for (int i = 0; i < WIG_REC; i++) r[i] = 0;
      atomicAdd(irregular, 1ull);
    }
  }
  __syncthreads();
  const int64_t left = n - j0, bytes = (left < WIG_THREADS ? left : WIG_THREADS) * WIG_REC;
  char *dst = out + j0 * WIG_REC;                     // 4 608 bytes per block: 16-byte aligned
  const int words = (int)(bytes >> 4);
  for (int i = threadIdx.x; i < words; i += WIG_THREADS)
    reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(rec)[i];
  for (int i = (words << 4) + threadIdx.x; i < bytes; i += WIG_THREADS) dst[i] = rec[i];
}

// fmt's %e / %f of a float64 next to C's: the same digits; only the names of the non-finite values differ
std::string go_float(const char *cfmt, double v, int width) {
  char buf[512];
  if (v != v) snprintf(buf, sizeof buf, "%*s", width, "NaN");
  else if (std::isinf(v)) snprintf(buf, sizeof buf, "%*s", width, v > 0 ? "+Inf" : "-Inf");
  else snprintf(buf, sizeof buf, cfmt, v);
  return buf;
}

struct File {
  FILE *f = nullptr;
  bool is_stdout = false;
  explicit File(const char *filename) {
    if (!filename || !*filename) { f = stdout; is_stdout = true; return; }       // "" = os.Stdout, as in the reference
    f = fopen(filename, "wb");
    if (!f) fail(KMERLR_ERR_ARG, std::string("cannot create `") + filename + "': " + strerror(errno));
  }
  void write(const void *p, size_t n) {
    if (n && fwrite(p, 1, n, f) != n) fail(KMERLR_ERR_ARG, std::string("write failed: ") + strerror(errno));
  }
  void write(const std::string &s) { write(s.data(), s.size()); }
  void close() {
    if (!f) return;
    const int rc = is_stdout ? fflush(f) : fclose(f);
    f = nullptr;
    if (rc != 0) fail(KMERLR_ERR_ARG, std::string("write failed: ") + strerror(errno));
  }
  ~File() { if (f && !is_stdout) fclose(f); }
};

}  // namespace

// records of n predictions (host or device pointer), 18 n bytes into out (host); returns the irregular ones.
// Chunks of 2^21 records through two sets of device buffers: the copy of chunk i + 1 to the device (copy stream), the
// formatting of chunk i (main stream) and the copy of the text of chunk i - 1 to the host (third stream) overlap.
int64_t wiggle_records(const double *pred, bool pred_on_device, int64_t n, char *out) {
  require_ready();
  KL_REQUIRE(n >= 0 && (n == 0 || (pred && out)), "wiggle_records: null argument");
  if (n == 0) return 0;
  const int64_t cap = n < WIG_CHUNK ? n : WIG_CHUNK;
  DevBuf<double> dp[2];
  DevBuf<char> drec[2];
  for (int b = 0; b < 2; b++) {
    if (!pred_on_device) dp[b].alloc((size_t)cap);
    drec[b].alloc((size_t)cap * WIG_REC);
  }
  DevBuf<unsigned long long> dirr(1);
  dirr.zero();
  cudaEvent_t up[2], done[2], home[2];
  for (int b = 0; b < 2; b++) {
    KL_CUDA(cudaEventCreateWithFlags(&up[b], cudaEventDisableTiming));
    KL_CUDA(cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming));
    KL_CUDA(cudaEventCreateWithFlags(&home[b], cudaEventDisableTiming));
    KL_CUDA(cudaEventRecord(home[b], ctx().stream));          // (the buffers exist, the counter is zero)
  }
  int64_t i = 0;
  for (int64_t j0 = 0; j0 < n; j0 += cap, i++) {
    const int b = (int)(i & 1);
    const int64_t c = n - j0 < cap ? n - j0 : cap;
    const double *src = pred + j0;
    if (!pred_on_device) {
      KL_CUDA(cudaStreamWaitEvent(ctx().copy_stream, home[b], 0));      // the buffers of chunk i - 2 are free again
      KL_CUDA(cudaMemcpyAsync(dp[b].p, pred + j0, (size_t)c * sizeof(double), cudaMemcpyHostToDevice, ctx().copy_stream));
      KL_CUDA(cudaEventRecord(up[b], ctx().copy_stream));
      KL_CUDA(cudaStreamWaitEvent(ctx().stream, up[b], 0));
      src = dp[b].p;
    } else {
      KL_CUDA(cudaStreamWaitEvent(ctx().stream, home[b], 0));
    }
    KL_LAUNCH(wiggle_kernel, (unsigned)((c + WIG_THREADS - 1) / WIG_THREADS), WIG_THREADS, 0, src, c, drec[b].p, dirr.p);
    KL_CUDA(cudaEventRecord(done[b], ctx().stream));
    KL_CUDA(cudaStreamWaitEvent(ctx().alt_stream, done[b], 0));
    KL_CUDA(cudaMemcpyAsync(out + j0 * WIG_REC, drec[b].p, (size_t)c * WIG_REC, cudaMemcpyDeviceToHost, ctx().alt_stream));
    KL_CUDA(cudaEventRecord(home[b], ctx().alt_stream));
  }
  KL_CUDA(cudaStreamSynchronize(ctx().alt_stream));
  unsigned long long h = 0;
  dirr.download(&h, 1);
  sync_stream();
  for (int b = 0; b < 2; b++) { cudaEventDestroy(up[b]); cudaEventDestroy(done[b]); cudaEventDestroy(home[b]); }
  return (int64_t)h;
}

// saveWindowPredictionsWiggle (kmerLr_predict_genomic.go:37-60); predictions of region i = pred[slot_off[i], slot_off[i+1])
void save_wiggle(const char *filename, const char *track_name, int64_t n_regions, const char *const *seqnames,
                 const int64_t *from, const int64_t *slot_off, const double *pred, bool pred_on_device, int64_t window_size,
                 int64_t window_step) {
  KL_REQUIRE(n_regions >= 0 && track_name && (n_regions == 0 || (seqnames && from && slot_off)), "save_wiggle: null argument");
  const int64_t total = n_regions ? slot_off[n_regions] : 0;
  for (int64_t i = 0; i < n_regions; i++) KL_REQUIRE(slot_off[i] <= slot_off[i + 1], "save_wiggle: slot offsets must not decrease");
  std::vector<char> rec((size_t)total * WIG_REC);
  std::vector<double> hpred;                      // only read for the irregular records
  const int64_t irregular = wiggle_records(pred, pred_on_device, total, rec.data());
  if (irregular && pred_on_device) {
    hpred.resize((size_t)total);
    KL_CUDA(cudaMemcpyAsync(hpred.data(), pred, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
    sync_stream();
  }
  const double *hp = pred_on_device ? hpred.data() : pred;
  File out(filename);
  out.write(std::string("track type=wiggle_0 name=") + track_name + "\n");
  for (int64_t i = 0; i < n_regions; i++) {
    char head[64];
    out.write(std::string("fixedStep chrom=") + seqnames[i]);
    snprintf(head, sizeof head, " start=%lld step=%lld span=%lld\n", (long long)(from[i] + window_size / 2),
             (long long)window_step, (long long)window_step);
    out.write(head);
    const int64_t a = slot_off[i], b = slot_off[i + 1];
    if (!irregular) { out.write(rec.data() + a * WIG_REC, (size_t)(b - a) * WIG_REC); continue; }
    for (int64_t j = a; j < b; j++) {
      if (rec[(size_t)j * WIG_REC]) out.write(rec.data() + j * WIG_REC, WIG_REC);
      else out.write(go_float("%.15f", go_exp(hp[j]), 0) + "\n");             // a value outside [0, 10): any width
    }
  }
  out.close();
}

// printed name of a class (gonetics KmerClass.String: the members joined by '|', smaller index first, a palindrome
// twice -- "gntanc|gntanc", kmerLr_test.go:40-43)
std::string class_name(const kmerlr_config &cfg, int k, uint64_t code) {
  static const char L[] = "acgtn";
  const uint64_t A = cfg.alphabet == 1 ? 5 : 4;
  auto comp = [](int c) { return c == 4 ? 4 : 3 - c; };
  int l[32];
  KL_REQUIRE(k >= 1 && k <= 31, "class_name: k out of range");
  { uint64_t c = code; for (int i = k - 1; i >= 0; i--) { l[i] = (int)(c % A); c /= A; } }
  std::vector<std::pair<uint64_t, std::string>> mem;
  auto add = [&](auto letter) {
    uint64_t c = 0; std::string s;
    for (int i = 0; i < k; i++) { const int x = letter(i); c = c * A + (uint64_t)x; s.push_back(L[x]); }
    mem.emplace_back(c, s);
  };
  add([&](int i) { return l[i]; });
  if (cfg.complement) add([&](int i) { return comp(l[i]); });
  if (cfg.reverse) add([&](int i) { return l[k - 1 - i]; });
  if (cfg.revcomp) add([&](int i) { return comp(l[k - 1 - i]); });
  for (size_t i = 1; i < mem.size(); i++)             // stable, by index
    for (size_t j = i; j > 0 && mem[j].first < mem[j - 1].first; j--) std::swap(mem[j], mem[j - 1]);
  std::string r;
  for (size_t i = 0; i < mem.size(); i++) { if (i) r.push_back('|'); r += mem[i].second; }
  return r;
}

// export_kmers (kmerLr_data.go:127-174): the class names, then every row DENSE -- "%d" of the counts, or "%e" of the
// values when the data went through a transform (as_float)
void export_kmers(Matrix &M, const kmerlr_config &cfg, const char *filename, bool as_float) {
  matrix_class_list(M);
  std::vector<int64_t> rowptr((size_t)M.n + 1);
  std::vector<int32_t> col((size_t)(M.nnz ? M.nnz : 1));
  std::vector<double> val((size_t)(M.nnz ? M.nnz : 1));
  matrix_rows(M, rowptr.data(), col.data(), val.data());
  File out(filename);
  std::string line;
  for (size_t j = 0; j < M.class_k.size(); j++) {
    if (j) line.push_back(',');
    line += class_name(cfg, M.class_k[j], M.class_code[j]);
  }
  line.push_back('\n');
  out.write(line);
  const std::string zero = as_float ? "0.000000e+00" : "0";
  char buf[64];
  for (int64_t i = 0; i < M.n; i++) {
    line.clear();
    int64_t p = rowptr[i];
    for (int64_t j = 0; j < M.m; j++) {
      if (j) line.push_back(',');
      if (p < rowptr[i + 1] && col[p] == j) {
        if (as_float) line += go_float("%e", val[p], 0);
        else { snprintf(buf, sizeof buf, "%lld", (long long)val[p]); line += buf; }      // IntAt: int(value)
        p++;
      } else line += zero;
    }
    line.push_back('\n');
    out.write(line);
  }
  out.close();
}

// KmerRegularizationPath.Export (kmerLr_estimator_path.go:41-73); theta of entry i = theta[theta_off[i], theta_off[i+1])
void export_path(const char *filename, int64_t n, const int64_t *estimator, const double *lambda, const double *norm,
                 const int64_t *theta_off, const double *theta) {
  KL_REQUIRE(filename && *filename, "export_path: no file name");
  KL_REQUIRE(n >= 0 && (n == 0 || (lambda && norm && theta_off)), "export_path: null argument");
  File out(filename);
  char buf[128];
  std::string line;
  if (estimator) { snprintf(buf, sizeof buf, "%9s ", "estimator"); line += buf; }
  snprintf(buf, sizeof buf, "%13s %13s %s\n", "lambda", "norm", "theta");
  line += buf;
  out.write(line);
  for (int64_t i = 0; i < n; i++) {
    line.clear();
    if (estimator) { snprintf(buf, sizeof buf, "%9lld ", (long long)estimator[i]); line += buf; }
    line += go_float("%13e", lambda[i], 13) + " " + go_float("%13e", norm[i], 13);
    for (int64_t j = theta_off[i]; j < theta_off[i + 1]; j++) {
      line.push_back(j == theta_off[i] ? ' ' : ',');
      line += go_float("%e", theta[j], 0);
    }
    line.push_back('\n');
    out.write(line);
  }
  out.close();
}

// format_duration (kmerLr_estimator_trace.go:28-35) on a time.Duration in nanoseconds: the float64 arithmetic of
// Duration.Hours / Minutes / Seconds (integer part + remainder / unit) and math.Mod, truncated like int()
static std::string format_duration(int64_t ns) {
  auto split = [&](int64_t unit) { return (double)(ns / unit) + (double)(ns % unit) / (double)unit; };
  const double hours = split(3600000000000LL), minutes = split(60000000000LL), seconds = split(1000000000LL);
  const double millis = (double)(ns / 1000000LL);
  char buf[96];
  snprintf(buf, sizeof buf, "%02lld:%02lld:%02lld:%02lld.%03lld", (long long)(hours / 24.0), (long long)std::fmod(hours, 24.0),
           (long long)std::fmod(minutes, 60.0), (long long)std::fmod(seconds, 60.0), (long long)std::fmod(millis, 1000.0));
  return buf;
}

// Trace.Export (kmerLr_estimator_trace.go:49-80); lambda / loss columns only when given
void export_trace(const char *filename, int64_t n, const int64_t *duration_ns, const int64_t *iteration, const double *change,
                  const int64_t *nonzero, const double *lambda, const double *loss) {
  KL_REQUIRE(filename && *filename, "export_trace: no file name");
  KL_REQUIRE(n >= 0 && (n == 0 || (duration_ns && iteration && change && nonzero)), "export_trace: null argument");
  File out(filename);
  char buf[160];
  std::string line;
  snprintf(buf, sizeof buf, "%15s %9s %12s %8s", "duration", "iteration", "change", "nonzero");
  line = buf;
  if (lambda) { snprintf(buf, sizeof buf, " %12s", "lambda"); line += buf; }
  if (loss) { snprintf(buf, sizeof buf, " %12s", "loss"); line += buf; }
  line.push_back('\n');
  out.write(line);
  for (int64_t i = 0; i < n; i++) {
    snprintf(buf, sizeof buf, "%15s %9lld ", format_duration(duration_ns[i]).c_str(), (long long)iteration[i]);
    line = buf;
    line += go_float("%12e", change[i], 12);
    snprintf(buf, sizeof buf, " %8lld", (long long)nonzero[i]);
    line += buf;
    if (lambda) line += " " + go_float("%12e", lambda[i], 12);
    if (loss) line += " " + go_float("%12e", loss[i], 12);
    line.push_back('\n');
    out.write(line);
  }
  out.close();
}

}  // namespace kl
