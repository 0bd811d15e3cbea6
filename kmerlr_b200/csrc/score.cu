// score.cu -- stage 3 of the hot path: sliding-window genomic scoring.
//
// Replaces genomicKmerLr.Predict + predict_window_genomic (kmerLr_predict_genomic.go:134-171):
// for every window start j = 0, step, 2 step, ... < len - W of every region, count the model's
// k-mer classes in the window, build the feature vector (singles, pair products, binarize),
// evaluate log sigma(x . theta) per ensemble member, summarize, and sum over models.
// Quirks kept: slot count n/step+1 with a strict loop bound (the last slot may
// stay 0.0), regions with len <= W give no output, the model Transform is not applied.
//
// Two kernels:
//   score_linear   models whose score is linear in the counts (no binarize, no pair features):
//                  per tile of the region, per-level prefix sums of the per-position coefficient
//                  in shared memory; a window is 2 reads per level
//   score_generic  everything else: one warp per window, counts in shared memory
#include "common.cuh"

#include <cmath>

namespace kl {

namespace {

constexpr uint32_t HEMPTY = 0xFFFFFFFFu;

struct DevModel {
  int M, N, op, binarize, summary;
  int n_classes, n_features, n_members;
  uint32_t hmask;          // hash table size - 1
  uint32_t levels;         // bit k set: the model has a class of length k
  const uint32_t *hkeys;   // (k << 26) | code, HEMPTY = free
  const int32_t *hvals;    // class index
  const int32_t *feat;     // n_features x 2
  const double *theta;     // n_members x (n_features + 1)
  const double *cweight;   // n_members x n_classes: sum of theta over single features of the class
};

__device__ __forceinline__ uint32_t hash_u32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ int model_lookup(const DevModel &md, int k, uint32_t code) {
  uint32_t key = ((uint32_t)k << 26) | code, h = hash_u32(key) & md.hmask;
  while (true) {
    uint32_t kk = __ldg(md.hkeys + h);
    if (kk == key) return __ldg(md.hvals + h);
    if (kk == HEMPTY) return -1;
    h = (h + 1) & md.hmask;
  }
}
// the same probe sequence over a copy of the table (shared memory)
__device__ __forceinline__ int table_lookup(const uint32_t *keys, const int32_t *vals, uint32_t hmask, int k, uint32_t code) {
  uint32_t key = ((uint32_t)k << 26) | code, h = hash_u32(key) & hmask;
  while (true) {
    uint32_t kk = keys[h];
    if (kk == key) return vals[h];
    if (kk == HEMPTY) return -1;
    h = (h + 1) & hmask;
  }
}

__device__ __forceinline__ double summarize(int summary, const double *x, int n) {
  double r;
  switch (summary) {
    case KMERLR_SUMMARY_MEAN: r = 0.0; for (int j = 0; j < n; j++) r += x[j]; return r / (double)n;
    case KMERLR_SUMMARY_PRODUCT: r = 1.0; for (int j = 0; j < n; j++) r *= x[j]; return r;
    case KMERLR_SUMMARY_MIN: r = x[0]; for (int j = 1; j < n; j++) if (r > x[j]) r = x[j]; return r;
    case KMERLR_SUMMARY_MAX: r = x[0]; for (int j = 1; j < n; j++) if (r < x[j]) r = x[j]; return r;
    default: return x[0];
  }
}

constexpr int MAX_MEMBERS = 16;

// one warp per window
__global__ void score_generic(const DevModel *__restrict__ models, int n_models, const int64_t *__restrict__ len,
                              const int64_t *__restrict__ blk, const uint32_t *__restrict__ bits2,
                              const uint16_t *__restrict__ inv16, int64_t n_regions,
                              const int64_t *__restrict__ win_off, const int64_t *__restrict__ slot_off,
                              int64_t total_windows, int64_t W, int64_t step, int max_classes,
                              int64_t slot_stride, double *__restrict__ out) {
  extern __shared__ uint32_t smem[];
  const unsigned lane = lane_id();
  uint32_t *cnt = smem + (size_t)(threadIdx.x >> 5) * max_classes;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t wi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < total_windows; wi += nwarps) {
    // region of window wi
    int64_t lo = 0, hi = n_regions;
    while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (win_off[mid] <= wi) lo = mid; else hi = mid; }
    const int64_t r = lo, w = wi - win_off[r], a = w * step;
    const uint32_t *b2 = bits2 + blk[r] * 4;
    const uint16_t *iv = inv16 + blk[r] * 4;
    double total = 0.0;
    for (int mi = 0; mi < n_models; mi++) {
      const DevModel md = models[mi];
      for (int i = lane; i < md.n_classes; i += 32) cnt[i] = 0;
      __syncwarp();
      // CountKmers / IdentifyKmers restricted to the model classes
      for (int64_t i = lane; i < W; i += 32) {
        uint32_t fw = 0, im = 0;
        for (int k = 1; k <= md.N && i + k <= W; k++) {
          int64_t idx = a + i + k - 1;
          uint32_t word = __ldg(b2 + (idx >> 4)), inv = (__ldg(iv + (idx >> 4)) >> (idx & 15)) & 1u;
          if (inv) break;
          uint32_t x = (word >> (2 * (idx & 15))) & 3u;
          fw = (fw << 2) | x;
          if (md.op == 1) im |= (3u - x) << (2 * (k - 1));
          else if (md.op == 2) im = (im << 2) | (3u - x);
          else if (md.op == 3) im |= x << (2 * (k - 1));
          if (k >= md.M && ((md.levels >> k) & 1u)) {
            uint32_t code = md.op ? min(fw, im) : fw;
            int ci = model_lookup(md, k, code);
            if (ci >= 0) atomicAdd(cnt + ci, 1u);
          }
        }
      }
      __syncwarp();
      // convert_counts(counts, features, false) (kmerLr_data.go:210-229) and x . theta per member
      double z[MAX_MEMBERS];
      for (int e = 0; e < md.n_members; e++) z[e] = 0.0;
      for (int j = lane; j < md.n_features; j += 32) {
        int i1 = md.feat[2 * j], i2 = md.feat[2 * j + 1];
        uint32_t c1 = cnt[i1], c2 = cnt[i2];
        if (md.binarize) { c1 = c1 ? 1 : 0; c2 = c2 ? 1 : 0; }
        double v = i1 == i2 ? (double)c1 : ((c1 && c2) ? (double)(c1 * c2) : 0.0);
        if (v != 0.0)
          for (int e = 0; e < md.n_members; e++) z[e] += v * md.theta[(int64_t)e * (md.n_features + 1) + j + 1];
      }
      double lp[MAX_MEMBERS];
      for (int e = 0; e < md.n_members; e++) {
        double zz = md.theta[(int64_t)e * (md.n_features + 1)] + warp_sum(z[e]);
        lp[e] = -log_add0(-zz);
      }
      total += summarize(md.summary, lp, md.n_members);
      __syncwarp();
    }
    if (lane == 0) out[slot_off[r] + w * slot_stride] = total;
  }
}

// ---- linear fast path -----------------------------------------------------------------------------
// One block per tile of TP window-start positions of one region.  Shared memory holds, for every
// level k of the model, the inclusive prefix sum Q_k over tile positions of
//   t_k(p) = class weight of the k-mer starting at p (0 if it is not a model class / invalid),
// so that  z(window at a) = theta_0 + sum_k Q_k[a + W - k] - Q_k[a - 1].
struct LinearTile { int64_t region, p0; };

constexpr int SCORE_THREADS = 1024;
// levels k <= SCORE_KD of a small model: the class weight of every code in a direct table in shared
// memory (one LDS per position and level instead of a hash probe sequence + a gather of the weight)
constexpr int SCORE_KD = 5;
constexpr int SCORE_WT = (4 + 16 + 64 + 256 + 1024);     // sum_{k=1..5} 4^k entries
__device__ __forceinline__ uint32_t score_doff(int k) { return ((1u << (2 * k)) - 4u) / 3u; }   // sum_{j<k} 4^j
__device__ __forceinline__ int level_slot(uint32_t levels, int k) {   // index of level k among the model's levels
  return ((levels >> k) & 1u) ? __popc(levels & ((1u << k) - 1u)) : -1;
}

__global__ void __launch_bounds__(SCORE_THREADS) score_linear(const DevModel *__restrict__ models, int n_models, int member,
                                                    int accumulate, const int64_t *__restrict__ len,
                                                    const int64_t *__restrict__ blk,
                                                    const uint32_t *__restrict__ bits2,
                                                    const uint16_t *__restrict__ inv16,
                                                    const LinearTile *__restrict__ tiles,
                                                    const int64_t *__restrict__ slot_off, int64_t W, int64_t step,
                                                    int TP, int model_index, int hs_smem, int64_t slot_stride,
                                                    double *__restrict__ out) {
  extern __shared__ double q[];   // nlev x (TP + 1), q[.][0] = 0; then a copy of the class table
  __shared__ double warp_tot[SCORE_THREADS / 32];
  const DevModel md = models[model_index];
  const LinearTile tl = tiles[blockIdx.x];
  const int64_t L = len[tl.region];
  const uint32_t *b2 = bits2 + blk[tl.region] * 4;
  const uint16_t *iv = inv16 + blk[tl.region] * 4;
  const int N = md.N;
  const uint32_t maskN = (1u << (2 * N)) - 1u;
  const int nlev = __popc(md.levels);
  const int stride = TP + 1;
  // small class tables are probed in shared memory (most probes miss: one LDS instead of one LDG each)
  uint32_t *skeys = reinterpret_cast<uint32_t *>(q + (size_t)nlev * stride);
  int32_t *svals = reinterpret_cast<int32_t *>(skeys + hs_smem);
  double *wtab = reinterpret_cast<double *>(svals + hs_smem);
  for (int i = threadIdx.x; i < hs_smem; i += blockDim.x) { skeys[i] = md.hkeys[i]; svals[i] = md.hvals[i]; }
  if (hs_smem) {
    for (int i = threadIdx.x; i < SCORE_WT; i += blockDim.x) wtab[i] = 0.0;
    __syncthreads();
    for (int i = threadIdx.x; i < hs_smem; i += blockDim.x) {
      const uint32_t key = skeys[i];
      if (key != HEMPTY && (int)(key >> 26) <= SCORE_KD)
        wtab[score_doff((int)(key >> 26)) + (key & 0x3FFFFFFu)] =
            __ldg(md.cweight + (int64_t)member * md.n_classes + svals[i]);
    }
    __syncthreads();
  }
  // phase 1: t_k(p) for p in [p0, p0 + TP).  The N bases starting at p are 2N consecutive bits of the
  // packed words (one funnel shift, first base in the low bits): no rolling state, so the positions
  // are simply dealt to the threads
  const int per = (TP + blockDim.x - 1) / blockDim.x;
  for (int i = threadIdx.x; i < TP; i += blockDim.x) {
    const int64_t p = tl.p0 + i;
    int len_f = 0;
    uint32_t FW = 0, IM = 0;
    if (p < L) {
      const int64_t wi = p >> 4;
      const int sh = (int)(p & 15);
      const uint32_t e = __funnelshift_r(__ldg(b2 + wi), __ldg(b2 + wi + 1), 2 * sh) & maskN;
      const int64_t rest = L - p;
      const uint32_t ivb = (((uint32_t)__ldg(iv + wi) | ((uint32_t)__ldg(iv + wi + 1) << 16)) >> sh) |
                           (1u << (rest < N ? (int)rest : N));
      len_f = __ffs(ivb) - 1;
      FW = swap_pairs(__brev(e)) >> (32 - 2 * N);
      IM = md.op == 1 ? (~e) & maskN : e;       // image of the window under revcomp / reverse
    }
    for (int k = md.M; k <= N; k++) {
      const int sl = level_slot(md.levels, k);
      if (sl < 0) continue;
      double t = 0.0;
      if (len_f >= k) {
        uint32_t fw = FW >> (2 * (N - k)), code = fw;
        if (md.op == 1 || md.op == 3) code = min(fw, IM & ((1u << (2 * k)) - 1u));
        else if (md.op == 2) code = min(fw, (~fw) & ((1u << (2 * k)) - 1u));
        if (hs_smem && k <= SCORE_KD) {
          t = wtab[score_doff(k) + code];
        } else {
          int ci = hs_smem ? table_lookup(skeys, svals, md.hmask, k, code) : model_lookup(md, k, code);
          if (ci >= 0) t = __ldg(md.cweight + (int64_t)member * md.n_classes + ci);
        }
      }
      q[sl * stride + 1 + i] = t;
    }
  }
  if ((int)threadIdx.x < nlev) q[threadIdx.x * stride] = 0.0;
  __syncthreads();
  // phase 2: inclusive scan per level (blocked: thread-local runs, warp scan, block carry)
  for (int sl = 0; sl < nlev; sl++) {
    double *row = q + sl * stride + 1;
    int i0 = threadIdx.x * per, i1 = min(TP, i0 + per);
    double s = 0.0;
    for (int i = i0; i < i1; i++) { s += row[i]; row[i] = s; }
    double x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double y = __shfl_up_sync(0xffffffffu, x, o);
      if ((int)lane_id() >= o) x += y;
    }
    if (lane_id() == 31) warp_tot[threadIdx.x >> 5] = x;
    __syncthreads();
    double carry = x - s;
    for (int wv = 0; wv < (int)(threadIdx.x >> 5); wv++) carry += warp_tot[wv];
    for (int i = i0; i < i1; i++) row[i] += carry;
    __syncthreads();
  }
  // phase 3: windows starting inside this tile whose k-mers all lie inside the tile
  const int64_t nwin_region = (L - W > 0) ? (L - W + step - 1) / step : 0;
  const int64_t w_first = (tl.p0 + step - 1) / step;
  for (int64_t w = w_first + threadIdx.x; w < nwin_region; w += blockDim.x) {
    int64_t a = w * step - tl.p0;           // tile-relative start
    if (a + W > TP) break;
    double z = md.theta[(int64_t)member * (md.n_features + 1)];
    for (int k = md.M; k <= N; k++) {
      const int sl = level_slot(md.levels, k);
      if (sl < 0 || W < k) continue;
      const double *row = q + sl * stride;  // row[i+1] = inclusive prefix through position i
      z += row[a + W - k + 1] - row[a];
    }
    double lp = -log_add0(-z);
    int64_t o = slot_off[tl.region] + w * slot_stride;
    if (accumulate) out[o] += lp; else out[o] = lp;
  }
}

struct HostModel {
  DevModel d;
  DevBuf<uint32_t> hkeys;
  DevBuf<int32_t> hvals, feat;
  DevBuf<double> theta, cweight;
  bool linear;
};

}  // namespace

// layout 0: predict_window_genomic -- n/step + 1 slots per region, window j*step in slot j
// layout 1: predict_window (kmerLr_predict.go:89-124) -- n = len - W slots, window starting at j in slot j
// feed: the sequences are still on the host and arrive in chunks of whole rows (kmerlr_score_windows with host
// buffers).  Count models without pair features then work chunk by chunk: the copy of chunk c + 1 (copy stream)
// and the copy of the scores of chunk c - 1 back to the host (third stream) overlap the scoring of chunk c.
void score_windows(const kmerlr_model *models, int n_models, const SeqSet &s, int64_t W, int64_t step,
                   double *out_host, std::shared_ptr<Object> *out_dev, int layout, SeqFeed *feed) {
  require_ready();
  KL_REQUIRE(n_models >= 1 && W >= 1 && step >= 1, "score_windows: bad arguments");
  std::vector<std::unique_ptr<HostModel>> hm;
  int max_classes = 1;
  bool all_linear = true;
  for (int mi = 0; mi < n_models; mi++) {
    const kmerlr_model &m = models[mi];
    KL_REQUIRE(m.cfg.alphabet == 0, "score_windows: only the nucleotide alphabet is implemented on the GPU path");
    KL_REQUIRE(m.cfg.M >= 1 && m.cfg.M <= m.cfg.N && m.cfg.N <= 13, "score_windows: need 1 <= M <= N <= 13");
    int nops = (m.cfg.complement != 0) + (m.cfg.reverse != 0) + (m.cfg.revcomp != 0);
    KL_REQUIRE(nops <= 1, "score_windows: at most one of complement / reverse / revcomp");
    KL_REQUIRE(m.n_members >= 1 && m.n_members <= MAX_MEMBERS, "score_windows: 1..16 ensemble members");
    KL_REQUIRE(m.n_members == 1 || m.summary != KMERLR_SUMMARY_NONE, "no summary given for ensemble classifier");
    KL_REQUIRE(m.n_classes >= 0 && m.n_classes <= 8192, "score_windows: at most 8192 model classes");
    auto h = std::make_unique<HostModel>();
    DevModel &d = h->d;
    d.M = m.cfg.M; d.N = m.cfg.N; d.binarize = m.cfg.binarize != 0; d.summary = m.summary;
    d.op = m.cfg.revcomp ? 1 : (m.cfg.complement ? 2 : (m.cfg.reverse ? 3 : 0));
    d.n_classes = (int)m.n_classes; d.n_features = (int)m.n_features; d.n_members = (int)m.n_members;
    uint32_t hs = 16; while (hs < 2 * (uint32_t)m.n_classes + 2) hs <<= 1;
    std::vector<uint32_t> keys(hs, HEMPTY); std::vector<int32_t> vals(hs, -1);
    d.levels = 0;
    auto mix = [](uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; };
    for (int64_t c = 0; c < m.n_classes; c++) {
      int k = m.class_k[c];
      KL_REQUIRE(k >= d.M && k <= d.N && m.class_code[c] < (1ull << (2 * k)), "score_windows: class outside [M,N]");
      uint32_t key = ((uint32_t)k << 26) | (uint32_t)m.class_code[c], hh = mix(key) & (hs - 1);
      while (keys[hh] != HEMPTY) { KL_REQUIRE(keys[hh] != key, "score_windows: duplicate model class"); hh = (hh + 1) & (hs - 1); }
      keys[hh] = key; vals[hh] = (int32_t)c;
      d.levels |= 1u << k;
    }
    d.hmask = hs - 1;
    h->linear = !d.binarize;
    std::vector<double> cwt((size_t)(m.n_members * (m.n_classes ? m.n_classes : 1)), 0.0);
    for (int64_t j = 0; j < m.n_features; j++) {
      int32_t i1 = m.features[2 * j], i2 = m.features[2 * j + 1];
      KL_REQUIRE(i1 >= 0 && i1 < m.n_classes && i2 >= 0 && i2 < m.n_classes, "score_windows: feature index out of range");
      if (i1 != i2) h->linear = false;
      else for (int64_t e = 0; e < m.n_members; e++) cwt[(size_t)(e * m.n_classes + i1)] += m.theta[e * (m.n_features + 1) + j + 1];
    }
    if (!h->linear) all_linear = false;
    h->hkeys.alloc(hs); h->hvals.alloc(hs);
    h->hkeys.upload(keys.data(), hs); h->hvals.upload(vals.data(), hs);
    h->feat.alloc((size_t)(m.n_features ? 2 * m.n_features : 1));
    h->feat.upload(m.features, (size_t)(2 * m.n_features));
    h->theta.alloc((size_t)(m.n_members * (m.n_features + 1)));
    h->theta.upload(m.theta, (size_t)(m.n_members * (m.n_features + 1)));
    h->cweight.alloc(cwt.size());
    h->cweight.upload(cwt.data(), cwt.size());
    d.hkeys = h->hkeys.p; d.hvals = h->hvals.p; d.feat = h->feat.p; d.theta = h->theta.p; d.cweight = h->cweight.p;
    if (d.n_classes > max_classes) max_classes = d.n_classes;
    hm.push_back(std::move(h));
  }
  sync_stream();
  // a summary of several members is not linear in the per-member scores unless it is the mean
  for (auto &h : hm)
    if (h->d.n_members > 1) all_linear = false;
  std::vector<DevModel> dm;
  for (auto &h : hm) dm.push_back(h->d);
  DevBuf<DevModel> dmodels((size_t)n_models);
  dmodels.upload(dm.data(), (size_t)n_models);
  // slots and windows per region
  std::vector<int64_t> len((size_t)s.n);
  s.len.download(len.data(), (size_t)s.n);
  sync_stream();
  std::vector<int64_t> slot_off((size_t)s.n + 1, 0), win_off((size_t)s.n + 1, 0);
  for (int64_t r = 0; r < s.n; r++) {
    slot_off[r + 1] = slot_off[r] + (layout == 1 ? (len[r] - W > 0 ? len[r] - W : 0) : kmerlr_window_slots(len[r], W, step));
    win_off[r + 1] = win_off[r] + (len[r] - W > 0 ? (len[r] - W + step - 1) / step : 0);
  }
  const int64_t total_slots = slot_off[s.n], total_windows = win_off[s.n];
  auto outbuf = std::make_shared<Matrix>();   // reuse Matrix.val_f64 as a plain device vector
  outbuf->val_f64.alloc((size_t)(total_slots ? total_slots : 1));
  outbuf->val_f64.zero();
  outbuf->n = total_slots;
  DevBuf<int64_t> dslot((size_t)s.n + 1), dwin((size_t)s.n + 1);
  dslot.upload(slot_off.data(), (size_t)s.n + 1);
  dwin.upload(win_off.data(), (size_t)s.n + 1);
  bool copied_home = false;       // the scores went to out_host chunk by chunk
  if (total_windows > 0) {
    if (all_linear) {
      for (int mi = 0; mi < n_models; mi++) {
        const DevModel &d = dm[mi];
        int nlev = __builtin_popcount(d.levels);
        if (nlev == 0) nlev = 1;
        // tile: as many positions as 200 KB of prefix sums allow, at least W + step
        const int64_t hs = (int64_t)d.hmask + 1;
        const int hs_smem = hs * 8 <= 16 * 1024 ? (int)hs : 0;
        const int wt_bytes = hs_smem ? SCORE_WT * 8 : 0;      // direct weight tables of the levels <= SCORE_KD
        int64_t TP = (200 * 1024 - hs_smem * 8 - wt_bytes) / (8 * nlev) - 1;
        if (TP > 16384) TP = 16384;
        KL_REQUIRE(TP >= W + step, "score_windows: window too large for the shared-memory tile");
        // windows per tile: starts a with a + W <= TP (a tile advances by that many steps)
        int64_t starts = (TP - W) / step + 1;
        std::vector<LinearTile> tiles;
        for (int64_t r = 0; r < s.n; r++) {
          int64_t nw = win_off[r + 1] - win_off[r];
          for (int64_t w0 = 0; w0 < nw; w0 += starts) tiles.push_back(LinearTile{r, w0 * step});
        }
        DevBuf<LinearTile> dt(tiles.size() ? tiles.size() : 1);
        dt.upload(tiles.data(), tiles.size());
        size_t smem = (size_t)nlev * (size_t)(TP + 1) * sizeof(double) + (size_t)hs_smem * 8 + (size_t)wt_bytes;
        KL_CUDA(cudaFuncSetAttribute(score_linear, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // chunk by chunk (one chunk when the sequences are resident): the tiles are sorted by region
        const int nch = feed ? feed->chunks() : 1;
        size_t t0 = 0;
        for (int c = 0; c < nch; c++) {
          const int64_t r0 = feed ? feed->first_row(c) : 0, r1 = feed ? feed->first_row(c + 1) : s.n;
          if (feed && mi == 0) feed->feed(c);
          size_t t1 = t0;
          while (t1 < tiles.size() && tiles[t1].region < r1) t1++;
          if (t1 > t0)
            KL_LAUNCH(score_linear, (unsigned)(t1 - t0), SCORE_THREADS, smem, dmodels.p, n_models, 0, mi > 0 ? 1 : 0, s.len.p,
                      s.blk.p, s.bits2.p, s.inv16.p, dt.p + t0, dslot.p, W, step, (int)TP, mi, hs_smem,
                      layout == 1 ? step : (int64_t)1, outbuf->val_f64.p);
          t0 = t1;
          // the scores of the chunk go home while the next chunk is scored (last model only: the models add up)
          if (feed && out_host && mi == n_models - 1 && slot_off[r1] > slot_off[r0]) {
            cudaEvent_t ev;
            KL_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            KL_CUDA(cudaEventRecord(ev, ctx().stream));
            KL_CUDA(cudaStreamWaitEvent(ctx().alt_stream, ev, 0));
            KL_CUDA(cudaMemcpyAsync(out_host + slot_off[r0], outbuf->val_f64.p + slot_off[r0],
                                    (size_t)(slot_off[r1] - slot_off[r0]) * sizeof(double), cudaMemcpyDeviceToHost, ctx().alt_stream));
            KL_CUDA(cudaEventDestroy(ev));            // (released once the event has completed)
            copied_home = true;
          }
        }
        sync_stream();
      }
      if (copied_home) KL_CUDA(cudaStreamSynchronize(ctx().alt_stream));
    } else {
      if (feed) feed->feed(-1);
      size_t smem = (size_t)4 * (size_t)max_classes * sizeof(uint32_t);
      KL_CUDA(cudaFuncSetAttribute(score_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int64_t blocks = (total_windows + 3) / 4, cap = (int64_t)ctx().sm_count * 16;
      if (blocks > cap) blocks = cap;
      KL_LAUNCH(score_generic, (unsigned)blocks, 128, smem, dmodels.p, n_models, s.len.p, s.blk.p, s.bits2.p,
                s.inv16.p, s.n, dwin.p, dslot.p, total_windows, W, step, max_classes, layout == 1 ? step : (int64_t)1,
                outbuf->val_f64.p);
    }
  }
  else if (feed) feed->feed(-1);
  if (out_host && !copied_home) outbuf->val_f64.download(out_host, (size_t)total_slots);
  sync_stream();
  if (out_dev) *out_dev = outbuf;
}

}  // namespace kl
