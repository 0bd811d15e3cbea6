// extract.cu -- stage 1 of the hot path: sequences -> sparse k-mer count rows.
//
// Replaces scan_sequences + NewKmerCountsList/SetKmers + convert_counts_list
// (kmerLr_data.go:197-284,306-357; k-mer semantics of gonetics restated in SURVEY.md 8c).
//
// Design (one warp per sequence, no hash tables, no float atomics):
//   * sequences are 2-bit packed in HBM (+1 invalid bit per base), see SeqSet;
//   * "two strand suffix sort": the k-mer class counts of ALL k in [M,N] follow from ONE sort of
//     the 2L suffix keys (N-prefix of every suffix of S and of op(S), op = revcomp / complement /
//     reverse): the k-mers of level k are the distinct k-prefixes of the sorted keys, the class
//     count of canonical u is #forward keys with prefix u + #op-strand keys with prefix u
//     (op-strand keys are ignored when u == op(u));
//   * the sort is a bitonic network over 32*E keys held in registers (E per lane), in-register
//     compare-exchanges for distances < E and warp shuffles above;
//   * levels k <= 5 never touch the sort: per-warp direct-address count tables in shared memory;
//   * run boundaries / counts / output slots come from warp ballots, so rows leave the kernel
//     already sorted by (k, code) = by final column index;
//   * columns are ranks in the bitmap of observed (or frozen) classes.
#include "common.cuh"

namespace kl {

namespace {

constexpr uint32_t SENT = 0xFFFFFFFFu;
constexpr int KS_MAX = 5;          // levels <= KS_MAX use direct tables
constexpr int MAX_N = 13;          // 2*13 code bits + strand + 4 len bits = 31 bits
constexpr int TAB_WORDS = 688;     // (4+16+64+256+1024)/2 = 682 packed u16 pairs, padded

struct XParams {
  int M, N, op, binarize;
  int ks_lo, ks_hi;                // table levels (empty if ks_lo > ks_hi)
  int big_lo;                      // sorted levels [big_lo, N] (empty if big_lo > N)
  int mark;                        // mark observed classes in the bitmap
  uint32_t level_off[MAX_N + 2];   // dense id of (k, code 0)
  int64_t stride, n;
  const int64_t *len, *blk;
  const uint32_t *bits2;
  const uint16_t *inv16;
  uint32_t *st_id, *st_cnt, *rowcnt, *bitmap;
};

__device__ __forceinline__ uint32_t swap_pairs(uint32_t y) {
  return ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
}
// image of the k-mer code u under the strand operation
__device__ __forceinline__ uint32_t kmer_op(uint32_t u, int k, int op) {
  if (op == 1) return swap_pairs(__brev(~u)) >> (32 - 2 * k);   // reverse complement
  if (op == 2) return (~u) & ((1u << (2 * k)) - 1u);            // complement
  return swap_pairs(__brev(u)) >> (32 - 2 * k);                 // reverse
}

__device__ __forceinline__ uint32_t tab_off(int k) {  // sum_{j=1}^{k-1} 4^j
  return ((1u << (2 * k)) - 4u) / 3u;
}

// ---- pack: ASCII -> 2 bit codes + invalid mask --------------------------------------------------
__global__ void pack_kernel(const uint8_t *__restrict__ seq, const int64_t *__restrict__ off,
                            const int64_t *__restrict__ blk, int64_t n, int64_t total_words,
                            uint32_t *__restrict__ bits2, uint16_t *__restrict__ inv16) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= total_words) return;
  // sequence owning word w: last i with blk[i]*4 <= w
  int64_t lo = 0, hi = n;
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (blk[mid] * 4 <= w) lo = mid; else hi = mid;
  }
  int64_t i = lo, j0 = (w - blk[i] * 4) * 16, L = off[i + 1] - off[i];
  const uint8_t *s = seq + off[i];
  uint32_t bits = 0, inv = 0;
#pragma unroll
  for (int j = 0; j < 16; j++) {
    if (j0 + j < L) {
      uint32_t ch = s[j0 + j] | 0x20u, c;   // fold case
      if (ch == 'a') c = 0; else if (ch == 'c') c = 1; else if (ch == 'g') c = 2; else if (ch == 't') c = 3;
      else { c = 0; inv |= 1u << j; }
      bits |= c << (2 * j);
    }
  }
  bits2[w] = bits;
  inv16[w] = (uint16_t)inv;
}

// ---- bitonic sort of 32*E keys, index = lane*E + r, ascending ----------------------------------
__device__ __forceinline__ void ce(uint32_t &a, uint32_t &b) {
  uint32_t lo = min(a, b), hi = max(a, b);
  a = lo; b = hi;
}

template <int E>
__device__ __forceinline__ void warp_sort(uint32_t (&K)[E], unsigned lane) {
#pragma unroll
  for (int k = 2; k <= 32 * E; k <<= 1) {
    // first stage of the merge: partner = i ^ (k-1)
    if (k <= E) {
#pragma unroll
      for (int r = 0; r < E; r++) {
        int pr = r ^ (k - 1);
        if (pr > r) ce(K[r], K[pr]);
      }
    } else {
      const int lm = k / E - 1;                       // lane xor mask
      const bool keepmin = (lane & ((k / E) >> 1)) == 0;
#pragma unroll
      for (int r = 0; r < E / 2; r++) {
        uint32_t a = K[r], b = K[E - 1 - r];
        uint32_t va = __shfl_xor_sync(0xffffffffu, b, lm);   // partner's K[E-1-r]
        uint32_t vb = __shfl_xor_sync(0xffffffffu, a, lm);   // partner's K[r]
        K[r] = keepmin ? min(a, va) : max(a, va);
        K[E - 1 - r] = keepmin ? min(b, vb) : max(b, vb);
      }
    }
    // remaining half-cleaners
#pragma unroll
    for (int j = k / 4; j >= 1; j >>= 1) {
      if (j < E) {
#pragma unroll
        for (int r = 0; r < E; r++)
          if ((r & j) == 0) ce(K[r], K[r | j]);
      } else {
        const int lm = j / E;
        const bool keepmin = (lane & lm) == 0;
#pragma unroll
        for (int r = 0; r < E; r++) {
          uint32_t v = __shfl_xor_sync(0xffffffffu, K[r], lm);
          K[r] = keepmin ? min(K[r], v) : max(K[r], v);
        }
      }
    }
  }
}

// ---- per-lane streaming reader of the packed sequence -------------------------------------------
struct BaseReader {
  const uint32_t *bits2;
  const uint16_t *inv16;
  int64_t L;
  int64_t cur_word;
  uint32_t w, iv;
  __device__ __forceinline__ void init(const uint32_t *b, const uint16_t *m, int64_t len) {
    bits2 = b; inv16 = m; L = len; cur_word = -1; w = 0; iv = 0;
  }
  // base idx -> (code, invalid); out of range = invalid
  __device__ __forceinline__ void get(int64_t idx, uint32_t &x, uint32_t &inv) {
    if (idx < 0 || idx >= L) { x = 0; inv = 1; return; }
    int64_t wi = idx >> 4;
    if (wi != cur_word) { cur_word = wi; w = __ldg(bits2 + wi); iv = __ldg(inv16 + wi); }
    int sh = (int)(idx & 15);
    x = (w >> (2 * sh)) & 3u;
    inv = (iv >> sh) & 1u;
  }
};

struct Emitter {
  uint32_t *st_id, *st_cnt, *bitmap;
  int64_t rowbase;
  uint32_t cursor;
  int binarize, mark;
  __device__ __forceinline__ void emit(bool flag, uint32_t id, uint32_t cnt) {
    unsigned em = __ballot_sync(0xffffffffu, flag);
    if (flag) {
      int64_t pos = rowbase + cursor + __popc(em & lanemask_lt());
      st_id[pos] = id;
      if (!binarize) st_cnt[pos] = cnt;
      if (mark) {
        uint32_t bit = 1u << (id & 31);
        if (!(bitmap[id >> 5] & bit)) atomicOr(bitmap + (id >> 5), bit);
      }
    }
    cursor += __popc(em);
  }
};

// ---- the extraction kernel: one warp per sequence ------------------------------------------------
// E = keys per lane (0: no sorted levels, sequences of any length)
template <int E>
__global__ void __launch_bounds__(128) extract_kernel(XParams P) {
  constexpr int EE = E > 0 ? E : 1;
  extern __shared__ uint32_t smem[];
  const unsigned lane = lane_id();
  const int warp_in_block = threadIdx.x >> 5;
  constexpr int KEYW = E > 0 ? 32 * (E + 1) : 0;
  uint32_t *sk = smem + (size_t)warp_in_block * (KEYW + TAB_WORDS);
  uint32_t *tab = sk + KEYW;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int N = P.N, M = P.M, op = P.op;
  const bool two = op != 0;
  const uint32_t maskN = (N == 16) ? 0xFFFFFFFFu : ((1u << (2 * N)) - 1u);
  const uint32_t maskNb = (1u << N) - 1u;

  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp_in_block; row < P.n; row += nwarps) {
    const int64_t L = P.len[row];
    const uint32_t *b2 = P.bits2 + P.blk[row] * 4;
    const uint16_t *iv16 = P.inv16 + P.blk[row] * 4;
    // reset count tables
    if (P.ks_lo <= P.ks_hi)
      for (int i = lane; i < TAB_WORDS; i += 32) tab[i] = 0;
    __syncwarp();

    // ---- key generation: lane handles steps u in [u0, u0+steps) of [0, L+N-1) -------------------
    const int64_t total_steps = L + N - 1;
    const int64_t steps = (total_steps + 31) / 32;
    const int64_t u0 = (int64_t)lane * steps;
    uint32_t K[EE];
#pragma unroll
    for (int r = 0; r < EE; r++) K[r] = SENT;
    {
      BaseReader rd; rd.init(b2, iv16, L);
      uint32_t FW = 0, S2 = 0, IV = 0xFFFFFFFFu;
      auto consume = [&](int64_t idx) {
        uint32_t x, inv; rd.get(idx, x, inv);
        FW = ((FW << 2) | x) & maskN;
        if (op == 1) S2 = (S2 >> 2) | ((3u - x) << (2 * (N - 1)));
        else if (op == 3) S2 = (S2 >> 2) | (x << (2 * (N - 1)));
        IV = (IV << 1) | inv;
      };
      auto produce = [&](uint32_t &kf, uint32_t &k2) {
        uint32_t xw = IV & maskNb;
        int len_f = N - 32 + __clz(xw);                 // valid bases forward from window start
        int len_r = xw ? (__ffs(xw) - 1) : N;           // valid bases backward from window end
        kf = SENT; k2 = SENT;
        if (len_f >= M) {
          uint32_t code = FW & ~((1u << (2 * (N - len_f))) - 1u);
          kf = (code << 5) | (uint32_t)len_f;
          if (op == 2) k2 = ((((~FW) & maskN) & ~((1u << (2 * (N - len_f))) - 1u)) << 5) | 16u | (uint32_t)len_f;
          // direct tables for the small levels (forward strand only)
          for (int k = P.ks_lo; k <= P.ks_hi; k++) {
            if (len_f >= k) {
              uint32_t idx = tab_off(k) + (FW >> (2 * (N - k)));
              atomicAdd(tab + (idx >> 1), 1u << (16 * (idx & 1)));
            }
          }
        }
        if ((op == 1 || op == 3) && len_r >= M) {
          uint32_t code = S2 & ~((1u << (2 * (N - len_r))) - 1u);
          k2 = (code << 5) | 16u | (uint32_t)len_r;
        }
      };
      for (int64_t idx = u0 - N + 1; idx < u0; idx++) consume(idx);
      if (E > 0) {
        constexpr int PMAX = EE;   // slots: two strands -> 2 per step, else 1
#pragma unroll
        for (int i = 0; i < PMAX; i++) {
          bool active = two ? (2 * i + 1 < EE) : true;
          if (active && i < steps && u0 + i < total_steps) {
            consume(u0 + i);
            uint32_t kf, k2; produce(kf, k2);
            if (P.big_lo <= N) {
              if (two) { if (2 * i + 1 < EE) { K[(2 * i) % EE] = k2; K[(2 * i + 1) % EE] = kf; } }
              else K[i] = kf;
            }
          }
        }
      } else {
        for (int64_t i = 0; i < steps && u0 + i < total_steps; i++) {
          consume(u0 + i);
          uint32_t kf, k2; produce(kf, k2);
        }
      }
    }
    __syncwarp();

    Emitter em;
    em.st_id = P.st_id; em.st_cnt = P.st_cnt; em.bitmap = P.bitmap;
    em.rowbase = row * P.stride; em.cursor = 0; em.binarize = P.binarize; em.mark = P.mark;

    // ---- small levels from the tables ---------------------------------------------------------
    for (int k = P.ks_lo; k <= P.ks_hi; k++) {
      const uint32_t nk = 1u << (2 * k), toff = tab_off(k);
      for (uint32_t base = 0; base < nk; base += 32) {
        uint32_t u = base + lane;
        bool flag = false; uint32_t cnt = 0;
        if (u < nk) {
          uint32_t i1 = toff + u;
          cnt = (tab[i1 >> 1] >> (16 * (i1 & 1))) & 0xFFFFu;
          if (two) {
            uint32_t ru = kmer_op(u, k, op);
            if (u > ru) cnt = 0;
            else if (u < ru) { uint32_t i2 = toff + ru; cnt += (tab[i2 >> 1] >> (16 * (i2 & 1))) & 0xFFFFu; }
          }
          flag = cnt > 0;
        }
        em.emit(flag, P.level_off[k] + u, cnt);
      }
    }

    // ---- sorted levels ---------------------------------------------------------------------------
    if (E > 0 && P.big_lo <= N) {
      warp_sort<EE>(K, lane);
#pragma unroll
      for (int r = 0; r < EE; r++) sk[lane * (EE + 1) + r] = K[r];
      __syncwarp();
      for (int k = P.big_lo; k <= N; k++) {
        const int s = 2 * (N - k) + 5;
        uint32_t carry_f = 0, carry_r = 0, last_key = 0;
        for (int g = 0; g < EE; g++) {
          const unsigned sidx = (unsigned)g * 32u + lane;
          const uint32_t key = sk[(sidx / EE) * (EE + 1) + (sidx % EE)];
          uint32_t kprev = __shfl_up_sync(0xffffffffu, key, 1);
          if (lane == 0) kprev = last_key;
          const uint32_t p = key >> s, pp = kprev >> s;
          const bool first = (g == 0 && lane == 0);
          const bool head = first || (p != pp);
          const unsigned hm = __ballot_sync(0xffffffffu, head);
          const bool valid = key != SENT && (int)(key & 15u) >= k;
          const unsigned vf = __ballot_sync(0xffffffffu, valid && !(key & 16u));
          const unsigned vr = __ballot_sync(0xffffffffu, valid && (key & 16u));
          // a head closes the run that ended just before it
          bool flag = false; uint32_t cnt = 0;
          if (head && !first) {
            unsigned below = hm & lanemask_lt();
            unsigned range; uint32_t cf, cr;
            if (below) {
              int lower = 31 - __clz(below);
              range = lanemask_lt() & ~((1u << lower) - 1u);
              cf = __popc(vf & range); cr = __popc(vr & range);
            } else {
              range = lanemask_lt();
              cf = __popc(vf & range) + carry_f; cr = __popc(vr & range) + carry_r;
            }
            if (two) {
              uint32_t ru = kmer_op(pp, k, op);
              if (pp > ru) cnt = 0; else if (pp == ru) cnt = cf; else cnt = cf + cr;
            } else cnt = cf;
            flag = cnt > 0;
          }
          em.emit(flag, P.level_off[k] + pp, cnt);
          // carry of the run that is still open at the end of this group
          if (hm) {
            int hl = 31 - __clz(hm);
            unsigned range = ~((1u << hl) - 1u);
            carry_f = __popc(vf & range); carry_r = __popc(vr & range);
          } else {
            carry_f += __popc(vf); carry_r += __popc(vr);
          }
          last_key = __shfl_sync(0xffffffffu, key, 31);
        }
        // the last run of the array is the sentinel run (at least 2(N-1) slots are never valid)
      }
      __syncwarp();
    }
    if (lane == 0) P.rowcnt[row] = em.cursor;
  }
}

// ---- bitmap helpers -------------------------------------------------------------------------------
__global__ void popc_words(const uint32_t *__restrict__ bm, int64_t nw, uint32_t *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nw) out[i] = __popc(bm[i]);
}
__global__ void bitmap_to_bytes(const uint32_t *__restrict__ bm, int64_t nbits, uint8_t *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nbits) out[i] = (bm[i >> 5] >> (i & 31)) & 1u;
}
__global__ void bytes_to_bitmap(const uint8_t *__restrict__ in, int64_t nbits, uint32_t *__restrict__ bm, int64_t nw) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  uint32_t v = 0;
  for (int b = 0; b < 32; b++) {
    int64_t i = w * 32 + b;
    if (i < nbits && in[i]) v |= 1u << b;
  }
  bm[w] = v;
}
// ids of the set bits at their rank positions
__global__ void enumerate_bits(const uint32_t *__restrict__ bm, const uint32_t *__restrict__ rank, int64_t nw,
                               uint32_t *__restrict__ ids) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  uint32_t v = bm[w], r = rank[w];
  while (v) {
    int b = __ffs(v) - 1;
    ids[r++] = (uint32_t)(w * 32 + b);
    v &= v - 1;
  }
}

// ---- compaction: staging -> final CSR with rank-mapped columns (warp per row) ----------------------
__global__ void count_kept(const uint32_t *__restrict__ st_id, const uint32_t *__restrict__ rowcnt, int64_t stride,
                           int64_t n, const uint32_t *__restrict__ bm, uint32_t *__restrict__ kept) {
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  unsigned lane = lane_id();
  uint32_t c = rowcnt[row], tot = 0;
  for (uint32_t j = lane; j < c; j += 32) {
    uint32_t id = st_id[row * stride + j];
    tot += (bm[id >> 5] >> (id & 31)) & 1u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  if (lane == 0) kept[row] = tot;
}

template <bool FILTER>
__global__ void compact_rows(const uint32_t *__restrict__ st_id, const uint32_t *__restrict__ st_cnt,
                             const uint32_t *__restrict__ rowcnt, int64_t stride, int64_t n,
                             const uint32_t *__restrict__ bm, const uint32_t *__restrict__ rank,
                             const int64_t *__restrict__ rowptr, uint32_t *__restrict__ col,
                             uint32_t *__restrict__ val) {
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  unsigned lane = lane_id();
  uint32_t c = rowcnt[row];
  int64_t outp = rowptr[row];
  for (uint32_t j0 = 0; j0 < c; j0 += 32) {
    uint32_t j = j0 + lane;
    bool keep = false; uint32_t id = 0, w = 0;
    if (j < c) {
      id = st_id[row * stride + j];
      w = __ldg(bm + (id >> 5));
      keep = FILTER ? ((w >> (id & 31)) & 1u) : true;
    }
    unsigned km = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      int64_t pos = outp + __popc(km & lanemask_lt());
      col[pos] = __ldg(rank + (id >> 5)) + __popc(w & ((1u << (id & 31)) - 1u));
      if (val) val[pos] = st_cnt[row * stride + j];
    }
    outp += __popc(km);
  }
}

// ---- explicit feature lists: singles and pair products (convert_counts, kmerLr_data.go:210-229) ----
__device__ __forceinline__ uint32_t row_lookup(const uint32_t *col, const uint32_t *val, int64_t a, int64_t b, uint32_t c) {
  int64_t lo = a, hi = b;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (col[mid] < c) lo = mid + 1; else hi = mid;
  }
  if (lo < b && col[lo] == c) return val ? val[lo] : 1u;
  return 0u;
}
template <bool WRITE>
__global__ void features_rows(const int64_t *__restrict__ rowptr, const uint32_t *__restrict__ col,
                              const uint32_t *__restrict__ val, int64_t n, const int32_t *__restrict__ feat,
                              int64_t nf, uint32_t *__restrict__ cnt_out, const int64_t *__restrict__ orowptr,
                              uint32_t *__restrict__ ocol, uint32_t *__restrict__ oval) {
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  unsigned lane = lane_id();
  int64_t a = rowptr[row], b = rowptr[row + 1];
  int64_t outp = WRITE ? orowptr[row] : 0;
  uint32_t tot = 0;
  for (int64_t j0 = 0; j0 < nf; j0 += 32) {
    int64_t j = j0 + lane;
    uint32_t v = 0;
    if (j < nf) {
      int32_t i1 = feat[2 * j], i2 = feat[2 * j + 1];
      uint32_t c1 = row_lookup(col, val, a, b, (uint32_t)i1);
      if (i1 == i2) v = c1;
      else if (c1) v = c1 * row_lookup(col, val, a, b, (uint32_t)i2);
    }
    unsigned km = __ballot_sync(0xffffffffu, v != 0);
    if (WRITE && v != 0) {
      int64_t pos = outp + __popc(km & lanemask_lt());
      ocol[pos] = (uint32_t)j;
      if (oval) oval[pos] = v;
    }
    outp += __popc(km);
    tot += __popc(km);
  }
  if (!WRITE && lane == 0) cnt_out[row] = tot;
}

template <int E>
void launch_extract(const XParams &P) {
  constexpr int KEYW = E > 0 ? 32 * (E + 1) : 0;
  size_t smem = (size_t)4 * (KEYW + TAB_WORDS) * sizeof(uint32_t);
  KL_CUDA(cudaFuncSetAttribute(extract_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  KL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, extract_kernel<E>, 128, smem));
  if (per_sm < 1) per_sm = 1;
  int64_t blocks = (int64_t)ctx().sm_count * per_sm;
  int64_t need = (P.n + 3) / 4;
  if (blocks > need) blocks = need;
  if (blocks < 1) blocks = 1;
  KL_LAUNCH((extract_kernel<E>), (unsigned)blocks, 128, smem, P);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
std::shared_ptr<SeqSet> sequences_create(const uint8_t *seq, const int64_t *off, int64_t n) {
  require_ready();
  KL_REQUIRE(n >= 0 && off != nullptr, "sequences: bad arguments");
  auto s = std::make_shared<SeqSet>();
  s->n = n;
  std::vector<int64_t> len((size_t)n), blk((size_t)n + 1);
  blk[0] = 0;
  for (int64_t i = 0; i < n; i++) {
    len[i] = off[i + 1] - off[i];
    KL_REQUIRE(len[i] >= 0, "sequences: offsets must be non-decreasing");
    if (len[i] > s->max_len) s->max_len = len[i];
    blk[i + 1] = blk[i] + (len[i] + 63) / 64;
  }
  s->total_bases = n ? off[n] - off[0] : 0;
  s->total_blocks = blk[n];
  s->len.alloc((size_t)(n ? n : 1));
  s->blk.alloc((size_t)n + 1);
  s->len.upload(len.data(), (size_t)n);
  s->blk.upload(blk.data(), (size_t)n + 1);
  int64_t words = s->total_blocks * 4;
  s->bits2.alloc((size_t)(words ? words : 1));
  s->inv16.alloc((size_t)(words ? words : 1));
  if (words > 0) {
    DevBuf<uint8_t> raw((size_t)s->total_bases);
    DevBuf<int64_t> doff((size_t)n + 1);
    std::vector<int64_t> rel((size_t)n + 1);
    for (int64_t i = 0; i <= n; i++) rel[i] = off[i] - off[0];
    raw.upload(seq + off[0], (size_t)s->total_bases);
    doff.upload(rel.data(), (size_t)n + 1);
    KL_LAUNCH(pack_kernel, (unsigned)((words + 255) / 256), 256, 0, raw.p, doff.p, s->blk.p, n, words,
              s->bits2.p, s->inv16.p);
    sync_stream();
  } else {
    sync_stream();
  }
  return s;
}

static std::shared_ptr<Matrix> apply_features(Matrix &cls, const int32_t *features, int64_t nf) {
  for (int64_t j = 0; j < nf; j++) {
    KL_REQUIRE(features[2 * j] >= 0 && features[2 * j] < cls.m && features[2 * j + 1] >= 0 && features[2 * j + 1] < cls.m,
               "features: class index out of range");
  }
  auto out = std::make_shared<Matrix>();
  out->n = cls.n; out->m = nf; out->vt = cls.vt;
  out->class_k = cls.class_k; out->class_code = cls.class_code;
  out->sharded = cls.sharded; out->n_global = cls.n_global;
  DevBuf<int32_t> dfeat((size_t)(2 * nf));
  dfeat.upload(features, (size_t)(2 * nf));
  DevBuf<uint32_t> cnt((size_t)(cls.n ? cls.n : 1));
  const uint32_t *val = cls.vt == VAL_U32 ? cls.val_u32.p : nullptr;
  unsigned grid = (unsigned)((cls.n * 32 + 127) / 128);
  out->rowptr.alloc((size_t)cls.n + 1);
  if (cls.n > 0) {
    KL_LAUNCH((features_rows<false>), grid, 128, 0, cls.rowptr.p, cls.col.p, val, cls.n, dfeat.p, nf, cnt.p, nullptr,
              nullptr, nullptr);
    exclusive_scan_u32_to_i64(cnt.p, out->rowptr.p, cls.n);
    KL_CUDA(cudaMemcpyAsync(&out->nnz, out->rowptr.p + cls.n, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx().stream));
    sync_stream();
  } else {
    out->rowptr.zero(); out->nnz = 0;
  }
  out->col.alloc((size_t)(out->nnz ? out->nnz : 1));
  if (out->vt == VAL_U32) out->val_u32.alloc((size_t)(out->nnz ? out->nnz : 1));
  if (cls.n > 0)
    KL_LAUNCH((features_rows<true>), grid, 128, 0, cls.rowptr.p, cls.col.p, val, cls.n, dfeat.p, nf, nullptr,
              out->rowptr.p, out->col.p, out->vt == VAL_U32 ? out->val_u32.p : nullptr);
  sync_stream();
  return out;
}

std::shared_ptr<Matrix> extract(const kmerlr_config &cfg, const SeqSet &s, const int32_t *frozen_k,
                                const uint64_t *frozen_code, int64_t n_frozen, const int32_t *features,
                                int64_t n_features, int flags) {
  require_ready();
  KL_REQUIRE(cfg.alphabet == 0, "only the nucleotide alphabet is implemented on the GPU path (gapped: SURVEY 8f-3)");
  KL_REQUIRE(cfg.M >= 1 && cfg.M <= cfg.N, "need 1 <= M <= N");
  KL_REQUIRE(cfg.N <= MAX_N, "k-mer length above 13 is not supported on the GPU path");
  int nops = (cfg.complement != 0) + (cfg.reverse != 0) + (cfg.revcomp != 0);
  KL_REQUIRE(nops <= 1, "at most one of complement / reverse / revcomp is supported on the GPU path");
  KL_REQUIRE(n_features == 0 || n_frozen > 0, "an explicit feature list needs a frozen class list");
  const bool sharded = (flags & KMERLR_FLAG_SHARDED) != 0 && ctx().world > 1;

  XParams P{};
  P.M = cfg.M; P.N = cfg.N; P.binarize = cfg.binarize != 0;
  P.op = cfg.revcomp ? 1 : (cfg.complement ? 2 : (cfg.reverse ? 3 : 0));
  P.ks_lo = cfg.M; P.ks_hi = cfg.N < KS_MAX ? cfg.N : KS_MAX;
  P.big_lo = cfg.M > KS_MAX + 1 ? cfg.M : KS_MAX + 1;
  P.n = s.n;
  P.len = s.len.p; P.blk = s.blk.p; P.bits2 = s.bits2.p; P.inv16 = s.inv16.p;
  uint64_t dense = 0;
  for (int k = cfg.M; k <= cfg.N; k++) { P.level_off[k] = (uint32_t)dense; dense += 1ull << (2 * k); }
  P.level_off[cfg.N + 1] = (uint32_t)dense;
  const int64_t nbits = (int64_t)dense, nw = (nbits + 31) / 32;
  // staging stride: upper bound on distinct classes of one sequence
  int64_t stride = 0;
  for (int k = cfg.M; k <= cfg.N; k++) {
    int64_t inst = s.max_len - k + 1; if (inst < 0) inst = 0;
    int64_t cls = (int64_t)1 << (2 * k);
    stride += inst < cls ? inst : cls;
  }
  if (stride < 1) stride = 1;
  P.stride = stride;
  KL_REQUIRE(s.max_len < 65536 || P.ks_lo > P.ks_hi, "sequences of 65536 bp or more need M > 5 on the GPU path");
  // keys per lane
  int E = 0;
  if (P.big_lo <= cfg.N) {
    int64_t steps = (s.max_len + cfg.N - 1 + 31) / 32;
    int64_t slots = P.op ? 2 * steps : steps;
    E = 2; while (E < slots && E < 128) E <<= 1;
    KL_REQUIRE(E <= 64, "sequence too long for the register sort path (k > 5 needs L <= ~1000 bp with an "
                        "equivalence flag, ~2000 bp without)");
  }
  DevBuf<uint32_t> st_id((size_t)(s.n ? s.n * stride : 1));
  DevBuf<uint32_t> st_cnt((size_t)(P.binarize ? 1 : (s.n ? s.n * stride : 1)));
  DevBuf<uint32_t> rowcnt((size_t)(s.n ? s.n : 1));
  DevBuf<uint32_t> bitmap((size_t)nw);
  bitmap.zero();
  P.st_id = st_id.p; P.st_cnt = st_cnt.p; P.rowcnt = rowcnt.p; P.bitmap = bitmap.p;
  P.mark = n_frozen == 0;

  if (s.n > 0) {
    switch (E) {
      case 0: launch_extract<0>(P); break;
      case 2: launch_extract<2>(P); break;
      case 4: launch_extract<4>(P); break;
      case 8: launch_extract<8>(P); break;
      case 16: launch_extract<16>(P); break;
      case 32: launch_extract<32>(P); break;
      default: launch_extract<64>(P); break;
    }
  }
  // class set: observed union (all ranks) or the frozen list
  if (n_frozen > 0) {
    std::vector<uint32_t> hb((size_t)nw, 0u);
    uint64_t prev = 0;
    for (int64_t j = 0; j < n_frozen; j++) {
      int k = frozen_k[j];
      KL_REQUIRE(k >= cfg.M && k <= cfg.N && frozen_code[j] < (1ull << (2 * k)), "frozen class outside [M,N]");
      uint64_t id = (uint64_t)P.level_off[k] + frozen_code[j];
      KL_REQUIRE(j == 0 || id > prev, "frozen class list must be sorted by (k, code) without duplicates");
      prev = id;
      hb[id >> 5] |= 1u << (id & 31);
    }
    bitmap.upload(hb.data(), (size_t)nw);
    sync_stream();
  } else if (sharded) {
    DevBuf<uint8_t> bytes((size_t)nbits);
    KL_LAUNCH(bitmap_to_bytes, (unsigned)((nbits + 255) / 256), 256, 0, bitmap.p, nbits, bytes.p);
    comm_allreduce_max_u8(bytes.p, nbits);
    KL_LAUNCH(bytes_to_bitmap, (unsigned)((nw + 255) / 256), 256, 0, bytes.p, nbits, bitmap.p, nw);
    sync_stream();
  }
  DevBuf<uint32_t> pc((size_t)nw), rank((size_t)nw + 1);
  KL_LAUNCH(popc_words, (unsigned)((nw + 255) / 256), 256, 0, bitmap.p, nw, pc.p);
  exclusive_scan_u32(pc.p, rank.p, nw);
  uint32_t m32 = 0;
  KL_CUDA(cudaMemcpyAsync(&m32, rank.p + nw, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx().stream));
  sync_stream();

  auto out = std::make_shared<Matrix>();
  out->n = s.n; out->m = (int64_t)m32; out->vt = P.binarize ? VAL_ONE : VAL_U32;
  out->sharded = sharded; out->n_global = s.n;
  if (sharded) {
    DevBuf<int64_t> tmp(1);
    int64_t nn = s.n;
    tmp.upload(&nn, 1);
    comm_allreduce_sum_i64(tmp.p, 1);
    tmp.download(&nn, 1);
    sync_stream();
    out->n_global = nn;
  }
  // class list
  {
    DevBuf<uint32_t> ids((size_t)(m32 ? m32 : 1));
    KL_LAUNCH(enumerate_bits, (unsigned)((nw + 255) / 256), 256, 0, bitmap.p, rank.p, nw, ids.p);
    std::vector<uint32_t> hid((size_t)m32);
    ids.download(hid.data(), (size_t)m32);
    sync_stream();
    out->class_k.resize((size_t)m32); out->class_code.resize((size_t)m32);
    int k = cfg.M;
    for (size_t j = 0; j < (size_t)m32; j++) {
      while (k < cfg.N && hid[j] >= P.level_off[k + 1]) k++;
      out->class_k[j] = k; out->class_code[j] = hid[j] - P.level_off[k];
    }
  }
  // final CSR
  out->rowptr.alloc((size_t)s.n + 1);
  DevBuf<uint32_t> kept;
  const uint32_t *cnt_final = rowcnt.p;
  unsigned wgrid = (unsigned)((s.n * 32 + 127) / 128);
  if (s.n > 0) {
    if (n_frozen > 0) {
      kept.alloc((size_t)s.n);
      KL_LAUNCH(count_kept, wgrid, 128, 0, st_id.p, rowcnt.p, stride, s.n, bitmap.p, kept.p);
      cnt_final = kept.p;
    }
    exclusive_scan_u32_to_i64(cnt_final, out->rowptr.p, s.n);
    KL_CUDA(cudaMemcpyAsync(&out->nnz, out->rowptr.p + s.n, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx().stream));
    sync_stream();
  } else {
    out->rowptr.zero(); out->nnz = 0;
  }
  out->col.alloc((size_t)(out->nnz ? out->nnz : 1));
  if (out->vt == VAL_U32) out->val_u32.alloc((size_t)(out->nnz ? out->nnz : 1));
  if (s.n > 0) {
    uint32_t *vp = out->vt == VAL_U32 ? out->val_u32.p : nullptr;
    if (n_frozen > 0)
      KL_LAUNCH((compact_rows<true>), wgrid, 128, 0, st_id.p, st_cnt.p, rowcnt.p, stride, s.n, bitmap.p, rank.p,
                out->rowptr.p, out->col.p, vp);
    else
      KL_LAUNCH((compact_rows<false>), wgrid, 128, 0, st_id.p, st_cnt.p, rowcnt.p, stride, s.n, bitmap.p, rank.p,
                out->rowptr.p, out->col.p, vp);
  }
  sync_stream();
  if (n_features > 0) return apply_features(*out, features, n_features);
  return out;
}

}  // namespace kl
