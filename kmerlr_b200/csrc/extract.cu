// extract.cu -- stage 1 of the hot path: sequences -> sparse k-mer count rows.
//
// Replaces scan_sequences + NewKmerCountsList/SetKmers + convert_counts_list
// (kmerLr_data.go:197-284,306-357; k-mer semantics of gonetics restated in SURVEY.md 8c).
//
// Design (one warp per sequence, no hash tables, no float atomics).  Sequences are 2-bit packed in
// HBM (+1 invalid bit per base, see SeqSet); every lane rolls the forward code and the image under
// the strand operation (revcomp / complement / reverse) over its stretch of positions, so the
// canonical code min(u, op(u)) of the k-mer starting at a position costs a shift and a min.  Levels:
//   * k <= 5   "table levels": forward counts of the deepest table level in a per-warp shared
//              memory table (one atomic per position), shallower levels are sums of 4 children; the
//              classes are walked through a precomputed list (code, image, id);
//   * k = 6, 7 "bitmap levels": one bit per canonical code in a per-warp shared memory bitmap
//              (4^k bits), set with atomicOr; the rare repeats go to a small list and are added to
//              the counts afterwards.  Walking the bitmap emits the row in code order.  Level 8 is a
//              bitmap level too when the rows are too long for the register sort;
//   * k >= 8   "sorted levels": per level ONE bitonic sort of the canonical codes in registers --
//              in-register compare-exchanges below distance E, warp shuffles above; a run of equal
//              keys is a class, its length the count.
// Rows leave the kernel sorted by (k, code) = by final column index; columns are ranks in the
// bitmap of observed (or frozen) classes; observed classes are gathered per block in shared memory.
// Host sequences (kmerlr_extract) arrive in chunks on a copy stream while earlier chunks are
// packed and extracted.
#include "common.cuh"

#include <type_traits>

namespace kl {

namespace {

constexpr uint32_t SENT = 0xFFFFFFFFu;    // sort key of a position without a valid k-mer (sorts last)
constexpr uint32_t NOKEY = 0xFFFFFFFEu;   // "no previous key" in front of the sorted array
constexpr int KT_MAX = 5;          // levels <= KT_MAX use direct count tables
constexpr int KB_MAX = 8;          // highest level that can use a per-row bitmap
constexpr int KB_DEFAULT = 7;      // ... level 8 does so only when the rows are too long for the register sort
constexpr int MAX_N = 13;          // 2*13 code bits + 4 length bits per position
constexpr int TAB_WORDS = 688;     // (4+16+64+256+1024)/2 = 682 packed u16 pairs, padded
constexpr int DUP_SMEM = 108;      // repeats of the bitmap levels kept in shared memory (the rest: global list)
constexpr int OBS_MAX_LEVEL = 8;   // marks of observed classes of levels <= 8 are gathered per block in smem

struct XParams {
  int M, N, op, binarize;
  int t_lo, t_hi;                  // table levels (empty if t_lo > t_hi)
  int b_lo, b_hi;                  // bitmap levels (empty if b_lo > b_hi)
  int s_lo;                        // sorted levels [s_lo, N] (empty if s_lo > N)
  int mark;                        // mark observed classes in the bitmap
  uint32_t level_off[MAX_N + 2];   // dense id of (k, code 0), multiples of 32
  uint32_t bm_off[KB_MAX + 2];     // word offset of level k's bitmap inside the per-warp bitmap area
  uint32_t pf_off[KB_MAX + 2];     // word offset of level k's prefix array (one entry per 4 words)
  uint32_t tl_cnt;                 // classes of the table levels (flat list tl, in (k, code) order)
  int bm_words, pf_words, ts_words;// per-warp shared memory areas, in 32-bit words
  int warp_words;                  // total per-warp shared memory, in 32-bit words
  int obs_words;                   // per-block bitmap of observed classes (levels <= OBS_MAX_LEVEL)
  int64_t stride, n, row0, ovf_stride;   // the launch covers the rows [row0, n)
  const int64_t *len, *blk;
  const uint32_t *bits2;
  const uint16_t *inv16;
  const uint2 *tl;                 // x = table index of the code | table index of its image << 16, y = class id
  uint32_t *st_id, *st_cnt, *rowcnt, *bitmap, *ovf;
  const uint2 *nbx;                // numbering set (all classes of the configuration, or the frozen list)
  int filter;                      // drop ids outside the numbering set
  unsigned long long *stats;       // [0] max_i sum_j v_ij^2, [1] max v_ij
};

__device__ __forceinline__ uint32_t tab_off(int k) {  // sum_{j=1}^{k-1} 4^j
  return ((1u << (2 * k)) - 4u) / 3u;
}

// ---- pack: ASCII -> 2 bit codes + invalid mask --------------------------------------------------
__global__ void pack_kernel(const uint8_t *__restrict__ seq, const int64_t *__restrict__ off,
                            const int64_t *__restrict__ blk, int64_t n, int64_t w0, int64_t w1,
                            uint32_t *__restrict__ bits2, uint16_t *__restrict__ inv16) {
  int64_t w = w0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= w1) return;
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

// ---- bitonic sort of 32*E keys in registers, index = lane*E + r, ascending ----------------------
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
  int L, cur_word;
  uint32_t w, iv;
  __device__ __forceinline__ void init(const uint32_t *b, const uint16_t *m, int len) {
    bits2 = b; inv16 = m; L = len; cur_word = -1; w = 0; iv = 0;
  }
  // base idx -> (code, invalid); out of range = invalid
  __device__ __forceinline__ void get(int idx, uint32_t &x, uint32_t &inv) {
    if ((unsigned)idx >= (unsigned)L) { x = 0; inv = 1; return; }
    int wi = idx >> 4;
    if (wi != cur_word) { cur_word = wi; w = __ldg(bits2 + wi); iv = __ldg(inv16 + wi); }
    int sh = idx & 15;
    x = (w >> (2 * sh)) & 3u;
    inv = (iv >> sh) & 1u;
  }
};

struct Emitter {
  uint32_t *sid, *scnt;            // this row's slices of the column / count arrays
  uint32_t *obs, *bitmap;          // observed classes: per-block (shared) and global bitmap
  const uint2 *nbx;                // numbering set, per 32 ids: x = member bits, y = column of the first member
  uint32_t obs_bits;               // ids below this are marked in obs
  uint32_t cursor;
  int binarize, mark, filter;      // filter: ids outside the numbering set are dropped (frozen class list)
  unsigned long long sq;           // per lane: sum of squared counts / largest count emitted
  uint32_t vm;
  // mark one class id as observed
  __device__ __forceinline__ void mark_id(uint32_t id) {
    if (!mark) return;
    const uint32_t bit = 1u << (id & 31), w = id >> 5;
    if (id < obs_bits) { if (!(obs[w] & bit)) atomicOr(obs + w, bit); }
    else if (!(bitmap[w] & bit)) atomicOr(bitmap + w, bit);
  }
  // OR a word of 32 consecutive class ids (a level <= OBS_MAX_LEVEL) into the block's bitmap
  __device__ __forceinline__ void mark_word(uint32_t word_index, uint32_t bits) {
    if (mark && (bits & ~obs[word_index])) atomicOr(obs + word_index, bits);
  }
  // column of a class id; false when the id is not in the numbering set
  __device__ __forceinline__ bool column(uint32_t id, uint32_t &col) const {
    const uint2 e = __ldg(nbx + (id >> 5));
    const uint32_t bit = 1u << (id & 31);
    col = e.y + __popc(e.x & (bit - 1u));
    return (e.x & bit) != 0u;
  }
  __device__ __forceinline__ void stat(uint32_t cnt) {
    sq += (unsigned long long)cnt * cnt;
    vm = max(vm, cnt);
  }
  // ballot-compacted emission (coalesced stores); col_known >= 0: the column when no class is dropped
  __device__ __forceinline__ void emit(bool flag, uint32_t id, uint32_t cnt, int64_t col_known = -1) {
    uint32_t col = (uint32_t)col_known;
    if (flag && (filter || col_known < 0)) { const bool member = column(id, col); if (filter) flag = member; }
    unsigned em = __ballot_sync(0xffffffffu, flag);
    if (flag) {
      uint32_t pos = cursor + __popc(em & lanemask_lt());
      sid[pos] = col;
      if (!binarize) scnt[pos] = cnt;
      mark_id(id);
      stat(binarize ? 1u : cnt);
    }
    cursor += __popc(em);
  }
};

__host__ __device__ constexpr int ext_threads(int E) { return E <= 16 ? 512 : (E == 32 ? 256 : 128); }

// ---- sorted level: the canonical codes of one level, sorted in registers -----------------------------
// K[r] sits at sorted index lane*E + r.  A run (equal keys) is one class, its length the count; a run
// is emitted by the lane holding the position right after it.  Ids and counts go through the warp's
// shared-memory buffer so that the global stores are coalesced.
template <int E>
__device__ __forceinline__ void emit_sorted(const uint32_t (&K)[E], unsigned lane, uint32_t *sbuf, Emitter &em,
                                            uint32_t idbase) {
  using MaskT = typename std::conditional<(E <= 32), uint32_t, unsigned long long>::type;
  uint32_t prevlast = __shfl_up_sync(0xffffffffu, K[E - 1], 1);
  if (lane == 0) prevlast = NOKEY;
  // bit r of bnd: a run starts at this lane's key r (sorted index lane*E + r)
  MaskT bnd = (MaskT)(K[0] != prevlast);
#pragma unroll
  for (int r = 1; r < E; r++) bnd |= (MaskT)(K[r] != K[r - 1]) << r;
  // a run is emitted by the position right after it: every boundary but sorted index 0
  const MaskT cls = lane == 0 ? (bnd & ~(MaskT)1) : bnd;
  const int c = sizeof(MaskT) == 4 ? __popc((uint32_t)cls) : __popcll(cls);
  const int lastb = bnd ? (int)lane * E + (sizeof(MaskT) == 4 ? 31 - __clz((uint32_t)bnd) : 63 - __clzll(bnd)) : -1;
  int incl = c, mx = lastb;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, incl, o), z = __shfl_up_sync(0xffffffffu, mx, o);
    if (lane >= (unsigned)o) { incl += y; mx = max(mx, z); }
  }
  const int base = incl - c, total = __shfl_sync(0xffffffffu, incl, 31);
  int start0 = __shfl_up_sync(0xffffffffu, mx, 1);   // last boundary before this lane's keys
  if (lane == 0) start0 = 0;
  // round 1: ids
#pragma unroll
  for (int r = 0; r < E; r++) {
    if ((cls >> r) & 1) {
      const MaskT below = cls & (((MaskT)1 << r) - 1);
      const int j = base + (sizeof(MaskT) == 4 ? __popc((uint32_t)below) : __popcll(below));
      sbuf[j] = idbase + (r == 0 ? prevlast : K[r > 0 ? r - 1 : 0]);
    }
  }
  __syncwarp();
  if (!em.filter) {
    // no class is dropped: slot i of the buffer is entry i of the level
    for (int i = lane; i < total; i += 32) {
      const uint32_t id = sbuf[i];
      uint32_t col;
      em.column(id, col);
      em.sid[em.cursor + i] = col;
      em.mark_id(id);
    }
    __syncwarp();
    if (!em.binarize) {
#pragma unroll
      for (int r = 0; r < E; r++) {
        if ((cls >> r) & 1) {
          const MaskT below = cls & (((MaskT)1 << r) - 1), bbelow = bnd & (((MaskT)1 << r) - 1);
          const int j = base + (sizeof(MaskT) == 4 ? __popc((uint32_t)below) : __popcll(below));
          const int st = bbelow ? (int)lane * E + (sizeof(MaskT) == 4 ? 31 - __clz((uint32_t)bbelow) : 63 - __clzll(bbelow))
                                : start0;
          const uint32_t cnt = (uint32_t)((int)lane * E + r - st);
          sbuf[j] = cnt;
          em.stat(cnt);
        }
      }
      __syncwarp();
      for (int i = lane; i < total; i += 32) em.scnt[em.cursor + i] = sbuf[i];
      __syncwarp();
    } else {
      em.sq += (unsigned long long)c; if (c) em.vm = max(em.vm, 1u);
    }
    em.cursor += (uint32_t)total;
    return;
  }
  // copy out: id -> column; with a frozen class list the ids outside the list are dropped here and the
  // keep decisions (one bit per copy iteration) are replayed for the counts
  unsigned long long keepbits = 0;
  uint32_t kept = 0;
  for (int i0 = 0, it = 0; i0 < total; i0 += 32, it++) {
    const int i = i0 + (int)lane;
    bool keep = false; uint32_t col = 0, id = 0;
    if (i < total) { id = sbuf[i]; keep = em.column(id, col) || !em.filter; }
    const unsigned km = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      em.sid[em.cursor + kept + __popc(km & lanemask_lt())] = col;
      em.mark_id(id);
      keepbits |= 1ull << it;
    }
    kept += __popc(km);
  }
  __syncwarp();
  // round 2: counts = distance to the previous boundary
  if (!em.binarize) {
#pragma unroll
    for (int r = 0; r < E; r++) {
      if ((cls >> r) & 1) {
        const MaskT below = cls & (((MaskT)1 << r) - 1), bbelow = bnd & (((MaskT)1 << r) - 1);
        const int j = base + (sizeof(MaskT) == 4 ? __popc((uint32_t)below) : __popcll(below));
        const int st = bbelow ? (int)lane * E + (sizeof(MaskT) == 4 ? 31 - __clz((uint32_t)bbelow) : 63 - __clzll(bbelow))
                              : start0;
        sbuf[j] = (uint32_t)((int)lane * E + r - st);
      }
    }
    __syncwarp();
    uint32_t k2 = 0;
    for (int i0 = 0, it = 0; i0 < total; i0 += 32, it++) {
      const int i = i0 + (int)lane;
      const bool keep = (keepbits >> it) & 1ull;
      const unsigned km = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const uint32_t cnt = sbuf[i];
        em.scnt[em.cursor + k2 + __popc(km & lanemask_lt())] = cnt;
        em.stat(cnt);
      }
      k2 += __popc(km);
    }
    __syncwarp();
  } else {
    em.sq += __popcll(keepbits); if (keepbits) em.vm = max(em.vm, 1u);
  }
  em.cursor += kept;
}


// ---- the extraction kernel: one warp per sequence ------------------------------------------------
// E = keys per lane of the register sort (0: no sorted levels, sequences of any length)
template <int E>
__global__ void __launch_bounds__(ext_threads(E), 2) extract_kernel(const XParams P) {
  constexpr int EE = E > 0 ? E : 1;
  extern __shared__ __align__(16) uint32_t smem[];
  const unsigned lane = lane_id();
  const int warp_in_block = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  // block: [observed classes]; per warp: [bitmaps][prefix][count tables | sort buffer][dup count + dups]
  uint32_t *obs = smem;
  uint32_t *bm = smem + P.obs_words + (size_t)warp_in_block * P.warp_words;
  uint32_t *pf = bm + P.bm_words;
  uint32_t *tab = pf + P.pf_words;
  uint32_t *dupn = tab + P.ts_words;
  uint32_t *dups = dupn + 4;
  const int64_t gwarp = (int64_t)blockIdx.x * wpb + warp_in_block, nwarps = (int64_t)gridDim.x * wpb;
  uint32_t *ovf = P.ovf ? P.ovf + gwarp * P.ovf_stride : nullptr;
  const int N = P.N, M = P.M, op = P.op;
  const bool two = op != 0;
  const uint32_t maskN = (1u << (2 * N)) - 1u;
  const uint32_t maskNb = (1u << N) - 1u;
  const bool has_tab = P.t_lo <= P.t_hi, has_bm = P.b_lo <= P.b_hi, has_sort = E > 0 && P.s_lo <= N;

  // shared memory starts clean; every row leaves its bitmaps clean again
  for (int i = threadIdx.x; i < P.obs_words; i += blockDim.x) obs[i] = 0;
  for (int i = lane; i < P.bm_words; i += 32) bm[i] = 0;
  if (lane == 0) dupn[0] = 0;
  __syncthreads();

  for (int64_t row = P.row0 + gwarp; row < P.n; row += nwarps) {
    const int L = (int)P.len[row];
    const uint32_t *b2 = P.bits2 + P.blk[row] * 4;
    const uint16_t *iv16 = P.inv16 + P.blk[row] * 4;
    if (has_tab)
      for (int i = lane; i < TAB_WORDS; i += 32) tab[i] = 0;
    __syncwarp();

    // ---- pass over the positions: lane handles the start positions p = lane, lane + 32, ... ----------
    // The N bases starting at p are 2N consecutive bits of the packed words (first base in the low
    // bits): one funnel shift.  Read that way the k-mer starting at p is the LOW 2k bits, its
    // reverse complement (first base most significant, the code order) is the complement of those
    // bits and its forward code is their digit reversal -- no rolling state, no warm-up.
    const int steps = (L + 31) / 32;
    uint32_t FWL[EE];                 // (forward code << 4) | valid length, per position of this lane
#pragma unroll
    for (int r = 0; r < EE; r++) FWL[r] = 0;
    {
      auto visit = [&](int p) -> uint32_t {
        const int wi = p >> 4, sh = p & 15;
        const uint32_t e = __funnelshift_r(__ldg(b2 + wi), __ldg(b2 + wi + 1), 2 * sh) & maskN;
        const uint32_t ivb = (((uint32_t)__ldg(iv16 + wi) | ((uint32_t)__ldg(iv16 + wi + 1) << 16)) >> sh) |
                             (1u << min(N, L - p));       // the sequence ends: no k-mer reaches past L
        const int len_f = __ffs(ivb) - 1;                 // valid bases from p on (<= N)
        const uint32_t FW = swap_pairs(__brev(e)) >> (32 - 2 * N);
        const uint32_t S2 = op == 1 ? (~e) & maskN : e;   // image of the window under revcomp / reverse
        if (len_f >= M) {
          // table levels: one count at the deepest table level this suffix reaches
          if (has_tab) {
            int kk = len_f < P.t_hi ? len_f : P.t_hi;
            if (kk >= P.t_lo) {
              uint32_t idx = tab_off(kk) + (FW >> (2 * (N - kk)));
              atomicAdd(tab + (idx >> 1), 1u << (16 * (idx & 1)));
            }
          }
          // bitmap levels: canonical code = min(prefix, image)
          if (has_bm) {
            for (int k = P.b_lo; k <= P.b_hi; k++) {
              if (len_f < k) break;
              const uint32_t mk = (1u << (2 * k)) - 1u;
              uint32_t c = FW >> (2 * (N - k));
              if (op == 1 || op == 3) c = min(c, S2 & mk);
              else if (op == 2) c = min(c, (~c) & mk);
              const uint32_t bit = 1u << (c & 31);
              const uint32_t old = atomicOr(bm + P.bm_off[k] + (c >> 5), bit);
              if ((old & bit) && !P.binarize) {
                const uint32_t at = atomicAdd(dupn, 1u), ee = ((uint32_t)k << 26) | c;
                if (at < (uint32_t)DUP_SMEM) dups[at] = ee; else ovf[at - DUP_SMEM] = ee;
              }
            }
          }
        }
        return (FW << 4) | (uint32_t)len_f;
      };
      if (E > 0) {
#pragma unroll
        for (int i = 0; i < EE; i++) {
          const int p = i * 32 + (int)lane;
          if (p < L) FWL[i] = visit(p);
        }
      } else {
        for (int i = 0; i < steps; i++) {
          const int p = i * 32 + (int)lane;
          if (p < L) visit(p);
        }
      }
    }
    __syncwarp();

    Emitter em;
    em.sid = P.st_id + row * P.stride; em.scnt = P.st_cnt + (P.binarize ? 0 : row * P.stride);
    em.obs = obs; em.bitmap = P.bitmap; em.obs_bits = (uint32_t)P.obs_words * 32u;
    em.nbx = P.nbx; em.filter = P.filter;
    em.cursor = 0; em.binarize = P.binarize; em.mark = P.mark;
    em.sq = 0; em.vm = 0;

    // ---- table levels ---------------------------------------------------------------------------
    if (has_tab) {
      // a k-mer count is the sum of its 4 extensions plus the suffixes that end right after it
      for (int k = P.t_hi - 1; k >= P.t_lo; k--) {
        const uint32_t nk = 1u << (2 * k), toff = tab_off(k), coff = tab_off(k + 1);
        for (uint32_t u = lane; u < nk; u += 32) {
          uint32_t c0 = coff + 4 * u;                 // 4 children = two aligned words
          uint32_t w0 = tab[c0 >> 1], w1 = tab[(c0 >> 1) + 1];
          uint32_t sum = (w0 & 0xFFFFu) + (w0 >> 16) + (w1 & 0xFFFFu) + (w1 >> 16);
          uint32_t idx = toff + u;
          if (sum) atomicAdd(tab + (idx >> 1), sum << (16 * (idx & 1)));
        }
        __syncwarp();
      }
      // walk the classes of the table levels in (k, code) order: count = code + image
      for (uint32_t base = 0; base < P.tl_cnt; base += 32) {
        const uint32_t j = base + lane;
        uint32_t id = 0, cnt = 0;
        if (j < P.tl_cnt) {
          const uint2 e = __ldg(P.tl + j);
          const uint32_t i1 = e.x & 0xFFFFu, i2 = e.x >> 16;
          cnt = (tab[i1 >> 1] >> (16 * (i1 & 1))) & 0xFFFFu;
          if (i2 != i1) cnt += (tab[i2 >> 1] >> (16 * (i2 & 1))) & 0xFFFFu;
          id = e.y;
        }
        // without a frozen list the numbering set is ALL classes: the j-th class of the list is column j
        em.emit(cnt > 0, id, cnt, P.filter ? -1 : (int64_t)j);
      }
      __syncwarp();
    }

    // ---- bitmap levels: walk the bitmap 128 words at a time (4 consecutive words per lane) ---------
    if (has_bm) {
      for (int k = P.b_lo; k <= P.b_hi; k++) {
        const int W = 1 << (2 * k - 5);
        uint32_t *bk = bm + P.bm_off[k];
        uint32_t *pk = pf + P.pf_off[k];
        const uint32_t idbase = P.level_off[k], gword = P.level_off[k] >> 5;
        for (int it = 0; it * 128 < W; it++) {
          const int wi = it * 128 + (int)lane * 4;
          uint4 w4 = make_uint4(0, 0, 0, 0);
          if (wi < W) w4 = *reinterpret_cast<const uint4 *>(bk + wi);
          // numbering of the lane's 4 words: member bits and column of the first member (coalesced loads)
          uint2 nx[4];
#pragma unroll
          for (int j = 0; j < 4; j++) nx[j] = wi < W ? __ldg(P.nbx + gword + wi + j) : make_uint2(0u, 0u);
          if (P.filter && wi < W) {
            // frozen class list: classes outside the list vanish here, before any slot is assigned
            w4.x &= nx[0].x; w4.y &= nx[1].x; w4.z &= nx[2].x; w4.w &= nx[3].x;
            *reinterpret_cast<uint4 *>(bk + wi) = w4;
          }
          if (w4.x | w4.y | w4.z | w4.w) {
            em.mark_word(gword + wi, w4.x); em.mark_word(gword + wi + 1, w4.y);
            em.mark_word(gword + wi + 2, w4.z); em.mark_word(gword + wi + 3, w4.w);
          }
          const uint32_t c = __popc(w4.x) + __popc(w4.y) + __popc(w4.z) + __popc(w4.w);
          uint32_t incl = c;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += y;
          }
          const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
          if (wi < W) pk[wi >> 2] = em.cursor + incl - c;   // slot of this lane's first class
          const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
          if (tot <= (uint32_t)P.ts_words) {
            // ids through the warp's shared buffer: coalesced global stores of the columns
            uint32_t pos = incl - c;
#pragma unroll
            for (int j = 0; j < 4; j++) {
              uint32_t w = ws[j];
              while (w) {
                const uint32_t below = (w & (0u - w)) - 1u;            // bits under the lowest set bit
                tab[pos++] = nx[j].y + __popc(nx[j].x & below);         // its column
                w &= w - 1;
              }
            }
            __syncwarp();
            for (uint32_t i = lane; i < tot; i += 32) {
              em.sid[em.cursor + i] = tab[i];
              if (!P.binarize) em.scnt[em.cursor + i] = 1;
            }
            __syncwarp();
          } else {
            uint32_t pos = em.cursor + incl - c;
#pragma unroll
            for (int j = 0; j < 4; j++) {
              uint32_t w = ws[j];
              while (w) {
                const uint32_t below = (w & (0u - w)) - 1u;
                w &= w - 1;
                em.sid[pos] = nx[j].y + __popc(nx[j].x & below);
                if (!P.binarize) em.scnt[pos] = 1;
                pos++;
              }
            }
          }
          em.sq += c; if (c) em.vm = max(em.vm, 1u);        // every class of the chunk enters with count 1
          em.cursor += tot;
        }
      }
      __syncwarp();
      __threadfence_block();
      // repeats: add one to the count of the class, found by its rank in the bitmap
      if (!P.binarize) {
        const uint32_t nd = dupn[0];
        for (uint32_t i = lane; i < nd; i += 32) {
          const uint32_t e = i < (uint32_t)DUP_SMEM ? dups[i] : ovf[i - DUP_SMEM];
          const uint32_t k = e >> 26, c = e & 0x3FFFFFFu, wi = c >> 5;
          const uint32_t *bk = bm + P.bm_off[k];
          if (!((bk[wi] >> (c & 31)) & 1u)) continue;       // class dropped by the frozen list
          uint32_t slot = pf[P.pf_off[k] + (wi >> 2)] + __popc(bk[wi] & ((1u << (c & 31)) - 1u));
          for (uint32_t j = wi & ~3u; j < wi; j++) slot += __popc(bk[j]);
          const uint32_t old = atomicAdd(em.scnt + slot, 1u);
          em.sq += 2ull * old + 1ull;                       // (old+1)^2 - old^2
          em.vm = max(em.vm, old + 1u);
        }
        __syncwarp();
        if (lane == 0) dupn[0] = 0;
      }
      for (int i = (int)lane * 4; i < P.bm_words; i += 128) *reinterpret_cast<uint4 *>(bm + i) = make_uint4(0, 0, 0, 0);
      __syncwarp();
    }

    // ---- sorted levels: one register sort of the canonical codes per level ---------------------------
    if (has_sort) {
      for (int k = P.s_lo; k <= N; k++) {
        uint32_t K[EE];
#pragma unroll
        for (int r = 0; r < EE; r++) {
          const uint32_t f = FWL[r];
          uint32_t c = (f >> 4) >> (2 * (N - k));
          if (two) c = min(c, kmer_op(c, k, op));
          K[r] = (int)(f & 15u) >= k ? c : SENT;
        }
        warp_sort<EE>(K, lane);
        emit_sorted<EE>(K, lane, tab, em, P.level_off[k]);
      }
    }
    // row statistics (exact integers): what the step size and the fixed-point scale of the gradient need
    {
      unsigned long long sq = em.sq; uint32_t vm = em.vm;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        vm = max(vm, __shfl_xor_sync(0xffffffffu, vm, o));
      }
      if (lane == 0) {
        P.rowcnt[row] = em.cursor;
        if (sq > P.stats[0]) atomicMax(P.stats, sq);
        if ((unsigned long long)vm > P.stats[1]) atomicMax(P.stats + 1, (unsigned long long)vm);
      }
    }
  }
  // flush the block's observed classes
  __syncthreads();
  if (P.mark)
    for (int i = threadIdx.x; i < P.obs_words; i += blockDim.x) {
      const uint32_t w = obs[i];
      if (w && (w & ~P.bitmap[i])) atomicOr(P.bitmap + i, w);
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

// ---- numbering set helpers --------------------------------------------------------------------------
struct LevelInfo {
  int M, N, op;
  uint32_t level_off[16];
};
// every class the configuration can produce: the codes u <= image(u) of every level
__global__ void all_classes_bits(const LevelInfo I, uint32_t *__restrict__ bm, int64_t nw) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  const uint32_t id0 = (uint32_t)(w * 32);
  int k = I.M;
  while (k < I.N && id0 >= I.level_off[k + 1]) k++;
  const uint32_t u0 = id0 - I.level_off[k], nk = 1u << (2 * k);   // levels start on multiples of 32
  uint32_t v = 0;
  for (int b = 0; b < 32; b++) {
    const uint32_t u = u0 + b;
    if (u < nk && (I.op == 0 || u <= kmer_op(u, k, I.op))) v |= 1u << b;
  }
  bm[w] = v;
}
__global__ void interleave_numbering(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ rank, int64_t nw,
                                     uint2 *__restrict__ out) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < nw) out[w] = make_uint2(bits[w], rank[w]);
}
__global__ void bitmaps_differ(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, int64_t nw,
                               uint32_t *__restrict__ flag) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < nw && a[w] != b[w]) atomicOr(flag, 1u);
}
// columns numbered in the numbering set -> ranks among the observed classes, in place (warp per row)
__global__ void renumber_rows(uint32_t *__restrict__ col, const uint32_t *__restrict__ rowcnt, int64_t stride, int64_t n,
                              const uint32_t *__restrict__ ids, const uint32_t *__restrict__ bm,
                              const uint32_t *__restrict__ rank) {
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const unsigned lane = lane_id();
  uint32_t *c = col + row * stride;
  const uint32_t cnt = rowcnt[row];
  for (uint32_t j = lane; j < cnt; j += 32) {
    const uint32_t id = __ldg(ids + c[j]), w = __ldg(bm + (id >> 5));
    c[j] = __ldg(rank + (id >> 5)) + __popc(w & ((1u << (id & 31)) - 1u));
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
__global__ void features_rows(const Rows R, const uint32_t *__restrict__ col,
                              const uint32_t *__restrict__ val, int64_t n, const int32_t *__restrict__ feat,
                              int64_t nf, uint32_t *__restrict__ cnt_out, const int64_t *__restrict__ orowptr,
                              uint32_t *__restrict__ ocol, uint32_t *__restrict__ oval) {
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  unsigned lane = lane_id();
  int64_t a, b;
  R.range(row, a, b);
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
  // warps per block: the choice that keeps the most warps resident per SM (the block shares one
  // bitmap of observed classes, every warp brings its own working set)
  KL_CUDA(cudaFuncSetAttribute(extract_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  int best_wpb = 0, best_per_sm = 0;
  for (int wpb = ext_threads(E) / 32; wpb >= 1; wpb--) {
    size_t smem = ((size_t)P.obs_words + (size_t)wpb * P.warp_words) * sizeof(uint32_t);
    if (smem > (size_t)227 * 1024) continue;
    int per_sm = 0;
    KL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, extract_kernel<E>, 32 * wpb, smem));
    if (per_sm * wpb > best_per_sm * best_wpb) { best_wpb = wpb; best_per_sm = per_sm; }
  }
  KL_REQUIRE(best_wpb > 0, "sequence too long for the shared-memory working set of one warp");
  const int wpb = best_wpb;
  size_t smem = ((size_t)P.obs_words + (size_t)wpb * P.warp_words) * sizeof(uint32_t);
  int64_t blocks = (int64_t)ctx().sm_count * best_per_sm;
  int64_t need = (P.n - P.row0 + wpb - 1) / wpb;
  if (blocks > need) blocks = need;
  if (blocks < 1) blocks = 1;
  KL_LAUNCH((extract_kernel<E>), (unsigned)blocks, 32 * wpb, smem, P);
}

// the classes of the table levels [t_lo, t_hi] in (k, code) order: table index of the code, table
// index of its image under the strand operation (the same index when the code is its own image), id
const uint2 *table_classes(int op, int t_lo, int t_hi, const uint32_t *level_off, uint32_t *count) {
  struct Entry { int op, lo, hi; uint32_t off0; DevBuf<uint2> dev; uint32_t cnt; };
  static std::vector<std::unique_ptr<Entry>> cache;
  for (auto &e : cache)
    if (e->op == op && e->lo == t_lo && e->hi == t_hi && e->off0 == level_off[t_lo]) { *count = e->cnt; return e->dev.p; }
  std::vector<uint2> list;
  for (int k = t_lo; k <= t_hi; k++) {
    const uint32_t toff = ((1u << (2 * k)) - 4u) / 3u;
    for (uint32_t u = 0; u < (1u << (2 * k)); u++) {
      uint32_t r = u;
      if (op) {
        r = 0;
        for (int i = 0; i < k; i++) {
          uint32_t d = (u >> (2 * i)) & 3u;               // digit i from the right
          if (op == 1) r |= (3u - d) << (2 * (k - 1 - i));
          else if (op == 2) r |= (3u - d) << (2 * i);
          else r |= d << (2 * (k - 1 - i));
        }
      }
      if (u <= r) list.push_back(make_uint2((toff + u) | ((toff + r) << 16), level_off[k] + u));
    }
  }
  auto e = std::make_unique<Entry>();
  e->op = op; e->lo = t_lo; e->hi = t_hi; e->off0 = level_off[t_lo]; e->cnt = (uint32_t)list.size();
  e->dev.alloc(list.size());
  e->dev.upload(list.data(), list.size());
  sync_stream();
  *count = e->cnt;
  cache.push_back(std::move(e));
  return cache.back()->dev.p;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// lengths and block offsets on the device, packed arrays allocated but not filled yet; the host copies
// of the block offsets and of the offsets relative to the first base come back for the feeding loop
static std::shared_ptr<SeqSet> sequences_alloc(const int64_t *off, int64_t n, std::vector<int64_t> &blk,
                                               std::vector<int64_t> &rel) {
  require_ready();
  KL_REQUIRE(n >= 0 && off != nullptr, "sequences: bad arguments");
  auto s = std::make_shared<SeqSet>();
  s->n = n;
  std::vector<int64_t> len((size_t)n);
  blk.assign((size_t)n + 1, 0); rel.assign((size_t)n + 1, 0);
  for (int64_t i = 0; i < n; i++) {
    len[i] = off[i + 1] - off[i];
    KL_REQUIRE(len[i] >= 0, "sequences: offsets must be non-decreasing");
    if (len[i] > s->max_len) s->max_len = len[i];
    blk[i + 1] = blk[i] + (len[i] + 63) / 64;
    rel[i + 1] = off[i + 1] - off[0];
  }
  s->total_bases = n ? off[n] - off[0] : 0;
  s->total_blocks = blk[n];
  s->len.alloc((size_t)(n ? n : 1));
  s->blk.alloc((size_t)n + 1);
  s->len.upload(len.data(), (size_t)n);
  s->blk.upload(blk.data(), (size_t)n + 1);
  int64_t words = s->total_blocks * 4;
  s->bits2.alloc((size_t)words + 1);     // + 1: the kernels read the word after the last base
  s->inv16.alloc((size_t)words + 1);
  KL_CUDA(cudaMemsetAsync(s->bits2.p + words, 0, sizeof(uint32_t), ctx().stream));
  KL_CUDA(cudaMemsetAsync(s->inv16.p + words, 0, sizeof(uint16_t), ctx().stream));
  sync_stream();                  // len / blk are stack vectors of this call
  return s;
}

// Host sequences feeding an extraction: the ASCII bytes go to the device in chunks on the copy stream
// while the previous chunk is packed and extracted on the main stream.
struct HostFeed {
  const uint8_t *seq = nullptr;   // first base of sequence 0
  std::vector<int64_t> blk, rel;
  DevBuf<uint8_t> raw;
  DevBuf<int64_t> doff;
  // copy the bases of the sequences [r0, r1) and pack them (main stream waits for the copy)
  void feed(SeqSet &s, int64_t r0, int64_t r1, int slot) {
    const int64_t b0 = rel[r0], b1 = rel[r1];
    if (b1 > b0) {
      KL_CUDA(cudaMemcpyAsync(raw.p + b0, seq + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, ctx().copy_stream));
      KL_CUDA(cudaEventRecord(ctx().copy_ev[slot], ctx().copy_stream));
      KL_CUDA(cudaStreamWaitEvent(ctx().stream, ctx().copy_ev[slot], 0));
    }
    const int64_t w0 = blk[r0] * 4, w1 = blk[r1] * 4;
    if (w1 > w0)
      KL_LAUNCH(pack_kernel, (unsigned)((w1 - w0 + 255) / 256), 256, 0, raw.p, doff.p, s.blk.p, s.n, w0, w1, s.bits2.p,
                s.inv16.p);
  }
};

static std::shared_ptr<SeqSet> sequences_begin(const uint8_t *seq, const int64_t *off, int64_t n, HostFeed &hf) {
  auto s = sequences_alloc(off, n, hf.blk, hf.rel);
  hf.seq = seq + (n ? off[0] : 0);
  hf.raw.alloc((size_t)(s->total_bases ? s->total_bases : 1));
  hf.doff.alloc((size_t)n + 1);
  hf.doff.upload(hf.rel.data(), (size_t)n + 1);
  // the copy stream must not run ahead of the allocation / earlier work on the main stream
  KL_CUDA(cudaEventRecord(ctx().copy_ev[CTX_COPY_EVENTS - 1], ctx().stream));
  KL_CUDA(cudaStreamWaitEvent(ctx().copy_stream, ctx().copy_ev[CTX_COPY_EVENTS - 1], 0));
  return s;
}

std::shared_ptr<SeqSet> sequences_create(const uint8_t *seq, const int64_t *off, int64_t n) {
  HostFeed hf;
  auto s = sequences_begin(seq, off, n, hf);
  if (n > 0) hf.feed(*s, 0, n, 0);
  sync_stream();
  return s;
}

static std::shared_ptr<Matrix> apply_features(Matrix &cls, const int32_t *features, int64_t nf) {
  for (int64_t j = 0; j < nf; j++) {
    KL_REQUIRE(features[2 * j] >= 0 && features[2 * j] < cls.m && features[2 * j + 1] >= 0 && features[2 * j + 1] < cls.m,
               "features: class index out of range");
  }
  auto out = std::make_shared<Matrix>();
  out->n = cls.n; out->m = nf; out->vt = cls.vt;
  matrix_class_list(cls);
  out->class_k = cls.class_k; out->class_code = cls.class_code; out->n_classes = cls.n_classes;
  out->sharded = cls.sharded; out->n_global = cls.n_global;
  DevBuf<int32_t> dfeat((size_t)(2 * nf));
  dfeat.upload(features, (size_t)(2 * nf));
  DevBuf<uint32_t> cnt((size_t)(cls.n ? cls.n : 1));
  const uint32_t *val = cls.vt == VAL_U32 ? cls.val_u32.p : nullptr;
  unsigned grid = (unsigned)((cls.n * 32 + 127) / 128);
  out->rowptr.alloc((size_t)cls.n + 1);
  if (cls.n > 0) {
    KL_LAUNCH((features_rows<false>), grid, 128, 0, cls.rows(), cls.col.p, val, cls.n, dfeat.p, nf, cnt.p, nullptr,
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
    KL_LAUNCH((features_rows<true>), grid, 128, 0, cls.rows(), cls.col.p, val, cls.n, dfeat.p, nf, nullptr,
              out->rowptr.p, out->col.p, out->vt == VAL_U32 ? out->val_u32.p : nullptr);
  sync_stream();
  return out;
}

static std::shared_ptr<Matrix> extract_impl(const kmerlr_config &cfg, std::shared_ptr<SeqSet> seqs,
                                            const int32_t *frozen_k, const uint64_t *frozen_code, int64_t n_frozen,
                                            const int32_t *features, int64_t n_features, int flags, HostFeed *feed) {
  require_ready();
  const SeqSet &s = *seqs;
  KL_REQUIRE(cfg.alphabet == 0 || cfg.alphabet == 1, "alphabet: 0 = nucleotide, 1 = gapped nucleotide (IUPAC is not implemented)");
  if (cfg.alphabet == 1) {
    // gapped alphabet: the sort-based path of gapped.cu (needs the whole set packed)
    KL_REQUIRE(n_features == 0 || n_frozen > 0, "an explicit feature list needs a frozen class list");
    if (feed && s.n > 0) feed->feed(*seqs, 0, s.n, 0);
    auto g = extract_gapped(cfg, seqs, frozen_k, frozen_code, n_frozen, flags);
    if (n_features > 0) return apply_features(*g, features, n_features);
    return g;
  }
  KL_REQUIRE(cfg.M >= 1 && cfg.M <= cfg.N, "need 1 <= M <= N");
  KL_REQUIRE(cfg.N <= MAX_N, "k-mer length above 13 is not supported on the GPU path");
  int nops = (cfg.complement != 0) + (cfg.reverse != 0) + (cfg.revcomp != 0);
  KL_REQUIRE(nops <= 1, "at most one of complement / reverse / revcomp is supported on the GPU path");
  KL_REQUIRE(n_features == 0 || n_frozen > 0, "an explicit feature list needs a frozen class list");
  const bool sharded = (flags & KMERLR_FLAG_SHARDED) != 0 && ctx().world > 1;

  XParams P{};
  P.M = cfg.M; P.N = cfg.N; P.binarize = cfg.binarize != 0;
  P.op = cfg.revcomp ? 1 : (cfg.complement ? 2 : (cfg.reverse ? 3 : 0));
  P.t_lo = cfg.M; P.t_hi = cfg.N < KT_MAX ? cfg.N : KT_MAX;
  // levels above the tables: per-row bitmaps up to level 7, a register sort per level above that;
  // rows too long for the register sort (more than 64 positions per lane) keep level 8 on a bitmap
  const int64_t steps = (s.max_len + 31) / 32;           // start positions per lane
  const int kb = steps > 64 ? KB_MAX : KB_DEFAULT;
  P.b_lo = cfg.M > KT_MAX + 1 ? cfg.M : KT_MAX + 1; P.b_hi = cfg.N < kb ? cfg.N : kb;
  P.s_lo = cfg.M > kb + 1 ? cfg.M : kb + 1;
  P.n = s.n;
  P.len = s.len.p; P.blk = s.blk.p; P.bits2 = s.bits2.p; P.inv16 = s.inv16.p;
  // dense class ids: every level starts on a multiple of 32 so that bitmap words never straddle levels
  uint64_t dense = 0;
  for (int k = cfg.M; k <= cfg.N; k++) {
    P.level_off[k] = (uint32_t)dense;
    dense += ((1ull << (2 * k)) + 31) / 32 * 32;
  }
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
  KL_REQUIRE(s.max_len < 65536 || P.t_lo > P.t_hi, "sequences of 65536 bp or more need M > 5 on the GPU path");
  KL_REQUIRE(s.max_len < ((int64_t)1 << 30), "sequence too long");
  // keys per lane of the register sort
  int E = 0;
  if (P.s_lo <= cfg.N) {
    E = 2; while (E < steps && E < 128) E <<= 1;
    KL_REQUIRE(E <= 64, "sequence too long for the register sort path (k > 8 needs L <= ~2000 bp)");
  }
  // per-warp shared memory
  P.bm_words = 0; P.pf_words = 0;
  int nbl = 0;
  for (int k = P.b_lo; k <= P.b_hi; k++) {
    P.bm_off[k] = (uint32_t)P.bm_words; P.pf_off[k] = (uint32_t)P.pf_words;
    P.bm_words += 1 << (2 * k - 5); P.pf_words += 1 << (2 * k - 7);
    nbl++;
  }
  P.pf_words = (P.pf_words + 3) / 4 * 4;
  P.ts_words = 32 * E > TAB_WORDS ? 32 * E : TAB_WORDS;
  P.warp_words = P.bm_words + P.pf_words + P.ts_words + 4 + DUP_SMEM;
  P.warp_words = (P.warp_words + 3) / 4 * 4;
  // the block's bitmap of observed classes covers the levels up to OBS_MAX_LEVEL
  {
    int top = cfg.N < OBS_MAX_LEVEL ? cfg.N : OBS_MAX_LEVEL;
    P.obs_words = top >= cfg.M ? (int)(P.level_off[top + 1] / 32) : 0;
    P.obs_words = (P.obs_words + 3) / 4 * 4;
  }
  // classes of the table levels
  if (P.t_lo <= P.t_hi) P.tl = table_classes(P.op, P.t_lo, P.t_hi, P.level_off, &P.tl_cnt);
  Trace tr("extract");
  auto out = std::make_shared<Matrix>();
  out->n = s.n; out->vt = P.binarize ? VAL_ONE : VAL_U32;
  out->sharded = sharded; out->n_global = s.n;
  // the kernel writes the rows of the matrix itself: fixed stride, final column numbers
  out->row_stride = stride;
  out->col.alloc((size_t)(s.n ? s.n * stride : 1));
  if (!P.binarize) out->val_u32.alloc((size_t)(s.n ? s.n * stride : 1));
  out->rowcnt.alloc((size_t)(s.n ? s.n : 1));
  DevBuf<uint32_t> dummy_cnt(1);
  // numbering set: the frozen list, or every class the configuration can produce (then the columns
  // are final unless some class never shows up, see below)
  DevBuf<uint32_t> nb((size_t)nw), nbrank((size_t)nw + 1), pc((size_t)nw), bitmap((size_t)nw);
  std::vector<uint32_t> frozen_bits;       // host image of the frozen class set (alive until the next sync)
  LevelInfo LI{};
  LI.M = cfg.M; LI.N = cfg.N; LI.op = P.op;
  for (int k = cfg.M; k <= cfg.N + 1; k++) LI.level_off[k] = P.level_off[k];
  if (n_frozen > 0) {
    frozen_bits.assign((size_t)nw, 0u);
    uint64_t prev = 0;
    for (int64_t j = 0; j < n_frozen; j++) {
      int k = frozen_k[j];
      KL_REQUIRE(k >= cfg.M && k <= cfg.N && frozen_code[j] < (1ull << (2 * k)), "frozen class outside [M,N]");
      uint64_t id = (uint64_t)P.level_off[k] + frozen_code[j];
      KL_REQUIRE(j == 0 || id > prev, "frozen class list must be sorted by (k, code) without duplicates");
      prev = id;
      frozen_bits[id >> 5] |= 1u << (id & 31);
    }
    nb.upload(frozen_bits.data(), (size_t)nw);
  } else {
    KL_LAUNCH(all_classes_bits, (unsigned)((nw + 255) / 256), 256, 0, LI, nb.p, nw);
  }
  KL_LAUNCH(popc_words, (unsigned)((nw + 255) / 256), 256, 0, nb.p, nw, pc.p);
  exclusive_scan_u32(pc.p, nbrank.p, nw);
  bitmap.zero();
  DevBuf<unsigned long long> stats(2);
  stats.zero();
  P.st_id = out->col.p; P.st_cnt = P.binarize ? dummy_cnt.p : out->val_u32.p; P.rowcnt = out->rowcnt.p;
  DevBuf<uint2> nbx((size_t)nw);
  KL_LAUNCH(interleave_numbering, (unsigned)((nw + 255) / 256), 256, 0, nb.p, nbrank.p, nw, nbx.p);
  P.bitmap = bitmap.p; P.nbx = nbx.p; P.filter = n_frozen > 0; P.stats = stats.p;
  P.mark = n_frozen == 0;
  // repeats beyond the shared-memory list of a warp (low-complexity rows): one global list per warp
  DevBuf<uint32_t> ovf;
  P.ovf_stride = nbl * (s.max_len > 0 ? s.max_len : 1);
  if (nbl && !P.binarize && s.n > 0) {
    ovf.alloc((size_t)ctx().sm_count * 64 * (size_t)P.ovf_stride);
    P.ovf = ovf.p;
  }
  tr.mark("alloc");

  if (s.n > 0) {
    // host sequences arrive in chunks: copy of chunk c+1 overlaps pack + extraction of chunk c
    const int nchunks = (feed && s.n >= 4096) ? CTX_COPY_EVENTS - 1 : 1;
    for (int c = 0; c < nchunks; c++) {
      const int64_t r0 = s.n * c / nchunks, r1 = s.n * (c + 1) / nchunks;
      if (feed) feed->feed(*seqs, r0, r1, c);
      P.row0 = r0; P.n = r1;
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
    P.row0 = 0; P.n = s.n;
  }
  tr.mark("extract_kernel");
  // observed classes of all ranks; do they cover the numbering set?
  DevBuf<uint32_t> differ(1);
  differ.zero();
  if (n_frozen == 0) {
    if (sharded) {
      DevBuf<uint8_t> bytes((size_t)nbits);
      KL_LAUNCH(bitmap_to_bytes, (unsigned)((nbits + 255) / 256), 256, 0, bitmap.p, nbits, bytes.p);
      comm_allreduce_max_u8(bytes.p, nbits);
      KL_LAUNCH(bytes_to_bitmap, (unsigned)((nw + 255) / 256), 256, 0, bytes.p, nbits, bitmap.p, nw);
    }
    KL_LAUNCH(bitmaps_differ, (unsigned)((nw + 255) / 256), 256, 0, bitmap.p, nb.p, nw, differ.p);
  }
  // one round trip: classes, stored entries, "columns are final", row statistics, global number of rows
  DevBuf<int64_t> rp((size_t)s.n + 1);
  if (s.n > 0) exclusive_scan_u32_to_i64(out->rowcnt.p, rp.p, s.n); else rp.zero();
  DevBuf<int64_t> nglob(1);
  int64_t nn = s.n;
  if (sharded) {
    nglob.upload(&nn, 1);
    comm_allreduce_sum_i64(nglob.p, 1);
    nglob.download(&nn, 1);
  }
  uint32_t m32 = 0, hdiffer = 0;
  unsigned long long hstats[2] = {0, 0};
  KL_CUDA(cudaMemcpyAsync(&m32, nbrank.p + nw, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx().stream));
  KL_CUDA(cudaMemcpyAsync(&out->nnz, rp.p + s.n, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx().stream));
  differ.download(&hdiffer, 1);
  stats.download(hstats, 2);
  sync_stream();
  out->n_global = nn;
  out->has_local_stats = true;
  out->local_maxsq = (double)hstats[0]; out->local_vmax = (double)hstats[1];
  tr.mark("sizes");
  if (hdiffer) {
    // some class of the configuration never occurs: the columns are the ranks among the OBSERVED
    // classes (kmerLr_data.go:316-324), renumber the stored columns in place
    DevBuf<uint32_t> rank2((size_t)nw + 1), ids_nb((size_t)(m32 ? m32 : 1));
    KL_LAUNCH(popc_words, (unsigned)((nw + 255) / 256), 256, 0, bitmap.p, nw, pc.p);
    exclusive_scan_u32(pc.p, rank2.p, nw);
    KL_LAUNCH(enumerate_bits, (unsigned)((nw + 255) / 256), 256, 0, nb.p, nbrank.p, nw, ids_nb.p);
    if (s.n > 0)
      KL_LAUNCH(renumber_rows, (unsigned)((s.n * 32 + 127) / 128), 128, 0, out->col.p, out->rowcnt.p, stride, s.n, ids_nb.p,
                bitmap.p, rank2.p);
    KL_CUDA(cudaMemcpyAsync(&m32, rank2.p + nw, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx().stream));
    sync_stream();
    nb = std::move(bitmap); nbrank = std::move(rank2);
    tr.mark("renumber");
  }
  out->m = (int64_t)m32;
  // class list: dense ids stay on the device, (k, code) pairs are decoded on demand
  out->n_classes = (int64_t)m32;
  out->class_ids.alloc((size_t)(m32 ? m32 : 1));
  KL_LAUNCH(enumerate_bits, (unsigned)((nw + 255) / 256), 256, 0, nb.p, nbrank.p, nw, out->class_ids.p);
  out->class_M = cfg.M; out->class_N = cfg.N; out->classes_on_host = false;
  for (int k = cfg.M; k <= cfg.N + 1; k++) out->class_level_off[k] = P.level_off[k];
  if (n_features > 0) { sync_stream(); return apply_features(*out, features, n_features); }
  // count matrices with one column per class keep what the matrix-free logistic pass needs
  if (!P.binarize && cfg.N <= IMP_MAX_N && s.n > 0 && m32 > 0) {
    auto imp = std::make_shared<Implicit>();
    imp->seqs = seqs; imp->M = cfg.M; imp->N = cfg.N; imp->op = P.op;
    uint32_t fo = 0;
    for (int k = cfg.M; k <= cfg.N + 1; k++) {
      imp->level_off[k] = P.level_off[k];
      imp->fo[k] = fo;
      if (k <= cfg.N) fo += 1u << (2 * k);
    }
    imp->bitmap = std::move(nb); imp->rank = std::move(nbrank);
    out->imp = imp;
  }
  sync_stream();
  return out;
}

void matrix_class_list(Matrix &M) {
  if (M.classes_on_host) return;
  std::vector<uint32_t> hid((size_t)M.n_classes);
  M.class_ids.download(hid.data(), (size_t)M.n_classes);
  sync_stream();
  M.class_k.resize((size_t)M.n_classes); M.class_code.resize((size_t)M.n_classes);
  int k = M.class_M;
  for (size_t j = 0; j < hid.size(); j++) {
    while (k < M.class_N && hid[j] >= M.class_level_off[k + 1]) k++;
    M.class_k[j] = k; M.class_code[j] = hid[j] - M.class_level_off[k];
  }
  M.classes_on_host = true;
}

std::shared_ptr<Matrix> extract(const kmerlr_config &cfg, std::shared_ptr<SeqSet> seqs, const int32_t *frozen_k,
                                const uint64_t *frozen_code, int64_t n_frozen, const int32_t *features,
                                int64_t n_features, int flags) {
  return extract_impl(cfg, seqs, frozen_k, frozen_code, n_frozen, features, n_features, flags, nullptr);
}

// kmerlr_extract: sequences in host memory; upload, packing and extraction are pipelined
std::shared_ptr<Matrix> extract_host(const kmerlr_config &cfg, const uint8_t *seq, const int64_t *off, int64_t n,
                                     const int32_t *frozen_k, const uint64_t *frozen_code, int64_t n_frozen,
                                     const int32_t *features, int64_t n_features, int flags) {
  HostFeed hf;
  auto s = sequences_begin(seq, off, n, hf);
  auto out = extract_impl(cfg, s, frozen_k, frozen_code, n_frozen, features, n_features, flags, &hf);
  sync_stream();
  return out;
}

}  // namespace kl
