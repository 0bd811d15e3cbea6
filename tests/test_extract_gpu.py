"""Stage 1 parity: CUDA k-mer extraction (through the C ABI) vs the CPU oracle -- bit exact counts,
column indices and class lists (north_star).  Needs a B200."""
import numpy as np
import pytest

from conftest import cat

pytestmark = pytest.mark.gpu


def same_matrix(d, ref):
    assert (d.n, d.m, d.nnz) == (ref.n, ref.m, ref.nnz)
    k, code = d.Kmers()
    ok, ocode = ref.classes()
    assert np.array_equal(k, ok) and np.array_equal(code, ocode)
    for x, y in zip(d.rows(), ref.rows()):
        assert np.array_equal(x, y)


def cfg_pair(K, O, M, N, **kw):
    return K.NewKmerCounter(M, N, **kw), O.make_config(M, N, **kw)


@pytest.mark.parametrize("M,N,flags", [
    (1, 6, dict(revcomp=True)), (2, 6, dict(revcomp=True)), (8, 8, dict(revcomp=True)), (1, 8, dict(revcomp=True)),
    (1, 8, dict()), (3, 7, dict(complement=True)), (3, 7, dict(reverse=True)), (1, 4, dict(revcomp=True)),
    (6, 6, dict()), (1, 1, dict(revcomp=True)), (2, 10, dict(revcomp=True, binarize=True)),
])
def test_reference_fixtures(K, oracle, fixtures, M, N, flags):
    """C1: the bundled kmerLr_test_{fg,bg}.fa (11 + 11 x 1000 bp, mixed case)"""
    kc, oc = cfg_pair(K, oracle, M, N, **flags)
    buf, off, _ = cat(fixtures, "kmerLr_test_fg", "kmerLr_test_bg")
    d = K.compile_test_data(None, kc, None, None, True, flags.get("binarize", False), (buf, off))
    same_matrix(d, oracle.extract(oc, (buf, off)))


def test_cooccurrence_fixture(K, oracle, fixtures):
    """C1: kmerLr_test_co_{fg,bg}.fa, 2 6 --binarize --revcomp (TestKmers6 data): m = 70"""
    kc, oc = cfg_pair(K, oracle, 2, 6, revcomp=True, binarize=True)
    buf, off, y = cat(fixtures, "kmerLr_test_co_fg", "kmerLr_test_co_bg")
    d = K.compile_training_data(None, kc, None, None, True, True, (buf[:off[10]], off[:11]),
                                (buf[off[10]:], off[10:] - off[10]))
    assert d.m == 70
    same_matrix(d, oracle.extract(oc, (buf, off)))
    assert np.all(d.rows()[2] == 1.0)


@pytest.mark.parametrize("L,M,N,flags", [
    (500, 1, 8, dict(revcomp=True)), (200, 1, 10, dict(revcomp=True, binarize=True)), (37, 1, 8, dict(revcomp=True)),
    (64, 5, 9, dict()), (300, 6, 12, dict(revcomp=True)), (120, 1, 13, dict(revcomp=True)), (500, 1, 5, dict(revcomp=True)),
    (3000, 1, 5, dict(revcomp=True)), (1900, 6, 7, dict()), (2500, 1, 8, dict(revcomp=True)), (700, 4, 9, dict(revcomp=True)),
])
def test_synthetic(K, oracle, L, M, N, flags):
    from kmerlr_b200 import synth
    kc, oc = cfg_pair(K, oracle, M, N, **flags)
    buf, off, _ = synth.training_set(150, 150, L)
    d = K.compile_test_data(None, kc, None, None, True, flags.get("binarize", False), (buf, off))
    same_matrix(d, oracle.extract(oc, (buf, off)))


def test_ragged_empty_and_invalid(K, oracle):
    """ragged lengths, empty sequences, sequences shorter than k, lower case, N and other bytes"""
    rng = np.random.default_rng(7)
    seqs = []
    for L in [0, 1, 2, 5, 7, 8, 9, 31, 32, 33, 63, 64, 65, 100, 257, 400, 0, 3]:
        s = rng.choice(list("ACGTacgt"), size=L)
        seqs.append("".join(s))
    seqs[9] = seqs[9][:10] + "N" + seqs[9][11:]
    seqs[13] = "NNNN" + seqs[13][4:50] + "nX-" + seqs[13][53:]
    seqs[14] = seqs[14][:100] + "N" * 30 + seqs[14][130:]
    seqs.append("N" * 40)
    for M, N, flags in [(1, 8, dict(revcomp=True)), (2, 7, dict()), (1, 3, dict(revcomp=True)), (6, 9, dict(reverse=True)),
                        (4, 8, dict(complement=True, binarize=True))]:
        kc, oc = cfg_pair(K, oracle, M, N, **flags)
        d = K.compile_test_data(None, kc, None, None, True, flags.get("binarize", False), seqs)
        same_matrix(d, oracle.extract(oc, seqs))
    # no sequences at all
    d = K.compile_test_data(None, K.NewKmerCounter(1, 6, revcomp=True), None, None, True, False, [])
    assert (d.n, d.m, d.nnz) == (0, 0, 0)


def test_chunked_host_feed_ragged_nonzero_first_offset(K, oracle):
    """n >= 16384 takes the pipelined feed (n / 8192 chunks of growing size, 4 here); the offsets need not start at 0,
    rows are ragged and some are empty; the resident path (K.Sequences) must agree"""
    rng = np.random.default_rng(23)
    n = 33003
    lens = rng.integers(0, 70, size=n)
    lens[rng.integers(0, n, size=200)] = 0
    lens[:3] = [0, 0, 1]
    lens[-2:] = [0, 65]
    lead = 37
    off = np.concatenate([[lead], lead + np.cumsum(lens)]).astype(np.int64)
    buf = rng.choice(np.frombuffer(b"ACGTacgtN", dtype=np.uint8), size=int(off[-1]) + 11,
                     p=[.2, .2, .2, .2, .045, .045, .045, .045, .02])
    seqs = [bytes(buf[off[i]:off[i + 1]]) for i in range(n)]
    for M, N, flags in [(1, 6, dict(revcomp=True)), (3, 8, dict())]:
        kc, oc = cfg_pair(K, oracle, M, N, **flags)
        ref = oracle.extract(oc, seqs)
        d = K.compile_test_data(None, kc, None, None, True, False, (buf, off))
        same_matrix(d, ref)
        d.free()
        res = K.Sequences((buf, off))
        d = K.compile_test_data(None, kc, None, None, True, False, res)
        same_matrix(d, ref)
        d.free()
    with pytest.raises(K.KmerLrError):
        bad = off.copy(); bad[100] = bad[101] + 5
        K.compile_test_data(None, K.NewKmerCounter(1, 6), None, None, True, False, (buf, bad))


def test_low_complexity_repeats(K, oracle):
    """homopolymers and short tandem repeats: almost every k-mer instance repeats an earlier one (the
    repeat lists of the bitmap levels overflow their shared-memory part), counts up to L"""
    rng = np.random.default_rng(11)
    seqs = ["A" * 600, "AC" * 300, "ACG" * 333, "T" * 2100, "ACGTTGCA" * 250, "G" * 9,
            "".join(rng.choice(list("AC"), size=800)), "AAAAAAAAC" * 100 + "N" + "GT" * 200]
    for M, N, flags in [(1, 8, dict(revcomp=True)), (6, 7, dict()), (5, 8, dict(reverse=True)), (1, 8, dict(revcomp=True, binarize=True))]:
        kc, oc = cfg_pair(K, oracle, M, N, **flags)
        d = K.compile_test_data(None, kc, None, None, True, flags.get("binarize", False), seqs)
        same_matrix(d, oracle.extract(oc, seqs))
    short = [s[:900] for s in seqs]
    kc, oc = cfg_pair(K, oracle, 7, 10, revcomp=True)
    d = K.compile_test_data(None, kc, None, None, True, False, short)
    same_matrix(d, oracle.extract(oc, short))


def test_frozen_counter_and_features(K, oracle, fixtures):
    """TestKmers2 property on the nucleotide alphabet + explicit feature lists with pair products
    (kmerLr_data.go:210-229, the shape of kmerLr_test.go:36-66)"""
    kc, oc = cfg_pair(K, oracle, 2, 7, revcomp=True)
    buf, off, _ = cat(fixtures, "kmerLr_test_fg", "kmerLr_test_bg")
    a = K.compile_test_data(None, kc, None, None, True, False, (buf, off))
    classes = a.Kmers()
    b = K.compile_test_data(None, kc, classes, None, True, False, (buf, off))
    for x, y in zip(a.rows(), b.rows()):
        assert np.array_equal(x, y)
    # frozen to a subset, applied to other sequences
    sub = (classes[0][::3], classes[1][::3])
    from kmerlr_b200 import synth
    sbuf, soff, _ = synth.training_set(20, 20, 300)
    c = K.compile_test_data(None, kc, sub, None, True, False, (sbuf, soff))
    same_matrix(c, oracle.extract(oc, (sbuf, soff), frozen=sub))
    # explicit features: every 5th single plus some pairs
    m = len(sub[0])
    rng = np.random.default_rng(3)
    feats = [(i, i) for i in range(0, m, 5)] + [tuple(sorted(rng.choice(m, 2, replace=False))) for _ in range(300)]
    e = K.compile_test_data(None, kc, sub, feats, False, False, (sbuf, soff))
    ref = oracle.extract(oc, (sbuf, soff), frozen=sub, features=feats)
    assert (e.n, e.m, e.nnz) == (ref.n, ref.m, ref.nnz)
    for x, y in zip(e.rows(), ref.rows()):
        assert np.array_equal(x, y)


def test_gapped_alphabet_reference_pins(K, oracle, fixtures):
    """TestKmers1 / TestKmers2 on the GPU (kmerLr_test.go:30-97): gapped alphabet, k = 4..8, revcomp --
    the six (index, class, count) pins, the four pair products, frozen = unfrozen"""
    kc, oc = cfg_pair(K, oracle, 4, 8, revcomp=True, alphabet="gapped-nucleotide")
    buf, off, _ = cat(fixtures, "kmerLr_test", "kmerLr_test")
    d = K.compile_test_data(None, kc, None, None, True, False, (buf, off))
    ref = oracle.extract(oc, (buf, off))
    assert d.n == 4 and d.m == 58308
    same_matrix(d, ref)
    names = ref.class_names()
    rp, col, val = d.rows()
    row0 = dict(zip(col[rp[0]:rp[1]].tolist(), val[rp[0]:rp[1]].tolist()))
    pins = [(4671, "gntanc|gntanc", 3), (4672, "gntcaa|ttganc", 0), (5068, "aaagaaa|tttcttt", 1),
            (5486, "aagannt|anntctt", 7), (19270, "aacgcgna|tncgcgtt", 1), (57071, "tgaatgca|tgcattca", 1)]
    for idx, name, count in pins:                                   # kmerLr_test.go:40-43
        assert names[idx] == name and row0.get(idx, 0) == count
    classes = d.Kmers()
    d2 = K.compile_test_data(None, kc, classes, None, True, False, (buf, off))       # frozen counter
    for x, y in zip(d.rows(), d2.rows()):
        assert np.array_equal(x, y)
    m = d.m
    feats = [(i, i) for i in range(0, m, 97)] + [(4671, 4672), (5068, 5486), (19270, 57071), (4671, 5486)]
    d3 = K.compile_test_data(None, kc, classes, feats, False, False, (buf, off))
    r3 = d3.rows()
    last = dict(zip(r3[1][r3[0][0]:r3[0][1]].tolist(), r3[2][r3[0][0]:r3[0][1]].tolist()))
    nf = len(feats)
    assert [last.get(nf - 4 + i, 0) for i in range(4)] == [0, 7, 1, 21]      # kmerLr_test.go:55-66


def test_gapped_alphabet_variants(K, oracle):
    """gapped alphabet on ragged synthetic rows: max_ambiguous, other strand operations, binarize, invalid bases"""
    from kmerlr_b200 import synth
    buf, off, _ = synth.training_set(9, 8, 90)
    buf = buf.copy()
    buf[off[2] + 7] = ord("N")
    seqs = (buf, off)
    for M, N, flags in [(2, 6, dict(revcomp=True)), (3, 5, dict(max_ambiguous=1)), (4, 7, dict(reverse=True, max_ambiguous=2)),
                        (1, 4, dict(complement=True, binarize=True)), (5, 6, dict(revcomp=True, max_ambiguous=0))]:
        flags = dict(flags, alphabet="gapped-nucleotide")
        ma = flags.get("max_ambiguous")
        kc = K.NewKmerCounter(M, N, **flags)
        oc = oracle.make_config(M, N, **{k: v for k, v in flags.items() if k != "max_ambiguous"},
                                max_ambiguous=-1 if ma is None else ma)
        d = K.compile_test_data(None, kc, None, None, True, flags.get("binarize", False), seqs)
        same_matrix(d, oracle.extract(oc, seqs))


def test_unsupported_configurations_fail_loudly(K):
    with pytest.raises(K.KmerLrError):
        K.compile_test_data(None, K.NewKmerCounter(4, 12, revcomp=True, alphabet="gapped-nucleotide"), None, None, True,
                            False, ["ACGTACGTACGTACGT"])
    with pytest.raises(K.KmerLrError):
        K.compile_test_data(None, K.NewKmerCounter(1, 15), None, None, True, False, ["ACGTACGTACGTACGT"])


@pytest.mark.parametrize("M,N,flags,L", [
    (1, 6, dict(complement=True, reverse=True), 300),                      # two strand flags, not closed under composition
    (2, 8, dict(complement=True, reverse=True, revcomp=True), 250),         # all three: orbits of the Klein four-group
    (3, 9, dict(reverse=True, revcomp=True, binarize=True), 120),
    (1, 14, dict(revcomp=True), 90),                                        # k-mers of 14 bases
    (12, 14, dict(), 200),
    (6, 9, dict(revcomp=True), 8000),                                       # k > 8 on rows beyond the register sort
    (1, 10, dict(revcomp=True, binarize=True), 2600),
    (4, 7, dict(revcomp=True), 70000),                                      # rows beyond the 16-bit count tables
])
def test_sort_based_path_lifts_the_cliffs(K, oracle, M, N, flags, L):
    """what the warp-per-row kernel does not take goes through the sort-based path (gapped.cu): several strand flags
    at once (kmerLr_learn.go:94 passes them independently; class = min over every enabled image), k = 14, long rows
    with k > 8.  Bit exact against the oracle, frozen subsets and gradient included."""
    from kmerlr_b200 import synth
    nrow = 6 if L >= 2600 else 40
    buf, off, y = synth.training_set(nrow // 2, nrow // 2, L)
    buf = buf.copy()
    buf[off[1] + 7] = ord("N")
    buf[off[2]:off[2] + L // 3] = ord("A")
    binz = flags.get("binarize", False)
    kc, oc = cfg_pair(K, oracle, M, N, **flags)
    d = K.compile_test_data(None, kc, None, None, True, binz, (buf, off))
    ref = oracle.extract(oc, (buf, off))
    same_matrix(d, ref)
    k, code = d.Kmers()
    sub = (k[::3], code[::3])
    same_matrix(K.compile_test_data(None, kc, sub, None, True, binz, (buf, off)), oracle.extract(oc, (buf, off), frozen=sub))
    d2 = K.compile_test_data(None, kc, None, None, True, binz, (buf, off))
    d2.SetLabels(y)
    theta = np.random.default_rng(4).normal(scale=0.01, size=d2.m + 1)
    g, og = K.logisticRegression(theta).Gradient(None, d2), oracle.gradient(ref, y, theta)
    assert np.max(np.abs(g - og)) <= 1e-10 * np.max(np.abs(og))


def test_full_size_properties(K):
    """C2-shaped rows at a reduced count: size independent properties (sortedness, no explicit zeros,
    row sums = number of k-mer instances, revcomp invariance)"""
    from kmerlr_b200 import synth
    buf, off, _ = synth.training_set(2048, 2048, 500)
    kc = K.NewKmerCounter(1, 8, revcomp=True)
    d = K.compile_test_data(None, kc, None, None, True, False, (buf, off))
    rp, col, val = d.rows()
    assert d.m == 43860 or d.m <= 43860
    assert np.all(val >= 1)
    inst = sum(500 - k + 1 for k in range(1, 9))
    sums = np.add.reduceat(val, rp[:-1])
    assert np.all(sums == inst)
    for i in (0, 17, 4095):
        c = col[rp[i]:rp[i + 1]]
        assert np.all(np.diff(c) > 0)
    # reverse complementing every sequence leaves the revcomp-merged rows unchanged
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    rc = comp[buf.reshape(-1, 500)[:, ::-1]].reshape(-1)
    d2 = K.compile_test_data(None, kc, d.Kmers(), None, True, False, (rc, off))
    for x, y in zip(d.rows(), d2.rows()):
        assert np.array_equal(x, y)


def test_c2_full_size_invariants(K):
    """BASELINE config C2 at full size (100 000 + 100 000 x 500 bp, k = 1..8, revcomp) through size independent
    properties, without copying the 3 GB matrix back: every row sums to the number of k-mer instances
    (x . 1 = 3 972), the class count is the closed form sum_k (4^k + [k even] 4^(k/2)) / 2 = 43 860, the matrix
    does not change when every sequence is reverse complemented, and the matrix-free pass agrees with the
    CSR pass on the full matrix"""
    from kmerlr_b200 import synth
    n_half, L = 100000, 500
    buf, off, y = synth.training_set(n_half, n_half, L)
    kc = K.NewKmerCounter(1, 8, revcomp=True)
    d = K.compile_test_data(None, kc, None, None, True, False, (buf, off))
    assert d.n == 2 * n_half
    assert d.m == sum((4 ** k + (4 ** (k // 2) if k % 2 == 0 else 0)) // 2 for k in range(1, 9)) == 43860
    inst = sum(L - k + 1 for k in range(1, 9))
    ones = np.ones(d.m + 1)
    z = K.logisticRegression(ones).LinearPdf(d)
    assert np.all(z == 1.0 + inst)
    assert d.nnz == int(K.logisticRegression(np.concatenate([[0.0], np.ones(d.m)])).LinearPdf(
        K.compile_test_data(None, K.NewKmerCounter(1, 8, revcomp=True, binarize=True), d.Kmers(), None, True, True, (buf, off))).sum())
    d.SetLabels(y)
    rng = np.random.default_rng(5)
    theta = rng.normal(scale=0.01, size=d.m + 1)
    lr = K.logisticRegression(theta, (1.0, 1.0), 0.0)
    g = lr.Gradient(None, d)
    K.option("implicit", 0)
    try:
        g_csr = lr.Gradient(None, d)
    finally:
        K.option("implicit", 1)
    assert np.max(np.abs(g - g_csr)) <= 1e-12 * np.max(np.abs(g))
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    rc = comp[buf.reshape(-1, L)[:, ::-1]].reshape(-1)
    d2 = K.compile_test_data(None, kc, None, None, True, False, (rc, off))
    d2.SetLabels(y)
    assert (d2.n, d2.m, d2.nnz) == (d.n, d.m, d.nnz)
    K.option("implicit", 0)
    try:
        assert np.array_equal(lr.Gradient(None, d2), g_csr)  # identical stored rows -> bit-identical gradient
    finally:
        K.option("implicit", 1)
    assert np.max(np.abs(lr.Gradient(None, d2) - g)) <= 1e-12 * np.max(np.abs(g))


def test_compile_data_shares_one_numbering(K, oracle):
    """compile_data (kmerLr_data.go:339-358): the sets get the column numbering of their union"""
    from kmerlr_b200 import synth
    a = synth.sequences(40, 90, 1)
    b = synth.sequences(25, 60, 2)
    c = synth.sequences(1, 30, 3)
    kc, oc = cfg_pair(K, oracle, 2, 7, revcomp=True)
    parts = K.compile_data(None, kc, None, None, True, False, [a, b, c])
    buf = np.concatenate([a[0][:a[1][-1]], b[0][:b[1][-1]], c[0][:c[1][-1]]])
    off = np.concatenate([a[1], b[1][1:] + a[1][-1], c[1][1:] + a[1][-1] + b[1][-1]])
    ref = oracle.extract(oc, (buf, off))
    rp, rc, rv = ref.rows()
    lo = 0
    for part in parts:
        assert part.m == ref.m
        k, code = part.Kmers(); ok, ocode = ref.classes()
        assert np.array_equal(k, ok) and np.array_equal(code, ocode)
        p, cc, v = part.rows()
        assert np.array_equal(p, rp[lo:lo + part.n + 1] - rp[lo])
        assert np.array_equal(cc, rc[rp[lo]:rp[lo + part.n]]) and np.array_equal(v, rv[rp[lo]:rp[lo + part.n]])
        lo += part.n
        part.free()
    assert lo == ref.n


def test_generate_features_off_with_empty_feature_list_gives_bias_only_rows(K):
    """convert_counts (kmerLr_data.go:197-235): `len(features) == 0 && generate_features` is the only way to one
    column per class; an empty list with generate_features off leaves the bias alone, Kmers stays the class list"""
    from kmerlr_b200 import synth
    seqs = synth.sequences(30, 80, 5)
    kc = K.NewKmerCounter(1, 4, revcomp=True)
    full = K.compile_test_data(None, kc, None, None, True, False, seqs)
    d = K.compile_test_data(None, kc, None, None, False, False, seqs)
    assert (d.n, d.m, d.nnz) == (30, 0, 0) and d.Dim() == 1
    k, code = d.Kmers(); fk, fcode = full.Kmers()
    assert np.array_equal(k, fk) and np.array_equal(code, fcode)
    lr = K.logisticRegression(np.array([0.37]), (1.0, 1.0), 0.0)
    assert np.array_equal(lr.LinearPdf(d), np.full(30, 0.37))
    d.free(); full.free()


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,flags,L", [
    (1, 4, dict(revcomp=True), 60), (1, 6, dict(), 40), (2, 9, dict(revcomp=True, binarize=True), 48), (5, 8, dict(revcomp=True), 64),
])
def test_marks_of_observed_classes_end_only_when_all_are_seen(K, oracle, M, N, flags, L):
    """Many row groups per block: the kernel stops marking observed classes once the device-wide bitmap covers the
    numbering set.  (a) random rows: every class of a small configuration is seen early, the marks end, nothing may
    change; (b) a class that shows up only in the LAST rows (homopolymer rows before them) must still get its column;
    (c) a configuration whose classes are never all seen keeps marking (columns = ranks among the observed classes)."""
    from kmerlr_b200 import synth
    kc, oc = cfg_pair(K, oracle, M, N, **flags)
    n = 40000
    buf, off, _ = synth.training_set(n // 2, n // 2, L)
    d = K.compile_test_data(None, kc, None, None, True, flags.get("binarize", False), (buf, off))
    same_matrix(d, oracle.extract(oc, (buf, off), threads=8))
    d.free()
    late = buf.copy()
    late[:off[n - 40]] = ord("A")                                  # only the last 40 rows carry anything but poly-A
    d = K.compile_test_data(None, kc, None, None, True, flags.get("binarize", False), (late, off))
    same_matrix(d, oracle.extract(oc, (late, off), threads=8))
    d.free()
