"""Stage 2 parity: sparse logistic loss / gradient, leapfrog selection, reduced matrices and the
proximal-gradient estimator (through the C ABI) vs the CPU oracle.  Needs a B200.

Tolerances (north_star): loss <= 1e-6 relative, theta <= 1e-5; selected feature sets, indices
and the first lambda are exact.  The observed agreement is far tighter and asserted as such."""
import numpy as np
import pytest

from conftest import cat

pytestmark = pytest.mark.gpu

RTOL_G = 1e-10      # gradient entries, relative to max |g|
RTOL_LOSS = 1e-12


def build(K, O, fixtures, M, N, fg="kmerLr_test_fg", bg="kmerLr_test_bg", **flags):
    buf, off, y = cat(fixtures, fg, bg)
    nfg = int(y.sum())
    kc, oc = K.NewKmerCounter(M, N, **flags), O.make_config(M, N, **flags)
    d = K.compile_training_data(None, kc, None, None, True, flags.get("binarize", False), (buf[:off[nfg]], off[:nfg + 1]),
                                (buf[off[nfg]:], off[nfg:] - off[nfg]))
    return d, O.extract(oc, (buf, off)), y


def close_g(g, og):
    scale = max(np.max(np.abs(og)), 1e-300)
    assert np.max(np.abs(g - og)) <= RTOL_G * scale


def test_linear_logpdf_gradient_loss(K, oracle, fixtures):
    d, ref, y = build(K, oracle, fixtures, 1, 6, revcomp=True)
    rng = np.random.default_rng(1)
    for cw, lam in [((1.0, 1.0), 0.0), ((0.7, 1.9), 0.05), ((1.0, 1.0), float("nan"))]:
        theta = rng.normal(scale=0.01, size=d.m + 1)
        theta[rng.integers(1, d.m + 1, size=d.m // 2)] = 0.0
        lr = K.logisticRegression(theta, cw, lam)
        assert np.allclose(lr.LinearPdf(d), oracle.linear_pdf(ref, theta), rtol=1e-12, atol=1e-13)
        assert np.allclose(lr.LogPdf(d), oracle.log_pdf(ref, theta), rtol=1e-12, atol=1e-13)
        close_g(lr.Gradient(None, d), oracle.gradient(ref, y, theta, cw, lam))
        lo, olo = lr.Loss(d), oracle.loss(ref, y, theta, cw, lam)
        assert abs(lo - olo) <= RTOL_LOSS * abs(olo)
    with pytest.raises(K.KmerLrError):       # panic("internal error") on a theta of the wrong length
        K.logisticRegression(np.zeros(d.m), (1, 1)).LinearPdf(d)


@pytest.mark.parametrize("M,N,L,flags", [
    (1, 8, 500, dict(revcomp=True)), (2, 6, 120, dict()), (3, 7, 333, dict(complement=True)),
    (4, 9, 64, dict(reverse=True)), (1, 8, 700, dict(revcomp=True)), (6, 10, 200, dict(revcomp=True)),
])
def test_matrix_free_pass_equals_csr_pass(K, oracle, M, N, L, flags):
    """count matrices straight from the extraction take the matrix-free pass (positions instead of stored
    entries); it must agree with the CSR kernel and the oracle -- ragged rows, invalid bases, frozen subsets"""
    from kmerlr_b200 import synth
    buf, off, y = synth.training_set(70, 58, L)
    buf = buf.copy()
    buf[off[3] + 10] = ord("N")
    buf[off[5]:off[5] + min(L, 40)] = ord("n")
    buf[off[7]:off[8]] = ord("A")                      # a homopolymer row
    kc, oc = K.NewKmerCounter(M, N, **flags), oracle.make_config(M, N, **flags)
    d = K.compile_test_data(None, kc, None, None, True, False, (buf, off))
    ref = oracle.extract(oc, (buf, off))
    d.SetLabels(y)
    rng = np.random.default_rng(M * 100 + N)
    theta = rng.normal(scale=0.02, size=d.m + 1)
    theta[rng.integers(1, d.m + 1, size=d.m // 3)] = 0.0
    cw = (0.8, 1.3)
    lr = K.logisticRegression(theta, cw, 0.0)
    K.option("implicit", 1)
    g1, l1 = lr.Gradient(None, d), lr.Loss(d)
    K.option("implicit", 0)
    try:
        g0, l0 = lr.Gradient(None, d), lr.Loss(d)
    finally:
        K.option("implicit", 1)
    og = oracle.gradient(ref, y, theta, cw)
    close_g(g1, og)
    close_g(g0, og)
    assert np.max(np.abs(g1 - g0)) <= 1e-13 * np.max(np.abs(og))
    assert abs(l1 - l0) <= 1e-13 * abs(l0)
    # frozen subset of the classes: absent classes contribute nothing
    k, code = d.Kmers()
    sub = (k[::2], code[::2])
    ds = K.compile_test_data(None, kc, sub, None, True, False, (buf, off))
    ds.SetLabels(y)
    refs = oracle.extract(oc, (buf, off), frozen=sub)
    ths = rng.normal(scale=0.02, size=ds.m + 1)
    close_g(K.logisticRegression(ths, cw).Gradient(None, ds), oracle.gradient(refs, y, ths, cw))
    # proximal-gradient iterations on the full space follow the oracle's
    est = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=0.0, MaxIterations=5)
    est.Theta = np.zeros(d.m + 1)
    est.ClassWeights = np.array(cw)
    est.estimate_proximal(d, 1e-3)
    oth, _, _ = oracle.proxgrad(ref, y, np.zeros(ref.m + 1), cw, lam=1e-3, epsilon=0.0, epsilon_loss=0.0, max_iter=5)
    assert np.allclose(est.Theta, oth, rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize("M,N,L,flags", [
    (1, 10, 200, dict(revcomp=True)), (1, 8, 500, dict(revcomp=True)), (4, 9, 150, dict()), (6, 10, 90, dict(reverse=True)),
    (1, 5, 300, dict(revcomp=True)), (2, 7, 260, dict(complement=True)), (7, 9, 333, dict(revcomp=True)),
])
def test_binarized_matrix_free_pass_equals_stored_pass(K, oracle, M, N, L, flags):
    """binarized matrices straight from the extraction: the table levels as per-row class bitmaps, the levels
    k >= 6 matrix-free with one correction per repeat of a class inside a row (events left by the extraction).
    Against the pass over the stored rows and the oracle -- ragged rows, invalid bases, low-complexity rows
    (nearly every instance a repeat), frozen subsets, unobserved classes (renumbered columns)"""
    from kmerlr_b200 import synth
    buf, off, y = synth.training_set(70, 58, L)
    buf = buf.copy()
    buf[off[3] + 10] = ord("N")
    buf[off[5]:off[5] + min(L, 40)] = ord("n")
    buf[off[7]:off[8]] = ord("A")                      # a homopolymer row
    buf[off[9]:off[10]] = np.frombuffer(b"AC" * L, dtype=np.uint8)[:L]
    buf[off[11]:off[12]] = np.frombuffer(b"ACGTTGCA" * L, dtype=np.uint8)[:L]
    buf[off[13]:off[13] + L // 2] = buf[off[13] + L // 2:off[13] + 2 * (L // 2)]     # two equal halves: every k-mer twice
    kc, oc = K.NewKmerCounter(M, N, **flags), oracle.make_config(M, N, binarize=True, **flags)
    d = K.compile_test_data(None, kc, None, None, True, True, (buf, off))
    ref = oracle.extract(oc, (buf, off))
    rp, col, val = d.rows()
    orp, ocol, _ = ref.rows()
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol) and np.all(val == 1.0)
    d.SetLabels(y)
    rng = np.random.default_rng(M * 100 + N)
    cw = (0.8, 1.3)
    for super_len in (-1, 0, 11):
        K.option("super_len", super_len)
        try:
            theta = rng.normal(scale=0.02, size=d.m + 1)
            theta[rng.integers(1, d.m + 1, size=d.m // 3)] = 0.0
            lr = K.logisticRegression(theta, cw, 0.0)
            K.option("implicit", 1)
            g1, l1 = lr.Gradient(None, d), lr.Loss(d)
            K.option("implicit", 0)
            try:
                g0, l0 = lr.Gradient(None, d), lr.Loss(d)
            finally:
                K.option("implicit", 1)
            og = oracle.gradient(ref, y, theta, cw)
            close_g(g1, og)
            close_g(g0, og)
            assert np.max(np.abs(g1 - g0)) <= 1e-13 * np.max(np.abs(og))
            assert abs(l1 - l0) <= 1e-13 * abs(l0)
            assert abs(l1 - oracle.loss(ref, y, theta, cw)) <= RTOL_LOSS * abs(l0)
        finally:
            K.option("super_len", -1)
    # identical columns get identical bits on the matrix-free path too (integer sums)
    g = K.logisticRegression(np.zeros(d.m + 1)).Gradient(None, d)[1:]
    X = ref.dense()
    groups = {}
    for j in range(d.m):
        groups.setdefault(X[:, j].tobytes(), []).append(j)
    for cols in groups.values():
        assert len({g[j] for j in cols}) == 1
    # frozen subset of the classes: absent classes contribute nothing, their repeats leave no events
    k, code = d.Kmers()
    sub = (k[::2], code[::2])
    ds = K.compile_test_data(None, kc, sub, None, True, True, (buf, off))
    ds.SetLabels(y)
    refs = oracle.extract(oc, (buf, off), frozen=sub)
    ths = rng.normal(scale=0.02, size=ds.m + 1)
    close_g(K.logisticRegression(ths, cw).Gradient(None, ds), oracle.gradient(refs, y, ths, cw))
    # proximal-gradient iterations on the full space follow the oracle's
    est = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=0.0, MaxIterations=5)
    est.Theta = np.zeros(d.m + 1)
    est.ClassWeights = np.array(cw)
    est.estimate_proximal(d, 1e-3)
    oth, _, _ = oracle.proxgrad(ref, y, np.zeros(ref.m + 1), cw, lam=1e-3, epsilon=0.0, epsilon_loss=0.0, max_iter=5)
    assert np.allclose(est.Theta, oth, rtol=1e-9, atol=1e-13)


def test_matrix_free_pass_long_rows_and_super_lengths(K, oracle):
    """rows too long for the register-cached variants (groups decoded twice), every super k-mer length"""
    from kmerlr_b200 import synth
    buf, off, y = synth.training_set(20, 20, 5000)
    buf = buf.copy()
    buf[off[2] + 2500] = ord("N")
    for binarize, M, N in ((False, 1, 8), (True, 1, 8), (False, 3, 6)):
        kc, oc = K.NewKmerCounter(M, N, revcomp=True), oracle.make_config(M, N, revcomp=True, binarize=binarize)
        d = K.compile_test_data(None, kc, None, None, True, binarize, (buf, off))
        ref = oracle.extract(oc, (buf, off))
        d.SetLabels(y)
        theta = np.random.default_rng(3).normal(scale=0.005, size=d.m + 1)
        og = oracle.gradient(ref, y, theta)
        for super_len in (0, 8, 9, 10, 11, -1):
            K.option("super_len", super_len)
            try:
                close_g(K.logisticRegression(theta).Gradient(None, d), og)
            finally:
                K.option("super_len", -1)


def test_kmers5_standardizer_golden_and_transforms(K, oracle, fixtures):
    """kmerLr_test.go:155-190 through the C ABI: k = 2..6 revcomp counts + standardizer, Loss at the golden theta
    with lambda = 4.460029 is the reference's 1.107745182633717.  The rows stay sparse counts in HBM, the
    transform is a reparameterisation of theta (api.Transform); all four transforms against the numpy
    restatement of kmerLr_transform.go, which densifies like the reference does"""
    d, ref, y = build(K, oracle, fixtures, 2, 6, revcomp=True)
    t = K.TransformFull().Fit(d, "standardizer")
    ooff, osc = oracle.fit_transform(ref, "standardizer")
    assert np.allclose(t.Offset, ooff, rtol=1e-13, atol=0.0) and np.allclose(t.Scale, osc, rtol=1e-12, atol=0.0)
    names = ref.class_names()
    sel = [0, names.index("aaaatt|aatttt") + 1, names.index("caggag|ctcctg") + 1]
    assert sel == [0, 703, 1673]
    rd = K.select_data(d, sel)
    rd.SetLabels(y)
    theta = [5.552570741538388e-05, -0.00772452196477929, 0.09287154394711336]       # :166-174
    lr = K.logisticRegression(theta, (1.0, 1.0), 4.460029e+00, Transform=t.Select(sel))
    assert abs(lr.Loss(rd) - 1.107745182633717) <= 1e-13                               # :186
    rng = np.random.default_rng(12)
    cw = (0.8, 1.3)
    for kind in ("standardizer", "variance-scaler", "max-abs-scaler", "mean-scaler"):
        t = K.TransformFull().Fit(d, kind)
        oo, os_ = oracle.fit_transform(ref, kind)
        th = rng.normal(scale=0.01, size=d.m + 1)
        th[rng.integers(1, d.m + 1, size=d.m // 2)] = 0.0
        lr = K.logisticRegression(th, cw, 0.01, Transform=t)
        olo = oracle.transformed_loss(ref, y, th, oo, os_, cw, 0.01)
        assert abs(lr.Loss(d) - olo) <= 1e-11 * abs(olo)
        close_g(lr.Gradient(None, d), oracle.transformed_gradient(ref, y, th, oo, os_, cw, 0.01))
        z = th[0] + oracle._transformed_dense(ref, oo, os_) @ th[1:]
        assert np.allclose(lr.LinearPdf(d), z, rtol=1e-10, atol=1e-11)
    with pytest.raises(K.KmerLrError):
        K.TransformFull().Fit(d, "no-such-transform")
    # first leapfrog epoch of TestKmers5 (--lambda-auto=2 with the standardizer): the two classes the reference's
    # test finds in the model (kmerLr_test.go:175-184) and the lambda of SURVEY section 0
    t = K.TransformFull().Fit(d, "standardizer")
    fs = K.featureSelector((1.0, 1.0), False, 2, d.m, tie=K.TIE_GO118, Transform=t)
    selection, lam, ok = fs.Select(d, 0.0, [], [], 0.0)
    assert ok and selection.sel.tolist() == [0, 703, 1673]
    assert abs(lam - 0.3345505506971979) <= 1e-12


def test_kmers5_leapfrog_under_standardizer_end_to_end(K, oracle, fixtures):
    """TestKmers5 (kmerLr_test.go:155-190) end to end through the C ABI: `learn --lambda-auto=2 --epsilon=0
    --epsilon-loss=1e-10 --revcomp --data-transform=standardizer 2 6`.  The selection gradient is taken under the
    transform, every reduced data set goes through Transform.Apply before the proximal-gradient solver
    (kmerLr_estimator.go:148).  Against the oracle path: same epochs, lambda sequence, iteration counts, theta;
    against the reference's goldens: Features [[0,0],[1,1]] = the two classes, theta within 1e-4 (two different
    early-stopped solvers), loss_ with lambda = 4.460029 within 1e-4."""
    d, ref, y = build(K, oracle, fixtures, 2, 6, revcomp=True)
    t = K.TransformFull().Fit(d, "standardizer")
    est = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=1e-10, MaxIterations=10 ** 7, tie=K.TIE_GO118)
    epochs = est.estimate_loop(d, 2, transform=t)
    oo, os_ = oracle.fit_transform(ref, "standardizer")
    o = oracle.estimate_loop_transformed(ref, y, (1.0, 1.0), 2, oo, os_, tie=oracle.TIE_GO118, epsilon=0.0,
                                         epsilon_loss=1e-10, max_iter=10 ** 7)
    names = ref.class_names()
    assert est.active_idx.tolist() == o["active_idx"].tolist() == [names.index("aaaatt|aatttt") + 1, names.index("caggag|ctcctg") + 1]
    assert epochs == len(o["lambdas"]) and [p[1] for p in est.path] == list(o["iters"])
    assert abs(est.path[0][0] - 0.3345505506971979) <= 1e-12
    assert np.allclose([p[0] for p in est.path], o["lambdas"], rtol=1e-9, atol=0.0)
    assert np.allclose(est.Theta, o["theta"], rtol=1e-7, atol=1e-12)            # north_star: <= 1e-5
    golden = np.array([5.552570741538388e-05, -0.00772452196477929, 0.09287154394711336])   # :166-174
    # (the golden is an early-stopped SAGA iterate, ours an ISTA iterate stopped by the same loss rule: 8e-5 apart;
    # the tight optimum of the objective is 2.1e-5 from the golden)
    assert np.max(np.abs(est.Theta - golden)) <= 1e-4
    # loss_ of the resulting model (kmerLr_loss.go:53-60): frozen counter + Features + Transform.Select + Loss
    sel = np.concatenate([[0], est.active_idx])
    rd = K.select_data(d, sel)
    rd.SetLabels(y)
    lo = K.logisticRegression(est.Theta, (1.0, 1.0), 4.460029e+00, Transform=est.Transform).Loss(rd)
    assert abs(lo - 1.107745182633717) <= 1e-4                                   # :186
    # Transform.Apply itself: dense rows (v - offset) scale, against the numpy restatement
    td = est.Transform.Apply(rd)
    rp, col, val = td.rows()
    X = oracle._transformed_dense(oracle.reduce(ref, sel), oo[sel], os_[sel])
    assert td.nnz == X.size and np.array_equal(np.diff(rp), np.full(d.n, 2))
    assert np.allclose(val.reshape(d.n, 2), X, rtol=1e-13, atol=1e-15)
    # a scale-only transform keeps the sparsity
    t2 = K.TransformFull().Fit(d, "max-abs-scaler")
    ts = t2.Select(sel).Apply(rd)
    rp2, col2, val2 = ts.rows()
    rp0, col0, val0 = rd.rows()
    assert np.array_equal(rp2, rp0) and np.array_equal(col2, col0)
    assert np.allclose(val2, val0 * t2.Scale[sel][col0 + 1], rtol=1e-15, atol=0.0)


def test_pair_features_under_a_transform(K, oracle, fixtures):
    """kmerLr_logistic_regression.go:91-107,200-216 + kmerLr_transform.go:90-99,118-127: the transform covers the
    pair features v_a v_b as well (offsets / scales in CoeffIndex order); on the device it stays a
    reparameterisation of theta.  Against a dense numpy restatement on the TestKmers6 fixture."""
    O = oracle
    d, ref, y = build(K, O, fixtures, 2, 6, fg="kmerLr_test_co_fg", bg="kmerLr_test_co_bg", revcomp=True)
    nt = K.CoeffIndex(d.m).Dim()
    rng = np.random.default_rng(8)
    cw = (0.8, 1.3)
    for kind in ("standardizer", "variance-scaler", "max-abs-scaler", "mean-scaler"):
        t = K.TransformFull().Fit(d, kind, cooccurrence=True)
        oo, os_ = O.fit_transform(ref, kind, cooccurrence=True)
        if oo is not None:
            assert len(t.Offset) == nt and np.allclose(t.Offset, oo, rtol=1e-13, atol=0.0)
        fin = np.isfinite(os_)
        assert len(t.Scale) == nt and np.array_equal(np.isfinite(t.Scale), fin)
        assert np.allclose(t.Scale[fin], os_[fin], rtol=1e-11, atol=0.0)
        # pairs that never occur together have an infinite max-abs / mean scale in the reference too: keep theta off them
        theta = np.zeros(nt)
        pick = rng.choice(np.nonzero(fin)[0][1:], size=60, replace=False)
        theta[pick] = rng.normal(scale=0.02, size=60)
        theta[0] = 0.1
        sc = np.where(fin, os_, 0.0)
        lr = K.logisticRegression(theta, cw, 0.01, Cooccurrence=True, Transform=K.Transform(t.Offset, np.where(fin, t.Scale, 0.0)))
        z = theta[0] + O._transformed_dense(ref, oo, sc, True) @ theta[1:]
        assert np.allclose(lr.LinearPdf(d), z, rtol=1e-10, atol=1e-11)
        olo = O.transformed_loss(ref, y, theta, oo, sc, cw, 0.01, cooccurrence=True)
        assert abs(lr.Loss(d) - olo) <= 1e-11 * abs(olo)
        close_g(lr.Gradient(None, d), O.transformed_gradient(ref, y, theta, oo, sc, cw, 0.01, cooccurrence=True))


def test_identical_columns_get_identical_gradients(K, oracle, fixtures):
    """what leapfrog tie handling rests on: the column reduction depends on the column only"""
    d, ref, y = build(K, oracle, fixtures, 2, 6, fg="kmerLr_test_co_fg", bg="kmerLr_test_co_bg", revcomp=True, binarize=True)
    X = ref.dense()
    g = K.logisticRegression(np.zeros(d.m + 1)).Gradient(None, d)[1:]
    groups = {}
    for j in range(d.m):
        groups.setdefault(X[:, j].tobytes(), []).append(j)
    assert any(len(v) > 1 for v in groups.values())
    for cols in groups.values():
        assert len({g[j] for j in cols}) == 1


def test_cooccurrence_gradient_and_select_go118(K, oracle, fixtures):
    """TestKmers6 data (kmerLr_test.go:192-268) on the GPU: 57-way tie, lambda, the golden pair"""
    O = oracle
    d, ref, y = build(K, O, fixtures, 2, 6, fg="kmerLr_test_co_fg", bg="kmerLr_test_co_bg", revcomp=True, binarize=True)
    nt = K.CoeffIndex(d.m).Dim()
    assert nt == 2486
    rng = np.random.default_rng(5)
    theta = np.zeros(nt)
    theta[rng.integers(0, nt, 40)] = rng.normal(scale=0.05, size=40)
    lr = K.logisticRegression(theta, (1, 1), 0.0, Cooccurrence=True)
    assert np.allclose(lr.LinearPdf(d), O.linear_pdf(ref, theta, True), rtol=1e-12, atol=1e-13)
    close_g(lr.Gradient(None, d), O.gradient(ref, y, theta, cooccurrence=True))
    assert abs(lr.Loss(d) - O.loss(ref, y, theta, cooccurrence=True)) < 1e-13
    s = K.featureSelector((1, 1), True, 2, d.m, tie=K.TIE_GO118)
    sel, lam, ok = s.Select(d, 0.0, [], [], 0.0, want_gradient=True)
    r = O.select(ref, y, (1, 1), 2, 0.0, [], [], cooccurrence=True, tie=O.TIE_GO118)
    assert ok and lam == 0.2375 and np.array_equal(sel.b, r["mask"])
    assert np.sum(np.abs(sel.g[1:]) == np.max(np.abs(sel.g[1:]))) == 57
    pairs = [K.CoeffIndex(d.m).Sub2Ind(int(j) - 1) for j in sel.sel[1:]]
    cls = sorted({k for p in pairs for k in p})
    assert [[cls.index(a), cls.index(b)] for a, b in pairs] == [[0, 1], [1, 2]]     # kmerLr_test.go:205-206
    # loss / predictions at the golden theta (kmerLr_test.go:201-204,224,228,248)
    red = sel.Data(d)
    red.SetLabels(y)
    th = np.array([-0.1000970529629098, 0.09995715710821684, 0.09995715710821684])
    lrr = K.logisticRegression(th)
    assert abs(lrr.Loss(red) - 0.644417014007959) < 1e-14
    lp = lrr.LogPdf(red)
    assert np.allclose(lp[:10], -0.6444834689451768, atol=1e-14) and np.allclose(lp[10:], -0.744447612033651, atol=1e-14)
    # the index tie rule agrees with the oracle in the same mode
    s2 = K.featureSelector((1, 1), True, 2, d.m, tie=K.TIE_INDEX)
    sel2, lam2, _ = s2.Select(d, 0.0, [], [], 0.0)
    r2 = O.select(ref, y, (1, 1), 2, 0.0, [], [], cooccurrence=True, tie=O.TIE_INDEX)
    assert lam2 == r2["lam"] and np.array_equal(sel2.b, r2["mask"])


def test_scores_lambda_from_csr(K, oracle, fixtures):
    """README.md:39: first leapfrog lambda of the 16 x 7 example, features {1,6} (scoresLr_test.go:47-56)"""
    X = np.vstack([fixtures["scoresLr_test_fg"], fixtures["scoresLr_test_bg"]])
    y = np.array([1] * 8 + [0] * 8, dtype=np.uint8)
    d = K.from_dense(X)
    d.SetLabels(y)
    s = K.featureSelector((1, 1), False, 2, d.m, tie=K.TIE_GO118)
    sel, lam, ok = s.Select(d, 0.0, [], [], 0.0)
    assert "%e" % lam == "2.496875e+00" and sel.sel.tolist() == [0, 2, 7] and ok
    red = sel.Data(d)
    red.SetLabels(y)
    th = np.array([0.842178566751775, -0.05466291047449, -0.03026279836545])
    assert abs(K.logisticRegression(th, (1, 1), 4.647556e+00).Loss(red) - 0.813659729805629) < 1e-9


def test_reduce_matches_oracle(K, oracle, fixtures):
    d, ref, y = build(K, oracle, fixtures, 1, 5, revcomp=True)
    sel = np.array([0, 1, 3, 10, 50, 200, d.m], dtype=np.int64)
    red = K.select_data(d, sel)
    oref = oracle.reduce(ref, sel)
    assert (red.n, red.m, red.nnz) == (oref.n, oref.m, oref.nnz)
    for x, z in zip(red.rows(), oref.rows()):
        assert np.array_equal(x, z)


@pytest.mark.parametrize("tie", ["go118", "index"])
def test_leapfrog_path_fixtures(K, oracle, fixtures, tie):
    """C1: kmerLr learn on the bundled fixtures, k=1..6 revcomp, --lambda-auto=2,5,10: same selected
    sets, same lambda sequence, theta within 1e-5, loss within 1e-6 of the CPU oracle"""
    O = oracle
    d, ref, y = build(K, O, fixtures, 1, 6, revcomp=True)
    kt, ot = (K.TIE_GO118, O.TIE_GO118) if tie == "go118" else (K.TIE_INDEX, O.TIE_INDEX)
    est = K.KmerLrEstimator(EpsilonLoss=1e-8, tie=kt, MaxIterations=20000)
    oest = O.EstimatorState()
    for N in (2, 5, 10):
        est.path = []
        est.estimate_loop(d, N)
        res = O.estimate_loop(ref, y, (1, 1), N, oest, tie=ot, epsilon_loss=1e-8, max_iter=20000)
        assert len(est.path) == res["epochs"]
        assert np.array_equal(est.active_idx, oest.active_idx)
        lams = np.array([p[0] for p in est.path])
        assert np.allclose(lams, res["lambdas"], rtol=1e-9, atol=0)
        assert abs(lams[0] - res["lambdas"][0]) <= 1e-12 * abs(lams[0])     # first lambda: a function of g(0) only
        assert [p[1] for p in est.path] == res["iters"].tolist()
        assert np.max(np.abs(est.Theta - oest.active_theta)) <= 1e-5 * max(1.0, np.max(np.abs(oest.active_theta)))
    # loss of the final model on the reduced data
    sel = np.concatenate([[0], est.active_idx])
    rm = O.reduce(ref, sel)
    red = K.select_data(d, sel)
    red.SetLabels(y)
    lo = K.logisticRegression(est.Theta, (1, 1), est.L1Reg / d.n).Loss(red)
    olo = O.loss(rm, y, oest.active_theta, (1, 1), est.L1Reg / d.n)
    assert abs(lo - olo) <= 1e-6 * abs(olo)


def test_proxgrad_matches_oracle_iteration_by_iteration(K, oracle):
    """same step size, update, stopping rule (kmerLr_estimator_proximal.go:30-120): same iteration count"""
    from kmerlr_b200 import synth
    O = oracle
    buf, off, y = synth.training_set(300, 300, 200)
    kc, oc = K.NewKmerCounter(1, 5, revcomp=True), O.make_config(1, 5, revcomp=True)
    d = K.compile_training_data(None, kc, None, None, True, False, (buf[:off[300]], off[:301]), (buf[off[300]:], off[300:] - off[300]))
    ref = O.extract(oc, (buf, off))
    cw = K.compute_class_weights(y)
    assert np.allclose(cw, d.class_weights()) and np.allclose(cw, O.class_weights(y))
    step = 0.0
    import ctypes as C
    out = C.c_double()
    from kmerlr_b200 import _lib
    _lib.check(_lib.lib().kmerlr_step_size(d.h, 0.0, 1.0, out))
    step = out.value
    assert step == O.step_size(ref)
    for eps, eps_loss, lam, cap in [(0.0, 1e-7, 0.002, 5000), (1e-4, 0.0, 0.0005, 5000), (0.0, 1e-300, 0.001, 37)]:
        est = K.KmerLrEstimator(Epsilon=eps, EpsilonLoss=eps_loss, MaxIterations=cap)
        est.Theta = np.zeros(d.m + 1)
        iters, delta = est.estimate_proximal(d, lam)
        oth, oit, odelta = O.proxgrad(ref, y, np.zeros(d.m + 1), (1, 1), lam, epsilon=eps, epsilon_loss=eps_loss, max_iter=cap)
        assert iters == oit
        assert np.max(np.abs(est.Theta - oth)) <= 1e-9
        assert np.array_equal(est.Theta == 0.0, oth == 0.0)
        assert abs(delta - odelta) <= 1e-6 * max(abs(odelta), 1e-300)


def test_persistent_reduced_iterations_equal_per_launch(K, oracle):
    """reduced matrices on one GPU: the cooperative many-iterations launch (grid barriers) must give the
    iterates of the one-launch-per-iteration path bit for bit, whatever stops the loop"""
    from kmerlr_b200 import synth
    buf, off, y = synth.training_set(700, 650, 150)
    kc = K.NewKmerCounter(1, 6, revcomp=True)
    d = K.compile_test_data(None, kc, None, None, True, False, (buf, off))
    sel = np.unique(np.concatenate([[0], np.linspace(1, d.m, 40).astype(np.int64)]))
    rd = K.select_data(d, sel)
    rd.SetLabels(y)
    cw = np.array([0.8, 1.3])
    try:
        for eps, eps_loss, lam, cap in [(0.0, 0.0, 1e-3, 37), (0.0, 0.0, 1e-4, 4500), (1e-6, 0.0, 1e-3, 100000),
                                        (0.0, 1e-9, 1e-3, 100000), (0.0, 0.0, 1e-3, 0), (0.0, 0.0, 1e-3, 1)]:
            res = []
            for mode in (0, 1):
                K.option("persistent", mode)
                est = K.KmerLrEstimator(Epsilon=eps, EpsilonLoss=eps_loss, MaxIterations=cap)
                est.Theta = np.zeros(len(sel)); est.ClassWeights = cw
                it, delta = est.estimate_proximal(rd, lam)
                res.append((it, delta, est.Theta.copy()))
            assert res[0][0] == res[1][0], (cap, res[0][0], res[1][0])
            assert res[0][1] == res[1][1] or (np.isnan(res[0][1]) and np.isnan(res[1][1]))
            assert np.array_equal(res[0][2], res[1][2])
            if cap in (37, 4500):
                assert res[0][0] == cap
    finally:
        K.option("persistent", 1)
        rd.free(); d.free()


def test_reduced_solver_long_rows_same_bits(K, oracle):
    """reduced matrices with long rows run on the sliced + column-major views (packed counts; real values unpacked) or on
    the block-local view (one grid barrier per iteration; counts only, real values fall back to the views):
    the iterates must be those of the row-wise solver bit for bit, and those of the oracle to rounding"""
    from kmerlr_b200 import synth
    O = oracle
    buf, off, y = synth.training_set(900, 833, 180)
    kc, oc = K.NewKmerCounter(1, 6, revcomp=True), O.make_config(1, 6, revcomp=True)
    d = K.compile_test_data(None, kc, None, None, True, False, (buf, off))
    ref = O.extract(oc, (buf, off))
    # the classes of k <= 3 are in (almost) every row: 44 dense columns + 30 spread ones
    sel = np.unique(np.concatenate([[0], np.arange(1, 45), np.linspace(50, d.m, 30).astype(np.int64)]))
    rd = K.select_data(d, sel)
    rd.SetLabels(y)
    assert rd.nnz >= 30 * rd.n
    scaled = K.Transform(Scale=np.concatenate([[1.0], 1.0 / (1.0 + np.arange(len(sel) - 1) % 7)])).Apply(rd)   # real values
    scaled.SetLabels(y)
    cw = np.array([0.8, 1.3])
    try:
        for data in (rd, scaled):
            for eps, eps_loss, lam, cap in [(0.0, 0.0, 1e-3, 61), (1e-6, 0.0, 1e-3, 100000), (0.0, 1e-9, 2e-3, 100000), (0.0, 0.0, 1e-3, 0)]:
                res = []
                for mode in (0, 1, 2):              # rows; sliced + column-major views; block-local view (counts only)
                    K.option("small_long", mode)
                    est = K.KmerLrEstimator(Epsilon=eps, EpsilonLoss=eps_loss, MaxIterations=cap)
                    est.Theta = np.zeros(len(sel)); est.ClassWeights = cw
                    it, delta = est.estimate_proximal(data, lam)
                    res.append((it, delta, est.Theta.copy()))
                for other in (1, 2):
                    assert res[0][0] == res[other][0], (cap, res[0][0], res[other][0])
                    assert res[0][1] == res[other][1] or (np.isnan(res[0][1]) and np.isnan(res[other][1]))
                    assert np.array_equal(res[0][2], res[other][2])
                if data is rd and cap == 61:
                    rm = O.reduce(ref, sel)
                    oth, oit, _ = O.proxgrad(rm, y, np.zeros(len(sel)), tuple(cw), lam, epsilon=eps, epsilon_loss=eps_loss, max_iter=cap)
                    assert oit == res[1][0] and np.max(np.abs(res[1][2] - oth)) <= 1e-9
    finally:
        K.option("small_long", -1)
        scaled.free(); rd.free(); d.free()


def test_coordinate_estimator_matches_restatement(K, oracle):
    """estimate_coordinate (kmerLr_estimator_coordinate.go:31-139, slices de-aliased) on reduced matrices: the
    CUDA path (fixed-point Gram matrix, one-block sweeps) against the numpy restatement -- same number of
    sweeps, same theta.  The reference holds no golden for this function (it is never called)."""
    from kmerlr_b200 import synth
    O = oracle
    buf, off, y = synth.training_set(900, 700, 150)
    kc, oc = K.NewKmerCounter(1, 6, revcomp=True), O.make_config(1, 6, revcomp=True)
    d = K.compile_test_data(None, kc, None, None, True, False, (buf, off))
    ref = O.extract(oc, (buf, off))
    sel = np.unique(np.concatenate([[0], np.linspace(1, d.m, 25).astype(np.int64)]))
    rd = K.select_data(d, sel); rd.SetLabels(y)
    rr = O.reduce(ref, sel)
    cwh = np.array([0.9, 1.2])
    try:
        for eps, eps_loss, l1, l2, cap in [(0.0, 0.0, 0.0, 0.0, 25), (1e-4, 0.0, 2.0, 0.0, 5000), (1e-3, 0.0, 20.0, 0.5, 5000),
                                           (0.0, 1e-6, 1.0, 0.0, 5000), (0.0, 0.0, 1.0, 0.0, 0)]:
            est = K.KmerLrEstimator(Epsilon=eps, EpsilonLoss=eps_loss, L2Reg=l2, MaxIterations=cap)
            est.Theta = np.zeros(len(sel)); est.ClassWeights = cwh; est.L1Reg = l1
            sweeps, delta = est.estimate_coordinate(rd)
            oth, osw, odelta = O.coordinate(rr, y, np.zeros(len(sel)), cwh, l1, l2, eps, eps_loss, cap)
            assert sweeps == osw, (eps, eps_loss, l1, cap, sweeps, osw)
            assert np.allclose(est.Theta, oth, rtol=1e-8, atol=1e-10)
            assert np.array_equal(est.Theta == 0.0, oth == 0.0)
            assert abs(delta - odelta) <= 1e-6 * max(abs(odelta), 1e-12)
        # warm start from a fitted theta; binarized (VAL_ONE) matrix
        kb = K.NewKmerCounter(1, 6, revcomp=True, binarize=True)
        db = K.compile_test_data(None, kb, None, None, True, True, (buf, off))
        rb = K.select_data(db, sel); rb.SetLabels(y)
        rrb = O.reduce(O.extract(O.make_config(1, 6, revcomp=True, binarize=True), (buf, off)), sel)
        est = K.KmerLrEstimator(Epsilon=1e-5, EpsilonLoss=0.0, MaxIterations=300)
        est.Theta = oth.copy(); est.L1Reg = 0.5
        sweeps, _ = est.estimate_coordinate(rb)
        oth2, osw2, _ = O.coordinate(rrb, y, oth.copy(), (1.0, 1.0), 0.5, 0.0, 1e-5, 0.0, 300)
        assert sweeps == osw2 and np.allclose(est.Theta, oth2, rtol=1e-8, atol=1e-10)
        rb.free(); db.free()
        # not a reduced matrix: refused loudly
        with pytest.raises(K.KmerLrError):
            est = K.KmerLrEstimator(MaxIterations=3); est.Theta = np.zeros(d.m + 1)
            d.SetLabels(y); est.estimate_coordinate(d)
    finally:
        rd.free(); d.free()


def test_proxgrad_is_resumable_in_slices(K, oracle):
    """the Go hook (trace / verbose / adaptive step, kmerLr_estimator_hook.go:46-99) wants to see theta every k
    iterations: kmerlr_proxgrad run in slices of k iterations (theta and hook_state carried over) must give the
    iterates of one long call bit for bit, full-space path and reduced path, and stop at the same place"""
    from kmerlr_b200 import synth
    buf, off, y = synth.training_set(400, 400, 150)
    kc = K.NewKmerCounter(1, 6, revcomp=True)
    d = K.compile_test_data(None, kc, None, None, True, False, (buf, off))
    d.SetLabels(y)
    sel = np.unique(np.concatenate([[0], np.linspace(1, d.m, 30).astype(np.int64)]))
    rd = K.select_data(d, sel); rd.SetLabels(y)
    try:
        for data, ntheta, lam in [(d, d.m + 1, 2e-3), (rd, len(sel), 1e-3)]:
            one = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=1e-300, MaxIterations=30)
            one.Theta = np.zeros(ntheta)
            it1, _ = one.estimate_proximal(data, lam)
            sl = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=1e-300, MaxIterations=10)
            sl.Theta = np.zeros(ntheta)
            total = 0
            for _ in range(3):
                it, _ = sl.estimate_proximal(data, lam)
                total += it
            assert (it1, total) == (30, 30)
            assert np.array_equal(one.Theta, sl.Theta)
            assert np.array_equal(one.hook_state, sl.hook_state)
            # a loss-based stop inside a slice stops the sliced run at the same iteration
            one = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=1e-5, MaxIterations=5000)
            one.Theta = np.zeros(ntheta)
            it1, _ = one.estimate_proximal(data, lam)
            sl = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=1e-5, MaxIterations=7)
            sl.Theta = np.zeros(ntheta)
            total = 0
            while True:
                it, _ = sl.estimate_proximal(data, lam)
                total += it
                if it < 7 or total > 6000:
                    break
            assert 0 < it1 < 5000 and total == it1
            assert np.array_equal(one.Theta, sl.Theta)
    finally:
        rd.free(); d.free()
