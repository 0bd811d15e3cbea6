"""The oracle against every golden vector the reference's own tests hold for the hot path
(SURVEY.md 8c).  CPU only.  Values are transcribed from kmerLr_test.go / scoresLr_test.go / README.md
(file:line next to each)."""
import numpy as np

from conftest import cat


def test_kmers1_index_name_count_pins(oracle, fixtures):
    """kmerLr_test.go:30-68 (gapped alphabet, k=4..8, revcomp): 6 (index, class, count) pins, 4 pair products."""
    O = oracle
    cfg = O.make_config(4, 8, revcomp=True, alphabet="gapped-nucleotide")
    buf, off, _ = cat(fixtures, "kmerLr_test", "kmerLr_test")
    data1 = O.extract(cfg, (buf, off))
    assert data1.n == 4 and data1.m == 58308
    names = data1.class_names()
    d = data1.rows()
    row0 = dict(zip(d[1][d[0][0]:d[0][1]].tolist(), d[2][d[0][0]:d[0][1]].tolist()))
    pins = [(4671, "gntanc|gntanc", 3), (4672, "gntcaa|ttganc", 0), (5068, "aaagaaa|tttcttt", 1),
            (5486, "aagannt|anntctt", 7), (19270, "aacgcgna|tncgcgtt", 1), (57071, "tgaatgca|tgcattca", 1)]
    for idx, name, count in pins:                       # kmerLr_test.go:40-43
        assert names[idx] == name
        assert row0.get(idx, 0) == count
    # kmerLr_test.go:38-66: explicit feature list = all singles + 4 pairs, frozen class list
    m = data1.m
    feats = [(i, i) for i in range(m)] + [(4671, 4672), (5068, 5486), (19270, 57071), (4671, 5486)]
    data2 = O.extract(cfg, (buf, off), frozen=data1.classes(), features=feats)
    d1, d2 = data1.dense(), data2.dense()
    assert np.array_equal(d1[:2, :], d2[:2, :m])       # :47-54
    assert d2[0, m + 0] == 0 and d2[0, m + 1] == 7 and d2[0, m + 2] == 1 and d2[0, m + 3] == 21   # :55-66


def test_kmers2_frozen_equals_unfrozen(oracle, fixtures):
    """kmerLr_test.go:70-97"""
    O = oracle
    cfg = O.make_config(4, 8, revcomp=True, alphabet="gapped-nucleotide")
    buf, off, _ = cat(fixtures, "kmerLr_test", "kmerLr_test")
    a = O.extract(cfg, (buf, off))
    b = O.extract(cfg, (buf, off), frozen=a.classes())
    assert a.m == b.m
    for x, y in zip(a.rows(), b.rows()):
        assert np.array_equal(x, y)


def test_kmers3_no_explicit_zeros_and_faithful_convert(oracle, fixtures):
    """kmerLr_test.go:117-123; the O(n*m) convert_counts walk gives the same rows"""
    O = oracle
    cfg = O.make_config(8, 8, revcomp=True)
    buf, off, _ = cat(fixtures, "kmerLr_test_fg", "kmerLr_test_bg")
    a = O.extract(cfg, (buf, off))
    b = O.extract(cfg, (buf, off), faithful=True)
    assert np.all(a.rows()[2] != 0)
    for x, y in zip(a.rows(), b.rows()):
        assert np.array_equal(x, y)


def test_coeff_index_round_trip(oracle):
    """kmerLr_coefficients_index.go:26-54"""
    O = oracle
    for n in (1, 2, 3, 7, 70, 2772):
        assert O.coeff_dim(n) == (n + 1) * n // 2 + 1
        seen = set()
        rng = np.random.default_rng(n)
        pairs = [(i, i) for i in range(min(n, 50))] + [tuple(sorted(rng.integers(0, n, 2))) for _ in range(200)]
        for k1, k2 in pairs:
            j = O.ind2sub(n, int(k1), int(k2))
            assert 1 <= j < O.coeff_dim(n)
            assert O.sub2ind(n, j - 1) == (k1, k2)
            seen.add(j)
    n = 9
    assert sorted(O.ind2sub(n, a, b) for a in range(n) for b in range(a, n)) == list(range(1, O.coeff_dim(n)))


def test_scores1_lambda_features_loss(oracle, fixtures):
    """README.md:39 lambda=2.496875e+00; scoresLr_test.go:28-61 features {1,6}, loss at the golden theta"""
    O = oracle
    X = np.vstack([fixtures["scoresLr_test_fg"], fixtures["scoresLr_test_bg"]])
    y = np.array([1] * 8 + [0] * 8)
    m = O.from_dense(X)
    r = O.select(m, y, (1, 1), 2, 0.0, [], [], tie=O.TIE_GO118)
    assert "%e" % r["lam"] == "2.496875e+00"
    assert np.nonzero(r["mask"])[0].tolist() == [0, 2, 7]          # Index 1 and 6 (+1 for the bias)
    rm = O.reduce(m, [0, 2, 7])
    theta = [0.842178566751775, -0.05466291047449, -0.03026279836545]    # scoresLr_test.go:38-46
    assert abs(O.loss(rm, y, theta, lam=4.647556e+00) - 0.813659729805629) < 1e-9   # :57
    # a tight solve of the same objective: the golden is an early-stopped SAGA iterate
    est = O.EstimatorState()
    res = O.estimate_loop(m, y, (1, 1), 2, est, epsilon_loss=1e-13, max_iter=3000000)
    assert res["epochs"] == 1 and est.active_idx.tolist() == [2, 7]
    assert np.allclose(est.active_theta[1:], theta[1:], atol=5e-4)
    assert abs(est.active_theta[0] - theta[0]) < 1e-2


def test_scores2_pair_feature(oracle, fixtures):
    """scoresLr_test.go:63-119: co-occurrence on dense scores selects the pair (1,2)"""
    O = oracle
    X = np.vstack([fixtures["scoresLr_test_co_fg"], fixtures["scoresLr_test_co_bg"]])
    y = np.array([1] * 8 + [0] * 8)
    m = O.from_dense(X)
    r = O.select(m, y, (1, 1), 1, 0.0, [], [], cooccurrence=True, tie=O.TIE_GO118)
    sel = np.nonzero(r["mask"])[0]
    assert len(sel) == 2 and O.sub2ind(m.m, int(sel[1]) - 1) == (1, 2)
    rm = O.reduce(m, sel)
    theta = [-3.6698336905701286e-06, 0.000247759511599005]
    assert abs(O.loss(rm, y, theta) - 0.6699931965725273) < 1e-4
    w = [-0.6662065007234194, -0.6477490377012717, -0.6539634473226117, -0.6063887787494636,
         -0.5995933447812485, -0.4461679042826637, -0.6290620668667339, -0.6750691863398406]
    assert np.allclose(O.log_pdf(rm, theta)[:8], w, atol=1e-4)


def test_kmers6_tie_group_go118_order(oracle, fixtures):
    """kmerLr_test.go:192-268: 57-way tie at |g| = 0.25, lambda 0.2375, Features [[0,1],[1,2]],
    loss and predictions at the golden theta to every printed digit"""
    O = oracle
    cfg = O.make_config(2, 6, revcomp=True, binarize=True)
    buf, off, y = cat(fixtures, "kmerLr_test_co_fg", "kmerLr_test_co_bg")
    m6 = O.extract(cfg, (buf, off))
    assert m6.m == 70 and O.coeff_dim(m6.m) == 2486
    r = O.select(m6, y, (1, 1), 2, 0.0, [], [], cooccurrence=True, tie=O.TIE_GO118)
    g = r["g"][1:]
    assert np.sum(np.abs(g) == 0.24999999999999997) == 57 and np.max(np.abs(g)) == 0.24999999999999997
    assert r["lam"] == 0.2375
    sel = np.nonzero(r["mask"])[0]
    names = m6.class_names()
    pairs = [tuple(names[k] for k in O.sub2ind(m6.m, int(j) - 1)) for j in sel[1:]]
    assert pairs == [("ca|tg", "agag|ctct"), ("agag|ctct", "ggaga|tctcc")]
    # reduced class list (ca, agag, ggaga) -> Features [[0,1],[1,2]]  (kmerLr_test.go:205-206)
    cls = sorted({k for j in sel[1:] for k in O.sub2ind(m6.m, int(j) - 1)})
    feats = [[cls.index(a), cls.index(b)] for a, b in (O.sub2ind(m6.m, int(j) - 1) for j in sel[1:])]
    assert feats == [[0, 1], [1, 2]]
    rm = O.reduce(m6, sel)
    theta = [-0.1000970529629098, 0.09995715710821684, 0.09995715710821684]        # :201-204
    assert O.loss(rm, y, theta) == 0.644417014007959                                # :224
    lp = O.log_pdf(rm, theta)
    assert np.all(lp[:10] == -0.6444834689451768) and np.all(lp[10:] == -0.744447612033651)   # :228,248
    # the |g| desc / index asc rule picks a different pair: the goldens encode the legacy sort
    r2 = O.select(m6, y, (1, 1), 2, 0.0, [], [], cooccurrence=True, tie=O.TIE_INDEX)
    assert not np.array_equal(r2["mask"], r["mask"]) and r2["lam"] == 0.2375
    # tight solve: closed form theta0 = -logit(1-2 lambda), theta1 = theta2 = +0.1000834
    est = O.EstimatorState()
    O.estimate_loop(m6, y, (1, 1), 2, est, cooccurrence=True, epsilon_loss=1e-14, max_iter=1000000)
    assert np.allclose(est.active_theta, [-0.1000834, 0.1000834, 0.1000834], atol=2e-6)
    assert abs(O.loss(rm, y, est.active_theta) - 0.644417014007959) < 1e-4


def test_kmers5_nucleotide_counts_and_first_lambda(oracle, fixtures):
    """kmerLr_test.go:155-190 without the standardizer (a 'next' row): class count and nnz structure"""
    O = oracle
    cfg = O.make_config(2, 6, revcomp=True)
    buf, off, y = cat(fixtures, "kmerLr_test_fg", "kmerLr_test_bg")
    m = O.extract(cfg, (buf, off))
    assert m.n == 22 and m.m == 2660        # m = 2 660 observed classes
    names = m.class_names()
    assert names.index("aaaatt|aatttt") == 702 and names.index("caggag|ctcctg") == 1672


def test_kmers5_standardizer_loss_golden(oracle, fixtures):
    """kmerLr_test.go:155-190: k = 2..6 revcomp counts + standardizer; Loss at the golden theta with
    lambda = 4.460029 is 1.107745182633717 -- reproduced to the last digit (counting, transform and loss)"""
    O = oracle
    cfg = O.make_config(2, 6, revcomp=True)
    buf, off, y = cat(fixtures, "kmerLr_test_fg", "kmerLr_test_bg")
    m = O.extract(cfg, (buf, off))
    names = m.class_names()
    c1, c2 = names.index("aaaatt|aatttt"), names.index("caggag|ctcctg")
    offset, scale = O.fit_transform(m, "standardizer")
    sel = [0, c1 + 1, c2 + 1]
    rm = O.reduce(m, sel)
    theta = [5.552570741538388e-05, -0.00772452196477929, 0.09287154394711336]       # :166-174
    assert O.transformed_loss(rm, y, theta, offset[sel], scale[sel], lam=4.460029e+00) == 1.107745182633717   # :186


def test_go118_sort_matches_reference_structure(oracle):
    """small arrays go through the shell pass + insertion sort: equal keys keep a defined order"""
    O = oracle
    x = np.array([0.5, -0.5, 0.25, 0.5, -0.25, 0.0, 0.5, 1.0, -1.0, 0.125, 0.5, 0.75])
    v, i = O.nlargest_abs(x, O.TIE_GO118)
    assert np.all(np.diff(np.abs(v)) <= 0) and sorted(i.tolist()) == list(range(len(x)))
    assert np.array_equal(np.abs(x[i]), np.abs(v))
    v2, i2 = O.nlargest_abs(x, O.TIE_INDEX)
    assert i2.tolist() == [7, 8, 11, 0, 1, 3, 6, 10, 2, 4, 9, 5]
    rng = np.random.default_rng(0)
    y = np.round(rng.normal(size=5000), 1)
    v, i = O.nlargest_abs(y, O.TIE_GO118)
    assert np.all(np.diff(np.abs(v)) <= 0) and np.array_equal(np.sort(i), np.arange(5000))


def test_window_scoring_quirks(oracle):
    """kmerLr_predict_genomic.go:152-159: slots n/step+1, strict loop bound, short regions"""
    O = oracle
    assert O.window_slots(100, 100, 10) == 0 and O.window_slots(50, 100, 10) == 0
    assert O.window_slots(300, 200, 10) == 11          # j = 0..90, slot 10 stays 0.0
    cfg = O.make_config(2, 3, revcomp=True)
    seq = "ACGTTGCAAGGCTTAACGGATCGATTTACGCGCGATATCGGCTA" * 8
    m = O.extract(cfg, [seq])
    k, code = m.classes()
    feats = [(i, i) for i in range(m.m)]
    theta = np.linspace(-0.2, 0.2, m.m + 1)
    md = dict(cfg=cfg, class_k=k, class_code=code, features=feats, theta=theta)
    out = O.score_windows([md], [seq[:300], seq[:150]], 200, 10)
    assert len(out) == 11 and out[10] == 0.0 and np.all(out[:10] < 0)
    # equals KmerLrEnsemble.Predict on each window
    w3 = O.extract(cfg, [seq[30:230]], frozen=(k, code))
    assert abs(O.log_pdf(w3, theta)[0] - out[3]) < 1e-15


def test_coordinate_estimator_restatement_properties(oracle):
    """estimate_coordinate has no golden in the reference (dead code, never called): the numpy restatement is
    pinned by what the algorithm must do.  (1) With L1Reg = 0 and a tolerance the sweeps stop at, repeated IRLS
    steps reach the maximum-likelihood point: the gradient under the computed class weights vanishes.
    (2) A fixed point of the sweeps satisfies the KKT conditions of the weighted lasso problem of that IRLS
    step.  (3) The Go loop structure: sweeps and outer steps share ONE iteration counter."""
    O = oracle
    rng = np.random.default_rng(3)
    n, m = 500, 10
    D = ((rng.random((n, m)) < 0.3) * rng.integers(1, 4, size=(n, m))).astype(float)
    y = (D[:, 0] * 0.8 - D[:, 1] * 0.6 + 0.4 * rng.normal(size=n)) > 0
    mat = O.from_dense(D)
    cw = O.class_weights(y)
    th, sweeps, delta = O.coordinate(mat, y, np.zeros(m + 1), l1reg=0.0, epsilon=1e-7, max_iter=5000)
    g = O.gradient(mat, y, th, cw)
    assert np.max(np.abs(g)) < 1e-6 and delta <= 1e-7 and sweeps < 5000
    # (2) one IRLS step, many sweeps: KKT of  1/2 sum w (z - x theta)^2 + l1 |theta_{1:}|
    l1 = 3.0
    th0 = np.zeros(m + 1)
    th1, sweeps1, _ = O.coordinate(mat, y, th0, l1reg=l1, epsilon=0.0, max_iter=400)
    assert sweeps1 == 400                # the inner loop uses up the shared counter: ONE IRLS step, 400 sweeps
    X = np.hstack([np.ones((n, 1)), D])
    r = X @ th0; p = 1 / (1 + np.exp(-r)); w = p * (1 - p); z = r + (y - p) / w; w = w * np.where(y, cw[1], cw[0])
    corr = X.T @ (w * (z - X @ th1))
    assert abs(corr[0]) < 1e-6
    for j in range(1, m + 1):
        if th1[j] != 0.0:
            assert abs(corr[j] - l1 * np.sign(th1[j])) < 1e-6
        else:
            assert abs(corr[j]) <= l1 + 1e-9
