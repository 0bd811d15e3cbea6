"""Parity hardening: a multi-epoch leapfrog path on a C2-shaped sample against the oracle (the reduced
matrices there are large enough for the one-launch persistent solver), the near-tie hazard of the selection, and
predict --sliding-window against the oracle directly.  Needs a B200."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tie", ["go118", "index"])
def test_leapfrog_path_on_a_c2_shaped_sample(K, oracle, tie):
    """2 000 sequences x 500 bp, k = 1..8 revcomp (m = 43 860 columns when all classes occur), --lambda-auto=5,10:
    same epochs, selected sets, lambda sequence, iteration counts as the oracle's estimate_loop, theta <= 1e-5"""
    from kmerlr_b200 import synth
    O = oracle
    buf, off, y = synth.training_set(1000, 1000, 500)
    kc, oc = K.NewKmerCounter(1, 8, revcomp=True), O.make_config(1, 8, revcomp=True)
    d = K.compile_test_data(None, kc, None, None, True, False, (buf, off))
    d.SetLabels(y)
    ref = O.extract(oc, (buf, off), threads=0)
    assert (d.n, d.m, d.nnz) == (ref.n, ref.m, ref.nnz)
    kt, ot = (K.TIE_GO118, O.TIE_GO118) if tie == "go118" else (K.TIE_INDEX, O.TIE_INDEX)
    est = K.KmerLrEstimator(EpsilonLoss=1e-8, tie=kt, MaxIterations=200000)
    oest = O.EstimatorState()
    for N in (5, 10):
        est.path = []
        est.estimate_loop(d, N)
        res = O.estimate_loop(ref, y, (1, 1), N, oest, tie=ot, epsilon_loss=1e-8, max_iter=200000)
        assert len(est.path) == res["epochs"] and res["epochs"] >= 1
        assert np.array_equal(est.active_idx, oest.active_idx)
        lams = np.array([p[0] for p in est.path])
        if N == 5:
            assert abs(lams[0] - res["lambdas"][0]) <= 1e-12 * abs(lams[0])      # a function of g(0) only
        assert np.allclose(lams, res["lambdas"], rtol=1e-9, atol=0)
        # iteration counts: the stopping rule compares loss differences with 1e-8; allow the last step to fall either way
        its, oits = np.array([p[1] for p in est.path]), res["iters"]
        assert np.all(np.abs(its - oits) <= 1)
        assert np.max(np.abs(est.Theta - oest.active_theta)) <= 1e-5 * max(1.0, np.max(np.abs(oest.active_theta)))
    assert sum(p[1] for p in est.path) > 100            # really iterated (persistent reduced solver)


def test_near_tie_hazard_permuted_duplicate_columns(K, oracle):
    """columns with the same multiset of terms.  The device sums 64-bit fixed point (exact, any order): such
    columns get IDENTICAL bits, whatever the order of the samples.  The reference (and the oracle) add fp64 terms in
    sample order, where two such columns can differ in the last bit -- a tie there can be split by rounding.  The
    product follows the exact arithmetic; this test pins that and measures how far the serial sums drift apart."""
    rng = np.random.default_rng(31)
    n, groups, copies = 400, 30, 4
    labels = (rng.random(n) < 0.5).astype(np.uint8)
    # a base column per group; its copies see the SAME rows of each label class but the samples are shuffled inside
    # the class, so every copy has the same multiset of (w_i v_i) terms at theta = 0 in a different order
    cols = []
    for g in range(groups):
        base = (rng.random(n) < 0.3) * rng.integers(1, 6, size=n)
        for c in range(copies):
            v = np.zeros(n)
            for lab in (0, 1):
                idx = np.nonzero(labels == lab)[0]
                v[rng.permutation(idx)] = base[idx]
            cols.append(v)
    X = np.stack(cols, axis=1)
    d = K.from_dense(X)
    d.SetLabels(labels)
    ref = oracle.from_dense(X)
    cw = (0.7, 1.6)
    theta = np.zeros(X.shape[1] + 1)
    theta[0] = 0.3                                       # every row of a label class has the same weight
    g = K.logisticRegression(theta, cw).Gradient(None, d)[1:]
    og = oracle.gradient(ref, labels, theta, cw)[1:]
    assert np.max(np.abs(g - og)) <= 1e-12 * np.max(np.abs(og))
    split = 0
    for grp in range(groups):
        vals = g[grp * copies:(grp + 1) * copies]
        assert len(set(vals.tolist())) == 1              # identical bits on the device
        split += len(set(og[grp * copies:(grp + 1) * copies].tolist())) > 1
    # the serial fp64 sums do split some of these ties (which ones depends on the sample order)
    print("groups whose serial fp64 sums differ in the last bits: %d of %d" % (split, groups))
    # under the index tie rule the product therefore selects the lowest indices of a tied group
    s = K.featureSelector(cw, False, 3, d.m, tie=K.TIE_INDEX)
    sel, lam, ok = s.Select(d, 0.3, [], [], 0.0)
    order = np.lexsort((np.arange(d.m), -np.abs(g)))
    assert ok and sel.sel[1:].tolist() == sorted((order[:3] + 1).tolist())


def test_predict_window_against_the_oracle(K, oracle):
    """predict --sliding-window (kmerLr_predict.go:89-124): len - W slots per sequence, the window starting at j lands
    in slot j for j = 0, step, 2 step, ... < len - W, every other slot stays 0.0.  Against the oracle's per-window
    scores (the same windows as predict-genomic, kmerLr_predict_genomic.go:147-171)."""
    import test_score_gpu as T
    seqs = T.regions()
    for (M, N, nf, flags, W, step, pairs) in [(1, 8, 60, dict(revcomp=True), 200, 10, 0), (2, 6, 30, dict(), 64, 1, 0),
                                              (1, 7, 25, dict(revcomp=True, binarize=True), 100, 7, 5)]:
        km, om = T.make_model(K, oracle, seqs[:1], M, N, nf, pairs=pairs, seed=M + N, **flags)
        out = K.genomicKmerLr([km]).predict_window(seqs, W, step)
        ref = oracle.score_windows([om], seqs, W, step)
        p = 0
        for s, o in zip(seqs, out):
            slots = oracle.window_slots(len(s), W, step)
            r = ref[p:p + slots]
            p += slots
            L = len(s)
            assert len(o) == max(L - W, 0)
            exp = np.zeros(max(L - W, 0))
            nwin = len(range(0, L - W, step)) if L > W else 0
            exp[0:L - W:step] = r[:nwin] if nwin else []
            assert np.allclose(o, exp, rtol=1e-12, atol=1e-12)
            assert np.array_equal(o == 0.0, exp == 0.0)
