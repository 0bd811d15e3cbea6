"""On-disk formats either side of the path: wiggle tracks, export_kmers tables, .path and .trace
tables -- the bytes of the reference's fmt verbs.  The reference holds no golden FILES for them (TestScores4 only
round-trips an export), so the restatement in oracle/oracle.py is pinned to the verbs themselves: exact decimal
arithmetic (`decimal`) for %0.15f / %e, the documented layout for the tables."""
import math
import os
import random
from decimal import ROUND_HALF_EVEN, Decimal

import numpy as np
import pytest

from conftest import cat


# ---- CPU: the oracle's formatting against exact decimal arithmetic, the host-side tables through the C ABI ----------
def test_oracle_fixed_and_exp_formatting_is_exact(oracle):
    rnd = random.Random(7)
    xs = [0.0, 1.0, 0.5, 1 / 65536, 3 / 65536, 12345 / 65536, 5e-16, 4.9999999999999994e-16, 1e-300, 0.9999999999999999,
          9.9999999999999991] + [rnd.random() for _ in range(2000)] + [rnd.random() * 10 ** -rnd.randint(0, 20) for _ in range(2000)]
    for x in xs:
        want = format(Decimal(x).quantize(Decimal("1e-15"), rounding=ROUND_HALF_EVEN), "f")
        assert oracle.go_fmt("%0.15f", x) == want, x
    # %e: six digits after the point, exponent of at least two digits
    for x in [0.0, 1.0, -2.5e-7, 2.496875, 4.870316e-02, 1.5e300, 123456789.0] + [rnd.uniform(-1, 1) * 10 ** rnd.randint(-30, 30) for _ in range(500)]:
        got = oracle.go_fmt("%e", x)
        if x == 0:
            assert got == "0.000000e+00"
            continue
        d = Decimal(x)
        e = d.adjusted()
        mant = (d.scaleb(-e)).quantize(Decimal("1e-6"), rounding=ROUND_HALF_EVEN)
        if abs(mant) >= 10:
            mant = (mant / 10).quantize(Decimal("1e-6"), rounding=ROUND_HALF_EVEN); e += 1
        assert got == "%se%s%02d" % (format(mant, "f"), "+" if e >= 0 else "-", abs(e)), x
    assert oracle.go_fmt("%13e", float("nan")) == " " * 10 + "NaN" and oracle.go_fmt("%e", float("-inf")) == "-Inf"
    assert oracle.go_fmt("%12e", float("inf")) == " " * 8 + "+Inf"
    # Go's portable math.Exp stays within one ulp of the correctly rounded value
    for _ in range(5000):
        x = -rnd.random() * 700
        a, b = oracle.go_exp(x), math.exp(x)
        assert a == b or a == math.nextafter(b, 0) or a == math.nextafter(b, 2), x
    assert oracle.go_exp(0.0) == 1.0 and oracle.go_exp(-800.0) == 0.0 and oracle.go_exp(1e-10) == 1.0 + 1e-10


def test_path_and_trace_tables(oracle, tmp_path):
    """KmerRegularizationPath.Export / Trace.Export through the C ABI (host-side entry points: no GPU needed)"""
    import kmerlr_b200 as K
    rnd = random.Random(3)
    for with_est in (False, True):
        P = K.KmerRegularizationPath()
        for i in range(7):
            P.Lambda.append(rnd.uniform(0, 3) * 10 ** rnd.randint(-6, 1))
            P.Norm.append(rnd.uniform(0, 50))
            P.Theta.append([rnd.uniform(-1, 1) * 10 ** rnd.randint(-9, 2) if rnd.random() < 0.8 else 0.0 for _ in range(5)])
            if with_est:
                P.Estimator.append(i % 3)
        P.Lambda[2] = float("nan"); P.Norm[3] = float("inf"); P.Theta[4][1] = float("-inf"); P.Theta[5] = []
        f = str(tmp_path / ("p%d.path" % with_est))
        P.Export(f)
        assert open(f).read() == oracle.path_text(P.Estimator, P.Lambda, P.Norm, P.Theta)
    P = K.KmerRegularizationPath()
    f = str(tmp_path / "empty.path")
    P.Export(f)
    assert open(f).read() == "%13s %13s %s\n" % ("lambda", "norm", "theta")
    # the README's lambdas print the way the reference prints them (README.md:39,64-68)
    assert oracle.go_fmt("%e", 2.496875) == "2.496875e+00" and oracle.go_fmt("%e", 4.870316e-02) == "4.870316e-02"
    for with_lambda, with_loss in [(False, False), (True, False), (True, True), (False, True)]:
        T = K.Trace()
        for i in range(9):
            T.Iteration.append(i * 37)
            T.Nonzero.append(rnd.randint(0, 120))
            T.Change.append(rnd.uniform(0, 1) * 10 ** rnd.randint(-12, 0))
            T.Duration.append(rnd.choice([0, 999999, 1000000, 59999999999, 3 * 86400 * 10 ** 9 + 5 * 3600 * 10 ** 9 + 7 * 60 * 10 ** 9 + 9123456789,
                                          rnd.randint(0, 10 ** 15)]))
            if with_lambda:
                T.Lambda.append(rnd.uniform(0, 1))
            if with_loss:
                T.Loss.append(rnd.uniform(0, 2))
        f = str(tmp_path / "t.trace")
        T.Export(f)
        assert open(f).read() == oracle.trace_text(T.Duration, T.Iteration, T.Change, T.Nonzero, T.Lambda, T.Loss)
    assert oracle.format_duration(3 * 86400 * 10 ** 9 + 5 * 3600 * 10 ** 9 + 7 * 60 * 10 ** 9 + 9123456789) == "03:05:07:09.123"
    with pytest.raises(K.KmerLrError):
        T.Export(str(tmp_path / "no_such_dir" / "t.trace"))


def test_class_names(oracle):
    """KmerClass.String: members joined by '|', smaller index first, palindromes twice (kmerLr_test.go:40-43)"""
    import kmerlr_b200 as K
    rnd = random.Random(11)
    kc = K.NewKmerCounter(4, 8, revcomp=True, alphabet="gapped-nucleotide")
    code = lambda s: sum("acgtn".index(c) * 5 ** (len(s) - 1 - i) for i, c in enumerate(s))
    assert K.class_name(kc, 6, code("gntanc")) == "gntanc|gntanc"
    assert K.class_name(kc, 8, code("tgaatgca")) == "tgaatgca|tgcattca"
    for flags in [dict(), dict(revcomp=True), dict(complement=True), dict(reverse=True), dict(complement=True, reverse=True, revcomp=True)]:
        for alphabet in ("nucleotide", "gapped-nucleotide"):
            kc = K.NewKmerCounter(1, 10, alphabet=alphabet, **flags)
            oc = oracle.make_config(1, 10, alphabet=alphabet, **flags)
            A = 4 if alphabet == "nucleotide" else 5
            for _ in range(60):
                k = rnd.randint(1, 10)
                c = rnd.randrange(A ** k)
                assert K.class_name(kc, k, c) == oracle.class_name(oc, k, c)


# ---- GPU: the wiggle records are formatted on the device ---------------------------------------------------------------
def tie_predictions(oracle):
    """log-probabilities whose exp (Go's algorithm) is an odd multiple of 2^-16: x 10^15 ends in exactly .5"""
    out = []
    for q in range(1, 65536, 2):
        p = math.log(q / 65536)
        for cand in (p, math.nextafter(p, 0), math.nextafter(p, -1000)):
            if oracle.go_exp(cand) * 65536 == q:
                out.append(cand)
                break
        if len(out) >= 64:
            break
    return out


@pytest.mark.gpu
def test_wiggle_records_on_the_device(K, oracle):
    rnd = random.Random(5)
    ties = tie_predictions(oracle)
    assert len(ties) >= 8
    pred = [0.0, -0.0, -1e-17, -1e-9, -0.6444834689451768, -0.744447612033651, -36.0, -40.0, -700.0, -744.0, -746.0, -1e9,
            float("-inf")] + ties
    pred += [-rnd.random() * 10 ** rnd.randint(-8, 2) for _ in range(20000)]
    pred += [math.log(rnd.random()) for _ in range(20000)]
    rec, irr = K.wiggle_records(np.array(pred))
    assert irr == 0 and len(rec) == 18 * len(pred)
    want = "".join(oracle.go_fmt("%0.15f", oracle.go_exp(p)) + "\n" for p in pred)
    assert rec.decode() == want
    # the device's exp is the portable Go algorithm bit for bit: a correctly rounded exp prints other digits now and then
    assert sum(oracle.go_fmt("%0.15f", math.exp(p)) != oracle.go_fmt("%0.15f", oracle.go_exp(p)) for p in pred) > 0
    # values whose record is not 18 bytes are counted, not printed
    rec, irr = K.wiggle_records(np.array([-1.0, 3.0, float("nan"), float("inf"), -2.0, math.log(9.9999999999999995)]))
    assert irr == 3 + (0 if oracle.go_fmt("%0.15f", oracle.go_exp(math.log(9.9999999999999995))).startswith("9.") else 1)
    assert rec[:18].decode() == oracle.go_fmt("%0.15f", oracle.go_exp(-1.0)) + "\n" and rec[18] == 0 and rec[36] == 0
    assert K.wiggle_records(np.zeros(0)) == (b"", 0)


@pytest.mark.gpu
def test_save_wiggle_file(K, oracle, fixtures, tmp_path):
    """predict_window_genomic -> saveWindowPredictionsWiggle on the bundled sequences, host and resident scores"""
    from kmerlr_b200 import _lib
    buf, off, y = cat(fixtures, "kmerLr_test_fg", "kmerLr_test_bg")
    kc = K.NewKmerCounter(1, 6, revcomp=True)
    d = K.compile_test_data(None, kc, None, None, True, False, (buf, off))
    ck, cc = d.Kmers()
    rng = np.random.default_rng(2)
    sel = np.sort(rng.choice(d.m, 40, replace=False))
    model = dict(counter=kc, class_k=ck[sel], class_code=cc[sel], features=[[i, i] for i in range(40)],
                 theta=rng.normal(size=41) * 0.1, summary="")
    g = K.genomicKmerLr([model])
    W, step = 200, 10
    preds = g.predict_window_genomic((buf, off), W, step)
    regions = [("chr%d" % (i + 1), 1000 * i + 17) for i in range(len(preds))]
    f = str(tmp_path / "a.wig")
    K.saveWindowPredictionsWiggle(f, regions, preds, "track one", W, step)
    want = oracle.wiggle_text(regions, preds, "track one", W, step)
    assert open(f).read() == want
    assert want.count("\n1.000000000000000\n") >= len(preds)       # the unused last slot of every region (exp(0))
    # scores that never leave the device
    seqs = K.Sequences((buf, off))
    h = _lib.C.c_uint64(0)
    K.api.check(K.api.lib().kmerlr_score_windows_resident(g._arr, 1, seqs.h, W, step, None, _lib.C.byref(h)))
    f2 = str(tmp_path / "b.wig")
    K.saveWindowPredictionsWiggle(f2, regions, [len(p) for p in preds], "track one", W, step, scores=h.value)
    assert open(f2).read() == want
    K.api.check(K.api.lib().kmerlr_free(h.value))
    # a prediction that is no log-probability still prints what the reference prints
    odd = [np.array([0.5, -1.0, 5.0, float("nan")]), np.array([]), np.array([float("inf"), -3.0])]
    regions = [("a", 0), ("b", 5), ("c", 7)]
    f3 = str(tmp_path / "c.wig")
    K.saveWindowPredictionsWiggle(f3, regions, odd, "t", 201, 7)
    assert open(f3).read() == oracle.wiggle_text(regions, odd, "t", 201, 7)
    # no regions at all: the track line only
    f4 = str(tmp_path / "d.wig")
    K.saveWindowPredictionsWiggle(f4, [], [], "empty", 200, 10)
    assert open(f4).read() == "track type=wiggle_0 name=empty\n"


@pytest.mark.gpu
def test_export_kmers_tables(K, oracle, fixtures, tmp_path):
    """export_kmers on the bundled sequences: counts ("%d"), and after the standardizer ("%e", dense rows)"""
    buf, off, y = cat(fixtures, "kmerLr_test_fg", "kmerLr_test_bg")
    for M, N, flags in [(2, 4, dict(revcomp=True)), (1, 3, dict()), (2, 3, dict(revcomp=True, binarize=True))]:
        kc = K.NewKmerCounter(M, N, **flags)
        oc = oracle.make_config(M, N, **flags)
        d = K.compile_test_data(None, kc, None, None, True, flags.get("binarize", False), (buf, off))
        om = oracle.extract(oc, (buf, off))
        f = str(tmp_path / "e.table")
        K.export_kmers(kc, f, d)
        assert open(f).read() == oracle.export_kmers_text(om.class_names(), om.dense(), False)
        # export under a data transform (kmerLr_export.go:45-56): Fit, Apply, then "%e" of every entry
        t = K.TransformFull()
        t.Fit(d, "standardizer")
        h = K._lib.C.c_uint64(0)
        K.api.check(K.api.lib().kmerlr_matrix_transform(d.h, K.api._p(t.Offset), K.api._p(t.Scale), len(t.Offset), K._lib.C.byref(h)))
        td = K.KmerDataSet(h.value)
        K.export_kmers(kc, f, td, transformed=True)
        offset, scale = oracle.fit_transform(om, "standardizer")
        dense = (om.dense() - offset[1:]) * scale[1:]
        assert open(f).read() == oracle.export_kmers_text(om.class_names(), dense, True)


@pytest.mark.gpu
def test_wiggle_records_at_scale_round_trip(K):
    """size-independent property at the size of a chromosome's track (8 M records, several pipeline chunks): every
    record is well formed and, read back as a number, is exp(prediction) to within the last printed digit"""
    rng = np.random.default_rng(12)
    n = 8_000_003
    pred = np.concatenate([-rng.random(n // 2) * 3.0, np.log(rng.random(n - n // 2))])
    rec, irr = K.wiggle_records(pred)
    assert irr == 0 and len(rec) == 18 * n
    b = np.frombuffer(rec, dtype=np.uint8).reshape(n, 18)
    assert np.all(b[:, 1] == ord(".")) and np.all(b[:, 17] == ord("\n"))
    digits = np.delete(b[:, :17], 1, axis=1).astype(np.int64) - 48
    assert digits.min() >= 0 and digits.max() <= 9
    value = (digits * (10 ** np.arange(15, -1, -1, dtype=np.int64))).sum(axis=1)          # x 10^15, exact
    want = np.exp(pred) * 1e15
    assert np.max(np.abs(value - want)) <= 1.0 + 1e15 * 2.3e-16       # half a unit of the last digit + one ulp of exp near 1
