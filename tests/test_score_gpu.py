"""Stage 3 parity: sliding-window genomic scoring (through the C ABI) vs the CPU oracle, which
restates genomicKmerLr.Predict per window (kmerLr_predict_genomic.go:134-171).  Needs a B200."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_model(K, O, seqs, M, N, n_feat, members=1, pairs=0, summary="", seed=0, **flags):
    """a model over classes observed in seqs: n_feat singles (+ pairs), random theta"""
    rng = np.random.default_rng(seed)
    ref = O.extract(O.make_config(M, N, **flags), seqs)
    k, code = ref.classes()
    pick = np.sort(rng.choice(ref.m, size=min(n_feat, ref.m), replace=False))
    ck, cc = k[pick], code[pick]
    feats = [(i, i) for i in range(len(pick))]
    for _ in range(pairs):
        a, b = sorted(rng.choice(len(pick), 2, replace=False))
        feats.append((int(a), int(b)))
    theta = rng.normal(scale=0.3, size=(members, len(feats) + 1))
    kmd = dict(counter=K.NewKmerCounter(M, N, **flags), class_k=ck, class_code=cc, features=feats, theta=theta, summary=summary)
    omd = dict(cfg=O.make_config(M, N, **flags), class_k=ck, class_code=cc, features=feats, theta=theta, summary=summary)
    return kmd, omd


def regions(seed=3):
    from kmerlr_b200 import synth
    rng = np.random.default_rng(seed)
    lens = [1000, 200, 201, 150, 777, 4096, 0, 210, 5000]
    seqs = []
    for i, L in enumerate(lens):
        s = bytes(synth.random_bases(L, 3, offset=100000 * i)).decode()
        seqs.append(s)
    s = list(seqs[4])
    for p in rng.choice(777, 12, replace=False):
        s[p] = "N"
    seqs[4] = "".join(s).lower()
    return seqs


def check(K, O, kms, oms, seqs, W, step, tol=1e-12):
    out = K.genomicKmerLr(kms).predict_window_genomic(seqs, W, step)
    ref = O.score_windows(oms, seqs, W, step)
    flat = np.concatenate(out) if len(out) else np.zeros(0)
    assert len(flat) == len(ref)
    assert np.allclose(flat, ref, rtol=tol, atol=tol)
    # slots the strict loop bound never writes stay exactly 0.0 (kmerLr_predict_genomic.go:153-159)
    assert np.array_equal(flat == 0.0, ref == 0.0)
    return flat


def test_linear_models(K, oracle):
    """count features, singles only: the prefix-sum kernel"""
    seqs = regions()
    km, om = make_model(K, oracle, seqs[:1], 1, 8, 100, seed=1, revcomp=True)
    f = check(K, oracle, [km], [om], seqs, 200, 10)
    assert len(f) == sum(oracle.window_slots(len(s), 200, 10) for s in seqs) and np.all(f <= 0)
    km2, om2 = make_model(K, oracle, seqs[:1], 3, 6, 40, seed=2)
    check(K, oracle, [km2], [om2], seqs, 100, 1)
    check(K, oracle, [km, km2], [om, om2], seqs, 200, 7)       # two models are summed
    km3, om3 = make_model(K, oracle, seqs[:1], 2, 5, 30, seed=4, complement=True)
    check(K, oracle, [km3], [om3], seqs, 64, 3)


def test_generic_models(K, oracle):
    """binarized counts, pair features, ensembles with a summary: one warp per window"""
    seqs = regions()[:5]
    km, om = make_model(K, oracle, seqs[:1], 2, 6, 60, pairs=20, seed=5, revcomp=True, binarize=True)
    check(K, oracle, [km], [om], seqs, 200, 10)
    km, om = make_model(K, oracle, seqs[:1], 1, 7, 50, pairs=10, seed=6, revcomp=True)
    check(K, oracle, [km], [om], seqs, 120, 5)
    for summary in ("mean", "product", "min", "max"):
        km, om = make_model(K, oracle, seqs[:1], 2, 5, 30, members=3, summary=summary, seed=7, revcomp=True)
        check(K, oracle, [km], [om], seqs, 150, 25)
    km2, om2 = make_model(K, oracle, seqs[:1], 4, 4, 20, seed=8)
    check(K, oracle, [km, km2], [om, om2], seqs, 150, 25)


def test_predict_window_layout(K, oracle):
    """predict --sliding-window (kmerLr_predict.go:89-124): len - W slots per sequence, the window starting at j
    in slot j, the slots in between stay 0.0; same scores as the genomic layout"""
    seqs = regions()
    km, om = make_model(K, oracle, seqs[:1], 1, 8, 60, seed=4, revcomp=True)
    g = K.genomicKmerLr([km])
    W, step = 200, 7
    pw = g.predict_window(seqs, W, step)
    gen = g.predict_window_genomic(seqs, W, step)
    for s, a, b in zip(seqs, pw, gen):
        n = len(s) - W
        assert len(a) == max(n, 0)
        if n <= 0:
            continue
        starts = np.arange(0, n, step)
        assert np.array_equal(a[starts], b[:len(starts)])
        rest = np.ones(n, dtype=bool)
        rest[starts] = False
        assert np.all(a[rest] == 0.0)


def test_ensemble_without_summary_fails(K, oracle):
    seqs = regions()[:1]
    km, _ = make_model(K, oracle, seqs, 2, 4, 10, members=2, summary="")
    with pytest.raises(K.KmerLrError):
        K.genomicKmerLr([km]).predict_window_genomic(seqs, 100, 10)


def test_host_buffers_arrive_in_chunks(K, oracle):
    """kmerlr_score_windows with host buffers above 16 MB works chunk by chunk (copies in both directions overlap the
    scoring): ragged regions -- empty, shorter than the window, one as long as several chunks -- and two summed
    models; the result equals the resident path bit for bit and the oracle on the regions it finishes in seconds"""
    from kmerlr_b200 import synth
    lens = [300_000, 0, 150, 5_000_000, 17, 11_000_000, 201, 2_500_000, 64, 9_000_000, 123_457, 3_000_000, 1_000]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    buf = synth.random_bases(int(off[-1]), 7)
    buf[1000:1040] = ord("N")
    small = [bytes(buf[off[i]:off[i + 1]]).decode() for i in (0, 2)]
    km, om = make_model(K, oracle, small[:1], 1, 8, 100, seed=1, revcomp=True)
    km2, om2 = make_model(K, oracle, small[:1], 2, 6, 30, seed=2)
    g = K.genomicKmerLr([km, km2])
    W, step = 200, 10
    out = g.predict_window_genomic((buf, off), W, step)
    slots = [oracle.window_slots(int(L), W, step) for L in lens]
    assert [len(o) for o in out] == slots
    seqs = K.Sequences((buf, off))
    res = g.predict_resident(seqs, W, step, fetch=True, total_slots=sum(slots))
    seqs.free()
    assert np.array_equal(np.concatenate(out), res[:sum(slots)])
    for i in (0, 2, 4, 6, 8, 10, 12):
        s = bytes(buf[off[i]:off[i + 1]]).decode()
        ref = oracle.score_windows([om, om2], [s], W, step)
        assert np.allclose(out[i], ref, rtol=1e-12, atol=1e-12), i
