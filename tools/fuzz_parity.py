"""Randomised parity sweep (run on a B200): random k ranges, strand operations, lengths, invalid bases, frozen
subsets, binarize, both alphabets -- extraction, gradient and loss against the oracle.
python tools/fuzz_parity.py [n_cases] [seed]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import kmerlr_b200 as K
from oracle import oracle as O



def run(n_cases, seed, verbose=True):
    """returns the number of mismatching cases"""
    rng = np.random.default_rng(seed)
    K.init(0)
    bad = 0
    for case in range(n_cases):
        gapped = rng.random() < 0.15
        if gapped:
            M = int(rng.integers(1, 6)); N = int(rng.integers(M, min(M + 3, 7) + 1))
        else:
            M = int(rng.integers(1, 10)); N = int(rng.integers(M, min(M + int(rng.integers(0, 8)), 14) + 1))
        op = rng.choice(["none", "revcomp", "complement", "reverse"])
        flags = {} if op == "none" else {op: True}
        if not gapped and rng.random() < 0.12:                           # several strand flags at once (sort-based path)
            for extra in rng.choice(["revcomp", "complement", "reverse"], size=2, replace=False):
                flags[str(extra)] = True
            op = "+".join(sorted(flags))
        binz = bool(rng.random() < 0.3)
        if gapped:
            flags["alphabet"] = "gapped-nucleotide"
        maxL = 150 if gapped else int(rng.choice([40, 200, 600, 1500, 2300]))
        nseq = int(rng.integers(1, 40 if not gapped else 8))
        seqs = []
        for _ in range(nseq):
            L = int(rng.integers(0, maxL + 1))
            alpha = list("ACGT") if rng.random() < 0.7 else list("ACGTacgt")
            if rng.random() < 0.2:
                alpha = alpha[:int(rng.integers(1, 3))]                  # low complexity
            s = rng.choice(alpha, size=L)
            if L and rng.random() < 0.3:
                for p in rng.integers(0, L, size=int(rng.integers(1, 4))):
                    s[p] = rng.choice(list("NnX-"))
            seqs.append("".join(s))
        kc, oc = K.NewKmerCounter(M, N, binarize=binz, **flags), O.make_config(M, N, binarize=binz, **flags)
        tag = "case %d: M=%d N=%d op=%s bin=%d gapped=%d nseq=%d maxL=%d" % (case, M, N, op, binz, gapped, nseq, maxL)
        try:
            d = K.compile_test_data(None, kc, None, None, True, binz, seqs)
            ref = O.extract(oc, seqs)
            ok = (d.n, d.m, d.nnz) == (ref.n, ref.m, ref.nnz)
            if ok:
                k, c = d.Kmers(); rk, rc = ref.classes()
                ok = np.array_equal(k, rk) and np.array_equal(c, rc) and all(np.array_equal(x, y) for x, y in zip(d.rows(), ref.rows()))
            if ok and d.m > 0 and d.n > 0:
                y = rng.integers(0, 2, size=d.n).astype(np.uint8)
                d2 = K.compile_test_data(None, kc, None, None, True, binz, seqs)    # fresh: keeps the extraction layout
                d2.SetLabels(y)
                th = rng.normal(scale=0.05, size=d.m + 1)
                lr = K.logisticRegression(th, (0.7, 1.4), 0.0)
                g, og = lr.Gradient(None, d2), O.gradient(ref, y, th, (0.7, 1.4))
                # fixed-point accumulation: absolute precision nnz(column) * 2^-60 * max(cw) * max|v|, which
                # only shows when every weight underflows it (one saturated row)
                vmax = float(np.max(d.rows()[2])) if d.nnz else 1.0
                ok = np.max(np.abs(g - og)) <= 1e-9 * np.max(np.abs(og)) + 1e-15 * d.n * 1.4 * vmax
                lo, olo = lr.Loss(d2), O.loss(ref, y, th, (0.7, 1.4))
                ok = ok and abs(lo - olo) <= 1e-11 * abs(olo)
                if ok and d.m > 3:
                    sub = (k[::2], c[::2])
                    ds = K.compile_test_data(None, kc, sub, None, True, binz, seqs)
                    rs = O.extract(oc, seqs, frozen=sub)
                    ok = (ds.n, ds.m, ds.nnz) == (rs.n, rs.m, rs.nnz) and all(np.array_equal(x, y2) for x, y2 in zip(ds.rows(), rs.rows()))
            if not ok:
                bad += 1
                print("MISMATCH", tag, (d.n, d.m, d.nnz), (ref.n, ref.m, ref.nnz), flush=True)
                os.makedirs("gpurun_out", exist_ok=True)
                with open("gpurun_out/fuzz_case_%d.txt" % case, "w") as f:
                    f.write(repr(dict(M=M, N=N, flags=flags, binz=binz, seqs=seqs)))
        except K.KmerLrError as e:
            # documented limits (k > 8 with rows beyond the register sort, gapped size limits) must fail loudly
            print("refused", tag, "--", str(e)[:90], flush=True)
    return bad


if __name__ == "__main__":
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    bad = run(n_cases, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    print("FUZZ", "PASS" if bad == 0 else "FAIL (%d)" % bad, n_cases, "cases")
    sys.exit(1 if bad else 0)
