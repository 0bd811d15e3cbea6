"""Reduced-matrix iteration time against the row length (run on a B200): C2 rows, columns chosen so that the rows of the
reduced matrix are short (classes spread over all levels) or long (the classes of the short k-mers, present in every
row), with the solver on the compact rows (small_long = 0) and on the sliced + column-major views (1).  The iterates
of the two must be the same bits.   python tools/reduced_probe.py [iterations]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import kmerlr_b200 as K
from kmerlr_b200 import synth

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
if len(sys.argv) > 2:
    K_BPS = int(sys.argv[2])
else:
    K_BPS = 0
K.init(0)
K.option("persist_bps", K_BPS)
buf, off, y = synth.training_set(100000, 100000, 500)
kc = K.NewKmerCounter(1, 8, revcomp=True)
d = K.compile_test_data(None, kc, None, None, True, False, (buf, off))
d.SetLabels(y)
cases = {
    "spread 100": np.linspace(1, d.m, 100).astype(np.int64),
    "first 10 + spread 90": np.concatenate([np.arange(1, 11), np.linspace(50, d.m, 90).astype(np.int64)]),
    "first 20 + spread 80": np.concatenate([np.arange(1, 21), np.linspace(50, d.m, 80).astype(np.int64)]),
    "first 47 + spread 53": np.concatenate([np.arange(1, 48), np.linspace(50, d.m, 53).astype(np.int64)]),
    "first 100": np.arange(1, 101),
    "first 300": np.arange(1, 301),
}
for name, cols in cases.items():
    sel = np.unique(np.concatenate([[0], cols]))
    rd = K.select_data(d, sel)
    rd.SetLabels(y)
    out = {}
    for mode in (0, 1, 2):
        K.option("small_long", mode)
        est = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=0.0, MaxIterations=50)
        est.Theta = np.zeros(len(sel)); est.ClassWeights = np.array([1.0, 1.0])
        est.estimate_proximal(rd, 1e-6)
        est.MaxIterations = iters
        est.Theta = np.zeros(len(sel))
        n_it, _ = est.estimate_proximal(rd, 1e-6)
        out[mode] = (K.last_device_ms() / max(n_it, 1), est.Theta.copy())
    same = np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][1], out[2][1])
    print("%-22s nnz %9d (%5.1f per row): rows %7.2f us  sliced+columns %7.2f us  block-local %7.2f us  same bits: %s" % (
        name, rd.nnz, rd.nnz / rd.n, 1e3 * out[0][0], 1e3 * out[1][0], 1e3 * out[2][0], same), flush=True)
    rd.free()
K.option("small_long", -1)
