"""The command ncu wraps for the reduced-matrix solver: C2 rows, the 47 classes of the short k-mers + 53 spread
columns (50 entries per row), `iters` iterations in one cooperative launch.
python tools/prof_reduced.py [iters] [small_long] [dense columns]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import kmerlr_b200 as K
from kmerlr_b200 import synth

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
mode = int(sys.argv[2]) if len(sys.argv) > 2 else -1
dense = int(sys.argv[3]) if len(sys.argv) > 3 else 47
K.init(0)
buf, off, y = synth.training_set(100000, 100000, 500)
kc = K.NewKmerCounter(1, 8, revcomp=True)
d = K.compile_test_data(None, kc, None, None, True, False, (buf, off))
d.SetLabels(y)
sel = np.unique(np.concatenate([[0], np.arange(1, dense + 1), np.linspace(50, d.m, 100 - dense).astype(np.int64)]))
rd = K.select_data(d, sel)
rd.SetLabels(y)
K.option("small_long", mode)
for rep in range(2):
    est = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=0.0, MaxIterations=iters)
    est.Theta = np.zeros(len(sel)); est.ClassWeights = np.array([1.0, 1.0])
    n_it, _ = est.estimate_proximal(rd, 1e-6)
    print("nnz %d, %d iterations, %.2f us per iteration" % (rd.nnz, n_it, 1e3 * K.last_device_ms() / max(n_it, 1)), flush=True)
