"""CSR fused pass: how many columns should accumulate in shared memory?  python tools/hot_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kmerlr_b200 as K
from kmerlr_b200 import api, synth

K.init(0)
for name, (nf, nb, L, M, N, binz) in {"c2": (100000, 100000, 500, 1, 8, False), "c3/4": (250000, 250000, 200, 1, 10, True)}.items():
    buf, off, y = synth.training_set(nf, nb, L)
    d = K.compile_test_data(None, K.NewKmerCounter(M, N, revcomp=True, binarize=binz), None, None, True, binz, (buf, off))
    d.SetLabels(y)
    K.option("implicit", 0)
    for hot in (16384, 12288, 8192, 6144, 4096):
        K.option("hot_cols", hot)
        est = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=0.0, MaxIterations=2)
        est.Theta = np.zeros(d.m + 1)
        est.estimate_proximal(d, 1e-3)
        est.MaxIterations = 10
        est.Theta = np.zeros(d.m + 1)
        est.estimate_proximal(d, 1e-3)
        print("%s n=%d m=%d nnz=%d hot_cols=%d: %.3f ms per CSR iteration" % (name, d.n, d.m, d.nnz, hot, K.last_device_ms() / 10), flush=True)
    K.option("implicit", 1); K.option("hot_cols", 6144)
    d.free()
