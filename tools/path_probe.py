"""C2 leapfrog path probe: time to N = 10, 25, 50, 100 features on the full C2 matrix (1 GPU).
python tools/path_probe.py [n_per_class] [max_iter]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time

import numpy as np

import kmerlr_b200 as K
from kmerlr_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
K.init(0)
buf, off, y = synth.training_set(n, n, 500)
kc = K.NewKmerCounter(1, 8, revcomp=True)
t0 = time.perf_counter()
d = K.compile_test_data(None, kc, None, None, True, False, (buf, off))
d.SetLabels(y)
print("extract: n=%d m=%d nnz=%d  %.1f ms" % (d.n, d.m, d.nnz, 1e3 * (time.perf_counter() - t0)), flush=True)
est = K.KmerLrEstimator(EpsilonLoss=1e-8, MaxIterations=max_iter, tie=K.TIE_INDEX)
t_all = time.perf_counter()
for N in (10, 25, 50, 100):
    t0 = time.perf_counter()
    l0 = K.launch_count()
    epochs = est.estimate_loop(d, N)
    dt = time.perf_counter() - t0
    its = sum(p[1] for p in est.path[-epochs:])
    print("N=%3d: %d epochs, %d prox-grad iterations, lambda=%.6g, %d active, %.1f ms (%d launches)" %
          (N, epochs, its, est.path[-1][0], len(est.active_idx), 1e3 * dt, K.launch_count() - l0), flush=True)
print("path to 100 features: %.1f ms" % (1e3 * (time.perf_counter() - t_all)))
