"""C4-shaped probe (pair features): timing of the co-occurrence gradient and one leapfrog epoch at a
reduced sample count.  Not part of the test suite; run on a B200: python tools/c4_probe.py [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time

import numpy as np

import kmerlr_b200 as K
from kmerlr_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
K.init(0)
buf, off, y = synth.training_set(n // 2, n // 2, 500)
kc = K.NewKmerCounter(1, 6, revcomp=True)
t0 = time.perf_counter()
d = K.compile_test_data(None, kc, None, None, True, False, (buf, off))
d.SetLabels(y)
print("extract: n=%d m=%d nnz=%d  %.1f ms" % (d.n, d.m, d.nnz, 1e3 * (time.perf_counter() - t0)), flush=True)
nt = K.CoeffIndex(d.m).Dim()
theta = np.zeros(nt)
lr = K.logisticRegression(theta, (1.0, 1.0), 0.0, Cooccurrence=True)
for rep in range(3):
    if rep == 2:
        K.api.profile(True)
    t0 = time.perf_counter()
    g = lr.Gradient(None, d)
    print("pair gradient (%d coefficients): wall %.1f ms, device %.1f ms" % (nt, 1e3 * (time.perf_counter() - t0), K.last_device_ms()), flush=True)
for k, v in sorted(K.api.profile_dump().items(), key=lambda kv: -kv[1][0])[:6]:
    print("  %-60s %9.3f ms / %d launches" % (k[:60], v[0], v[1]))
K.api.profile(False)
s = K.featureSelector((1.0, 1.0), True, 20, d.m, tie=K.TIE_INDEX)
t0 = time.perf_counter()
sel, lam, ok = s.Select(d, 0.0, [], [], 0.0)
print("select N=20: lambda=%g, %d selected, wall %.1f ms" % (lam, sel.c, 1e3 * (time.perf_counter() - t0)), flush=True)
