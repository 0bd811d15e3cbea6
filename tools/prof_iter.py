"""A few full-space prox-grad iterations on a resident matrix (C2 / C3 shape) -- the command ncu wraps to
capture the logistic pass.  python tools/prof_iter.py [c2|c3] [iters] [rows_divisor] [implicit 0/1] [super_len]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kmerlr_b200 as K
from kmerlr_b200 import synth

CONFIGS = {"c2": (100000, 100000, 500, 1, 8, False), "c3": (1000000, 1000000, 200, 1, 10, True)}
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
div = int(sys.argv[3]) if len(sys.argv) > 3 else 1
implicit = int(sys.argv[4]) if len(sys.argv) > 4 else 1
nf, nb, L, M, N, binz = CONFIGS[name]
nf //= div; nb //= div
K.init(0)
K.option("implicit", implicit)
if len(sys.argv) > 5:
    K.option("super_len", int(sys.argv[5]))
if len(sys.argv) > 6:
    K.option("fused_ticket", int(sys.argv[6]))
buf, off, y = synth.training_set(nf, nb, L)
d = K.compile_test_data(None, K.NewKmerCounter(M, N, revcomp=True, binarize=binz), None, None, True, binz, (buf, off))
print("extract: %.3f ms" % K.last_device_ms(), flush=True)
d.SetLabels(y)
est = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=0.0, MaxIterations=iters)
est.Theta = np.zeros(d.m + 1)
est.estimate_proximal(d, 1e-3)
est.Theta = np.zeros(d.m + 1)
K.api.profile(True)
est.estimate_proximal(d, 1e-3)
ms = K.last_device_ms()
prof = K.api.profile_dump()
print("%s n=%d m=%d nnz=%d implicit=%d: %.3f ms per iteration" % (name, d.n, d.m, d.nnz, implicit, ms / iters), flush=True)
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:8]:
    print("  %-60s %9.3f ms / %d launches" % (k[:60], v[0], v[1]))
