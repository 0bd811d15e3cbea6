"""A few extractions of a C2 / C3 shaped set (resident sequences) -- the command ncu wraps to capture the
extraction kernel.  python tools/prof_extract.py [c2|c3] [reps] [rows_divisor]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kmerlr_b200 as K
from kmerlr_b200 import api, synth

CONFIGS = {"c2": (100000, 100000, 500, 1, 8, False), "c3": (1000000, 1000000, 200, 1, 10, True)}
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
div = int(sys.argv[3]) if len(sys.argv) > 3 else 1
nf, nb, L, M, N, binz = CONFIGS[name]
nf //= div; nb //= div
K.init(0)
buf, off, y = synth.training_set(nf, nb, L)
seqs = K.Sequences((buf, off))
kc = K.NewKmerCounter(M, N, revcomp=True, binarize=binz)
K.api.profile(True)
for i in range(reps):
    d = api._extract(kc, seqs, None, None, False)
    print("%s extract %d: %.3f ms (n=%d m=%d nnz=%d)" % (name, i, K.last_device_ms(), d.n, d.m, d.nnz), flush=True)
    d.free()
for k, v in sorted(K.api.profile_dump().items(), key=lambda kv: -kv[1][0])[:6]:
    print("  %-60s %9.3f ms / %d launches" % (k[:60], v[0], v[1]))
