"""Wall-clock split of one e2e step (host buffers -> extract -> labels -> one prox-grad iteration)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kmerlr_b200 as K
from kmerlr_b200 import api, synth

K.init(0)
if len(sys.argv) > 1:
    K.option("feed_growth", int(sys.argv[1]))
n_fg = n_bg = 100000
buf, off, labels = synth.training_set(n_fg, n_bg, 500)
pin = torch.empty(len(buf), dtype=torch.uint8).pin_memory(); pin.numpy()[:] = buf
hbuf = pin.numpy()
counter = K.NewKmerCounter(1, 8, revcomp=True)
cw = np.ones(2)
acc = {}
for rep in range(8):
    t = [time.perf_counter()]
    data = api._extract(counter, (hbuf, off), None, None, False); t.append(time.perf_counter())
    data.SetLabels(labels); t.append(time.perf_counter())
    est = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=1e-300, MaxIterations=1)
    est.Theta = np.zeros(data.m + 1); est.ClassWeights = cw; t.append(time.perf_counter())
    est.estimate_proximal(data, 1e-3); t.append(time.perf_counter())
    c = float(est.Theta[0]); t.append(time.perf_counter())
    data.free(); t.append(time.perf_counter())
    if rep >= 3:
        for name, a, b in zip(["extract", "labels", "est", "proxgrad", "read", "free"], t, t[1:]):
            acc.setdefault(name, []).append((b - a) * 1e3)
for k, v in acc.items():
    print("%-10s %.3f ms" % (k, np.mean(v)))
print("total %.3f" % sum(np.mean(v) for k, v in acc.items() if k != "free"))
