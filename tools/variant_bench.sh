#!/bin/bash
# A/B timing of kernel variants: every kmerlr_b200/libv_*.so replaces the in-tree library for one short bench run.
cp kmerlr_b200/libkmerlr_b200.so /tmp/lib_orig.so
for f in kmerlr_b200/libv_*.so; do
  cp "$f" kmerlr_b200/libkmerlr_b200.so
  printf "%s " "$f"
  timeout 300 python bench.py --no-cpu-baseline --score-mbp 0 --iters 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step', d['ms_per_step'], 'extract', d['kernels_ms_per_step']['extract_kernel'], 'e2e', d['e2e']['ms_per_step'], 'reduced us/iter', 1e3*d['reduced_proxgrad'].get('ms_per_iter', 0))"
done
cp /tmp/lib_orig.so kmerlr_b200/libkmerlr_b200.so
