#!/bin/bash
# A/B timing of logistic-pass variants: every kmerlr_b200/libv_*.so replaces the in-tree library for one run of
# tools/prof_iter.py (C2 and a quarter of C3).
cp kmerlr_b200/libkmerlr_b200.so /tmp/lib_orig.so
for f in /tmp/lib_orig.so kmerlr_b200/libv_*.so; do
  [ "$f" != /tmp/lib_orig.so ] && cp "$f" kmerlr_b200/libkmerlr_b200.so
  for cfg in "c2 5 1" "c3 5 4"; do
    printf "%s %s: " "$f" "$cfg"
    timeout 300 python tools/prof_iter.py $cfg 2>&1 | grep -E "per iteration|imp_pass|low_acc" | tr -s ' ' | tr '\n' ' '
    echo
  done
done
cp /tmp/lib_orig.so kmerlr_b200/libkmerlr_b200.so
