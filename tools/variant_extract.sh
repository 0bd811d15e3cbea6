#!/bin/bash
# A/B timing of extraction kernel variants: every kmerlr_b200/libv_*.so replaces the in-tree library for one run of
# tools/prof_extract.py (C2 and a quarter of C3).
cp kmerlr_b200/libkmerlr_b200.so /tmp/lib_orig.so
for f in /tmp/lib_orig.so kmerlr_b200/libv_*.so; do
  [ "$f" != /tmp/lib_orig.so ] && cp "$f" kmerlr_b200/libkmerlr_b200.so
  for cfg in "c2 4 1" "c3 3 4"; do
    printf "%s %s: " "$f" "$cfg"
    timeout 300 python tools/prof_extract.py $cfg 2>&1 | grep "extract_kernel" | tr -s ' '
  done
done
cp /tmp/lib_orig.so kmerlr_b200/libkmerlr_b200.so
