cp kmerlr_b200/libkmerlr_b200.so /tmp/lib_orig.so
for f in /tmp/lib_orig.so kmerlr_b200/libv_*.so; do
  [ "$f" != /tmp/lib_orig.so ] && cp "$f" kmerlr_b200/libkmerlr_b200.so
  echo "== $f"
  timeout 300 python tools/reduced_probe.py 500 2>&1 | tail -6 | cut -c1-150
done
cp /tmp/lib_orig.so kmerlr_b200/libkmerlr_b200.so
