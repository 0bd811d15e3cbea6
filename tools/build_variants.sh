#!/bin/bash
# Build kernel variants into variants/lib<name>.so: tools/build_variants.sh name "-DKL_X_WALK=1 ..." [name flags ...]
set -e
cd "$(dirname "$0")/../kmerlr_b200/csrc"
mkdir -p ../../variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  rm -rf /tmp/vb_$name; mkdir -p /tmp/vb_$name
  for f in extract_e0 extract_e2 extract_e4 extract_e8 extract_e16 extract_e32 extract_e64 extract; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr --fmad=false $flags -c $f.cu -o /tmp/vb_$name/$f.o &
  done
  wait
  objs=""
  for f in abi comm coordinate formats gapped logistic matrix scan score select; do objs="$objs build/$f.o"; done
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libv_$name.so /tmp/vb_$name/*.o $objs -lcudart -ldl
  echo built kmerlr_b200/libv_$name.so
done
