#!/bin/bash
# build_variant.sh DIR OUT.so : compile DIR/neurecon_b200/csrc/*.cu (headers in DIR/include) into OUT.so
set -e
D=$1; OUT=$2
objs=""
for src in $D/neurecon_b200/csrc/*.cu; do
  o=$D/$(basename ${src%.cu}).o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 \
      --expt-relaxed-constexpr $NVCC_EXTRA -c $src -o $o 2>/dev/null &
  objs="$objs $o"
done
wait
/usr/local/cuda/bin/nvcc -shared -o $OUT $objs -lcudart -lcuda 2>/dev/null
echo $OUT
