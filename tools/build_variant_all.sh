#!/bin/bash
# build_variant_all.sh NAME "-DFLAGS..." : the whole production library recompiled under extra flags ->
# neurecon_b200/lib/variants/NAME.so (kernel experiments; run with NEURECON_B200_LIB=that path)
set -e
cd "$(dirname "$0")/.."
D=neurecon_b200/lib/variants/$1.d
mkdir -p $D
objs=""
for src in neurecon_b200/csrc/*.cu; do
  o=$D/$(basename ${src%.cu}).o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 \
      --expt-relaxed-constexpr -Iinclude $2 -c $src -o $o 2>/dev/null &
  objs="$objs $o"
done
wait
/usr/local/cuda/bin/nvcc -shared -o neurecon_b200/lib/variants/$1.so $objs -lcudart 2>/dev/null
rm -rf $D
echo neurecon_b200/lib/variants/$1.so
