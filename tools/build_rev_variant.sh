#!/bin/bash
# build_rev_variant.sh NAME "-DFLAGS..." [SRC]: libneurecon_b200.so with csrc/SRC.cu (default mlp_rev) recompiled under extra
# flags -> neurecon_b200/lib/variants/NAME.so (kernel experiments; run with NEURECON_B200_LIB=that path)
set -e
cd "$(dirname "$0")/.."
SRC=${3:-mlp_rev}
mkdir -p neurecon_b200/lib/variants
O=neurecon_b200/lib/variants/$1.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 \
    --expt-relaxed-constexpr -Iinclude $2 -c neurecon_b200/csrc/$SRC.cu -o $O 2>/dev/null
objs=$(ls neurecon_b200/lib/*.o | grep -v "inject\|devtools\|/$SRC.o")
/usr/local/cuda/bin/nvcc -shared -o neurecon_b200/lib/variants/$1.so $objs $O -lcudart 2>/dev/null
rm -f $O
echo neurecon_b200/lib/variants/$1.so
