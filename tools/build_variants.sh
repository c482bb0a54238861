#!/bin/bash
# Experiment helper (not product): build libjsgpu variants that differ in the compile flags of ONE source file.
# usage: tools/build_variants.sh <source.cu> name1="<flags>" name2="<flags>" ...   -> jpegsnoop_b200/variants/libjsgpu_<name>.so
set -e
SRC=$1; shift
cd "$(dirname "$0")/../jpegsnoop_b200/csrc"
make -s >/dev/null
mkdir -p build/var ../variants
BASE=$(basename $SRC .cu)
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
for kv in "$@"; do
  n="${kv%%=*}"; fl="${kv#*=}"
  ( $NV $fl -c $SRC -o build/var/${BASE}_$n.o 2> build/var/${BASE}_$n.log
    OBJS=$(ls build/*.o | grep -v "build/$BASE.o")
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libjsgpu_$n.so $OBJS build/var/${BASE}_$n.o -ldl
    echo "$n: $fl" ) &
done
wait
