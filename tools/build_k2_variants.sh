#!/bin/bash
# Experiment helper (not product): build libjsgpu variants that differ only in the K2 (k_idct_tile) compile-time knobs.
# Usage: tools/build_k2_variants.sh ; then on the GPU box: JSGPU_LIB=jpegsnoop_b200/variants/libjsgpu_<name>.so python bench.py ...
set -e
cd "$(dirname "$0")/../jpegsnoop_b200/csrc"
make -s >/dev/null
mkdir -p build/var ../variants
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
declare -A V
V[w0]=""
V[w1]="-DIDCT_MIN_CTAS=6"
V[w2]="-DIDCT_MIN_CTAS=4"
V[w3]="-DIDCT_MIN_CTAS=4 -DIDCT_PREFETCH=1"
V[w4]="-DIDCT_THREADS=96 -DIDCT_FORCE_WARPS=0 -DIDCT_MIN_CTAS=6"
V[w5]="-DIDCT_FIN_PACKED=0"
V[w6]="-DIDCT_THREADS=160 -DIDCT_FORCE_WARPS=5 -DIDCT_MIN_CTAS=4"
V[w7]="-DIDCT_THREADS=256 -DIDCT_FORCE_WARPS=8 -DIDCT_MIN_CTAS=2"
for n in "${!V[@]}"; do
  ( $NV ${V[$n]} -c jsgpu_idct.cu -o build/var/idct_$n.o 2> build/var/idct_$n.log
    OBJS=$(ls build/*.o | grep -v jsgpu_idct.o)
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libjsgpu_$n.so $OBJS build/var/idct_$n.o
    echo "$n: ${V[$n]} :: $(grep -A1 'k_idct_tileILi2ELi1' build/var/idct_$n.log | grep -E 'Used' | head -1) $(grep -B1 'Used' build/var/idct_$n.log | grep -A1 'k_idct_tileILi2ELi1' | grep spill | head -1)" ) &
done
wait
