#!/bin/bash
# Tuning builds of librobotick_b200.so (CTA size x tick-loop unroll) into tools/variants/ (git-ignored).
cd "$(dirname "$0")/.."
SRC=roboken-fmskf-robot-controller_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false --shared -Xcompiler -fPIC -cudart static"
for th in ${THREADS:-64 128 256}; do for un in ${UNROLLS:-1 2 3 4}; do
  ( nvcc $FLAGS -DRK_FAST_THREADS=$th -DRK_FAST_UNROLL=$un -o tools/variants/lib_t${th}_u${un}.so $SRC/*.cu 2>&1 | grep -i error ) &
done; done; wait; ls -la tools/variants
