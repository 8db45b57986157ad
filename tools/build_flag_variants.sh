#!/bin/bash
# Tuning builds of librobotick_b200.so with extra -D flags: tools/build_flag_variants.sh name1:"-DX=1 -DY=0" name2:"..."
cd "$(dirname "$0")/.."
SRC=roboken-fmskf-robot-controller_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false --shared -Xcompiler -fPIC -cudart static"
mkdir -p tools/variants
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  ( nvcc $FLAGS $defs -o tools/variants/lib_${name}.so $SRC/*.cu 2>&1 | grep -i error ) &
done; wait; ls -la tools/variants
