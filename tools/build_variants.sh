#!/bin/bash
# builds A/B variants of the cell kernel: tools/build_variants.sh "name:flags" ...
set -e
cd "$(dirname "$0")/.."
OBJ=dftfe_b200/build
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
mkdir -p dftfe_b200/lib/variants
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  nvcc $FLAGS $defs -Xptxas=-v -c ${SRC:-dftfe_b200/csrc/cell_matvec.cu} -o /tmp/cell_$name.o 2>&1 | grep -A1 "persistent_kernelILi343ELb0" | grep -E "registers|spill" | head -3
  others=$(ls $OBJ/*.o | grep -v cell_matvec)
  nvcc -shared -o dftfe_b200/lib/variants/lib_$name.so /tmp/cell_$name.o $others -L/usr/local/cuda/lib64 -lcublas -lcusolver -ldl -Xlinker -rpath=/usr/local/cuda/lib64 2>/dev/null
  echo built $name
done
