#!/bin/bash
# Tuning builds of the library with other macro settings, made HERE (nvcc cross-compiles) so that a GPU call spends its
# minutes measuring: bash tools/build_variants.sh name1:"-DFOO=1 -DBAR=2" name2:"..."  ->  variants/<name>.so
# On the box: cp variants/<name>.so bonej2_b200/libbonejthickness.so && python tools/stage_bench.py ...
set -eu
cd "$(dirname "$0")/.."
mkdir -p variants
for spec in "$@"; do
    name=${spec%%:*}; flags=${spec#*:}
    ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $flags \
        -o variants/$name.so bonej2_b200/csrc/*.cu bonej2_b200/csrc/*.cpp && echo "built variants/$name.so ($flags)" ) &
done
wait
