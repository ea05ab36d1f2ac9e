#!/bin/bash
# Builds variants of the library that differ in the -D flags of ONE kernel file (kernel experiments):
#   tools/variants.sh pair_fwd2 name1 "-DFLAG1 -DFLAG2" name2 "" ...
# -> demovlp_b200/lib/variants/<name>.so ; run with DVLP_LIB=<path>
set -e
cd "$(dirname "$0")/.."
file=$1; shift
make -s >/dev/null
mkdir -p demovlp_b200/lib/variants build/variants
others=$(ls build/*.o | grep -v "build/$file.o")
while [ $# -gt 0 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $flags \
       -c demovlp_b200/csrc/$file.cu -o build/variants/$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o demovlp_b200/lib/variants/$name.so build/variants/$name.o $others \
       -lcudart_static -lpthread -ldl -lrt
  echo "built $name ($flags)"
done
