#!/bin/bash
# scratch: build a library variant with extra -D flags:  scripts/build_variant.sh <suffix> <units|all> <nvcc flags...>
# <units> = comma-separated translation units to recompile with the flags (the rest come from marbler_b200/build/)
set -e
cd "$(dirname "$0")/../marbler_b200"
suffix=$1; shift
units=$1; shift
mkdir -p build/$suffix
for f in csrc/*.cu; do
  b=$(basename $f .cu)
  if [ "$units" = all ] || [[ ",$units," == *",$b,"* ]]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c -o build/$suffix/$b.o $f &
  else
    cp build/$b.o build/$suffix/$b.o
  fi
done
wait
nvcc -shared -o libmarbler_b200_$suffix.so build/$suffix/*.o
echo built libmarbler_b200_$suffix.so
