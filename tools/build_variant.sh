#!/bin/bash
# Build an experiment variant of the product library: the named sources recompiled with extra flags, the other objects
# taken from posenet-pytorch_b200/build.  usage: tools/build_variant.sh <suffix> "<nvcc flags>" <file.cu> [...]
set -e
sfx=$1; flags=$2; shift 2
P=posenet-pytorch_b200
objs=""
for o in $P/build/*.o; do
  b=$(basename $o .o)
  case "$b" in diag_*) continue;; esac
  skip=0; for f in "$@"; do [ "$b" = "$(basename $f .cu)" ] && skip=1; done
  [ $skip = 0 ] && objs="$objs $o"
done
for f in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c $P/csrc/$f -o /tmp/var_${sfx}_$(basename $f .cu).o &
  objs="$objs /tmp/var_${sfx}_$(basename $f .cu).o"
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $P/lib/libposenet_b200$sfx.so $objs
echo $P/lib/libposenet_b200$sfx.so
