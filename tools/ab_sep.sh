#!/bin/bash
# A/B device timing of the fused-block shapes of C2 / C4 over several builds of the library.
# usage: bash tools/ab_sep.sh "<lib-suffix> ..."   ("" = the product library; e.g. "_base" = lib/libposenet_b200_base.so)
SHAPES="${SHAPES:-64,257,257,64,128,2,1 64,129,129,128,128,1,1 64,129,129,128,256,2,1 64,65,65,256,512,2,1 64,33,33,512,512,1,1 512,17,17,384,384,1,1}"
for sfx in ${1:-"-"}; do
  [ "$sfx" = "-" ] && sfx=""
  echo "== lib$sfx"
  POSENET_B200_LIB=$PWD/posenet-pytorch_b200/lib/libposenet_b200$sfx.so timeout 300 python tools/time_sep.py $SHAPES 2>&1 | tail -8 | cut -c1-250
done
