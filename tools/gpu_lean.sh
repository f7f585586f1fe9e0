#!/bin/bash
# A/B of the PN_SEP_LEAN builds (tools/build_variant.sh _l<bits> "-DPN_SEP_LEAN=<bits>" sepconv.cu) on the fused-block shapes of
# C2 / C4 / C3: correctness of the combined build first, then device times per variant.
mkdir -p gpurun_out
L=$PWD/posenet-pytorch_b200/lib
POSENET_B200_LIB=$L/libposenet_b200_l15.so timeout 600 python -m pytest tests/test_gpu_sepconv.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
C2="64,257,257,64,128,2,1 64,129,129,128,128,1,1 64,129,129,128,256,2,1 64,65,65,256,512,2,1 64,33,33,512,512,1,1"
C4="512,129,129,48,96,2,1 512,65,65,96,96,1,1 512,65,65,96,192,2,1 512,33,33,192,192,1,1 512,33,33,192,384,2,1 512,17,17,384,384,1,1"
C3="32,361,641,32,64,2,1 32,181,321,64,64,1,1 32,181,321,64,128,2,1 32,91,161,128,128,1,1 32,91,161,128,256,1,1"
for v in "" _l1 _l2 _l4 _l8 _l15; do
  echo "== lib$v"
  POSENET_B200_LIB=$L/libposenet_b200$v.so timeout 300 python tools/time_sep.py $C2 $C4 $C3 2>&1 | grep median | cut -c1-150
done
echo "== l15 with the baseline's tiles"
t() { PN_SEP_TILE=$1 POSENET_B200_LIB=$L/libposenet_b200_l15.so timeout 100 python tools/time_sep.py $2 2>&1 | grep median | cut -c1-150; }
t 10,12,1 64,257,257,64,128,2,1
t 22,4,1 64,129,129,128,256,2,1
t 11,11,1 512,129,129,48,96,2,1
t 3,33,1 512,65,65,96,192,2,1
t 7,12,1 32,361,641,32,64,2,1
t 19,4,1 32,181,321,64,128,2,1
echo "== 256 -> 256: tensor-pipe depthwise (default) vs CUDA-core depthwise (PN_SEP_TC=0), base and l15"
for v in "" _l15; do for tc in 1 0; do
  PN_SEP_TC=$tc POSENET_B200_LIB=$L/libposenet_b200$v.so timeout 100 python tools/time_sep.py 32,91,161,256,256,1,2 64,65,65,256,256,1,1 2>&1 | grep median | cut -c1-150
done; done
