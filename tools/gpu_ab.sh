#!/bin/bash
L=$PWD/posenet-pytorch_b200/lib
timeout 300 python -m pytest tests/test_gpu_sepconv.py -x -q -m gpu -p no:cacheprovider -k "sepconv_block" 2>&1 | tail -2
for v in "" _swold; do
  echo "== lib$v"; POSENET_B200_LIB=$L/libposenet_b200$v.so timeout 200 python tools/time_sep.py 64,257,257,32,64,1,1 512,129,129,24,48,1,1 32,361,641,16,32,1,1 2>&1 | grep median | cut -c1-100
done
