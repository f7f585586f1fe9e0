#!/bin/bash
# correctness of the tuned / resident-weight build on the fused-block tests and the full-size model tests, then device times of
# the fused-block shapes with and without the table
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sepconv.py tests/test_gpu_full_size.py tests/test_gpu_model.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
C2="64,257,257,64,128,2,1 64,129,129,128,128,1,1 64,129,129,128,256,2,1 64,65,65,256,512,2,1 64,33,33,512,512,1,1"
C4="512,129,129,48,96,2,1 512,65,65,96,96,1,1 512,65,65,96,192,2,1 512,33,33,192,192,1,1 512,33,33,192,384,2,1 512,17,17,384,384,1,1"
C3="32,361,641,32,64,2,1 32,181,321,64,64,1,1 32,181,321,64,128,2,1 32,91,161,128,128,1,1 32,91,161,128,256,1,1"
echo "== tuned"; timeout 300 python tools/time_sep.py $C2 $C4 $C3 2>&1 | grep median | cut -c1-150
echo "== untuned"; PN_SEP_TUNED=0 timeout 300 python tools/time_sep.py $C2 $C4 $C3 2>&1 | grep median | cut -c1-150
TUNE_TILES=150 python tools/tune_sep.py 64,65,65,256,512,2,1 2>&1 | tail -12
