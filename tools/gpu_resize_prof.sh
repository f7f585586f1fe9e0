#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -p no:cacheprovider -k "resize or preprocess" 2>&1 | tail -2
echo "== words"; python tools/bench_resize.py 2>&1 | tail -2
echo "== bytes"; POSENET_B200_LIB=$PWD/posenet-pytorch_b200/lib/libposenet_b200_rb.so python tools/bench_resize.py 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:"resize_linear" -s 3 -c 1 -f -o gpurun_out/prof_resize python tools/bench_resize.py > gpurun_out/ncu_resize.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_resize.ncu-rep --page details 2>/dev/null | grep -E "Duration|Throughput|Issue|Active Warps|Registers|Theoretical Occ|Achieved Occ|Hit Rate|Stall|Mem Busy|Max Bandwidth|Mem Pipes|Executed Ipc|No Eligible|Bank|Eligible" | head -50
