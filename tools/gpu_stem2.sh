#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -p no:cacheprovider -k "stem" 2>&1 | tail -4
echo "== time_stem"; timeout 300 python tools/time_stem.py 2>&1 | tail -16
CMD="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --sustain 0"
$CMD > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sep|dw|gemm|stem|decode|candidates" -s 0 -c 19 -f -o gpurun_out/prof_full $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out | tail -5
