#!/bin/bash
# Runs on the GPU box (via gpurun): every GPU test file in its own process, then smoke and a short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for t in ops gemm_tc decode model; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -q -m gpu -p no:cacheprovider > gpurun_out/t_$t.log 2>&1
  echo "== test_gpu_$t exit $? =="; tail -n ${TAILN:-12} gpurun_out/t_$t.log
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $? =="; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench exit $? =="
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
