#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-sepconv}" -s ${KSKIP:-0} -c ${KCOUNT:-13} -f -o gpurun_out/prof_sep $CMD > gpurun_out/ncu_sep.log 2>&1
echo "capture exit $?"; tail -3 gpurun_out/ncu_sep.log
