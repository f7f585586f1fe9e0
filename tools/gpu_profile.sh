#!/bin/bash
# ncu evidence (B200_PROFILING.md recipe): launch list, then one fully profiled forward.  1 GPU only.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-dwconv|gemm_tc|stem|decode|candidates}" -s ${KSKIP:-0} -c ${KCOUNT:-31} -f -o gpurun_out/prof_full $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out
