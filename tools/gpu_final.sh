#!/bin/bash
# round-end evidence in one call: the whole GPU suite + every bench line (gpu_full.sh), then the ncu launch list of the same build
bash tools/gpu_full.sh
CMD="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
