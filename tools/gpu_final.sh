#!/bin/bash
# round-end evidence on one B200: full -m gpu suite, smoke, the reference arm, bench lines of configs[1..3], ncu launch list + full capture
bash tools/gpu_round2.sh 1
KREGEX="sep|dw|gemm|stem|decode|candidates" KCOUNT=19 bash tools/gpu_profile.sh 2>&1 | tail -5
