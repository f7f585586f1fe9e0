#!/bin/bash
# GPU suite + ncu evidence for the tensor-pipe depthwise kernel (one launch of each of three bench shapes) + the UMMA cost table
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "== pytest -m gpu exit $? =="; tail -n 5 gpurun_out/t_all.log
timeout 120 python tools/umma_cost.py > gpurun_out/umma_cost.txt 2>&1; echo "== umma cost exit $? =="
PN_SEP_TC=1 timeout 100 python tools/check_septc.py --time > gpurun_out/septc_time.txt 2>&1; echo "== septc time exit $? =="; cat gpurun_out/septc_time.txt
i=0
for skip in 1 24 47; do
  PN_SEP_TC=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:septc --launch-skip $skip -c 1 -f -o gpurun_out/prof_septc_$i python tools/check_septc.py --time > gpurun_out/ncu_septc_$i.log 2>&1
  echo "== ncu septc $i exit $? =="
  i=$((i+1))
done
ls -la gpurun_out/*.ncu-rep
