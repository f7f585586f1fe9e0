#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_septc.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -2
for cfg in "0,0,0,0,0" "100,32,64,300,200" "200,64,128,500,300" "300,100,200,1000,500" "100,0,64,300,200"; do
  for wl in c3 c2; do
    PN_SEP_TC=1 PN_TCS_SLEEP=$cfg timeout 300 python bench.py --workload $wl --skip-cpu --skip-e2e --steps 5 --warmup 3 > gpurun_out/ts.json 2> gpurun_out/ts.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ts.json").read().strip().splitlines()[-1])
    k = {x["name"]: x["ms"] for x in d["kernels"]}
    print("$cfg $wl", d["value"], " ".join("%s=%.3f" % (n, k[n]) for n in ("sep3", "sep5", "sep7", "sep13") if n in k))
except Exception as e:
    print("no line", e); print(open("gpurun_out/ts.err").read()[-800:])
PY
  done
done
