#!/bin/bash
mkdir -p gpurun_out
export PN_SEP_TC=1
for b in 0 1 2 3 4; do
  if [ "$b" = "0" ]; then unset PN_TCS_BANDS; else export PN_TCS_BANDS=$b; fi
  timeout 300 python bench.py --workload c2 --skip-cpu --skip-e2e --steps 5 --warmup 3 > gpurun_out/tb.json 2> gpurun_out/tb.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/tb.json").read().strip().splitlines()[-1])
    k = {x["name"]: x["ms"] for x in d["kernels"]}
    print("bands $b c2 (all TC)", d["value"], " ".join("%s=%.3f" % (n, k[n]) for n in ("sep3", "sep5", "sep7") if n in k))
except Exception as e:
    print("bands $b: no line", open("gpurun_out/tb.err").read()[-200:].replace("\n", " "))
PY
done
