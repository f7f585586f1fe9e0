#!/bin/bash
# septc iteration: correctness, then per-kernel device times of the bench plan with and without the tensor-pipe depthwise
mkdir -p gpurun_out
timeout 200 python tools/check_septc.py 2>&1 | cut -c1-110 | tail -14
for tc in 1 0; do
  PN_SEP_TC=$tc timeout 300 python bench.py --skip-cpu --skip-e2e --steps 5 --warmup 3 ${BENCH_ARGS} > gpurun_out/tc_$tc.json 2> gpurun_out/tc_$tc.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/tc_$tc.json").read().strip().splitlines()[-1])
    print("PN_SEP_TC=$tc", d["value"], "img/s", " ".join("%s=%.3f" % (k["name"], k["ms"]) for k in d["kernels"]))
except Exception as e:
    print("no line", e); print(open("gpurun_out/tc_$tc.err").read()[-1500:])
PY
done
