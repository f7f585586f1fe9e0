#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sepconv.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
for ns in ${SLEEPS:-0 100 200 500 1000}; do
  PN_SEP_EPI_SLEEP=$ns timeout 300 python bench.py --skip-cpu --skip-e2e --steps 10 --warmup 3 ${BENCH_ARGS} > gpurun_out/sl_$ns.json 2> gpurun_out/sl_$ns.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/sl_$ns.json").read().strip().splitlines()[-1])
    print("sleep $ns ns:", d["value"], "img/s", " ".join("%s=%.3f" % (k["name"], k["ms"]) for k in d["kernels"]))
except Exception as e:
    print("no line", e); print(open("gpurun_out/sl_$ns.err").read()[-1500:])
PY
done
