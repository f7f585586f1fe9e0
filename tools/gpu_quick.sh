#!/bin/bash
# quick check on one B200: the tests named by $K (pytest -k), then bench lines of the three workloads (no CPU leg)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_sepconv.py -x -q -m gpu -p no:cacheprovider -k "${K:-stem}" 2>&1 | tail -4
for w in ${W:-c2 c3 c4}; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --skip-cpu ${BENCH_ARGS} > gpurun_out/q_$w.json 2> gpurun_out/q_$w.err; echo "== bench $w exit $?"; tail -2 gpurun_out/q_$w.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/q_c*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value", d["value"], "sustained", d.get("value_sustained"), "e2e", d.get("e2e", {}).get("value"), "ms/step", d["ms_per_step"])
        print("   ", " ".join("%s=%.3f(%.2f)" % (k["name"], k["ms"], k["frac"]) for k in d["kernels"]))
    except Exception as ex:
        print(f, "ERR", ex)
PY
