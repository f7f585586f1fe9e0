#!/bin/bash
# quick GPU check: build, selected tests, short bench
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
for t in ${TESTS:-model}; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -x -q -m gpu -p no:cacheprovider > gpurun_out/t_$t.log 2>&1
  echo "== test_gpu_$t exit $? =="; tail -n ${TAILN:-6} gpurun_out/t_$t.log
done
timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS:---skip-cpu} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "== bench exit $? =="
tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print(d["value"], "img/s  e2e", d["e2e"], "fwd ms", d["forward_ms_sum_of_kernels"], d["clocks"])
print("  ", " ".join("%s=%.3f"%(k["name"],k["ms"]) for k in d["kernels"]))
PY
