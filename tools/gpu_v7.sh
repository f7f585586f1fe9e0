#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1
echo "== pytest -m gpu exit $? =="; tail -n 3 gpurun_out/t_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $? =="; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "== bench c2 exit $? =="
for w in c3 c4; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 --skip-cpu > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "== bench $w exit $? =="
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_c?.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1]); e = d["e2e"]
    print(f.split("/")[-1], "value", d["value"], "sustained", d.get("value_sustained"), "e2e", round(e["value"]), "ms/step", d["ms_per_step"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"])
    print("   ", " ".join("%s=%.3f(%.2f)" % (k["name"], k["ms"], k["frac"]) for k in d["kernels"]))
PY
