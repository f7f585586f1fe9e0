#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -p no:cacheprovider -k stem 2>&1 | tail -4
for w in c2 c3 c4; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --skip-cpu > gpurun_out/b5_$w.json 2> gpurun_out/b5_$w.err; echo "== bench $w exit $?"; tail -2 gpurun_out/b5_$w.err
done
PN_STEM_WHOLE_ROWS=1 timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 --skip-cpu --skip-e2e > gpurun_out/b5_c3_wholerows.json 2> gpurun_out/b5_c3_wr.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/b5_c*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value", d["value"], "sustained", d.get("value_sustained"), "e2e", d.get("e2e", {}).get("value"), "ms/step", d["ms_per_step"])
        print("   ", " ".join("%s=%.3f(%.2f)" % (k["name"], k["ms"], k["frac"]) for k in d["kernels"]))
    except Exception as ex:
        print(f, "ERR", ex)
PY
