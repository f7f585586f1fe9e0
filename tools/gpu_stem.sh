#!/bin/bash
# stem A/B on one B200: bit-equality tests, per-variant device times (tools/time_stem.py), resize word/byte loads A/B, step-level bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -p no:cacheprovider -k "stem or resize or preprocess" 2>&1 | tail -6
echo "== time_stem"; timeout 300 python tools/time_stem.py 2>&1 | tail -14
if [ -f posenet-pytorch_b200/lib/libposenet_b200_rb.so ]; then
  echo "== resize words"; timeout 120 python tools/bench_resize.py 2>&1 | tail -2
  echo "== resize bytes"; POSENET_B200_LIB=$PWD/posenet-pytorch_b200/lib/libposenet_b200_rb.so timeout 120 python tools/bench_resize.py 2>&1 | tail -2
fi
for w in c2 c3 c4; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --skip-cpu > gpurun_out/b6_$w.json 2> gpurun_out/b6_$w.err; echo "== bench $w exit $?"; tail -2 gpurun_out/b6_$w.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/b6_c*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value", d["value"], "sustained", d.get("value_sustained"), "e2e", d.get("e2e", {}).get("value"), "ms/step", d["ms_per_step"])
        print("   ", " ".join("%s=%.3f(%.2f)" % (k["name"], k["ms"], k["frac"]) for k in d["kernels"]))
    except Exception as ex:
        print(f, "ERR", ex)
PY
