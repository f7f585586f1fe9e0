#!/bin/bash
# Runs on the GPU box (via gpurun): the whole GPU suite, smoke, the bench at every workload, the decode stress and the reference arm.
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "== pytest -m gpu exit $? =="; tail -n 6 gpurun_out/t_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $? =="; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "== bench c2 exit $? =="
for wl in c3 c4; do
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "== bench $wl exit $? =="
done
timeout 600 python tools/bench_decode.py > gpurun_out/bench_c5_decode.jsonl 2> gpurun_out/bench_c5.err; echo "== decode stress exit $? =="
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "== reference arm exit $? =="
python - <<'PY'
import json
for wl in ("c2", "c3", "c4"):
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % wl).read().strip().splitlines()[-1])
        print(wl, d["value"], "img/s e2e", round(d["e2e"]["value"], 1), d["clocks"], d["roofline"])
        print("   " + " ".join("%s=%.3f" % (k["name"], k["ms"]) for k in d["kernels"]))
    except Exception as e:
        print(wl, "no line", e)
print(open("gpurun_out/bench_c5_decode.jsonl").read())
print(open("gpurun_out/bench_ref.json").read()[:600])
PY
tail -n 3 gpurun_out/bench_c2.err gpurun_out/bench_c5.err; true
