#!/bin/bash
# GPU box: fused-block tests first (bounded), then the rest of the suite and a fused / unfused bench pair.
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_sepconv.py -x -q -m gpu -p no:cacheprovider > gpurun_out/t_sep.log 2>&1
rc=$?; echo "== test_gpu_sepconv exit $rc =="; tail -n 40 gpurun_out/t_sep.log
[ $rc -ne 0 ] && exit $rc
for t in ops gemm_tc decode model; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -q -m gpu -p no:cacheprovider > gpurun_out/t_$t.log 2>&1
  echo "== test_gpu_$t exit $? =="; tail -n 4 gpurun_out/t_$t.log
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $? =="; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; echo "== bench fused exit $? =="
tail -c 4500 gpurun_out/bench_fused.json; tail -5 gpurun_out/bench_fused.err
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --unfused > gpurun_out/bench_unfused.json 2> gpurun_out/bench_unfused.err; echo "== bench unfused exit $? =="
python - <<'PY'
import json
for f in ("fused","unfused"):
    try:
        d=json.loads(open("gpurun_out/bench_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], "img/s e2e", round(d["e2e"]["value"]), "fwd ms", d["forward_ms_sum_of_kernels"])
        print("  ", " ".join("%s=%.3f"%(k["name"],k["ms"]) for k in d["kernels"]))
    except Exception as e: print(f, "ERR", e)
PY
