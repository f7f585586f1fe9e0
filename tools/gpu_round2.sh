#!/bin/bash
# GPU box (1 or more GPUs): the whole -m gpu suite, smoke, then bench lines.  usage: bash tools/gpu_round2.sh [ngpu]
NG=${1:-1}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1
echo "== pytest -m gpu exit $? =="; tail -n 15 gpurun_out/t_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $? =="; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "== bench reference exit $? =="; cut -c1-400 gpurun_out/bench_ref.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "== bench c2 exit $? =="; tail -3 gpurun_out/bench_c2.err
for w in c3 c4; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 --skip-cpu > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "== bench $w exit $? =="; tail -3 gpurun_out/bench_$w.err
done
if [ "$NG" -gt 1 ]; then
  for w in c2 c3 c4; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --workload $w --steps 20 --warmup 5 \
      > gpurun_out/bench_${w}_${NG}gpu.json 2> gpurun_out/bench_${w}_${NG}gpu.err; echo "== bench $w x$NG exit $? =="; tail -3 gpurun_out/bench_${w}_${NG}gpu.err
  done
fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_c*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d["e2e"]
        print(f.split("/")[-1], "value", d["value"], "sustained", d.get("value_sustained"), "e2e", round(e["value"]), "ceil GB/s", e.get("h2d_ceiling_gbs"), "frac", e.get("frac_of_h2d_ceiling"),
              "gather_ms", e.get("gather_ms"), "ms/step", d["ms_per_step"], "clk", d["clocks"]["sm_mhz"], d.get("sustained", {}).get("clocks", {}).get("sm_mhz"))
        print("   ", " ".join("%s=%.3f(%.2f)" % (k["name"], k["ms"], k["frac"]) for k in d["kernels"]))
    except Exception as ex:
        print(f, "ERR", ex)
PY
