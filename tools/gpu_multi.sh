#!/bin/bash
# N-GPU box: NCCL sharding tests, then bench.py at N GPUs for configs[1..3].  usage: bash tools/gpu_multi.sh N [tests]
NG=${1:-8}
mkdir -p gpurun_out
if [ "$2" = "tests" ]; then
  timeout 900 python -m pytest tests/test_gpu_sharding.py tests/test_gpu_ref_scripts.py -q -m gpu -p no:cacheprovider > gpurun_out/t_multi_${NG}.log 2>&1
  echo "== sharding + ref scripts tests exit $? =="; tail -n 6 gpurun_out/t_multi_${NG}.log
fi
for w in c2 c3 c4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --workload $w --steps 20 --warmup 5 --sustain 2 \
    > gpurun_out/bench_${w}_${NG}gpu.json 2> gpurun_out/bench_${w}_${NG}gpu.err; echo "== bench $w x$NG exit $? =="; tail -2 gpurun_out/bench_${w}_${NG}gpu.err | cut -c1-300
done
python - <<PY
import json
for w in ("c2","c3","c4"):
    try:
        d = json.loads(open("gpurun_out/bench_%s_${NG}gpu.json" % w).read().strip().splitlines()[-1])
        e = d["e2e"]
        print(w, "x${NG}", "value", d["value"], "sustained", d.get("value_sustained"), "e2e", round(e["value"]), "ceil GB/s", e.get("h2d_ceiling_gbs"), "frac", e.get("frac_of_h2d_ceiling"),
              "gather_ms", e.get("gather_ms"), "ms/step", d["ms_per_step"])
    except Exception as ex:
        print(w, "ERR", ex)
PY
