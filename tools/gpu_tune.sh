#!/bin/bash
# A/B the fused-block pipeline depths: one short bench per PN_SEP_STAGES setting ("p,w,a,stg").
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_sepconv.py -x -q -m gpu -p no:cacheprovider > gpurun_out/t_sep.log 2>&1
rc=$?; echo "== test_gpu_sepconv exit $rc =="; tail -n 5 gpurun_out/t_sep.log
[ $rc -ne 0 ] && { tail -40 gpurun_out/t_sep.log; exit $rc; }
for cfg in ${CFGS:-default 3,3,2,1 2,3,3,2 3,3,3,1 2,4,2,2}; do
  if [ "$cfg" = default ]; then unset PN_SEP_STAGES; else export PN_SEP_STAGES=$cfg; fi
  PN_SEP_ALL=1 timeout 300 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/tune_$cfg.json 2> gpurun_out/tune_$cfg.err
  python - "$cfg" <<'PY'
import json,sys
cfg=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/tune_%s.json"%cfg).read().strip().splitlines()[-1])
    print(cfg, d["value"], "img/s fwd ms", d["forward_ms_sum_of_kernels"], " ".join("%s=%.3f"%(k["name"],k["ms"]) for k in d["kernels"]))
except Exception as e: print(cfg, "ERR", e, open("gpurun_out/tune_%s.err"%cfg).read()[-600:])
PY
done
