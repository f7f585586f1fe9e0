#!/bin/bash
# decode-only iteration loop on the GPU box: decode parity tests, phase trace, decode stress, the three bench workloads
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_full_size.py -q -x -p no:cacheprovider 2>&1 | tail -n 5
POSENET_B200_LIB=$PWD/posenet-pytorch_b200/lib/libposenet_b200_trace.so timeout 200 python tools/trace_decode.py 2>&1 | grep "==\|img 0\|rep 1" | cut -c1-300
timeout 600 python tools/bench_decode.py 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['workload'], 'batch', d['batch'], 'us/img', d['gpu_us_per_image'], 'exact', d['checked_bit_exact'])"
for wl in c2 c3 c4; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --skip-cpu > gpurun_out/q_$wl.json 2> gpurun_out/q_$wl.err
  python -c "
import json
d = json.loads(open('gpurun_out/q_$wl.json').read().strip().splitlines()[-1])
print('$wl', d['value'], 'e2e', round(d['e2e']['value']), 'cand+decode ms', [k['ms'] for k in d['kernels'] if k['name'].startswith('cand')])"
done
