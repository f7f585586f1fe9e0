#!/bin/bash
# stem on one B200: bit-equality tests, then device times per staging variant (whole rows / segments, ring depth)
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -p no:cacheprovider -k "stem" 2>&1 | tail -4
timeout 300 python tools/time_stem.py 2>&1 | tail -16
