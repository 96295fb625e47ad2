#!/bin/bash
# usage: python tools/build_variant.py NAME -D...; gpurun -- bash tools/sweep_rollout.sh NAME...   (headline bench per variant library)
for v in "$@"; do
  AZB_LIB=gpurun_variants/$v.so python bench.py --no-extras --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_min'])"
done
