#!/bin/bash
# usage: python tools/build_variant.py NAME -D...; gpurun -- bash tools/sweep_step.sh NAME...
# step bench (2 players unless PLAYERS is set) for every variant library named on the command line
for v in "$@"; do
  for p in ${PLAYERS:-2}; do
    AZB_LIB=gpurun_variants/$v.so python bench.py --mode step --players $p --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['config']['players'], d['ms_per_step'], round(d['roofline']['frac'],4), d['roofline']['kernel_ms_min'])"
  done
done
