#!/bin/bash
# tests of azb_step + the step bench for 2 / 3 / 4 players (one line each: players, median ms, fraction of the HBM peak, min ms)
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k step 2>&1 | tail -1
for p in ${PLAYERS:-2 3 4}; do python bench.py --mode step --players $p --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['players'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_min'])"; done
