#!/bin/bash
# rollout parity tests, the 18-configuration validation against the C oracle, then the headline bench (device value, e2e)
python -m pytest tests/test_gpu_parity.py tests/test_variant_gpu.py tests/test_facade_gpu.py -m gpu -x -q 2>&1 | tail -1
timeout 300 python tools/validate_rollout.py 65536 2>&1 | tail -2
for g in ${GAMES:-65536 262144}; do python bench.py --no-extras --games $g --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rollout', d['config']['games_per_gpu'], d['value'], 'e2e', d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms_min'])"; done
