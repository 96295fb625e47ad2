#!/bin/bash
# policy / self-play / golden tests, then the policy and training bench sections (one line each)
python -m pytest tests/test_policy_gpu.py tests/test_selfplay_gpu.py tests/test_reference_golden_gpu.py -m gpu -x -q 2>&1 | tail -2
python bench.py --mode policy --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('policy', d['value'], 'e2e', d['e2e']['value'], d['ms_per_step'], d.get('accuracy'))"
python bench.py --mode train --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train', d['games_per_sec'], d['ms_per_step'], d.get('rollout_ms'), d.get('update_ms'))"
