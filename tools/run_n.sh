#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
MODE=${2:-random}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 --mode $MODE > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
tail -c 800 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'n_gpus', d['n_gpus'], {k:d.get(k) for k in ('rollout_ms','update_ms','cuda_graph','last_batch','games_per_sec')})
for k,v in d.get('extra',{}).items():
    if k=='config3':
        for p,m in v.items(): print('config3',p,m['value'],m['roofline']['frac'])
    else: print(k,v['value'],(v.get('e2e') or {}).get('value'),(v.get('roofline') or {}).get('frac'),v.get('rollout_ms'),v.get('update_ms'),v.get('allreduce_and_stats_ms'))
PY
