"""world_size-2 gloo test of the N > 1 host logic (CPU): id-range sharding + counter reduction.

Each rank rolls out its shard of the global game-id range (with the host build of the rules header
standing in for the GPU), the counters are summed with the product's ``parallel.reduce_counters``
and both ranks must see exactly what a single process computes for the whole range."""
import os
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world_size, port, n_per_rank, k, seed, out_dir):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size),
                      LOCAL_RANK=str(rank))
    from azul_deep_reinforcement_learning_b200 import parallel
    from oracle import oracle as O
    from tests import harness as H
    parallel.init("gloo")
    r, w, _ = parallel.world()
    assert (r, w) == (rank, world_size)
    base = parallel.shard(r, n_per_rank)
    recs = O.fresh_records(n_per_rank, 2, 1, 0, seed, base)
    cnt = torch.from_numpy(H.rollout(recs, 2, 1, 0, seed, base, k))
    parallel.reduce_counters(cnt)
    slow = parallel.max_over_ranks(1.0 + rank)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), recs=recs, cnt=cnt.numpy(), slow=slow)
    torch.distributed.destroy_process_group()


def test_two_rank_sharding_and_reduction(tmp_path):
    from oracle import oracle as O
    from tests import harness as H
    H.lib(); O.lib()                      # build once before forking
    from tests.helpers import free_port
    n, k, seed, port = 96, 90, 31337, free_port()
    mp.spawn(_worker, args=(2, port, n, k, seed, str(tmp_path)), nprocs=2, join=True)
    whole = O.fresh_records(2 * n, 2, 1, 0, seed, 0)
    cnt = O.rollout_random(whole, 2, 1, 0, seed, 0, k)
    r0 = np.load(tmp_path / "rank0.npz"); r1 = np.load(tmp_path / "rank1.npz")
    assert np.array_equal(np.concatenate([r0["recs"], r1["recs"]]), whole)
    assert np.array_equal(r0["cnt"], cnt) and np.array_equal(r1["cnt"], cnt)
    assert r0["slow"] == 2.0 and r1["slow"] == 2.0
