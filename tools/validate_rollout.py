"""Wider one-off check of the fused rollout against the C oracle than the test-suite's: every (players, pool, first-player
rule), 20,000 games x 400 env steps, records and counters bit-exact.  python tools/validate_rollout.py [n_games]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from azul_deep_reinforcement_learning_b200.engine import BatchedAzul  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    n, k = (int(sys.argv[1]) if len(sys.argv) > 1 else 20000), 400     # 65536: one block of 14 game warps + 2 helper warps per SM
    t0 = time.time()
    for players in (2, 3, 4):
        for pool in (0, 1):
            for first in (0, 1, players):
                seed, base = 1000 * players + 10 * pool + first, 12345 * players
                eng = BatchedAzul(n, players, pool, first, seed=seed, game_id_base=base)
                ref = O.fresh_records(n, players, pool, first, seed, base)
                assert np.array_equal(eng.export_records().cpu().numpy(), ref)
                eng.rollout_random(k // 2)
                eng.rollout_random(k - k // 2)
                cnt = O.rollout_random(ref, players, pool, first, seed, base, k, threads=os.cpu_count() or 1)
                got = eng.export_records().cpu().numpy()
                assert np.array_equal(got, ref), (players, pool, first, np.nonzero((got != ref).any(axis=1))[0][:5])
                assert np.array_equal(eng.counters.cpu().numpy(), cnt), (players, pool, first)
                print("ok players %d pool %d first %d: %d games finished, %d stuck" % (players, pool, first, cnt[1], cnt[6]), flush=True)
    print("all rollouts bit-exact in %.1f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
