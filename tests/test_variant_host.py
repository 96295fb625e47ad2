"""The opt-in "factory count by player count" rules (csrc/azb_variant.cuh, compiled for the host by tests/harness) against
the C oracle with ao_set_factories: seeded Philox rollouts (records and counters), legal masks, single steps with illegal
actions -- for 7 / 9 displays, and for 5 displays, where the variant code must also equal the DEFAULT packed rules
(azb_rules.cuh), which are pinned to the reference.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import harness as H

VARIANTS = [(3, 7), (4, 9)]
ALL = [(2, 5), (3, 5), (4, 5)] + VARIANTS


@pytest.mark.parametrize("players,factories", ALL)
@pytest.mark.parametrize("pool", [0, 1])
@pytest.mark.parametrize("first_rule", [0, 1])
def test_variant_rollout_equals_oracle(players, factories, pool, first_rule):
    n, k, seed, gid0 = 96, 230, 0xBEEF + factories, 1000
    with O.factories(factories):
        recs = O.fresh_records(n, players, pool, first_rule, seed, gid0)
        mine = recs.copy()
        want_cnt = O.rollout_random(recs, players, pool, first_rule, seed, gid0, k)
        got_cnt = H.v_rollout(mine, players, factories, pool, first_rule, seed, gid0, k)
        assert np.array_equal(mine, recs)
        assert np.array_equal(got_cnt, want_cnt) and want_cnt[1] > 0
        for i in range(0, n, 7):                                   # masks of the final states
            _, m = H.v_op(mine[i].copy(), players, factories, pool, H.V_OP_MASK)
            assert np.array_equal(m, O.legal_mask64(recs[i], players))
        if pool == 1:                                              # 100 tiles stay in the game (Lid pool)
            S, U = factories + 1, recs.shape[1]
            off = 5 * factories
            tiles = recs[:, :off + 5].sum(1) + recs[:, off + 6: off + 6 + 25 * players].sum(1)
            walls = recs[:, off + 6 + 25 * players: off + 6 + 50 * players].sum(1)
            sc = off + 6 + 52 * players
            box_lid = recs[:, sc + 5: sc + 15].sum(1)
            # tiles on the floor line are not in the record by colour: they went to the lid when they fell (azul.py:156-161)
            assert ((tiles + walls + box_lid == 100) | (recs[:, sc + 16 + 6 * players] != 0)).all()


@pytest.mark.parametrize("players", [2, 3, 4])
@pytest.mark.parametrize("pool", [0, 1])
def test_variant_with_five_displays_equals_default_rules(players, pool):
    """F = 5: the variant's packed rules == the default packed rules (which replay the reference's golden traces)."""
    n, k, seed = 64, 300, 77
    recs = O.fresh_records(n, players, pool, 0, seed, 5)
    a, b = recs.copy(), recs.copy()
    ca = H.rollout(a, players, pool, 0, seed, 5, k)
    cb = H.v_rollout(b, players, 5, pool, 0, seed, 5, k)
    assert np.array_equal(a, b) and np.array_equal(ca, cb)


@pytest.mark.parametrize("players,factories", VARIANTS)
def test_variant_step_status_and_injected_draws(players, factories):
    S = factories + 1
    with O.factories(factories):
        rng = np.random.default_rng(5)
        rec = O.fresh_records(1, players, 1, 1, 9, 0)[0]
        g = O.Game(players, 1, record=rec.copy())
        mine = rec.copy()
        for t in range(400):
            m = O.legal_mask64(g.rec, players)
            if g.rec[5 * factories + 6 + 52 * players + 3]:        # end_of_game
                rc, _ = H.v_op(mine, players, factories, 1, H.V_OP_STEP, a=0)
                assert rc == -2
                break
            legal = [p * 5 * S + b for p in range(6) for b in range(5 * S) if (int(m[p]) >> b) & 1]
            illegal = [a for a in range(30 * S) if a not in set(legal)]
            if illegal:
                bad = int(rng.choice(illegal))
                before = mine.copy()
                rc, _ = H.v_op(mine, players, factories, 1, H.V_OP_STEP, a=bad)
                assert rc == -1 and np.array_equal(mine, before) and g.step(bad) == -1
            a = int(rng.choice(legal))
            draws = rng.integers(0, 5, size=4 * factories).astype(np.int8)
            # injected draws must respect the box: let the oracle decide with its own Philox refill instead
            assert g.step(a, None, 9, 0) == 0
            rc, mm = H.v_op(mine, players, factories, 1, H.V_OP_STEP, a=a, seed=9, gid=0)
            assert rc == 0 and np.array_equal(mine, g.rec), t
            assert np.array_equal(mm, O.legal_mask64(g.rec, players))
        else:
            raise AssertionError("game did not end")
