"""Pins the C restatement (oracle/azul_oracle.c) to the reference.

CPU-only.  The golden vectors under tests/golden were recorded from the UNMODIFIED reference by
oracle/record_golden.py; here every recorded game is replayed through the restatement with the
recorded tile draws and actions, and every pre-step legal mask and post-step record must match
bit for bit (full streams for the first 8 games of each configuration, SHA-256 of the stream and
the final record for all 64).  The known-answer scenarios mirror the reference's tests/test_azul.py.
"""
import numpy as np
import pytest

from azul_deep_reinforcement_learning_b200.layout import UnpackedLayout
from oracle import oracle as O
from tests.helpers import TRACE_CONFIGS, TraceGame, load_kat, load_trace, stream_digest


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert [hex(x) for x in O.philox4x32_10([0] * 4, [0] * 2)] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    assert [hex(x) for x in O.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2)] == \
        ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
    assert [hex(x) for x in O.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                                           [0xa4093822, 0x299f31d0])] == \
        ['0xd16cfe09', '0x94fdcceb', '0x5001e420', '0x24126ea1']


@pytest.mark.parametrize("players,rules", TRACE_CONFIGS)
def test_replay_golden_traces(players, rules):
    tr = load_trace(players, rules)
    pool = int(tr["tile_pool"])
    n_games = len(tr["first_player"])
    for i in range(n_games):
        tg = TraceGame(tr, i)
        g = O.Game(players, pool, tg.first_player)
        g.new_round(tg.draws[0])
        assert np.array_equal(g.rec, tg.initial), "initial state game %d" % i
        rnd = 1
        masks, states = [], []
        for t, a in enumerate(tg.actions):
            m = g.legal_mask()
            turn_before = g.rec[UnpackedLayout(players).turn_counter]
            draws = tg.draws[rnd] if rnd < len(tg.draws) else np.full(20, -1, np.int8)
            rc = g.step(int(a), draws)
            assert rc == 0
            if g.rec[UnpackedLayout(players).turn_counter] != turn_before:
                rnd += 1
            masks.append(m.copy())
            states.append(g.rec.copy())
            if tg.full:
                assert np.array_equal(m, tg.masks[t]), "mask game %d step %d" % (i, t)
                assert np.array_equal(g.rec, tg.states[t + 1]), "state game %d step %d" % (i, t)
        assert rnd == len(tg.draws)
        assert np.array_equal(g.rec, tg.final)
        assert stream_digest(masks, states) == tg.sha
        assert g.step(0, None) == -2          # GameEnded, azul.py:298-299


def test_reference_fixture_masks():
    kat = load_kat()
    for name, rec, mask in zip(kat["fixture_names"], kat["fixture_records"], kat["fixture_masks"]):
        g = O.Game(2, 0, record=rec.astype(np.int32))
        assert np.array_equal(g.legal_mask(), mask), name


def test_seed1_new_round():
    # tests/test_azul.py:36-39 : random.seed(1); Azul().new_round() == game_first_round_seed_1.json
    kat = load_kat()
    g = O.Game(2, 0, 1)
    g.new_round(kat["seed1_draws"])
    assert np.array_equal(g.rec, kat["seed1_record"].astype(np.int32))
    names = list(kat["fixture_names"])
    fix = kat["fixture_records"][names.index("game_first_round_seed_1")].astype(np.int32)
    L = UnpackedLayout(2)
    # the JSON fixture carries no statistics (azul.py:105-116): compare the __eq__ fields (azul.py:63)
    assert np.array_equal(g.rec[:L.box], fix[:L.box])


def test_known_answer_scenarios():
    """Op sequences of the reference's tests/test_azul.py:123-331, replayed op by op."""
    kat = load_kat()
    n = len(kat["kat_names"])
    for s in range(n):
        pool = int(kat["kat_pool"][s])
        rec0 = kat["fixture_records"][int(kat["kat_fixture"][s])].astype(np.int32).copy()
        L = UnpackedLayout(2)
        if pool == 1:
            rec0[L.box:L.box + 5] = 20          # Azul(rules={"tile_pool":"Lid"}) then import_JSON
        g = O.Game(2, pool, record=rec0)
        for k in range(int(kat["kat_op_offsets"][s]), int(kat["kat_op_offsets"][s + 1])):
            code, a, b, c = [int(x) for x in kat["kat_ops"][k]]
            ret = 0
            if code == 0:
                g.move(a, b, c)
            elif code == 1:
                draws = kat["kat_draws"][k]
                ret = g.step(a + 6 * b + 30 * c, draws if draws[0] >= 0 else np.full(20, -1, np.int8))
            elif code == 2:
                g.next_player()
            elif code == 3:
                g.count_score()
            elif code == 4:
                ret = int(g.is_legal_move(a, b, c))
            assert ret == int(kat["kat_returns"][k]), (kat["kat_names"][s], k)
            assert np.array_equal(g.rec, kat["kat_records"][k].astype(np.int32)), (kat["kat_names"][s], k)


def test_reference_count_score_values():
    """The literal expectations of tests/test_azul.py:243-286 (not just recorded outputs)."""
    kat = load_kat()
    names = list(kat["fixture_names"])
    L = UnpackedLayout(2)

    def load(n):
        return O.Game(2, 0, record=kat["fixture_records"][names.index(n)].astype(np.int32).copy())

    g = load("game_end_of_round_1")
    prev = g.rec[L.score:L.score + 2].copy()
    g.count_score()
    assert list(g.rec[L.score:L.score + 2] - prev) == [5 + 5 + 1 - 2, 4 + 2 + 3 - 8]
    assert list(g.rec[L.floors:L.floors + 2]) == [0, 0]
    g = load("game_end_of_round_2")
    prev = g.rec[L.score:L.score + 2].copy()
    g.move(0, 4, 1); g.next_player(); g.move(0, 0, 3); g.count_score()
    assert list(g.rec[L.score:L.score + 2] - prev) == [5 + 7 + 10, 5 + 7]
    g = load("game_end_of_round_2")
    g.move(0, 0, 1); g.next_player(); g.move(0, 4, 1); g.count_score()
    assert list(g.rec[L.score:L.score + 2]) == [2 - 2, 5 + 2]
    assert g.is_end_of_game()
    g = load("game_end_of_round_2")
    g.move(0, 0, 0); g.next_player(); g.move(0, 4, 0); g.count_score()
    assert list(g.rec[L.score:L.score + 2]) == [0, 0]


def test_random_agent_sampler_only_legal_and_weighted():
    """game_runner.py:87-97: floor actions (p = 0) carry weight 0.01, others 1.0 -> 1 : 100."""
    rng = np.random.default_rng(0)
    mask = np.array([0b101, 0b11, 0, 0, 0, 0], dtype=np.uint32)   # 2 floor actions, 2 heavy actions
    counts = {}
    n = 200000
    for w in rng.integers(0, 2 ** 32, size=n, dtype=np.uint64):
        a = O.random_action(mask, int(w))
        counts[a] = counts.get(a, 0) + 1
    assert set(counts) == {0, 2, 30, 31}
    heavy = counts[30] + counts[31]
    light = counts[0] + counts[2]
    assert abs(light / n - 2 / 202) < 0.002 and abs(heavy / n - 200 / 202) < 0.002
    assert O.random_action(np.zeros(6, np.uint32), 123) == -1


def test_random_boards_against_reference():
    """tests/golden/fuzz.npz: random representable boards (dense walls, multi-bonus scoring, floor caps, full-row
    quirk) with the reference's legal mask, count_score result and one random legal step."""
    from tests.helpers import FUZZ_CONFIGS, fuzz_key, load_fuzz
    fz = load_fuzz()
    for players, pool in FUZZ_CONFIGS:
        k = fuzz_key(players, pool)
        for i in range(len(fz[k + "_action"])):
            rec = fz[k + "_before"][i].astype(np.int32)
            g = O.Game(players, pool, record=rec.copy())
            assert np.array_equal(g.legal_mask(), fz[k + "_mask"][i]), (k, i)
            h = g.copy()
            h.count_score()
            assert np.array_equal(h.rec, fz[k + "_scored"][i].astype(np.int32)), (k, i)
            a = int(fz[k + "_action"][i])
            if a != 255:
                assert g.step(a, fz[k + "_draws"][i]) == 0
                assert np.array_equal(g.rec, fz[k + "_stepped"][i].astype(np.int32)), (k, i)
