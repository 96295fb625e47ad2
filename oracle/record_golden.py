#!/usr/bin/env python
"""Generate tests/golden/*.npz by driving the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Run as ``python -m oracle.record_golden`` from the repo root; needs
/root/reference.  Nothing here is imported by the product package.

What is recorded (SURVEY.md §8c "trace format"):

* ``trace_p{P}_{rules}.npz`` -- random-agent games of the reference ``Azul`` (``azul.py:296``
  step loop, ``game_runner.py:94-97`` RandomAgent, ``game_runner.py:113-117`` mask) for
  P in {2,3,4} and rules in {``{}`` ("default": seat-1 start, Random pool),
  ``{"first_player":"Random","tile_pool":"Lid"}`` ("lid": GameRunner's default,
  ``game_runner.py:23``)}.  Per game: initial ``next_first_player``, the 20 tile draws of every
  ``new_round`` in (display, slot) order (``azul.py:74-75``), every action and -- for the first
  ``N_FULL`` games -- the pre-step 180-bit legal mask and the full post-step state as an unpacked
  record (``azul_deep_reinforcement_learning_b200/layout.py``).  The remaining games keep only
  draws, actions, the final record and a SHA-256 over the (mask, post-state) stream.
* ``fuzz.npz`` -- 1,500 random (not necessarily reachable) representable boards for P in {2,3,4} x both pools with
  the reference's legal mask, its ``count_score`` result (dense walls: multi-bonus and long-run cases) and one
  random legal ``step`` including the draws it consumed.
* ``kat.npz`` -- the reference's own board fixtures (``tests/resources/*.json``) as unpacked
  records and the op sequences of ``tests/test_azul.py:123-331`` / ``tests/test_game_runner.py:36-69,
  89-114`` replayed on the reference, with the record after every op.

The draws are captured by swapping the ``random`` module object seen by ``azulnet.azul`` for a
recording proxy; the reference code itself is untouched.
"""
import hashlib
import os
import random as _random
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from azul_deep_reinforcement_learning_b200.layout import UnpackedLayout, unpacked_size  # noqa: E402
from oracle.ref_loader import load_reference, REFERENCE_ROOT  # noqa: E402

N_GAMES = 64
N_FULL = 8
RULESETS = {"default": {}, "lid": {"first_player": "Random", "tile_pool": "Lid"}}


class RecordingRandom:
    """Proxy for the ``random`` module inside azulnet.azul: logs what the rules engine draws."""

    def __init__(self):
        self.draws = []          # colours drawn by new_round (azul.py:78,87)
        self.first = []          # random.choice results (azul.py:37)

    def seed(self, *a):
        return _random.seed(*a)

    def randrange(self, *a):
        v = _random.randrange(*a)
        self.draws.append(int(v))
        return v

    def choices(self, population, weights=None, **kw):
        v = _random.choices(population, weights=weights, **kw)
        self.draws.append(int(v[0]))
        return v

    def choice(self, seq):
        v = _random.choice(seq)
        self.first.append(int(v))
        return v


def ref_to_record(game, total_steps=0, status=0):
    """Reference ``Azul`` object -> unpacked int32 record (layout.py)."""
    P = int(game.players)
    L = UnpackedLayout(P)
    r = np.zeros(L.size, dtype=np.int32)
    r[L.displays:L.displays + 25] = np.asarray(game.game_board_displays).reshape(-1)
    r[L.center:L.center + 6] = np.asarray(game.game_board_center).reshape(-1)
    r[L.pattern_lines:L.pattern_lines + 25 * P] = np.asarray(game.pattern_lines).reshape(-1)
    r[L.walls:L.walls + 25 * P] = np.asarray(game.walls).astype(np.int32).reshape(-1)
    r[L.floors:L.floors + P] = np.asarray(game.floors)
    r[L.score:L.score + P] = np.asarray(game.score)
    r[L.current_player] = game.current_player
    r[L.next_first_player] = game.next_first_player
    r[L.n_players] = P
    r[L.end_of_game] = int(bool(game.end_of_game))
    r[L.turn_counter] = game.turn_counter
    if game.tile_pool == "Lid":
        r[L.box:L.box + 5] = np.asarray(game.box_tiles).astype(np.int64)
        r[L.lid:L.lid + 5] = np.asarray(game.lid_tiles).astype(np.int64)
    r[L.first_player_stats:L.first_player_stats + P] = np.asarray(game.first_player_stats).astype(np.int64)
    r[L.floor_penalty:L.floor_penalty + P] = np.asarray(game.floor_penalty).astype(np.int64)
    r[L.max_combo:L.max_combo + P] = np.asarray(game.max_combo).astype(np.int64)
    r[L.completed_lines:L.completed_lines + 3 * P] = np.asarray(game.completed_lines).astype(np.int64).reshape(-1)
    r[L.total_steps] = total_steps
    r[L.status] = status
    return r


def mask_words(valid180):
    """bool[180] -> uint32[6]; word p bit (d + 6c) (game_runner.py:102-103 index = d + 6c + 30p)."""
    v = np.asarray(valid180, dtype=np.uint64).reshape(6, 30)
    return (v << np.arange(30, dtype=np.uint64)).sum(axis=1).astype(np.uint32)


def record_traces(ref, players, rules_name, out_dir):
    azul_mod = sys.modules["azulnet.azul"]
    rules = RULESETS[rules_name]
    agent = ref.RandomAgent()
    import torch

    per_game = []
    for seed in range(N_GAMES):
        proxy = RecordingRandom()
        azul_mod.random = proxy
        try:
            _random.seed(1000 * players + seed)
            game = ref.Azul(players=players, rules=dict(rules))
            first = int(game.next_first_player)
            game.new_round()
            states = [ref_to_record(game, 0)]
            masks, actions = [], []
            t = 0
            while not game.end_of_game:
                valid = ref.check_all_valid(game)
                assert valid.any(), "stuck round in golden seed %d" % seed
                a = agent.get_a_output(None, torch.from_numpy(valid.reshape(1, 180)))
                game.step(*ref.nn_deserialize(a))
                t += 1
                masks.append(mask_words(valid))
                actions.append(a)
                states.append(ref_to_record(game, t))
        finally:
            azul_mod.random = _random
        draws = np.asarray(proxy.draws, dtype=np.int8).reshape(-1, 20)
        assert draws.shape[0] == game.turn_counter
        per_game.append(dict(first=first, draws=draws, actions=np.asarray(actions, dtype=np.uint8),
                             masks=np.asarray(masks, dtype=np.uint32),
                             states=np.asarray(states, dtype=np.int32)))

    U = unpacked_size(players)
    step_off = np.cumsum([0] + [len(g["actions"]) for g in per_game]).astype(np.int64)
    round_off = np.cumsum([0] + [g["draws"].shape[0] for g in per_game]).astype(np.int64)
    digests = []
    for g in per_game:
        h = hashlib.sha256()
        for t in range(len(g["actions"])):
            h.update(g["masks"][t].astype("<u4").tobytes())
            h.update(g["states"][t + 1].astype("<i4").tobytes())
        digests.append(np.frombuffer(h.digest(), dtype=np.uint8))
    full_states = np.concatenate([g["states"] for g in per_game[:N_FULL]]).astype(np.int16)
    full_masks = np.concatenate([g["masks"] for g in per_game[:N_FULL]])
    out = dict(
        players=np.int32(players),
        tile_pool=np.int32(1 if rules.get("tile_pool") == "Lid" else 0),
        first_player_rule=np.int32(0 if rules.get("first_player") == "Random" else 1),
        n_full=np.int32(N_FULL),
        first_player=np.asarray([g["first"] for g in per_game], dtype=np.int8),
        step_offsets=step_off, round_offsets=round_off,
        actions=np.concatenate([g["actions"] for g in per_game]),
        draws=np.concatenate([g["draws"] for g in per_game]),
        initial_states=np.stack([g["states"][0] for g in per_game]).astype(np.int16),
        final_states=np.stack([g["states"][-1] for g in per_game]).astype(np.int16),
        stream_sha256=np.stack(digests),
        full_states=full_states,       # games 0..N_FULL-1: [sum(T_g + 1), U]
        full_masks=full_masks,         # games 0..N_FULL-1: [sum(T_g), 6]
    )
    assert full_states.shape[1] == U
    path = os.path.join(out_dir, "trace_p%d_%s.npz" % (players, rules_name))
    np.savez_compressed(path, **out)
    steps = int(step_off[-1])
    print("%s: %d games, %d steps, %d rounds, %.1f KiB" % (
        os.path.basename(path), N_GAMES, steps, int(round_off[-1]), os.path.getsize(path) / 1024))


# ---- known-answer scenarios: op sequences of the reference's own tests, replayed on the reference ----
# op codes: 0 move(d,c,p) azul.py:118 | 1 step(d,c,p) azul.py:296 | 2 next_player azul.py:177 |
#           3 count_score azul.py:192 | 4 is_legal_move(d,c,p) azul.py:162 (result recorded) |
#           5 new_round with recorded draws azul.py:64
KAT_SCENARIOS = [
    # tests/test_azul.py:123-165 (move)
    ("move_display_to_center", "game_first_round", [(0, 5, 0, 2)]),
    ("move_mono_display", "game_first_round", [(0, 2, 3, 4)]),
    ("move_overflow_floor", "game_first_round", [(0, 2, 3, 2)]),
    ("move_center_token", "game_first_round", [(0, 1, 0, 2), (0, 0, 1, 1)]),
    ("move_center_stack", "game_first_round", [(0, 1, 0, 3), (0, 3, 0, 3)]),
    ("move_floor_cap", "game_first_round", [(0, 3, 0, 0), (0, 4, 0, 0), (0, 1, 0, 0), (0, 2, 3, 1)]),
    # tests/test_azul.py:167-188 (is_legal_move)
    ("legal_first_round", "game_first_round", [(4, 5, 0, 2), (4, 1, 4, 2), (0, 5, 0, 2), (4, 0, 1, 1), (4, 0, 0, 0)]),
    ("legal_first_round_center", "game_first_round", [(4, 0, 0, 0)]),
    ("legal_sample_1", "game_sample_1", [(4, 0, 0, 5), (4, 0, 1, 5), (4, 5, 0, 3), (4, 5, 2, 3)]),
    # tests/test_azul.py:212-223 (end of round)
    ("eor_sample_1", "game_sample_1", [(0, 0, 0, 5)]),
    ("eor_1", "game_end_of_round_1", [(0, 0, 3, 3)]),
    # tests/test_azul.py:225-241 (end of game)
    ("eog_no", "game_end_of_round_2", [(0, 0, 4, 1), (2,), (0, 0, 0, 3), (3,)]),
    ("eog_yes", "game_end_of_round_2", [(0, 0, 0, 1), (2,), (0, 0, 4, 1), (3,)]),
    # tests/test_azul.py:243-286 (count_score)
    ("score_plain", "game_end_of_round_1", [(3,)]),
    ("score_same_round", "game_end_of_round_1", [(0, 0, 3, 3), (3,)]),
    ("score_clamp", "game_end_of_round_2", [(0, 0, 0, 0), (2,), (0, 0, 4, 0), (3,)]),
    # tests/test_azul.py:288-331 (step)
    ("step_pass_turn", "game_first_round", [(1, 5, 0, 2)]),
    ("step_illegal", "game_first_round", [(1, 1, 4, 2)]),
    ("step_rollover", "game_end_of_round_1", [(1, 0, 3, 3)]),
    ("step_game_end", "game_end_of_round_2", [(1, 0, 0, 1), (1, 0, 4, 1), (1, 0, 0, 0)]),
    # tests/test_game_runner.py:36-69 underlying engine steps (Lid pool object)
    ("step_eor3_floor", "game_end_of_round_3", [(1, 0, 3, 0)]),
]
FIXTURES = ["game_empty", "game_first_round", "game_first_round_seed_1", "game_sample_1",
            "game_end_of_round_1", "game_end_of_round_2", "game_end_of_round_3"]


def record_kats(ref, out_dir):
    azul_mod = sys.modules["azulnet.azul"]
    res = os.path.join(REFERENCE_ROOT, "tests", "resources")
    out = {}
    fixture_records, fixture_masks = [], []
    for name in FIXTURES:
        g = ref.Azul(state_file=os.path.join(res, name + ".json"))
        fixture_records.append(ref_to_record(g))
        fixture_masks.append(mask_words(ref.check_all_valid(g)))
    out["fixture_names"] = np.asarray(FIXTURES)
    out["fixture_records"] = np.stack(fixture_records).astype(np.int16)
    out["fixture_masks"] = np.stack(fixture_masks)          # game_runner.py:113-117 on each fixture

    names, fix_idx, pools = [], [], []
    ops_all, ops_off = [], [0]
    recs_all, ret_all, draws_all = [], [], []
    for pool in (0, 1):
        rules = {} if pool == 0 else {"first_player": 1, "tile_pool": "Lid"}
        for (name, fixture, ops) in KAT_SCENARIOS:
            proxy = RecordingRandom()
            azul_mod.random = proxy
            try:
                _random.seed(7)
                g = ref.Azul(rules=dict(rules))
                g.import_JSON(os.path.join(res, fixture + ".json"))
                steps = 0
                for op in ops:
                    code, args = op[0], op[1:]
                    ret, n_before = 0, len(proxy.draws)
                    try:
                        if code == 0:
                            g.move(*args)
                        elif code == 1:
                            g.step(*args)
                            steps += 1
                        elif code == 2:
                            g.next_player()
                        elif code == 3:
                            g.count_score()
                        elif code == 4:
                            ret = int(bool(g.is_legal_move(*args)))
                    except azul_mod.IllegalMove:
                        ret = -1
                    except azul_mod.GameEnded:
                        ret = -2
                    new_draws = proxy.draws[n_before:]
                    assert len(new_draws) in (0, 20)
                    draws_all.append(np.asarray(new_draws if new_draws else [-1] * 20, dtype=np.int8))
                    ops_all.append(np.asarray(list(op) + [0] * (4 - len(op)), dtype=np.int8))
                    recs_all.append(ref_to_record(g, steps))
                    ret_all.append(ret)
                    # mask after the op (undefined for ended games in the reference too: still well-formed)
            finally:
                azul_mod.random = _random
            names.append(name)
            fix_idx.append(FIXTURES.index(fixture))
            pools.append(pool)
            ops_off.append(len(ops_all))
    out["kat_names"] = np.asarray(names)
    out["kat_fixture"] = np.asarray(fix_idx, dtype=np.int8)
    out["kat_pool"] = np.asarray(pools, dtype=np.int8)
    out["kat_op_offsets"] = np.asarray(ops_off, dtype=np.int32)
    out["kat_ops"] = np.stack(ops_all)
    out["kat_records"] = np.stack(recs_all).astype(np.int16)   # record after each op
    out["kat_returns"] = np.asarray(ret_all, dtype=np.int8)    # is_legal result / -1 IllegalMove / -2 GameEnded
    out["kat_draws"] = np.stack(draws_all)                     # draws consumed by that op (-1: none)

    # seed-1 KAT (tests/test_azul.py:36-39): random.seed(1); Azul().new_round() == game_first_round_seed_1
    proxy = RecordingRandom()
    azul_mod.random = proxy
    try:
        _random.seed(1)
        g = ref.Azul()
        g.new_round()
    finally:
        azul_mod.random = _random
    out["seed1_draws"] = np.asarray(proxy.draws, dtype=np.int8)
    out["seed1_record"] = ref_to_record(g).astype(np.int16)
    path = os.path.join(out_dir, "kat.npz")
    np.savez_compressed(path, **out)
    print("kat.npz: %d fixtures, %d scenarios, %d ops, %.1f KiB" % (
        len(FIXTURES), len(names), len(ops_all), os.path.getsize(path) / 1024))


def random_board(ref, rng, players, pool):
    """A random (not necessarily reachable) but representable board: one colour per pattern row, counts within
    capacity, no pattern colour already on that wall row; dense walls so that bonuses and long adjacency runs occur."""
    g = ref.Azul(players=players, rules={"first_player": 1, "tile_pool": "Lid"} if pool else {})
    g.turn_counter = int(rng.integers(1, 9))
    g.current_player = int(rng.integers(1, players + 1))
    g.next_first_player = int(rng.integers(0, players + 1))
    density = rng.uniform(0.1, 0.9)
    g.walls = rng.random((players, 5, 5)) < density
    for p in range(players):                      # never start from an already complete row (the game would be over)
        for r in range(5):
            if g.walls[p, r].all():
                g.walls[p, r, int(rng.integers(0, 5))] = False
    g.pattern_lines = np.zeros((players, 5, 5), dtype=int)
    for p in range(players):
        for r in range(5):
            free = [c for c in range(5) if not g.walls[p, r, c]]
            if free and rng.random() < 0.75:
                c = int(rng.choice(free))
                g.pattern_lines[p, r, c] = int(rng.integers(1, r + 2)) if rng.random() < 0.5 else r + 1
    g.floors = rng.integers(0, 8, size=players)
    g.score = rng.integers(0, 90, size=players)
    empty_table = rng.random() < 0.3
    g.game_board_displays = np.zeros((5, 5), dtype=int)
    g.game_board_center = np.zeros(6, dtype=int)
    if not empty_table:
        for i in range(5):
            if rng.random() < 0.6:
                for _ in range(4):
                    g.game_board_displays[i, int(rng.integers(0, 5))] += 1
        for c in range(5):
            g.game_board_center[c] = int(rng.integers(0, 4)) if rng.random() < 0.5 else 0
        g.game_board_center[5] = int(rng.random() < 0.5)
    if pool:
        g.box_tiles = rng.integers(0, 21, size=5)
        g.lid_tiles = rng.integers(0, 21, size=5)
    return g


def record_fuzz(ref, out_dir, n_per_config=250):
    """Random boards: legal mask, count_score result, and one random legal step (with its draws) on the reference."""
    azul_mod = sys.modules["azulnet.azul"]
    rng = np.random.default_rng(20240607)
    out = {}
    for players in (2, 3, 4):
        for pool in (0, 1):
            before, masks, scored, stepped, actions, draws = [], [], [], [], [], []
            for _ in range(n_per_config):
                g = random_board(ref, rng, players, pool)
                before.append(ref_to_record(g))
                valid = ref.check_all_valid(g)
                masks.append(mask_words(valid))
                import copy
                h = copy.deepcopy(g)
                h.count_score()
                scored.append(ref_to_record(h))
                if valid.any():
                    a = int(rng.choice(np.nonzero(valid)[0]))
                    proxy = RecordingRandom()
                    azul_mod.random = proxy
                    try:
                        _random.seed(int(rng.integers(0, 2 ** 31)))
                        g.step(*ref.nn_deserialize(a))
                    finally:
                        azul_mod.random = _random
                    actions.append(a)
                    draws.append(np.asarray(proxy.draws if proxy.draws else [-1] * 20, dtype=np.int8))
                    stepped.append(ref_to_record(g, 1))
                else:
                    actions.append(255)
                    draws.append(np.full(20, -1, np.int8))
                    stepped.append(ref_to_record(g, 0))
            key = "p%d_%s" % (players, "lid" if pool else "default")
            out[key + "_before"] = np.stack(before).astype(np.int16)
            out[key + "_mask"] = np.stack(masks)
            out[key + "_scored"] = np.stack(scored).astype(np.int16)
            out[key + "_action"] = np.asarray(actions, dtype=np.uint8)
            out[key + "_draws"] = np.stack(draws)
            out[key + "_stepped"] = np.stack(stepped).astype(np.int16)
    path = os.path.join(out_dir, "fuzz.npz")
    np.savez_compressed(path, **out)
    print("fuzz.npz: %d boards, %.1f KiB" % (6 * n_per_config, os.path.getsize(path) / 1024))


def main():
    ref = load_reference()
    out_dir = os.path.join(REPO, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    record_kats(ref, out_dir)
    record_fuzz(ref, out_dir)
    for players in (2, 3, 4):
        for rules_name in RULESETS:
            record_traces(ref, players, rules_name, out_dir)


if __name__ == "__main__":
    main()
