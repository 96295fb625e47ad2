"""``azulnet.Azul`` drop-in: one game = a batch of one on the GPU.

Public surface and attribute set follow the reference class (``azulnet/azul.py:17-315``): numpy
attributes that callers read and assign, the same method names and argument meaning, the same
exceptions raised before any mutation.  Every rule is executed by the CUDA kernels behind
``include/azb.h``: before an operation the (possibly user-modified) attributes are packed on the
device (``azb_import_state``), afterwards the result is unpacked back into the attributes.  Nothing
here evaluates a rule on the host; the only host logic is drawing the 20 tile colours of a new
round from Python's ``random`` in the reference's call order (``azul.py:74-89``) so that seeded
runs (``tests/test_azul.py:36-39``) reproduce, and injecting them into ``azb_new_round``.
"""
import json
import random

import numpy as np
import torch

from ..layout import (FIRST_PLAYER_RANDOM, STATUS_BAD_IMPORT, TILE_POOL_LID, TILE_POOL_RANDOM, UnpackedLayout)


class IllegalMove(Exception):
    pass


class GameEnded(Exception):
    pass


class IllegalRule(Exception):
    pass


def _parse_rules(players, rules):
    """azul.py:35-56 -> (tile_pool, first_player rule, initial next_first_player)."""
    first_rule, nfp = 1, 1
    if "first_player" in rules:
        fp = rules["first_player"]
        if fp == "Random":
            first_rule = FIRST_PLAYER_RANDOM
            nfp = random.choice(list(range(1, players + 1)))         # same draw as azul.py:37
        elif type(fp) == int and 1 <= fp <= players:
            first_rule = nfp = fp
        else:
            raise IllegalRule
    pool = TILE_POOL_RANDOM
    if "tile_pool" in rules:
        if rules["tile_pool"] == "Random":
            pool = TILE_POOL_RANDOM
        elif rules["tile_pool"] == "Lid":
            pool = TILE_POOL_LID
        else:
            raise IllegalRule
    return pool, first_rule, nfp


_ENGINES = {}


def _engine(players, pool):
    """One single-game device engine per (players, tile pool); it keeps no state between operations
    (every façade call re-imports the host view), so all ``Azul`` objects of a kind share it."""
    key = (players, pool)
    if key not in _ENGINES:
        from ..engine import BatchedAzul
        _ENGINES[key] = BatchedAzul(1, players, pool, 1, reset=False)
    return _ENGINES[key]


class Azul:
    def __init__(self, players=2, state_file=None, rules={}):
        self.players = players
        self.rules = rules
        pool, first_rule, nfp = _parse_rules(players, rules)
        self.tile_pool = "Lid" if pool == TILE_POOL_LID else "Random"
        self._pool = pool
        self._eng = _engine(players, pool)
        self._layout = UnpackedLayout(players)
        self.game_board_displays = np.zeros((5, 5), dtype=int)
        self.game_board_center = np.zeros(6, dtype=int)
        self.pattern_lines = np.zeros((players, 5, 5), dtype=int)
        self.walls = np.zeros((players, 5, 5), dtype=bool)
        self.floors = np.zeros(players, dtype=int)
        self.score = np.zeros(players, dtype=int)
        self.current_player = 0
        self.end_of_game = False
        self.turn_counter = 0
        self.first_player_stats = np.zeros(players)
        self.floor_penalty = np.zeros(players)
        self.max_combo = np.zeros(players)
        self.completed_lines = np.zeros((players, 3))
        self.next_first_player = nfp
        if pool == TILE_POOL_LID:
            self.box_tiles = np.array([20, 20, 20, 20, 20])
            self.lid_tiles = np.array([0, 0, 0, 0, 0])
        self._total_steps = 0
        if state_file is not None:
            self.import_JSON(state_file)

    # ---- host view <-> device state ----------------------------------------------------------
    def _record(self):
        L, P = self._layout, self.players
        r = np.zeros(L.size, dtype=np.int32)
        r[L.displays:L.displays + 25] = np.asarray(self.game_board_displays).reshape(-1)
        r[L.center:L.center + 6] = np.asarray(self.game_board_center).reshape(-1)
        r[L.pattern_lines:L.pattern_lines + 25 * P] = np.asarray(self.pattern_lines).reshape(-1)
        r[L.walls:L.walls + 25 * P] = np.asarray(self.walls).astype(np.int32).reshape(-1)
        r[L.floors:L.floors + P] = np.asarray(self.floors)
        r[L.score:L.score + P] = np.asarray(self.score)
        r[L.current_player] = self.current_player
        r[L.next_first_player] = self.next_first_player
        r[L.n_players] = P
        r[L.end_of_game] = int(bool(self.end_of_game))
        r[L.turn_counter] = self.turn_counter
        if self._pool == TILE_POOL_LID:
            r[L.box:L.box + 5] = np.asarray(self.box_tiles).astype(np.int64)
            r[L.lid:L.lid + 5] = np.asarray(self.lid_tiles).astype(np.int64)
        r[L.first_player_stats:L.first_player_stats + P] = np.asarray(self.first_player_stats).astype(np.int64)
        r[L.floor_penalty:L.floor_penalty + P] = np.asarray(self.floor_penalty).astype(np.int64)
        r[L.max_combo:L.max_combo + P] = np.asarray(self.max_combo).astype(np.int64)
        r[L.completed_lines:L.completed_lines + 3 * P] = np.asarray(self.completed_lines).astype(np.int64).reshape(-1)
        r[L.total_steps] = self._total_steps
        return r

    def _to_device(self):
        ok = self._eng.import_records(self._record()[None, :])
        if not bool(ok.cpu()[0]):
            raise ValueError("board state cannot be represented in the packed device format "
                             "(status %d: two colours in a pattern row or a count out of range)" % STATUS_BAD_IMPORT)

    def _from_device(self):
        L, P = self._layout, self.players
        r = self._eng.export_records().cpu().numpy()[0].astype(np.int64)
        self.game_board_displays = r[L.displays:L.displays + 25].reshape(5, 5).astype(int)
        self.game_board_center = r[L.center:L.center + 6].astype(int)
        self.pattern_lines = r[L.pattern_lines:L.pattern_lines + 25 * P].reshape(P, 5, 5).astype(int)
        self.walls = r[L.walls:L.walls + 25 * P].reshape(P, 5, 5).astype(bool)
        self.floors = r[L.floors:L.floors + P].astype(int)
        self.score = r[L.score:L.score + P].astype(int)
        self.current_player = int(r[L.current_player])
        self.next_first_player = int(r[L.next_first_player])
        self.end_of_game = bool(r[L.end_of_game])
        self.turn_counter = int(r[L.turn_counter])
        if self._pool == TILE_POOL_LID:
            self.box_tiles = r[L.box:L.box + 5].astype(int)
            self.lid_tiles = r[L.lid:L.lid + 5].astype(int)
        self.first_player_stats = r[L.first_player_stats:L.first_player_stats + P].astype(float)
        self.floor_penalty = r[L.floor_penalty:L.floor_penalty + P].astype(float)
        self.max_combo = r[L.max_combo:L.max_combo + P].astype(float)
        self.completed_lines = r[L.completed_lines:L.completed_lines + 3 * P].reshape(P, 3).astype(float)
        self._total_steps = int(r[L.total_steps])

    def _mask_bits(self):
        m = self._eng.legal_mask().cpu().numpy().astype(np.uint32)[:, 0]
        return m

    # ---- reference API -------------------------------------------------------------------------
    def __eq__(self, other):
        # same field set as azul.py:63 (box/lid, tile_pool and the statistics are not compared)
        return (np.array_equal(self.game_board_displays, other.game_board_displays)
                and np.array_equal(self.game_board_center, other.game_board_center)
                and np.array_equal(self.pattern_lines, other.pattern_lines)
                and np.array_equal(self.walls, other.walls)
                and np.array_equal(self.floors, other.floors)
                and np.array_equal(self.score, other.score)
                and self.current_player == other.current_player
                and self.next_first_player == other.next_first_player
                and self.players == other.players
                and self.end_of_game == other.end_of_game
                and self.turn_counter == other.turn_counter)

    def _draw_round(self):
        """The 20 colours of a new round from Python ``random``, in the reference's call order."""
        draws = []
        if self._pool == TILE_POOL_RANDOM:
            for _ in range(20):
                draws.append(random.randrange(0, 5, 1))                  # azul.py:78
            return draws
        box = [int(x) for x in self.box_tiles]
        lid = [int(x) for x in self.lid_tiles]
        for _ in range(20):
            if sum(box) == 0:                                            # azul.py:81-83
                box, lid = lid, [0] * 5
            total = sum(box)
            if total == 0:
                # the reference divides 0/0 here and random.choices raises (its azul.py:86 TODO)
                raise ValueError("Total of weights must be finite")
            c = random.choices([0, 1, 2, 3, 4], weights=[b / total for b in box])[0]   # azul.py:87
            draws.append(c)
            box[c] -= 1
        return draws

    def _new_round_on_device(self):
        draws = self._draw_round()
        self._eng.new_round(torch.tensor(draws, dtype=torch.int8).reshape(1, 20))

    def new_round(self):
        self._to_device()
        self._new_round_on_device()
        self._from_device()

    def import_JSON(self, path):
        with open(path) as f:
            data = json.load(f)
        self.game_board_displays = np.array(data["game_board_displays"], dtype=int)
        self.game_board_center = np.array(data["game_board_center"], dtype=int)
        self.pattern_lines = np.array(data["pattern_lines"], dtype=int)
        self.walls = np.array(data["walls"], dtype=bool)
        self.floors = np.array(data["floors"], dtype=int)
        self.score = np.array(data["score"], dtype=int)
        self.current_player = data["current_player"]
        self.next_first_player = data["next_first_player"]
        self.players = data["players"]
        self.turn_counter = data["turn_counter"]

    def export_JSON(self, path):
        with open(path, "w+") as f:
            f.write(json.dumps({
                "game_board_displays": np.asarray(self.game_board_displays).tolist(),
                "game_board_center": np.asarray(self.game_board_center).tolist(),
                "pattern_lines": np.asarray(self.pattern_lines).tolist(),
                "walls": np.asarray(self.walls).tolist(),
                "floors": np.asarray(self.floors).tolist(),
                "score": np.asarray(self.score).tolist(),
                "current_player": self.current_player,
                "next_first_player": self.next_first_player,
                "players": self.players,
                "turn_counter": self.turn_counter,
            }))

    def move(self, display, color, pattern):
        self._to_device()
        seat = (self.current_player - 1) % self.players
        other = [c for c in range(5) if c != color and self.pattern_lines[seat, pattern - 1, c] != 0] if pattern else []
        if other:
            raise ValueError("move() would put two colours on one pattern line; the packed device state "
                             "cannot hold that (the reference only reaches it through an illegal direct move())")
        a = display + 6 * color + 30 * pattern
        self._eng.move(torch.tensor([a], dtype=torch.uint8))
        self._from_device()

    def is_legal_move(self, display, color, pattern):
        self._to_device()
        m = self._mask_bits()
        return bool((int(m[pattern]) >> (display + 6 * color)) & 1)

    def next_player(self):
        self._to_device()
        self._eng.next_player()
        self._from_device()

    def _flags(self):
        return int(self._eng.round_flags().cpu()[0])

    def is_end_of_round(self):
        self._to_device()
        return bool(self._flags() & 1)

    def is_end_of_game(self):
        self._to_device()
        return bool(self._flags() & 2)

    def count_score(self):
        self._to_device()
        self._eng.count_score()
        self._from_device()

    def step(self, display, color, pattern):
        if self.end_of_game:
            raise GameEnded                                              # azul.py:298-299
        self._to_device()
        a = display + 6 * color + 30 * pattern
        m = self._mask_bits()
        if not (0 <= display <= 5 and 0 <= color <= 4 and 0 <= pattern <= 5) or not (int(m[pattern]) >> (display + 6 * color)) & 1:
            raise IllegalMove                                            # azul.py:301-302, state untouched
        self._eng.move(torch.tensor([a], dtype=torch.uint8))             # azul.py:304
        self._total_steps += 1
        steps = self._total_steps                                        # azb_move does not count steps: carried on the host
        ended = False
        if self._flags() & 1:                                            # azul.py:306
            self._eng.count_score()                                      # azul.py:307
            if self._flags() & 2:                                        # azul.py:308-309
                ended = True
            else:
                if self._pool == TILE_POOL_LID:
                    self._from_device()                                  # box / lid for the weighted draws
                self._new_round_on_device()                              # azul.py:311
        else:
            self._eng.next_player()                                      # azul.py:313
        self._from_device()
        self._total_steps = steps
        if ended:
            self.end_of_game = True

    def score_preview(self):
        """Scores after a ``count_score`` on a copy (what ``GameRunner.step`` needs, game_runner.py:48-50)."""
        self._to_device()
        return self._eng.score_preview().cpu().numpy()[:, 0].astype(int)

    def observation(self, perspective=0):
        self._to_device()
        return self._eng.observe(perspective).cpu().numpy()[0]

    def legal_mask_bool(self):
        from ..engine import mask_to_bool
        self._to_device()
        return mask_to_bool(self._eng.legal_mask()).cpu().numpy()[0]

    def get_statistics(self):
        # azul.py:314-315
        return {"player_score": self.score[0], "opponent_score": self.score[1], "rounds": self.turn_counter,
                "percent_first_player": self.first_player_stats[0] / self.first_player_stats.sum() * 100,
                "floor_penalty": -self.floor_penalty[0], "max_combo": self.max_combo[0],
                "completed_rows": self.completed_lines[0, 0], "completed_columns": self.completed_lines[0, 2],
                "completed_colors": self.completed_lines[0, 1], "win_percent": self.score[0] > self.score[1]}
