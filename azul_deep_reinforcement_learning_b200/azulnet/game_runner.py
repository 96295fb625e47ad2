"""``azulnet.game_runner`` drop-in: the 2-player env wrapper on the CUDA engine.

Mirrors the reference module's surface (``azulnet/game_runner.py:9-117``): ``GameRunner`` with
``step / get_state / get_valid_moves / reset / opponent_move``, its nested ``GameStatistics``,
``RandomAgent``, the action codec and ``check_all_valid``.  The rules, the legal mask, the
observation and the reward preview all come from kernels (``azb_move``, ``azb_legal_mask``,
``azb_observe``, ``azb_score_preview`` ...) through the ``Azul`` façade; the host only sequences
calls and keeps the statistics buffers.
"""
import random

import numpy as np
import torch

from .azul import Azul

_STAT_KEYS = ["player_score", "opponent_score", "rounds", "percent_first_player", "floor_penalty", "max_combo",
              "completed_rows", "completed_columns", "completed_colors", "win_percent"]


def nn_serialize(display, color, pattern):
    """(source 0..5, colour 0..4, destination 0..5) -> action 0..179 (game_runner.py:102-103)."""
    return display + 6 * color + 30 * pattern


def nn_deserialize(i):
    """Inverse of :func:`nn_serialize` (game_runner.py:107-111)."""
    i = int(i)
    return (i % 6, (i // 6) % 5, i // 30)


def check_all_valid(game):
    """180-entry legal mask of ``game`` for the player to move (game_runner.py:113-117), from K2."""
    return game.legal_mask_bool()


class RandomAgent:
    """Uniform over legal actions, straight-to-floor actions down-weighted 100x (game_runner.py:87-97)."""

    def __init__(self):
        self.weight_table = np.ones(180)
        self.weight_table[:30] = 0.01                      # the 30 actions with destination 0 (floor)

    def get_a_output(self, state, valid_moves):
        w = self.weight_table * valid_moves.numpy()[0]
        return random.choices(range(180), weights=w)[0]


class GameRunner:
    class GameStatistics:
        """Per-game statistics buffer; ``get_stats`` folds the buffer into one mean per key."""

        def __init__(self):
            self.statisticsBuffer = {k: np.empty(0) for k in _STAT_KEYS}
            self.statistics = {k: np.empty(0) for k in _STAT_KEYS}

        def update(self, statistics):
            for k, v in statistics.items():
                self.statisticsBuffer[k] = np.append(self.statisticsBuffer[k], v)

        def get_stats(self):
            for k in self.statistics:
                if len(self.statisticsBuffer[k]) > 0:
                    self.statistics[k] = np.append(self.statistics[k], self.statisticsBuffer[k].mean())
                    self.statisticsBuffer[k] = np.empty(0)
            return self.statistics

    def __init__(self, opponent=None, rules={"first_player": "Random", "tile_pool": "Lid"}):
        self.game = Azul(rules=rules)
        self.rules = rules
        self.game_statistics = GameRunner.GameStatistics()
        self.opponent = opponent if opponent is not None else RandomAgent()
        self.game.new_round()
        self.player_score = 0
        self.move_counter = 0

    def opponent_move(self):
        state = self.get_state(perspective=self.game.current_player - 1)
        valid = torch.from_numpy(self.get_valid_moves().reshape(1, 180))
        action = self.opponent.get_a_output(state, valid)
        self.game.step(*nn_deserialize(action))
        self.move_counter += 1

    def step(self, i):
        self.game.step(*nn_deserialize(i))
        self.move_counter += 1
        # the opponent also plays seat 1's forced moves (fewer than two legal actions), game_runner.py:46
        while (self.game.current_player != 1 or np.count_nonzero(self.get_valid_moves()) < 2) \
                and not self.game.is_end_of_game():
            self.opponent_move()
        preview = self.game.score_preview()                 # count_score on a copy, game_runner.py:48-50
        new_player_score = int(preview[0] - preview[1])
        reward = new_player_score - self.player_score
        self.player_score = new_player_score
        done = self.game.is_end_of_game()
        if done:
            self.game_statistics.update(self.game.get_statistics())
        return reward, done

    def get_state(self, perspective=0):
        return self.game.observation(perspective).astype(np.int64)

    def get_valid_moves(self):
        return check_all_valid(self.game)

    def reset(self):
        self.game = Azul(rules=self.rules)
        self.game.new_round()
        self.player_score = 0
        self.move_counter = 0
        while self.game.current_player != 1:
            self.opponent_move()
