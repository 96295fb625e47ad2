"""ctypes wrapper around the C restatement (oracle/azul_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never from the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libazul_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the restatement with gcc (seconds)."""
    src = os.path.join(_HERE, "azul_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c11", "-shared", "-o", _SO, src, "-lpthread"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        i32p = ctypes.POINTER(ctypes.c_int32)
        u32p = ctypes.POINTER(ctypes.c_uint32)
        i8p = ctypes.POINTER(ctypes.c_int8)
        i64p = ctypes.POINTER(ctypes.c_int64)
        I, U64, U32 = ctypes.c_int, ctypes.c_uint64, ctypes.c_uint32
        L.ao_record_size.argtypes = [I]
        L.ao_init.argtypes = [i32p, I, I, I]
        L.ao_new_round.argtypes = [i32p, I, I, i8p]
        L.ao_reset_philox.argtypes = [i32p, I, I, I, U64, U32]
        L.ao_move.argtypes = [i32p, I, I, I, I, I]
        L.ao_is_legal_move.argtypes = [i32p, I, I, I, I]
        L.ao_next_player.argtypes = [i32p, I]
        L.ao_is_end_of_round.argtypes = [i32p, I]
        L.ao_is_end_of_game.argtypes = [i32p, I]
        L.ao_count_score.argtypes = [i32p, I, I]
        L.ao_score_preview.argtypes = [i32p, I, I, i32p]
        L.ao_legal_mask.argtypes = [i32p, I, u32p]
        L.ao_step.argtypes = [i32p, I, I, I, i8p, U64, U32]
        L.ao_random_action.argtypes = [u32p, U32]
        L.ao_opponent_random.argtypes = [i32p, I, I, U64, U32, I, u32p]
        L.ao_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.ao_set_factories.argtypes = [I]
        L.ao_legal_mask64.argtypes = [i32p, I, ctypes.POINTER(ctypes.c_uint64)]
        L.ao_observe.argtypes = [i32p, I, I, i32p]
        L.ao_runner_continues.argtypes = [i32p, I, I]
        L.ao_statistics.argtypes = [i32p, I, i32p]
        L.ao_rollout_random.argtypes = [i32p, ctypes.c_int64, I, I, I, U64, U32, I, i64p]
        L.ao_rollout_random_mt.argtypes = [i32p, ctypes.c_int64, I, I, I, U64, U32, I, i64p, I]
        for f in ("ao_init", "ao_new_round", "ao_reset_philox", "ao_move", "ao_next_player", "ao_count_score",
                  "ao_score_preview", "ao_legal_mask", "ao_philox4x32_10", "ao_rollout_random",
                  "ao_rollout_random_mt", "ao_observe", "ao_statistics", "ao_set_factories", "ao_legal_mask64"):
            getattr(L, f).restype = None
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _rec(rec):
    assert rec.dtype == np.int32 and rec.flags.c_contiguous
    return _p(rec, ctypes.c_int32)


def _draws(draws):
    if draws is None:
        return None
    d = np.ascontiguousarray(draws, dtype=np.int8)
    assert d.size == 20
    return d, _p(d, ctypes.c_int8)


class Game:
    """One game held as an unpacked int32 record; methods mirror the reference ``Azul`` object."""

    def __init__(self, players=2, tile_pool=0, first_player=1, record=None):
        self.players, self.tile_pool = players, tile_pool
        self.rec = np.zeros(lib().ao_record_size(players), dtype=np.int32)
        if record is None:
            lib().ao_init(_rec(self.rec), players, tile_pool, first_player)
        else:
            self.rec[:] = np.asarray(record, dtype=np.int32)

    def copy(self):
        return Game(self.players, self.tile_pool, record=self.rec.copy())

    def new_round(self, draws):
        d, dp = _draws(draws)
        lib().ao_new_round(_rec(self.rec), self.players, self.tile_pool, dp)

    def reset_philox(self, first_rule, seed, gid):
        lib().ao_reset_philox(_rec(self.rec), self.players, self.tile_pool, first_rule, seed, gid)

    def move(self, d, c, p):
        lib().ao_move(_rec(self.rec), self.players, self.tile_pool, d, c, p)

    def is_legal_move(self, d, c, p):
        return bool(lib().ao_is_legal_move(_rec(self.rec), self.players, d, c, p))

    def next_player(self):
        lib().ao_next_player(_rec(self.rec), self.players)

    def is_end_of_round(self):
        return bool(lib().ao_is_end_of_round(_rec(self.rec), self.players))

    def is_end_of_game(self):
        return bool(lib().ao_is_end_of_game(_rec(self.rec), self.players))

    def count_score(self):
        lib().ao_count_score(_rec(self.rec), self.players, self.tile_pool)

    def score_preview(self):
        out = np.zeros(self.players, dtype=np.int32)
        lib().ao_score_preview(_rec(self.rec), self.players, self.tile_pool, _p(out, ctypes.c_int32))
        return out

    def legal_mask(self):
        m = np.zeros(6, dtype=np.uint32)
        lib().ao_legal_mask(_rec(self.rec), self.players, _p(m, ctypes.c_uint32))
        return m

    def step(self, action, draws=None, seed=0, gid=0):
        """0 ok, -1 IllegalMove, -2 GameEnded (record untouched on failure)."""
        dd = _draws(draws)
        return lib().ao_step(_rec(self.rec), self.players, self.tile_pool, int(action),
                             dd[1] if dd else None, seed, gid)


class factories:
    """``with factories(7): ...`` -- run the oracle with 7 / 9 factory displays (the opt-in "factory count by player
    count" variant); the reference itself always uses 5 (azul.py:19), which is restored on exit."""

    def __init__(self, f):
        self.f = f

    def __enter__(self):
        lib().ao_set_factories(self.f)
        return self

    def __exit__(self, *a):
        lib().ao_set_factories(5)


def legal_mask64(rec, players):
    m = np.zeros(6, dtype=np.uint64)
    lib().ao_legal_mask64(_rec(np.ascontiguousarray(rec, dtype=np.int32)), players, _p(m, ctypes.c_uint64))
    return m


def observe(rec, players, perspective=0):
    """``GameRunner.get_state`` (game_runner.py:56-72) of one record: int32 [32 + 52P]; perspective -1 = the mover."""
    out = np.zeros(32 + 52 * players, dtype=np.int32)
    lib().ao_observe(_rec(np.ascontiguousarray(rec, dtype=np.int32)), players, perspective, _p(out, ctypes.c_int32))
    return out


def runner_continues(rec, players, require_two=True):
    """The ``while`` test of ``GameRunner.step`` (:46, require_two) / ``GameRunner.reset`` (:84) on this state."""
    return bool(lib().ao_runner_continues(_rec(np.ascontiguousarray(rec, dtype=np.int32)), players, int(require_two)))


def statistics(rec, players):
    out = np.zeros(10, dtype=np.int32)
    lib().ao_statistics(_rec(np.ascontiguousarray(rec, dtype=np.int32)), players, _p(out, ctypes.c_int32))
    return out


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().ao_philox4x32_10(_p(c, ctypes.c_uint32), _p(k, ctypes.c_uint32), _p(o, ctypes.c_uint32))
    return o


def random_action(mask6, word):
    m = np.ascontiguousarray(mask6, dtype=np.uint32)
    return lib().ao_random_action(_p(m, ctypes.c_uint32), int(word) & 0xFFFFFFFF)


def fresh_records(n, players, tile_pool, first_rule, seed, gid0=0):
    """n freshly reset games (Philox schedule at total_steps = 0), as records [n, U]."""
    U = lib().ao_record_size(players)
    recs = np.zeros((n, U), dtype=np.int32)
    for i in range(n):
        lib().ao_reset_philox(_rec(recs[i]), players, tile_pool, first_rule, seed, gid0 + i)
    return recs


def rollout_random(recs, players, tile_pool, first_rule, seed, gid0, k_steps, threads=1):
    """In-place K-step random-agent rollout with auto-reset; returns the int64[16] counters."""
    assert recs.dtype == np.int32 and recs.flags.c_contiguous
    cnt = np.zeros(16, dtype=np.int64)
    if threads <= 1:
        lib().ao_rollout_random(_rec(recs), recs.shape[0], players, tile_pool, first_rule, seed, gid0,
                                k_steps, _p(cnt, ctypes.c_int64))
    else:
        lib().ao_rollout_random_mt(_rec(recs), recs.shape[0], players, tile_pool, first_rule, seed, gid0,
                                   k_steps, _p(cnt, ctypes.c_int64), threads)
    return cnt


def opponent_random(rec, players, tile_pool, seed, gid, require_two=True):
    """In place on one record; returns (score diff seat1 - seat2 after a count_score on a copy, next mask)."""
    m = np.zeros(6, dtype=np.uint32)
    d = lib().ao_opponent_random(_rec(rec), players, tile_pool, seed, gid, int(require_two), _p(m, ctypes.c_uint32))
    return int(d), m
