"""Board-state interchange at batch level: the reference's JSON schema (``Azul.export_JSON`` /
``import_JSON``, azul.py:90-117) <-> unpacked int32 records <-> packed device state (kernel K7)."""
import json

import numpy as np

from .layout import UnpackedLayout


def record_from_json_dict(data, tile_pool=0):
    """One dict in the reference's 10-key schema -> unpacked record (box = 20 per colour for the Lid pool,
    like ``Azul(rules={"tile_pool": "Lid"})`` followed by ``import_JSON``)."""
    P = int(data["players"])
    L = UnpackedLayout(P)
    r = np.zeros(L.size, dtype=np.int32)
    r[L.displays:L.displays + 25] = np.asarray(data["game_board_displays"], dtype=np.int64).reshape(-1)
    r[L.center:L.center + 6] = np.asarray(data["game_board_center"], dtype=np.int64)
    r[L.pattern_lines:L.pattern_lines + 25 * P] = np.asarray(data["pattern_lines"], dtype=np.int64).reshape(-1)
    r[L.walls:L.walls + 25 * P] = np.asarray(data["walls"], dtype=np.int64).reshape(-1) != 0
    r[L.floors:L.floors + P] = np.asarray(data["floors"], dtype=np.int64)
    r[L.score:L.score + P] = np.asarray(data["score"], dtype=np.int64)
    r[L.current_player] = data["current_player"]
    r[L.next_first_player] = data["next_first_player"]
    r[L.n_players] = P
    r[L.turn_counter] = data["turn_counter"]
    if tile_pool:
        r[L.box:L.box + 5] = 20
    return r


def json_dict_from_record(rec):
    P = int(rec[UnpackedLayout(2).n_players]) if len(rec) == UnpackedLayout(2).size else (len(rec) - 48) // 58
    L = UnpackedLayout(P)
    r = np.asarray(rec, dtype=np.int64)
    return {
        "game_board_displays": r[L.displays:L.displays + 25].reshape(5, 5).tolist(),
        "game_board_center": r[L.center:L.center + 6].tolist(),
        "pattern_lines": r[L.pattern_lines:L.pattern_lines + 25 * P].reshape(P, 5, 5).tolist(),
        "walls": r[L.walls:L.walls + 25 * P].reshape(P, 5, 5).tolist(),
        "floors": r[L.floors:L.floors + P].tolist(), "score": r[L.score:L.score + P].tolist(),
        "current_player": int(r[L.current_player]), "next_first_player": int(r[L.next_first_player]),
        "players": P, "turn_counter": int(r[L.turn_counter]),
    }


def import_json_files(engine, paths):
    """Load one JSON board per game slot into ``engine`` (len(paths) == engine.n_games); returns the ok flags."""
    recs = []
    for p in paths:
        with open(p) as f:
            recs.append(record_from_json_dict(json.load(f), engine.tile_pool))
    return engine.import_records(np.stack(recs))


def export_json_files(engine, paths):
    recs = engine.export_records().cpu().numpy()
    for rec, p in zip(recs, paths):
        with open(p, "w") as f:
            json.dump(json_dict_from_record(rec), f)


def records_equal(a, b):
    """``Azul.__eq__`` (azul.py:63) on unpacked records: displays, centre, pattern lines, walls, floors, score,
    current / next first player, players, end_of_game, turn_counter -- box / lid and the statistics are not compared.
    ``a`` and ``b``: int arrays ``[..., U]`` of the same player count; returns a bool array over the leading dims."""
    a, b = np.asarray(a), np.asarray(b)
    players = (a.shape[-1] - 48) // 58
    end = UnpackedLayout(players).box            # everything before box/lid is exactly the __eq__ field set
    return (a[..., :end] == b[..., :end]).all(axis=-1)
