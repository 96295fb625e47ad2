"""Shared helpers for the parity tests (golden-trace access; no game logic)."""
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRACE_CONFIGS = [(p, r) for p in (2, 3, 4) for r in ("default", "lid")]


def load_trace(players, rules):
    z = np.load(os.path.join(GOLDEN, "trace_p%d_%s.npz" % (players, rules)))
    return {k: z[k] for k in z.files}


def load_kat():
    z = np.load(os.path.join(GOLDEN, "kat.npz"))
    return {k: z[k] for k in z.files}


class TraceGame:
    """View of game ``i`` inside a loaded trace."""

    def __init__(self, tr, i):
        s0, s1 = int(tr["step_offsets"][i]), int(tr["step_offsets"][i + 1])
        r0, r1 = int(tr["round_offsets"][i]), int(tr["round_offsets"][i + 1])
        self.actions = tr["actions"][s0:s1]
        self.draws = tr["draws"][r0:r1]                # [rounds, 20], round 0 = the initial new_round
        self.first_player = int(tr["first_player"][i])
        self.initial = tr["initial_states"][i].astype(np.int32)
        self.final = tr["final_states"][i].astype(np.int32)
        self.sha = bytes(tr["stream_sha256"][i].tobytes())
        self.full = i < int(tr["n_full"])
        if self.full:
            # full_states rows: game j contributes T_j + 1 rows
            off = s0 + i
            self.states = tr["full_states"][off:off + (s1 - s0) + 1].astype(np.int32)
            self.masks = tr["full_masks"][s0:s1]


def stream_digest(masks, states):
    h = hashlib.sha256()
    for m, s in zip(masks, states):
        h.update(np.asarray(m, dtype="<u4").tobytes())
        h.update(np.asarray(s, dtype="<i4").tobytes())
    return h.digest()


def load_fuzz():
    z = np.load(os.path.join(GOLDEN, "fuzz.npz"))
    return {k: z[k] for k in z.files}


FUZZ_CONFIGS = [(p, pool) for p in (2, 3, 4) for pool in (0, 1)]


def fuzz_key(players, pool):
    return "p%d_%s" % (players, "lid" if pool else "default")
