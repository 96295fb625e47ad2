"""Replay recorded games (tile draws + actions) through the CUDA engine and compare with the recording.

    python -m azul_deep_reinforcement_learning_b200.replay tests/golden/trace_p2_lid.npz [...]

Trace files are the ``.npz`` written by ``oracle/record_golden.py`` from the unmodified reference (format in its
docstring): per game the initial ``next_first_player``, the 20 draws of every ``new_round`` and every action; for
the first ``n_full`` games also each pre-step legal mask and post-step unpacked record, for all games the final
record and a SHA-256 of the (mask, record) stream.  All games of a file advance in lock-step in one batch.
"""
import hashlib
import sys

import numpy as np
import torch

from .engine import BatchedAzul
from .layout import UnpackedLayout


def replay_trace(tr, device=0):
    """``tr``: dict of arrays from a trace ``.npz``.  Returns a report dict; ``report["ok"]`` is the verdict."""
    P, pool = int(tr["players"]), int(tr["tile_pool"])
    n, n_full = len(tr["first_player"]), int(tr["n_full"])
    L = UnpackedLayout(P)
    so, ro = tr["step_offsets"], tr["round_offsets"]
    eng = BatchedAzul(n, P, pool, 1, device=device, reset=False)
    rec0 = np.zeros((n, L.size), np.int32)
    rec0[:, L.n_players] = P
    rec0[:, L.next_first_player] = tr["first_player"]
    if pool:
        rec0[:, L.box:L.box + 5] = 20
    assert bool(eng.import_records(rec0).all())
    eng.new_round(torch.from_numpy(np.stack([tr["draws"][ro[i]] for i in range(n)])))
    rec = eng.export_records().cpu().numpy()
    report = {"games": n, "steps": int(so[-1]), "mismatches": []}
    if not np.array_equal(rec, tr["initial_states"].astype(np.int32)):
        report["mismatches"].append("initial states")
    T = int(np.max(np.diff(so)))
    rnd = np.ones(n, np.int64)
    hashes = [hashlib.sha256() for _ in range(n)]
    for t in range(T):
        act = np.full(n, 255, np.uint8)
        draws = np.full((n, 20), -1, np.int8)
        for i in range(n):
            if so[i] + t < so[i + 1]:
                act[i] = tr["actions"][so[i] + t]
                if ro[i] + rnd[i] < ro[i + 1]:
                    draws[i] = tr["draws"][ro[i] + rnd[i]]
        mask = eng.legal_mask().cpu().numpy().astype(np.uint32)
        turn = rec[:, L.turn_counter].copy()
        out = eng.step(torch.from_numpy(act), torch.from_numpy(draws), want_mask=False)
        rec = eng.export_records().cpu().numpy()
        st = out["status"].cpu().numpy()
        rnd += rec[:, L.turn_counter] != turn
        for i in range(n):
            if so[i] + t >= so[i + 1]:
                continue
            if st[i] & 3:
                report["mismatches"].append("game %d step %d: status %d" % (i, t, st[i]))
            hashes[i].update(mask[:, i].astype("<u4").tobytes())
            hashes[i].update(rec[i].astype("<i4").tobytes())
            if i < n_full:
                if not np.array_equal(mask[:, i], tr["full_masks"][so[i] + t]):
                    report["mismatches"].append("game %d step %d: legal mask" % (i, t))
                if not np.array_equal(rec[i], tr["full_states"][so[i] + i + t + 1].astype(np.int32)):
                    report["mismatches"].append("game %d step %d: state" % (i, t))
    if not np.array_equal(rec, tr["final_states"].astype(np.int32)):
        report["mismatches"].append("final states")
    for i in range(n):
        if hashes[i].digest() != bytes(tr["stream_sha256"][i].tobytes()):
            report["mismatches"].append("game %d: stream digest" % i)
    report["ok"] = not report["mismatches"]
    return report


def main(argv=None):
    paths = (argv if argv is not None else sys.argv[1:])
    if not paths:
        print(__doc__)
        return 2
    bad = 0
    for p in paths:
        z = np.load(p)
        rep = replay_trace({k: z[k] for k in z.files})
        print("%s: %d games, %d steps: %s" % (p, rep["games"], rep["steps"], "bit-exact" if rep["ok"] else
                                              "MISMATCH (%s ...)" % "; ".join(rep["mismatches"][:3])))
        bad += not rep["ok"]
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
