"""Parity of the CUDA path (through the C ABI) with the reference: golden traces recorded from the
unmodified reference, the reference's own known-answer scenarios, and seeded rollouts against the
C oracle.  Needs a B200: run with ``-m gpu``.  Everything is integer work: the bar is bit-exact."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from azul_deep_reinforcement_learning_b200.layout import UnpackedLayout  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.helpers import TRACE_CONFIGS, TraceGame, load_kat, load_trace, stream_digest  # noqa: E402

pytestmark = pytest.mark.gpu


def engine(*a, **k):
    from azul_deep_reinforcement_learning_b200.engine import BatchedAzul
    return BatchedAzul(*a, **k)


def empty_records(n, players, pool, first):
    L = UnpackedLayout(players)
    rec = np.zeros((n, L.size), np.int32)
    rec[:, L.n_players] = players
    rec[:, L.next_first_player] = first
    if pool:
        rec[:, L.box:L.box + 5] = 20
    return rec


@pytest.mark.parametrize("players,rules", TRACE_CONFIGS)
def test_replay_golden_traces(players, rules):
    """All 64 recorded games of a configuration advance in lock-step in one batch."""
    tr = load_trace(players, rules)
    pool = int(tr["tile_pool"])
    n = len(tr["first_player"])
    L = UnpackedLayout(players)
    games = [TraceGame(tr, i) for i in range(n)]
    eng = engine(n, players, pool, 1, reset=False)
    ok = eng.import_records(empty_records(n, players, pool, tr["first_player"].astype(np.int32)))
    assert bool(ok.all())
    eng.new_round(torch.from_numpy(np.stack([g.draws[0] for g in games])))
    rec = eng.export_records().cpu().numpy()
    assert np.array_equal(rec, tr["initial_states"].astype(np.int32))
    T = max(len(g.actions) for g in games)
    rnd = np.ones(n, np.int64)
    masks = [[] for _ in range(n)]
    states = [[] for _ in range(n)]
    for t in range(T):
        act = np.full(n, 255, np.uint8)
        draws = np.full((n, 20), -1, np.int8)
        for i, g in enumerate(games):
            if t < len(g.actions):
                act[i] = g.actions[t]
                if rnd[i] < len(g.draws):
                    draws[i] = g.draws[rnd[i]]
        pre_mask = eng.legal_mask().cpu().numpy().astype(np.uint32)
        turn_before = rec[:, L.turn_counter].copy()
        out = eng.step(torch.from_numpy(act), torch.from_numpy(draws))
        rec = eng.export_records().cpu().numpy()
        status = out["status"].cpu().numpy()
        done = out["done"].cpu().numpy()
        post_mask = out["mask"].cpu().numpy().astype(np.uint32)
        rnd += rec[:, L.turn_counter] != turn_before
        for i, g in enumerate(games):
            if t < len(g.actions):
                assert status[i] == 0, (i, t, status[i])
                masks[i].append(pre_mask[:, i].copy()); states[i].append(rec[i].copy())
                if g.full:
                    assert np.array_equal(pre_mask[:, i], g.masks[t]), (i, t)
                    assert np.array_equal(rec[i], g.states[t + 1]), (i, t)
                    if t + 1 < len(g.actions):
                        assert np.array_equal(post_mask[:, i], g.masks[t + 1]), (i, t)
                assert done[i] == (t + 1 == len(g.actions))
    for i, g in enumerate(games):
        assert np.array_equal(rec[i], g.final), i
        assert stream_digest(masks[i], states[i]) == g.sha, i
        assert rnd[i] == len(g.draws)
    # GameEnded (azul.py:298-299): state untouched, status bit set
    out = eng.step(torch.zeros(n, dtype=torch.uint8), None)
    assert bool((out["status"].cpu() & 2).all())
    assert np.array_equal(eng.export_records().cpu().numpy(), rec)


def test_known_answer_scenarios():
    """tests/test_azul.py:123-331 op sequences: all scenarios run side by side in one batch."""
    kat = load_kat()
    L = UnpackedLayout(2)
    for pool in (0, 1):
        idx = [s for s in range(len(kat["kat_names"])) if int(kat["kat_pool"][s]) == pool]
        n = len(idx)
        rec0 = np.stack([kat["fixture_records"][int(kat["kat_fixture"][s])].astype(np.int32) for s in idx])
        if pool:
            rec0[:, L.box:L.box + 5] = 20
        eng = engine(n, 2, pool, 1, reset=False)
        assert bool(eng.import_records(rec0).all())
        nops = max(int(kat["kat_op_offsets"][s + 1] - kat["kat_op_offsets"][s]) for s in idx)
        for j in range(nops):
            # one op per scenario per round; different op kinds go through different entry points
            sel = {c: np.full(n, 255, np.uint8) for c in (0, 1)}
            which = {c: [] for c in (0, 1, 2, 3, 4)}
            draws = np.full((n, 20), -1, np.int8)
            for i, s in enumerate(idx):
                k = int(kat["kat_op_offsets"][s]) + j
                if k >= int(kat["kat_op_offsets"][s + 1]):
                    continue
                code, a, b, c = [int(x) for x in kat["kat_ops"][k]]
                which[code].append((i, k, a + 6 * b + 30 * c))
                if code in (0, 1):
                    sel[code][i] = a + 6 * b + 30 * c
                    if code == 1 and kat["kat_draws"][k][0] >= 0:
                        draws[i] = kat["kat_draws"][k]
            before = eng.state.clone()
            if which[4]:
                m = eng.legal_mask().cpu().numpy().astype(np.uint32)
                for i, k, a in which[4]:
                    assert int(m[a // 30, i] >> (a % 30) & 1) == int(kat["kat_returns"][k]), (kat["kat_names"][idx[i]], k)
            if which[0]:
                eng.move(torch.from_numpy(sel[0]))
            if which[1]:
                out = eng.step(torch.from_numpy(sel[1]), torch.from_numpy(draws))
                st = out["status"].cpu().numpy()
                for i, k, a in which[1]:
                    want = int(kat["kat_returns"][k])
                    assert (st[i] & 3) == {0: 0, -1: 1, -2: 2}[want], (kat["kat_names"][idx[i]], k)
            for code, fn in ((2, eng.next_player), (3, eng.count_score)):
                if which[code]:
                    # these entry points act on the whole batch: apply, then restore the others
                    keep = torch.ones(n, dtype=torch.bool, device=eng.device)
                    keep[[i for i, _, _ in which[code]]] = False
                    snap = eng.state.clone()
                    fn()
                    eng.state[:, keep] = snap[:, keep]
            rec = eng.export_records().cpu().numpy()
            for code in which:
                for i, k, a in which[code]:
                    assert np.array_equal(rec[i], kat["kat_records"][k].astype(np.int32)), (kat["kat_names"][idx[i]], k)
            del before


def test_fixture_masks_roundtrip_preview_flags():
    kat = load_kat()
    recs = kat["fixture_records"].astype(np.int32)
    n = len(recs)
    eng = engine(n, 2, 0, 1, reset=False)
    assert bool(eng.import_records(recs).all())
    assert np.array_equal(eng.export_records().cpu().numpy(), recs)
    m = eng.legal_mask().cpu().numpy().astype(np.uint32)
    assert np.array_equal(m.T, kat["fixture_masks"])
    prev = eng.score_preview().cpu().numpy()
    flags = eng.round_flags().cpu().numpy()
    for i in range(n):
        g = O.Game(2, 0, record=recs[i].copy())
        assert np.array_equal(prev[:, i], g.score_preview()), kat["fixture_names"][i]
        assert flags[i] == int(g.is_end_of_round()) + 2 * int(g.is_end_of_game())
    assert np.array_equal(eng.export_records().cpu().numpy(), recs)     # preview did not mutate


def test_import_flags_unrepresentable_records():
    L = UnpackedLayout(2)
    rec = empty_records(4, 2, 0, 1)
    rec[1, L.pattern_lines + 10] = 1
    rec[1, L.pattern_lines + 13] = 1            # two colours in one row
    rec[2, L.center + 1] = 16                   # more than a centre field can hold
    eng = engine(4, 2, 0, 1, reset=False)
    ok = eng.import_records(rec).cpu().numpy()
    assert list(ok) == [1, 0, 0, 1]
    out = eng.export_records().cpu().numpy()
    assert out[1, L.status] & 16 and out[2, L.status] & 16 and out[0, L.status] == 0


@pytest.mark.parametrize("players", [2, 3, 4])
@pytest.mark.parametrize("pool", [0, 1])
@pytest.mark.parametrize("first_rule", [0, 1])
def test_rollout_matches_oracle(players, pool, first_rule):
    """Seeded Philox rollouts with auto-reset: records and counters identical to the C oracle."""
    seed, gid0, n, k = 0xABCDEF0123 + players, 777, 2048, 160
    eng = engine(n, players, pool, first_rule, seed=seed, game_id_base=gid0)
    a = O.fresh_records(n, players, pool, first_rule, seed, gid0)
    assert np.array_equal(eng.export_records().cpu().numpy(), a)          # K6 reset parity
    ca = O.rollout_random(a, players, pool, first_rule, seed, gid0, k, threads=8)
    mask = torch.empty((6, n), dtype=torch.int32, device=eng.device)
    eng.rollout_random(k, mask)
    b = eng.export_records().cpu().numpy()
    assert np.array_equal(a, b), np.nonzero((a != b).any(axis=1))[0][:10]
    cb = eng.counters.cpu().numpy()
    assert np.array_equal(ca, cb), (ca, cb)
    # the mask written by the fused kernel is the mask of the final state
    assert torch.equal(mask, eng.legal_mask())
    want = np.stack([O.Game(players, pool, record=a[i]).legal_mask() for i in range(64)])
    assert np.array_equal(mask[:, :64].cpu().numpy().astype(np.uint32).T, want)


def test_step_philox_matches_oracle_and_illegal_leaves_state():
    """azb_step with the Philox schedule (draws20 = NULL) against ao_step; illegal actions are flagged."""
    players, pool, seed, n = 2, 1, 99, 512
    eng = engine(n, players, pool, 0, seed=seed)
    recs = eng.export_records().cpu().numpy()
    rng = np.random.default_rng(5)
    for t in range(80):
        masks = eng.legal_mask().cpu().numpy().astype(np.uint32)
        act = np.zeros(n, np.uint8)
        expect = np.zeros(n, np.int64)
        for i in range(n):
            g = O.Game(players, pool, record=recs[i])
            if rng.random() < 0.15:
                a = int(rng.integers(0, 200))          # sometimes illegal / out of range
            else:
                a = O.random_action(masks[:, i], int(rng.integers(0, 2 ** 32)))
                a = 0 if a < 0 else a
            act[i] = a
            expect[i] = g.step(a, None, seed, i)
            recs[i] = g.rec
        out = eng.step(torch.from_numpy(act), None)
        got = eng.export_records().cpu().numpy()
        st = out["status"].cpu().numpy()
        assert np.array_equal(got, recs), t
        assert np.array_equal(st & 3, np.where(expect == 0, 0, np.where(expect == -1, 1, 2))), t


def _lowest_legal_action(mask):
    """One legal action per game from the mask words (torch, on device)."""
    m = mask.to(torch.int64) & 0xFFFFFFFF
    word = (m != 0).to(torch.int64).argmax(dim=0)
    w = m.gather(0, word[None, :]).squeeze(0)
    low = torch.zeros_like(w)
    for b in range(30):
        low = torch.where(((w >> b) & 1).bool() & ((w & ((1 << b) - 1)) == 0), torch.full_like(w, b), low)
    return (30 * word + low).to(torch.uint8)


@pytest.mark.parametrize("players,pool,n", [(2, 1, 300001), (2, 0, 299972), (4, 1, 200000), (3, 1, 150002)])
def test_step_many_rows_per_warp(players, pool, n):
    """azb_step on batches where every persistent warp walks many rows (queue drains in the middle of the walk,
    staged row tiles are recycled), with ragged / 16-byte-unaligned sizes, skipped, illegal and ended games:
    identical to the same games stepped in small separate batches, and to the oracle on a sample."""
    seed = 77
    eng = engine(n, players, pool, 0, seed=seed)
    eng.rollout_random(1 + (n % 13))
    rng = torch.Generator(device="cpu").manual_seed(n)
    for it in range(3):
        action = _lowest_legal_action(eng.legal_mask())
        r = torch.rand(n, generator=rng).to(eng.device)
        action = torch.where(r < 0.05, torch.full_like(action, 255), action)            # skipped slots
        action = torch.where((r >= 0.05) & (r < 0.08), torch.full_like(action, 199), action)   # out of range -> illegal
        before = eng.state.clone()
        out = eng.step(action, None, want_preview=(it == 1))
        after = eng.state
        for lo, m in ((0, 1000), (n - 999, 999), (n // 2 - 7, 4096)):
            sub = engine(m, players, pool, 0, seed=seed, game_id_base=lo, reset=False)
            sub.state.copy_(before[:, lo:lo + m])
            o2 = sub.step(action[lo:lo + m], None, want_preview=(it == 1))
            assert torch.equal(sub.state, after[:, lo:lo + m]), (it, lo)
            for key in o2:
                assert torch.equal(o2[key], out[key][..., lo:lo + m]), (it, lo, key)
            if lo == 0 and it < 2:
                recs = engine_records(before[:, :m], players, pool, seed)
                got = sub.export_records().cpu().numpy()
                st = o2["status"].cpu().numpy()
                act = action[:m].cpu().numpy()
                for i in range(m):
                    if act[i] == 255:
                        assert np.array_equal(got[i], recs[i]) and st[i] & 3 == 0
                        continue
                    g = O.Game(players, pool, record=recs[i])
                    code = g.step(int(act[i]), None, seed, i)
                    assert np.array_equal(got[i], g.rec), (it, i)
                    assert (st[i] & 3) == (0 if code == 0 else 1 if code == -1 else 2), (it, i)
        assert int((out["status"] & 1).sum()) > 0


def engine_records(state, players, pool, seed):
    e = engine(state.shape[1], players, pool, 0, seed=seed, reset=False)
    e.state.copy_(state)
    return e.export_records().cpu().numpy()


def test_observe_matches_reference_layout():
    """GameRunner.get_state (game_runner.py:56-72) rebuilt with numpy from exported records."""
    for players in (2, 3, 4):
        n = 256
        eng = engine(n, players, 1, 0, seed=3)
        eng.rollout_random(37)
        rec = eng.export_records().cpu().numpy()
        L = UnpackedLayout(players)
        for persp in list(range(players)) + [-1]:
            obs = eng.observe(persp).cpu().numpy()
            assert obs.shape == (n, 32 + 52 * players)
            for i in range(0, n, 17):
                r = rec[i]
                p0 = persp if persp >= 0 else (int(r[L.current_player]) - 1) % players
                order = [p0] + [q for q in range(players) if q != p0]
                pat = r[L.pattern_lines:L.pattern_lines + 25 * players].reshape(players, 25)
                wal = r[L.walls:L.walls + 25 * players].reshape(players, 25)
                nf = int(r[L.next_first_player])
                pn = ((nf - 1 - p0) % players) + 1 if nf > 0 else 0
                want = np.concatenate([r[0:25], r[25:31], pat[order].ravel(), wal[order].ravel(),
                                       r[L.floors:L.floors + players][order], r[L.score:L.score + players][order], [pn]])
                assert np.array_equal(obs[i], want.astype(np.float32)), (players, persp, i)
    # tests/test_game_runner.py:71-75: a fresh 2-player game observes 136 values summing to 21
    eng = engine(8, 2, 1, 0, seed=1)
    obs = eng.observe(0)
    assert obs.shape[1] == 136 and bool((obs.sum(dim=1) == 21).all())
    # the bfloat16 variant is the float32 observation rounded once (ragged batch: the last block is partial)
    for players in (2, 3, 4):
        eng = engine(1000 + players, players, 1, 0, seed=9)
        eng.rollout_random(37)
        for persp in (-1, 0, players - 1):
            assert torch.equal(eng.observe_bf16(persp), eng.observe(persp).to(torch.bfloat16)), (players, persp)


def test_stats_match_records():
    eng = engine(512, 2, 0, 1, seed=11)
    eng.rollout_random(45)
    rec = eng.export_records().cpu().numpy()
    st = eng.stats().cpu().numpy()
    L = UnpackedLayout(2)
    assert np.array_equal(st[:, 0], rec[:, L.score]) and np.array_equal(st[:, 1], rec[:, L.score + 1])
    assert np.array_equal(st[:, 2], rec[:, L.turn_counter])
    assert np.array_equal(st[:, 3], rec[:, L.first_player_stats])
    assert np.array_equal(st[:, 4], rec[:, L.first_player_stats:L.first_player_stats + 2].sum(axis=1))
    assert np.array_equal(st[:, 5], -rec[:, L.floor_penalty])
    assert np.array_equal(st[:, 6], rec[:, L.max_combo])
    assert np.array_equal(st[:, 7], rec[:, L.completed_lines + 0])
    assert np.array_equal(st[:, 8], rec[:, L.completed_lines + 2])
    assert np.array_equal(st[:, 9], rec[:, L.completed_lines + 1])


def test_full_size_properties():
    """BASELINE config 2 size (65,536 games): size-independent properties of the rollout."""
    n, k = 65536, 512
    for pool in (0, 1):
        eng = engine(n, 2, pool, 0, seed=0x5EED)
        eng.rollout_random(k)
        one = eng.state.clone()
        c1 = eng.counters.clone()
        # determinism + restartability: two half-length launches == one launch
        eng2 = engine(n, 2, pool, 0, seed=0x5EED)
        eng2.rollout_random(200); eng2.rollout_random(k - 200)
        assert torch.equal(one, eng2.state) and torch.equal(c1, eng2.counters)
        # sharding independence: the same id range split over two handles (multi-GPU layout)
        h = n // 2
        lo = engine(h, 2, pool, 0, seed=0x5EED, game_id_base=0)
        hi = engine(h, 2, pool, 0, seed=0x5EED, game_id_base=h)
        lo.rollout_random(k); hi.rollout_random(k)
        assert torch.equal(one[:, :h], lo.state) and torch.equal(one[:, h:], hi.state)
        assert torch.equal(c1, lo.counters + hi.counters)
        cnt = eng.read_counters()
        assert cnt["steps"] == n * k and cnt["games"] > 0 and cnt["stuck"] == 0
        assert 45 < cnt["steps"] / cnt["games"] < 75          # ~57 steps per 2-player game
        rec = eng.export_records().cpu().numpy()
        L = UnpackedLayout(2)
        assert (rec[:, L.status] == 0).all() and (rec[:, L.end_of_game] == 0).all()
        tiles_on_table = rec[:, 0:30].sum(axis=1)
        assert (tiles_on_table > 0).all()                     # auto-reset leaves every slot playable
        assert (rec[:, L.floors:L.floors + 2] <= 7).all() and (rec[:, L.score:L.score + 2] >= 0).all()
        if pool == 1:
            # tile conservation: box + lid + table + pattern lines + walls == 100
            total = (rec[:, L.box:L.box + 10].sum(axis=1) + tiles_on_table +
                     rec[:, L.pattern_lines:L.pattern_lines + 50].sum(axis=1) + rec[:, L.walls:L.walls + 50].sum(axis=1))
            # floor tiles are already in the lid (azul.py:155-161)
            assert (total == 100).all()
        # every legal-mask bit implies its source holds that colour
        m = eng.legal_mask().cpu().numpy().astype(np.uint32)
        src = m[0]
        assert ((m[1:] & ~src) == 0).all() and (src != 0).all()


def test_rollout_independent_of_launch_shape():
    """Ragged batch (not a multiple of the warp size), every block size and end-of-round batching
    threshold give the oracle's result: the warp-level scheduling never changes a game."""
    players, pool, seed, n, k = 3, 1, 4242, 1000, 120
    ref = O.fresh_records(n, players, pool, 0, seed, 5)
    cref = O.rollout_random(ref, players, pool, 0, seed, 5, k, threads=8)
    for block, defer in ((128, 1), (32, 7), (64, 16), (256, 32), (96, 24)):
        eng = engine(n, players, pool, 0, seed=seed, game_id_base=5)
        eng.set_block_threads(block)
        eng.set_rollout_defer(defer)
        eng.rollout_random(k)
        assert np.array_equal(eng.export_records().cpu().numpy(), ref), (block, defer)
        assert np.array_equal(eng.counters.cpu().numpy(), cref), (block, defer)


def test_maximum_batch_indexing():
    """2^27 games (9.1 GB of packed state, word offsets beyond 2^31): reset, a few fused steps, spot checks
    against the oracle at both ends of the id range."""
    n = 1 << 27
    free, _ = torch.cuda.mem_get_info()
    if free < 12 * (1 << 30):
        pytest.skip("needs 12 GB of free device memory")
    seed, k = 321, 3
    eng = engine(n, 2, 1, 0, seed=seed)
    eng.rollout_random(k)
    c = eng.read_counters()
    assert c["steps"] == n * k
    L = UnpackedLayout(2)
    for lo in (0, n - 64):
        ref = O.fresh_records(64, 2, 1, 0, seed, lo)
        O.rollout_random(ref, 2, 1, 0, seed, lo, k)
        sub = engine(64, 2, 1, 0, seed=seed, game_id_base=lo, reset=False)
        sub.state.copy_(eng.state[:, lo:lo + 64])
        assert np.array_equal(sub.export_records().cpu().numpy(), ref), lo
    del eng
    torch.cuda.empty_cache()


@pytest.mark.parametrize("players", [3, 4])
def test_full_size_three_and_four_players(players):
    """BASELINE config 3 batch (262,144 games per GPU) for the 3- and 4-player variants: conservation, counters."""
    n, k = 262144, 300
    eng = engine(n, players, 1, 0, seed=9)
    eng.rollout_random(k)
    c = eng.read_counters()
    assert c["steps"] == n * k and c["games"] > 0 and c["stuck"] == 0
    per_game = c["steps"] / c["games"]
    assert (55 < per_game < 80) if players == 3 else (60 < per_game < 90)      # ~63 / ~71 steps per game
    rec = eng.export_records().cpu().numpy()
    L = UnpackedLayout(players)
    total = (rec[:, L.box:L.box + 10].sum(axis=1) + rec[:, 0:30].sum(axis=1) +
             rec[:, L.pattern_lines:L.pattern_lines + 25 * players].sum(axis=1) +
             rec[:, L.walls:L.walls + 25 * players].sum(axis=1))
    ok = (rec[:, L.status] & 8) == 0
    assert (total[ok] == 100).all() and ok.mean() > 0.99
    assert ((rec[:, L.current_player] >= 1) & (rec[:, L.current_player] <= players)).all()


def test_replay_cli_and_json_io(tmp_path):
    """The trace replay tool and the batch JSON import/export (azul.py:90-117 schema) of the product package."""
    import json
    from azul_deep_reinforcement_learning_b200 import io as azio
    from azul_deep_reinforcement_learning_b200.replay import replay_trace
    tr = load_trace(2, "lid")
    rep = replay_trace(tr)
    assert rep["ok"] and rep["games"] == 64, rep["mismatches"][:3]
    tr["actions"] = tr["actions"].copy()
    tr["actions"][5] = (int(tr["actions"][5]) + 1) % 180          # a corrupted recording must be detected
    assert not replay_trace(tr)["ok"]
    kat = load_kat()
    recs = kat["fixture_records"].astype(np.int32)
    paths = []
    for i, r in enumerate(recs):
        p = tmp_path / ("b%d.json" % i)
        p.write_text(json.dumps(azio.json_dict_from_record(r)))
        paths.append(str(p))
    eng = engine(len(recs), 2, 0, 1, reset=False)
    assert bool(azio.import_json_files(eng, paths).all())
    L = UnpackedLayout(2)
    got = eng.export_records().cpu().numpy()
    assert np.array_equal(got[:, :L.end_of_game], recs[:, :L.end_of_game])      # the 10 JSON keys round-trip
    out = [str(tmp_path / ("o%d.json" % i)) for i in range(len(recs))]
    azio.export_json_files(eng, out)
    for a, b in zip(paths, out):
        assert json.load(open(a)) == json.load(open(b))


def test_random_boards_against_reference():
    """tests/golden/fuzz.npz through the C ABI: mask, count_score and one step of 250 random boards per configuration."""
    from tests.helpers import FUZZ_CONFIGS, fuzz_key, load_fuzz
    fz = load_fuzz()
    for players, pool in FUZZ_CONFIGS:
        k = fuzz_key(players, pool)
        before = fz[k + "_before"].astype(np.int32)
        n = len(before)
        eng = engine(n, players, pool, 1, reset=False)
        assert bool(eng.import_records(before).all())
        assert np.array_equal(eng.export_records().cpu().numpy(), before)
        assert np.array_equal(eng.legal_mask().cpu().numpy().astype(np.uint32).T, fz[k + "_mask"]), k
        prev = eng.score_preview().cpu().numpy().T
        L = UnpackedLayout(players)
        assert np.array_equal(prev, fz[k + "_scored"][:, L.score:L.score + players]), k
        snap = eng.state.clone()
        eng.count_score()
        assert np.array_equal(eng.export_records().cpu().numpy(), fz[k + "_scored"].astype(np.int32)), k
        eng.state.copy_(snap)
        out = eng.step(torch.from_numpy(fz[k + "_action"]), torch.from_numpy(fz[k + "_draws"]))
        assert int((out["status"] & 3).max()) == 0
        got = eng.export_records().cpu().numpy()
        want = fz[k + "_stepped"].astype(np.int32)
        assert np.array_equal(got, want), (k, np.nonzero((got != want).any(axis=1))[0][:5])
