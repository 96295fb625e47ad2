"""The opt-in rule variant "factory count by player count" (7 / 9 factory displays for 3 / 4 players, 240 / 300 actions)
on the GPU through the C ABI (azb_v_*): against the C oracle run with the same number of displays, and -- with five
displays -- against the default engine, which is pinned to the reference."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu

VARIANTS = [(3, 7), (4, 9)]


@pytest.mark.parametrize("players,factories", VARIANTS)
@pytest.mark.parametrize("pool,first_rule", [(0, 1), (1, 0)])
def test_variant_rollout_matches_oracle(players, factories, pool, first_rule):
    from azul_deep_reinforcement_learning_b200.engine import BatchedAzulByPlayers
    n, k, seed, gid0 = 3000, 260, 0xFACE + factories, 50
    eng = BatchedAzulByPlayers(n, players, pool, first_rule, seed=seed, game_id_base=gid0)
    assert eng.factories == factories and eng.n_actions == 30 * (factories + 1)
    with O.factories(factories):
        recs = O.fresh_records(n, players, pool, first_rule, seed, gid0)
        assert np.array_equal(eng.export_records().cpu().numpy(), recs)
        assert (recs[:, :5 * factories].sum(1) == 4 * factories).all()
        eng.rollout_random(k)
        cnt = O.rollout_random(recs, players, pool, first_rule, seed, gid0, k, threads=4)
        assert np.array_equal(eng.export_records().cpu().numpy(), recs)
        assert np.array_equal(eng.counters.cpu().numpy(), cnt) and cnt[1] > 0
        mask = eng.legal_mask().cpu().numpy().astype(np.uint64)
        for i in range(0, n, 211):
            assert np.array_equal(mask[:, i], O.legal_mask64(recs[i], players))
    # a second launch continues from the same states (split invariance)
    twin = BatchedAzulByPlayers(n, players, pool, first_rule, seed=seed, game_id_base=gid0)
    twin.rollout_random(100)
    twin.rollout_random(k - 100)
    assert torch.equal(twin.state, eng.state) and torch.equal(twin.counters, eng.counters)


@pytest.mark.parametrize("players", [2, 3, 4])
@pytest.mark.parametrize("pool", [0, 1])
def test_variant_kernels_with_five_displays_equal_the_default_engine(players, pool):
    """factories = 5: the variant's kernels reproduce the default (reference-pinned) engine: records, counters, masks."""
    from azul_deep_reinforcement_learning_b200.engine import BatchedAzul, BatchedAzulByPlayers
    n, k = 4096, 300
    a = BatchedAzul(n, players, pool, 0, seed=41)
    b = BatchedAzulByPlayers(n, players, pool, 0, seed=41, factories=5)
    assert torch.equal(a.export_records(), b.export_records())
    a.rollout_random(k)
    b.rollout_random(k)
    assert torch.equal(a.export_records(), b.export_records())
    assert torch.equal(a.counters, b.counters) and int(a.counters[1]) > 0
    assert torch.equal(a.legal_mask().to(torch.int64) & 0xFFFFFFFF, b.legal_mask())


@pytest.mark.parametrize("players,factories", VARIANTS)
def test_variant_step_statuses_and_record_round_trip(players, factories):
    from azul_deep_reinforcement_learning_b200.engine import BatchedAzulByPlayers
    n = 512
    S = factories + 1
    eng = BatchedAzulByPlayers(n, players, 1, 0, seed=3)
    eng.rollout_random(17)
    rec = eng.export_records().cpu().numpy()
    twin = BatchedAzulByPlayers(n, players, 1, 0, seed=3, reset=False)
    assert bool(twin.import_records(rec).all()) and torch.equal(twin.state, eng.state)
    mask = eng.legal_mask().cpu().numpy().astype(np.uint64)
    rng = np.random.default_rng(1)
    act = np.zeros(n, np.int64)
    kind = rng.integers(0, 3, size=n)                 # 0 legal, 1 illegal, 2 skip
    with O.factories(factories):
        want = rec.copy()
        for i in range(n):
            bits = [(p * 5 * S + b) for p in range(6) for b in range(5 * S) if (int(mask[p, i]) >> b) & 1]
            if kind[i] == 0:
                act[i] = rng.choice(bits)
                g = O.Game(players, 1, record=want[i])
                assert g.step(int(act[i]), None, 3, i) == 0
                want[i] = g.rec
            elif kind[i] == 1:
                illegal = sorted(set(range(30 * S)) - set(bits))
                act[i] = rng.choice(illegal) if illegal else 30 * S + 5
            else:
                act[i] = 0xFFFF
        out = eng.step(torch.from_numpy(act))
        st = out["status"].cpu().numpy()
        assert ((st[kind == 0] & 3) == 0).all() and ((st[kind == 1] & 1) == 1).all() and ((st[kind == 2] & 3) == 0).all()
        got = eng.export_records().cpu().numpy()
        assert np.array_equal(got, want)
        m2 = out["mask"].cpu().numpy().astype(np.uint64)
        for i in range(0, n, 37):
            assert np.array_equal(m2[:, i], O.legal_mask64(want[i], players))
