"""The ``azulnet`` drop-in façade on the CUDA path: the expectations of the reference's own test-suite
(tests/test_azul.py, test_game_runner.py, test_model.py, test_nn_runner.py, test_random_agent.py),
restated against ``azul_deep_reinforcement_learning_b200.azulnet``.  Board fixtures come from
tests/golden/kat.npz (the reference's tests/resources/*.json as unpacked records)."""
import copy
import json
import random

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from azul_deep_reinforcement_learning_b200.layout import UnpackedLayout  # noqa: E402
from tests.helpers import load_kat  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def az():
    import azul_deep_reinforcement_learning_b200.azulnet as m
    return m


@pytest.fixture(scope="module")
def res(tmp_path_factory):
    """name -> path of a JSON board file in the reference's export_JSON schema (azul.py:105-116)."""
    kat = load_kat()
    d = tmp_path_factory.mktemp("resources")
    L = UnpackedLayout(2)
    out = {}
    for name, rec in zip(kat["fixture_names"], kat["fixture_records"].astype(int)):
        data = {
            "game_board_displays": rec[0:25].reshape(5, 5).tolist(), "game_board_center": rec[25:31].tolist(),
            "pattern_lines": rec[L.pattern_lines:L.pattern_lines + 50].reshape(2, 5, 5).tolist(),
            "walls": rec[L.walls:L.walls + 50].reshape(2, 5, 5).tolist(),
            "floors": rec[L.floors:L.floors + 2].tolist(), "score": rec[L.score:L.score + 2].tolist(),
            "current_player": int(rec[L.current_player]), "next_first_player": int(rec[L.next_first_player]),
            "players": 2, "turn_counter": int(rec[L.turn_counter])}
        p = d / (str(name) + ".json")
        p.write_text(json.dumps(data))
        out[str(name)] = str(p)
    return out


# ---- tests/test_azul.py ------------------------------------------------------------------------
def test_azul_init(az, res):
    Azul = az.Azul
    assert Azul().game_board_displays.shape == (5, 5)
    assert Azul().game_board_center.shape == (6,)
    for i in range(2, 5):
        g = Azul(players=i)
        assert g.pattern_lines.shape == (i, 5, 5) and np.count_nonzero(g.pattern_lines) == 0
        assert g.walls.shape == (i, 5, 5) and np.count_nonzero(g.walls) == 0
        assert g.floors.shape == (i,) and g.score.shape == (i,) and np.count_nonzero(g.score) == 0
    assert Azul().turn_counter == 0
    random.seed(1)
    game = Azul()
    game.new_round()
    assert game == Azul(state_file=res["game_first_round_seed_1"])       # seed-1 KAT, tests/test_azul.py:36-39
    for fp in (1, 2):
        game = Azul(rules={"first_player": fp})
        game.new_round()
        assert game.current_player == fp
    with pytest.raises(az.IllegalRule):
        Azul(rules={"first_player": 3})
    with pytest.raises(az.IllegalRule):
        Azul(rules={"tile_pool": "Bag"})
    firsts = []
    for _ in range(60):
        game = Azul(rules={"first_player": "Random"})
        game.new_round()
        firsts.append(game.current_player)
    assert set(firsts) == {1, 2}


def test_azul_new_round_eq_json(az, res, tmp_path):
    Azul = az.Azul
    game = Azul()
    prev = game.next_first_player
    game.new_round()
    assert all(np.sum(t) == 4 for t in game.game_board_displays)
    assert np.count_nonzero(game.game_board_center[:5]) == 0 and game.game_board_center[5] == 1
    assert game.current_player == prev
    game.next_first_player = 1
    game.new_round()
    assert game.next_first_player == 0 and game.turn_counter == 2
    g1, g2 = Azul(), Azul()
    assert g1 == g2
    g1.new_round(); g2.new_round()
    assert g1 != g2
    imported = Azul()
    imported.import_JSON(res["game_empty"])
    assert Azul() == imported
    game = Azul(); game.new_round()
    game.export_JSON(tmp_path / "x.json")
    other = Azul(); other.import_JSON(tmp_path / "x.json")
    assert game == other
    game.import_JSON(res["game_sample_1"]); game.export_JSON(tmp_path / "x.json"); other.import_JSON(tmp_path / "x.json")
    assert game == other


def test_azul_move(az, res):
    game = az.Azul()
    fr = res["game_first_round"]
    game.import_JSON(fr); game.move(5, 0, 2)
    assert np.array_equal(game.game_board_displays[4], np.zeros(5))
    assert np.array_equal(game.game_board_center, [0, 1, 2, 0, 0, 1])
    assert np.array_equal(game.pattern_lines[game.current_player - 1, 1], [1, 0, 0, 0, 0])
    game.import_JSON(fr); game.move(2, 3, 4)
    assert np.array_equal(game.game_board_displays[1], np.zeros(5))
    assert np.array_equal(game.game_board_center, [0, 0, 0, 0, 0, 1])
    assert np.array_equal(game.pattern_lines[game.current_player - 1, 3], [0, 0, 0, 4, 0])
    game.import_JSON(fr); game.move(2, 3, 2)
    assert np.array_equal(game.pattern_lines[game.current_player - 1, 1], [0, 0, 0, 2, 0])
    assert game.floors[game.current_player - 1] == 2
    game.import_JSON(fr); game.move(1, 0, 2); game.move(0, 1, 1)
    assert np.array_equal(game.game_board_center, [0, 0, 1, 0, 0, 0])
    assert np.array_equal(game.pattern_lines[game.current_player - 1, 0], [0, 1, 0, 0, 0])
    assert game.next_first_player == game.current_player and game.floors[game.current_player - 1] == 1
    game.import_JSON(fr); game.move(1, 0, 3); game.move(3, 0, 3)
    assert np.array_equal(game.game_board_center, [0, 2, 1, 1, 0, 1])
    assert np.array_equal(game.pattern_lines[game.current_player - 1, 2], [3, 0, 0, 0, 0])
    assert game.floors[game.current_player - 1] == 1
    game.import_JSON(fr); game.move(3, 0, 0)
    assert game.floors[game.current_player - 1] == 2
    game.move(4, 0, 0)
    assert game.floors[game.current_player - 1] == 3
    game.move(1, 0, 0); game.move(2, 3, 1)
    assert game.floors[game.current_player - 1] == 7


def test_azul_legal_next_player_round_game_end(az, res):
    Azul = az.Azul
    game = Azul()
    game.import_JSON(res["game_first_round"])
    assert game.is_legal_move(5, 0, 2) and not game.is_legal_move(1, 4, 2)
    game.move(5, 0, 2)
    assert game.is_legal_move(0, 1, 1) and not game.is_legal_move(0, 0, 0)
    game.import_JSON(res["game_first_round"])
    assert not game.is_legal_move(0, 0, 0)
    game.import_JSON(res["game_sample_1"])
    assert game.is_legal_move(0, 0, 5) and not game.is_legal_move(0, 1, 5)
    assert game.is_legal_move(5, 0, 3) and not game.is_legal_move(5, 2, 3)
    game = Azul(); game.new_round()
    seq = [game.current_player]
    for _ in range(2):
        game.next_player(); seq.append(game.current_player)
    assert seq == [1, 2, 1]
    game = Azul(players=4); game.new_round()
    seq = [game.current_player]
    for _ in range(4):
        game.next_player(); seq.append(game.current_player)
    assert seq == [1, 2, 3, 4, 1]
    game = Azul()
    game.import_JSON(res["game_sample_1"])
    assert not game.is_end_of_round()
    game.move(0, 0, 5)
    assert not game.is_end_of_round()
    game.import_JSON(res["game_end_of_round_1"])
    assert not game.is_end_of_round()
    game.move(0, 3, 3)
    assert game.is_end_of_round()
    game.import_JSON(res["game_end_of_round_2"])
    assert not game.is_end_of_game()
    game.move(0, 4, 1); game.next_player(); game.move(0, 0, 3); game.count_score()
    assert not game.is_end_of_game()
    game.import_JSON(res["game_end_of_round_2"])
    game.move(0, 0, 1); game.next_player(); game.move(0, 4, 1); game.count_score()
    assert game.is_end_of_game()


def test_azul_count_score(az, res):
    game = az.Azul()
    game.import_JSON(res["game_end_of_round_1"])
    prev = np.copy(game.score)
    game.count_score()
    assert np.array_equal(game.score, prev + np.array([5 + 5 + 1 - 2, 4 + 2 + 3 - 8]))
    assert np.array_equal(game.floors, np.zeros(2))
    assert np.array_equal(game.pattern_lines[0], [[0] * 5, [0] * 5, [0, 0, 2, 0, 0], [0] * 5, [0, 0, 3, 0, 0]])
    assert np.array_equal(game.pattern_lines[1], [[0] * 5, [0] * 5, [0] * 5, [2, 0, 0, 0, 0], [0] * 5])
    assert np.array_equal(game.walls[0], [[1, 1, 1, 0, 0], [1, 1, 1, 0, 0], [0] * 5, [0, 1, 0, 0, 1], [0] * 5])
    assert np.array_equal(game.walls[1], [[1, 1, 1, 1, 0], [0, 0, 0, 0, 1], [0, 0, 1, 0, 0], [0, 1, 0, 0, 0], [1, 1, 0, 0, 0]])
    game.import_JSON(res["game_end_of_round_1"])
    prev = np.copy(game.score)
    game.move(0, 3, 3); game.count_score()
    assert np.array_equal(game.score, prev + np.array([5 + 5 + 1 - 2, 4 + 2 + 3 + 3 - 8]))
    game.import_JSON(res["game_end_of_round_2"])
    prev = np.copy(game.score)
    game.move(0, 4, 1); game.next_player(); game.move(0, 0, 3); game.count_score()
    assert np.array_equal(game.score, prev + np.array([5 + 7 + 10, 5 + 7]))
    game.import_JSON(res["game_end_of_round_2"])
    prev = np.copy(game.score)
    game.move(0, 0, 1); game.next_player(); game.move(0, 4, 1); game.count_score()
    assert np.array_equal(game.score, prev + np.array([2 - 2, 5 + 2]))
    game.import_JSON(res["game_end_of_round_2"])
    game.move(0, 0, 0); game.next_player(); game.move(0, 4, 0); game.count_score()
    assert np.array_equal(game.score, [0, 0])


def test_azul_step(az, res):
    game = az.Azul()
    game.import_JSON(res["game_first_round"])
    game.step(5, 0, 2)
    assert np.array_equal(game.game_board_displays[4], np.zeros(5))
    assert np.array_equal(game.game_board_center, [0, 1, 2, 0, 0, 1])
    assert np.array_equal(game.pattern_lines[0, 1], [1, 0, 0, 0, 0]) and game.current_player == 2
    game.import_JSON(res["game_first_round"])
    before = copy.copy(game)
    with pytest.raises(az.IllegalMove):
        game.step(1, 4, 2)
    assert game == before
    game.import_JSON(res["game_end_of_round_1"])
    prev, nfp = np.copy(game.score), game.next_first_player
    game.step(0, 3, 3)
    assert all(np.sum(t) == 4 for t in game.game_board_displays)
    assert np.count_nonzero(game.game_board_center[:5]) == 0 and game.game_board_center[5] == 1
    assert np.array_equal(game.score, prev + np.array([5 + 5 + 1 - 2, 4 + 2 + 3 + 3 - 8]))
    assert game.current_player == nfp and game.next_first_player == 0
    game.import_JSON(res["game_end_of_round_2"])
    prev = np.copy(game.score)
    game.step(0, 0, 1)
    assert not game.end_of_game
    game.step(0, 4, 1)
    assert np.array_equal(game.score, prev + np.array([2 - 2, 5 + 2])) and game.end_of_game
    with pytest.raises(az.GameEnded):
        game.step(0, 0, 0)


# ---- tests/test_game_runner.py -----------------------------------------------------------------
def test_game_runner(az, res):
    gr = az.GameRunner()
    assert gr.move_counter == 0 and gr.player_score == 0
    assert all(np.sum(gr.game.game_board_displays[i]) == 4 for i in range(5))
    assert np.array_equal(gr.game.game_board_center, [0, 0, 0, 0, 0, 1])
    st = gr.get_state()
    assert np.sum(st) == 4 * 5 + 1 and np.size(st) == 136
    random.seed(1)
    gr = az.GameRunner()
    reward, end = gr.step(az.nn_serialize(1, 0, 2))
    assert gr.game.current_player == 1 and not end and np.array_equal(gr.game.score, np.zeros(2))
    gr = az.GameRunner()
    gr.game.import_JSON(res["game_end_of_round_2"])
    reward, end = gr.step(az.nn_serialize(0, 4, 1))
    assert gr.game.current_player == 1 and not end
    gr = az.GameRunner()
    gr.game.import_JSON(res["game_end_of_round_2"])
    gr.player_score = 49 - 32
    random.seed(1)
    reward, end = gr.step(az.nn_serialize(0, 0, 1))
    assert end and gr.player_score == gr.game.score[0] - gr.game.score[1]
    gr = az.GameRunner()
    gr.game.import_JSON(res["game_end_of_round_3"])
    gr.player_score = 49 - 32
    reward, end = gr.step(az.nn_serialize(0, 3, 0))
    assert not end and reward == -6                                     # tests/test_game_runner.py:52-62
    assert gr.player_score == gr.game.score[0] - gr.game.score[1]
    assert all(np.sum(gr.game.game_board_displays[i]) == 4 for i in range(5))
    assert np.array_equal(gr.game.game_board_center, [0, 0, 0, 0, 0, 1]) and gr.game.current_player == 1


def test_codec_and_check_all_valid(az, res):
    for i in range(6):
        for j in range(5):
            for k in range(6):
                assert (i, j, k) == az.nn_deserialize(az.nn_serialize(i, j, k))
    for i in range(180):
        assert i == az.nn_serialize(*az.nn_deserialize(i))
    game = az.Azul()
    assert np.array_equal(az.check_all_valid(game), np.zeros(180, dtype=bool))
    game.import_JSON(res["game_first_round"])
    v = az.check_all_valid(game)
    for d, cols in ((1, [0, 1, 2]), (2, [3]), (3, [0, 1, 3]), (4, [0, 3]), (5, [0, 1, 2])):
        for j in cols:
            for k in range(6):
                assert v[az.nn_serialize(d, j, k)]
    assert v.sum() == 6 * 12
    game.import_JSON(res["game_sample_1"])
    assert az.check_all_valid(game)[az.nn_serialize(0, 0, 4)]
    kat = load_kat()
    names = list(kat["fixture_names"])
    want = kat["fixture_masks"][names.index("game_sample_1")]
    bits = np.array([(int(want[a // 30]) >> (a % 30)) & 1 for a in range(180)], dtype=bool)
    assert np.array_equal(az.check_all_valid(game), bits)


# ---- tests/test_model.py, test_random_agent.py, test_nn_runner.py ------------------------------
def test_model_and_agents(az):
    net = az.ActorCritic(136, 180)
    state = torch.rand(1, 136)
    mask = torch.zeros((1, 180), dtype=torch.bool)
    with pytest.raises(az.IllegalMask):
        net.forward_actor(state, mask)
    mask[0, 17] = True
    p, logp = net.forward_actor(state, mask)
    assert float(p[0, 17]) == 1.0 and abs(float(p.sum()) - 1) < 1e-6
    mask[:] = True
    p, logp = net.forward_actor(state, mask)
    assert abs(float(p.sum()) - 1) < 1e-5 and net.forward_critic(state).shape == (1, 1)
    gr = az.GameRunner()
    ra = az.RandomAgent()
    for _ in range(200):
        valid = gr.get_valid_moves()
        a = ra.get_a_output(gr.get_state(), torch.from_numpy(valid.reshape(1, 180)))
        assert valid[a]


def test_nn_runner_episode_and_train(az):
    agent = az.Agent()
    runner = az.NNRunner(agent, az.GameRunner())
    r, v, lp, e = runner.run_episode()
    assert len(r) == len(v) == len(lp) == len(e) > 0
    before = [p.detach().clone() for p in agent.ac_net.parameters()]
    runner.train(batch_size=2, batches=1)
    assert any(not torch.equal(a, b) for a, b in zip(before, agent.ac_net.parameters()))
    assert len(agent.agent_statistics.statisticsBuffer["ac_loss"]) == 1
    runner = az.NNRunner(az.Agent(), az.GameRunner(opponent=az.Agent()))
    runner.run_batch(1)
    assert len(runner.game_runner.game_statistics.statistics["player_score"]) == 1
