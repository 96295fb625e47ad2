"""Batched counterparts of ``GameRunner`` / ``NNRunner`` (reference game_runner.py:9-85, nn_runner.py:13-84).

``BatchedGameRunner`` is G copies of the reference's 2-player env wrapper advancing in lock-step on one
GPU: seat 1 is the learning agent, seat 2 a random agent (``RandomAgent``, game_runner.py:87-97, drawing
from the Philox schedule).  One agent decision costs three launches: observation (``azb_observe``, only
when the caller records it), the fused policy kernel (``azb_policy_step``: MLP on tcgen05 + masked
softmax + sampling + the agent's ``Azul.step``) and the opponent loop + reward (``azb_opponent_random``).
"""
import torch

from .engine import (BatchedAzul, EpisodeRecords, discounted_returns_records, mask_to_bool, policy_step, rules_to_ints,
                     runner_rollout)

DEFAULT_RULES = {"first_player": "Random", "tile_pool": "Lid"}      # game_runner.py:23


class BatchedGameRunner:
    def __init__(self, n_games, rules=None, seed=0, device=0, game_id_base=0, opponent=None, opponent_mode=0,
                 record_obs=False):
        """``opponent``: None = RandomAgent (game_runner.py:29-30) or a ``PackedPolicy`` = a frozen ``Agent``
        (scripts/run_batch.py:6-8) whose moves come from the fused policy kernel.  ``record_obs``: the opponent /
        reward kernel also writes the next decision's observation (``self.obs``, bfloat16 [G, 136])."""
        self.opponent, self.opponent_mode, self.record_obs = opponent, opponent_mode, record_obs
        self.obs = None
        rules = DEFAULT_RULES if rules is None else rules
        pool, first = rules_to_ints(2, rules)
        self.engine = BatchedAzul(n_games, 2, pool, first, seed=seed, device=device, game_id_base=game_id_base, reset=False)
        self.n_games, self.device = n_games, self.engine.device
        self.player_score = torch.zeros(n_games, dtype=torch.int16, device=self.device)
        self.mask = None
        self.reset()

    def reset(self):
        """``GameRunner.reset`` (game_runner.py:76-85): fresh games, the opponent plays until seat 1 is to move."""
        self.engine.reset()
        self.player_score.zero_()
        if self.opponent is not None:                 # a fresh board always offers seat 1 >= 2 actions, so the
            self._opponent_policy_moves()             # ":46" and ":84" loop conditions coincide here
        out = self.engine.opponent_random(self.player_score, require_two=False, want_obs=self.record_obs)
        self.player_score.zero_()                     # game_runner.py:81: the score baseline restarts at 0
        self.mask, self.obs = out["mask"], out.get("obs")
        return out

    def get_state(self):
        """``GameRunner.get_state()`` for every game: float32 [G, 136] from seat 1's perspective."""
        return self.engine.observe(0)

    def get_valid_moves(self):
        return mask_to_bool(self.mask)

    def _opponent_policy_moves(self, max_rounds=64, check_every=1):
        """``while (current_player != 1 or #valid < 2) and not end: opponent_move()`` (game_runner.py:46-47) with an
        Agent opponent: each launch lets every game that is still in the opponent's hands take one policy decision
        (``act_filter = 1``); typically one or two launches per agent decision.  The loop runs until no game acted
        (checked on the host every ``check_every`` launches: a launch in which nobody acts changes nothing); a round
        is at most 20 takes long, so ``max_rounds`` launches can only be exceeded by a bug -- that raises."""
        for i in range(max_rounds):
            out = policy_step(self.engine, self.opponent, mode=self.opponent_mode, apply_step=True, want_mask=False,
                              act_filter=1)
            if (i + 1) % check_every == 0 and not bool((out["action"] != 255).any()):
                return
        if bool((out["action"] != 255).any()):
            raise RuntimeError("opponent loop did not hand the move back to seat 1 within %d policy launches" % max_rounds)

    def _after_agent_move(self):
        if self.opponent is not None:
            self._opponent_policy_moves()
        out = self.engine.opponent_random(self.player_score, require_two=True, want_obs=self.record_obs)
        self.mask, self.obs = out["mask"], out.get("obs")
        return out

    def step(self, actions):
        """``GameRunner.step`` (game_runner.py:43-55) with one action per game (uint8, 255 = skip that game).
        Returns dict(reward int16, done uint8, status uint8 [step status | opponent status], mask)."""
        st = self.engine.step(actions, None, want_mask=False)
        out = self._after_agent_move()
        out["status"] = out["status"] | st["status"]
        return out

    def step_policy(self, packed, mode=0):
        """Agent decision by the fused policy kernel + its env step, then the opponent loop and the reward."""
        pol = policy_step(self.engine, packed, mode=mode, apply_step=True, want_mask=False)
        out = self._after_agent_move()
        out.update(action=pol["action"], logp=pol["logp"], value=pol["value"], entropy=pol["entropy"],
                   policy_status=pol["status"])
        return out


RAW_KEYS = ("obs", "mask", "reward", "value", "logp", "entropy", "action", "status", "done")


def _one_decision(runner, packed, rec, record_obs):
    """One agent decision for every game; appends the raw outputs of the kernels (what ``NNRunner.run_episode``
    records, nn_runner.py:27-45) without any per-decision bookkeeping -- see :func:`_alive_chain`."""
    if record_obs:                           # the previous opponent / reward launch already wrote it when asked to
        rec["obs"].append(runner.obs if runner.obs is not None else runner.engine.observe_bf16(0))
    rec["mask"].append(runner.mask)          # every opponent_random call returns a new tensor: no copy needed
    out = runner.step_policy(packed)
    rec["reward"].append(out["reward"])
    rec["value"].append(out["value"])
    rec["logp"].append(out["logp"])
    rec["entropy"].append(out["entropy"])
    rec["action"].append(out["action"])
    rec["status"].append(out["policy_status"])
    rec["done"].append(out["done"])


def _alive_chain(status, done, alive0):
    """``active[t]`` = the game was still running when decision t was taken and a decision was taken
    (policy status without ENDED / STUCK); a game stays alive while it acts and is not done.  Vectorised over the
    decisions of a chunk: status, done uint8 [C, G]; alive0 bool [G].  Returns (active bool [C, G], alive bool [G])."""
    ok = (status & 6) == 0
    live = (ok & (done == 0)).to(torch.uint8)
    chain = torch.cummin(live, dim=0).values.bool()                       # alive after decision t (given alive0)
    before = torch.cat([torch.ones_like(chain[:1]), chain[:-1]])
    return before & ok & alive0, chain[-1] & alive0


def _finish_batch(chunks, alive):
    out = {k: torch.cat([c[k] for c in chunks]) for k in chunks[0]}
    out["reward"] = out["reward"].to(torch.float32)
    out["action"] = out["action"].to(torch.int64)
    del out["status"], out["done"]
    out["unfinished"] = int(alive.sum())
    return out


def _stack_chunk(rec, alive):
    chunk = {k: torch.stack(v) for k, v in rec.items() if v}
    chunk["active"], alive = _alive_chain(chunk["status"], chunk["done"], alive)
    return chunk, alive


def run_episodes(runner, packed, max_decisions=160, check_every=8, record_obs=True):
    """``NNRunner.run_episode`` for every game of ``runner`` at once (nn_runner.py:17-47).

    Returns a dict of [T, G] tensors: reward (float32), value, logp, entropy (from the kernel), action
    (int64), active (bool: the game was still running when the decision was taken), mask (int32
    [T, 6, G] legal-mask words) and, when ``record_obs``, obs (bfloat16 [T, G, 136]) for the autograd
    recomputation."""
    G, dev = runner.n_games, runner.device
    runner.reset()
    alive = torch.ones(G, dtype=torch.bool, device=dev)
    chunks, t = [], 0
    while t < max_decisions:
        rec = {k: [] for k in RAW_KEYS}
        for _ in range(min(check_every, max_decisions - t)):
            _one_decision(runner, packed, rec, record_obs)
        t += check_every
        chunk, alive = _stack_chunk(rec, alive)
        chunks.append(chunk)
        if not bool(alive.any()):
            break
    return _finish_batch(chunks, alive)


class GraphedEpisodes:
    """``run_episodes`` with its launch-bound inner loop captured in CUDA graphs.

    One agent decision is five small kernels (observation record, policy, round finisher, opponent loop + reward,
    and nothing else: the bookkeeping is vectorised over the decisions afterwards); the loop is bound by launch
    latency, not by the GPU.  The first ``decisions`` decisions of an episode batch (reset included) are recorded
    into one CUDA graph, and a second graph holds ``more`` further decisions; batches in which some game is still
    running replay the second graph until all are done (its outputs are copied out after every replay).  Results
    are identical to :func:`run_episodes`."""

    def __init__(self, runner, packed, decisions=48, record_obs=True, more=8):
        self.runner, self.packed, self.decisions, self.record_obs, self.more = runner, packed, decisions, record_obs, more
        dev = runner.device
        # warm-up on a side stream (lazy CUDA initialisation must not happen during capture)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            runner.reset()
            self._body(2)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            runner.reset()
            self.rec = self._body(decisions)
            self.mask_out, self.obs_out = runner.mask, runner.obs
        # continuation: starts from the engine's current state and the legal mask / observation in self.mask_in / self.obs_in (copied in before a
        # replay: a graph reads fixed addresses)
        self.mask_in = torch.zeros_like(self.mask_out)
        self.obs_in = None if self.obs_out is None else torch.zeros_like(self.obs_out)
        self.graph_more = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_more, pool=self.graph.pool()):
            runner.mask, runner.obs = self.mask_in, self.obs_in
            self.rec_more = self._body(more)
            self.mask_more, self.obs_more = runner.mask, runner.obs
        # the first replay of a graph pays its upload (tens of ms): pay it here, not in the middle of training
        # (every run() starts with the reset inside the first graph, so the games played here are discarded)
        self.graph.replay()
        self.graph_more.replay()
        torch.cuda.synchronize(dev)

    def _body(self, decisions):
        rec = {k: [] for k in RAW_KEYS}
        for _ in range(decisions):
            _one_decision(self.runner, self.packed, rec, self.record_obs)
        return {k: torch.stack(v) for k, v in rec.items() if v}

    def run(self, max_decisions=160, check_every=8):
        G, dev = self.runner.n_games, self.runner.device
        self.graph.replay()
        chunk = dict(self.rec)
        chunk["active"], alive = _alive_chain(chunk["status"], chunk["done"], torch.ones(G, dtype=torch.bool, device=dev))
        chunks, mask, obs, t = [chunk], self.mask_out, self.obs_out, self.decisions
        while t < max_decisions and bool(alive.any()):
            self.mask_in.copy_(mask)
            if obs is not None:
                self.obs_in.copy_(obs)
            self.graph_more.replay()
            chunk = {k: v.clone() for k, v in self.rec_more.items()}
            chunk["active"], alive = _alive_chain(chunk["status"], chunk["done"], alive)
            chunks.append(chunk)
            mask, obs = self.mask_more, self.obs_more
            t += self.more
        self.runner.mask, self.runner.obs = mask, obs
        return _finish_batch(chunks, alive)


class PersistentEpisodes:
    """``NNRunner.run_episode`` for every game slot with the whole episode loop on the device: ``GameRunner.reset`` (two
    launches), then ONE persistent launch of the fused policy kernel that takes every agent decision, plays the random
    opponent's moves, takes the rewards and writes the decision records, then the discounted returns (one launch).  The
    recorded decisions are those of :func:`run_episodes` (same Philox schedule); only their storage differs (compact
    slots in arbitrary order instead of [T, G])."""

    def __init__(self, runner, packed, max_decisions=160, capacity=None, want_logp_value=False):
        assert runner.opponent is None, "the persistent rollout plays the RandomAgent opponent"
        self.runner, self.packed = runner, packed
        n = runner.n_games
        if capacity is None:                       # ~30 decisions per episode on average; small batches get the full K x G
            capacity = max_decisions * n if n <= 4096 else 56 * n
        self.records = EpisodeRecords(runner.engine, max_decisions, capacity, want_logp_value)
        self.out = {"mask": torch.zeros((6, n), dtype=torch.int32, device=runner.device),
                    "done": torch.zeros(n, dtype=torch.uint8, device=runner.device),
                    "status": torch.zeros(n, dtype=torch.uint8, device=runner.device)}

    def run(self, gamma=0.99, mode=0):
        r = self.records
        self.runner.reset()
        r.clear()
        runner_rollout(self.runner.engine, self.packed, r, self.runner.player_score, mode=mode, out=self.out)
        discounted_returns_records(self.runner.engine, r, gamma)
        self.runner.mask = self.out["mask"]
        return r


def discounted_returns(reward, active, gamma):
    """q_t = r_t + gamma * q_{t+1} per episode (nn_runner.py:72-75), [T, G]; inactive slots contribute 0."""
    T = reward.shape[0]
    q = torch.zeros_like(reward)
    run = torch.zeros_like(reward[0])
    r = reward * active
    for t in range(T - 1, -1, -1):
        run = r[t] + gamma * run
        q[t] = run
    return q
