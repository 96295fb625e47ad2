"""Self-play A2C training on the GPU rollout engine -- the ``scripts/training.py`` / ``NNRunner.train``
equivalent (reference scripts/training.py:8-22, nn_runner.py:53-84, agent.py:39-62).

    python -m azul_deep_reinforcement_learning_b200.train [batch_size] [net_name] [--batches N]
    torchrun --nproc-per-node 8 -m azul_deep_reinforcement_learning_b200.train 1024 run1

``batch_size`` episodes are played IN PARALLEL per rank (one game slot each) instead of one after the
other; everything else follows the reference: discounted returns with gamma = 0.99, the loss of
``Agent.update`` (advantage not detached in the actor term, "entropy" = -mean(log pi over legal moves)
added with +0.1), Adam(lr 3e-4).  Rollouts use the fused tensor-core policy kernel (fp16 operands, fp32 accumulation); the update
recomputes the forward pass in fp32 with PyTorch autograd on the recorded observations.  With several
ranks each one plays its own shard of the global game-id range and the only collectives are one flat
gradient all-reduce per update (C1, 82,081 fp32 values) and the statistics reduction (C2).
"""
import argparse
import csv
import time

import torch
import torch.distributed as dist

from . import parallel
from .azulnet.model import ActorCritic
from .engine import PackedPolicy
from .selfplay import BatchedGameRunner, GraphedEpisodes, discounted_returns, run_episodes

ACTOR_COEFF, CRITIC_COEFF, ENTROPY_COEFF = 1.0, 0.5, 0.1          # agent.py:47-49


def a2c_loss_terms(net, obs, mask, action, qval):
    """Sums (not means) of the three loss terms of ``Agent.update`` over N transitions (agent.py:45-56).

    obs [N,136] float32, mask [N,180] bool, action [N] int64, qval [N] float32.  Returned as sums so that
    several ranks can divide by the GLOBAL transition count and all-reduce exact global-batch gradients."""
    value = net.forward_critic(obs).squeeze(1)
    logits = net.actor_linear2(torch.relu(net.actor_linear1(obs))).masked_fill(~mask, float("-inf"))
    logp_all = torch.log_softmax(logits, dim=1)
    log_prob = logp_all.gather(1, action[:, None]).squeeze(1)                     # nn_runner.py:32
    entropy = -(logp_all.masked_fill(~mask, 0.0).sum(dim=1) / mask.sum(dim=1))    # nn_runner.py:36-40
    advantage = qval - value                                                      # agent.py:45 (not detached)
    return (-log_prob * advantage).sum(), advantage.pow(2).sum(), entropy.sum()


class _Linear(torch.autograd.Function):
    """``x @ w.T + b`` whose bias gradient is a GEMM with a row of ones (tensor cores) instead of autograd's column
    reduction over the [N, out] gradient -- a fifth of the update's GPU time at 5*10^5 decisions."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return torch.addmm(b, x, w.t())

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = dy @ w if ctx.needs_input_grad[0] else None
        dw = dy.t() @ x
        db = (torch.ones(1, dy.shape[0], dtype=dy.dtype, device=dy.device) @ dy).squeeze(0)
        return dx, dw, db


def network_outputs(net, x):
    """(logits [N,180] raw, value [N]) of ``ActorCritic`` (model.py:23-41 before the mask) through :class:`_Linear`."""
    lin = _Linear.apply
    logits = lin(torch.relu(lin(x, net.actor_linear1.weight, net.actor_linear1.bias)), net.actor_linear2.weight, net.actor_linear2.bias)
    value = lin(torch.relu(lin(x, net.critic_linear1.weight, net.critic_linear1.bias)), net.critic_linear2.weight, net.critic_linear2.bias)
    return logits, value.squeeze(1)


def allreduce_gradients(params):
    """C1: one flat all-reduce (sum) over every gradient; the caller already divided by the global count."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


def allreduce_gradients_and_stats(params, stats):
    """C1 + C2 in ONE collective: every gradient (sums over the local decisions, not yet divided by a count) and the
    statistics vector travel as one float64 buffer; returns the reduced statistics.  No host synchronisation."""
    flat = torch.cat([p.grad.reshape(-1).double() for p in params] + [stats.double().reshape(-1)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat[off:]


def global_count(n_local, device):
    t = torch.tensor([float(n_local)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


class SelfPlayTrainer:
    def __init__(self, games_per_rank=1024, learning_rate=3e-4, gamma=0.99, seed=0, device=0, rank=0, world=1,
                 rules=None, max_decisions=160, use_cuda_graph=True, tf32_update=True):
        self.rank, self.world, self.gamma, self.max_decisions = rank, world, gamma, max_decisions
        self.tf32_update = tf32_update
        self.device = torch.device("cuda", device)
        torch.manual_seed(seed)                              # identical initial weights on every rank
        self.net = ActorCritic(136, 180).to(self.device)
        self.params = list(self.net.parameters())
        self.opt = torch.optim.Adam(self.params, lr=learning_rate)       # agent.py:37
        self.runner = BatchedGameRunner(games_per_rank, rules=rules, seed=seed, device=device,
                                        game_id_base=parallel.shard(rank, games_per_rank), record_obs=True)
        self.packed = PackedPolicy(self.runner.engine, self.net)
        self.graphed = GraphedEpisodes(self.runner, self.packed) if use_cuda_graph else None
        self.history = []

    def rollout(self):
        with torch.no_grad():
            self.packed.update(self.net)
            if self.graphed is not None:
                batch = self.graphed.run(max_decisions=self.max_decisions)
            else:
                batch = run_episodes(self.runner, self.packed, max_decisions=self.max_decisions)
            batch["qval"] = discounted_returns(batch["reward"], batch["active"], self.gamma)
            batch["stats"] = self.runner.engine.stats().to(torch.float64)
        return batch

    def update(self, batch, chunk=1 << 18):
        """``Agent.update`` (agent.py:39-62) on every recorded agent decision of the batch.

        Gradients are accumulated as SUMS over the local decisions; one all-reduce carries them together with the
        decision count and the batch statistics, then every rank divides by the global count (the means of
        agent.py:47-56 over the global batch) and takes the same Adam step.  The host is not synchronised before
        the statistics are read at the end."""
        act = batch["active"]
        T, G = act.shape
        sel = act.reshape(-1).nonzero(as_tuple=True)[0]
        n_local = int(sel.numel())
        obs = batch["obs"].reshape(T * G, -1)
        masks = batch["mask"].permute(0, 2, 1).reshape(T * G, 6)
        action = batch["action"].reshape(-1)
        qval = batch["qval"].reshape(-1)
        self.opt.zero_grad(set_to_none=False)
        for p in self.params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        sums = torch.zeros(3, dtype=torch.float64, device=self.device)
        # The dense layers run on the tensor cores through cuBLAS (TF32: fp32 storage and accumulation, 10-bit operand
        # mantissas -- about as fine as the fp16 operands the rollout's decisions were taken with); everything between the
        # network outputs and the loss -- masked log-softmax, the three terms, their gradient -- is one kernel
        # (azb_a2c_loss_grad), back-propagated through the layers by autograd.
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.tf32_update
        try:
            for lo in range(0, n_local, chunk):
                idx = sel[lo:lo + chunk]
                x = obs[idx].float()
                logits, value = network_outputs(self.net, x)
                dlogits, dvalue = self.runner.engine.a2c_loss_grad(
                    logits.detach(), value.detach().contiguous(), masks[idx].contiguous(), action[idx], qval[idx],
                    1.0, (ACTOR_COEFF, CRITIC_COEFF, ENTROPY_COEFF), sums)
                torch.autograd.backward([logits, value], [dlogits, dvalue])
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32
        t0 = time.perf_counter()
        one = torch.ones(1, dtype=torch.float64, device=self.device)
        wins = (batch["stats"][:, 0] > batch["stats"][:, 1]).double().sum().reshape(1)
        stats = torch.cat([one * n_local, sums, batch["reward"].double().mul(act).sum().reshape(1), one * G,
                           batch["stats"].sum(dim=0), wins])
        stats = allreduce_gradients_and_stats(self.params, stats)
        inv = (1.0 / stats[0].clamp_min(1.0)).float()
        for p in self.params:
            p.grad.mul_(inv)
        self.opt.step()
        s = stats.cpu().tolist()
        n_global, games = max(s[0], 1.0), s[5]
        out = {"transitions": s[0], "games": games, "actor_loss": s[1] / n_global, "critic_loss": s[2] / n_global,
               "entropy_loss": s[3] / n_global, "reward": s[4] / games, "update_sync_s": time.perf_counter() - t0}
        out["ac_loss"] = ACTOR_COEFF * out["actor_loss"] + CRITIC_COEFF * out["critic_loss"] + ENTROPY_COEFF * out["entropy_loss"]
        g = [x / games for x in s[6:16]]
        out.update(player_score=g[0], opponent_score=g[1], rounds=g[2],
                   percent_first_player=100.0 * s[9] / max(s[10], 1.0), floor_penalty=g[5], max_combo=g[6],
                   completed_rows=g[7], completed_columns=g[8], completed_colors=g[9])
        out["win_percent"] = s[16] / games
        return out

    def train(self, batches=1000, net_name=None, log=print):
        writer = fh = None
        for b in range(batches):
            t0 = time.perf_counter()
            batch = self.rollout()
            torch.cuda.synchronize(self.device)
            t1 = time.perf_counter()
            st = self.update(batch)
            torch.cuda.synchronize(self.device)
            t2 = time.perf_counter()
            st.update(batch=b + 1, rollout_s=t1 - t0, update_s=t2 - t1, games_per_sec=st["games"] / (t2 - t0),
                      unfinished=batch["unfinished"])
            self.history.append(st)
            if self.rank == 0:
                if net_name is not None:
                    if writer is None:
                        fh = open(net_name + ".csv", "w", newline="")
                        writer = csv.DictWriter(fh, fieldnames=list(st.keys()))
                        writer.writeheader()
                    writer.writerow(st)
                    fh.flush()
                    if (b + 1) % 1000 == 0 or b + 1 == batches:           # nn_runner.py:83-84, as a state_dict
                        torch.save({"ac_net": self.net.state_dict(), "optimizer": self.opt.state_dict(), "batch": b + 1},
                                   net_name + ".pt")
                if log:
                    log("batch %d: %.0f games/s  reward %.2f  score %.1f vs %.1f  win %.1f%%  loss %.3f" % (
                        b + 1, st["games_per_sec"], st["reward"], st["player_score"], st["opponent_score"],
                        100 * st["win_percent"], st["ac_loss"]))
        if fh:
            fh.close()
        return self.history


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("batch_size", nargs="?", type=int, default=10, help="episodes per batch and rank (scripts/training.py:8-11)")
    ap.add_argument("net_name", nargs="?", default=None, help="write <net_name>.csv / .pt (scripts/training.py:12-15)")
    ap.add_argument("--batches", type=int, default=1000)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args(argv)
    rank, world, local = parallel.world()
    if world > 1:
        parallel.init("nccl", local)
    torch.cuda.set_device(local)
    tr = SelfPlayTrainer(args.batch_size, learning_rate=args.lr, seed=args.seed, device=local, rank=rank, world=world)
    tr.train(batches=args.batches, net_name=args.net_name)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
