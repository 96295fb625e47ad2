"""Small end-to-end exercise of the kernels added in round 2 (for compute-sanitizer runs): persistent policy rollouts in
both modes, discounted returns, the tensor-core update, the factory-count variant."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from azul_deep_reinforcement_learning_b200.azulnet.model import ActorCritic
from azul_deep_reinforcement_learning_b200.engine import (BatchedAzul, BatchedAzulByPlayers, PackedPolicy, UpdateGradients)
from azul_deep_reinforcement_learning_b200.selfplay import BatchedGameRunner, PersistentEpisodes

torch.manual_seed(0)
net = ActorCritic(136, 180)
eng = BatchedAzul(300, 2, 1, 0, seed=5)
packed = PackedPolicy(eng, net)
eng.policy_rollout(packed, 12, want_last=True)
gr = BatchedGameRunner(300, seed=7)
pk = PackedPolicy(gr.engine, net)
recs = PersistentEpisodes(gr, pk, max_decisions=120).run(0.99)
upd = UpdateGradients(gr.engine, recs.cap)
upd.run(pk, recs.state_rec, recs.action_rec, recs.qval, n_dec=recs.meta[:1])
for players in (3, 4):
    v = BatchedAzulByPlayers(500, players, 1, 0, seed=3)
    v.rollout_random(80)
    v.legal_mask()
eng.step(torch.zeros(300, dtype=torch.uint8))
eng.rollout_random(50)
torch.cuda.synchronize()
print("ok", int(recs.meta[0]), float(upd.flat.abs().sum()), eng.read_counters()["steps"])
