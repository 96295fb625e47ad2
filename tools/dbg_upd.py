import torch, numpy as np, sys
sys.path.insert(0,'/root/repo')
from azul_deep_reinforcement_learning_b200.azulnet.model import ActorCritic
from azul_deep_reinforcement_learning_b200.engine import PARAM_ORDER, PackedPolicy, UpdateGradients
from azul_deep_reinforcement_learning_b200.selfplay import BatchedGameRunner, PersistentEpisodes
from tests.test_selfplay_gpu import _autograd_reference
for games, scale in ((300,1.0),(1500,2.5)):
    torch.manual_seed(3)
    net = ActorCritic(136, 180).cuda()
    with torch.no_grad():
        for p in net.parameters(): p.mul_(scale)
    gr = BatchedGameRunner(games, seed=8)
    packed = PackedPolicy(gr.engine, net)
    recs = PersistentEpisodes(gr, packed, max_decisions=160).run(gamma=0.99)
    n = int(recs.meta[0])
    coeffs=(1.0,0.5,0.1)
    want, want_sums, wl, wv = _autograd_reference(net, recs, n, coeffs)
    upd = UpdateGradients(gr.engine, recs.cap)
    logits, value = upd.run(packed, recs.state_rec, recs.action_rec, recs.qval, n_dec=recs.meta[:1], coeffs=coeffs, want_outputs=True)
    torch.cuda.synchronize()
    print('games',games,'n',n,'sums',upd.sums.tolist(), want_sums.tolist())
    for name in PARAM_ORDER:
        g,w = upd.grads[name].double().flatten(), want[name].double().flatten()
        cos = float((g*w).sum()/(g.norm()*w.norm()+1e-30))
        print('%-24s max|w| %10.4f maxerr %10.4f rel %.2e  cos %.6f  ratio %.4f' % (name, float(w.abs().max()), float((g-w).abs().max()), float((g-w).abs().max()/w.abs().max()), cos, float(g.norm()/w.norm())))
    # where is the error in w1a: by column (obs feature)
    e = (upd.grads['actor_linear1.weight']-want['actor_linear1.weight']).abs()
    print('w1a err by col top:', torch.topk(e.max(0).values,8))
    print('w1a |w| by col top:', torch.topk(want['actor_linear1.weight'].abs().max(0).values,8))
