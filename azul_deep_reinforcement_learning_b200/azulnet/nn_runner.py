"""``azulnet.nn_runner`` drop-in: episode rollout and the training loop (reference ``nn_runner.py:13-84``)."""
import numpy as np
import torch


class NNRunner:
    def __init__(self, agent, game_runner):
        self.agent = agent
        self.game_runner = game_runner

    def run_episode(self):
        """One episode of seat 1 vs the runner's opponent; per agent move: reward, value, log-prob of the
        chosen action and the entropy term -mean(log p over valid moves) (nn_runner.py:32-40)."""
        rewards, values, log_probs, entropy_terms = [], [], [], []
        self.game_runner.reset()
        state = self.game_runner.get_state()
        done = False
        while not done:
            valid = torch.from_numpy(self.game_runner.get_valid_moves().reshape(1, 180))
            action, _, log_policy, value = self.agent.get_ac_output(state, valid)
            reward, done = self.game_runner.step(action)
            state = self.game_runner.get_state()
            rewards.append(reward)
            values.append(value)
            log_probs.append(log_policy.squeeze(0)[action])
            entropy_terms.append(-log_policy.masked_select(valid).mean())
        return rewards, values, log_probs, entropy_terms

    def run_batch(self, episodes):
        for _ in range(episodes):
            self.run_episode()
        stats = self.game_runner.game_statistics.get_stats()
        for k in stats:
            print(k + ": " + str(stats[k][-1]))

    def train(self, net_name=None, batch_size=1000, batches=1000):
        """``batches`` updates of ``batch_size`` episodes each: discounted returns (gamma = agent.gamma)
        then ``agent.update``.  CSV / checkpoint output (dead in the reference: nn_runner.py:55-57,79-84
        reference a missing attribute) is written to ``net_name`` + .csv / .pt when given."""
        log = None
        if net_name is not None:
            log = open(net_name + ".csv", "w")
            log.write(",".join(["batch"] + list(self.agent.agent_statistics.statistics.keys()) +
                               list(self.game_runner.game_statistics.statistics.keys())) + "\n")
        for batch in range(batches):
            rewards, values, log_probs, entropy, qvals = [], [], [], [], []
            for _ in range(batch_size):
                ep_r, ep_v, ep_lp, ep_e = self.run_episode()
                rewards.append(np.sum(ep_r))
                values += ep_v
                log_probs += ep_lp
                entropy += ep_e
                q, ep_q = 0.0, np.zeros(len(ep_r))
                for t in reversed(range(len(ep_r))):
                    q = ep_r[t] + self.agent.gamma * q                       # nn_runner.py:72-75
                    ep_q[t] = q
                qvals.append(ep_q)
            self.agent.update(np.concatenate(qvals).reshape(-1, 1), rewards, values, log_probs, entropy)
            if log is not None:
                a = [v[-1] for v in self.agent.agent_statistics.get_stats().values()]
                g = [v[-1] if len(v) else float("nan") for v in self.game_runner.game_statistics.get_stats().values()]
                log.write(",".join(str(x) for x in [batch + 1] + a + g) + "\n")
                log.flush()
                if (batch + 1) % 1000 == 0 or batch + 1 == batches:
                    torch.save(self.agent.ac_net.state_dict(), net_name + ".pt")
        if log is not None:
            log.close()
