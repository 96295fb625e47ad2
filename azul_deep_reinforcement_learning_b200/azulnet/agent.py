"""``azulnet.agent`` drop-in: the A2C agent (reference ``azulnet/agent.py:9-81``)."""
import os

import numpy as np
import torch
import torch.optim as optim

from .model import ActorCritic

_KEYS = ["reward", "actor_loss", "critic_loss", "entropy_loss", "ac_loss"]


def load_ac_net(base_net_file, num_in=136, num_out=180):
    """Load a saved network.  ``base_net_file`` is a path, or -- like the reference (agent.py:36) -- a bare name that
    resolves to ``/results/<name>.mx``; ``<name>.pt`` next to it is tried too.  Accepted contents: a pickled
    ``ActorCritic`` module (what the reference writes, nn_runner.py:83-84), a bare ``state_dict`` (what this package's
    ``NNRunner.train`` writes) or the trainer's checkpoint dict ``{"ac_net": state_dict, "optimizer": ..., "batch": n}``."""
    candidates = [base_net_file, base_net_file + ".pt", base_net_file + ".mx",
                  "/results/" + base_net_file + ".mx", "/results/" + base_net_file + ".pt"]
    path = next((c for c in candidates if os.path.isfile(c)), None)
    if path is None:
        raise FileNotFoundError("no saved network among " + ", ".join(candidates))
    obj = torch.load(path, map_location="cpu", weights_only=False)
    if isinstance(obj, torch.nn.Module):
        return obj
    net = ActorCritic(num_in, num_out)
    net.load_state_dict(obj["ac_net"] if isinstance(obj, dict) and "ac_net" in obj else obj)
    return net


class Agent:
    class AgentStatistics:
        def __init__(self):
            self.statisticsBuffer = {k: np.empty(0) for k in _KEYS}
            self.statistics = {k: np.empty(0) for k in _KEYS}

        def update(self, statistics):
            for k, v in statistics.items():
                self.statisticsBuffer[k] = np.append(self.statisticsBuffer[k], v)

        def get_stats(self):
            for k in self.statistics:
                if len(self.statisticsBuffer[k]) > 0:
                    self.statistics[k] = np.append(self.statistics[k], self.statisticsBuffer[k].mean())
                    self.statisticsBuffer[k] = np.empty(0)
            return self.statistics

    def __init__(self, base_net_file=None, base_net="Blue Adam", learning_rate=3e-4, gamma=0.99):
        self.agent_statistics = Agent.AgentStatistics()
        self.learning_rate = learning_rate
        self.gamma = gamma
        self.num_in = 136
        self.num_out = 180
        if base_net_file is None:
            if base_net == "Blue Adam":
                self.ac_net = ActorCritic(self.num_in, self.num_out)
        else:
            self.ac_net = load_ac_net(base_net_file, self.num_in, self.num_out)                  # agent.py:36
        self.ac_optimizer = optim.Adam(self.ac_net.parameters(), lr=learning_rate)

    def update(self, qvals, rewards, values, log_probs, entropy):
        """One A2C step with the reference's loss (agent.py:39-62): advantage NOT detached in the actor
        term, "entropy" = mean of -mean(log p over valid moves) added with a positive 0.1 coefficient."""
        values = torch.stack(values).squeeze(2)
        qvals = torch.as_tensor(np.asarray(qvals), dtype=torch.float32)
        log_probs = torch.stack(log_probs)
        entropy = torch.stack(entropy)
        advantage = qvals - values
        actor_loss = (-log_probs * advantage.squeeze(1)).mean()
        critic_loss = advantage.pow(2).mean()
        entropy_loss = entropy.mean()
        ac_loss = 1 * actor_loss + 0.5 * critic_loss + 0.1 * entropy_loss
        self.ac_optimizer.zero_grad()
        ac_loss.backward()
        self.ac_optimizer.step()
        self.agent_statistics.update({
            "reward": np.mean(rewards), "actor_loss": float(actor_loss.detach()), "critic_loss": float(critic_loss.detach()),
            "entropy_loss": float(entropy_loss.detach()), "ac_loss": float(ac_loss.detach())})

    def _select(self, policy_dist, action_selection):
        p = policy_dist.detach().numpy().squeeze(0)
        if action_selection == "Distribution":
            return np.random.choice(self.num_out, p=p)      # agent.py:69
        if action_selection == "Max":
            return int(np.argmax(p))
        raise ValueError(action_selection)

    def get_ac_output(self, state, valid_moves, action_selection="Distribution"):
        state = torch.from_numpy(np.asarray(state)).float().unsqueeze(0)
        value = self.ac_net.forward_critic(state)
        policy_dist, log_policy_dist = self.ac_net.forward_actor(state, valid_moves)
        return self._select(policy_dist, action_selection), policy_dist, log_policy_dist, value

    def get_a_output(self, state, valid_moves, action_selection="Distribution"):
        state = torch.from_numpy(np.asarray(state)).float().unsqueeze(0)
        policy_dist, _ = self.ac_net.forward_actor(state, valid_moves)
        return self._select(policy_dist, action_selection)
