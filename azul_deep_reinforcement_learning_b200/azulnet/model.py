"""``azulnet.model`` drop-in: the actor-critic MLP (reference ``azulnet/model.py:12-41``).

Two independent two-layer heads: actor 136 -> hidden -> ReLU -> 180 logits, critic 136 -> hidden ->
ReLU -> 1.  Parameter names match the reference so ``state_dict``s are interchangeable.  This class
is the single-sample host form used by ``Agent`` / ``NNRunner``; the batched forward + masked
softmax + sampling over 10^5 games is the fused CUDA policy kernel (DESIGN.md "next").
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class IllegalMask(Exception):
    pass


class ActorCritic(nn.Module):
    def __init__(self, num_inputs, num_actions, hidden_size=180, learning_rate=3e-6):
        super().__init__()
        self.num_actions = num_actions
        self.critic_linear1 = nn.Linear(num_inputs, hidden_size)
        self.critic_linear2 = nn.Linear(hidden_size, 1)
        self.actor_linear1 = nn.Linear(num_inputs, hidden_size)
        self.actor_linear2 = nn.Linear(hidden_size, num_actions)

    def forward_critic(self, state_tensor):
        return self.critic_linear2(F.relu(self.critic_linear1(state_tensor)))

    def forward_actor(self, state_tensor, mask=None):
        if mask is None:
            mask = torch.ones((1, self.num_actions), dtype=torch.bool)
        elif int(mask.sum()) == 0:
            raise IllegalMask                              # model.py:33-34: softmax of all -inf would be NaN
        logits = self.actor_linear2(F.relu(self.actor_linear1(state_tensor)))
        logits = logits.masked_fill(~mask, float("-inf"))  # model.py:37
        return F.softmax(logits, dim=1), F.log_softmax(logits, dim=1)
