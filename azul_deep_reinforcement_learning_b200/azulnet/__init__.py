"""Drop-in surface of the reference ``azulnet`` package (azulnet/__init__.py:1-5), on the CUDA engine."""
from .agent import Agent  # noqa: F401
from .azul import Azul, GameEnded, IllegalMove, IllegalRule  # noqa: F401
from .game_runner import GameRunner, RandomAgent, check_all_valid, nn_deserialize, nn_serialize  # noqa: F401
from .model import ActorCritic, IllegalMask  # noqa: F401
from .nn_runner import NNRunner  # noqa: F401
