"""B200-native batched Azul engine behind the ``azulnet`` API of patello/azul_deep_reinforcement_learning."""
from .layout import (  # noqa: F401
    FIRST_PLAYER_RANDOM, N_ACTIONS, TILE_POOL_LID, TILE_POOL_RANDOM, UnpackedLayout, algorithmic_bytes_per_step,
    state_bytes, state_words, unpacked_size)


def __getattr__(name):
    # torch / CUDA are only needed once the engine is touched
    if name in ("BatchedAzul", "mask_to_bool", "COUNTER_NAMES", "PackedPolicy", "policy_step"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
