"""Drop-in surface of the reference ``azulnet`` package (azulnet/__init__.py:1-5)."""
from .azul import Azul, GameEnded, IllegalMove, IllegalRule  # noqa: F401
