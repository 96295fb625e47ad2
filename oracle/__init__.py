"""TEST INFRASTRUCTURE ONLY -- not part of the product.

Everything under ``oracle/`` is the checker for the CUDA hot path:

* ``ref_loader.py``   imports the UNMODIFIED reference (``/root/reference``) in the
  build container; it does not travel to the GPU box.
* ``record_golden.py`` drives the reference and writes ``tests/golden/*.npz``.
* ``azul_oracle.c``   a plain-C restatement of ``azulnet/azul.py`` +
  ``azulnet/game_runner.py`` (rules, legal mask, random agent), pinned against the
  golden vectors by ``tests/test_oracle_*.py``.
* ``oracle.py``       ctypes wrapper around the compiled restatement.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import from here.  The product package
(``azul_deep_reinforcement_learning_b200``) never does: it fails loudly when its
CUDA library is missing instead of falling back to this code.
"""
