"""The reference's OWN test files, unchanged (baseline/_ref/tests, staged by __graft_entry__.build from /root/reference),
run against this package's ``azulnet`` drop-in on the CUDA path (28 tests; all pass on the reference itself).

Mechanism: a scratch directory on PYTHONPATH holds an ``azulnet`` alias package whose ``__init__`` replaces itself (and its
five submodules) in ``sys.modules`` by ``azul_deep_reinforcement_learning_b200.azulnet``, and a two-line pytest plugin that
supplies the ``benchmark`` fixture of pytest-benchmark (not installed here: the three benchmark tests just call their
target).  No reference file is edited."""
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(REPO, "baseline", "_ref", "tests")

pytestmark = pytest.mark.gpu

ALIAS_INIT = '''
import importlib, sys
_pkg = importlib.import_module("azul_deep_reinforcement_learning_b200.azulnet")
for _sub in ("azul", "agent", "game_runner", "model", "nn_runner"):
    sys.modules["azulnet." + _sub] = importlib.import_module("azul_deep_reinforcement_learning_b200.azulnet." + _sub)
sys.modules["azulnet"] = _pkg
'''

PLUGIN = '''
import pytest

class _Benchmark:
    def __call__(self, fn, *args, **kwargs):
        return fn(*args, **kwargs)
    def pedantic(self, fn, args=(), kwargs=None, rounds=1, **_):
        out = None
        for _i in range(min(rounds, 2)):
            out = fn(*args, **(kwargs or {}))
        return out

@pytest.fixture
def benchmark():
    return _Benchmark()
'''


def test_reference_test_suite_passes_unchanged_on_the_facade(tmp_path):
    if not os.path.isdir(REF_TESTS):
        pytest.skip("baseline/_ref not staged (python __graft_entry__.py in the build container)")
    alias = tmp_path / "alias"
    (alias / "azulnet").mkdir(parents=True)
    (alias / "azulnet" / "__init__.py").write_text(ALIAS_INIT)
    (alias / "refsuite_plugin.py").write_text(PLUGIN)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(alias), REPO, os.environ.get("PYTHONPATH", "")]))
    out = subprocess.run([sys.executable, "-m", "pytest", REF_TESTS, "-q", "-p", "refsuite_plugin", "-p", "no:cacheprovider",
                          "--rootdir", str(tmp_path)], capture_output=True, text=True, timeout=1500, cwd=str(tmp_path), env=env)
    tail = out.stdout[-4000:] + out.stderr[-2000:]
    assert out.returncode == 0, tail
    assert " passed" in out.stdout and "failed" not in out.stdout.splitlines()[-1], tail
    n_passed = int(out.stdout.splitlines()[-1].split(" passed")[0].split()[-1])
    assert n_passed >= 28, tail
