"""bench.py's reference arm (the C oracle port on the host cores) runs without a GPU and prints one JSON line
with the contract's keys."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--k-steps", "64", "--pyref-seconds", "0"], capture_output=True, text=True, timeout=300, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "env_steps/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and line["vs_baseline"] is None


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--k-steps", "64"], capture_output=True, text=True, timeout=120, cwd=REPO, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_python_reference_leg_when_the_reference_is_staged():
    """cpu_baseline.python_reference: the UNMODIFIED reference's GameRunner loop from baseline/_ref (staged by build());
    reports 'unavailable' instead of failing where the tree was not staged."""
    import pytest
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--k-steps", "64", "--pyref-seconds", "2"], capture_output=True, text=True, timeout=300, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    py = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]["python_reference"]
    if not os.path.isdir(os.path.join(REPO, "baseline", "_ref", "azulnet")):
        assert "unavailable" in py
        pytest.skip("baseline/_ref not staged")
    assert py["kind"] == "reference" and py["cores"] >= 1 and py["games"] > 0 and py["value"] > 100
    assert py["single_core_value"] > 100 and py["unit"] == "env_steps/s" and "cpu_model" in py
