"""GPU: the reference's two Monte Carlo drivers rewritten against the C++ facade (examples/) print the reference's own
checked-in results. The programs draw their inputs themselves with std::mt19937 (not through the oracle), so this is
an end-to-end check of the C++ host path: facade -> C ABI -> CUDA."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run(exe, *args):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples")], stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(ROOT, "examples", exe), *args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout, r.stderr)
    return r.stdout


def _value(out, label):
    return re.search(re.escape(label) + r"\s*=\s*(\S+)", out).group(1)


def test_forest_performance_eval_prints_the_reference_percentages(golden_counts):
    out = _run("forest_performance_eval")  # 10^4 batches x 100 candidates, as in the reference
    # data/ForestPerfEval.txt:28-29
    assert _value(out, "Percent of trajectories that were input feasible") == "63.9533"
    assert _value(out, "Percent of trajectories that were collision free") == "60.381"


def test_performance_eval_prints_the_reference_percentages():
    out = _run("performance_eval")  # 10^7 accepted trials, as in the reference
    # data/PerformanceEval.txt:23-25
    assert _value(out, "Percent of trajectories that were collision-free ") == "96.0038"
    assert _value(out, "Percent of trajectories with collision ") == "3.99612"
    assert _value(out, "Percent of trajectories with collision indeterminable ") == "8e-05"
