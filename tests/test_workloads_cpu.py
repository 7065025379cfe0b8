"""CPU: the drivers' mt19937(0) input streams as the product package draws them (numpy) against the oracle's restatement,
which tests/test_oracle_golden.py pins to the reference's own binaries and checked-in results."""
import json
import os

import numpy as np

from rapidquadcoptercollisiondetection_b200 import workloads as W


def test_c1_trial_stream_equals_the_oracles(oracle):
    bc, sph, trials = oracle.c1_inputs(4000, filter=False)
    assert trials == 4000
    b2, s2 = W.c1_trials(4000)
    np.testing.assert_array_equal(bc, b2)
    np.testing.assert_array_equal(sph, s2)
    # and the accepted ones are a subsequence of it
    acc, sph_a, t = oracle.c1_inputs(1000, filter=True)
    full, _ = W.c1_trials(t)
    keep = oracle.input_feasibility(full) == 0
    assert int(keep.sum()) == 1000
    np.testing.assert_array_equal(full[:, keep], acc)


def test_c2_candidate_stream_equals_the_oracles(oracle):
    np.testing.assert_array_equal(oracle.c2_inputs(23, 100), W.c2_candidates(23, 100))


def test_golden_counts_record_the_trial_count():
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_counts.json")))
    assert g["c1_trials"] == 15_571_728 and g["c1_n"] == 10_000_000
