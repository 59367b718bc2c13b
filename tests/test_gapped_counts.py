"""SURVEY.md 8(f) row 4 (first half): the oracle's restatement of ElongCandidates::countAllOccurrences' counting loop against
golden vectors dumped from the UNMODIFIED reference (tests/golden/make_gapped.py -> gapped_counts.npz): every gap pattern of
the three seed types, with and without --revcomp, on a ragged set with N runs."""
import os

import numpy as np
import pytest

import oracle_lib as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gapped_counts.npz")
CASES = ["five_fw", "five_rc", "pal_fw", "pal_rc", "tandem_rc"]


def load_golden():
    g = np.load(GOLD)
    return g, g["n"], g["codes"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_counts_match_the_reference(case):
    g, n, codes = load_golden()
    s = O.SeqSet(n, codes, revcomp=case.endswith("_rc"))
    half = int(g[case + "_half"])
    assert half == (1 if case == "pal_rc" else 0)
    for off, want in zip(g[case + "_offsets"], g[case + "_counts"]):
        got = s.count_gapped_kmers(len(n), off, half)
        assert np.array_equal(got, want), (case, off)
    assert g[case + "_counts"].sum() > 10000
