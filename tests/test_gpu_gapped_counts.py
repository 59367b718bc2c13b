"""SURVEY.md 8(f) row 4 (first half) on the GPU: xxm_count_gapped_kmers against the reference's golden vectors (every gap
pattern of the three seed types, +/- --revcomp, ragged set with N runs), against the oracle on larger sets, on a prefix of the
set, through the multi-device layer, and at the size of the metric through a size-independent property."""
import os

import numpy as np
import pytest

import oracle_lib as O
from test_gapped_counts import CASES, load_golden
from xxmotif_b200 import capi, hostprep
from xxmotif_b200 import workloads as W

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CASES)
def test_device_counts_match_the_reference(case):
    g, n, codes = load_golden()
    ctx = capi.Context(0)
    try:
        ctx.load_seqset(n, codes, revcomp=case.endswith("_rc"))
        half = int(g[case + "_half"])
        for off, want in zip(g[case + "_offsets"], g[case + "_counts"]):
            assert np.array_equal(ctx.count_gapped_kmers(len(n), off, half), want), (case, off)
    finally:
        ctx.close()


@pytest.mark.parametrize("revcomp", [False, True])
def test_device_counts_match_the_oracle_on_a_larger_ragged_set(revcomp):
    w = W.Workload("g", 20000, 0, 91, ragged=(5, 700))
    n, codes = w.positive_set()
    codes = codes.copy()
    rng = np.random.default_rng(3)
    codes[rng.random(codes.size) < 0.005] = 0
    codes[1000:1400] = 0  # a run of N across several sequences
    s = O.SeqSet(n, codes, revcomp=revcomp)
    ctx = capi.Context(0)
    try:
        ctx.load_seqset(n, codes, revcomp=revcomp)
        for off, half in (([0, 1, 2, 3, 4], 0), ([0, 2, 3, 5, 7], 0), ([0, 1, 2, 18, 19, 20], 1 if revcomp else 0), ([0, 1, 2, 3, 4, 5], 0),
                          ([0, 4, 8, 12, 16, 20, 31], 0)):
            for nseq in (len(n), 7777, 1):
                assert np.array_equal(ctx.count_gapped_kmers(nseq, off, half), s.count_gapped_kmers(nseq, off, half)), (off, half, nseq)
    finally:
        ctx.close()


def test_multi_device_layer_adds_the_shards():
    import torch
    g, n, codes = load_golden()
    nd = min(8, torch.cuda.device_count())
    many = capi.MultiContext(nd)
    try:
        many.load_seqset(n, codes, revcomp=True)
        for case in ("five_rc", "pal_rc"):
            half = int(g[case + "_half"])
            for off, want in list(zip(g[case + "_offsets"], g[case + "_counts"]))[::5]:
                assert np.array_equal(many.count_gapped_kmers(len(n), off, half), want), (case, off, nd)
    finally:
        many.close()


def test_counts_at_the_size_of_the_metric():
    """1M x 1 kb without N: every start position with the whole pattern inside its sequence is counted exactly once, and the
    counts of a pattern are the counts of its mirrored pattern on the reversed k-mer ids only if ... (not assumed); the sum is
    the property.  Also prints the device time of one pattern."""
    nseq, length = 1000000, 1000
    n, codes = hostprep.synthetic_set(nseq, length)
    ctx = capi.Context(0)
    try:
        ctx.load_seqset(n, codes)
        for off in ([0, 1, 2, 3, 4], [0, 1, 3, 5, 7], [0, 1, 2, 13, 14, 15]):
            c = ctx.count_gapped_kmers(nseq, off, 0)
            assert int(c.astype(np.int64).sum()) == nseq * (length - off[-1])
            # a 10 000-sequence prefix against the oracle
        s = O.SeqSet(n[:10000], codes[:10000 * length])
        assert np.array_equal(ctx.count_gapped_kmers(10000, [0, 1, 3, 5, 7], 0), s.count_gapped_kmers(10000, [0, 1, 3, 5, 7], 0))
    finally:
        ctx.close()
