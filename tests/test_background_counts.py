"""CPU: the NullModel split into (k-mer counts) + (model from counts) -- the seam the device histogram plugs into
 -- gives bit-identical tables to the one-piece host path that is pinned against the reference."""
import numpy as np
import pytest


def _ragged(seed, nseq=300, lo=1, hi=80, pn=0.1):
    rng = np.random.default_rng(seed)
    n = rng.integers(lo, hi, size=nseq).astype(np.int32)
    codes = rng.integers(1, 5, size=int(n.sum())).astype(np.uint8)
    codes[rng.random(codes.size) < pn] = 0
    return n, codes


@pytest.mark.parametrize("order", [0, 1, 2, 3, 5])
def test_counts_then_model_equals_host_null_model(order):
    from xxmotif_b200 import hostprep
    n, codes = _ragged(order)
    c = hostprep.kmer_counts_host(n, codes, order)
    # mono-nucleotide counts are plain symbol counts; every longer k-mer count is bounded by its prefix count
    assert [int(c[a]) for a in range(1, 5)] == [int((codes == a).sum()) for a in range(1, 5)]
    rp, rc = hostprep.null_model_from_counts(c, order)
    rp2, rc2 = hostprep.null_model(n, codes, order)
    assert (rp == rp2).all() and (rc == rc2).all()


def test_counts_do_not_cross_sequences_or_n():
    from xxmotif_b200 import hostprep
    n = np.array([3, 2, 4], dtype=np.int32)
    codes = np.array([1, 2, 3, 4, 4, 1, 0, 2, 2], dtype=np.uint8)  # ACG | TT | A N C C
    c = hostprep.kmer_counts_host(n, codes, 2)
    idx = lambda *d: int(sum(x * 5 ** (len(d) - 1 - i) for i, x in enumerate(d)))
    assert c[idx(1, 2)] == 1 and c[idx(2, 3)] == 1 and c[idx(3, 4)] == 0 and c[idx(4, 4)] == 1 and c[idx(4, 1)] == 0
    assert c[idx(1, 2, 3)] == 1 and c[idx(2, 2)] == 1 and c[idx(1, 2, 2)] == 0 and int(c[1:5].sum()) == 8
