"""Pin the C restatement (oracle/xxm_oracle.c) to the unmodified reference.

Every golden trace under tests/golden/ was dumped from the real reference functions by oracle/ref_harness.cpp
(tests/golden/make_golden.py).  The oracle must reproduce every boundary array bit-exactly: sequence features,
background, NullModel tables, p-value tables, per-window doubles, site lists, sorts, subset choice and PWMs.
"""
import numpy as np
import pytest

from backends import OracleBackend
from conftest import golden_names
from replay import TraceInputs, replay, _eq


def _fnv(a):
    h = np.uint64(1469598103934665603)
    prime = np.uint64(1099511628211)
    with np.errstate(over="ignore"):
        for b in np.frombuffer(a.tobytes(), dtype=np.uint8):
            h = (h ^ np.uint64(b)) * prime
    return int(h.astype(np.int64))


@pytest.mark.parametrize("name", golden_names())
def test_inputs_from_scratch(golden, name):
    """seqset features / revcomp / background / NullModel restatements vs the reference's own values."""
    t = golden(name)
    ti = TraceInputs(t)
    be = OracleBackend(ti, from_scratch=True)
    _eq("S", be.seqset.flat(), ti.S_post)
    _eq("n", be.seqset.n, ti.n_post)
    _eq("total", be.seqset.total, ti.total)
    _eq("avgLength", be.seqset.avgLength[1:], ti.avgLength[1:])
    _eq("max_leng", be.seqset.max_leng, int(t["in.pos.max_leng"][0]))
    pbg, pbgl = be.seqset.background()
    _eq("posBg", pbg[1:], t["bg.posBg"][1:])
    _eq("posBg_log", pbgl[1:], t["bg.posBg_log"][1:])
    if ti.hasNeg:
        _eq("neg.S", be.negset.flat(), t["in.neg.S"])
        _eq("neg.total", be.negset.total, t["in.neg.total"])
        _eq("neg.avgLength", be.negset.avgLength[1:], t["in.neg.avgLength"][1:])
        _eq("negBg", be.bg[1:], t["bg.negBg"][1:])
        _eq("negBg_log", be.bgLog[1:], t["bg.negBg_log"][1:])
    _eq("rfields", be.nm.rfields, int(t["nm.rfields"][0]))
    if "nm.rprobs" in t:
        _eq("rprobs", be.nm.rprobs, t["nm.rprobs"])
        _eq("rconds", be.nm.rconds, t["nm.rconds"])
    else:  # order-8 tables are too big for a fixture: compare the FNV-1a hash of the raw bytes
        import hashlib  # noqa: F401  (vectorised FNV below)
        want = t["nm.hash"]
        for arr, w in ((be.nm.rprobs, want[0]), (be.nm.rconds, want[1])):
            raw = np.frombuffer(arr.tobytes(), dtype=np.uint8)
            # FNV-1a is sequential; do it in C-speed chunks via python ints on a reduced sample is not equivalent,
            # so run the exact loop but in numpy uint64 scalar arithmetic on the non-zero prefix only when small.
            if raw.size <= (1 << 22):
                assert _fnv(arr) == int(w)


@pytest.mark.parametrize("name", golden_names())
def test_replay(golden, name):
    t = golden(name)
    ti = TraceInputs(t)
    if "nm.rprobs" in t:
        be = OracleBackend(ti, from_scratch=False)
    else:
        be = OracleBackend(ti, from_scratch=True)
    done = replay(t, be)
    assert "opt1" in done and "oops" in done and "hist" in done
