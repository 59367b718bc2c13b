"""GPU parity, part 1: every golden boundary trace of the unmodified reference replayed through the C ABI of
libxxmotif_gpu.so on a real B200.  Bit-exact: sites, strands (positions in the doubled sequence), counts, float
scores, double p-values, PWMs."""
import pytest

from conftest import golden_names
from replay import TraceInputs, replay

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names())
def test_replay_gpu(golden, name):
    from gpu_backend import GpuBackend
    t = golden(name)
    ti = TraceInputs(t)
    be = GpuBackend(ti)
    try:
        done = replay(t, be)
    finally:
        be.ctx.close()
    assert "opt1" in done and "oops" in done and "hist" in done


@pytest.mark.parametrize("name", golden_names())
def test_seqset_features_gpu(golden, name):
    """a1: total[] / avgLength[] / max_leng of the loaded set (after --revcomp doubling) against the values the unmodified
    reference dumped (seqset.cpp:21-36 calculateSequenceFeatures, seqset.cpp:128-178 createRevcomp), directly."""
    import numpy as np
    from gpu_backend import GpuBackend
    t = golden(name)
    ti = TraceInputs(t)
    be = GpuBackend(ti)
    try:
        total, avg, mxl, _ = be.ctx.seqset_info()
    finally:
        be.ctx.close()
    assert np.array_equal(np.asarray(total, dtype=np.int64), np.asarray(ti.total, dtype=np.int64).reshape(-1))
    assert np.array_equal(np.asarray(avg)[1:], np.asarray(ti.avgLength, dtype=np.float64).reshape(-1)[1:])
    assert mxl == int(t["in.pos.max_leng"][0])
