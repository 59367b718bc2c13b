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
