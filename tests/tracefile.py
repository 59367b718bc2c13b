"""Reader for the tagged binary traces written by oracle/ref_harness.cpp (golden vectors)."""
import gzip
import struct
from collections import OrderedDict

import numpy as np

_DT = {b"i": np.int32, b"l": np.int64, b"f": np.float32, b"d": np.float64, b"b": np.uint8}


def read_trace(path):
    """Return an OrderedDict name -> numpy array (scalars are 1-element arrays)."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as fh:
        buf = fh.read()
    assert buf[:8] == b"XXMTRC1\n", "not a trace file"
    off = 8
    out = OrderedDict()
    while off < len(buf):
        (ln,) = struct.unpack_from("<I", buf, off)
        off += 4
        name = buf[off:off + ln].decode()
        off += ln
        dt = _DT[buf[off:off + 1]]
        off += 1
        (cnt,) = struct.unpack_from("<Q", buf, off)
        off += 8
        nbytes = cnt * np.dtype(dt).itemsize
        arr = np.frombuffer(buf, dtype=dt, count=cnt, offset=off).copy()
        off += nbytes
        assert name not in out, "duplicate record " + name
        out[name] = arr
    return out
