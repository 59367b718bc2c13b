"""Writes past the end of a device buffer: every buffer of the library gets a 256-byte canary (XXM_DEBUG_GUARDS) and the golden
replays, the ragged / low-complexity stress case and small matrix workloads run in a fresh process (tools/guard_run.py); no
canary may be touched.  (compute-sanitizer is closed on this GPU pool; this covers the part of memcheck that matters most for
kernels with vectorised stores near buffer ends.)"""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_canary_overwritten():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "guard_run.py")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    lines = [x for x in r.stdout.decode().splitlines() if x.startswith("{")]
    assert lines, "guard run produced no report: %s" % r.stderr.decode()[-2000:]
    rep = json.loads(lines[-1])
    assert rep["checked"] > 500, rep
    assert rep["damaged"] == 0, [c for c in rep["cases"] if c["damaged"]]
    assert r.returncode == 0
