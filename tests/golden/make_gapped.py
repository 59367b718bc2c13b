#!/usr/bin/env python
"""Golden vectors for SURVEY.md 8(f) row 4: the UNMODIFIED reference's ElongCandidates::countAllOccurrences
(elongationCandidates.cpp:234-262) run by oracle/_ref/ref_harness (mode `gapped`) on one small ragged set with N runs, for every
gap pattern of the three seed types, with and without --revcomp.  Writes tests/golden/gapped_counts.npz (input set + counts).

    make -C oracle ref && python tests/golden/make_gapped.py
"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from xxmotif_b200 import workloads as W  # noqa: E402

CASES = [("five", False), ("five", True), ("pal", False), ("pal", True), ("tandem", True)]


def input_set():
    w = W.Workload("g", 160, 0, 77, ragged=(9, 260))
    n, codes = w.positive_set()
    codes = codes.copy()
    rng = np.random.default_rng(5)
    codes[rng.random(codes.size) < 0.02] = 0
    return n, codes


def run_reference(kind, revcomp, fasta, tmp):
    out = os.path.join(tmp, "g.bin")
    cmd = [os.path.join(ROOT, "oracle", "_ref", "ref_harness"), "gapped", kind, out, os.path.join(tmp, "out"), fasta] + (["--revcomp"] if revcomp else [])
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    b = open(out, "rb").read()
    nc, npat, half = struct.unpack("iii", b[:12])
    body = np.frombuffer(b[20:], dtype=np.int32).reshape(npat, nc + 4 ** nc)
    return half, body[:, :nc].copy(), body[:, nc:].astype(np.uint32)


def main():
    n, codes = input_set()
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "out"))
    fa = os.path.join(tmp, "g.fa")
    W.write_fasta(fa, n, codes)
    out = {"n": n.astype(np.int32), "codes": codes.astype(np.uint8)}
    for kind, rc in CASES:
        half, off, cnt = run_reference(kind, rc, fa, tmp)
        key = "%s_%s" % (kind, "rc" if rc else "fw")
        out[key + "_half"] = np.int32(half)
        out[key + "_offsets"] = off
        out[key + "_counts"] = cnt
        print(key, "firstHalfOnly", half, "patterns", len(off), "counted", cnt.sum(axis=1)[:4])
    np.savez_compressed(os.path.join(HERE, "gapped_counts.npz"), **out)


if __name__ == "__main__":
    main()
