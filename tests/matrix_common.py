"""Shared pieces of the workload-matrix parity tests (tests/test_matrix_oracle.py on the CPU, tests/test_gpu_matrix.py on the
GPU): run the UNMODIFIED reference (oracle/_ref/ref_harness bench = K passes of the optimizeMotif loop body,
Iterate_Motif.cpp:26-32) on one of xxmotif_b200.workloads and hand back what it ended on."""
import json
import os
import shutil
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


def have_reference():
    return os.path.exists(HARNESS)


def run_reference(w, steps, n=None, codes=None, neg=None):
    """K optimizeMotif steps of the reference on workload `w`; returns (json, n, codes, neg) -- neg = (n, codes) or None."""
    from xxmotif_b200 import workloads as W
    tmp = tempfile.mkdtemp(prefix="xxm_matrix_")
    try:
        if n is None:
            n, codes = w.positive_set()
            neg = w.negative_set(n, codes)
        fa = os.path.join(tmp, "in.fa")
        W.write_fasta(fa, n, codes)
        negfa = None
        if neg is not None:
            negfa = os.path.join(tmp, "neg.fa")
            W.write_fasta(negfa, neg[0], neg[1])
        cmd = [HARNESS, "bench", str(steps), "0", os.path.join(tmp, "out"), fa] + w.reference_flags(tmp, negfa)
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        lines = [x for x in r.stdout.decode().splitlines() if x.startswith("{")]
        if r.returncode != 0 or not lines:
            raise RuntimeError("ref_harness failed for %s: %s" % (w.name, r.stderr.decode()[-1500:]))
        return json.loads(lines[-1]), n, codes, neg
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def background_inputs(w, n, codes, neg):
    """mono-nucleotide background + NullModel tables the way the reference derives them: from the negative set when there is
    one, else from the positive set; with --revcomp from the doubled sequences.  Host entry points of the
    library (pinned against the reference's dumps by the golden traces)."""
    from xxmotif_b200 import hostprep
    bn, bc = neg if neg is not None else (n, codes)
    if w.revcomp:
        bn, bc = hostprep.revcomp_double(bn, bc)
    bg, bgl = hostprep.background(bn, bc)
    rp, rc = hostprep.null_model(bn, bc, order=w.order)
    return bg, bgl, rp, rc


def compare_with_reference(ref, got):
    """`got`: dict with the same keys as the harness line (nbSites, nbBestSites, maxSites, bestMotifNb, bestPval, pwm, enr)"""
    for k in ("nbSites", "nbBestSites", "maxSites", "bestMotifNb"):
        assert int(got[k]) == int(ref[k]), "%s: %s != reference %s" % (k, got[k], ref[k])
    assert float(got["bestPval"]) == float(ref["bestPval"]), "bestPval: %r != reference %r" % (got["bestPval"], ref["bestPval"])
    a = np.asarray(got["pwm"], dtype=np.float64).reshape(31, 5)[1:, 1:]
    b = np.asarray(ref["pwm"], dtype=np.float64).reshape(31, 5)[1:, 1:]
    bad = np.flatnonzero((a != b).reshape(-1))
    assert bad.size == 0, "PWM differs from the reference in %d entries (first %s: %r vs %r)" % (bad.size, bad[:3], a.reshape(-1)[bad[:3]], b.reshape(-1)[bad[:3]])
    if "enr" in got:
        assert [int(x) for x in got["enr"]] == [int(x) for x in ref["enr"]], "enriched region %s != reference %s" % (got["enr"], ref["enr"])
