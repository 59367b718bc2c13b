"""The drop-in boundary, linked: oracle/_ref/XXmotif_bound is the UNMODIFIED reference compiled with the reference-side binding
of libxxmotif_gpu.so (integration/xxmotif_gpu_binding.{h,cpp}: the four window loops of StartPosUpdater replaced by calls
through the C ABI; everything else -- command line, refinement control flow, getBestSubset, the output writers -- is the
reference's own code).  Same command line, same input: every output file (_MotifFile.txt, _Pvals.txt, .pwm, .meme,
_sequence.txt, tmp/*.dat) must be byte-identical to what oracle/_ref/XXmotif (the plain reference) writes."""
import filecmp
import os
import shutil
import subprocess
import tempfile

import pytest

import matrix_common as mc
from xxmotif_b200 import workloads as W

pytestmark = pytest.mark.gpu

REF = os.path.join(mc.ROOT, "oracle", "_ref", "XXmotif")
BOUND = os.path.join(mc.ROOT, "oracle", "_ref", "XXmotif_bound")

CASES = [
    ("zoops_m", W.Workload("a", 2000, 300, 31, frac=0.6), []),
    ("mops_revcomp", W.Workload("b", 1500, 400, 32, revcomp=True, mops=True), []),
    ("profile", W.Workload("c", 3000, 500, 33, start=("p", 8)), []),
    ("localization_gapped", W.Workload("d", 3000, 300, 34, frac=0.6, planted="TATAW.AWR", start=("m", "TATAW.AWR"), localization=True, anchor=170,
                                       downstream=100), []),
    ("negset_order4_revcomp", W.Workload("e", 2000, 200, 35, revcomp=True, order=4, neg=6000), []),
    ("profile_12_columns", W.Workload("f", 2000, 600, 36, start=("p", 12)), []),
    ("oops", W.Workload("g", 1500, 250, 37, frac=0.9), ["--oops"]),
    ("ragged", W.Workload("h", 2000, 0, 38, ragged=(60, 700), start=("p", 8)), []),
]


def _files(root):
    out = []
    for d, _, fs in os.walk(root):
        for f in fs:
            if f.endswith(".R"):  # the plotting script embeds its own output directory
                continue
            out.append(os.path.relpath(os.path.join(d, f), root))
    return sorted(out)


def _compare_with_reference(w, extra, env=None):
    if not (os.path.exists(REF) and os.path.exists(BOUND)):
        pytest.skip("oracle/_ref/XXmotif or XXmotif_bound not built (make -C oracle ref bound)")
    tmp = tempfile.mkdtemp(prefix="xxm_bound_")
    try:
        n, codes = w.positive_set()
        fa = os.path.join(tmp, "in.fa")
        W.write_fasta(fa, n, codes)
        neg = w.negative_set(n, codes)
        negfa = None
        if neg is not None:
            negfa = os.path.join(tmp, "neg.fa")
            W.write_fasta(negfa, neg[0], neg[1])
        flags = w.reference_flags(tmp, negfa) + extra
        outs = {}
        for tag, exe in (("ref", REF), ("gpu", BOUND)):
            out = os.path.join(tmp, tag)
            r = subprocess.run([exe, out, fa] + flags, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, **(env or {})))
            assert r.returncode == 0, "%s failed: %s" % (exe, r.stderr.decode()[-1500:])
            outs[tag] = out
        fr, fg = _files(outs["ref"]), _files(outs["gpu"])
        assert fr == fg and any(f.endswith("_MotifFile.txt") for f in fr) and any(f.endswith("_Pvals.txt") for f in fr)
        assert os.path.getsize(os.path.join(outs["ref"], [f for f in fr if f.endswith("_Pvals.txt")][0])) > 200
        diff = [f for f in fr if not filecmp.cmp(os.path.join(outs["ref"], f), os.path.join(outs["gpu"], f), shallow=False)]
        assert not diff, "files differ from the reference's: %s" % diff
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


@pytest.mark.parametrize("name,w,extra", CASES, ids=[c[0] for c in CASES])
def test_bound_binary_writes_the_reference_files(name, w, extra):
    _compare_with_reference(w, extra)


@pytest.mark.parametrize("name,w,extra", CASES, ids=[c[0] for c in CASES])
def test_bound_binary_on_several_devices(name, w, extra):
    """XXM_DEVICES=N (all GPUs of the box): the reference's single host thread drives them through xxm_multi (the set cut into
    N blocks of sequences, the counts of the whole set on each, lists concatenated / merged) -- the files stay the reference's."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs (gpurun --gpus 2)")
    _compare_with_reference(w, extra, {"XXM_DEVICES": str(min(8, torch.cuda.device_count()))})


@pytest.mark.parametrize("flags", [[], ["--type", "PALINDROME", "--revcomp"], ["--type", "TANDEM"]], ids=["all_seed_types", "palindrome_revcomp", "tandem"])
def test_bound_binary_full_discovery(flags):
    """No start motif: the reference's whole pipeline -- seed discovery, elongation, merging, refinement of ~150 motifs,
    ranking, output -- with the counting loop of the seed search (ElongCandidates::countAllOccurrences) and
    every window loop of the refinement path served by the library (tens of thousands of small calls through the binding).
    ~2 x 1 minute of host work per case, so only with XXM_BOUND_FULL=1 (profiles/ holds the record of a run)."""
    if not os.environ.get("XXM_BOUND_FULL"):
        pytest.skip("set XXM_BOUND_FULL=1 to run the full-discovery comparison (minutes of CPU seed search)")
    if not (os.path.exists(REF) and os.path.exists(BOUND)):
        pytest.skip("oracle/_ref/XXmotif or XXmotif_bound not built")
    import time
    tmp = tempfile.mkdtemp(prefix="xxm_bound_full_")
    try:
        n, codes = W.Workload("s", 80, 100, 1, frac=0.8).positive_set()
        fa = os.path.join(tmp, "s.fa")
        W.write_fasta(fa, n, codes)
        secs = {}
        for tag, exe in (("ref", REF), ("gpu", BOUND)):
            t0 = time.time()
            r = subprocess.run([exe, os.path.join(tmp, tag), fa, "--batch"] + flags, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            secs[tag] = time.time() - t0
            assert r.returncode == 0, r.stderr.decode()[-1500:]
        fr, fg = _files(os.path.join(tmp, "ref")), _files(os.path.join(tmp, "gpu"))
        assert fr == fg
        diff = [f for f in fr if not filecmp.cmp(os.path.join(tmp, "ref", f), os.path.join(tmp, "gpu", f), shallow=False)]
        motifs = open(os.path.join(tmp, "ref", "s_MotifFile.txt")).read().count("Motif Number")
        print("full discovery %s: %d motifs reported, reference %.1f s, bound %.1f s, differing files: %s" % (" ".join(flags) or "(defaults)", motifs, secs["ref"], secs["gpu"], diff))
        assert not diff and motifs > 10
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
