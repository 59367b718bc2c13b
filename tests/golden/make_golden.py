#!/usr/bin/env python
"""Generate the golden boundary traces under tests/golden/ from the UNMODIFIED reference.

Runs oracle/_ref/ref_harness (the reference compiled in place from /root/reference, see oracle/Makefile)
on small seeded synthetic inputs that mirror BASELINE.json's five configs plus the edge cases the path has
(N runs, ragged lengths, gapped columns, >8 and >17 match columns, OOPS, localization, negative set).
Only this script and the harness touch the reference; the committed *.trace.gz files are what tests read.

    python tests/golden/make_golden.py            # regenerate everything
"""
import gzip
import os
import random
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")

IUPAC = {"A": "A", "C": "C", "G": "G", "T": "T", "W": "AT", "R": "AG", "Y": "CT", "S": "CG", "K": "GT", "M": "AC",
         "N": "ACGT", ".": "ACGT"}
TATA_PROF = """0.05 0.90 0.05 0.90 0.45 0.90 0.50 0.45
0.05 0.03 0.05 0.03 0.05 0.03 0.05 0.05
0.05 0.03 0.05 0.03 0.05 0.03 0.05 0.45
0.85 0.04 0.85 0.04 0.45 0.04 0.40 0.05
"""


def synth(rng, nseq, length, motif, frac, ragged=None, nruns=0, anchor=None):
    seqs = []
    for _ in range(nseq):
        L = length if ragged is None else rng.randint(ragged[0], ragged[1])
        s = [rng.choice("ACGT") for _ in range(L)]
        if rng.random() < frac and L >= len(motif):
            inst = [rng.choice(IUPAC[c]) for c in motif]
            if anchor is None:
                p = rng.randrange(0, L - len(motif) + 1)
            else:
                p = min(max(0, int(rng.gauss(anchor, 3))), L - len(motif))
            s[p:p + len(motif)] = inst
        for _ in range(nruns):
            if rng.random() < 0.3:
                p = rng.randrange(0, L)
                for q in range(p, min(L, p + rng.randint(1, 6))):
                    s[q] = "N"
        seqs.append("".join(s))
    return seqs


def write_fasta(path, seqs):
    with open(path, "w") as fh:
        for i, s in enumerate(seqs):
            fh.write(">seq%d\n%s\n" % (i + 1, s))


CASES = [
    # name, seed, nseq, length, planted motif, fraction, extra kwargs, harness flags
    ("c1_zoops", 1, 60, 120, "TATAWAWR", 0.7, {}, ["-m", "TATAWAWR"]),
    ("c2_mops_rc", 2, 40, 150, "TATAWAWR", 0.5, {}, ["--revcomp", "--mops", "-m", "TATAWAWR"]),
    ("c3_neg_o4_rc", 3, 50, 100, "TATAWAWR", 0.5, {"neg": 150, "order": 4}, ["--revcomp", "-m", "TATAWAWR"]),
    ("c4_prof", 4, 80, 250, "TATAWAWR", 0.5, {"prof": True}, []),
    ("c5_loc_gap", 5, 60, 150, "TATAW.AWR", 0.6, {"anchor": 90}, ["--localization", "--downstream", "50", "-m", "TATAW.AWR"]),
    ("c5b_loc_mops_rc", 15, 50, 120, "TATAWAWR", 0.6, {"anchor": 60},
     ["--localization", "--downstream", "40", "--mops", "--revcomp", "-m", "TATAWAWR"]),
    ("c6_long12_mops", 6, 60, 150, "TATAWAWRGCGC", 0.6, {}, ["--mops", "-m", "TATAWAWRGCGC"]),
    ("c7_oops", 7, 40, 100, "TATAWAWR", 0.8, {}, ["--oops", "-m", "TATAWAWR"]),
    ("c8_ragged_n", 8, 50, 0, "TATAWAWR", 0.6, {"ragged": (30, 200), "nruns": 3}, ["-m", "TATAWAWR"]),
    ("c8b_ragged_n_mops_rc", 18, 40, 0, "TATAWAWR", 0.6, {"ragged": (30, 160), "nruns": 3}, ["--mops", "--revcomp", "-m", "TATAWAWR"]),
    ("c9_neg_o8", 9, 40, 100, "TATAWAWR", 0.5, {"neg": 120}, ["-m", "TATAWAWR"]),
    ("c10_long18", 10, 50, 160, "TATAWAWRGCGCATATWW", 0.7, {}, ["-m", "TATAWAWRGCGCATATWW"]),
    ("c11_short5", 11, 40, 100, "TATAW", 0.8, {}, ["--mops", "-m", "TATAW"]),
]


def main():
    if not os.path.exists(HARNESS):
        sys.exit("build the reference harness first: make -C oracle ref")
    only = set(sys.argv[1:])
    for name, seed, nseq, length, motif, frac, kw, flags in CASES:
        if only and name not in only:
            continue
        rng = random.Random(seed)
        tmp = tempfile.mkdtemp(prefix="xxm_golden_")
        try:
            pos = synth(rng, nseq, length, motif, frac, ragged=kw.get("ragged"), nruns=kw.get("nruns", 0), anchor=kw.get("anchor"))
            fa = os.path.join(tmp, "in.fa")
            write_fasta(fa, pos)
            args = []
            if "neg" in kw:
                neg = []
                for i in range(kw["neg"]):
                    s = list(pos[i % len(pos)])
                    rng.shuffle(s)
                    neg.append("".join(s))
                nfa = os.path.join(tmp, "neg.fa")
                write_fasta(nfa, neg)
                args += ["--negSet", nfa]
                if "order" in kw:
                    args += ["--background-model-order", str(kw["order"])]
            if kw.get("prof"):
                pf = os.path.join(tmp, "tata.prof")
                with open(pf, "w") as fh:
                    fh.write(TATA_PROF)
                args += ["-p", pf]
            trace = os.path.join(tmp, "trace.bin")
            cmd = [HARNESS, "trace", trace, os.path.join(tmp, "out"), fa, "--batch"] + args + flags
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            if r.returncode != 0 or not os.path.exists(trace):
                sys.stderr.write(r.stderr.decode()[-2000:])
                sys.exit("harness failed for %s (rc=%d)" % (name, r.returncode))
            out = os.path.join(HERE, name + ".trace.gz")
            with open(trace, "rb") as fi, gzip.GzipFile(out, "wb", compresslevel=9, mtime=0) as fo:
                shutil.copyfileobj(fi, fo)
            print("%-24s %8d bytes" % (name, os.path.getsize(out)))
        finally:
            shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
