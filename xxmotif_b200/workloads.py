"""The synthetic workloads of BASELINE.json's five configurations and of the cases the fast path of
round 1 skipped (more than 8 match columns, ragged lengths), as data + the reference's command-line flags.

bench.py times every one of them on the GPU next to a sample run of the unmodified reference, and tests/test_gpu_matrix.py
compares the two bit by bit.  Nothing here touches oracle/: running the reference is the caller's business.
"""
import numpy as np

from . import hostprep

# rows A, C, G, T in percent.  TATA: the README's 4 x 8 profile.  The longer ones append mildly informative columns, so that
# the motif has 12 / 17 match columns (two p-value tables, branch_and_bound-inl.h:190-193) while the planted TATAWAWR
# instances still carry it.
_MILD = np.array([[40, 20, 20, 20], [20, 40, 20, 20], [20, 20, 40, 20], [20, 20, 20, 40], [35, 15, 15, 35], [15, 35, 35, 15], [30, 30, 20, 20],
                  [20, 20, 30, 30], [40, 20, 20, 20]], dtype=np.float64).T


def profile(ncols):
    base = hostprep.TATA_PROFILE
    if ncols <= 8:
        return base[:, :ncols].copy()
    return np.concatenate([base, _MILD[:, :ncols - 8]], axis=1)


def profile_text(prof):
    return "".join(" ".join("%.2f" % (v / 100.0) for v in row) + "\n" for row in prof)


class Workload:
    def __init__(self, name, nseq, seqlen, seed, frac=0.5, planted="TATAWAWR", start=("m", "TATAWAWR"), revcomp=False, mops=False, localization=False,
                 order=2, neg=0, anchor=None, downstream=None, ragged=None, what=""):
        self.name, self.nseq, self.seqlen, self.seed, self.frac, self.planted = name, nseq, seqlen, seed, frac, planted
        self.start, self.revcomp, self.mops, self.localization, self.order = start, revcomp, mops, localization, order
        self.neg, self.anchor, self.downstream, self.ragged, self.what = neg, anchor, downstream, ragged, what

    def scaled(self, nseq, neg=None):
        w = Workload.__new__(Workload)
        w.__dict__.update(self.__dict__)
        w.nseq = nseq
        if neg is not None:
            w.neg = neg
        return w

    # ---- data
    def positive_set(self):
        """(n int32[nseq], codes uint8[sum n]) with codes 1..4, one planted instance in a fraction of the sequences"""
        if self.ragged is None and self.anchor is None:
            return hostprep.synthetic_set(self.nseq, self.seqlen, motif=self.planted, frac=self.frac, seed=self.seed)
        rng = np.random.default_rng(self.seed)
        m = len(self.planted)
        if self.ragged is None:
            lens = np.full(self.nseq, self.seqlen, dtype=np.int64)
        else:
            lens = rng.integers(self.ragged[0], self.ragged[1] + 1, size=self.nseq).astype(np.int64)
        off = np.concatenate(([0], np.cumsum(lens)))
        codes = rng.integers(1, 5, size=int(off[-1]), dtype=np.uint8)
        has = np.flatnonzero((rng.random(self.nseq) < self.frac) & (lens >= m))
        if self.anchor is None:
            pos = (rng.random(len(has)) * (lens[has] - m + 1)).astype(np.int64)
        else:  # planted position ~ Normal(anchor, 5), clipped to the sequence (cfg5: anchored sequences)
            pos = np.clip(np.rint(rng.normal(self.anchor, 5.0, size=len(has))).astype(np.int64), 0, lens[has] - m)
        for j, ch in enumerate(self.planted):
            alts = np.array(hostprep.IUPAC[ch], dtype=np.uint8)
            codes[off[has] + pos + j] = alts[rng.integers(0, len(alts), size=len(has))]
        return lens.astype(np.int32), codes

    def negative_set(self, n, codes):
        """`neg` sequences, each a mono-nucleotide shuffle of a positive sequence (cfg3: --negSet)"""
        if not self.neg:
            return None
        rng = np.random.default_rng(self.seed + 10)
        off = np.concatenate(([0], np.cumsum(n.astype(np.int64))))
        src = np.arange(self.neg) % len(n)
        lens = n[src]
        if self.ragged is None:
            L = int(n[0])
            rows = codes.reshape(len(n), L)[src]
            return lens.astype(np.int32), rng.permuted(rows, axis=1).reshape(-1)
        out = np.concatenate([rng.permutation(codes[off[s]:off[s + 1]]) for s in src])
        return lens.astype(np.int32), out

    # ---- start motif
    def start_pwm(self):
        """the 31 x 5 log PWM and match columns of a -p profile (Motif::InitProfMotif); -m starts are taken from the reference"""
        if self.start[0] != "p":
            return None
        return hostprep.profile_pwm(profile(self.start[1]))

    def reference_flags(self, tmpdir, negfa=None):
        """command-line flags of the reference for this workload (after OUTDIR SEQFILE)"""
        import os
        f = ["--batch"]
        if self.revcomp:
            f.append("--revcomp")
        if self.mops:
            f.append("--mops")
        if self.localization:
            f.append("--localization")
        if self.downstream is not None:
            f += ["--downstream", str(self.downstream)]
        if negfa:
            f += ["--negSet", negfa]
        if self.order != 2:
            f += ["--background-model-order", str(self.order)]
        if self.start[0] == "m":
            f += ["-m", self.start[1]]
        else:
            pf = os.path.join(tmpdir, "start.prof")
            with open(pf, "w") as fh:
                fh.write(profile_text(profile(self.start[1])))
            f += ["-p", pf]
        return f

    def describe(self):
        mode = "MOPS" if self.mops else "ZOOPS"
        start = ("-m %s" % self.start[1]) if self.start[0] == "m" else ("-p %d-column profile" % self.start[1])
        shape = ("%d-%d bp" % self.ragged) if self.ragged else ("%d bp" % self.seqlen)
        extra = ("%s%s%s" % (" --revcomp" if self.revcomp else "", " --localization --downstream %s" % self.downstream if self.localization else "",
                             " --negSet %d seqs" % self.neg if self.neg else ""))
        return "%s: %d x %s, %s, bg order %d,%s %s" % (self.name, self.nseq, shape, mode, self.order, extra, start)


def write_fasta(path, n, codes):
    lut = np.frombuffer(b"NACGT", dtype=np.uint8)
    off = np.concatenate(([0], np.cumsum(np.asarray(n, dtype=np.int64))))
    with open(path, "wb") as fh:
        if len(set(np.asarray(n).tolist())) == 1:
            rows = lut[codes.reshape(len(n), int(n[0]))]
            for i in range(len(n)):
                fh.write(b">seq%d\n" % (i + 1))
                fh.write(rows[i].tobytes())
                fh.write(b"\n")
        else:
            txt = lut[codes]
            for i in range(len(n)):
                fh.write(b">seq%d\n" % (i + 1))
                fh.write(txt[off[i]:off[i + 1]].tobytes())
                fh.write(b"\n")


# BASELINE.json configs[1..4] at their own sizes (configs[0], 500 x 200 bp, is the reference's CPU-runnable case and is covered by
# the golden traces) + the shapes the round-1 fast path skipped.  `scale` = sequences of the GPU-filling variant bench.py times.
MATRIX = [
    (Workload("cfg2", 10000, 500, 2, revcomp=True, mops=True, what="ChIP-seq peaks, --revcomp --mops"), 1000000),
    (Workload("cfg3", 100000, 200, 3, revcomp=True, order=8, neg=500000, what="peaks + shuffled negative set, bg order 8, --revcomp"), 2000000),
    (Workload("cfg4", 1000000, 1000, 4, start=("p", 8), what="-p TATA profile jump-start"), 1000000),
    (Workload("cfg5", 50000, 300, 5, frac=0.6, planted="TATAW.AWR", start=("m", "TATAW.AWR"), localization=True, anchor=170, downstream=100,
              what="--localization --downstream 100, gapped start motif"), 2000000),
    (Workload("cols12", 1000000, 1000, 4, start=("p", 12), what="12 match columns (two p-value tables) on the cfg4 set"), 1000000),
    (Workload("cols17", 1000000, 1000, 4, start=("p", 17), what="17 match columns on the cfg4 set"), 1000000),
    (Workload("ragged", 1000000, 1000, 6, start=("p", 8), ragged=(200, 1800), what="ragged lengths 200-1800 bp, -p TATA profile"), 1000000),
]


def by_name(name):
    for w, _ in MATRIX:
        if w.name == name:
            return w
    raise KeyError(name)
