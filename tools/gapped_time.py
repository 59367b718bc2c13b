#!/usr/bin/env python
"""Device time of xxm_count_gapped_kmers on a set of the metric's size, every gap pattern of the reference's
FIVEMERS seed type (35 patterns, elongationCandidates.cpp:101-115) and of the palindromic types (16 patterns, 6 characters),
beside the oracle's CPU loop on a sample:   python tools/gapped_time.py 1000000 1000
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402


def main():
    nseq, length = int(sys.argv[1]), int(sys.argv[2])
    from xxmotif_b200 import capi, hostprep
    import oracle_lib as O
    n, codes = hostprep.synthetic_set(nseq, length)
    g = np.load(os.path.join(ROOT, "tests", "golden", "gapped_counts.npz"))
    ctx = capi.Context(0)
    ctx.load_seqset(n, codes)
    out = {"what": "xxm_count_gapped_kmers on %d x %d" % (nseq, length), "peak_GBs": 6556.5}
    for key in ("five_fw", "pal_fw"):
        ms = []
        for off in g[key + "_offsets"]:
            ctx.count_gapped_kmers(nseq, off, 0)
            ms.append(hostprep.kmer_counts_ms(ctx))
        t0 = time.perf_counter()
        for off in g[key + "_offsets"]:
            ctx.count_gapped_kmers(nseq, off, 0)
        wall = time.perf_counter() - t0
        ms = np.array(ms)
        alg = 0.375 * nseq * length
        out[key] = {"patterns": len(ms), "kernel_ms_min": float(ms.min()), "kernel_ms_max": float(ms.max()), "kernel_ms_sum": float(ms.sum()),
                    "wall_s_all_patterns": wall, "GBs_at_0.375B_per_base": alg / (ms.mean() / 1e3) / 1e9, "frac_of_peak": alg / (ms.mean() / 1e3) / 1e9 / 6556.5}
    sample = 20000
    s = O.SeqSet(n[:sample], codes[:sample * length])
    t0 = time.perf_counter()
    for off in g["five_fw_offsets"][:5]:
        s.count_gapped_kmers(sample, off, 0)
    cpu = (time.perf_counter() - t0) / 5
    out["cpu_oracle"] = {"sample_sequences": sample, "s_per_pattern_sample": cpu, "s_per_pattern_scaled": cpu * nseq / sample, "cores": 1}
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
