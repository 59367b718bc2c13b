"""Host-side pieces of the product (table construction, optimal-subset sweep) against the reference's golden vectors.
These run without a GPU through the device-free entry points of the C ABI."""
import numpy as np
import pytest

import xxmotif_b200.capi as capi
from backends import tables_dump_dense
from conftest import golden_names
from replay import TraceInputs, _cmp_tables, _eq


def _table_prefixes(t):
    out = [k[:-len(".tab.L")] for k in t if k.endswith(".tab.L")]
    return out


@pytest.mark.parametrize("name", golden_names())
def test_host_tables_vs_reference(golden, name):
    t = golden(name)
    if "nm.rprobs" not in t:
        pytest.skip("order-8 NullModel tables are not in the fixture")
    ti = TraceInputs(t)
    n = 0
    for p in _table_prefixes(t):
        if p.startswith("opt"):
            pwm, cols, pss = t[p + ".in.pwm"], t[p + ".in.cols"], int(t[p + ".in.posSetSize"][0])
        elif p.startswith("el"):
            pwm, cols, pss = t[p + ".pwm_in"], t[p + ".cols"], int(t["opt1.in.posSetSize"][0])
        else:
            pwm, cols, pss = t["converged.pwm"], t["converged.cols"], int(t["converged.posSetSize"][0])
        d = capi.host_tables(pwm, cols, pss, int(ti.total[pss]), ti.bgLog, ti.order, t["nm.rprobs"], t["nm.rconds"])
        _cmp_tables(p + ".tab", tables_dump_dense(d["L"], d["maxk"], d["nparts"], d["pvalues"], d["probs"], d["classes"]), t, p + ".tab")
        n += 1
    assert n >= 3


@pytest.mark.parametrize("name", golden_names())
def test_host_best_subset_vs_reference(golden, name):
    t = golden(name)
    ti = TraceInputs(t)
    nent = len(ti.n_post)
    checked = 0
    for p in ["opt0", "opt1"] + ["el%d.%d" % (v, it) for v in range(int(t["el.count"][0])) for it in range(3)]:
        if p + ".sorted.score" not in t:
            continue
        sc, sq = t[p + ".sorted.score"], t[p + ".sorted.seq"]
        nb_key = p + (".bestMotifNb" if p.startswith("opt") else ".motifNb")
        bp_key = p + (".bestPval" if p.startswith("opt") else ".overrepPval")
        if len(sc) == 0 or sc[0] == 1:
            continue
        pss = int(t["opt1.in.posSetSize"][0])
        nb, bp = capi.host_best_subset(sc, np.where(sc < 1, sq, 1), pss, int(t[p + ".maxSites"][0]), nent, oneOccurrence=bool(ti.oops))
        _eq(p + ".bestMotifNb", nb, int(t[nb_key][0]))
        _eq(p + ".bestPval", bp, float(t[bp_key][0]))
        checked += 1
    assert checked >= 2


def _pair_flags(tab, codes):
    """numpy model of k_scan5: hit flag of every window of a base sequence (codes 0..3) from the 5-base pair table."""
    n = len(codes)
    c = np.concatenate([codes, np.zeros(16, dtype=np.int64)])
    four = c[:n + 8] | (c[1:n + 9] << 2) | (c[2:n + 10] << 4) | (c[3:n + 11] << 6)  # 4-mer starting at every base
    # entries are fetched per even base p through the 5-base index; a/b of base p and p+1 sit in bytes 3,2 / 1,0
    q = c[:n + 8] | (c[1:n + 9] << 2) | (c[2:n + 10] << 4) | (c[3:n + 11] << 6) | (c[4:n + 12] << 8)
    E = tab[q[0::2]].astype(np.int64)
    a = np.empty(len(q), dtype=np.int64)
    b = np.empty(len(q), dtype=np.int64)
    a[0::2], b[0::2] = (E >> 24) & 255, (E >> 16) & 255
    a[1::2], b[1::2] = ((E >> 8) & 255)[:len(a[1::2])], (E & 255)[:len(b[1::2])]
    del four
    return (a[:n] + b[4:n + 4]) >= 128


@pytest.mark.parametrize("name", golden_names())
def test_pair_table_exact(golden, name):
    """The pair table of the scan (csrc/pair_table.h) is an EXACT restatement of 'left k-mer is in the p-value table' for
    every table of the golden traces whose left match columns span at most 8 bases; longer spans are declined."""
    t = golden(name)
    if "nm.rprobs" not in t:
        pytest.skip("order-8 NullModel tables are not in the fixture")
    ti = TraceInputs(t)
    rng = np.random.default_rng(7)
    seq = rng.integers(0, 4, size=20000).astype(np.int64)
    checked = 0
    for p in _table_prefixes(t):
        if p.startswith("opt"):
            pwm, cols, pss = t[p + ".in.pwm"], t[p + ".in.cols"], int(t[p + ".in.posSetSize"][0])
        elif p.startswith("el"):
            pwm, cols, pss = t[p + ".pwm_in"], t[p + ".cols"], int(t["opt1.in.posSetSize"][0])
        else:
            pwm, cols, pss = t["converged.pwm"], t["converged.cols"], int(t["converged.posSetSize"][0])
        d = capi.host_tables(pwm, cols, pss, int(ti.total[pss]), ti.bgLog, ti.order, t["nm.rprobs"], t["nm.rconds"])
        pv = d["pvalues"][0]
        len0 = int(round(np.log2(len(pv)) / 2))
        coloff = (np.asarray(cols[:len0]) - int(cols[0])).astype(np.int32)
        member = pv != 0
        ids = np.nonzero(member)[0].astype(np.uint32)
        ok, tab, levels = capi.host_pair_table(ids, coloff)
        if coloff[-1] > 7:
            assert not ok
            continue
        assert ok and levels <= 127, (p, levels)
        # brute force over a random sequence: flag == membership of the gathered k-mer id
        flags = _pair_flags(tab, seq)
        c = np.concatenate([seq, np.zeros(16, dtype=np.int64)])
        kid = np.zeros(len(seq), dtype=np.int64)
        for k, o in enumerate(coloff):
            kid |= c[o:o + len(seq)] << (2 * k)
        assert np.array_equal(flags, member[kid]), p
        # and over all 4^8 windows
        w = np.arange(65536, dtype=np.int64)
        kid = np.zeros(65536, dtype=np.int64)
        for k, o in enumerate(coloff):
            kid |= ((w >> (2 * o)) & 3) << (2 * k)
        x0, y0 = w & 255, w >> 8
        a = (tab[x0 | ((0) << 8)] >> 24) & 255  # entry q with bases 0..3 = x0 (base 4 irrelevant for a[x0])
        bb = (tab[y0] >> 16) & 255
        assert np.array_equal((a.astype(np.int64) + bb.astype(np.int64)) >= 128, member[kid]), p
        checked += 1
    assert checked >= 1 or name in ("c5_loc_gap",)


def test_pair_table_declines_non_threshold_sets():
    ids = np.array([0b00000000_00000001, 0b00000001_00000000], dtype=np.uint32)  # rows {x=1: y=0}, {x=0: y=1}: not nested
    ok, _, _ = capi.host_pair_table(ids, np.arange(8, dtype=np.int32))
    assert not ok
    ok, tab, lv = capi.host_pair_table(np.zeros(0, dtype=np.uint32), np.arange(8, dtype=np.int32))
    assert ok and lv == 0 and not tab.any()


def test_many_runs_sweep_matches_literal_sweep():
    """Lists whose scores are not table valued reach getBestSubset as tens of thousands of runs: the threaded evaluation over
    dense LOG_sum_i windows (host_stats.cpp: best_subset_from_runs) must pick what the literal site-by-site sweep picks."""
    from xxmotif_b200 import capi, sharded
    rng = np.random.default_rng(3)
    nruns = 30000
    scores = np.sort(np.unique((10.0 ** rng.uniform(-9, -0.5, size=nruns)).astype(np.float32)))
    counts = rng.integers(1, 4, size=len(scores)).astype(np.int64)
    nb = int(counts.sum())
    maxSites = 50 * nb
    K, bp, last = sharded.choose_subset(scores, counts, nb, nb * 10, maxSites)
    site_scores = np.repeat(scores, counts)
    seq = np.arange(1, nb + 1, dtype=np.int32)
    K2, bp2 = capi.host_best_subset(site_scores, seq, nb, maxSites, nb)
    assert (K, bp) == (K2, bp2) and K > 1
    assert last == site_scores[K - 1]
    # the same runs below the threading threshold (sequential path) agree with themselves on a prefix
    K3, bp3, _ = sharded.choose_subset(scores[:1500], counts[:1500], int(counts[:1500].sum()), nb * 10, maxSites)
    K4, bp4 = capi.host_best_subset(np.repeat(scores[:1500], counts[:1500]), seq[:int(counts[:1500].sum())], nb, maxSites, nb)
    assert (K3, bp3) == (K4, bp4)


def test_bench_side_measurements_cannot_take_the_line_down(tmp_path):
    """bench.tool_json: a tool that prints a JSON line, one that fails without printing, one that hangs, one that prints garbage"""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    ok = tmp_path / "ok.py"
    ok.write_text("print('noise')\nprint('{\"a\": 1}')\nprint('{\"ok\": true, \"v\": [1, 2]}')\n")
    assert bench.tool_json([str(ok)], 30) == {"ok": True, "v": [1, 2]}
    bad = tmp_path / "bad.py"
    bad.write_text("import sys\nsys.stderr.write('no CUDA device')\nsys.exit(4)\n")
    r = bench.tool_json([str(bad)], 30)
    assert "failed" in r and "rc 4" in r["failed"] and "no CUDA device" in r["failed"]
    hang = tmp_path / "hang.py"
    hang.write_text("import time\ntime.sleep(60)\n")
    r = bench.tool_json([str(hang)], 1)
    assert "failed" in r and "Timeout" in r["failed"]
    junk = tmp_path / "junk.py"
    junk.write_text("print('{not json')\n")
    assert "failed" in bench.tool_json([str(junk)], 30)
    half = tmp_path / "half.py"
    half.write_text("import sys\nprint('{\"ok\": false}')\nsys.exit(3)\n")
    assert bench.tool_json([str(half)], 30) == {"ok": False, "rc": 3}
    # the two tools themselves, in a container without a GPU: they fail loudly, bench keeps going
    r = bench.tool_json([os.path.join(root, "tools", "motif_batch.py"), "--shapes", "cfg1", "--contexts", "2", "--motifs", "2"], 120)
    import torch
    if not torch.cuda.is_available():
        assert "failed" in r and "no CUDA device" in r["failed"]
