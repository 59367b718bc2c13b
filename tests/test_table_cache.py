"""CPU: build_pvalue_tables with the background-probability cache of a context (BgProbCache, csrc/pvalue_tables.h) gives the tables
of a fresh build bit for bit, over 120 builds with drifting PWMs, 5-8 and 12 columns, gap columns and a change of the null model
(tests/host/table_cache_test.cpp, compiled against the product's pvalue_tables.cpp)."""
import os
import subprocess


def test_cached_table_builds_are_bit_identical(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "xxmotif_b200", "csrc")
    exe = str(tmp_path / "table_cache_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-I", csrc, os.path.join(root, "tests", "host", "table_cache_test.cpp"),
                           os.path.join(csrc, "pvalue_tables.cpp"), os.path.join(csrc, "background_model.cpp"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 differ" in r.stdout
