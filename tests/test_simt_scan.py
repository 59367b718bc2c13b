"""The kernels of a refinement step themselves -- k_dead_mask, k_scan5 (find, both mask orders), k_tile_scan, k_expand6 and k_expand7,
compiled from their CUDA sources (xxmotif_b200/csrc/*.cuh) for the SIMT emulator of tests/simt/ -- against a brute-force
enumeration of the reference's window loop (StartPosUpdater-inl.h:70-140, table membership branch_and_bound-inl.h:386-392):
bit-exact (sequence, position, class) records in scan order.  Runs without a GPU; the emulator is test infrastructure only
(nothing under xxmotif_b200/ includes it, and libxxmotif_gpu.so has no CPU path)."""
import os
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt"))
import build as simt_build  # noqa: E402

CASES = ["tile_scan", "uniform_1kb", "ragged_N", "revcomp_ragged", "dense_lowcomplexity", "gapped_6cols", "five_columns", "tiny", "every_warp_dense",
         "ten_tiles_uniform", "one_word_sequences", "small_capacity", "positive_set_prefix_ragged", "positive_set_prefix_uniform",
         "two_part_twelve", "two_part_seventeen_rc", "two_part_gapped_uniform"]


STEP_CASES = ["clustered_uniform", "ragged_short", "long_motif", "all_sites_dense", "few_sites", "tiny", "zoops_uniform", "zoops_ragged_shift", "zoops_long_runs",
              "zoops_sparse", "rc_uniform", "rc_ragged", "rc_long_motif", "rc_long_sequences", "rc_palindromic_dense", "sort_two", "sort_one_chunk", "sort_ties",
              "sort_many_chunks", "gen_table_clustered", "gen_table_ragged_long", "gen_localized", "gen_localized_shift", "gen_per_candidate", "gen_dense", "gen_tiny",
              "cmp_flags_sparse", "cmp_flags_dense", "cmp_run_heads", "cmp_one", "cmp_none_selected", "cmp_all_one_run"]


@pytest.fixture(scope="module", autouse=True)
def _all_binaries():
    """every harness in every mode this module uses, compiled side by side up front (the tests below then find them built)"""
    targets = [("scan_emu", ("host_pack.cpp", "pair_table.cpp")), ("step_emu", ("host_pack.cpp",)), ("scan4_emu", ("host_pack.cpp",)), ("hist_emu", ("host_pack.cpp",)),
               ("tables_emu", ("pvalue_tables.cpp", "background_model.cpp")), ("memo_emu", ()), ("load_emu", ()), ("sweep_emu", ("host_stats.cpp",)), ("elem_emu", ())]
    jobs = [(t, e, m) for t, e in targets for m in (False, "address", "thread-strict")]
    jobs += [("race_control", (), "thread"), ("race_control", (), "thread-strict")]
    simt_build.build_many(jobs)  # a failed sanitizer build is left to the test that needs it (it skips)


@pytest.fixture(scope="module")
def exe():
    return simt_build.build()


@pytest.fixture(scope="module")
def step_exe():
    return simt_build.build("step_emu", extra_sources=("host_pack.cpp",))


@pytest.mark.parametrize("case", STEP_CASES)
def test_select_and_count_kernels_under_emulator(step_exe, case):
    """k_select_cls_fwd vs the reference's updateMultOcc greedy (StartPosUtils.h:12-28), k_count_pwm3 and k_count_pwm2 (keep flags + float
    scores, index list, ZOOPS slots) vs a column-by-column count;
    zoops_*: the re-score kernels k_eval_list + k_select_zoops vs the ZOOPS branch of
    fill_elongSites_with_possible_startPositions (StartPosUpdater-inl.h:167-203); rc_*: k_select_cls_rc (forward rule + one cursor into
    the frozen forward list) vs the reference's updateMultOcc with --revcomp, literally (the look-back walk, StartPosUtils.h:29-58);
    sort_*: the stable radix sort of (score bits, index) pairs vs std::stable_sort; gen_*: k_select_mops_fwd (the forward rule on float-valued
    lists: finalPval from a table in any order, from the localized MOPS formula, per candidate) vs the same greedy; cmp_*: the ordered
    compaction k_cmp_count + k_tile_scan + k_cmp_scatter (by flag byte, by run heads of sorted keys) vs std::copy_if"""
    r = subprocess.run([step_exe, case], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "DIFFERENT" not in r.stdout and " 0 failed" in r.stdout, r.stdout


@pytest.mark.parametrize("case", CASES)
def test_scan_kernels_under_emulator(exe, case):
    r = subprocess.run([exe, case], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "DIFFERENT" not in r.stdout and " 0 failed" in r.stdout, r.stdout


ASAN_ENV = dict(os.environ, ASAN_OPTIONS="detect_stack_use_after_return=0:detect_leaks=0", UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")


@pytest.mark.parametrize("target,extra,cases", [
    ("scan_emu", ("host_pack.cpp", "pair_table.cpp"), ["ragged_N", "revcomp_ragged", "gapped_6cols", "one_word_sequences", "small_capacity", "positive_set_prefix_ragged",
                                                         "two_part_seventeen_rc"]),
    ("step_emu", ("host_pack.cpp",), ["ragged_short", "all_sites_dense", "zoops_ragged_shift", "rc_ragged", "rc_long_sequences", "sort_many_chunks", "gen_table_clustered",
                                   "gen_localized_shift", "gen_per_candidate", "cmp_flags_dense", "cmp_run_heads", "cmp_one"]),
])
def test_kernels_under_address_sanitizer(target, extra, cases):
    """The same kernels with every access bounds-checked (AddressSanitizer + UBSan over the emulated run; compute-sanitizer is closed on
    the GPU pool, profiles/r02_compute_sanitizer_closed.txt).  The harness' buffers have the library's sizes (stream + 2 guard words,
    record arrays of exactly the capacity passed to the kernel + 1).  The full list of 33 cases: profiles/r02h_emulator_asan.txt."""
    try:
        exe = simt_build.build(target, extra_sources=extra, sanitize=True)
    except subprocess.CalledProcessError:
        pytest.skip("no sanitizer runtime in this toolchain")
    for case in cases:
        r = subprocess.run([exe, case], capture_output=True, text=True, timeout=900, env=ASAN_ENV)
        if r.returncode and "Shadow memory range interleaves" in r.stderr:
            pytest.skip("AddressSanitizer cannot map its shadow memory here")
        assert r.returncode == 0, case + "\n" + r.stdout[-2000:] + r.stderr[-4000:]
        assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, case + "\n" + r.stderr[-4000:]
        assert "DIFFERENT" not in r.stdout and " 0 failed" in r.stdout, r.stdout


TSAN_ENV = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=66")


def _tsan_run(exe, case):
    r = subprocess.run([exe, case], capture_output=True, text=True, timeout=900, env=TSAN_ENV)
    if r.returncode and "FATAL: ThreadSanitizer" in r.stderr:
        pytest.skip("ThreadSanitizer cannot start here: " + r.stderr.strip().splitlines()[0])
    return r, r.stderr.count("WARNING: ThreadSanitizer: data race")


def test_race_detector_controls():
    """the emulator's racecheck (ThreadSanitizer, one fiber per CUDA thread, edges only where the programming model gives them) on
    kernels with a known hazard and their repaired twins (tests/simt/race_control.cpp)"""
    try:
        lenient = simt_build.build("race_control", extra_sources=(), sanitize="thread")
        strict = simt_build.build("race_control", extra_sources=(), sanitize="thread-strict")
    except subprocess.CalledProcessError:
        pytest.skip("no ThreadSanitizer runtime in this toolchain")
    expect = {"warp_racy": (True, True), "warp_synced": (False, False), "block_racy": (True, True), "block_synced": (False, False),
              "shfl": (False, True), "atomics": (False, False)}  # (reported by the default model, by the strict one)
    for case, (a, b) in expect.items():
        for exe, want in ((lenient, a), (strict, b)):
            r, n = _tsan_run(exe, case)
            assert (n > 0) == want, (case, exe, n, r.stderr[-3000:])
            assert " done" in r.stdout


@pytest.mark.parametrize("target,extra,cases", [
    ("scan_emu", ("host_pack.cpp", "pair_table.cpp"), ["uniform_1kb", "ragged_N", "revcomp_ragged", "every_warp_dense", "small_capacity", "positive_set_prefix_ragged",
                                                         "two_part_twelve"]),
    ("step_emu", ("host_pack.cpp",), ["clustered_uniform", "all_sites_dense", "zoops_ragged_shift", "rc_ragged", "rc_long_sequences", "sort_many_chunks", "gen_table_clustered",
                                   "gen_localized_shift", "gen_per_candidate", "cmp_flags_dense", "cmp_run_heads", "cmp_one"]),
])
def test_kernels_under_race_detector(target, extra, cases):
    """The kernels of a step under the STRICT model (only __syncthreads and __syncwarp order memory): no two accesses of a block to
    one shared or global address with a write among them and no fence between them.  All 33 cases, both models:
    profiles/r02h_emulator_tsan.txt."""
    try:
        exe = simt_build.build(target, extra_sources=extra, sanitize="thread-strict")
    except subprocess.CalledProcessError:
        pytest.skip("no ThreadSanitizer runtime in this toolchain")
    for case in cases:
        r, n = _tsan_run(exe, case)
        assert n == 0 and r.returncode == 0, case + "\n" + r.stderr[-6000:]
        assert "DIFFERENT" not in r.stdout and " 0 failed" in r.stdout, r.stdout


def test_race_detector_sees_a_removed_fence(tmp_path):
    """mutation control on a product kernel: k_expand7 without the __syncwarp between filling the hit queue and reading it"""
    import re
    import shutil
    simt_build.build("scan_emu")  # refreshes the rewritten headers
    mut = tmp_path / "mut"
    shutil.copytree(simt_build.BUILD, mut, ignore=shutil.ignore_patterns("*_emu*", "race_control*"))
    f = mut / "xxm_scan7.cuh"
    text = f.read_text()
    # the fence right after the lock-step extraction loop (the sparse path every case takes)
    m = re.search(r"pop_msb\(r3, q3, tag \+ 96\);\s*\}\s*__syncwarp\(\);", text)
    assert m, "k_expand7 changed: update the mutation"
    f.write_text(text[:m.start()] + m.group(0).replace("__syncwarp();", "/* fence removed */") + text[m.end():])
    try:
        exe = simt_build.build("scan_emu", sanitize="thread", kernel_dir=str(mut))
    except subprocess.CalledProcessError:
        pytest.skip("no ThreadSanitizer runtime in this toolchain")
    r, n = _tsan_run(exe, "uniform_1kb")
    assert n > 0 and "xxm_scan7.cuh" in r.stderr, r.stderr[-3000:]


@pytest.mark.parametrize("mode", [False, "address", "thread-strict"])
def test_histogram_kernels_under_emulator(mode):
    """k_kmer_hist (all three bin layouts: private bins without atomics, lane-replicated bins, plain bins in two passes),
    k_gapped_hist (contiguous and gapped patterns, a prefix of the set, the first-half rule of palindromic seeds under --revcomp) and
    k_logodds_hist (a14; contiguous and gapped columns, N penalty, a position limit below the longest sequence) against
    position-by-position counts over the raw sequences, and k_encode (a1: the device packing incl. createRevcomp) against the host
    packing word for word -- plain, bounds-checked, and under the strict race model"""
    try:
        exe = simt_build.build("hist_emu", extra_sources=("host_pack.cpp",), sanitize=mode)
    except subprocess.CalledProcessError:
        if mode:
            pytest.skip("no sanitizer runtime in this toolchain")
        raise
    env = dict(ASAN_ENV, TSAN_OPTIONS=TSAN_ENV["TSAN_OPTIONS"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900, env=env)
    if r.returncode and ("FATAL: ThreadSanitizer" in r.stderr or "Shadow memory range interleaves" in r.stderr):
        pytest.skip("the sanitizer cannot start here")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "DIFFERENT" not in r.stdout and " 0 failed" in r.stdout, r.stdout
    for mark in ("ERROR: AddressSanitizer", "runtime error", "WARNING: ThreadSanitizer"):
        assert mark not in r.stderr, r.stderr[-4000:]


@pytest.mark.parametrize("mode,cases", [(False, []), ("address", ["five_columns", "gapped_six", "ties_seven", "one_column"]),
                                        ("thread-strict", ["five_columns", "gapped_six", "ties_seven", "one_column"])])
def test_device_table_build_under_emulator(mode, cases):
    """k_tab_count + k_tile_scan + k_tab_leaves + k_tab_rank + k_tab_accumulate, launched as the library launches
    them, against the product's host build (ids, classes, integer scores, class sizes, cumulative FP64 class p-values: bit for bit) for
    5..8 columns, gap columns, null-model orders 0..4 and PWMs whose k-mers tie in score -- plain (all 10 cases), bounds-checked and
    under the strict race model (4 cases; all 10: profiles/r02h_emulator_tsan.txt)"""
    try:
        exe = simt_build.build("tables_emu", extra_sources=("pvalue_tables.cpp", "background_model.cpp"), sanitize=mode)
    except subprocess.CalledProcessError:
        if mode:
            pytest.skip("no sanitizer runtime in this toolchain")
        raise
    env = dict(ASAN_ENV, TSAN_OPTIONS=TSAN_ENV["TSAN_OPTIONS"])
    r = subprocess.run([exe] + cases, capture_output=True, text=True, timeout=900, env=env)
    if r.returncode and ("FATAL: ThreadSanitizer" in r.stderr or "Shadow memory range interleaves" in r.stderr):
        pytest.skip("the sanitizer cannot start here")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "DIFFERENT" not in r.stdout and " 0 failed" in r.stdout, r.stdout
    for mark in ("ERROR: AddressSanitizer", "runtime error", "WARNING: ThreadSanitizer"):
        assert mark not in r.stderr, r.stderr[-4000:]


@pytest.mark.parametrize("mode", [False, "address", "thread-strict"])
def test_memo_evaluation_under_emulator(mode):
    """k_memo_eval_list (the convolution over the left score classes for motifs with more than 8 match columns; the kernel bisects where
    the reference walks its class pointer down) against the reference's loop written out literally (branch_and_bound-inl.h:425-449,
    binarySearch :337-357), FP64 bit for bit, class arrays in shared and in global memory -- plain, bounds-checked, strict race model"""
    try:
        exe = simt_build.build("memo_emu", extra_sources=(), sanitize=mode)
    except subprocess.CalledProcessError:
        if mode:
            pytest.skip("no sanitizer runtime in this toolchain")
        raise
    env = dict(ASAN_ENV, TSAN_OPTIONS=TSAN_ENV["TSAN_OPTIONS"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900, env=env)
    if r.returncode and ("FATAL: ThreadSanitizer" in r.stderr or "Shadow memory range interleaves" in r.stderr):
        pytest.skip("the sanitizer cannot start here")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "DIFFERENT" not in r.stdout and " 0 failed" in r.stdout, r.stdout
    for mark in ("ERROR: AddressSanitizer", "runtime error", "WARNING: ThreadSanitizer"):
        assert mark not in r.stderr, r.stderr[-4000:]


@pytest.mark.parametrize("mode", [False, "address", "thread-strict"])
def test_generic_find_kernel_under_emulator(mode):
    """k_build_t9 + k_scan4<true/false> -- the find kernel of every model the pair table cannot represent (gapped left parts beyond 8
    bases whose scan must find sites only, tables that are not threshold-shaped) -- staged and launched as the library does, against the definition of a scan
    candidate taken window by window (no N in a match column, left k-mer in the table): exact and projected table entries, bitmap in
    shared and in global memory, --revcomp, one-word sequences, several blocks; then k_expand5<true/false> on those masks -- ordered
    (sequence, position, value index) records incl. the memo slots and rejections of two-part models, tiles with more hits than the
    queue holds, a look-back across 42 tiles, record arrays of exactly the right size and five records short -- plain, bounds-checked,
    strict race model"""
    try:
        exe = simt_build.build("scan4_emu", extra_sources=("host_pack.cpp",), sanitize=mode)
    except subprocess.CalledProcessError:
        if mode:
            pytest.skip("no sanitizer runtime in this toolchain")
        raise
    env = dict(ASAN_ENV, TSAN_OPTIONS=TSAN_ENV["TSAN_OPTIONS"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900, env=env)
    if r.returncode and ("FATAL: ThreadSanitizer" in r.stderr or "Shadow memory range interleaves" in r.stderr):
        pytest.skip("the sanitizer cannot start here")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "DIFFERENT" not in r.stdout and " 0 failed" in r.stdout, r.stdout
    for mark in ("ERROR: AddressSanitizer", "runtime error", "WARNING: ThreadSanitizer"):
        assert mark not in r.stderr, r.stderr[-4000:]


@pytest.mark.parametrize("mode", [False, "address", "thread-strict"])
def test_load_kernels_under_emulator(mode):
    """a1 on the device: k_seq_layout_sums + k_seq_layout (offsets and stored lengths of 1 .. 5000 sequences, blocks of 1024, --revcomp),
    k_blk2seq and k_seq_stats (residue counts, a code outside 0..4 flagged; raw array of exactly the residues) against the same
    quantities sequence by sequence -- plain, bounds-checked, strict race model"""
    try:
        exe = simt_build.build("load_emu", extra_sources=(), sanitize=mode)
    except subprocess.CalledProcessError:
        if mode:
            pytest.skip("no sanitizer runtime in this toolchain")
        raise
    env = dict(ASAN_ENV, TSAN_OPTIONS=TSAN_ENV["TSAN_OPTIONS"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900, env=env)
    if r.returncode and ("FATAL: ThreadSanitizer" in r.stderr or "Shadow memory range interleaves" in r.stderr):
        pytest.skip("the sanitizer cannot start here")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "DIFFERENT" not in r.stdout and " 0 failed" in r.stdout, r.stdout
    for mark in ("ERROR: AddressSanitizer", "runtime error", "WARNING: ThreadSanitizer"):
        assert mark not in r.stderr, r.stderr[-4000:]


@pytest.mark.parametrize("mode,cases", [(False, []), ("address", ["limit_binds", "few_runs", "tiny"]), ("thread-strict", ["strong_motif", "few_runs", "tiny"])])
def test_sweep_prefilter_under_emulator(mode, cases):
    """a8: the device prefilter of getBestSubset's sweep (k_pack_runs, k_sweep_eval, k_sweep_long, k_sweep_flag, k_sweep_gather with the
    ordered compaction in between, staged as sorted_score_runs stages them) must hand the exact host sweep (the product's
    best_subset_from_runs) runs that lead to the SAME subset -- size, log p-value bit for bit, last score -- as all runs do; decided
    approximations within 1e-9 of the sequential formula (the warp-scan composition of k_sweep_long), the minimum key right"""
    try:
        exe = simt_build.build("sweep_emu", extra_sources=("host_stats.cpp",), sanitize=mode)
    except subprocess.CalledProcessError:
        if mode:
            pytest.skip("no sanitizer runtime in this toolchain")
        raise
    env = dict(ASAN_ENV, TSAN_OPTIONS=TSAN_ENV["TSAN_OPTIONS"])
    r = subprocess.run([exe] + cases, capture_output=True, text=True, timeout=900, env=env)
    if r.returncode and ("FATAL: ThreadSanitizer" in r.stderr or "Shadow memory range interleaves" in r.stderr):
        pytest.skip("the sanitizer cannot start here")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "DIFFERENT" not in r.stdout and " 0 failed" in r.stdout, r.stdout
    for mark in ("ERROR: AddressSanitizer", "runtime error", "WARNING: ThreadSanitizer"):
        assert mark not in r.stderr, r.stderr[-4000:]


@pytest.mark.parametrize("mode", [False, "address", "thread-strict"])
def test_elementwise_kernels_under_emulator(mode):
    """the one-thread-per-element kernels (fills, gathers, flags, value conversions under the three kinds of finalPval, site bits, memo
    bookkeeping, site records, the dense rank table) against the obvious loops, arrays of exactly n elements"""
    try:
        exe = simt_build.build("elem_emu", extra_sources=(), sanitize=mode)
    except subprocess.CalledProcessError:
        if mode:
            pytest.skip("no sanitizer runtime in this toolchain")
        raise
    env = dict(ASAN_ENV, TSAN_OPTIONS=TSAN_ENV["TSAN_OPTIONS"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900, env=env)
    if r.returncode and ("FATAL: ThreadSanitizer" in r.stderr or "Shadow memory range interleaves" in r.stderr):
        pytest.skip("the sanitizer cannot start here")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "DIFFERENT" not in r.stdout and "0 failed" in r.stdout, r.stdout
    for mark in ("ERROR: AddressSanitizer", "runtime error", "WARNING: ThreadSanitizer"):
        assert mark not in r.stderr, r.stderr[-4000:]


def test_every_kernel_runs_under_the_emulator():
    """every __global__ function of the library is launched by one of the harnesses (a new kernel needs a case before it ships)"""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    kernels = set()
    for f in glob.glob(os.path.join(root, "xxmotif_b200", "csrc", "*")):
        if f.endswith((".cuh", ".cu", ".inc")):
            kernels |= set(re.findall(r"__global__\s+void\s+(?:__launch_bounds__\([^)]*\)\s*)?(\w+)", open(f).read()))
    harness = " ".join(open(f).read() for f in glob.glob(os.path.join(root, "tests", "simt", "*.cpp")))
    missing = sorted(k for k in kernels if not re.search(r"\b" + k + r"\b", harness))
    assert len(kernels) >= 60 and not missing, missing
    # ... and the kernels live in headers (what the emulator can compile), not next to host code
    for f in glob.glob(os.path.join(root, "xxmotif_b200", "csrc", "*")):
        if f.endswith((".cu", ".inc")):
            assert "__global__" not in open(f).read(), f


def test_product_does_not_reference_the_emulator():
    """the emulator hooks in the kernel sources are compile-time (#ifdef XXM_SIMT_EMU) and the library is built without them"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mk = open(os.path.join(root, "xxmotif_b200", "csrc", "Makefile")).read()
    assert "XXM_SIMT_EMU" not in mk and "simt" not in mk
    for dirpath, _, files in os.walk(os.path.join(root, "xxmotif_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".inc", ".h", ".cuh")):
                assert "simt_emu.h" not in open(os.path.join(dirpath, f), errors="ignore").read(), f
