"""xxm_iterate_motifs (csrc/refine.cpp), the part that runs without a GPU: argument checks of the motif-batch entry point
(SURVEY.md 8(f) row 2).  The refinement itself is GPU-only: tests/test_gpu_motif_batch.py."""
import ctypes as C



def _call(L, ctx_ptrs, motifs, nmotifs, nparams):
    from xxmotif_b200 import refine
    hs = (C.c_void_p * max(1, len(ctx_ptrs)))(*ctx_ptrs) if ctx_ptrs else (C.c_void_p * 1)()
    arr = (refine.XxmMotif * max(1, nmotifs))()
    p = (refine.XxmRefineParams * max(1, nparams))()
    rounds = (C.c_int32 * max(1, nmotifs))()
    bad, badCtx = C.c_int32(7), C.c_int32(7)
    rc = L.xxm_iterate_motifs(hs, len(ctx_ptrs), arr if motifs else None, nmotifs, p, nparams, rounds, C.byref(bad), C.byref(badCtx))
    return rc, bad.value, badCtx.value


def test_iterate_motifs_argument_checks():
    from xxmotif_b200 import refine
    L = refine._lib()
    fake = [0x1000, 0x2000]  # never dereferenced: no motif is started
    assert _call(L, [], True, 0, 1) == (1, -1, -1)                 # no context
    assert _call(L, [fake[0], fake[0]], True, 0, 1) == (1, -1, -1)  # two threads on one context
    assert _call(L, [fake[0], None], True, 0, 1)[0] == 1            # a null context
    assert _call(L, fake, True, 3, 2)[0] == 1                       # params: one entry or one per motif
    assert _call(L, fake, False, 2, 1)[0] == 1                      # motifs without an array
    assert _call(L, fake, True, 0, 1) == (0, -1, -1)                # nothing to do


def test_iterate_motifs_threading_against_a_stand_in(tmp_path):
    """csrc/refine.cpp linked against a deterministic stand-in for the data-sized calls (tests/host/refine_batch_test.cpp): 23 motifs
    on 1..6 contexts end where the serial loop ends, no context is entered by two threads, a failing motif is reported."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "refine_batch_test")
    srcs = [os.path.join(root, "tests", "host", "refine_batch_test.cpp"), os.path.join(root, "xxmotif_b200", "csrc", "refine.cpp")]
    r = None
    for flags in (["-fsanitize=thread"], []):  # under ThreadSanitizer where its runtime starts (it refuses some address-space layouts), else plain
        if subprocess.call(["g++", "-O1", "-g", "-std=c++17", "-pthread"] + flags + srcs + ["-o", exe], stderr=subprocess.DEVNULL):
            continue
        r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
        if flags and r.returncode and "FATAL: ThreadSanitizer" in r.stderr:
            continue
        break
    assert r is not None and r.returncode == 0, (r.stdout + r.stderr) if r else "does not compile"
    assert " 0 problems" in r.stdout
    assert "WARNING: ThreadSanitizer" not in r.stderr
