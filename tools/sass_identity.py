#!/usr/bin/env python
"""Is the device code of the library in this tree the device code of an earlier commit?  Builds that commit in a scratch worktree
(nvcc cross-compiles sm_100a without a GPU) and compares `cuobjdump -sass` of the two libraries; the path of the translation unit is
the only line allowed to differ.  Used after moving kernels between files when no GPU is at hand (profiles/r02h_sass_identity.txt).

    python tools/sass_identity.py <commit> [<commit> ...]"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    return [ln for ln in out.splitlines() if not ln.startswith("identifier = ")]


def main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "xxmotif_b200", "csrc")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    here = sass(os.path.join(ROOT, "xxmotif_b200", "libxxmotif_gpu.so"))
    rc = 0
    for commit in sys.argv[1:]:
        wt = tempfile.mkdtemp(prefix="xxm_wt_")
        try:
            subprocess.check_call(["git", "-C", ROOT, "worktree", "add", "-q", "--force", wt, commit])
            subprocess.check_call(["make", "-s", "-C", os.path.join(wt, "xxmotif_b200", "csrc")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            there = sass(os.path.join(wt, "xxmotif_b200", "libxxmotif_gpu.so"))
            same = here == there
            print("%s: %d lines of SASS, %s" % (commit, len(there), "identical" if same else "DIFFERENT (%d vs %d lines)" % (len(there), len(here))))
            rc |= 0 if same else 1
        finally:
            subprocess.call(["git", "-C", ROOT, "worktree", "remove", "--force", wt])
    return rc


if __name__ == "__main__":
    sys.exit(main())
