#!/usr/bin/env python
"""Is the device code of the library in this tree the device code of an earlier commit?  Builds that commit in a scratch worktree
(nvcc cross-compiles sm_100a without a GPU) and compares `cuobjdump -sass` of the two libraries kernel by kernel: every kernel of
that commit must be found here with the same machine code (exit code 1 otherwise); kernels only this tree has are listed.  Used after moving kernels between files when no GPU is at hand (profiles/r02h_sass_identity.txt).

    python tools/sass_identity.py <commit> [<commit> ...]"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass(lib):
    """{mangled kernel name: SASS lines}"""
    import re
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    fn, cur = {}, None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            fn[cur] = []
        elif cur is not None and not ln.startswith("identifier = "):
            fn[cur].append(ln)
    return fn


def compare(here, there):
    """kernels of `there` whose machine code is found unchanged in `here` (under the same name, or under another: a template that grew
    a defaulted parameter is mangled differently), changed ones, and those only one side has"""
    bodies = {}
    for k, v in here.items():
        bodies.setdefault("\n".join(v), []).append(k)
    same, renamed, changed, gone = [], [], [], []
    for k, v in there.items():
        if k in here and here[k] == v:
            same.append(k)
        elif "\n".join(v) in bodies:
            renamed.append((k, bodies["\n".join(v)][0]))
        elif k in here:
            changed.append(k)
        else:
            gone.append(k)
    taken = set(same) | {b for _, b in renamed} | set(changed)
    new = [k for k in here if k not in taken]
    return same, renamed, changed, gone, new


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
            same, renamed, changed, gone, new = compare(here, there)
            print("%s: %d kernels there; %d identical here, %d identical under another name, %d CHANGED, %d gone; %d only here"
                  % (commit, len(there), len(same), len(renamed), len(changed), len(gone), len(new)))
            for a, b in renamed:
                print("  same code, other name: %s -> %s" % (a, b))
            for k in changed + gone:
                print("  changed or gone: %s" % k)
            for k in new:
                print("  only here: %s" % k)
            rc |= 1 if (changed or gone) else 0
        finally:
            subprocess.call(["git", "-C", ROOT, "worktree", "remove", "--force", wt])
    return rc


if __name__ == "__main__":
    sys.exit(main())
