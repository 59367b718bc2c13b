"""TEST INFRASTRUCTURE ONLY: compiles the product's kernel SOURCES for the SIMT emulator (simt_emu.h).

The only rewrite of the sources is textual and mechanical: `extern __shared__` -> `extern` (the harness defines the arrays) and
`__shared__` -> `static` (one block runs at a time).  Everything else -- intrinsics, collectives, launch geometry -- is the
kernel's own code, compiled by g++ against the emulator's stand-in for <cuda_runtime.h>."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "xxmotif_b200", "csrc")
BUILD = os.path.join(HERE, "build")
KERNEL_HEADERS = ["xxm_device.cuh", "xxm_kernels.cuh", "xxm_scan4.cuh", "xxm_scan5.cuh", "xxm_scan6.cuh", "xxm_scan7.cuh", "xxm_select2.cuh", "xxm_select3.cuh",
                  "xxm_select4.cuh", "xxm_kmerhist.cuh", "xxm_gappedhist.cuh", "xxm_tables_device.cuh", "xxm_load.cuh", "xxm_region.cuh", "xxm_sweep.cuh", "xxm_stage.cuh", "xxm_xchg.cuh"]


def build(target="scan_emu", extra_sources=("host_pack.cpp", "pair_table.cpp"), sanitize=False, kernel_dir=None):
    """sanitize=True / "address": AddressSanitizer + UBSan over the emulated kernels -- every load and store of a kernel into the
    harness' "device" buffers (sized like the library's) and its `__shared__` arrays is bounds-checked, a memcheck without a GPU.
    sanitize="thread": ThreadSanitizer with one fiber per CUDA thread (simt_emu.h) -- accesses of one block to the same address
    that no barrier or warp fence orders are reported, a racecheck without a GPU; "thread-strict": only __syncthreads and
    __syncwarp order memory, the other *_sync collectives do not (what the CUDA programming guide promises).
    kernel_dir: take the rewritten kernel headers from there instead of BUILD (mutation controls)."""
    os.makedirs(BUILD, exist_ok=True)
    newest = 0.0
    for h in KERNEL_HEADERS:
        src = os.path.join(CSRC, h)
        text = re.sub(r"extern __shared__ (__align__\(\d+\) )?", "extern ", open(src).read()).replace("__shared__", "static")
        dst = os.path.join(BUILD, h)
        if not os.path.exists(dst) or open(dst).read() != text:
            with open(dst, "w") as fh:
                fh.write(text)
        newest = max(newest, os.path.getmtime(dst))
    mode = "address" if sanitize is True else (sanitize or "")
    exe = os.path.join(BUILD, target + {"": "", "address": "_asan", "thread": "_tsan", "thread-strict": "_tsan_strict"}[mode] + ("_mut" if kernel_dir else ""))
    deps = [os.path.join(HERE, target + ".cpp"), os.path.join(HERE, "simt_emu.h")] + [os.path.join(CSRC, s) for s in extra_sources]
    newest = max([newest] + [os.path.getmtime(d) for d in deps])
    if os.path.exists(exe) and os.path.getmtime(exe) >= newest and not kernel_dir:
        return exe
    flags = {"": [], "address": ["-fsanitize=address,undefined", "-fno-omit-frame-pointer"], "thread": ["-fsanitize=thread", "-fno-omit-frame-pointer"],
             "thread-strict": ["-fsanitize=thread", "-fno-omit-frame-pointer", "-DXXM_EMU_STRICT_SHFL"]}[mode]
    cmd = ["g++", "-std=c++17", "-O1", "-g1"] + flags + ["-ffp-contract=off", "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-unused-variable", "-Wno-maybe-uninitialized",
           "-I", kernel_dir or BUILD, "-I", os.path.join(HERE, "shim"), "-I", HERE, "-I", CSRC, os.path.join(HERE, target + ".cpp")]
    cmd += [os.path.join(CSRC, s) for s in extra_sources] + ["-o", exe, "-lpthread"]
    subprocess.check_call(cmd)
    return exe


def build_many(jobs, workers=None):
    """jobs: (target, extra_sources, sanitize) triples, compiled side by side (a fresh tree needs two dozen builds of 5-15 s each);
    returns {(target, sanitize): exe or the CalledProcessError}"""
    from concurrent.futures import ThreadPoolExecutor
    build("race_control", extra_sources=())  # refreshes the rewritten headers once, before the threads start
    out = {}

    def one(j):
        try:
            return j, build(j[0], extra_sources=j[1], sanitize=j[2])
        except subprocess.CalledProcessError as e:
            return j, e
    with ThreadPoolExecutor(max_workers=workers or max(1, min(8, (os.cpu_count() or 2) - 1))) as ex:
        for j, r in ex.map(one, jobs):
            out[(j[0], j[2])] = r
    return out


if __name__ == "__main__":
    import sys
    target = "scan_emu"
    args = sys.argv[1:]
    if args and args[0] in ("scan_emu", "step_emu"):
        target, args = args[0], args[1:]
    sys.exit(subprocess.call([build(target)] + args))
