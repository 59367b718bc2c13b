#!/usr/bin/env python
"""SASS summary of the shipped library (cuobjdump -sass, no GPU needed): mnemonic counts of the kernels of a refinement step, the
tensor-core / TMA mnemonics of the whole library, and the listings of the find and expand kernels.

    python tools/sass_summary.py > profiles/r02e_sass_scan_kernels.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "xxmotif_b200", "libxxmotif_gpu.so")
STEP = ["k_scan5ILb1", "k_tile_scan", "k_expand7ILb1ELb1", "k_expand7ILb1ELb0", "k_select_cls_fwd", "k_select_cls_rc", "k_count_pwm3"]
LIST = ["k_scan5ILb1", "k_expand7ILb1ELb1"]
SPECIAL = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UTCMMA", "UTCHMMA", "HMMA", "IMMA", "LDS", "LDG", "STG", "PRMT", "LOP3", "SHF", "IMAD", "REDUX", "ATOMS", "ATOMG", "RED"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    funcs = {}
    for part in re.split(r"\n\s*Function : ", txt)[1:]:
        name = part.split("\n", 1)[0].strip()
        ins = []
        for line in part.split("\n"):
            m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s*(.*?)\s*;", line)
            if m:
                ins.append((m.group(1), m.group(2)))
        funcs[name] = ins
    print("# SASS of the shipped libxxmotif_gpu.so (cuobjdump -sass, sm_100a): kernels of a refinement step with the generation-7 scan.")
    print("# Mnemonic counts per kernel; then the instruction listings of k_scan5<true> (window-order masks) and k_expand7<true, true>.\n")
    for key in STEP:
        for name, ins in funcs.items():
            if key in name:
                c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i).split(".")[0].split(" ")[0] for _, i in ins)
                print("## %s: %d instructions" % (name, len(ins)))
                print(" ".join("%s %d" % kv for kv in c.most_common(26)) + "\n")
    total = collections.Counter()
    for ins in funcs.values():
        for _, i in ins:
            total[re.sub(r"^@!?U?P\d+\s+", "", i).split(".")[0].split(" ")[0]] += 1
    print("## whole library (%d kernels): tensor-core / TMA mnemonics" % len(funcs))
    print(" ".join("%s %d" % (k, total.get(k, 0)) for k in SPECIAL))
    print("(no TMA and no tensor-core instruction: the tables a kernel stages are 4 KB .. 128 KB read once per CTA with plain loads, and the path is a\n gather / compaction, not a contraction)\n")
    for key in LIST:
        for name, ins in funcs.items():
            if key in name:
                print("\n## listing: %s\n" % name)
                for a, i in ins:
                    print("        /*%s*/                   %s ;" % (a, i))


if __name__ == "__main__":
    sys.exit(main())
