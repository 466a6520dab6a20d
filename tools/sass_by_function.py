#!/usr/bin/env python
"""Static SASS size of a kernel by source function (innermost inlined frame) and by source line.
usage: sass_by_function.py <libprl.so> <kernel mangled-name substring> [top]"""
import collections, os, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(__file__))
from ncu_lines import line_map
from ncu_regions import func_table

def main():
    lib, kernel = os.path.abspath(sys.argv[1]), sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
    lm = line_map(os.path.join(tmp, "prl_api.sm_100a.cubin"), kernel)
    src = os.path.join(os.path.dirname(__file__), "..", "physics_reinforcement_learning_environment_b200", "csrc")
    tabs = {f: func_table(os.path.join(src, f)) for f in ("prl_device.cuh", "prl_kernels.cuh", "prl_libm.h")}
    byfn, byline = collections.Counter(), collections.Counter()
    for addr, key in lm.items():
        if key is None:
            byfn["?"] += 1; continue
        f, ln = key
        t = tabs.get(f)
        fn = t[ln - 1] if t and 0 < ln <= len(t) else f
        byfn[f"{f}:{fn}"] += 1
        byline[f"{f}:{ln}"] += 1
    total = sum(byfn.values())
    print(f"{kernel}: {total} SASS instructions = {total * 16 / 1024:.1f} KB")
    for k, v in byfn.most_common(top):
        print(f"  {v:5d}  {k}")
    print("  -- by line")
    for k, v in byline.most_common(15):
        print(f"  {v:5d}  {k}")

main()
