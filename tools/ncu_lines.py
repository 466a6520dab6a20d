#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` SASS dump by CUDA source line.

usage: ncu_lines.py <src.csv> <kernel mangled-name substring> <cubin> [top]
Maps each SASS address to the `//## File ..., line N` marker nvdisasm -g prints (needs -lineinfo),
then sums samples / executed instructions / stall reasons per source line and per function-ish
region. Used to write the profiles/*.md summaries.
"""
import collections
import csv
import re
import subprocess
import sys


def line_map(cubin, kernel):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    m, cur, inside = {}, None, False
    for ln in txt:
        if ln.startswith(".text."):
            inside = kernel in ln
            continue
        if not inside:
            continue
        f = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if f:
            cur = (f.group(1).split("/")[-1], int(f.group(2)))
            continue
        a = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
        if a:
            m[int(a.group(1), 16)] = cur
    return m


def main():
    src, kernel, cubin = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    lm = line_map(cubin, kernel)
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
    h = rows[hi]
    ci = {n: i for i, n in enumerate(h)}
    base = None
    agg = collections.defaultdict(lambda: collections.Counter())
    stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    for r in rows[hi + 1:]:
        if len(r) < len(h):
            continue
        addr = int(r[0], 16)
        base = addr if base is None else base
        key = lm.get(addr - base, ("?", 0))
        a = agg[key]
        a["samples"] += int(r[ci["# Samples"]])
        a["inst"] += int(r[ci["Instructions Executed"]])
        a["thread_inst"] += int(r[ci["Thread Instructions Executed"]])
        a["sass"] += 1
        for c in stall_cols:
            a[c] += int(r[ci[c]] or 0)
    tot = collections.Counter()
    for a in agg.values():
        tot.update(a)
    print(f"total samples {tot['samples']}, warp inst {tot['inst']}, thread inst {tot['thread_inst']}, "
          f"avg active lanes {tot['thread_inst'] / max(1, tot['inst']):.2f}, SASS instructions {tot['sass']}")
    print("stalls:", ", ".join(f"{c[6:]} {100 * tot[c] / tot['samples']:.1f}%" for c in sorted(stall_cols, key=lambda c: -tot[c])[:8]))
    src_lines = {}
    print(f"{'file:line':28s} {'samp%':>6s} {'inst%':>6s} {'lanes':>5s} {'sass':>5s}  top stalls")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(stall_cols, key=lambda c: -a[c])[:3]
        print(f"{key[0] + ':' + str(key[1]):28s} {100 * a['samples'] / tot['samples']:6.2f} {100 * a['inst'] / max(1, tot['inst']):6.2f} "
              f"{a['thread_inst'] / max(1, a['inst']):5.1f} {a['sass']:5d}  " + ", ".join(f"{c[6:]} {a[c]}" for c in st))


if __name__ == "__main__":
    main()
