#!/usr/bin/env python
"""Joins an ncu SASS-page CSV (per-address 'Instructions Executed') with nvdisasm -gi line
info to give instruction counts per innermost source line.
usage: ncu_lines.py <all.sass from nvdisasm -gi> <mangled kernel name> <ncu --page source --csv file> [top]"""
import collections
import csv
import re
import sys


def main():
    sass, kernel, src = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    lines = open(sass).read().split('\n')
    start = next(i for i, l in enumerate(lines) if l.startswith('.text.' + kernel))
    pat_file = re.compile(r'//## File "([^"]+)", line (\d+)')
    pat_ins = re.compile(r'/\*([0-9a-f]{4,})\*/\s+(\S.*?);')
    off2line, pending = {}, None
    for l in lines[start + 1:]:
        if l.strip().startswith('.section') and kernel not in l:
            break
        m = pat_file.search(l)
        if m and pending is None:
            pending = (m.group(1).split('/')[-1], int(m.group(2)))
        m = pat_ins.search(l)
        if m:
            if pending is not None:
                last = pending
            off2line[int(m.group(1), 16)] = last if 'last' in dir() else ('?', 0)
            pending = None
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    ia, ie = hdr.index('Address'), hdr.index('Instructions Executed')
    base, agg, tot = None, collections.Counter(), 0
    for r in rows[2:]:
        if len(r) <= ie:
            continue
        a = int(r[ia], 16)
        base = a if base is None else base
        n = int(r[ie] or 0)
        tot += n
        agg[off2line.get(a - base, ('?', 0))] += n
    print('total warp instructions', tot)
    byfile = collections.Counter()
    for (f, ln), n in agg.items():
        byfile[f] += n
    for f, n in byfile.most_common():
        print(f"  {f:28s} {n / 1e6:9.2f}M {100 * n / tot:5.1f}%")
    for (f, ln), n in agg.most_common(top):
        print(f"{f:28s} {ln:5d} {n / 1e6:9.2f}M {100 * n / tot:5.1f}%")


main()
