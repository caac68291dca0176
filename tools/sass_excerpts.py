#!/usr/bin/env python
"""Writes profiles/r02/sass_excerpts.txt: the copy-engine (UTMALDG / UTMASTG), mbarrier (SYNCS) and programmatic-dependent-
launch (ACQBULK = griddepcontrol.wait, PREEXIT = griddepcontrol.launch_dependents) instructions of the tuned kernels, from
`cuobjdump -sass webrender_b200/libwrcu.so`.  No GPU needed."""
import os
import re
import subprocess
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "webrender_b200", "libwrcu.so")], capture_output=True, text=True).stdout
lines = sass.split("\n")


def block(fname):
    start, end = None, len(lines)
    for i, l in enumerate(lines):
        if start is None and "Function : " + fname in l:
            start = i
        elif start is not None and "Function : " in l:
            end = i
            break
    return lines[start:end] if start is not None else []


out = ["# cuobjdump -sass webrender_b200/libwrcu.so (sm_100a): copy engine, mbarrier and programmatic-dependent-launch\n"
       "# instructions of the tuned kernels (tools/sass_excerpts.py)\n"]
PAT = r"UTMALDG|UTMASTG|SYNCS|UTMACMDFLUSH|UBLKCP|ACQBULK|PREEXIT|FENCE|UTMACCTL"
for fn in ("_Z17wr_composite_copyILb0EEv10RasterArgs", "_Z17wr_composite_copyILb1EEv10RasterArgs", "wr_raster_solid_premult",
           "wr_raster_solid_flat", "_Z9wr_rasterI10QuadShaderLi1ELb0EEv10RasterArgs", "wr_setup_multi"):
    b = block(fn)
    n = sum(1 for l in b if re.search(r"/\*[0-9a-f]{4,6}\*/\s+\w", l))
    hits = [l.strip() for l in b if re.search(PAT, l)]
    ops = Counter(re.sub(r"/\*[0-9a-f]+\*/", "", h).split()[0].rstrip(";") for h in hits)
    out.append("== %s: %d SASS instructions (%.1f KB)\n" % (fn, n, n * 16 / 1024))
    out.append("   " + ", ".join("%s x%d" % kv for kv in ops.most_common(14)) + "\n")
    for h in hits[:12]:
        out.append("   " + h[:140] + "\n")
os.makedirs(os.path.join(ROOT, "profiles", "r02"), exist_ok=True)
open(os.path.join(ROOT, "profiles", "r02", "sass_excerpts.txt"), "w").write("".join(out))
print("".join(out)[:600])
