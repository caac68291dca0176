"""stdin: `ncu --page source --print-source cuda --csv`; stdout: per file the source lines that carry samples or
instructions, most expensive first (line, samples, instructions executed, source)."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
fname, hdr, out = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        fname = r[1]
        hdr = None
        continue
    if r and r[0] in ("Line No", "#"):
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))

        def num(*keys):
            for k in keys:
                if k in d:
                    try:
                        return float(d[k].replace(",", ""))
                    except ValueError:
                        return 0.0
            return 0.0
        samp = num("Warp Stall Sampling (All Samples)", "# Samples")
        inst = num("Instructions Executed")
        if samp or inst:
            out.append((samp, inst, fname, d.get("Line No", d.get("#", "")), d.get("Source", "")[:150]))
tot = sum(o[0] for o in out) or 1.0
toti = sum(o[1] for o in out) or 1.0
print("total samples %d, instructions %d" % (tot, toti))
for samp, inst, fname, line, src in sorted(out, reverse=True)[:90]:
    print("%5.1f%% %5.1f%%i %s:%s  %s" % (100 * samp / tot, 100 * inst / toti, (fname or "").split("/")[-1], line, src.strip()))
