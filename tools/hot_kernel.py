#!/usr/bin/env python
"""Writes profiles/hot_kernel.json from an `ncu --page raw --csv` dump of a config-B run: DRAM bytes and
duration of the dominant kernel, stamped with the SHA-1 of the source file that defines it, so that
bench.py can tell when `roofline.traffic` no longer describes the kernel in the tree.
usage: hot_kernel.py <raw.csv> [kernel substring = wr_raster_solid_premult]"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def source_sha():
    return hashlib.sha1(open(os.path.join(ROOT, "webrender_b200", "csrc", "raster.cuh"), "rb").read()).hexdigest()


def main():
    raw = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else "wr_raster_solid_premult"
    rows = list(csv.reader(open(raw)))
    hdr = rows[0]
    col = {n: i for i, n in enumerate(hdr)}
    best = None
    for r in rows[2:]:
        if want in r[col["Kernel Name"]]:
            t = float(r[col["gpu__time_duration.sum"]])
            if best is None or t > best[0]:
                best = (t, r)
    if best is None:
        raise SystemExit(f"no launch of {want} in {raw}")
    t, r = best
    unit_t = rows[1][col["gpu__time_duration.sum"]]
    scale_t = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}.get(unit_t, 1e-3)

    def nbytes(name):
        v, u = float(r[col[name]]), rows[1][col[name]]
        return int(v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1))
    out = {"kernel": want, "dram_bytes_read": nbytes("dram__bytes_read.sum"), "dram_bytes_write": nbytes("dram__bytes_write.sum"),
           "gpu_time_ms": t * scale_t, "raster_cuh_sha1": source_sha(),
           "source": f"ncu --set full, profiles/{os.path.basename(raw)} (one launch, config B)"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "hot_kernel.json"), "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()
