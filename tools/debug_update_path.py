"""GPU debugging aid: the two-frame update-path sequence (tests/update_path.py) on the CUDA backend vs the oracle, with
a description of where the frames differ."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.backends import OracleDevice  # noqa: E402
from update_path import run_sequence  # noqa: E402
from webrender_b200.device import CudaDevice  # noqa: E402


def run(cls):
    d = cls()
    try:
        return run_sequence(d, "text", 1)
    finally:
        d.close()


want = run(OracleDevice)
got = run(CudaDevice)
for name, g, w in zip(("frame1", "frame2", "atlas"), got, want):
    g = np.asarray(g); w = np.asarray(w)
    diff = g != w
    if not diff.any():
        print(name, "equal")
        continue
    ys, xs = np.nonzero(diff)
    print(name, "bytes differing", int(diff.sum()), "rows", ys.min(), ys.max(), "byte cols", xs.min(), xs.max())
    px = sorted({(int(y), int(x) // 4) for y, x in zip(ys[:4000], xs[:4000])})[:12]
    for y, x in px:
        print("   px", (x, y), "got", g[y, 4 * x:4 * x + 4], "want", w[y, 4 * x:4 * x + 4], "frame1", np.asarray(want[0])[y, 4 * x:4 * x + 4])
    # how many differing pixels equal frame 1's value (stale commands) ?
    g4 = g.reshape(g.shape[0], -1, 4); w4 = w.reshape(w.shape[0], -1, 4); f4 = np.asarray(want[0]).reshape(g4.shape)
    dpx = (g4 != w4).any(axis=2)
    print("   differing pixels", int(dpx.sum()), "of which equal to frame 1's pixel", int(((g4 == f4).all(axis=2) & dpx).sum()))
