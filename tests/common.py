"""Helpers shared by the parity tests."""
import numpy as np

from webrender_b200 import abi, draw_frame


def render(device_cls, frame, targets=None, tile_lists=False):
    """Run `frame` through a device; return {name: uint8 array} of target contents."""
    d = device_cls()
    try:
        handles = draw_frame(d, frame, tile_lists=tile_lists)
        out = {}
        names = targets or sorted({t.texture for p in frame.passes for t in p})
        for n in names:
            desc = frame.textures[n]
            out[n] = d.read_pixels(handles[n], 0, 0, desc.width, desc.height, abi.FMT_BPP[desc.fmt])
        return out
    finally:
        d.close()


def assert_same(a, b, what=""):
    for k in a:
        diff = a[k] != b[k]
        if diff.any():
            ys, xs = np.nonzero(diff)
            raise AssertionError(f"{what} target {k}: {int(diff.sum())} bytes differ; first at "
                                 f"(byte x={xs[0]}, y={ys[0]}): {a[k][ys[0], xs[0]]} vs {b[k][ys[0], xs[0]]}")
