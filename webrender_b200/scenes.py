"""Synthetic frames for the configurations BASELINE.json names.

Each builder returns a `Frame` holding exactly the binary tables and batches
the reference's frame builder would hand to `Renderer::draw_frame` for that
scene; deterministic (seeded numpy RNG), no file or network input.
"""
import numpy as np

from . import abi
from .frame import Batch, Clear, Frame, Target, TextureDesc
from .gpu_types import (FrameTables, INVALID_SEGMENT_INDEX, PART_ALL, QF_APPLY_DEVICE_CLIP, quad_instance)


def alpha_rects_frame(width=3840, height=2160, n_rects=1000, random_rects=False, seed=1,
                      color=(0.05, 0.05, 0.05, 0.05), clear_color=(0.3, 0.0, 0.0, 1.0),
                      blend=abi.BLEND_PREMULTIPLIED_ALPHA):
    """Config B — examples/alpha_perf.rs:35-49: N overlapping alpha rects in ONE
    batch of `Quad(ColorOrTexture)` instances (plain rects take the quad path,
    prepare.rs:216-259) over a cleared colour target.  `random_rects` gives the
    B' variant: seeded uniform origins, sizes in [64, 1024] px."""
    t = FrameTables()
    task = t.add_render_task((0.0, 0.0, float(width), float(height)), 1.0, (0.0, 0.0))
    rng = np.random.RandomState(seed)
    inst = []
    for i in range(n_rects):
        if random_rects:
            w, h = rng.randint(64, 1025, size=2)
            x0 = int(rng.randint(0, max(1, width - 32)))
            y0 = int(rng.randint(0, max(1, height - 32)))
            rect = (float(x0), float(y0), float(min(width, x0 + w)), float(min(height, y0 + h)))
        else:
            rect = (0.0, 0.0, float(width), float(height))
        if color is None:
            a = rng.uniform(0.05, 1.0)
            c = tuple(float(v) for v in (rng.uniform(0, a), rng.uniform(0, a), rng.uniform(0, a), a))
        else:
            c = color
        prim_f = t.add_quad_prim(rect, rect, c)
        prim_i = t.add_quad_header(0, i + 1)
        inst.append(quad_instance(prim_i, prim_f, QF_APPLY_DEVICE_CLIP, 0, PART_ALL, INVALID_SEGMENT_INDEX, task))
    inst = np.stack(inst)
    textures = {"target": TextureDesc(abi.FMT_RGBA8, width, height)}
    ops = [Clear(color=clear_color),
           Batch(abi.KIND_QUAD_TEXTURED, inst, blend=blend)]
    return Frame(t.arrays(), textures, [[Target("target", ops=ops)]])


def pixel_layers_of_quad_batch(frame: Frame):
    """Σ covered pixels over the quad instances of the first batch (axis-aligned,
    identity transform): the unit of work of the Mpix/s metric."""
    tgt = frame.passes[0][0]
    desc = frame.textures[tgt.texture]
    batch = [op for op in tgt.ops if isinstance(op, Batch)][0]
    gf = frame.tables["gpu_buffer_f"]
    total = 0
    for row in batch.instances:
        b = gf[row[1]]
        x0, y0 = max(0.0, b[0]), max(0.0, b[1])
        x1, y1 = min(float(desc.width), b[2]), min(float(desc.height), b[3])
        total += max(0, int(np.floor(x1 + 0.5)) - int(np.floor(x0 + 0.5))) * \
            max(0, int(np.floor(y1 + 0.5)) - int(np.floor(y0 + 0.5)))
    return total
