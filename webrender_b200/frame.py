"""Frame description consumed by the draw path: the subset of the reference's
`Frame` (frame_builder.rs:1129-1180) that `Renderer::draw_frame` reads — the
data tables, textures, and per pass the targets with their clears and batches.

`draw_frame(device, frame)` replays the reference's call sequence
(renderer/mod.rs:4525-4841) against any object with the wrcu device methods;
the product device is `webrender_b200.device.CudaDevice`.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import abi
from .gpu_types import ortho


@dataclass
class TextureDesc:
    fmt: int
    width: int
    height: int
    data: Optional[np.ndarray] = None   # (h, w*bpp) uint8 or None
    filter: int = abi.LINEAR


@dataclass
class Clear:
    """clear_target (device/gl.rs:3779): rect None = whole target."""
    color: Optional[Tuple[float, float, float, float]] = None
    depth: Optional[float] = None
    rect: Optional[Tuple[int, int, int, int]] = None


@dataclass
class Batch:
    """PrimitiveBatch / clip batch (batch.rs:233-237): one instanced draw."""
    kind: int
    instances: np.ndarray            # (n, stride) uint8 or (n, k) int32/float32
    blend: int = abi.BLEND_NONE
    depth: int = abi.DEPTH_OFF
    features: int = 0
    color: Tuple[str, str, str] = ("", "", "")   # texture names for sColor0..2
    clip_mask: str = ""
    scissor: Optional[Tuple[int, int, int, int]] = None
    blend_color: Tuple[float, float, float, float] = (0, 0, 0, 0)

    def instance_bytes(self):
        a = np.ascontiguousarray(self.instances)
        n = a.shape[0]
        return a.view(np.uint8).reshape(n, -1)


@dataclass
class Target:
    """A render target of a pass (picture-cache tile, colour or alpha target)."""
    texture: str
    depth: str = ""                   # name of a DEPTH24 texture or ""
    ops: List[object] = field(default_factory=list)   # Clear | Batch in order


@dataclass
class Frame:
    tables: Dict[str, np.ndarray]
    textures: Dict[str, TextureDesc]
    passes: List[List[Target]]


def draw_frame(dev, frame: Frame, handles: Optional[Dict[str, int]] = None, tile_lists: bool = False):
    """Renderer::draw_frame restated over the wrcu device calls.  Returns the
    name → texture handle map (textures are created on first use).  tile_lists: submit runs of
    composite batches through wrcu_draw_composite_tiles (what the C++ host's draw_tile_list does)."""
    handles = {} if handles is None else handles
    for name, t in frame.textures.items():
        if name not in handles:
            handles[name] = dev.texture_create(t.fmt, t.width, t.height)
            dev.texture_set_filter(handles[name], t.filter)
            if t.data is not None:
                dev.texture_upload(handles[name], 0, 0, t.width, t.height, t.data)
    dev.frame_begin(frame.tables)
    for rpass in frame.passes:
        for tgt in rpass:
            desc = frame.textures[tgt.texture]
            dev.target_bind(handles[tgt.texture], handles.get(tgt.depth, 0) if tgt.depth else 0,
                            ortho(desc.width, desc.height), (0, 0, desc.width, desc.height))
            ops = _fuse_tile_lists(tgt.ops) if tile_lists else tgt.ops
            for op in ops:
                if isinstance(op, Clear):
                    dev.clear(op.rect, op.color, op.depth)
                elif isinstance(op, _TileList):
                    b = op.batches[0]
                    inst = np.concatenate([x.instance_bytes() for x in op.batches])
                    tex = [handles[x.color[0]] for x in op.batches for _ in range(x.instance_bytes().shape[0])]
                    dev.draw_composite_tiles(b.features, b.blend, b.scissor, b.blend_color, inst, tex)
                else:
                    dev.draw_batch(op.kind, op.features, op.blend, op.depth,
                                   [handles.get(n, 0) if n else 0 for n in op.color],
                                   handles.get(op.clip_mask, 0) if op.clip_mask else 0,
                                   op.scissor, op.blend_color, op.instance_bytes())
    dev.frame_end()
    return handles


class _TileList:
    def __init__(self, batches):
        self.batches = batches


def _fuse_tile_lists(ops):
    """draw_tile_list's grouping for wrcu_draw_composite_tiles: consecutive RGBA composite batches with the
    same shader parameters and blend state become one submission, whatever their textures."""
    from . import abi
    out = []
    for op in ops:
        ok = (not isinstance(op, Clear) and op.kind == abi.KIND_COMPOSITE and not (op.features & abi.FEAT_YUV)
              and not op.clip_mask)
        if ok and out and isinstance(out[-1], _TileList):
            p = out[-1].batches[0]
            if (p.features, p.blend, p.scissor, tuple(p.blend_color)) == (op.features, op.blend, op.scissor, tuple(op.blend_color)):
                out[-1].batches.append(op)
                continue
        out.append(_TileList([op]) if ok else op)
    return out
