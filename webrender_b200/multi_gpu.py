"""Sharding one frame's picture-cache tiles across GPUs.

Every picture-cache tile is its own render target with its own batches
(frame_builder.rs:995-1057); tiles never read each other within a frame.  So a
frame that spans many tiles shards by tile: rank r renders tiles r, r+W, r+2W…
on its own GPU with no data-path communication, then ONE exchange step moves
the finished tiles to the rank that runs `composite` (renderer/mod.rs:3340):
a gather of ≤2 MiB tiles over NCCL (NVLink/NVSwitch), after which rank 0
composites them into the framebuffer exactly as a single GPU would.

One process per GPU; `torch.distributed` supplies the process group.  With the
NCCL backend the tiles go GPU→GPU from device memory; with gloo (CPU tests of
the host logic) the same gather runs on host copies of the tiles.
"""
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

from . import abi
from .frame import Batch, Clear, Frame, Target, TextureDesc, draw_frame
from .gpu_types import (FrameTables, INVALID_SEGMENT_INDEX, PART_ALL, QF_APPLY_DEVICE_CLIP, composite_instance,
                        ortho, quad_instance)


@dataclass
class TiledScene:
    """A frame cut into picture-cache tiles: one Frame per tile (rendering into
    a texture named "tile"), the tile's framebuffer rect, and the size of the
    framebuffer the tiles are composited into."""
    width: int
    height: int
    tile_w: int
    tile_h: int
    tiles: List[Frame]
    rects: List[Tuple[int, int, int, int]]
    pixel_layers: int


def assign_tiles(n_tiles: int, world: int) -> List[List[int]]:
    """Static round-robin tile → rank map (tiles are uniform cost in config E)."""
    return [list(range(r, n_tiles, world)) for r in range(world)]


def tiled_alpha_scene(width=8192, height=4096, tile_w=1024, tile_h=512, n_rects=1000, seed=3) -> TiledScene:
    """Config E: the config-B' instance list — n_rects seeded
    random alpha rects over the whole frame — cut per tile: each tile gets the
    rects that touch it, as Quad(ColorOrTexture) instances in a picture task
    whose content origin is the tile origin (batch.rs per-tile batchers)."""
    rng = np.random.RandomState(seed)
    rects = []
    for _ in range(n_rects):
        w, h = rng.randint(64, 1025, 2)
        x, y = rng.randint(-32, width - 32), rng.randint(-32, height - 32)
        col = rng.uniform(0.02, 0.3)
        rects.append((float(x), float(y), float(x + w), float(y + h), float(col)))
    tiles, trects, layers = [], [], 0
    for ty in range(0, height, tile_h):
        for tx in range(0, width, tile_w):
            tw, th = min(tile_w, width - tx), min(tile_h, height - ty)
            t = FrameTables()
            task = t.add_render_task((0.0, 0.0, float(tw), float(th)), 1.0, (float(tx), float(ty)))
            inst = []
            z = 1
            for (x0, y0, x1, y1, c) in rects:
                ix0, iy0, ix1, iy1 = max(x0, tx), max(y0, ty), min(x1, tx + tw), min(y1, ty + th)
                if ix1 <= ix0 or iy1 <= iy0:
                    continue
                layers += int(ix1 - ix0) * int(iy1 - iy0)
                prim_f = t.add_quad_prim((x0, y0, x1, y1), (x0, y0, x1, y1), (c, c, c, c))
                prim_i = t.add_quad_header(0, z)
                inst.append(quad_instance(prim_i, prim_f, QF_APPLY_DEVICE_CLIP, 0, PART_ALL, INVALID_SEGMENT_INDEX, task))
                z += 1
            ops = [Clear(color=(0.3, 0.0, 0.0, 1.0))]
            if inst:
                ops.append(Batch(abi.KIND_QUAD_TEXTURED, np.stack(inst), blend=abi.BLEND_PREMULTIPLIED_ALPHA))
            tiles.append(Frame(t.arrays(), {"tile": TextureDesc(abi.FMT_RGBA8, tw, th, filter=abi.NEAREST)},
                               [[Target("tile", ops=ops)]]))
            trects.append((tx, ty, tx + tw, ty + th))
    return TiledScene(width, height, tile_w, tile_h, tiles, trects, layers)


def composite_ops(scene: TiledScene, names: List[str]) -> List[object]:
    """composite_simple for opaque picture-cache tiles: clear, then one
    FAST_PATH composite instance per tile, blending off."""
    ops = [Clear(color=(0.0, 0.0, 0.0, 0.0))]
    for name, r in zip(names, scene.rects):
        rf = tuple(float(v) for v in r)
        ops.append(Batch(abi.KIND_COMPOSITE, composite_instance(rf, rf)[None, :], blend=abi.BLEND_NONE,
                         features=abi.FEAT_FAST_PATH | abi.FEAT_TEXTURE_2D, color=(name, "", "")))
    return ops


def rank_frames(scene: TiledScene, mine: List[int], composite_clear=None):
    """The two frames a rank draws for its share `mine` of the scene's tiles, the way one wr::Frame holds
    many picture-cache targets: (1) ONE frame whose single pass has a PictureCacheTarget per tile, all
    sharing one set of per-frame tables (one table upload per rank and frame instead of one per tile);
    (2) the `composite` frame: one tile list for the rank's tiles into the texture named "fb"
    (composite_simple, renderer/mod.rs:3340-3484; `composite_clear` = the framebuffer clear colour, or
    None when another context clears it)."""
    from .frame import Target as _T
    t = FrameTables()
    textures, targets = {}, []
    off = {k: 0 for k in ("prim_headers_f", "prim_headers_i", "transforms", "render_tasks", "gpu_cache",
                          "gpu_buffer_f", "gpu_buffer_i")}
    merged = {k: [] for k in off}
    for i in mine:
        f = scene.tiles[i]
        name = "tile%d" % i
        textures[name] = f.textures["tile"]
        tgt = f.passes[0][0]
        ops = []
        # re-base the tile's table addresses onto the merged tables (quad instances carry three of them:
        # prim_address_i → gpu_buffer_i, prim_address_f → gpu_buffer_f, render task address)
        base_i, base_f, base_t = off["gpu_buffer_i"], off["gpu_buffer_f"], off["render_tasks"] // 2
        for op in tgt.ops:
            if isinstance(op, Batch):
                inst = op.instances.copy()
                inst[:, 0] += base_i
                inst[:, 1] += base_f
                inst[:, 3] += base_t
                ops.append(Batch(op.kind, inst, blend=op.blend, depth=op.depth, features=op.features, color=op.color,
                                 clip_mask=op.clip_mask))
            else:
                ops.append(op)
        targets.append(_T(name, ops=ops))
        for k in off:
            arr = f.tables[k]
            if k == "transforms" and off[k]:
                continue  # every tile frame carries the identity palette entry only: keep one copy
            merged[k].append(arr)
            off[k] += len(arr)
    tables = {}
    for k, lst in merged.items():
        dt = np.int32 if k.endswith("_i") else np.float32
        tables[k] = np.ascontiguousarray(np.concatenate(lst)) if lst else np.zeros((0, 4), dt)
    tiles_frame = Frame(tables, textures, [targets])
    names = ["tile%d" % i for i in mine]
    ops = [Clear(color=composite_clear)] if composite_clear is not None else []
    for name, i in zip(names, mine):
        rf = tuple(float(v) for v in scene.rects[i])
        ops.append(Batch(abi.KIND_COMPOSITE, composite_instance(rf, rf)[None, :], blend=abi.BLEND_NONE,
                         features=abi.FEAT_FAST_PATH | abi.FEAT_TEXTURE_2D, color=(name, "", "")))
    ctex = {"fb": TextureDesc(abi.FMT_RGBA8, scene.width, scene.height)}
    ctex.update({n: textures[n] for n in names})
    comp_frame = Frame(FrameTables().arrays(), ctex, [[Target("fb", ops=ops)]])
    return tiles_frame, comp_frame


class DirectShardedRenderer:
    """One frame's tiles sharded over GPUs with NO gather: the compositing context (rank 0) exports its
    framebuffer, every other context imports it (CUDA IPC across processes, raw mapping inside one
    process) and runs `composite` for its own tiles with that framebuffer as the render target — the copy
    kernel's bulk-tensor stores go over NVLink straight into rank 0's memory.  Ordering is by flag words
    polled on the CUDA streams (wrcu_peer_signal / wrcu_peer_wait); the host never synchronises inside
    a frame.  Per frame and rank the host makes two native Renderer::render calls (C++ host mirror).

    Set-up plumbing, once: every rank publishes `self.blob` (a few hundred bytes: the framebuffer and flag
    handles) and calls connect() with the list of all ranks' blobs — torch.distributed.all_gather_object
    between processes, a plain list in-process."""

    CLEAR = (0.0, 0.0, 0.0, 0.0)

    def __init__(self, dev, scene: TiledScene, rank, world):
        from .host import HostRenderer
        self.dev, self.scene, self.rank, self.world = dev, scene, rank, world
        self.mine = assign_tiles(len(scene.tiles), world)[rank]
        self.frame_no = 0
        self.hr = HostRenderer(dev)
        fb_blob = b""
        if rank == 0:
            self.fb = dev.texture_create(abi.FMT_RGBA8, scene.width, scene.height)
            fb_blob = dev.texture_export(self.fb)
        self.blob = (fb_blob, dev.peer_flags_create(world + 1))

    def connect(self, blobs):
        dev, rank, world, scene = self.dev, self.rank, self.world, self.scene
        if rank != 0:
            self.fb = dev.texture_import(blobs[0][0])
            self.peer0 = dev.peer_flags_open(blobs[0][1])
            self.peers = {}
        else:
            self.peers = {r: dev.peer_flags_open(blobs[r][1]) for r in range(1, world)}
        tiles_frame, comp_frame = rank_frames(scene, self.mine)
        self.nf_tiles = self.hr.build(tiles_frame)
        handles = {"fb": self.fb}
        handles.update({n: h for n, h in self.nf_tiles.handles.items() if n.startswith("tile")})
        self.nf_comp = self.hr.build(comp_frame, handles)
        self.proj = ortho(scene.width, scene.height)

    # the frame is queued in three steps so several in-process contexts can be driven from one thread
    def begin(self):
        """rank 0: clear the framebuffer, then let the others composite into it"""
        self.frame_no += 1
        if self.rank == 0:
            s = self.scene
            self.dev.target_bind(self.fb, 0, self.proj, (0, 0, s.width, s.height))
            self.dev.clear(None, self.CLEAR, None)
            for r, pid in self.peers.items():
                self.dev.peer_signal(pid, 0, self.frame_no)

    def draw(self):
        """every rank: its tiles, then its tile list into the (shared) framebuffer"""
        if self.mine:
            self.hr.render_native(self.nf_tiles)
        if self.rank != 0:
            self.dev.peer_wait(0, self.frame_no)          # rank 0 has cleared this frame's framebuffer
        if self.mine:
            self.hr.render_native(self.nf_comp)
        if self.rank != 0:
            self.dev.peer_signal(self.peer0, self.rank, self.frame_no)

    def end(self):
        """rank 0: the frame is complete once every other rank's stores have landed"""
        if self.rank == 0:
            for r in self.peers:
                self.dev.peer_wait(r, self.frame_no)

    def render(self):
        self.begin()
        self.draw()
        self.end()

    def read_framebuffer(self):
        assert self.rank == 0
        return self.dev.read_pixels(self.fb, 0, 0, self.scene.width, self.scene.height, 4)

    def close(self):
        self.nf_comp.destroy()
        self.nf_tiles.destroy()
        self.hr.close()


class _DevMem:
    """__cuda_array_interface__ view of wrcu texture memory for torch."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _texture_tensor(dev, tex, w, h, bpp):
    import torch
    ptr, pitch = dev.texture_device_ptr(tex)
    flat = torch.as_tensor(_DevMem(ptr, pitch * h), device="cuda")
    return flat.view(h, pitch)[:, : w * bpp]


class ShardedRenderer:
    """Renders a TiledScene with tiles sharded over the ranks of a process
    group; rank 0 ends up with the composited framebuffer."""

    def __init__(self, dev, scene: TiledScene, rank=0, world=1, group=None, device_exchange=None):
        self.dev, self.scene, self.rank, self.world, self.group = dev, scene, rank, world, group
        self.mine = assign_tiles(len(scene.tiles), world)[rank]
        self.per_rank = (len(scene.tiles) + world - 1) // world
        # NCCL moves tiles GPU→GPU; gloo (CPU tests) moves host copies
        self.device_exchange = hasattr(dev, "texture_device_ptr") if device_exchange is None else device_exchange
        self.tile_handles: Dict[int, Dict[str, int]] = {}
        self.fb = None
        self.fb_handles: Dict[str, int] = {}
        self.stage = None
        self.gathered = None

    # -- step 1: every rank renders its own tiles (no communication) ---------------
    def render_tiles(self):
        for i in self.mine:
            self.tile_handles[i] = draw_frame(self.dev, self.scene.tiles[i], self.tile_handles.get(i))

    # -- step 2: the one exchange step: finished tiles → rank 0 --------------------
    def exchange(self):
        s = self.scene
        if self.world == 1:
            return
        import torch
        import torch.distributed as dist
        row = s.tile_w * 4
        if self.device_exchange:
            if self.stage is None:
                self.stage = torch.zeros((self.per_rank, s.tile_h, row), dtype=torch.uint8, device="cuda")
                if self.rank == 0:
                    self.gathered = [torch.empty_like(self.stage) for _ in range(self.world)]
            self.dev.finish()
            for k, i in enumerate(self.mine):
                x0, y0, x1, y1 = s.rects[i]
                src = _texture_tensor(self.dev, self.tile_handles[i]["tile"], x1 - x0, y1 - y0, 4)
                self.stage[k, : y1 - y0, : (x1 - x0) * 4].copy_(src)
            dist.gather(self.stage, self.gathered if self.rank == 0 else None, dst=0, group=self.group)
            if self.rank == 0:
                for r in range(1, self.world):
                    for k, i in enumerate(assign_tiles(len(s.tiles), self.world)[r]):
                        x0, y0, x1, y1 = s.rects[i]
                        h = self._ensure_tile(i)
                        dst = _texture_tensor(self.dev, h, x1 - x0, y1 - y0, 4)
                        dst.copy_(self.gathered[r][k, : y1 - y0, : (x1 - x0) * 4])
                torch.cuda.synchronize()
        else:
            stage = torch.zeros((self.per_rank, s.tile_h, row), dtype=torch.uint8)
            for k, i in enumerate(self.mine):
                x0, y0, x1, y1 = s.rects[i]
                px = self.dev.read_pixels(self.tile_handles[i]["tile"], 0, 0, x1 - x0, y1 - y0, 4)
                stage[k, : y1 - y0, : (x1 - x0) * 4] = torch.from_numpy(px)
            gathered = [torch.empty_like(stage) for _ in range(self.world)] if self.rank == 0 else None
            dist.gather(stage, gathered, dst=0, group=self.group)
            if self.rank == 0:
                for r in range(1, self.world):
                    for k, i in enumerate(assign_tiles(len(s.tiles), self.world)[r]):
                        x0, y0, x1, y1 = s.rects[i]
                        h = self._ensure_tile(i)
                        self.dev.texture_upload(h, 0, 0, x1 - x0, y1 - y0,
                                                gathered[r][k, : y1 - y0, : (x1 - x0) * 4].contiguous().numpy())

    def _ensure_tile(self, i):
        if i not in self.tile_handles:
            x0, y0, x1, y1 = self.scene.rects[i]
            h = self.dev.texture_create(abi.FMT_RGBA8, x1 - x0, y1 - y0)
            self.dev.texture_set_filter(h, abi.NEAREST)
            self.tile_handles[i] = {"tile": h}
        return self.tile_handles[i]["tile"]

    # -- step 3: rank 0 composites the tiles into the framebuffer ------------------
    def composite(self):
        if self.rank != 0:
            return
        s = self.scene
        names = ["tile%d" % i for i in range(len(s.tiles))]
        if self.fb is None:
            textures = {"fb": TextureDesc(abi.FMT_RGBA8, s.width, s.height)}
            for i, n in enumerate(names):
                x0, y0, x1, y1 = s.rects[i]
                textures[n] = TextureDesc(abi.FMT_RGBA8, x1 - x0, y1 - y0, filter=abi.NEAREST)
                self.fb_handles[n] = self._ensure_tile(i)
            self.fb = Frame(FrameTables().arrays(), textures, [[Target("fb", ops=composite_ops(s, names))]])
        self.fb_handles = draw_frame(self.dev, self.fb, self.fb_handles)

    def render(self):
        self.render_tiles()
        self.exchange()
        self.composite()

    def read_framebuffer(self):
        assert self.rank == 0
        return self.dev.read_pixels(self.fb_handles["fb"], 0, 0, self.scene.width, self.scene.height, 4)
