"""Shared driver for the update-path tests: the same
two-frame sequence on any device object.

Frame 1: the glyph/image atlas arrives as a batch of tile-sized rect uploads out
of one staging blob (update_texture_cache → upload_to_texture_cache), the GPU
cache as a GpuCacheUpdateList of Copy records (GpuCacheTexture::update), and the
frame is drawn with tables.gpu_cache = NULL, i.e. from the persistent cache.
Frame 2: a second update list patches some blocks (new colours), part of the
atlas is moved inside the texture with a texture-to-texture copy through a
scratch texture (texture-cache defragmentation), and the frame is drawn again.
"""
import numpy as np

from webrender_b200 import abi
from workloads import scenes
from webrender_b200.frame import draw_frame


def _cache_update_list(cache, rng):
    """Split a flat gpu_cache array into Copy records of 1..24 blocks that do not
    cross rows, in shuffled order, with the block array shuffled to match."""
    n = len(cache)
    spans, a = [], 0
    while a < n:
        c = int(min(rng.randint(1, 25), n - a, 1024 - (a % 1024)))
        spans.append((a, c))
        a += c
    order = rng.permutation(len(spans))
    blocks, updates = [], []
    for i in order:
        addr, cnt = spans[i]
        updates.append((len(blocks), cnt, addr % 1024, addr // 1024))
        blocks.extend(cache[addr:addr + cnt])
    return updates, np.asarray(blocks, np.float32).reshape(-1, 4)


def _tile_uploads(data, bpp, rng, tile=64):
    """Cut an image into tiles and pack them, shuffled, into one staging blob."""
    h, wb = data.shape
    w = wb // bpp
    tiles = [(x, y, min(tile, w - x), min(tile, h - y)) for y in range(0, h, tile) for x in range(0, w, tile)]
    order = rng.permutation(len(tiles))
    blob, rects = [], []
    off = 0
    for i in order:
        x, y, tw, th = tiles[i]
        stride = tw * bpp + int(rng.choice([0, 4, 16]))   # padded rows, as a PBO's aligned stride
        buf = np.zeros((th, stride), np.uint8)
        buf[:, : tw * bpp] = data[y:y + th, x * bpp:(x + tw) * bpp]
        off = (off + 15) & ~15
        rects.append((x, y, tw, th, off, stride))
        blob.append((off, buf.reshape(-1)))
        off += buf.size
    staging = np.zeros(off, np.uint8)
    for o, b in blob:
        staging[o:o + b.size] = b
    return rects, staging


def run_sequence(dev, kind="text", seed=1):
    """Returns the target bytes after frame 1 and after frame 2."""
    rng = np.random.RandomState(seed)
    if kind == "text":
        frame = scenes.text_frame(seed=seed, width=480, height=270, n_runs=8, glyphs_per_run=20)
        atlas_name = "atlas"
    else:
        frame = scenes.image_frame(seed=seed, one_to_one=True)
        atlas_name = "atlas"
    assert atlas_name in frame.textures, sorted(frame.textures)
    desc = frame.textures[atlas_name]
    bpp = abi.FMT_BPP[desc.fmt]
    data = np.ascontiguousarray(desc.data).view(np.uint8).reshape(desc.height, -1)[:, : desc.width * bpp]
    # texture arrives through batched uploads
    handles = {atlas_name: dev.texture_create(desc.fmt, desc.width, desc.height)}
    dev.texture_set_filter(handles[atlas_name], desc.filter)
    rects, staging = _tile_uploads(data, bpp, rng)
    dev.texture_upload_batch(handles[atlas_name], rects, staging)
    # GPU cache arrives as an update list
    cache = np.ascontiguousarray(frame.tables["gpu_cache"], np.float32).reshape(-1, 4)
    height = max(1, (len(cache) + 1023) // 1024)
    updates, blocks = _cache_update_list(cache, rng)
    dev.gpu_cache_update(height, True, updates, blocks)
    tables = dict(frame.tables)
    tables["gpu_cache"] = None
    f1 = type(frame)(tables, frame.textures, frame.passes)
    draw_frame(dev, f1, handles)
    tdesc = frame.textures["target"]
    out1 = dev.read_pixels(handles["target"], 0, 0, tdesc.width, tdesc.height, abi.FMT_BPP[tdesc.fmt]).copy()
    # frame 2: patch a few cache blocks (scale colours) — the list only carries the changed blocks
    changed = sorted(rng.choice(len(cache), size=min(12, len(cache)), replace=False))
    patch = cache[changed] * np.float32(0.5)
    dev.gpu_cache_update(height + 1, False, [(i, 1, a % 1024, a // 1024) for i, a in enumerate(changed)], patch)
    # ...and move a 64x64 block of the atlas through a scratch texture and back, shifted by 0: net no-op on content,
    # then copy one tile over its neighbour (content change both backends must agree on)
    scratch = dev.texture_create(desc.fmt, 64, 64)
    dev.texture_copy(handles[atlas_name], scratch, (0, 0, 64, 64), 0, 0)
    if desc.width >= 128:
        dev.texture_copy(scratch, handles[atlas_name], (0, 0, 64, 64), 64, 0)
    draw_frame(dev, f1, handles)
    out2 = dev.read_pixels(handles["target"], 0, 0, tdesc.width, tdesc.height, abi.FMT_BPP[tdesc.fmt]).copy()
    atlas = dev.read_pixels(handles[atlas_name], 0, 0, desc.width, desc.height, bpp).copy()
    return out1, out2, atlas
