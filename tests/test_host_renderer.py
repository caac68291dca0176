"""The C++ host mirror of the reference's Renderer (webrender_b200/host/): the
library must load and export its entry points on any box; on the GPU the
reference-shaped frame (passes → targets → batch containers → composite) drawn
by wr::Renderer::render must equal the oracle's rendering of the same scene."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle.backends import OracleDevice
from webrender_b200 import abi
from workloads import scenes

from common import render

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_LIB = os.path.join(ROOT, "webrender_b200", "libwrhost.so")

SYMBOLS = ["wrh_renderer_create", "wrh_renderer_destroy", "wrh_frame_create", "wrh_frame_destroy", "wrh_frame_add_pass",
           "wrh_pass_add_picture_cache_target", "wrh_pass_add_color_target", "wrh_pass_add_alpha_target",
           "wrh_pass_add_texture_cache_target", "wrh_texture_cache_target_add_clear", "wrh_texture_cache_target_add_tasks",
           "wrh_picture_target_add_batch", "wrh_color_target_add_batch", "wrh_alpha_target_add_clear",
           "wrh_alpha_target_add_clips", "wrh_target_add_blur_or_scale", "wrh_frame_set_framebuffer", "wrh_frame_add_composite_tile", "wrh_frame_add_composite_yuv_tile",
           "wrh_renderer_render", "wrh_renderer_last_error",
           "wrh_renderer_queue_gpu_cache_updates", "wrh_renderer_queue_texture_update", "wrh_renderer_queue_texture_copy"]


def test_host_library_exports():
    lib = C.CDLL(HOST_LIB)
    for s in SYMBOLS:
        assert hasattr(lib, s), s


CASES = [
    ("brush_solid", lambda: scenes.brush_solid_frame(seed=1), ["target"]),
    ("brush_solid_rotated", lambda: scenes.brush_solid_frame(seed=2, rotate=17.0, fractional=True, with_masks=False), ["target"]),
    ("alpha_rects", lambda: scenes.alpha_rects_frame(640, 360, 50, random_rects=True, seed=4), ["target"]),
    ("clip_masks", lambda: scenes.clip_mask_frame(seed=2, fractional=True), ["mask"]),
    ("box_shadows", lambda: scenes.box_shadow_frame(seed=1), ["mask"]),
    ("rounded_rects_indirect", lambda: scenes.rounded_rects_frame(seed=1), ["target"]),
    ("images_one_to_one", lambda: scenes.image_frame(seed=1, one_to_one=True), ["target"]),
    ("text", lambda: scenes.text_frame(seed=2, width=480, height=270, n_runs=8, glyphs_per_run=20), ["target"]),
    ("gradients", lambda: scenes.gradient_frame(seed=1, blend=abi.BLEND_PREMULTIPLIED_ALPHA), ["target"]),
    ("composite", lambda: scenes.composite_frame(seed=1), ["fb"]),
    ("page", lambda: scenes.page_frame(width=2048, height=1024, seed=2), ["mask", "tile0", "tile3", "fb"]),
    ("composite_yuv_planar", lambda: scenes.yuv_composite_frame("planar", 2, seed=1), ["fb"]),
    ("composite_yuv_nv12", lambda: scenes.yuv_composite_frame("nv12", 0, seed=2, opaque=False), ["fb"]),
    ("brush_yuv_image", lambda: scenes.yuv_image_frame("nv12", 2, seed=1), ["target"]),
    ("blur_a8", lambda: scenes.blur_frame(seed=1), ["mid", "target"]),
    ("blur_rgba8", lambda: scenes.blur_frame(seed=2, color=True), ["mid", "target"]),
    ("scale", lambda: scenes.scale_frame(seed=1), ["target"]),
    ("quad_radial_gradients", lambda: scenes.quad_gradient_frame(abi.KIND_QUAD_RADIAL_GRADIENT, seed=2), ["target"]),
    ("quad_conic_gradients", lambda: scenes.quad_gradient_frame(abi.KIND_QUAD_CONIC_GRADIENT, seed=1, fractional=True), ["target"]),
    ("texture_cache_target", lambda: scenes.texture_cache_frame(seed=1), ["target"]),
    ("texture_cache_linear_gradients", lambda: scenes.cached_gradient_frame(abi.KIND_LINEAR_GRADIENT, seed=2), ["target"]),
    ("texture_cache_conic_gradients", lambda: scenes.cached_gradient_frame(abi.KIND_CONIC_GRADIENT, seed=1), ["target"]),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,make,targets", CASES, ids=[c[0] for c in CASES])
def test_host_renderer_matches_oracle(name, make, targets):
    from webrender_b200.device import CudaDevice
    from webrender_b200.host import HostRenderer
    frame = make()
    dev = CudaDevice(0)
    hr = HostRenderer(dev)
    try:
        handles, calls = hr.render(frame)
        assert calls > 0
        want = render(OracleDevice, frame, targets)
        for t in targets:
            d = frame.textures[t]
            bpp = 4 if d.fmt == abi.FMT_RGBA8 else 1
            got = dev.read_pixels(handles[t], 0, 0, d.width, d.height, bpp)
            assert np.array_equal(got, want[t]), f"{name}/{t}: {(got != want[t]).sum()} bytes differ"
    finally:
        hr.close()
        dev.close()


@pytest.mark.gpu
def test_host_renderer_update_path():
    """Renderer::update_texture_cache / update_gpu_cache ahead of draw_frame: the atlas arrives as
    queued TextureCacheUpdates, the GPU cache as a GpuCacheUpdateList, a second frame patches both."""
    from webrender_b200.device import CudaDevice
    from webrender_b200.host import HostRenderer
    from update_path import _cache_update_list
    rng = np.random.RandomState(5)
    frame = scenes.text_frame(seed=2, width=480, height=270, n_runs=8, glyphs_per_run=20)
    desc = frame.textures["atlas"]
    bpp = abi.FMT_BPP[desc.fmt]
    data = np.ascontiguousarray(desc.data).view(np.uint8).reshape(desc.height, -1)[:, : desc.width * bpp]
    cache = np.ascontiguousarray(frame.tables["gpu_cache"], np.float32).reshape(-1, 4)
    height = max(1, (len(cache) + 1023) // 1024)
    dev, odev = CudaDevice(0), OracleDevice()
    hr = HostRenderer(dev)
    try:
        handles = {"atlas": dev.texture_create(desc.fmt, desc.width, desc.height)}
        dev.texture_set_filter(handles["atlas"], desc.filter)
        tile = 96
        for y in range(0, desc.height, tile):
            for x in range(0, desc.width, tile):
                w, h = min(tile, desc.width - x), min(tile, desc.height - y)
                hr.queue_texture_update(handles["atlas"], (x, y, x + w, y + h), data[y:y + h, x * bpp:], bpp)
        updates, blocks = _cache_update_list(cache, rng)
        hr.queue_gpu_cache_updates(height, True, updates, blocks)
        tables = dict(frame.tables)
        tables["gpu_cache"] = None
        f1 = type(frame)(tables, frame.textures, frame.passes)
        handles, _ = hr.render(f1, handles)
        want = render(OracleDevice, frame, ["target"])["target"]
        d = frame.textures["target"]
        got = dev.read_pixels(handles["target"], 0, 0, d.width, d.height, 4)
        assert np.array_equal(got, want)
        # frame 2: halve a few cached colours; only the changed blocks travel
        changed = sorted(rng.choice(len(cache), size=10, replace=False))
        patched = cache.copy()
        patched[changed] *= np.float32(0.5)
        hr.queue_gpu_cache_updates(height, False, [(i, 1, a % 1024, a // 1024) for i, a in enumerate(changed)],
                                   patched[changed])
        handles, _ = hr.render(f1, handles)
        t2 = dict(frame.tables)
        t2["gpu_cache"] = patched
        want2 = render(OracleDevice, type(frame)(t2, frame.textures, frame.passes), ["target"])["target"]
        got2 = dev.read_pixels(handles["target"], 0, 0, d.width, d.height, 4)
        assert np.array_equal(got2, want2)
    finally:
        hr.close()
        dev.close()
        odev.close()
