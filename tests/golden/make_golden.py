"""Generates tests/golden/*.npz from the UNMODIFIED reference rasteriser
(oracle/_ref/libswgl_ref.so built from /root/reference/swgl/src/gl.cc).

Run here (needs /root/reference at build time):  python tests/golden/make_golden.py
Each case stores the scene-builder call (name + kwargs) and the reference's output
bytes; tests rebuild the frame from the call and compare oracle / CUDA output to
the stored bytes.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [
    ("alpha_rects_premult", "alpha_rects_frame", dict(width=301, height=157, n_rects=37, random_rects=True, seed=3)),
    ("alpha_rects_full_cover", "alpha_rects_frame", dict(width=256, height=64, n_rects=100)),
    ("alpha_rects_softlight", "alpha_rects_frame", dict(width=160, height=64, n_rects=24, random_rects=True, seed=122,
                                                         blend=22, color=None, clear_color=(0.4, 0.7, 0.2, 0.8))),
    ("alpha_rects_hue", "alpha_rects_frame", dict(width=160, height=64, n_rects=24, random_rects=True, seed=125,
                                                   blend=25, color=None, clear_color=(0.4, 0.7, 0.2, 0.8))),
    ("brush_solid_masks_depth", "brush_solid_frame", dict(width=333, height=207, seed=1)),
    ("brush_solid_aa_fractional", "brush_solid_frame", dict(width=333, height=207, seed=2, fractional=True, force_aa=True)),
    ("clip_rect_integer", "clip_mask_frame", dict(seed=1)),
    ("clip_rect_fractional", "clip_mask_frame", dict(seed=2, fractional=True)),
    ("clip_rect_scaled", "clip_mask_frame", dict(seed=3, fractional=True, scale=1.25)),
    # config A: wrench/reftests/aa/rounded-rects.yaml (also checked against the reference's own PNG,
    # tests/test_golden.py::test_config_a_against_reference_png)
    ("config_a_rounded_rects", "config_a_frame", dict()),
    ("image_scaled", "image_frame", dict(seed=2, fractional=True, n_opaque=0)),
    ("image_one_to_one_depth", "image_frame", dict(seed=3, one_to_one=True)),
    ("text_run_fractional", "text_frame", dict(seed=3, width=480, height=270, n_runs=8, glyphs_per_run=20, fractional=True)),
    ("linear_gradient_alpha", "gradient_frame", dict(seed=2, blend=2)),
    ("box_shadow_fractional", "box_shadow_frame", dict(seed=2, fractional=True)),
    ("composite_external", "composite_frame", dict(seed=2, external=True)),
    ("brush_opacity_scaled", "opacity_frame", dict(seed=2)),
    ("brush_blend_filters", "blend_frame", dict(seed=1)),
    ("brush_mix_blend_modes", "mix_blend_frame", dict(seed=2)),
    ("cs_blur_a8", "blur_frame", dict(seed=2)),
    ("cs_blur_rgba8", "blur_frame", dict(seed=1, color=True)),
    ("cs_scale_rgba8", "scale_frame", dict(seed=2)),
    ("rotated_brush_solid", "brush_solid_frame", dict(seed=2, rotate=-33.5, fractional=True)),
    ("rotated_gradient", "gradient_frame", dict(seed=2, rotate=17.0, fractional=True, blend=2)),
    ("ps_clear_depth", "clear_frame", dict(seed=1)),
    ("reftest_clip_mode", "reftest_clip_frame", dict(which="clip-mode")),
    ("reftest_clip_ellipse", "reftest_clip_frame", dict(which="clip-ellipse")),
    ("reftest_gradient_linear", "reftest_gradient_frame", dict(which="linear")),
    ("reftest_gradient_hard_stop", "reftest_gradient_frame", dict(which="linear-hard-stop")),
    # cached gradient render tasks (kind = wrcu_kind value)
    ("cs_fast_linear_gradient", "cached_gradient_frame", dict(kind=16, width=512, height=256, n_tasks=4, seed=2)),
    ("cs_linear_gradient", "cached_gradient_frame", dict(kind=17, width=512, height=256, n_tasks=4, seed=2)),
    ("cs_radial_gradient", "cached_gradient_frame", dict(kind=18, width=512, height=256, n_tasks=4, seed=4)),
    ("cs_radial_gradient_repeat", "cached_gradient_frame", dict(kind=18, width=512, height=256, n_tasks=4, seed=3,
                                                                 repeat=True, hard=True)),
    ("cs_conic_gradient", "cached_gradient_frame", dict(kind=19, width=512, height=256, n_tasks=4, seed=2)),
    ("ps_quad_radial_gradient", "quad_gradient_frame", dict(kind=23, seed=2, width=480, height=270, fractional=True)),
    ("ps_quad_conic_gradient", "quad_gradient_frame", dict(kind=24, seed=3, width=480, height=270, rotate=-12.0)),
    ("brush_image_repetition", "image_repeat_frame", dict(seed=2, n_opaque=0, width=480, height=270, fractional=True)),
    ("text_run_glyph_transform", "text_frame", dict(seed=2, width=480, height=270, n_runs=8, glyphs_per_run=16, fractional=True,
                                                     glyph_transform=(-33.0, 1.3, 0.9), clip_runs=True)),
    ("reftest_premultiplied_radial", "reftest_cached_gradient_frame", dict(which="premultiplied-radial")),
    ("reftest_conic_center", "reftest_cached_gradient_frame", dict(which="conic-center")),
    ("reftest_radial_circle", "reftest_cached_gradient_frame2", dict(name="radial-circle")),
    ("reftest_radial_ellipse", "reftest_cached_gradient_frame2", dict(name="radial-ellipse")),
    ("reftest_conic_simple", "reftest_cached_gradient_frame2", dict(name="conic-simple")),
    ("composite_yuv_planar_rec709", "yuv_composite_frame", dict(fmt="planar", color_space=2, seed=1)),
    ("composite_yuv_nv12_rec601_full", "yuv_composite_frame", dict(fmt="nv12", color_space=1, seed=2, fractional=True)),
    ("composite_yuv_interleaved_rec2020", "yuv_composite_frame", dict(fmt="interleaved", color_space=4, seed=3)),
    ("composite_yuv_planar_nearest_gbr", "yuv_composite_frame", dict(fmt="planar", color_space=6, seed=4, linear=False)),
    ("brush_yuv_image_nv12_alpha", "yuv_image_frame", dict(fmt="nv12", color_space=2, seed=1)),
    ("brush_yuv_image_planar_opaque", "yuv_image_frame", dict(fmt="planar", color_space=0, seed=2, alpha_pass=False)),
    ("brush_yuv_image_interleaved_rotated", "yuv_image_frame", dict(fmt="interleaved", color_space=3, seed=3, rotate=17.0)),
    # perspective (draw_perspective) and plane-split polygons: checked against the reference build directly; the
    # plain-C port (oracle/wr_oracle.c) does not restate this path ("port": False → CPU tier skips it)
    ("perspective_brush_solid_aa", "perspective_frame", dict(kind="solid", d=800.0, ry=35.0, rx=10.0, seed=4, force_aa=True,
                                                              n_opaque=6, n_alpha=12), False),
    ("perspective_brush_image_clipped", "perspective_frame", dict(kind="image", d=220.0, ry=60.0, rx=-30.0, seed=2), False),
    ("perspective_brush_opacity", "perspective_frame", dict(kind="opacity", d=220.0, ry=60.0, rx=-30.0, seed=2, brush_flags=1), False),
    ("perspective_brush_mix_blend", "perspective_frame", dict(kind="mix_blend", height=400, d=800.0, ry=-20.0, rx=15.0, seed=2), False),
    ("split_composite", "split_composite_frame", dict(seed=1), False),
    ("split_composite_near_plane", "split_composite_frame", dict(seed=2, d=220.0, ry=65.0, rx=20.0, perspective_interpolate=1),
     False),
    ("reftest_border_overlapping", "reftest_border_overlapping_frame", dict()),
    ("reftest_border_no_bogus_line", "reftest_border_no_bogus_line_frame", dict()),
    ("reftest_border_radii", "reftest_border_frame", dict(name="border-radii")),
    ("reftest_border_clamp_corner_radius", "reftest_border_frame", dict(name="border-clamp-corner-radius")),
    ("reftest_clip_inverted_ellipse", "reftest_clip_inverted_ellipse_frame", dict()),
    # wrench/reftests/split/near-plane.yaml (also checked against the reference's own PNG: 0 pixels differ)
    ("reftest_split_near_plane", "reftest_split_near_plane_frame", dict(), False),
    ("reftest_inset_no_blur_radius", "reftest_box_shadow_frame", dict(which="inset-no-blur-radius")),
    ("reftest_box_shadow_spread", "reftest_box_shadow_frame", dict(which="box-shadow-spread")),
    ("reftest_boxshadow_spread_only", "reftest_box_shadow_frame", dict(which="boxshadow-spread-only")),
    ("reftest_box_shadow_suite_no_blur", "reftest_box_shadow_frame", dict(which="suite-no-blur")),
    ("reftest_box_shadow_suite_composited", "reftest_box_shadow_suite_composited_frame", dict()),
    ("reftest_filter_small_blur_radius", "reftest_filter_blur_frame", dict()),
    ("reftest_line_decorations", "reftest_line_decorations_frame", dict()),
    ("reftest_image_segments", "reftest_image_segments_frame", dict()),
    ("reftest_linear_aligned_border_radius", "reftest_gradient_border_radius_frame", dict()),
    ("cs_line_decoration", "line_decoration_frame", dict(seed=2)),
    ("cs_border_solid", "border_frame", dict(kind=21, width=512, height=512, n_borders=3, seed=2)),
    ("cs_border_segment", "border_frame", dict(kind=22, width=768, height=512, n_borders=5, seed=3, scale=1.5)),
]


def main():
    from common import render
    from oracle.backends import SwglDevice
    from workloads import scenes
    index = {}
    regenerate_all = "--all" in sys.argv   # default: only cases whose fixture is missing
    for case in CASES:
        name, builder, kwargs = case[:3]
        port = case[3] if len(case) > 3 else True   # False: the plain-C port does not restate this path
        path = os.path.join(HERE, name + ".npz")
        if os.path.exists(path) and not regenerate_all:
            index[name] = {"builder": builder, "kwargs": kwargs, "targets": sorted(np.load(path).files), "port": port}
            continue
        frame = getattr(scenes, builder)(**kwargs)
        out = render(SwglDevice, frame)
        np.savez_compressed(path, **out)
        index[name] = {"builder": builder, "kwargs": kwargs, "targets": sorted(out), "port": port}
        print(name, {k: v.shape for k, v in out.items()})
    json.dump(index, open(os.path.join(HERE, "index.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
