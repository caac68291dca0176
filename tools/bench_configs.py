#!/usr/bin/env python
"""Informational timings of the other SURVEY.md §8d configurations on one GPU
(the contract's bench line is config B, bench.py).  Prints one JSON object per
config: ms/frame (CUDA events on the wrcu stream, L2 flushed), fps, Mpix/s where
the unit applies.  Usage: python tools/bench_configs.py [--steps 10]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--only", default="", help="substring of the config name to run alone")
    args = ap.parse_args()
    import torch
    from webrender_b200 import abi, multi_gpu
    from workloads import scenes
    from webrender_b200.device import CudaDevice
    from webrender_b200.frame import draw_frame

    dev = CudaDevice(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    W, H = 3840, 2160
    cfgs = [
        ("B' 1000 random alpha rects 64-1024 px @4K", scenes.alpha_rects_frame(W, H, 1000, random_rects=True, seed=1)),
        ("B' varied colours/alphas (generic kernel)", scenes.alpha_rects_frame(W, H, 1000, random_rects=True, seed=1, color=None)),
        ("C text: 68 runs x 89 glyphs @4K, R8 atlas 2048^2",
         scenes.text_frame(width=W, height=H, n_runs=68, glyphs_per_run=89, seed=2, atlas_size=2048)),
        ("D gradients: 10 full-frame 2-stop linear gradients @4K (opaque)",
         scenes.gradient_frame(width=W, height=H, n_grad=10, full_frame=True)),
        ("D box shadow: one 1024^2 cs_clip_box_shadow instance",
         scenes.box_shadow_frame(width=1024, height=1024, n_clips=1, full_size=(1024, 1024), seed=7)),
        ("D clip rect: 10 rounded-rect masks 512x384 target", scenes.clip_mask_frame()),
        ("composite: 4x5 tiles 1024x512 -> 4K framebuffer", scenes.composite_frame(W, H, 1024, 512, seed=4)),
        ("images: 8 opaque + 20 alpha brush_image @4K-ish", scenes.image_frame(width=W, height=H, seed=1)),
    ]
    for name, frame in cfgs:
        if args.only and args.only not in name:
            continue
        handles = draw_frame(dev, frame)
        for _ in range(2):
            draw_frame(dev, frame, handles)
        dev.finish()
        ms = []
        for _ in range(args.steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            dev.timer_begin()
            draw_frame(dev, frame, handles)
            ms.append(dev.timer_end())
        ms.sort()
        med = ms[len(ms) // 2]
        out = {"config": name, "ms_per_frame": med, "fps": 1e3 / med}
        if name.startswith("B'"):
            layers = scenes.pixel_layers_of_quad_batch(frame)
            out["Mpix_s"] = layers / (med * 1e-3) / 1e6
        print(json.dumps(out), flush=True)
        for h in handles.values():
            dev.texture_destroy(h)
    dev.close()


if __name__ == "__main__":
    main()
