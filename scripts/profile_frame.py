#!/usr/bin/env python
"""Renders the bench workload (C4: dragon full-res, 3840x2160, 16 spp) `frames` times with device-resident
offsets and prints the stats of the last frame.  Used under ncu: every frame launches exactly
chunks*(2*passes+1) kernels in a fixed order (passes = maxDepth + 1, or 2 maxDepth + 1 with a Transparent material), so `-k regex:shadow_kernel -s N -c 1` picks a known chunk."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import rayhs_b200 as rh
from rayhs_b200 import capi

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=2)
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--spp", type=int, default=16)
ap.add_argument("--pack", default="dragon_full")
ap.add_argument("--count", action="store_true")
ap.add_argument("--chunk", type=int, default=0)
ap.add_argument("--no-profile", action="store_true", help="no per-launch events: lets two chunks be in flight")
ap.add_argument("--shadow", default="pooled", choices=["auto", "pooled", "split"],
                help="shadow-ray schedule; forced by default so that every frame launches the same kernels")
a = ap.parse_args()
rh.init(0)
L = capi.lib()
sc = rh.Scene.from_pack(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", a.pack + ".pack"))
job = rh.renderingFromScene(sc, a.width, a.height)
off = torch.empty((a.width * a.height, a.spp, 2), dtype=torch.float64)
L.rh_sample_offsets_f64(24, a.width * a.height, a.spp, off.data_ptr())
off_dev = off.cuda()
rgb = torch.empty((a.height, a.width, 3), dtype=torch.uint8, device="cuda")
for i in range(a.frames):
    st = rh.render_device(job, rgb, spp=a.spp, offsets_dev=off_dev, profile=not a.no_profile, count=a.count, chunk_samples=a.chunk,
                          shadow=None if a.shadow == "auto" else a.shadow)
print(json.dumps(st))
