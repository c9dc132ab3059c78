#!/usr/bin/env python
"""One shard of the bench frame on ONE GPU, as a rank of an N-GPU run sees it (C4: dragon full-res, 3840x2160, 16 spp;
shard 0 of `--shards` interleaved row bands): device time of the render call against the wall time per call (host
overhead: launches, the stream synchronisation, the control-block read-back) and the per-kernel times of one frame.
The ideal is the single-GPU frame divided by the number of shards."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import rayhs_b200 as rh
from rayhs_b200 import capi

ap = argparse.ArgumentParser()
ap.add_argument("--shards", type=int, default=8)
ap.add_argument("--frames", type=int, default=30)
ap.add_argument("--chunk", type=int, default=0)
ap.add_argument("--shadow", default="pooled", choices=["pooled", "split"])
ap.add_argument("--count", action="store_true", help="one instrumented frame per shard count at the end (RAYHS_B200_DEBUG=1 prints the histograms)")
a = ap.parse_args()
rh.init(0)
L = capi.lib()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sc = rh.Scene.from_pack(os.path.join(root, "tests", "golden", "dragon_full.pack"))
W, H, spp = 3840, 2160, 16
job = rh.renderingFromScene(sc, W, H)
off = torch.empty((W * H, spp, 2), dtype=torch.float64)
L.rh_sample_offsets_f64(24, W * H, spp, off.data_ptr())
off_dev = off.cuda()
out = {}
for G in (1, a.shards):
    bh = L.rh_default_band_height(H, G)
    rows = L.rh_shard_rows(H, G, bh)
    rgb = torch.empty((rows, W, 3), dtype=torch.uint8, device="cuda")
    kw = dict(spp=spp, offsets_dev=off_dev, shard_index=0, shard_count=G, band_height=bh, shadow=a.shadow, chunk_samples=a.chunk)
    for _ in range(4):
        rh.render_device(job, rgb, **kw)
    torch.cuda.synchronize()
    t0 = time.time()
    dev = 0.0
    for _ in range(a.frames):
        dev += rh.render_device(job, rgb, **kw)["ms_total"]
    torch.cuda.synchronize()
    wall = 1e3 * (time.time() - t0) / a.frames
    st = rh.render_device(job, rgb, profile=True, **kw)
    if a.count:
        print(f"--- counted frame, {G} shard(s)", file=sys.stderr, flush=True)
        rh.render_device(job, rgb, count=True, **kw)
    out[f"shards_{G}"] = {"wall_ms_per_call": wall, "device_ms_per_call": dev / a.frames, "kernels_ms": st["ms_trace"] + st["ms_shadow"] + st["ms_resolve"],
                          "trace": st["ms_trace"], "shadow": st["ms_shadow"], "resolve": st["ms_resolve"], "chunks": st["chunks"]}
one, many = out["shards_1"], out[f"shards_{a.shards}"]
out["efficiency_device"] = one["device_ms_per_call"] / (a.shards * many["device_ms_per_call"])
out["efficiency_wall"] = one["wall_ms_per_call"] / (a.shards * many["wall_ms_per_call"])
print(json.dumps(out))
