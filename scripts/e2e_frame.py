#!/usr/bin/env python
"""End-to-end time of the bench frame on one GPU: pinned host offsets in (2.1 GB, uploaded chunk by chunk inside the call),
RGB8 frame out to host memory.  Wall time per call over `--frames` calls after 4 warm-up calls, and the upload bytes."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rayhs_b200 as rh
from rayhs_b200 import capi

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=10)
a = ap.parse_args()
rh.init(0)
L = capi.lib()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sc = rh.Scene.from_pack(os.path.join(root, "tests", "golden", "dragon_full.pack"))
W, H, spp = 3840, 2160, 16
job = rh.renderingFromScene(sc, W, H)
off = torch.empty((W * H, spp, 2), dtype=torch.float64, pin_memory=True)
L.rh_sample_offsets_f64(24, W * H, spp, off.data_ptr())
out = torch.empty((H, W, 3), dtype=torch.uint8, pin_memory=True)
for _ in range(4):
    st = rh.render(job, spp=spp, offsets=off, out=out.numpy()).stats
torch.cuda.synchronize()
t0 = time.time()
for _ in range(a.frames):
    st = rh.render(job, spp=spp, offsets=off, out=out.numpy()).stats
wall = 1e3 * (time.time() - t0) / a.frames
print(json.dumps({"e2e_wall_ms": wall, "device_ms": st["ms_total"], "chunks": st["chunks"], "upload_GB": st["upload_bytes"] / 1e9,
                  "upload_GBps_if_alone": st["upload_bytes"] / 1e6 / wall}))
