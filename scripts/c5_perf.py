#!/usr/bin/env python
"""configs[4] shape on one GPU: n_tris random triangles + n_spheres spheres, WxH, spp with a tiled offset stream
(the per-pixel stream of 8K x 64 spp is 34 GB).  Prints one JSON line: frame time per kernel, rays, counters."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rayhs_b200 as rh
from rayhs_b200 import capi

ap = argparse.ArgumentParser()
ap.add_argument("--tris", type=int, default=10_000_000)
ap.add_argument("--spheres", type=int, default=1000)
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--spp", type=int, default=4)
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--count", action="store_true")
ap.add_argument("--shadow", default="auto", choices=["auto", "pooled", "split"])
a = ap.parse_args()
rh.init(0)
t0 = time.time()
sc = rh.Scene.synthetic(a.tris, a.spheres)
job = rh.renderingFromScene(sc, a.width, a.height)
_ = sc.device
t_build = time.time() - t0
tile = torch.from_numpy(rh.sample_offsets(64 * 64, a.spp, 24)).cuda()
rgb = torch.empty((a.height, a.width, 3), dtype=torch.uint8, device="cuda")
for i in range(a.frames):
    st = rh.render_device(job, rgb, spp=a.spp, offsets_dev=tile, offset_tile=64, profile=True, count=(a.count and i == a.frames - 1),
                          shadow=None if a.shadow == "auto" else a.shadow)
rays = st["rays_primary"] + st["rays_reflect"] + st["rays_probe"] + st["rays_exit"] + st["rays_shadow"]
st.update(build_s=t_build, mrays_per_s=rays / st["ms_total"] / 1e3, tris=a.tris, spheres=a.spheres, w=a.width, h=a.height, spp=a.spp,
          mem_GB=torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9)
print(json.dumps(st))
