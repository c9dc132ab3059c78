#!/usr/bin/env python
"""rh_multi_render on every GPU of the box vs rh_render on one: same bytes; prints frame times.
    gpurun --gpus N -- python scripts/gpu_multi_check.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rayhs_b200 as rh

n = torch.cuda.device_count()
sc = rh.Scene.from_pack(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dragon_full.pack"))
w, h, spp = 3840, 2160, 16
job = rh.renderingFromScene(sc, w, h)
base = rh.render(job, spp=spp, seed=24)
out = {"gpus": n}
for _ in range(4):
    t0 = time.time()
    img = rh.render_multi(job, n, spp=spp, seed=24)
    out["wall_ms"] = 1e3 * (time.time() - t0)
out["device_ms_slowest_shard"] = img.stats["ms_total"]
out["single_gpu_ms"] = base.stats["ms_total"]
out["equal_bytes"] = bool(np.array_equal(img.pixels, base.pixels))
out["rays_equal"] = all(img.stats[k] == base.stats[k] for k in ("rays_primary", "rays_reflect", "rays_shadow"))
off = rh.sample_offsets(w * h, spp, 24)
t0 = time.time()
img2 = rh.render_multi(job, n, spp=spp, offsets=off)
out["wall_ms_host_offsets"] = 1e3 * (time.time() - t0)
out["equal_bytes_host_offsets"] = bool(np.array_equal(img2.pixels, base.pixels))
rh.multi_shutdown([sc])
print(json.dumps(out))
