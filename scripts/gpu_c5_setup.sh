#!/bin/bash
# Set-up stages of rh_scene_create for configs[4]'s geometry (10 M triangles + 1 000 spheres), and the 10 M-triangle GPU test.
RAYHS_B200_DEBUG=s python - <<'PY' 2>&1 | grep -v "^$" | tail -14
import time, sys
sys.path.insert(0, ".")
import rayhs_b200 as rh
rh.init(0)
t0 = time.time(); sc = rh.Scene.synthetic(10_000_000, 1000); _ = sc.flat; t1 = time.time(); _ = sc.device; t2 = time.time()
print("generate + flatten %.2f s, rh_scene_create %.2f s" % (t1 - t0, t2 - t1), rh.scene_setup_ms(sc))
PY
timeout 600 python -m pytest tests/test_full_size_gpu.py -x -q 2>&1 | tail -2
