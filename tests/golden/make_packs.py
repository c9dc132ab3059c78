"""Generates the scene packs and oracle golden images under tests/golden/ (run in the dev container).

The GPU boxes have no /root/reference, so every BASELINE.json scene is loaded ONCE here through the
front end (rh_load_json: JSON.hs / Mesh.hs OBJ / Bitmap.hs PPM semantics) and saved as a binary
"pack" (rh_save_pack: the rh_raw_scene before any tree build).  Two textures that
data/texture.json names but the reference does not ship (data/checkerboard.ppm, data/tile.ppm;
SURVEY.md §7 hard part 7) are synthesised here as P3 files.

Also writes golden_<scene>.npz: the ORACLE's output (rgb_u8, primary hit ids, ray counts) at a
small resolution, used by the CPU tests to pin the oracle against regressions.  These are
outputs of oracle/oracle.cpp, not of the reference (GHC is absent): parity stays "unpinned".

    python tests/golden/make_packs.py
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def write_p3(path, img):
    h, w, _ = img.shape
    with open(path, "w") as f:
        f.write(f"P3\n{w} {h}\n255\n")
        f.write("\n".join(str(int(v)) for v in img.reshape(-1)))
        f.write("\n")


def synth_textures(data_dir):
    # 64x64, 8-pixel black/white checker
    y, x = np.mgrid[0:64, 0:64]
    c = (((x // 8) + (y // 8)) % 2) * 255
    write_p3(os.path.join(data_dir, "checkerboard.ppm"), np.stack([c, c, c], -1))
    # 64x64 two-tone tile with a darker 4-pixel grout, seeded noise
    rng = np.random.RandomState(7)
    t = np.full((64, 64, 3), (200, 120, 80), dtype=np.int64) + rng.randint(-20, 21, size=(64, 64, 1))
    t[(x % 32 < 2) | (y % 32 < 2)] = (60, 60, 60)
    write_p3(os.path.join(data_dir, "tile.ppm"), np.clip(t, 0, 255))


def main():
    from rayhs_b200 import Scene
    from oracle.orc import OracleScene

    work = tempfile.mkdtemp(prefix="rh_packs_")
    data = os.path.join(work, "data")
    shutil.copytree(os.path.join(REF, "data"), data)
    os.chmod(data, 0o755)
    synth_textures(data)
    # BASELINE configs C3/C4: dragon.json with the low / full-res mesh (the shipped file names dragon_superlow.obj)
    dj = json.load(open(os.path.join(data, "dragon.json")))
    for name, obj in (("dragon_low", "data/dragon_low.obj"), ("dragon_full", "data/dragon.obj")):
        j = json.loads(json.dumps(dj))
        n = 0
        for ob in j["scene"]["objects"]:
            if ob["geometry"]["type"] == "mesh":
                ob["geometry"]["fileName"] = obj
                n += 1
        assert n == 1
        json.dump(j, open(os.path.join(data, name + ".json"), "w"))
    scenes = {
        "cornellBox": "cornellBox.json", "texture": "texture.json", "transform": "transform.json",
        "dragon_superlow": "dragon.json", "dragon_low": "dragon_low.json", "dragon_full": "dragon_full.json",
        "outScene": "outScene.json",
    }
    golden_res = {"cornellBox": (96, 96), "texture": (128, 72), "transform": (128, 72), "dragon_superlow": (96, 96),
                  "dragon_low": (128, 72), "dragon_full": (128, 72), "outScene": (128, 72)}
    cwd = os.getcwd()
    os.chdir(work)  # file names inside the JSON are cwd-relative (JSON.hs:113, Descriptors.hs:52)
    try:
        for name, fn in scenes.items():
            sc = Scene.from_json(os.path.join("data", fn))
            pack = os.path.join(OUT, name + ".pack")
            sc.save_pack(pack)
            w, h = golden_res[name]
            o = OracleScene(sc.raw)
            r = o.render(sc.camera, w, h, sc.max_depth)
            np.savez_compressed(os.path.join(OUT, f"golden_{name}.npz"), rgb_u8=r["rgb_u8"], rgb_int=r["rgb_int"],
                                hit_ids=r["hit_ids"], rays=np.array([r["rays"][k] for k in ("primary", "reflect", "probe", "exit", "shadow")]),
                                size=np.array([w, h, sc.max_depth]))
            print(name, sc.width, sc.height, os.path.getsize(pack), r["rays"])
            o.close()
            sc.close()
    finally:
        os.chdir(cwd)
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
