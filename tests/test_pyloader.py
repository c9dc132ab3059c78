"""The C++ front end (csrc/frontend.cpp) against an independent Python loader written from the Haskell sources
(tests/pyloader.py): every scene pack the parity tests consume must hold exactly what the independent loader parses out
of the reference's own data/ files — vertices after de-duplication and transform, indices, planes, spheres, materials,
texels, lights, camera.  Runs where /root/reference exists (this container); the GPU boxes only see the packs."""
import json
import os
import shutil
import tempfile

import numpy as np
import pytest

from tests import pyloader
from tests.util import load_scene

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "data")), reason="the reference's data/ is not on this box")

KINDS = {0: "mirror", 1: "diffuse", 2: "plastic", 3: "emmit", 4: "transparent"}
CMAPS = {0: "flat", 1: "checker", 2: "texture"}


@pytest.fixture(scope="module")
def data_dir():
    """The reference's data/ plus the two textures texture.json names but the reference does not ship
    (tests/golden/make_packs.py synthesises the same two files for the packs)."""
    from tests.golden.make_packs import synth_textures

    work = tempfile.mkdtemp(prefix="rh_pyloader_")
    data = os.path.join(work, "data")
    shutil.copytree(os.path.join(REF, "data"), data)
    os.chmod(data, 0o755)
    synth_textures(data)
    dj = json.load(open(os.path.join(data, "dragon.json")))
    for name, obj in (("dragon_superlow", None), ("dragon_low", "data/dragon_low.obj"), ("dragon_full", "data/dragon.obj")):
        d2 = json.loads(json.dumps(dj))
        if obj:
            for o in d2["scene"]["objects"]:
                if o["geometry"]["type"] == "mesh":
                    o["geometry"]["fileName"] = obj
        json.dump(d2, open(os.path.join(data, name + ".json"), "w"))
    yield work
    shutil.rmtree(work, ignore_errors=True)


@pytest.mark.parametrize("name", ["cornellBox", "texture", "transform", "dragon_superlow", "dragon_low", "dragon_full", "outScene"])
def test_pack_holds_what_an_independent_loader_parses(name, data_dir):
    sc = load_scene(name)
    want = pyloader.load_scene(os.path.join(data_dir, "data", name + ".json"), data_dir)
    raw = sc.raw.contents
    assert (sc.width, sc.height, sc.max_depth) == (want["width"], want["height"], want["max_depth"])
    cam = sc.camera
    assert np.array_equal(np.array(cam.position[:]), want["camera"]["position"])
    assert np.array_equal(np.array(cam.target[:]), want["camera"]["target"])
    assert np.array_equal(np.array(cam.up[:]), want["camera"]["up"])
    assert (cam.projection == 1) == (want["camera"]["projection"] == "perspective")
    if cam.projection == 1:
        assert cam.fovy == want["camera"]["fovy"]
    assert raw.n_objects == len(want["objects"]) and raw.n_lights == len(want["lights"])
    for i, o in enumerate(want["objects"]):
        ro = raw.objects[i]
        assert ro.kind == {"plane": 0, "sphere": 1, "mesh": 2}[o["kind"]], (name, i)
        if o["kind"] == "plane":
            assert np.array_equal(ro.a[:], o["point"]) and np.array_equal(ro.b[:], o["normal"]) and np.array_equal(ro.c[:], o["tangent"])
        elif o["kind"] == "sphere":
            assert np.array_equal(ro.a[:], o["center"]) and ro.b[0] == o["radius"]
        else:
            nv, ni = ro.n_verts, ro.n_indices
            assert (nv, ni) == (len(o["positions"]), len(o["indices"])), (name, i, nv, ni)
            assert np.array_equal(np.ctypeslib.as_array(ro.indices, (ni,)), o["indices"])
            # bit-exact: the same IEEE operations in the same order (Transform.hs / Mat.hs), decimal parsing is correctly rounded on both sides
            assert np.array_equal(np.ctypeslib.as_array(ro.positions, (nv, 3)), o["positions"]), (name, i, "positions")
            assert np.array_equal(np.ctypeslib.as_array(ro.normals, (nv, 3)), o["normals"]), (name, i, "normals")
            assert np.array_equal(np.ctypeslib.as_array(ro.uvs, (nv, 2)), o["uvs"]), (name, i, "uvs")
        m, wm = raw.materials[ro.material], want["materials"][o["material"]]
        assert KINDS[m.kind] == wm["kind"], (name, i)
        if "ior" in wm:
            assert m.ior == wm["ior"]
        if wm["kind"] == "emmit":
            assert np.array_equal(m.color1[:], wm["color1"])
        if wm["kind"] in ("diffuse", "plastic"):
            assert CMAPS[m.cmap_kind] == wm["cmap"]
            if wm["cmap"] == "flat":
                assert np.array_equal(m.color1[:], wm["color1"])
            elif wm["cmap"] == "checker":
                assert np.array_equal(m.color1[:], wm["color1"]) and np.array_equal(m.color2[:], wm["color2"]) and m.size == wm["size"]
            else:
                t, tex = raw.textures[m.texture], want["textures"][wm["texture"]]
                assert (t.h, t.w) == tex.shape[:2]
                got = np.ctypeslib.as_array(raw.texels, (raw.n_texels, 3))[t.offset:t.offset + t.w * t.h]
                assert np.array_equal(got, tex.reshape(-1, 3))
    for i, l in enumerate(want["lights"]):
        rl = raw.lights[i]
        assert (rl.kind == 1) == (l["kind"] == "point")
        assert np.array_equal(rl.vec[:], l["vec"]) and np.array_equal(rl.color[:], l["color"])
        if l["kind"] == "point":
            assert rl.radius == l["radius"]
