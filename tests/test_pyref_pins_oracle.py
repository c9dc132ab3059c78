"""Pins the oracle with a second, independent restatement of the WHOLE path (tests/pyref.py: scalar Python written from
the Haskell sources, brute force over every triangle, no tree): same colours (to rounding), same bytes and the same
ray counts per class on the shipped scenes and on a hand-built scene with every material kind.  The reference has no
golden vectors and GHC is not installed, so agreement of two restatements that share no code is the
strongest pin available here."""
import numpy as np
import pytest

from oracle.orc import OracleScene
from tests import pyref
from tests.util import RawScene, load_scene, oracle_for

CASES = [("cornellBox", 26, 26), ("texture", 30, 17), ("transform", 24, 14), ("outScene", 24, 14), ("dragon_superlow", 14, 14)]


def check(raw, camera, w, h, depth, orc):
    ref = orc.render(camera, w, h, depth)
    img, u8, count = pyref.render(raw, camera, w, h, depth)
    assert count == ref["rays"], (count, ref["rays"])
    # the two sum and multiply in slightly different orders here and there: agreement to 1e-9, and in every byte
    assert np.allclose(img, ref["rgb_f64"], rtol=1e-9, atol=1e-12, equal_nan=True), np.abs(img - ref["rgb_f64"]).max()
    assert np.array_equal(u8, ref["rgb_u8"])
    return ref


@pytest.mark.parametrize("name,w,h", CASES)
def test_shipped_scenes(name, w, h):
    sc = load_scene(name)
    raw = sc.raw.contents if hasattr(sc.raw, "contents") else sc.raw
    ref = check(raw, sc.camera, w, h, sc.max_depth, oracle_for(sc))
    assert ref["rays"]["shadow"] > 0


def test_every_material_kind_and_both_projections():
    tex = np.random.RandomState(3).uniform(0, 1, size=(5, 7, 3))
    mats = [{"kind": "diffuse", "color1": (0.9, 0.2, 0.2)}, {"kind": "plastic", "ior": 1.9, "color1": (0.2, 0.9, 0.2)},
            {"kind": "mirror", "ior": 4.0}, {"kind": "emmit", "color1": (3, 2, 1)}, {"kind": "transparent", "ior": 1.5},
            {"kind": "shownormal"}, {"kind": "showuv"},
            {"kind": "plastic", "ior": 1.3, "cmap": "checker", "color1": (1, 1, 1), "color2": (0.1, 0.1, 0.1), "size": 0.4},
            {"kind": "diffuse", "cmap": "texture", "texture": 0}]
    objs = [{"kind": "sphere", "center": (-1.5 + 0.75 * i, 0.2 * (i % 2), 0.5), "radius": 0.33, "material": i} for i in range(7)]
    quad = dict(positions=[(-1, -1, 1.6), (1, -1, 1.6), (1, 1, 1.8), (-1, 1, 1.8)], normals=[(0, 0.1, -1), (0.1, 0, -1), (0, -0.1, -1), (-0.1, 0, -1)],
                uvs=[(0, 0), (1, 0), (1, 1), (0, 1)], indices=[0, 1, 2, 0, 2, 3])
    objs += [{"kind": "sphere", "center": (0.0, 0.9, 0.0), "radius": 0.4, "material": 4},
             {"kind": "sphere", "center": (0.9, 0.8, 0.3), "radius": 0.3, "material": 8},
             {"kind": "plane", "point": (0, -0.6, 0), "normal": (0, 1, 0), "tangent": (1, 0, 0), "material": 7},
             {"kind": "plane", "point": (0, 0, 3), "normal": (0, 0, -2), "tangent": (1, 0, 0), "material": 8},
             dict(kind="mesh", material=8, **quad)]
    lights = [{"kind": "point", "vec": (0, 2, -1), "color": (20, 20, 20), "radius": 0.5},
              {"kind": "directional", "vec": (0.3, 1, -0.5), "color": (0.6, 0.6, 0.7)}]
    rs = RawScene(objs, mats, lights, textures=[tex], camera={"position": (0.1, 0.3, -3), "target": (0, 0.1, 0.5), "up": (0.1, 1, 0)})
    for depth in (0, 2, 4):
        ref = check(rs.raw_struct, rs.camera, 28, 20, depth, OracleScene(rs.raw))
    assert ref["rays"]["probe"] > 0 and ref["rays"]["exit"] > 0 and ref["rays"]["reflect"] > 0
    rs2 = RawScene(objs, mats, lights, textures=[tex], camera={"position": (0.1, 0.3, -3), "target": (0, 0.1, 0.5), "projection": "orthographic"})
    # Projection.hs:30-32: the orthographic view plane is in pixel units — a 6 x 4 pixel frame spans the scene
    check(rs2.raw_struct, rs2.camera, 6, 4, 2, OracleScene(rs2.raw))


def test_multi_sample_pixels():
    """distributedRayTrace (RayHs.hs:169-195): spp samples at (i + x - 0.5, j + y - 0.5), unweighted mean before toIntC."""
    import rayhs_b200 as rh

    sc = load_scene("cornellBox")
    raw = sc.raw.contents if hasattr(sc.raw, "contents") else sc.raw
    w, h, spp = 12, 12, 5
    off = rh.sample_offsets(w * h, spp, seed=24)
    ref = oracle_for(sc).render(sc.camera, w, h, sc.max_depth, spp=spp, offsets=off)
    img, u8, count = pyref.render(raw, sc.camera, w, h, sc.max_depth, spp=spp, offsets=np.asarray(off).reshape(w * h, spp, 2))
    assert count == ref["rays"]
    assert np.allclose(img, ref["rgb_f64"], rtol=1e-9, atol=1e-12)
    assert np.array_equal(u8, ref["rgb_u8"])
