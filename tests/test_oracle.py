"""CPU tests of the oracle (oracle/oracle.cpp): the hand-derived known answers of SURVEY.md App. D,
analytic primitive cases, the tree shapes of App. B, the colour-map semantics, an independent
numpy restatement of the primary-hit search, and the committed golden outputs.  The reference ships
no golden vectors for this path and GHC is absent, so these are what pins the oracle (parity
"unpinned" in the sense of DESIGN.md)."""
import math
import os

import numpy as np
import pytest

from tests.util import GOLDEN, SCENES, RawScene, load_scene, oracle_for

FOVY = 0.9272952180016123


@pytest.fixture(scope="module")
def cornell():
    sc = load_scene("cornellBox")
    return sc, oracle_for(sc)


def test_focal_term_and_camera_rays(cornell):
    """Projection.hs:35 parses as (tan 0.5) * fovy (App. A-C1); App. D values."""
    sc, o = cornell
    assert math.tan(0.5) * FOVY == pytest.approx(0.5065836864145213, rel=1e-15)
    for h, f in ((512, 505.345921839503), (1080, 1065.9640538802016), (2160, 2131.9281077604032)):
        assert 0.5 * h / (math.tan(0.5) * FOVY) == pytest.approx(f, rel=1e-14)
    org, d = o.ray_from_pixel(sc.camera, 512, 512, 0, 0)
    assert np.allclose(org, (0, 0, -2), atol=0)
    assert np.allclose(d, (-0.4118084708498613, 0.4118084708498613, 0.8129130129802314), rtol=1e-14)
    _, d = o.ray_from_pixel(sc.camera, 512, 512, 256, 256)
    assert tuple(d) == (0.0, 0.0, 1.0)
    _, d = o.ray_from_pixel(sc.camera, 512, 512, 511, 300)
    assert np.allclose(d, (0.4491445255984047, -0.07749944755423453, 0.8900915855987924), rtol=1e-14)


def test_cornell_centre_pixel_known_answer(cornell):
    """App. D: pixel (256,256) hits object 5 (back wall) at t = 4; colour 0.7840383474358017 -> 199."""
    sc, o = cornell
    h = o.closest((0, 0, -2), (0, 0, 1))
    assert h["object"] == 5 and h["tri"] == -1 and h["t"] == 4.0
    assert np.allclose(h["p"], (0, 0, 2)) and np.allclose(h["n"], (0, 0, -1)) and np.allclose(h["uv"], (0, 0))
    r = o.render(sc.camera, 512, 512, 3, rows=(256, 257, 1))
    assert r["rgb_f64"][256, 256] == pytest.approx([0.7840383474358017] * 3, rel=1e-13)
    assert tuple(r["rgb_u8"][256, 256]) == (199, 199, 199)
    assert tuple(r["hit_ids"][256, 256, 0]) == (5, -1)


def test_light_terms_known_answer():
    """App. D light 0 at p = (0,0,2): Light.hs:15-17 and Material.hs:31-33 by hand."""
    L, p = np.array([0, 0.9, 0.75]), np.array([0.0, 0.0, 2.0])
    d = math.sqrt(((L - p) ** 2).sum())
    assert d == pytest.approx(1.5402921800749363, rel=1e-15)
    lc = 200 * (1.0 / (1.0 + d / 0.1) ** 2)
    assert lc == pytest.approx(0.7433401085918134, rel=1e-14)
    ld = (L - p) / d
    lam = max(ld @ np.array([0, 0, -1.0]), 0) * (1 / math.pi) * 2 * lc
    assert lam == pytest.approx(0.38403834743580173, rel=1e-13)


def test_tree_shapes_match_appendix_b():
    """KDTree.hs:79-90 build: inner / leaf counts and depths of the shipped meshes."""
    want = {"cornellBox": {10: (0, 1, 0, 0, 12), 11: (38, 39, 0, 6, 19)}, "dragon_superlow": (279, 280, 0, 11, 19),
            "dragon_low": (547, 548, 0, 12, 19), "dragon_full": (2199, 2200, 0, 16, 19), "transform": (5, 4, 2, 3, None)}  # cone.obj is rotated here: App. B leaf sizes are for the raw mesh
    for name, w in want.items():
        sc = load_scene(name)
        o = oracle_for(sc)
        raw = sc.raw.contents
        meshes = [i for i in range(raw.n_objects) if raw.objects[i].kind == 2]
        if isinstance(w, dict):
            for i, t in w.items():
                s = o.tree_stats(i)
                assert (s["inner"], s["leaves"], s["empties"], s["max_depth"], s["max_leaf"]) == t, (name, i, s)
        else:
            s = o.tree_stats(meshes[0])
            got = (s["inner"], s["leaves"], s["empties"], s["max_depth"], s["max_leaf"])
            assert all(b is None or a == b for a, b in zip(got, w)), (name, s)


def test_flattened_tree_agrees_with_oracle_tree():
    """host_build.cpp (product front end) and the oracle build the same tree from the same triangles."""
    for name in ("cornellBox", "dragon_low", "transform"):
        sc = load_scene(name)
        o = oracle_for(sc)
        d = sc.flat.contents
        for i in range(d.n_objects):
            ob = d.objects[i]
            s = o.tree_stats(i)
            if s is None:
                continue
            assert ob.n_leaves == s["leaves"] and ob.depth == s["max_depth"]
        leaves = [d.nodes[k] for k in range(d.n_nodes) if d.nodes[k].is_leaf]
        assert sum(n.right for n in leaves) == d.n_tris
        firsts = [n.left for n in leaves]
        assert firsts == sorted(firsts)  # triangles are stored in left-to-right leaf order (tie rule)


def _single(objects, materials=None, lights=None):
    return RawScene(objects, materials or [{"kind": "diffuse"}], lights or [])


def test_plane_intersection_semantics():
    """Geometry.hs:70-79: t = n.(p-o)/(d.n); hit iff |d.n| > 0 and t > 0; normal never flipped; uv from tangent."""
    from oracle.orc import OracleScene

    rs = _single([{"kind": "plane", "point": (0, 0, 2), "normal": (0, 0, -1), "tangent": (1, 0, 0)}])
    o = OracleScene(rs.raw)
    h = o.closest((0.25, 0.5, -2), (0, 0, 1))
    assert h["t"] == 4.0 and tuple(h["n"]) == (0, 0, -1) and np.allclose(h["uv"], (0.25, 0.5))  # b = t x n = (0,1,0)
    assert o.closest((0, 0, -2), (0, 0, -1)) is None       # behind
    assert o.closest((0, 0, -2), (1, 0, 0)) is None        # parallel: |d.n| = 0
    h = o.closest((0, 0, 5), (0, 0, -1))                   # from behind the plane: still a hit, same normal
    assert h["t"] == 3.0 and tuple(h["n"]) == (0, 0, -1)


def test_sphere_intersection_semantics():
    """Geometry.hs:81-96: first positive root; outward normal even from inside; uv uses atan (not atan2)."""
    from oracle.orc import OracleScene

    rs = _single([{"kind": "sphere", "center": (0, 0, 0), "radius": 1.0}])
    o = OracleScene(rs.raw)
    h = o.closest((0, 0, -3), (0, 0, 1))
    assert h["t"] == 2.0 and np.allclose(h["n"], (0, 0, -1))
    h = o.closest((0, 0, 0), (0, 0, 1))                    # origin inside: t1, normal outward
    assert h["t"] == 1.0 and np.allclose(h["n"], (0, 0, 1))
    assert o.closest((0, 2, -3), (0, 0, 1)) is None
    h = o.closest((0.5, 0.5, -3), (0, 0, 1))
    n = h["n"]
    assert h["uv"][0] == pytest.approx(math.atan(n[2] / n[0]) / math.pi, rel=1e-15)
    assert h["uv"][1] == pytest.approx(math.acos(n[1]) / math.pi, rel=1e-15)
    # non-unit direction: t scales (a = d.d), Geometry.hs:87
    h = o.closest((0, 0, -3), (0, 0, 2))
    assert h["t"] == 1.0


def test_triangle_intersection_semantics():
    """Mesh.hs:59-82: two-sided, eps on |det| and t, normal = u*n1 + v*n2 + (1-u-v)*n0 un-normalised."""
    from oracle.orc import OracleScene

    pos = [(0, 0, 0), (1, 0, 0), (0, 1, 0)]
    nrm = [(0, 0, -1), (0, 0, -2), (0, 0, -4)]
    uvs = [(0, 0), (1, 0), (0, 1)]
    rs = _single([{"kind": "mesh", "positions": pos, "normals": nrm, "uvs": uvs, "indices": [0, 1, 2]}])
    o = OracleScene(rs.raw)
    h = o.closest((0.25, 0.5, -1), (0, 0, 1))
    assert h["t"] == 1.0 and h["tri"] == 0
    assert np.allclose(h["uv"], (0.25, 0.5)) and np.allclose(h["n"], (0, 0, -(0.25 * 2 + 0.5 * 4 + 0.25 * 1)))
    assert o.closest((0.25, 0.5, 1), (0, 0, -1))["t"] == 1.0          # back face also hits
    assert o.closest((0.75, 0.75, -1), (0, 0, 1)) is None              # u + v > 1
    assert o.closest((0.25, 0.5, -1e-7), (0, 0, 1)) is None            # t < eps
    assert o.closest((0.25, 0.5, -1), (1, 0, 0)) is None               # |det| < eps (parallel)


def test_leaf_and_object_tie_rules():
    """Geometry.hs:54-57 first minimum inside a leaf; RayHs.hs:67-71 first object on equal t."""
    from oracle.orc import OracleScene

    pos = [(0, 0, 0), (1, 0, 0), (0, 1, 0)]
    rs = _single([{"kind": "mesh", "positions": pos, "indices": [0, 1, 2, 0, 1, 2]}])
    assert OracleScene(rs.raw).closest((0.25, 0.25, -1), (0, 0, 1))["tri"] == 0
    rs = _single([{"kind": "plane", "point": (0, 0, 1), "normal": (0, 0, -1), "tangent": (1, 0, 0)},
                  {"kind": "plane", "point": (0, 0, 1), "normal": (0, 0, -1), "tangent": (1, 0, 0)}])
    assert OracleScene(rs.raw).closest((0, 0, 0), (0, 0, 1))["object"] == 0


def test_mod1_and_checker_and_texture():
    """ColorMap.hs:18-58 with Data.Fixed.mod' (exact rational floor), round-half-even, floor-mod wrap."""
    from oracle import orc
    from oracle.orc import OracleScene

    assert orc.mod1(0.75, 0.5) == 0.25 and orc.mod1(-0.25, 1.0) == 0.75 and orc.mod1(2.0, 1.0) == 0.0
    assert orc.mod1(-1e-20, 1.0) == 1.0  # App. A-X2: can return exactly the divisor
    tex = np.zeros((2, 4, 3))
    tex[0, :, 0] = (0.0, 0.25, 0.5, 1.0)
    tex[1, :, 0] = (1.0, 0.5, 0.25, 0.0)
    mats = [{"kind": "diffuse", "cmap": "checker", "color1": (1, 0, 0), "color2": (0, 1, 0), "size": 0.5},
            {"kind": "diffuse", "cmap": "texture", "texture": 0}]
    rs = RawScene([{"kind": "sphere", "center": (0, 0, 0), "radius": 1, "material": 0}], mats, [], textures=[tex])
    o = OracleScene(rs.raw)
    assert tuple(o.color_at(0, 0.1, 0.3)) == (1, 0, 0)   # (0.1-0.25)*(0.3-0.25) < 0 -> color1
    assert tuple(o.color_at(0, 0.1, 0.1)) == (0, 1, 0)
    # texture: u=0.5 -> 2.0 -> ui=2; x0=1,x1=2; lx = 2 - 1 - 0.5 = 0.5 ; v=0.25 -> 0.5 -> round-half-even -> 0; y0=-1 mod 2 = 1, y1 = 0; ly = 0.5-(-1)-0.5 = 1
    c = o.color_at(1, 0.5, 0.25)
    assert c[0] == pytest.approx(1.0 * (0.5 * 0.5 + 0.5 * 0.25) + 0.0, rel=1e-15)


def test_numpy_restatement_of_primary_hits_agrees(cornell):
    """Second, independent restatement (numpy, brute force over all primitives, no tree): same primary hit
    object for every pixel of a 64x64 cornellBox frame."""
    sc, o = cornell
    w = h = 64
    ref = o.render(sc.camera, w, h, 0)["hit_ids"].reshape(h, w, 2)
    raw = sc.raw.contents
    f = 0.5 * h / (math.tan(0.5) * FOVY)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    v = np.stack([w * (xs - w / 2) / w, h * ((-ys) + h / 2) / h, np.full_like(xs, f)], -1)
    d = v / np.sqrt((v * v).sum(-1, keepdims=True))          # camera matrix is the identity for this scene
    org = np.array([0, 0, -2.0])
    best_t = np.full((h, w), np.inf)
    best_o = np.full((h, w), -1)
    for i in range(raw.n_objects):
        ob = raw.objects[i]
        a, b = np.array(ob.a[:]), np.array(ob.b[:])
        if ob.kind == 0:
            dn = d @ b
            with np.errstate(divide="ignore", invalid="ignore"):
                t = (b @ (a - org)) / dn
            hit = (np.abs(dn) > 0) & (t > 0)
        elif ob.kind == 1:
            oc = org - a
            A = (d * d).sum(-1)
            B = 2.0 * (d @ oc)
            Cc = oc @ oc - b[0] * b[0]
            delta = B * B - 4.0 * A * Cc
            with np.errstate(invalid="ignore"):
                sq = np.sqrt(np.where(delta >= 0, delta, 0))
            t0, t1 = 0.5 * (-B - sq) / A, 0.5 * (-B + sq) / A
            t = np.where(t0 > 0, t0, t1)
            hit = (delta >= 0) & (t > 0)
        else:
            P = np.ctypeslib.as_array(ob.positions, shape=(ob.n_verts, 3))
            I = np.ctypeslib.as_array(ob.indices, shape=(ob.n_indices,)).reshape(-1, 3)
            t = np.full((h, w), np.inf)
            for tri in I:
                p0, e1, e2 = P[tri[0]], P[tri[1]] - P[tri[0]], P[tri[2]] - P[tri[0]]
                pv = np.cross(d, e2)
                det = pv @ e1
                with np.errstate(divide="ignore", invalid="ignore"):
                    idet = 1 / det
                    t0v = org - p0
                    u = idet * (pv @ t0v)
                    q = np.cross(t0v, e1)
                    vv = idet * (d @ q)
                    tt = idet * (q @ e2)
                ok = ~((np.abs(det) < 1e-6) | (u < 0) | (u > 1) | (vv < 0) | (u + vv > 1) | (tt < 1e-6))
                t = np.where(ok & (tt < t), tt, t)
            hit = np.isfinite(t)
        upd = hit & (t < best_t)
        best_t = np.where(upd, t, best_t)
        best_o = np.where(upd, i, best_o)
    assert np.array_equal(best_o, ref[..., 0])


@pytest.mark.parametrize("name", SCENES)
def test_golden_outputs(name):
    """The committed oracle outputs (tests/golden/golden_*.npz, made by make_packs.py) still reproduce."""
    g = np.load(os.path.join(GOLDEN, f"golden_{name}.npz"))
    w, h, depth = (int(x) for x in g["size"])
    sc = load_scene(name)
    r = oracle_for(sc).render(sc.camera, w, h, depth)
    assert np.array_equal(r["rgb_u8"], g["rgb_u8"]) and np.array_equal(r["rgb_int"], g["rgb_int"])
    assert np.array_equal(r["hit_ids"], g["hit_ids"])
    assert [r["rays"][k] for k in ("primary", "reflect", "probe", "exit", "shadow")] == list(g["rays"])


def test_oracle_row_subsets_and_threads_agree(cornell):
    sc, o = cornell
    full = o.render(sc.camera, 48, 48, 3, threads=1)
    part = o.render(sc.camera, 48, 48, 3, rows=(1, 48, 2), threads=4)
    assert np.array_equal(full["rgb_u8"][1::2], part["rgb_u8"][1::2])
    assert not part["rgb_u8"][0::2].any()


def test_pack_reader_of_the_reference_arm_equals_the_front_end():
    """oracle/packio.py (what `bench.py --impl reference` uses, so that the arm never maps the product library) reads a
    scene pack into the same raw scene as the front end's rh_load_pack, and generates the same offset stream."""
    import rayhs_b200 as rh
    from oracle import packio
    from oracle.orc import OracleScene

    assert np.array_equal(packio.sample_offsets(777, 3, 24), rh.sample_offsets(777, 3, 24))
    for name in ("cornellBox", "texture", "dragon_low"):
        path = os.path.join(GOLDEN, name + ".pack")
        a, b = packio.PackScene(path), rh.Scene.from_pack(path)
        assert (a.width, a.height, a.max_depth) == (b.width, b.height, b.max_depth)
        oa, ob = OracleScene(a.raw), OracleScene(b.raw)
        ra = oa.render(a.camera, 40, 30, a.max_depth, want_ids=True)
        rb = ob.render(b.camera, 40, 30, b.max_depth, want_ids=True)
        assert np.array_equal(ra["rgb_u8"], rb["rgb_u8"]) and np.array_equal(ra["hit_ids"], rb["hit_ids"])
        assert ra["rays_total"] == rb["rays_total"]
        oa.close()
        ob.close()
