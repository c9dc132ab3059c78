"""The per-light cube maps of nearest possible occluder distance (csrc/light_maps.cpp) are conservative: whenever the
map clears a (shaded point, light) pair, a brute-force shadowIntersection over the mesh's triangles (RayHs.hs:74-87
with Mesh.hs:59-82's accept rule) finds no occluder in front of the light.  CPU only: the map comes from the C ABI's
validation hook, the cell lookup below is the float arithmetic of the kernels' light_map_cell restated in numpy."""
import ctypes as C

import numpy as np
import pytest

from rayhs_b200 import capi

from .util import load_scene

EPS = 1e-6


def scene_tris(name):
    sc = load_scene(name)
    d = sc.flat.contents
    n = d.n_tris
    raw = np.ctypeslib.as_array(C.cast(d.tris, C.POINTER(C.c_double)), shape=(n, 10)).copy()
    raw_scene = sc.raw.contents if hasattr(sc.raw, "contents") else sc.raw
    lights = [np.array(raw_scene.lights[i].vec[:]) for i in range(raw_scene.n_lights) if raw_scene.lights[i].kind == capi.RH_LIGHT_POINT]
    return sc, d.tris, n, raw[:, 0:3], raw[:, 3:6], raw[:, 6:9], lights


def build_map(L, tris_ptr, n, res):
    out = np.empty(6 * res * res, dtype=np.float32)
    useful, empty = C.c_int(0), C.c_double(0)
    capi.check(capi.lib().rh_light_map_build((C.c_double * 3)(*L), tris_ptr, n, res, out.ctypes.data, C.byref(useful), C.byref(empty)))
    return out, bool(useful.value), empty.value


def map_cells(p, L, res):
    """light_map_cell of kernels.cu, in float32, for an array of points."""
    v = (p - L).astype(np.float32)
    a = np.abs(v)
    k = np.where((a[:, 0] >= a[:, 1]) & (a[:, 0] >= a[:, 2]), 0, np.where(a[:, 1] >= a[:, 2], 1, 2))
    rows = np.arange(len(p))
    w = a[rows, k]
    face = 2 * k + (v[rows, k] < 0)
    ia = np.where(k == 0, 1, 0)
    ib = np.where(k == 2, 1, 2)
    ok = (w > np.float32(1e-30)) & (w < np.float32(1e30))
    iw = np.float32(1) / np.where(ok, w, np.float32(1))
    half = np.float32(0.5 * res)
    ci = np.clip(np.floor((v[rows, ia] * iw + np.float32(1)) * half).astype(np.int64), 0, res - 1)
    cj = np.clip(np.floor((v[rows, ib] * iw + np.float32(1)) * half).astype(np.int64), 0, res - 1)
    return np.where(ok, (face * res + cj) * res + ci, -1)


def round_up_f32(x):
    f = x.astype(np.float32)
    return np.where(f.astype(np.float64) < x, np.nextafter(f, np.float32(np.inf)), f)


def occluded(p, L, p0, e1, e2, chunk=256, directional=False):
    """shadowIntersection against the mesh alone, brute force: ray (p + eps*ld, ld), ld = (L - p)/dist for a point light at
    L (Light.hs:15-17), ld = L itself for a directional light (Light.hs:14), whose hits are always in front (RayHs.hs:85)."""
    out = np.zeros(len(p), dtype=bool)
    for s in range(0, len(p), chunk):
        pp = p[s:s + chunk]
        if directional:
            ld = np.broadcast_to(L, pp.shape).copy()
        else:
            dv = L - pp
            dd = np.sqrt((dv * dv).sum(1))
            ld = dv * (1 / dd)[:, None]
        o = pp + EPS * ld
        D, O = ld[:, None, :], o[:, None, :]
        pv = np.cross(D, e2[None])
        det = (e1[None] * pv).sum(-1)
        with np.errstate(divide="ignore", invalid="ignore"):
            idet = 1 / det
            t0 = O - p0[None]
            u = idet * (t0 * pv).sum(-1)
            q = np.cross(t0, e1[None])
            vv = idet * (D * q).sum(-1)
            tt = idet * (e2[None] * q).sum(-1)
        hit = ~((np.abs(det) < EPS) | (u < 0) | (u > 1) | (vv < 0) | (u + vv > 1) | (tt < EPS) | np.isnan(u) | np.isnan(vv) | np.isnan(tt))
        dl2 = ((o - L) ** 2).sum(1)  # inFrontOfLight, RayHs.hs:84-87
        front = dl2[:, None] > (tt * tt) * (ld * ld).sum(1)[:, None]
        if directional:
            front = np.ones_like(hit)
        out[s:s + chunk] = (hit & front).any(1)
    return out


def query_points(rng, p0, e1, e2, L, n):
    lo = np.minimum(np.minimum(p0, p0 + e1), p0 + e2).min(0)
    hi = np.maximum(np.maximum(p0, p0 + e1), p0 + e2).max(0)
    ext = (hi - lo).max()
    pts = [rng.uniform(lo - 2 * ext, hi + 2 * ext, size=(n, 3))]  # the volume around the mesh
    walls = rng.uniform(lo - 1.5 * ext, hi + 1.5 * ext, size=(n, 3))  # the faces of a room around it
    ax = rng.integers(0, 3, n)
    side = rng.integers(0, 2, n)
    walls[np.arange(n), ax] = np.where(side == 0, (lo - 1.5 * ext)[ax], (hi + 1.5 * ext)[ax])
    pts.append(walls)
    k = rng.integers(0, len(p0), n)  # points on the mesh itself, lifted a little off the surface either way
    a, b = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    flip = a + b > 1
    a, b = np.where(flip, 1 - a, a), np.where(flip, 1 - b, b)
    nrm = np.cross(e1[k], e2[k])
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1), 1e-300)[:, None]
    pts.append(p0[k] + a[:, None] * e1[k] + b[:, None] * e2[k] + nrm * rng.choice([-1e-3, 1e-9, 1e-3], n)[:, None])
    pts.append(L + (pts[0] - L) * 1e-3)  # close to the light
    return np.concatenate(pts)


@pytest.mark.parametrize("name,res", [("dragon_superlow", 256), ("dragon_superlow", 512), ("cornellBox", 512), ("transform", 64)])
def test_cleared_pairs_have_no_occluder(name, res):
    sc, tris_ptr, n, p0, e1, e2, lights = scene_tris(name)
    rng = np.random.default_rng(7)
    if not lights:  # transform.json has directional lights only: take points as lights
        lights = [np.array([0.3, 2.0, -1.0]), np.array([0.0, 0.2, 0.0])]
    checked = cleared_total = 0
    for L in lights:
        m, useful, empty = build_map(L, tris_ptr, n, res)
        if not useful:
            continue
        p = query_points(rng, p0, e1, e2, L, 1500)
        cell = map_cells(p, L, res)
        dd = np.sqrt(((L - p) ** 2).sum(1))
        cleared = (cell >= 0) & (round_up_f32(dd) < m[np.maximum(cell, 0)])
        occ = occluded(p, L, p0, e1, e2)
        assert not (cleared & occ).any(), (name, res, L, p[cleared & occ][:3])
        checked += len(p)
        cleared_total += int(cleared.sum())
        # the map is worth its lookups: most pairs without an occluder are cleared
        assert cleared[~occ].mean() > 0.5, (name, L, cleared[~occ].mean())
    assert checked and cleared_total


def test_rays_through_face_edges_and_corners():
    """Directions on the seams of the cube map (two or three equal components) land in marked cells."""
    sc, tris_ptr, n, p0, e1, e2, lights = scene_tris("cornellBox")
    L = p0.mean(0)  # in the open, between the cube and the torus
    res = 128
    m, useful, _ = build_map(L, tris_ptr, n, res)
    assert useful
    dirs = np.array([[sx, sy, sz] for sx in (-1, 0, 1) for sy in (-1, 0, 1) for sz in (-1, 0, 1) if (sx, sy, sz) != (0, 0, 0)], dtype=np.float64)
    rng = np.random.default_rng(3)
    p = np.concatenate([L + d[None, :] * rng.uniform(0.05, 6.0, (40, 1)) for d in dirs])
    p = np.concatenate([p, p + rng.normal(0, 1e-9, p.shape)])
    cell = map_cells(p, L, res)
    dd = np.sqrt(((L - p) ** 2).sum(1))
    cleared = (cell >= 0) & (round_up_f32(dd) < m[np.maximum(cell, 0)])
    occ = occluded(p, L, p0, e1, e2)
    assert not (cleared & occ).any()
    assert occ.any() and cleared.any()


def test_light_touching_a_triangle_turns_the_map_off():
    sc, tris_ptr, n, p0, e1, e2, _ = scene_tris("cornellBox")
    L = p0[0] + 0.25 * e1[0] + 0.25 * e2[0]
    _, useful, _ = build_map(L, tris_ptr, n, 64)
    assert not useful


def test_bad_arguments():
    out = np.empty(6, dtype=np.float32)
    L = (C.c_double * 3)(0, 0, 0)
    assert capi.lib().rh_light_map_build(L, None, 1, 1, out.ctypes.data, None, None) == capi.RH_ERR_ARG
    assert capi.lib().rh_light_map_build(L, None, 0, 0, out.ctypes.data, None, None) == capi.RH_ERR_ARG
    capi.check(capi.lib().rh_light_map_build(L, None, 0, 1, out.ctypes.data, None, None))
    assert np.isinf(out).all()


@pytest.mark.parametrize("name", ["dragon_superlow", "cornellBox"])
def test_points_of_lit_triangles_see_the_light(name):
    """Lit-triangle flags: from any point of a flagged triangle (corners and edges included, and a little outside, as a
    rounded hit point can be) the brute-force shadow query against the whole mesh finds no occluder — for the scene's
    point lights and for directional lights (un-normalised vectors, as the reference uses them)."""
    sc, tris_ptr, n, p0, e1, e2, lights = scene_tris(name)
    rng = np.random.default_rng(11)
    cases = [(capi.RH_LIGHT_POINT, L) for L in lights]
    cases += [(capi.RH_LIGHT_DIRECTIONAL, np.array(v)) for v in ((0.0, 1.0, 0.0), (0.3, 0.6, -1.2), (-2.0, 0.5, 0.4), (0.02, -0.03, 0.01))]
    total_lit = 0
    for kind, L in cases:
        directional = kind == capi.RH_LIGHT_DIRECTIONAL
        flags = np.zeros(n, dtype=np.uint8)
        capi.check(capi.lib().rh_lit_triangles(kind, (C.c_double * 3)(*L), tris_ptr, n, flags.ctypes.data))
        lit = np.flatnonzero(flags)
        total_lit += len(lit)
        if not len(lit):
            continue
        k = rng.choice(lit, 3000)
        a, b = rng.uniform(0, 1, len(k)), rng.uniform(0, 1, len(k))
        flip = a + b > 1
        a, b = np.where(flip, 1 - a, a), np.where(flip, 1 - b, b)
        snap = rng.integers(0, 4, len(k))  # a quarter each: interior, on edge e1, on edge e2, at the corner p0
        a = np.where((snap == 2) | (snap == 3), 0.0, a) - np.where(snap == 3, 1e-12, 0.0)
        b = np.where((snap == 1) | (snap == 3), 0.0, b) - np.where(snap == 1, 1e-12, 0.0)
        p = p0[k] + a[:, None] * e1[k] + b[:, None] * e2[k]
        occ = occluded(p, L, p0, e1, e2, directional=directional)
        assert not occ.any(), (name, kind, L, int(occ.sum()), k[occ][:5])
    assert total_lit > 0.1 * n, (total_lit, n)
    print(name, "lit (triangle, light) pairs:", total_lit, "of", n * len(cases))


def test_lit_triangles_bad_arguments():
    L = (C.c_double * 3)(0, 1, 0)
    assert capi.lib().rh_lit_triangles(7, L, None, 0, None) == capi.RH_ERR_ARG
    assert capi.lib().rh_lit_triangles(capi.RH_LIGHT_POINT, L, None, 1, None) == capi.RH_ERR_ARG
    capi.check(capi.lib().rh_lit_triangles(capi.RH_LIGHT_POINT, L, None, 0, None))
