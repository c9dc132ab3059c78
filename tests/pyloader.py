"""An independent scene loader, written from the Haskell sources alone — src/JSON.hs (schema), src/Descriptors.hs,
src/MaterialDescriptors.hs, src/Mesh.hs:118-221 (OBJ), src/Bitmap.hs:20-37 (PPM), src/Transform.hs, src/Mat.hs:55-81 —
and sharing no code with the C++ front end (csrc/frontend.cpp).  TEST INFRASTRUCTURE: tests/test_pyloader.py parses
the reference's own data/ files with it and compares the result, array by array, with what the front end put into
the committed scene packs, so that a loader bug cannot hide behind every parity test consuming the same parse.
"""
from __future__ import annotations

import json
import math
import os

import numpy as np


# ------------------------------------------------------------------ Vec / Mat (Vec.hs, Mat.hs)
def _vec(o):
    return np.array([o["x"], o["y"], o["z"]], dtype=np.float64)   # JSON.hs:22-27


def _color(o):
    return np.array([o["r"], o["g"], o["b"]], dtype=np.float64)   # JSON.hs:29-34


def _normalize(v):
    return (1.0 / math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])) * v   # Vec.hs:124-126: multiply by the reciprocal


def _cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def _apply(m, v):   # Mat.hs:40-44, rows a b c / d e f / g h i, left-associated sums
    return np.array([m[0] * v[0] + m[1] * v[1] + m[2] * v[2], m[3] * v[0] + m[4] * v[1] + m[5] * v[2], m[6] * v[0] + m[7] * v[1] + m[8] * v[2]])


def _matmul(p, q):  # Mat.hs:25-31
    a1, b1, c1, d1, e1, f1, g1, h1, i1 = p
    a2, b2, c2, d2, e2, f2, g2, h2, i2 = q
    return [a1 * a2 + b1 * d2 + c1 * g2, a1 * b2 + b1 * e2 + c1 * h2, a1 * c2 + b1 * f2 + c1 * i2,
            d1 * a2 + e1 * d2 + f1 * g2, d1 * b2 + e1 * e2 + f1 * h2, d1 * c2 + e1 * f2 + f1 * i2,
            g1 * a2 + h1 * d2 + i1 * g2, g1 * b2 + h1 * e2 + i1 * h2, g1 * c2 + h1 * f2 + i1 * i2]


def _rot_x(a):
    return [1, 0, 0, 0, math.cos(a), -math.sin(a), 0, math.sin(a), math.cos(a)]     # Mat.hs:70-74


def _rot_y(a):
    return [math.cos(a), 0, math.sin(a), 0, 1, 0, -math.sin(a), 0, math.cos(a)]     # Mat.hs:64-68


def _rot_z(a):
    return [math.cos(a), -math.sin(a), 0, math.sin(a), math.cos(a), 0, 0, 0, 1]     # Mat.hs:58-62


def _orthonormal(r):   # Vec.hs:151-158
    x, y, z = abs(r[0]), abs(r[1]), abs(r[2])
    if x < y and x < z:
        s = _normalize(np.array([0.0, -r[2], r[1]]))
    elif y < x and y < z:
        s = _normalize(np.array([-r[2], 0.0, r[0]]))
    else:
        s = _normalize(np.array([-r[1], r[0], 0.0]))
    return r, s, _cross(r, s)


def _rotate(axis, angle):   # Mat.hs:76-81: transpose m * rotateX angle * m, m = fromColumns r s t
    r, s, t = _orthonormal(axis)
    m = [r[0], s[0], t[0], r[1], s[1], t[1], r[2], s[2], t[2]]
    mt = [m[0], m[3], m[6], m[1], m[4], m[7], m[2], m[5], m[8]]
    return _matmul(_matmul(mt, _rot_x(angle)), m)


def transform_point(t, p):   # Transform.hs:16-24
    k = t["type"]
    if k == "translate":
        return p + _vec(t["vector"])
    if k == "scale":
        return _vec(t["vector"]) * p
    if k == "rotateX":
        return _apply(_rot_x(t["angle"]), p)
    if k == "rotateY":
        return _apply(_rot_y(t["angle"]), p)
    if k == "rotateZ":
        return _apply(_rot_z(t["angle"]), p)
    if k == "rotate":
        return _apply(_rotate(_vec(t["axis"]), t["angle"]), p)
    if k == "sequence":
        for s in t["transforms"]:   # foldl (flip transform) p transforms: first to last
            p = transform_point(s, p)
        return p
    raise ValueError("unknown transform " + k)


# ------------------------------------------------------------------ OBJ (Mesh.hs:118-221)
def read_obj(path):
    pos, nrm, uvs, inds = [], [], [], []
    with open(path) as f:
        for line in f.read().split("\n"):          # `lines`
            if line.startswith("v "):
                w = line[2:].split()
                if len(w) == 3:
                    pos.append([float(x) for x in w])
            elif line.startswith("vn "):
                w = line[3:].split()
                if len(w) == 3:
                    nrm.append([float(x) for x in w])
            elif line.startswith("vt "):
                w = line[3:].split()
                if len(w) == 2:
                    uvs.append([float(x) for x in w])
            elif line.startswith("f "):
                face = []
                for tok in line[2:].split():
                    parts = tok.split("/")
                    vals = []
                    for s in parts:
                        try:
                            vals.append(int(s))
                        except ValueError:
                            vals.append(None)           # readMaybe
                    if vals[0] is None:
                        continue                        # parseIndex _ = Nothing: the token is dropped (mapMaybe)
                    pi, ti, ni = vals[0], (vals[1] if len(vals) > 1 else None), (vals[2] if len(vals) > 2 else None)
                    face.append((pi, ti, ni))
                if len(face) > 2:
                    inds.extend(face)                   # all faces' indices in one list (Mesh.hs:183)
    # flattenVertices (Mesh.hs:195-212): one vertex per distinct (p, t, n) triple, numbered in order of first use
    verts_p, verts_n, verts_uv, ids, seen = [], [], [], [], {}
    for key in inds:
        k = seen.get(key)
        if k is None:
            k = len(verts_p)
            seen[key] = k
            pi, ti, ni = key
            verts_p.append(pos[pi - 1])
            verts_n.append(nrm[ni - 1] if ni is not None else [0.0, 0.0, 0.0])
            verts_uv.append(uvs[ti - 1] if ti is not None else [0.0, 0.0])
        ids.append(k)
    n_tri = len(ids) // 3     # `faces` regroups the index list in threes (Mesh.hs:105-109)
    return (np.array(verts_p, dtype=np.float64).reshape(-1, 3), np.array(verts_n, dtype=np.float64).reshape(-1, 3),
            np.array(verts_uv, dtype=np.float64).reshape(-1, 2), np.array(ids[:3 * n_tri], dtype=np.uint32))


def transform_mesh(p, n, t):   # Mesh.hs:89-101: Translate moves positions only; EVERY other transform also hits the normals
    p2 = np.array([transform_point(t, v) for v in p]).reshape(-1, 3)
    if t["type"] == "translate":
        return p2, n
    return p2, np.array([transform_point(t, v) for v in n]).reshape(-1, 3)


# ------------------------------------------------------------------ PPM texture (Bitmap.hs:20-37)
def read_ppm(path):
    with open(path) as f:
        lines = f.read().split("\n")
    w, h = [int(x) for x in lines[1].split()][:2]
    vals = [float(int(x)) / 255 for ln in lines[3:] for x in ln.split()]   # `/ 255` whatever the header's maxval says
    n = len(vals) // 3
    return w, h, np.array(vals[:3 * n], dtype=np.float64).reshape(n, 3)


# ------------------------------------------------------------------ scene (Descriptors.hs, MaterialDescriptors.hs)
def load_scene(json_path, base_dir):
    """Returns dict(width, height, max_depth, camera, objects, materials, lights, textures) in the form tests/util.RawScene
    takes: one material per object, in object order (the front end does the same: Object Geometry Material)."""
    with open(json_path) as f:
        d = json.load(f)
    objects, materials, textures = [], [], []
    for od in d["scene"]["objects"]:
        g, m = od["geometry"], od["material"]
        mat = {"kind": m["type"]}
        if m["type"] in ("mirror", "plastic", "transparent"):
            mat["ior"] = m["ior"]
        if m["type"] == "emmit":
            mat["color1"] = _color(m["ce"])
        if m["type"] in ("diffuse", "plastic"):
            cd = m["cd"]
            mat["cmap"] = cd["type"]
            if cd["type"] == "flat":
                mat["color1"] = _color(cd["color"])
            elif cd["type"] == "checker":
                mat["color1"], mat["color2"], mat["size"] = _color(cd["color1"]), _color(cd["color2"]), cd["size"]
            else:
                w, h, tex = read_ppm(os.path.join(base_dir, cd["fileName"]))
                mat["texture"] = len(textures)
                textures.append(tex.reshape(h, w, 3) if len(tex) == w * h else (w, h, tex))
        materials.append(mat)
        o = {"material": len(materials) - 1}
        if g["type"] == "sphere":
            o.update(kind="sphere", center=_vec(g["center"]), radius=g["radius"])
        elif g["type"] == "plane":
            o.update(kind="plane", point=_vec(g["point"]), normal=_vec(g["normal"]), tangent=_vec(g["tangent"]))
        else:
            p, n, uv, idx = read_obj(os.path.join(base_dir, g["fileName"]))
            p, n = transform_mesh(p, n, g["transform"])
            o.update(kind="mesh", positions=p, normals=n, uvs=uv, indices=idx)
        objects.append(o)
    lights = []
    for l in d["scene"]["lights"]:
        if l["type"] == "directional":
            lights.append({"kind": "directional", "vec": _vec(l["direction"]), "color": _color(l["color"])})
        else:
            lights.append({"kind": "point", "vec": _vec(l["position"]), "color": _color(l["color"]), "radius": l["radius"]})
    c = d["camera"]
    cam = {"position": _vec(c["position"]), "target": _vec(c["target"]), "up": _vec(c["up"]), "projection": c["projection"]["type"],
           "fovy": c["projection"].get("fovy", 0.0)}
    return dict(width=d["width"], height=d["height"], max_depth=d["maxDepth"], camera=cam, objects=objects, materials=materials,
                lights=lights, textures=textures)
