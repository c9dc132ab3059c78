"""A second, independent restatement of the reference's whole per-pixel path, for pinning the oracle.

Written from the Haskell sources alone (file:line cited per function), in scalar Python with numpy only over the
triangles of a mesh: no tree (every triangle of every mesh is tested, first minimum wins), no shared code with
oracle/oracle.cpp or with the product.  Slow by design — tests run it at 16..32 pixels a side.

Input is the same POD rh_raw_scene the oracle takes (objects before any tree build).
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np

EPS = 0.000001  # Geometry.hs:31-32
PI_INV = 1 / math.pi  # Math.hs


def vec(a):
    return np.array(a[:3], dtype=np.float64)


def dot(a, b):  # Vec.hs:103-105 (left to right)
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def cross(a, b):  # Vec.hs:107-110
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def normalize(v):  # Vec.hs:124-126: multiply by the reciprocal of the length
    return (1 / math.sqrt(dot(v, v))) * v


def mod1(n, d):  # Data.Fixed.mod': n - floor(n / d) * d with an exact rational floor
    if d == 0 or not math.isfinite(n) or not math.isfinite(d):
        return float("nan")
    f = math.floor(Fraction(n) / Fraction(d))
    return n - float(f) * d


class Scene:
    def __init__(self, raw):
        self.objects = []
        for i in range(raw.n_objects):
            o = raw.objects[i]
            ob = dict(kind=o.kind, material=o.material, a=vec(o.a), b=vec(o.b), c=vec(o.c))
            if o.kind == 2:  # mesh: `triangles mesh` = the index list in threes (Mesh.hs:105-109)
                nv, ni = o.n_verts, o.n_indices
                if ni:
                    P = np.ctypeslib.as_array(o.positions, shape=(nv, 3)).copy()
                    N = np.ctypeslib.as_array(o.normals, shape=(nv, 3)).copy()
                    U = np.ctypeslib.as_array(o.uvs, shape=(nv, 2)).copy()
                    I = np.ctypeslib.as_array(o.indices, shape=(ni,)).reshape(-1, 3).astype(np.int64)
                    ob.update(p0=P[I[:, 0]], p1=P[I[:, 1]], p2=P[I[:, 2]], n0=N[I[:, 0]], n1=N[I[:, 1]], n2=N[I[:, 2]],
                              uv0=U[I[:, 0]], uv1=U[I[:, 1]], uv2=U[I[:, 2]])
                else:
                    ob["p0"] = np.zeros((0, 3))
            self.objects.append(ob)
        self.materials = []
        for i in range(raw.n_materials):
            m = raw.materials[i]
            self.materials.append(dict(kind=m.kind, cmap=m.cmap_kind, ior=m.ior, c1=vec(m.color1), c2=vec(m.color2), size=m.size,
                                       texture=m.texture))
        self.lights = [dict(kind=raw.lights[i].kind, vec=vec(raw.lights[i].vec), color=vec(raw.lights[i].color),
                            radius=raw.lights[i].radius) for i in range(raw.n_lights)]
        self.textures = []
        for i in range(raw.n_textures):
            t = raw.textures[i]
            tex = np.ctypeslib.as_array(raw.texels, shape=(raw.n_texels * 3,))[t.offset * 3:(t.offset + t.w * t.h) * 3]
            self.textures.append(tex.reshape(t.h, t.w, 3).copy())


# ---------------------------------------------------------------- intersections (Geometry.hs, Mesh.hs)
def hit_plane(o, d, ob):  # Geometry.hs:70-79
    p, n, t = ob["a"], ob["b"], ob["c"]
    ddn = dot(d, n)
    with np.errstate(divide="ignore", invalid="ignore"):
        time = np.float64(dot(n, p - o)) / np.float64(ddn)
    if not (abs(ddn) > 0 and time > 0):
        return None
    pos = o + time * d
    b = cross(t, n)
    rel = pos - p
    return pos, n, (dot(t, rel), dot(b, rel)), float(time)


def hit_sphere(o, d, ob):  # Geometry.hs:81-96
    ct, r = ob["a"], ob["b"][0]
    a = dot(d, d)
    b = 2.0 * dot(d, o - ct)
    c = dot(o - ct, o - ct) - r * r
    delta = b * b - 4.0 * a * c
    if delta < 0.0:
        return None
    for t in (0.5 * ((-b) - math.sqrt(delta)) / a, 0.5 * ((-b) + math.sqrt(delta)) / a):
        if t > 0:
            p = o + t * d
            n = normalize(p - ct)
            with np.errstate(divide="ignore", invalid="ignore"):
                polar = (PI_INV * math.atan(np.float64(n[2]) / np.float64(n[0])), PI_INV * math.acos(n[1]))  # atan, not atan2
            return p, n, polar, t
    return None


def hit_mesh(o, d, ob):  # Mesh.hs:59-82 over every triangle; closestHit keeps the first minimum (Geometry.hs:54-57)
    if len(ob["p0"]) == 0:
        return None
    e1, e2 = ob["p1"] - ob["p0"], ob["p2"] - ob["p0"]
    p = np.cross(d[None], e2)
    det = e1[:, 0] * p[:, 0] + e1[:, 1] * p[:, 1] + e1[:, 2] * p[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        idet = 1 / det
        t0 = o[None] - ob["p0"]
        u = idet * (t0[:, 0] * p[:, 0] + t0[:, 1] * p[:, 1] + t0[:, 2] * p[:, 2])
        q = np.cross(t0, e1)
        v = idet * (d[0] * q[:, 0] + d[1] * q[:, 1] + d[2] * q[:, 2])
        t = idet * (e2[:, 0] * q[:, 0] + e2[:, 1] * q[:, 1] + e2[:, 2] * q[:, 2])
        miss = (np.abs(det) < EPS) | (u < 0) | (u > 1) | (v < 0) | ((u + v) > 1) | (t < EPS)
    miss |= np.isnan(u) | np.isnan(v) | np.isnan(t)
    if miss.all():
        return None
    k = int(np.argmin(np.where(miss, np.inf, t)))
    uu, vv, tt = float(u[k]), float(v[k]), float(t[k])
    w = 1 - uu - vv
    n = uu * ob["n1"][k] + vv * ob["n2"][k] + w * ob["n0"][k]  # not renormalised
    uv = uu * ob["uv1"][k] + vv * ob["uv2"][k] + w * ob["uv0"][k]
    return o + tt * d, n, (float(uv[0]), float(uv[1])), tt


def intersections(sc, o, d):  # RayHs.hs:58-65
    out = []
    for ob in sc.objects:
        h = (hit_plane, hit_sphere, hit_mesh)[ob["kind"]](o, d, ob)
        if h is not None:
            out.append(h + (sc.materials[ob["material"]],))
    return out


def closest(sc, o, d):  # RayHs.hs:67-71: minimumBy keeps the first of equal times
    hits = intersections(sc, o, d)
    return min(hits, key=lambda h: h[3]) if hits else None


def shadowed(sc, light, o, d):  # RayHs.hs:74-87
    for (pos, _, _, _, mat) in intersections(sc, o, d):
        front = True if light["kind"] == 0 else dot(o - light["vec"], o - light["vec"]) > dot(o - pos, o - pos)
        if front and mat["kind"] != 3:
            return True
    return False


# ---------------------------------------------------------------- shading (Light.hs, Material.hs, ColorMap.hs, RayHs.hs)
def light_at(light, p):  # Light.hs:12-17
    if light["kind"] == 0:
        return light["vec"], light["color"]
    dv = light["vec"] - p
    dd = math.sqrt(dot(dv, dv))
    s = 1.0 + dd / light["radius"]
    return (1 / dd) * dv, (1.0 / (s * s)) * light["color"]


def color_at(sc, mat, uv):  # ColorMap.hs:18-58
    if mat["cmap"] == 0:
        return mat["c1"]
    u, v = uv
    if mat["cmap"] == 1:
        s = mat["size"]
        return mat["c1"] if (mod1(u, s) - 0.5 * s) * (mod1(v, s) - 0.5 * s) < 0 else mat["c2"]
    tex = sc.textures[mat["texture"]]
    h, w = tex.shape[:2]
    uu, vv = mod1(u, 1) * w, mod1(v, 1) * h
    ui, vi = round(uu), round(vv)  # Python's round is half-to-even, like Haskell's
    x0, x1, y0, y1 = (ui - 1) % w, ui % w, (vi - 1) % h, vi % h
    lx, ly = uu - (ui - 1) - 0.5, vv - (vi - 1) - 0.5
    c0, c1, c2, c3 = tex[y0, x0], tex[y0, x1], tex[y1, x0], tex[y1, x1]  # Bitmap.hs:17-18: row-major, (x, y)
    cx0 = lx * c1 + (1 - lx) * c0
    cx1 = lx * c3 + (1 - lx) * c2
    return ly * cx1 + (1 - ly) * cx0


def r0(n1, n2):  # Material.hs:22-24
    q = (n1 - n2) / (n1 + n2)
    return q * q


def fresnel(ior, cos0):  # Material.hs:26-29; (^5) by repeated squaring
    r = r0(1.0, ior)
    x = 1 - cos0
    x2 = x * x
    return r + (1 - r) * ((x2 * x2) * x)


def reflect(v, n):  # Vec.hs:128-130
    return v - (2 * dot(v, n)) * n


def refract(i, n, n1, n2):  # Vec.hs:132-140
    n1n2 = n1 / n2
    cos0 = -dot(i, n)
    sin20 = n1n2 * n1n2 * (1 - cos0 * cos0)
    if sin20 > 1:
        return None
    return n1n2 * i + (n1n2 * cos0 - math.sqrt(1.0 - sin20)) * n


def accum_diffuse(sc, p, n, cd, count):  # RayHs.hs:89-97; diffuse: Material.hs:31-33
    c = np.zeros(3)
    for light in sc.lights:
        ld, lc = light_at(light, p)
        count["shadow"] += 1
        if not shadowed(sc, light, p + EPS * ld, ld):
            c = c + (max(dot(ld, n), 0) * PI_INV) * (cd * lc)
    return c


def specular(sc, depth, max_depth, v, p, n, count):  # RayHs.hs:99-104
    if depth < max_depth:
        r = reflect(v, n)
        count["reflect"] += 1
        return dot(r, n) * trace_ray(sc, depth + 1, max_depth, p + EPS * r, r, count)
    return np.zeros(3)


def irradiance(sc, d, max_depth, mat, v, p, n, uv, count):  # RayHs.hs:107-147
    k = mat["kind"]
    if k == 1:  # Diffuse
        cd = color_at(sc, mat, uv)
        return 0.2 * cd + accum_diffuse(sc, p, n, cd, count)
    if k == 2:  # Plastic
        cd = color_at(sc, mat, uv)
        return accum_diffuse(sc, p, n, cd, count) + fresnel(mat["ior"], dot(n, -v)) * specular(sc, d, max_depth, v, p, n, count)
    if k == 0:  # Mirror
        return fresnel(mat["ior"], dot(n, -v)) * specular(sc, d, max_depth, v, p, n, count)
    if k == 3:  # Emmit
        return mat["c1"]
    if k == 4:  # Transparent
        ior = mat["ior"]
        radiance = None
        if d != max_depth:
            rd = refract(v, n, 1.0, ior)
            if rd is not None:
                count["probe"] += 1
                h = closest(sc, p + EPS * rd, rd)
                if h is not None:
                    outp, outn = h[0], h[1]
                    od = refract(rd, -outn, ior, 1.0)
                    if od is not None:
                        count["exit"] += 1
                        radiance = trace_ray(sc, d + 1, max_depth, outp + EPS * od, od, count)
        spec = fresnel(ior, dot(n, -v)) * specular(sc, d, max_depth, v, p, n, count)
        return spec if radiance is None else (1 - r0(ior, 1.0)) * radiance + spec
    if k == 5:  # ShowNormal
        return np.array(n, dtype=np.float64)
    return np.array([uv[0], uv[1], 0.0])  # ShowUV


def trace_ray(sc, depth, max_depth, o, d, count):  # RayHs.hs:149-154
    h = closest(sc, o, d)
    if h is None:
        return np.zeros(3)
    p, n, uv, _, mat = h
    return irradiance(sc, depth, max_depth, mat, d, p, n, uv, count)


# ---------------------------------------------------------------- camera and image (Projection.hs, Mat.hs, Image.hs)
def ray_from_pixel(cam, w, h, px, py):  # Projection.hs:22-46, Mat.hs:83-93, 40-44
    pos, target, up = vec(cam.position), vec(cam.target), vec(cam.up)
    aspect = w / h
    apw, aph = (w, w / aspect) if aspect > 1 else (aspect * h, h)
    x, y = apw * (px - (w / 2)) / w, aph * ((-py) + (h / 2)) / h
    if cam.projection == 0:  # orthographic
        o, d = np.array([x, y, 0.0]), np.array([0.0, 0.0, 1.0])
    else:
        f = 0.5 * h / (math.tan(0.5) * cam.fovy)  # `tan 0.5 * fovy` parses as (tan 0.5) * fovy
        o, d = np.zeros(3), normalize(np.array([x, y, f]))
    forward = normalize(target - pos)
    right = normalize(cross(up, forward))
    upv = cross(forward, right)
    # fromColumns right up forward; apply: rows of the matrix dotted with the vector
    rows = [np.array([right[k], upv[k], forward[k]]) for k in range(3)]
    return o + pos, np.array([dot(rows[0], d), dot(rows[1], d), dot(rows[2], d)])


def render(raw, cam, w, h, max_depth, spp=1, offsets=None):
    """rayTrace (RayHs.hs:161-166) for pixels (i mod w, i div w), no half-pixel offset (Image.hs:31-32); with `offsets`
    [w*h, spp, 2] distributedRayTrace (RayHs.hs:169-195): samples at (i + ox, j + oy), average = (1/n) * foldl (+) black.
    Returns (rgb float64 [h, w, 3], rgb u8 via toIntC, ray counts)."""
    sc = Scene(raw)
    img = np.zeros((h, w, 3))
    count = dict(primary=0, reflect=0, probe=0, exit=0, shadow=0)
    for j in range(h):
        for i in range(w):
            if offsets is None:
                o, d = ray_from_pixel(cam, float(w), float(h), float(i), float(j))
                count["primary"] += 1
                img[j, i] = trace_ray(sc, 0, max_depth, o, d, count)
            else:
                acc = np.zeros(3)
                for s in range(spp):
                    ox, oy = offsets[j * w + i, s]
                    o, d = ray_from_pixel(cam, float(w), float(h), float(i) + ox, float(j) + oy)
                    count["primary"] += 1
                    acc = acc + trace_ray(sc, 0, max_depth, o, d, count)
                img[j, i] = (1 / spp) * acc
    with np.errstate(invalid="ignore"):
        u8 = np.clip(np.trunc(255 * np.where(np.isnan(img), 1.0, np.minimum(img, 1.0))), 0, 255).astype(np.uint8)  # Image.hs:54-55
    return img, u8, count
