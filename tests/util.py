"""Shared helpers of the parity tests: scene packs, the oracle, and the comparison metrics of
BASELINE.json (hit ids bit-exact; <= 1 LSB on >= 99.9 % of pixels; PSNR >= 50 dB)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENES = ["cornellBox", "texture", "transform", "dragon_superlow", "dragon_low", "dragon_full", "outScene"]


def load_scene(name):
    from rayhs_b200 import Scene

    return Scene.from_pack(os.path.join(GOLDEN, name + ".pack"))


def oracle_for(scene):
    from oracle.orc import OracleScene

    return OracleScene(scene.raw)


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def compare_images(gpu_u8, ref_u8):
    diff = np.abs(gpu_u8.astype(np.int32) - ref_u8.astype(np.int32)).max(axis=-1)
    return dict(exact=float(np.mean(diff == 0)), within1=float(np.mean(diff <= 1)), maxdiff=int(diff.max()),
                psnr=psnr(gpu_u8, ref_u8))


def assert_parity(gpu_u8, ref_u8, what=""):
    m = compare_images(gpu_u8, ref_u8)
    assert m["within1"] >= 0.999, (what, m)   # BASELINE.json: <= 1 LSB per channel on >= 99.9 % of pixels
    assert m["psnr"] >= 50.0, (what, m)       # BASELINE.json: PSNR >= 50 dB
    return m


# ---------------------------------------------------------------------------------------------
# Hand-built raw scenes (rh_raw_scene) for analytic and edge-case tests.
class RawScene:
    """Keeps the ctypes arrays of an rh_raw_scene alive.  objects: list of dicts
    {kind:'plane'|'sphere'|'mesh', material:int, a,b,c (plane: point, normal, tangent; sphere: center, radius),
     positions [n,3], normals [n,3], uvs [n,2], indices [m]}; materials: list of dicts; lights: list of dicts."""

    def __init__(self, objects, materials, lights, textures=(), camera=None, size=(64, 64, 3)):
        import ctypes as C

        from rayhs_b200 import capi

        self._keep = []
        n = len(objects)
        objs = (capi.rh_raw_object * max(n, 1))()
        for i, o in enumerate(objects):
            ro = objs[i]
            ro.kind = {"plane": capi.RH_OBJ_PLANE, "sphere": capi.RH_OBJ_SPHERE, "mesh": capi.RH_OBJ_MESH}[o["kind"]]
            ro.material = o.get("material", 0)
            if o["kind"] == "plane":
                ro.a[:], ro.b[:], ro.c[:] = o["point"], o["normal"], o["tangent"]
            elif o["kind"] == "sphere":
                ro.a[:] = o["center"]
                ro.b[:] = (o["radius"], 0.0, 0.0)
            else:
                pos = np.ascontiguousarray(o["positions"], dtype=np.float64).reshape(-1, 3)
                nrm = np.ascontiguousarray(o.get("normals", np.zeros_like(pos)), dtype=np.float64).reshape(-1, 3)
                uv = np.ascontiguousarray(o.get("uvs", np.zeros((len(pos), 2))), dtype=np.float64).reshape(-1, 2)
                idx = np.ascontiguousarray(o["indices"], dtype=np.uint32).reshape(-1)
                self._keep += [pos, nrm, uv, idx]
                ro.n_verts, ro.n_indices = len(pos), len(idx)
                ro.positions = pos.ctypes.data_as(C.POINTER(C.c_double))
                ro.normals = nrm.ctypes.data_as(C.POINTER(C.c_double))
                ro.uvs = uv.ctypes.data_as(C.POINTER(C.c_double))
                ro.indices = idx.ctypes.data_as(C.POINTER(C.c_uint32))
        mats = (capi.rh_material * max(len(materials), 1))()
        kinds = {"mirror": 0, "diffuse": 1, "plastic": 2, "emmit": 3, "transparent": 4, "shownormal": 5, "showuv": 6}
        cmaps = {"flat": 0, "checker": 1, "texture": 2}
        for i, m in enumerate(materials):
            mm = mats[i]
            mm.kind = kinds[m["kind"]]
            mm.cmap_kind = cmaps[m.get("cmap", "flat")]
            mm.ior = m.get("ior", 1.5)
            mm.color1[:] = m.get("color1", (1, 1, 1))
            mm.color2[:] = m.get("color2", (0, 0, 0))
            mm.size = m.get("size", 1.0)
            mm.texture = m.get("texture", -1)
        lts = (capi.rh_light * max(len(lights), 1))()
        for i, l in enumerate(lights):
            ll = lts[i]
            ll.kind = capi.RH_LIGHT_POINT if l["kind"] == "point" else capi.RH_LIGHT_DIRECTIONAL
            ll.vec[:] = l["vec"]
            ll.color[:] = l.get("color", (1, 1, 1))
            ll.radius = l.get("radius", 1.0)
        texs = (capi.rh_texture * max(len(textures), 1))()
        texels = []
        for i, t in enumerate(textures):
            arr = np.asarray(t, dtype=np.float64)  # [h, w, 3]
            texs[i].h, texs[i].w, texs[i].offset = arr.shape[0], arr.shape[1], sum(len(x) for x in texels) // 3
            texels.append(arr.reshape(-1))
        tex_arr = np.ascontiguousarray(np.concatenate(texels) if texels else np.zeros(3))
        self._keep += [objs, mats, lts, texs, tex_arr]
        raw = capi.rh_raw_scene()
        raw.n_objects, raw.n_materials, raw.n_lights, raw.n_textures = n, len(materials), len(lights), len(textures)
        raw.objects, raw.materials, raw.lights, raw.textures = objs, mats, lts, texs
        raw.texels = tex_arr.ctypes.data_as(C.POINTER(C.c_double))
        raw.n_texels = (len(tex_arr) // 3) if texels else 0
        self.raw_struct = raw
        self.raw = C.pointer(raw)
        cam = capi.rh_camera()
        c = camera or {}
        cam.position[:] = c.get("position", (0, 0, -2))
        cam.target[:] = c.get("target", (0, 0, 0))
        cam.up[:] = c.get("up", (0, 1, 0))
        cam.projection = capi.RH_PROJ_ORTHOGRAPHIC if c.get("projection") == "orthographic" else capi.RH_PROJ_PERSPECTIVE
        cam.fovy = c.get("fovy", 0.9272952180016123)
        self.camera = cam
        self.width, self.height, self.max_depth = size
        self._flat = None
        self._dev = None

    # the same views rayhs_b200.Scene offers, so that rh.render() can take a RawScene
    @property
    def flat(self):
        import ctypes as C

        from rayhs_b200 import capi

        if self._flat is None:
            self._flat = C.c_void_p()
            capi.check(capi.lib().rh_flatten(self.raw, C.byref(self._flat)))
        return capi.lib().rh_flat_desc(self._flat)

    @property
    def device(self):
        import ctypes as C

        import rayhs_b200 as rh
        from rayhs_b200 import capi

        if self._dev is None:
            rh.init()
            self._dev = C.c_void_p()
            capi.check(capi.lib().rh_scene_create(self.flat, C.byref(self._dev)))
        return self._dev
