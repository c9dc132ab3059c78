"""Scene packs and sample offsets for the oracle, without the product library.

TEST INFRASTRUCTURE ONLY, like everything under oracle/: `bench.py --impl reference` times the CPU restatement of the
reference on the bench scene, and that arm should not map librayhs_b200.so at all.  This module reads a
tests/golden/*.pack file (the format csrc/frontend.cpp's rh_save_pack writes: header, camera, material / light / texture
tables, texels, then per object the POD record and — for a mesh — positions, normals, uvs, indices) into the ctypes
structs of rayhs_b200/capi.py (definitions only; nothing is loaded), and generates the harness's SplitMix64 offset stream
(the values rh_sample_offsets_f64 writes) with numpy.
"""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from rayhs_b200 import capi   # struct definitions only: capi.lib() is never called here

MAGIC = b"RHPK0001"
RH_OBJ_MESH = 2


class PackScene:
    """What tests/util's Scene offers the oracle: `.raw` (pointer to an rh_raw_scene), `.camera`, `.width`, `.height`,
    `.max_depth`.  Owns the arrays the raw scene points into."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            data = f.read()
        if data[:8] != MAGIC:
            raise ValueError(f"{path}: not a scene pack")
        at = 8
        self.width, self.height, self.max_depth = struct.unpack_from("<3i", data, at)
        at += 12

        def take(ctype, n=1):
            nonlocal at
            arr = (ctype * n).from_buffer_copy(data, at) if n else (ctype * 0)()
            at += C.sizeof(ctype) * n
            return arr

        self.camera = take(capi.rh_camera)[0]
        n_obj, n_mat, n_light, n_tex = struct.unpack_from("<4I", data, at)
        at += 16
        self._materials = take(capi.rh_material, n_mat)
        self._lights = take(capi.rh_light, n_light)
        self._textures = take(capi.rh_texture, n_tex)
        (n_texels,) = struct.unpack_from("<Q", data, at)   # doubles in the file: three per texel
        at += 8
        as_bytes = data[at]
        at += 1
        if as_bytes:   # one byte per value k / 255 (Bitmap.hs:28-29)
            self._texels = np.frombuffer(data, dtype=np.uint8, count=n_texels, offset=at).astype(np.float64) / 255
            at += n_texels
        else:
            self._texels = np.frombuffer(data, dtype=np.float64, count=n_texels, offset=at).copy()
            at += 8 * n_texels
        self._objects = (capi.rh_raw_object * n_obj)()
        self._mesh_arrays = []
        for i in range(n_obj):
            o = capi.rh_raw_object.from_buffer_copy(data, at)
            at += C.sizeof(capi.rh_raw_object)
            if o.kind == RH_OBJ_MESH:
                nv, ni = int(o.n_verts), int(o.n_indices)
                pos = np.frombuffer(data, dtype=np.float64, count=nv * 3, offset=at).copy()
                at += 24 * nv
                nrm = np.frombuffer(data, dtype=np.float64, count=nv * 3, offset=at).copy()
                at += 24 * nv
                uv = np.frombuffer(data, dtype=np.float64, count=nv * 2, offset=at).copy()
                at += 16 * nv
                idx = np.frombuffer(data, dtype=np.uint32, count=ni, offset=at).copy()
                at += 4 * ni
                self._mesh_arrays.append((pos, nrm, uv, idx))
                dp = C.POINTER(C.c_double)
                o.positions = pos.ctypes.data_as(dp)
                o.normals = nrm.ctypes.data_as(dp)
                o.uvs = uv.ctypes.data_as(dp)
                o.indices = idx.ctypes.data_as(C.POINTER(C.c_uint32))
            self._objects[i] = o
        if at != len(data):
            raise ValueError(f"{path}: {len(data) - at} bytes left over")
        self._raw = capi.rh_raw_scene()
        self._raw.n_objects, self._raw.n_materials, self._raw.n_lights, self._raw.n_textures = n_obj, n_mat, n_light, n_tex
        self._raw.objects = C.cast(self._objects, C.POINTER(capi.rh_raw_object))
        self._raw.materials = C.cast(self._materials, C.POINTER(capi.rh_material))
        self._raw.lights = C.cast(self._lights, C.POINTER(capi.rh_light))
        self._raw.textures = C.cast(self._textures, C.POINTER(capi.rh_texture))
        self._raw.texels = self._texels.ctypes.data_as(C.POINTER(C.c_double))
        self._raw.n_texels = n_texels // 3
        self.raw = C.pointer(self._raw)


def sample_offsets(n_pixels: int, spp: int, seed: int = 24) -> np.ndarray:
    """[n_pixels, spp, 2] float64: value i of the stream is mix(seed + (i + 1) * golden) >> 11, scaled by 2^-53, minus 0.5
    (SplitMix64; x before y per sample, RayHs.hs:185-188) — what rh_sample_offsets_f64(seed, ...) writes."""
    n = n_pixels * spp * 2
    out = np.empty(n, dtype=np.float64)
    golden, c1, c2 = np.uint64(0x9E3779B97F4A7C15), np.uint64(0xBF58476D1CE4E5B9), np.uint64(0x94D049BB133111EB)
    block = 1 << 24
    with np.errstate(over="ignore"):
        for b in range(0, n, block):
            k = np.arange(b + 1, min(n, b + block) + 1, dtype=np.uint64)
            z = np.uint64(seed) + k * golden
            z = (z ^ (z >> np.uint64(30))) * c1
            z = (z ^ (z >> np.uint64(27))) * c2
            z ^= z >> np.uint64(31)
            out[b:b + len(k)] = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 0.5
    return out.reshape(n_pixels, spp, 2)
