"""CPU tests of the boundary: librayhs_b200.so loads, exports every symbol include/rayhs_b200.h declares,
the struct layouts match, and the product fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from rayhs_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rayhs_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rh_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    L = capi.lib()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/rayhs_b200.h but not exported"
        assert s in capi.SIGNATURES, f"{s} has no ctypes signature"
    assert L.rh_abi_version() == 2


def test_struct_layouts_match_the_header():
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirror."""
    names = ["rh_material", "rh_light", "rh_texture", "rh_camera", "rh_raw_object", "rh_raw_scene", "rh_node", "rh_tri",
             "rh_tri_shade", "rh_object", "rh_scene_desc", "rh_render_opts", "rh_stats"]
    src = '#include <stdio.h>\n#include "rayhs_b200.h"\nint main(void){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    exe = os.path.join(ROOT, "build", "abi_sizes")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    with open(exe + ".c", "w") as f:
        f.write(src)
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), exe + ".c", "-o", exe])  # header is plain C
    out = dict(line.split() for line in subprocess.check_output([exe], text=True).splitlines())
    for n in names:
        assert int(out[n]) == C.sizeof(getattr(capi, n)), n


def test_no_gpu_means_loud_failure_not_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = capi.lib()
    rc = L.rh_init(-1)
    assert rc == capi.RH_ERR_CUDA
    assert b"no CPU fallback" in L.rh_last_error()
    out = C.c_void_p()
    assert L.rh_scene_create(None, C.byref(out)) == capi.RH_ERR_STATE
    assert L.rh_render(None, None, None, None, None, None) == capi.RH_ERR_STATE


def test_product_never_references_the_oracle():
    """rayhs_b200/ (the product) must not import, link or load anything under oracle/ (build.py only compiles it)."""
    pkg = os.path.join(ROOT, "rayhs_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cpp", ".cu", ".cuh", ".h")) and fn != "build.py":
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "liborc" not in text and "oracle.orc" not in text and "from oracle" not in text, fn
    deps = subprocess.check_output(["ldd", capi.LIB_PATH], text=True)
    assert "liborc" not in deps


def test_shard_row_arithmetic():
    L = capi.lib()
    assert L.rh_shard_rows(2160, 8, 16) == 272 and L.rh_shard_rows(2160, 8, 4) == 272 and L.rh_shard_rows(2160, 1, 2160) == 2160
    assert L.rh_shard_rows(10, 2, 2) == 6
    assert L.rh_default_band_height(2160, 1) == 2160 and L.rh_default_band_height(2160, 8) == 4
    import rayhs_b200 as rh

    for H, G, bh in ((10, 2, 2), (2160, 8, 16), (7, 4, 1), (150, 8, 4)):
        seen = []
        for g in range(G):
            rows = rh.shard_global_rows(H, g, G, bh)
            assert len(rows) == L.rh_shard_rows(H, G, bh)
            seen += [r for r in rows if r >= 0]
        assert sorted(seen) == list(range(H))


def test_struct_offsets_the_haskell_shim_pokes():
    """INTEGRATION.md's B200FFI.hs fills rh_scene_desc and rh_render_opts with pokeByteOff at fixed offsets; they must be
    the offsets of the C structs (ctypes lays the mirrored structs out like the C compiler does)."""
    import ctypes as C

    from rayhs_b200 import capi

    d = capi.rh_scene_desc
    assert [getattr(d, f).offset for f in ("n_objects", "n_materials", "n_lights", "n_textures", "n_nodes", "n_tris")] == [0, 4, 8, 12, 16, 20]
    assert [getattr(d, f).offset for f in ("objects", "materials", "lights", "textures", "texels", "n_texels", "nodes", "tris", "tri_shade")] == \
        [24, 32, 40, 48, 56, 64, 72, 80, 88]
    assert C.sizeof(d) == 96
    o = capi.rh_render_opts
    assert [getattr(o, f).offset for f in ("width", "height", "max_depth", "spp", "offset_mode", "offset_tile")] == [0, 4, 8, 12, 16, 20]
    assert o.offsets.offset == 24 and o.shard_index.offset == 32 and o.shard_count.offset == 36 and o.flags.offset == 48
    assert o.peer_frames.offset == 56 and C.sizeof(o) == 64
    assert C.sizeof(capi.rh_camera) == 112
