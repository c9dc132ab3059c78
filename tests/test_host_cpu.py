"""CPU tests of the host logic around the C ABI: the front end (JSON / OBJ / PPM loaders, Transform.hs
semantics), the pack format, the sample-offset stream, the P3 writer, the CLI argument parsing, and
the N > 1 shard path over gloo (world_size 2) with the oracle standing in for the renderer."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import rayhs_b200 as rh
from rayhs_b200 import capi
from tests.util import load_scene, oracle_for

TINY_OBJ = """# a square of two triangles
v -1 -1 0
v 1 -1 0
v 1 1 0
v -1 1 0
vn 0 0 -1
vt 0 0
vt 1 0
vt 1 1
vt 0 1
f 1/1/1 2/2/1 3/3/1
f 1/1/1 3/3/1 4/4/1
"""


def tiny_scene(tmp_path, transform):
    (tmp_path / "data").mkdir(exist_ok=True)
    (tmp_path / "data" / "sq.obj").write_text(TINY_OBJ)
    (tmp_path / "data" / "t.ppm").write_text("P3\n2 1\n255\n255\n0\n0\n0\n255\n51\n")
    scene = {"width": 32, "height": 16, "maxDepth": 2,
             "camera": {"position": {"x": 0, "y": 0, "z": -3}, "target": {"x": 0, "y": 0, "z": 0}, "up": {"x": 0, "y": 1, "z": 0},
                        "projection": {"type": "perspective", "fovy": 0.9, "width": 1, "height": 1, "near": 1}},
             "scene": {"objects": [
                 {"geometry": {"type": "mesh", "fileName": "data/sq.obj", "transform": transform},
                  "material": {"type": "plastic", "ior": 1.9, "cd": {"type": "texture", "fileName": "data/t.ppm"}}},
                 {"geometry": {"type": "sphere", "center": {"x": 0, "y": 0, "z": 0}, "radius": 7.0e-2},
                  "material": {"type": "emmit", "ce": {"r": 1, "g": 2, "b": 3}}},
                 {"geometry": {"type": "plane", "point": {"x": 0, "y": -1, "z": 0}, "normal": {"x": 0, "y": 1, "z": 0},
                               "tangent": {"x": 1, "y": 0, "z": 0}},
                  "material": {"type": "diffuse", "cd": {"type": "checker", "color1": {"r": 1, "g": 1, "b": 1},
                                                         "color2": {"r": 0, "g": 0, "b": 0}, "size": 0.5}}}],
                 "lights": [{"type": "point", "position": {"x": 0, "y": 2, "z": -1}, "color": {"r": 9, "g": 9, "b": 9}, "radius": 0.5},
                            {"type": "directional", "direction": {"x": 0, "y": -1, "z": 1}, "color": {"r": 1, "g": 1, "b": 1}}]}}
    p = tmp_path / "scene.json"
    p.write_text(json.dumps(scene))
    return str(p)


def test_json_obj_ppm_front_end(tmp_path):
    """JSON.hs schema, Mesh.hs OBJ faces p/t/n, Bitmap.hs texels = byte/255, Descriptors.hs object order."""
    path = tiny_scene(tmp_path, {"type": "translate", "vector": {"x": 0, "y": 0, "z": 1}})
    sc = rh.Scene.from_json(path, str(tmp_path))
    assert (sc.width, sc.height, sc.max_depth) == (32, 16, 2)
    raw = sc.raw.contents
    assert raw.n_objects == 3 and raw.n_lights == 2 and raw.n_textures == 1 and raw.n_texels == 2
    assert [raw.objects[i].kind for i in range(3)] == [capi.RH_OBJ_MESH, capi.RH_OBJ_SPHERE, capi.RH_OBJ_PLANE]
    m = raw.objects[0]
    assert m.n_indices == 6 and m.n_verts == 4
    pos = np.ctypeslib.as_array(m.positions, shape=(4, 3))
    assert np.array_equal(pos[:, 2], [1, 1, 1, 1])                        # Translate moves positions ...
    nrm = np.ctypeslib.as_array(m.normals, shape=(4, 3))
    assert np.array_equal(nrm, [[0, 0, -1]] * 4)                          # ... and leaves normals alone (Mesh.hs:89-94)
    tex = np.ctypeslib.as_array(raw.texels, shape=(2, 3))
    assert np.allclose(tex, [[1, 0, 0], [0, 1, 0.2]])                     # byte / 255 (Bitmap.hs:28-29)
    assert raw.objects[1].b[0] == 0.07                                    # 7.0e-2 parses
    mat0 = raw.materials[raw.objects[0].material]
    assert mat0.kind == capi.RH_MAT_PLASTIC and mat0.cmap_kind == capi.RH_CMAP_TEXTURE and mat0.ior == 1.9
    assert raw.lights[1].kind == capi.RH_LIGHT_DIRECTIONAL and tuple(raw.lights[1].vec) == (0, -1, 1)   # un-normalised (Light.hs:14)


def test_non_translate_transforms_also_hit_normals(tmp_path):
    """Mesh.hs:95-101 / SURVEY App. A-P4: scale and sequences are applied to normals like to positions."""
    path = tiny_scene(tmp_path, {"type": "sequence", "transforms": [{"type": "scale", "vector": {"x": 2, "y": 2, "z": 2}},
                                                                   {"type": "translate", "vector": {"x": 0, "y": 0, "z": 1}}]})
    sc = rh.Scene.from_json(path, str(tmp_path))
    m = sc.raw.contents.objects[0]
    nrm = np.ctypeslib.as_array(m.normals, shape=(4, 3))
    assert np.array_equal(nrm, [[0, 0, -1]] * 4)   # (0,0,-1)*2 + (0,0,1): the translate inside a sequence moves normals too
    pos = np.ctypeslib.as_array(m.positions, shape=(4, 3))
    assert np.array_equal(pos[0], [-2, -2, 1])


def test_bad_scene_is_an_error_not_a_crash(tmp_path):
    bad = tmp_path / "bad.json"
    bad.write_text('{"width": 4}')
    with pytest.raises(capi.RayHsError) as e:
        rh.Scene.from_json(str(bad), str(tmp_path))
    assert e.value.code == capi.RH_ERR_IO and "Failed to read scene" in str(e.value)
    with pytest.raises(capi.RayHsError):
        rh.Scene.from_json(str(tmp_path / "missing.json"))


def test_pack_round_trip(tmp_path):
    sc = load_scene("cornellBox")
    p = str(tmp_path / "c.pack")
    sc.save_pack(p)
    sc2 = rh.Scene.from_pack(p)
    a, b = sc.flat.contents, sc2.flat.contents
    assert (a.n_nodes, a.n_tris, a.n_objects) == (b.n_nodes, b.n_tris, b.n_objects)
    assert bytes(C.string_at(a.tris, a.n_tris * 80)) == bytes(C.string_at(b.tris, b.n_tris * 80))
    r1 = oracle_for(sc).render(sc.camera, 24, 24, 3)
    r2 = oracle_for(sc2).render(sc2.camera, 24, 24, 3)
    assert np.array_equal(r1["rgb_u8"], r2["rgb_u8"])


def test_sample_offsets_stream_shape():
    """RayHs.hs:173-188: one stream, pixel-major, x before y, values in [-0.5, 0.5)."""
    a = rh.sample_offsets(6, 4, seed=24)
    b = rh.sample_offsets(3, 4, seed=24)
    assert a.shape == (6, 4, 2) and np.array_equal(a[:3], b)      # a prefix of the stream is the stream of fewer pixels
    assert a.min() >= -0.5 and a.max() < 0.5 and len(np.unique(a)) == a.size
    f = rh.sample_offsets(6, 4, seed=24, dtype=np.float32)
    assert np.allclose(f, a, atol=1e-7)
    assert not np.array_equal(rh.sample_offsets(6, 4, seed=25), a)


def test_ppm_writer_is_byte_exact(tmp_path):
    """Image.hs:60-75: 'P3\\nW H\\n255\\n', rows joined by '\\n', each pixel 'R G B' + two spaces, no trailing newline."""
    px = np.array([[[0, 10, 255], [1, 2, 3]], [[100, 99, 9], [7, 8, 200]]], dtype=np.uint8)
    p = str(tmp_path / "o.ppm")
    rh.writePPM(p, rh.Image(2, 2, px))
    assert open(p, "rb").read() == b"P3\n2 2\n255\n0 10 255  1 2 3  \n100 99 9  7 8 200  "


def test_assemble_bands_inverts_the_shard_mapping():
    H, W, bh = 22, 5, 3
    full = np.random.RandomState(0).randint(0, 255, size=(H, W, 3)).astype(np.uint8)
    for G in (1, 2, 3, 8):
        parts = []
        for g in range(G):
            rows = rh.shard_global_rows(H, g, G, bh)
            parts.append(np.stack([full[r] if r >= 0 else np.zeros((W, 3), np.uint8) for r in rows]))
        assert np.array_equal(rh.assemble_bands(parts, H, bh), full)


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = load_scene("cornellBox")
        o = oracle_for(sc)
        W, H, bh = 40, 30, 4
        rows = rh.shard_global_rows(H, rank, world, bh)
        mine = np.zeros((len(rows), W, 3), dtype=np.uint8)
        for lr, g in enumerate(rows):          # the oracle renders exactly this rank's rows
            if g >= 0:
                mine[lr] = o.render(sc.camera, W, H, sc.max_depth, rows=(g, g + 1, 1), want_ids=False)["rgb_u8"][g]
        t = torch.from_numpy(mine)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)              # the one exchange step of the path (NCCL on the GPU box)
        full = rh.assemble_bands([p.numpy() for p in parts], H, bh)
        ref = o.render(sc.camera, W, H, sc.max_depth, want_ids=False)["rgb_u8"]
        q.put((rank, bool(np.array_equal(full, ref))))
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_path_over_gloo():
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_cli_option_parsing(monkeypatch, capsys):
    """RayHs.hs:204-234: -oFILE glued, RequireOrder, default out.ppm."""
    calls = {}

    def fake_build(path, *a, **k):
        calls["scene"] = path
        return "job"

    monkeypatch.setattr(rh.host, "buildRendering", fake_build)
    monkeypatch.setattr(rh.host, "rayTrace", lambda job: "img")
    monkeypatch.setattr(rh.host, "writePPM", lambda path, img: calls.setdefault("out", path))
    assert rh.main(["-ocornell.ppm", "data/cornellBox.json"]) == 0
    assert calls == {"scene": "data/cornellBox.json", "out": "cornell.ppm"}
    calls.clear()
    assert rh.main(["scene.json"]) == 0 and calls["out"] == "out.ppm"
    assert rh.main([]) == 1
