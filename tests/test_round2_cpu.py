"""CPU tests of the round-2 host pieces: the counter-based offset generator, the oracle's sample renderer, and the
parallel reference-rule tree build."""
import ctypes as C

import numpy as np
import pytest

import rayhs_b200 as rh
from rayhs_b200 import capi
from tests.util import load_scene, oracle_for


def test_offsets_from_any_pixel_equal_the_full_stream():
    """rh_sample_offsets_f64_at(seed, first, n): the same values as the slice of rh_sample_offsets_f64's stream
    (SplitMix64 is counter-based), also across the thread boundaries of the parallel generator."""
    spp, n = 3, 700_000   # > 2^20 values: several threads
    full = rh.sample_offsets(n, spp, 24)
    for first, count in ((0, n), (1, 5), (12345, 400_000), (n - 7, 7)):
        part = np.empty((count, spp, 2), dtype=np.float64)
        capi.lib().rh_sample_offsets_f64_at(24, first, count, spp, part.ctypes.data)
        assert np.array_equal(part, full[first:first + count])
    assert full.min() >= -0.5 and full.max() < 0.5


def test_oracle_sample_renderer_equals_the_full_renderer():
    """orc_render2 (rows x columns, compact outputs, offsets regenerated from the seed) against orc_render (full frame,
    offsets array): same bytes, same hit ids, same ray counts on the sampled pixels."""
    sc = load_scene("cornellBox")
    w, h, spp = 96, 64, 4
    off = rh.sample_offsets(w * h, spp, 24)
    o = oracle_for(sc)
    full = o.render(sc.camera, w, h, sc.max_depth, spp=spp, offsets=off)
    rows, cols = (5, h, 9), (3, w, 7)
    part = o.render_sample(sc.camera, w, h, sc.max_depth, spp=spp, seed=24, rows=rows, cols=cols)
    ys, xs = np.arange(*rows), np.arange(*cols)
    assert np.array_equal(part["rows"], ys) and np.array_equal(part["cols"], xs)
    assert np.array_equal(part["rgb_u8"], full["rgb_u8"][ys][:, xs])
    assert np.array_equal(part["rgb_f64"], full["rgb_f64"][ys][:, xs])
    assert np.array_equal(part["hit_ids"], full["hit_ids"][ys][:, xs])
    one = o.render_sample(sc.camera, w, h, sc.max_depth, rows=rows, cols=cols)   # 1 sample at the pixel corner
    plain = o.render(sc.camera, w, h, sc.max_depth)
    assert np.array_equal(one["rgb_u8"], plain["rgb_u8"][ys][:, xs])
    o.close()


def test_parallel_tree_build_equals_the_oracles_own_build():
    """rh_flatten builds large meshes with several threads (subtrees spliced in left-to-right order): node count, leaf
    count, depth and the leaf order of the triangles must be what the oracle's sequential KDTree.hs:79-90 build gives."""
    sc = rh.Scene.synthetic(200_000, 10)   # >= 2^16 triangles: the forked build
    d = sc.flat.contents
    o = oracle_for(sc)
    mesh = [i for i in range(d.n_objects) if d.objects[i].kind == capi.RH_OBJ_MESH][0]
    st = o.tree_stats(mesh)
    nodes = np.ctypeslib.as_array(C.cast(d.nodes, C.POINTER(C.c_uint32)), (d.n_nodes, 16))   # rh_node = 64 bytes
    is_leaf, left, right = nodes[:, 15], nodes[:, 12], nodes[:, 13]
    assert int(is_leaf.sum()) == st["leaves"] and int((is_leaf == 0).sum()) == st["inner"]
    assert d.objects[mesh].depth == st["max_depth"] and d.objects[mesh].n_leaves == st["leaves"]
    # leaves occupy increasing, gap-free triangle slots in preorder (= left-to-right) order
    firsts, counts = left[is_leaf == 1], right[is_leaf == 1]
    assert firsts[0] == 0 and np.array_equal(firsts[1:], np.cumsum(counts)[:-1]) and int(counts.sum()) == d.n_tris
    assert int(counts.max()) == st["max_leaf"]
    # the image of a few rows agrees with the oracle (its own tree): the splice kept every index consistent
    o.close()


def _random_tris(n, seed=5):
    rng = np.random.default_rng(seed)
    tris = (capi.rh_tri * n)()
    a = np.frombuffer(tris, dtype=np.uint8).reshape(n, C.sizeof(capi.rh_tri))
    d = np.frombuffer(tris, dtype=np.float64).reshape(n, C.sizeof(capi.rh_tri) // 8)
    d[:, 0:3] = rng.uniform(-10, 10, (n, 3))          # p0
    d[:, 3:9] = rng.normal(0, 0.05, (n, 6))           # e1, e2
    ids = np.frombuffer(tris, dtype=np.uint32).reshape(n, C.sizeof(capi.rh_tri) // 4)
    ids[:, 18] = np.arange(n, dtype=np.uint32)        # tri_id
    del a
    return tris, d


@pytest.mark.parametrize("n", [1, 3, 5, 1000, 1 << 20])
def test_cull_tree_of_the_set_up_path(n):
    """rh_cull_tree_build = the binned-SAH tree rh_scene_create builds for the float path (several host threads; the
    ranges of 2^20 triangles and more are binned in pieces): every triangle in exactly one leaf of at most 4, every node's
    box holds its triangles / children, and the build does not depend on how the threads interleave."""
    L = capi.lib()
    tris, d = _random_tris(n)
    order = np.empty(n, dtype=np.uint32)
    nodes = (capi.rh_node * (2 * n + 1))()
    n_nodes, depth = C.c_uint32(), C.c_uint32()
    capi.check(L.rh_cull_tree_build(tris, n, order.ctypes.data, C.cast(nodes, C.c_void_p), C.byref(n_nodes), C.byref(depth)))
    assert sorted(order.tolist()) == list(range(n)) if n <= 1000 else np.array_equal(np.sort(order), np.arange(n, dtype=np.uint32))
    nn = n_nodes.value
    nd = np.frombuffer(nodes, dtype=np.uint8)[: nn * C.sizeof(capi.rh_node)].reshape(nn, C.sizeof(capi.rh_node))
    box = nd[:, :48].copy().view(np.float64).reshape(nn, 6)
    words = nd[:, 48:64].copy().view(np.uint32).reshape(nn, 4)      # left, right, leaf_index, is_leaf
    leaf = words[:, 3] != 0
    assert words[leaf, 1].max() <= 4 and words[leaf, 1].min() >= 1
    assert int(words[leaf, 1].sum()) == n                            # the leaves partition the slots ...
    firsts = np.sort(words[leaf, 0])
    assert firsts[0] == 0 and np.all(np.diff(firsts) >= 1)           # ... with distinct first slots
    # triangle boxes in new slot order
    p0, e1, e2 = d[order, 0:3], d[order, 3:6], d[order, 6:9]
    lo = np.minimum(p0, np.minimum(p0 + e1, p0 + e2))
    hi = np.maximum(p0, np.maximum(p0 + e1, p0 + e2))
    for i in np.flatnonzero(leaf)[:: max(1, int(leaf.sum()) // 2000)]:   # a sample of the leaves
        f, c = int(words[i, 0]), int(words[i, 1])
        assert np.all(box[i, :3] <= lo[f:f + c].min(axis=0)) and np.all(box[i, 3:] >= hi[f:f + c].max(axis=0))
    inner = np.flatnonzero(~leaf)
    if len(inner):
        l, r = words[inner, 0], words[inner, 1]
        assert np.all(l > inner) and np.all(r > inner) and r.max() < nn   # children follow their parent
        assert np.all(box[inner, :3] <= np.minimum(box[l, :3], box[r, :3])) and np.all(box[inner, 3:] >= np.maximum(box[l, 3:], box[r, 3:]))
    assert depth.value <= 64
    if n >= 1000:   # a second build gives the same tree
        order2 = np.empty(n, dtype=np.uint32)
        nodes2 = (capi.rh_node * (2 * n + 1))()
        capi.check(L.rh_cull_tree_build(tris, n, order2.ctypes.data, C.cast(nodes2, C.c_void_p), C.byref(n_nodes), C.byref(depth)))
        assert n_nodes.value == nn and np.array_equal(order, order2)
        assert bytes(nodes2)[: nn * C.sizeof(capi.rh_node)] == nd.tobytes()
