"""CPU tests of the round-2 host pieces: the counter-based offset generator, the oracle's sample renderer, and the
parallel reference-rule tree build."""
import ctypes as C

import numpy as np

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
