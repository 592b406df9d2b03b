"""The device walkers, compiled for the host from the same source (qengine_b200/csrc/walkers.cuh, C entry points
qrad_host_tree_* of include/qrad_cuda.h), against the reference's golden answers and against each other:

  * TestLine / PointInLeaf of the product reproduce the reference's (golden rays and leafs);
  * the cell proofs are SOUND: whenever segment_through_cell / shaft_blocked call a ray (a bundle of rays) blocked, the
    reference's TestLine_r says blocked -- on every ray of the goldens and on thousands of patch pairs per map -- and
    they are not vacuous (they settle a large share of the blocked rays);
  * starting a walk at common_start_node() changes no result.
The kernels use these very functions; the GPU parity tests cover what the kernels make of them.
"""
from __future__ import annotations

import numpy as np
import pytest

import golden_util as G
from qengine_b200 import qrad as Q

CASES = ["samplec", "grid2", "grid3_chop16", "hall", "angled", "styles"]


def setup(case, tmp):
    meta, flags = G.load_case(case)
    bsp = G.stage_case(case, tmp)
    sc = Q.Scene(bsp).prepare(G.params_from_flags(flags))
    return sc, Q.HostTree(sc)


@pytest.mark.parametrize("case", CASES)
def test_host_testline_and_point_in_leaf_match_the_reference(case, tmp_path):
    sc, tree = setup(case, tmp_path)
    rays = G.golden_rays(case)
    res, blk = tree.test_lines(rays["start"], rays["stop"])
    assert np.array_equal(res != 0, rays["result"] != 0)
    assert ((blk >= 0) == (res != 0)).all()
    pts, leafs = G.golden_rayleafs(case)
    assert np.array_equal(tree.point_in_leaf(pts), leafs)


@pytest.mark.parametrize("case", CASES)
def test_cell_proofs_never_block_a_clear_ray(case, tmp_path):
    sc, tree = setup(case, tmp_path)
    rays = G.golden_rays(case)
    # golden rays: the reference's own answers
    leaf = tree.prove_segments(rays["start"], rays["stop"])
    proven = leaf >= 0
    assert (rays["result"][proven] != 0).all(), "a ray the reference calls clear was proven blocked"
    blocked = rays["result"] != 0
    if case != "angled":      # maps of axial brushes: most blockers have a box cell
        assert proven.sum() >= 0.5 * blocked.sum(), (int(proven.sum()), int(blocked.sum()))
    # many more segments between patch origins, checked against the host TestLine (itself pinned above)
    org = sc.patches()["origin"]
    rng = np.random.default_rng(3)
    a = rng.integers(0, len(org), 20000)
    b = rng.integers(0, len(org), 20000)
    res, blk = tree.test_lines(org[a], org[b])
    leaf = tree.prove_segments(org[a], org[b])
    assert (res[leaf >= 0] != 0).all()
    # the library's margin is what the kernels use; a proof with a smaller margin must still be sound down to ON_EPSILON
    leaf2 = tree.prove_segments(org[a], org[b], margin=0.12)
    assert (res[leaf2 >= 0] != 0).all()
    assert (leaf2 >= 0).sum() >= (leaf >= 0).sum()


@pytest.mark.parametrize("case", ["grid2", "grid3_chop16", "hall", "angled"])
def test_box_proofs_and_start_nodes(case, tmp_path):
    sc, tree = setup(case, tmp_path)
    org = sc.patches()["origin"]
    n = len(org)
    rng = np.random.default_rng(5)
    proven_boxes = deeper = 0
    for _ in range(400):
        k = int(rng.choice([4, 8, 16, 32]))
        i, j = (int(x) for x in rng.integers(0, n - k, 2))
        A, B = org[i:i + k], org[j:j + k]
        amn, amx, bmn, bmx = A.min(0), A.max(0), B.min(0), B.max(0)
        s = np.repeat(A, k, axis=0)
        e = np.tile(B, (k, 1))
        res, _ = tree.test_lines(s, e)
        if tree.prove_boxes(amn, amx, bmn, bmx) >= 0:
            proven_boxes += 1
            assert (res != 0).all(), "a box proof covers a clear ray"
        node = tree.common_start_node(amn, amx, bmn, bmx)
        assert 0 <= node < tree.numnodes
        deeper += node != 0
        res2, _ = tree.test_lines(s, e, start_nodes=np.full(len(s), node, np.int32))
        assert np.array_equal(res, res2), "starting at the common node changed a result"
    assert proven_boxes > 0 and deeper > 0
