"""Ingest on the device: stacks built from serialized protos / a pbstream equal stacks
built from the raw cells; an updated stack equals a freshly built one; a 3D matcher built
from proto::HybridGrid messages gives the oracle's match."""
import math

import numpy as np
import pytest

from benchmarks import synthetic
from tests import protowire as pw
from tests import worlds

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sm():
    from cartographer_b200 import scan_matching
    return scan_matching


def test_stack_from_proto_and_pbstream(sm, tmp_path):
    opts = sm.FastCorrelativeScanMatcherOptions2D(2.0, 0.5, 5)
    grids = [worlds.small_world(700 + k, size_cells=120 + 10 * k, beams=91, max_range=5.0)[0]
             for k in range(3)]
    blobs = [pw.grid2d(g.cells, g.resolution, g.max_x, g.max_y, np.float32(g.min_cost),
                       np.float32(g.max_cost)) for g in grids]
    path = tmp_path / "map.pbstream"
    pw.write_pbstream(str(path), [(0, 5 + k, b) for k, b in enumerate(blobs)])
    loaded, ids = sm.load_pbstream_matchers2d(str(path), opts)
    assert ids == [(0, 5), (0, 6), (0, 7)]
    for g, blob, from_file in zip(grids, blobs, loaded):
        direct = sm.FastCorrelativeScanMatcher2D(g, opts)
        from_proto = sm.FastCorrelativeScanMatcher2D.from_proto(blob, opts)
        for level in range(5):
            want = direct.precomputation_grid(level)
            np.testing.assert_array_equal(from_proto.precomputation_grid(level), want)
            np.testing.assert_array_equal(from_file.precomputation_grid(level), want)
        for m in (direct, from_proto, from_file):
            m.close()


def test_stack_update_equals_rebuild(oracle, sm):
    grid, occ, pose, scan = worlds.small_world(710, size_cells=150, beams=181, max_range=6.0)
    opts = sm.FastCorrelativeScanMatcherOptions2D(1.5, 0.4, 5)
    m = sm.FastCorrelativeScanMatcher2D(grid, opts)
    rng = np.random.RandomState(3)
    cells2 = grid.cells.copy()
    mask = rng.uniform(size=cells2.shape) < 0.2
    cells2[mask] = rng.randint(1, 32768, size=int(mask.sum())).astype(np.uint16)
    m.update(cells2)
    g2 = synthetic.GridSpec(cells2, grid.resolution, grid.max_x, grid.max_y)
    fresh = sm.FastCorrelativeScanMatcher2D(g2, opts)
    for level in range(5):
        np.testing.assert_array_equal(m.precomputation_grid(level), fresh.precomputation_grid(level))
    init = pose + np.array([0.3, -0.2, 0.05])
    a = m.Match(init, scan, 0.3)
    b = fresh.Match(init, scan, 0.3)
    og = oracle.Grid2D(cells2, grid.resolution, grid.max_x, grid.max_y)
    want = oracle.FastCorrelativeScanMatcher2D(og, 1.5, 0.4, 5).match(init, scan, 0.3)
    assert bool(a[0]) == bool(b[0]) == want["found"]
    if want["found"]:
        assert a[1] == b[1] == want["score"]
        np.testing.assert_array_equal(a[2], want["pose"])
        np.testing.assert_array_equal(b[2], want["pose"])
    m.close()
    fresh.close()


def test_matcher3d_from_proto(oracle, sm):
    hi, lo, hist, world = synthetic.make_submap3d(46, 30.0, 8, 256, 15.0)
    rng = np.random.RandomState(5)
    node = synthetic.make_node3d(world, rng, 8, 256, 15.0, seed=3)
    o3 = sm.FastCorrelativeScanMatcherOptions3D(min_rotational_score=0.3)
    direct = sm.FastCorrelativeScanMatcher3D(hi, lo, hist, o3)
    from_proto = sm.FastCorrelativeScanMatcher3D.from_proto(
        pw.hybrid_grid(hi.resolution, hi.indices, hi.values),
        pw.hybrid_grid(lo.resolution, lo.indices, lo.values), hist, o3)
    init = node["pose"].copy()
    init[:3] += [0.5, -0.4, 0.1]
    ident = [0, 0, 0, 1, 0, 0, 0]
    data = sm.TrajectoryNodeData3D(node["cloud"], node["low"], node["hist"])
    a = direct.Match(init, ident, data, 0.4)
    b = from_proto.Match(init, ident, data, 0.4)
    assert (a is None) == (b is None)
    if a is not None:
        assert a["score"] == b["score"]
        np.testing.assert_array_equal(a["pose_estimate"], b["pose_estimate"])
    direct.close()
    from_proto.close()
