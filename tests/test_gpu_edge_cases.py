"""Edge cases of the 2D fast matcher on the device vs the oracle: degenerate depths,
tiny grids and clouds, scans outside the map, windows larger than the map, min_score
boundaries, update-marker bits in the cells."""
import math

import numpy as np
import pytest

from benchmarks import synthetic
from tests import worlds

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sm():
    from cartographer_b200 import scan_matching
    return scan_matching


def _run(oracle, sm, grid, cloud, init, lin, ang, depth, min_score, full=False):
    m = sm.FastCorrelativeScanMatcher2D(grid, sm.FastCorrelativeScanMatcherOptions2D(lin, ang, depth))
    om = oracle.FastCorrelativeScanMatcher2D(
        oracle.Grid2D(grid.cells, grid.resolution, grid.max_x, grid.max_y), lin, ang, depth)
    if full:
        want = om.match_full_submap(cloud, min_score)
        found, score, est = m.MatchFullSubmap(cloud, min_score)
    else:
        want = om.match(init, cloud, min_score)
        found, score, est = m.Match(init, cloud, min_score)
    st = m.last_stats
    m.close()
    assert bool(found) == want["found"], (found, want)
    if want["found"]:
        assert np.float32(score) == want["score"]
        np.testing.assert_array_equal(est, want["pose"])
        assert (st["best_scan_index"], st["best_x_offset"], st["best_y_offset"]) == (
            want["best_scan_index"], want["best_x_offset"], want["best_y_offset"])
    return want


@pytest.mark.parametrize("depth", [1, 2, 3])
def test_shallow_depths(oracle, sm, depth):
    grid, occ, pose, scan = worlds.small_world(31, size_cells=120, beams=97, max_range=4.0)
    init = pose + np.array([0.1, -0.15, 0.03])
    w = _run(oracle, sm, grid, scan, init, 0.4, 0.1, depth, 0.2)
    assert w["found"]


@pytest.mark.parametrize("n", [1, 2, 31, 33, 257])
def test_ragged_point_counts(oracle, sm, n):
    grid, occ, pose, scan = worlds.small_world(32, size_cells=100, beams=300, max_range=4.0)
    _run(oracle, sm, grid, scan[:n], pose + np.array([0.05, 0.05, 0.01]), 0.5, 0.2, 4, 0.15)


def test_scan_completely_outside_the_map(oracle, sm):
    grid, occ, pose, scan = worlds.small_world(33, size_cells=100, beams=61, max_range=3.0)
    far = np.array([200.0, -300.0, 0.3])
    w = _run(oracle, sm, grid, scan, far, 1.0, 0.2, 4, 0.11)
    assert not w["found"]
    # min_score below the all-unknown score (0.1): everything "matches" with score 0.1... the
    # strict > keeps parity on which candidate wins among a window of exact ties
    _run(oracle, sm, grid, scan, far, 0.3, 0.05, 3, 0.05)


def test_window_larger_than_map_and_tiny_grids(oracle, sm):
    rng = np.random.RandomState(3)
    for (nx, ny) in [(1, 1), (2, 5), (7, 3), (16, 16)]:
        cells = rng.randint(1, 32768, size=(ny, nx)).astype(np.uint16)
        cells[rng.uniform(size=cells.shape) < 0.3] = 0
        cells[rng.uniform(size=cells.shape) < 0.2] |= 0x8000
        grid = synthetic.GridSpec(cells, 0.05, 0.3, 0.4)
        cloud = rng.uniform(-0.3, 0.3, size=(9, 3)).astype(np.float32)
        cloud[:, 2] = 0
        _run(oracle, sm, grid, cloud, (0.1, 0.2, 0.1), 1.0, 0.5, 4, 0.12)
        _run(oracle, sm, grid, cloud, (0.0, 0.0, 0.0), 0.0, 0.0, 3, 0.12, full=True)


def test_zero_windows(oracle, sm):
    grid, occ, pose, scan = worlds.small_world(34, size_cells=100, beams=61, max_range=3.0)
    _run(oracle, sm, grid, scan, pose, 0.0, 0.0, 3, 0.2)      # a single candidate
    _run(oracle, sm, grid, scan, pose, 0.0, 0.3, 3, 0.2)      # rotations only
    _run(oracle, sm, grid, scan, pose, 0.6, 0.0, 5, 0.2)      # translations only


def test_min_score_boundary_is_strict(oracle, sm):
    grid, occ, pose, scan = worlds.small_world(35, size_cells=100, beams=121, max_range=3.0)
    w = _run(oracle, sm, grid, scan, pose, 0.3, 0.1, 4, 0.2)
    assert w["found"]
    s = float(w["score"])
    # exactly the best score as threshold: `score > min_score` must fail
    w2 = _run(oracle, sm, grid, scan, pose, 0.3, 0.1, 4, s)
    assert not w2["found"]
    w3 = _run(oracle, sm, grid, scan, pose, 0.3, 0.1, 4, float(np.nextafter(np.float32(s), np.float32(0))))
    assert w3["found"]


def test_custom_cost_bounds(oracle, sm):
    """Grid2D with non-default min/max correspondence cost (the TSDF-style bounds the
    precomputation grid is also built from, fast...2d.cc:97-98)."""
    rng = np.random.RandomState(5)
    cells = rng.randint(0, 32768, size=(64, 80)).astype(np.uint16)
    grid = synthetic.GridSpec(cells, 0.05, 2.0, 1.6)
    grid.min_cost, grid.max_cost = 0.0, 0.5
    og = oracle.Grid2D(cells, 0.05, 2.0, 1.6, 0.0, 0.5)
    m = sm.FastCorrelativeScanMatcher2D(grid, sm.FastCorrelativeScanMatcherOptions2D(0.5, 0.2, 4))
    for level in range(4):
        want = oracle.precompute_grid2d(cells, 0.0, 0.5, 1 << level)
        np.testing.assert_array_equal(m.precomputation_grid(level), want)
    cloud = rng.uniform(-1.0, 1.0, size=(50, 3)).astype(np.float32)
    om = oracle.FastCorrelativeScanMatcher2D(og, 0.5, 0.2, 4)
    want = om.match((0.3, -0.2, 0.1), cloud, 0.55)
    found, score, est = m.Match((0.3, -0.2, 0.1), cloud, 0.55)
    assert bool(found) == want["found"]
    if found:
        assert np.float32(score) == want["score"]
        np.testing.assert_array_equal(est, want["pose"])
    m.close()


def test_invalid_arguments_on_device(sm):
    from cartographer_b200._lib import CsmError
    grid = synthetic.GridSpec(np.zeros((8, 8), np.uint16), 0.05, 0.2, 0.2)
    m = sm.FastCorrelativeScanMatcher2D(grid, sm.FastCorrelativeScanMatcherOptions2D(0.5, 0.2, 3))
    with pytest.raises(CsmError):
        m.Match((0, 0, 0), np.zeros((0, 3), np.float32), 0.1)      # empty cloud
    with pytest.raises(ValueError):
        m.Match((0, 0, 0), np.zeros((4, 2), np.float32), 0.1)      # not N x 3
    m.close()
