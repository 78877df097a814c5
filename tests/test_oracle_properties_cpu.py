"""Size-independent properties of the path, checked on the CPU oracle: the invariants the
branch-and-bound (and therefore the device engine's pruning) relies on.  CPU only."""
import numpy as np
import pytest

from tests import worlds


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_precomputed_score_bounds_every_descendant(oracle, seed):
    """fast_correlative_scan_matcher_2d.cc:335-378 prunes with `score <= min_score`: that
    is only sound if a node's score at level h is >= the score of each of its four
    children at level h-1 (max over a 2^h window >= max over each 2^(h-1) sub-window)."""
    grid, occ, pose, scan = worlds.small_world(seed)
    og = oracle.Grid2D(grid.cells, grid.resolution, grid.max_x, grid.max_y)
    depth = 5
    om = oracle.FastCorrelativeScanMatcher2D(og, 2.0, 0.4, depth)
    fe = oracle.frontend2d(og, scan, pose, full=False, lin=2.0, ang=0.4)
    ds = fe["discrete_scans"]
    rng = np.random.RandomState(seed)
    for h in range(depth - 1, 0, -1):
        half = 1 << (h - 1)
        cand = np.stack([rng.randint(0, len(ds), 64), rng.randint(-40, 40, 64) * 1,
                         rng.randint(-40, 40, 64)], 1).astype(np.int32)
        parent, _ = om.score_candidates(h, ds, cand)
        for dx in (0, half):
            for dy in (0, half):
                child, _ = om.score_candidates(h - 1, ds, cand + np.array([0, dx, dy], np.int32))
                assert np.all(parent >= child)


@pytest.mark.parametrize("seed", [4, 5])
def test_levels_are_monotone_and_level0_is_the_grid(oracle, seed):
    """PrecomputationGridStack2D (:171-186): level h+1 dominates level h cell-wise at the
    same grid position (its window contains level h's window)."""
    grid, occ, pose, scan = worlds.small_world(seed, size_cells=120)
    og = oracle.Grid2D(grid.cells, grid.resolution, grid.max_x, grid.max_y)
    om = oracle.FastCorrelativeScanMatcher2D(og, 1.0, 0.2, 5)
    prev = om.level(0)
    assert prev.shape == grid.cells.shape
    for h in range(1, 5):
        cur = om.level(h)
        w, wp = (1 << h) - 1, (1 << (h - 1)) - 1
        ny, nx = grid.cells.shape
        # same window START (grid cell (x, y)): wide index + offset w / wp
        assert np.all(cur[w:w + ny, w:w + nx] >= prev[wp:wp + ny, wp:wp + nx])
        prev = cur


def test_match_result_is_idempotent_and_window_monotone(oracle):
    """Matching again from the found pose with a zero window returns the same score at the
    same discretised pose; a wider window can only find an equal or better score."""
    grid, occ, pose, scan = worlds.small_world(7)
    og = oracle.Grid2D(grid.cells, grid.resolution, grid.max_x, grid.max_y)
    narrow = oracle.FastCorrelativeScanMatcher2D(og, 0.5, 0.2, 4).match(pose, scan, 0.1)
    wide = oracle.FastCorrelativeScanMatcher2D(og, 1.5, 0.4, 4).match(pose, scan, 0.1)
    assert narrow["found"] and wide["found"]
    assert wide["score"] >= narrow["score"]
    again = oracle.FastCorrelativeScanMatcher2D(og, 0.0, 0.0, 1).match(narrow["pose"], scan, 0.0)
    assert again["found"] and again["score"] == narrow["score"]
    np.testing.assert_allclose(again["pose"], narrow["pose"], atol=1e-12)
