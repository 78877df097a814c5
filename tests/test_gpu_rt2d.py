"""RealTimeCorrelativeScanMatcher2D on the device: grid-resident handle, batched
matches, the three kernel forms (TMA-staged box, staged CROP of a larger grid, global
gathers) and the public ScoreCandidates — all compared with the oracle for equality
(score bits, pose doubles).  Reference: real_time_correlative_scan_matcher_2d.cc:117-176,
real_time_correlative_scan_matcher_2d_test.cc:125-160."""
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


def _ogrid(oracle, g):
    return oracle.Grid2D(g.cells, g.resolution, g.max_x, g.max_y)


def _jobs(grid, occ, n, beams, max_range, seed, margin=15):
    rng = np.random.RandomState(seed)
    scans, inits = [], []
    for i in range(n):
        pose = synthetic.random_free_pose(occ, grid, rng, margin_cells=margin)
        scans.append(synthetic.cast_scan(occ, grid, pose, beams=beams, seed=seed * 1000 + i,
                                         max_range=max_range))
        inits.append(pose + rng.uniform(-1, 1, 3) * [0.05, 0.05, math.radians(3)])
    return scans, inits


def _check_batch(oracle, sm, grid, scans, inits, lin, ang, w_t=0.1, w_r=0.1):
    opts = sm.RealTimeCorrelativeScanMatcherOptions(lin, ang, w_t, w_r)
    rt = sm.RealTimeCorrelativeScanMatcher2D(opts)
    dg = sm.RealTimeGrid2D(grid)
    scores, poses, st = rt.MatchBatch(inits, scans, dg)
    og = _ogrid(oracle, grid)
    total = 0
    for k in range(len(scans)):
        want = oracle.rt2d_match(og, scans[k], inits[k], lin, ang, w_t, w_r)
        assert np.float32(scores[k]) == np.float32(want["score"]), k
        np.testing.assert_array_equal(poses[k], want["pose"])
        r = rt.last_results[k]
        assert (r.best_scan_index, r.best_x_offset, r.best_y_offset) == (
            want["best_scan_index"], want["best_x_offset"], want["best_y_offset"])
        assert r.candidates_scored == want["candidates_scored"]
        total += want["candidates_scored"]
    assert st["candidates_scored"] == total
    assert st["host_syncs"] == 1
    dg.close()
    return rt


def test_rt_batch_config1_shape(oracle, sm):
    """BASELINE config 1 shape, batched: 1081 beams, 200x200 @5 cm, +-0.1 m / +-7 deg.
    The whole grid is one TMA box (200 x 200 uint16 = 80 KB of shared memory)."""
    big, occ = synthetic.make_grid2d(7, 1000)
    grid, occ2 = synthetic.crop_grid(big, occ, 400, 400, 200, 200)
    scans, inits = _jobs(grid, occ2, 24, 1081, 30.0, 3)
    rt = _check_batch(oracle, sm, grid, scans, inits, 0.1, math.radians(7.0))
    # the single-call form (grid passed per call) gives the same answers
    score, est = rt.Match(inits[0], scans[0], grid)
    want = oracle.rt2d_match(_ogrid(oracle, grid), scans[0], inits[0], 0.1, math.radians(7.0),
                             0.1, 0.1)
    assert np.float32(score) == np.float32(want["score"])
    np.testing.assert_array_equal(est, want["pose"])


def test_rt_batch_default_options(oracle, sm):
    """trajectory_builder_2d.lua:37-43 defaults: 0.1 m / 20 deg, weights 1e-1 / 1e-1."""
    grid, occ, _, _ = worlds.small_world(410, size_cells=160, beams=181, max_range=5.0)
    scans, inits = _jobs(grid, occ, 12, 361, 6.0, 5, margin=10)
    _check_batch(oracle, sm, grid, scans, inits, 0.1, math.radians(20.0))


def test_rt_wide_window_several_candidates_per_lane(oracle, sm):
    """+-0.3 m => 13 x 13 = 169 offsets per rotated scan: more than one candidate per lane
    and a second pass of the lane accumulators."""
    grid, occ, _, _ = worlds.small_world(411, size_cells=120, beams=181, max_range=4.0)
    scans, inits = _jobs(grid, occ, 6, 241, 4.0, 7, margin=10)
    _check_batch(oracle, sm, grid, scans, inits, 0.3, math.radians(4.0), 0.3, 0.7)


def test_rt_large_grid_crop_and_global_forms(oracle, sm):
    """A 1000 x 1000 grid does not fit the staged box.  Short-range scans reach only a
    crop of it (TMA box placed at the crop origin); long-range scans fall back to global
    gathers.  Both must equal the oracle."""
    grid, occ = synthetic.make_grid2d(2, 1000)
    # reach = 2 * (ceil(3.0 / 0.05) + 2 + 3) + 1 = 131 cells < the box: staged crop
    scans, inits = _jobs(grid, occ, 8, 361, 3.0, 11, margin=80)
    _check_batch(oracle, sm, grid, scans, inits, 0.1, math.radians(7.0))
    # 30 m scans reach the whole grid: global-gather form
    scans, inits = _jobs(grid, occ, 4, 1081, 30.0, 12, margin=40)
    _check_batch(oracle, sm, grid, scans, inits, 0.1, math.radians(7.0))


def test_rt_scan_outside_the_grid(oracle, sm):
    """Every point outside the limits: all candidates score kMinProbability * weight and
    the first maximum in generation order wins (the zero-offset, smallest-rotation ones
    carry the largest weight)."""
    grid, occ, pose, scan = worlds.small_world(412, size_cells=100, beams=91, max_range=3.0)
    init = np.array([grid.max_x + 40.0, grid.max_y + 40.0, 0.3])
    _check_batch(oracle, sm, grid, [scan], [init], 0.1, math.radians(5.0))


def test_rt_grid_update(oracle, sm):
    """csm_rt_grid2d_update: the device copy follows the host grid after a scan insertion."""
    grid, occ, _, _ = worlds.small_world(413, size_cells=140, beams=181, max_range=5.0)
    scans, inits = _jobs(grid, occ, 3, 361, 6.0, 9, margin=10)
    opts = sm.RealTimeCorrelativeScanMatcherOptions(0.1, math.radians(7.0), 0.1, 0.1)
    rt = sm.RealTimeCorrelativeScanMatcher2D(opts)
    dg = sm.RealTimeGrid2D(grid)
    rng = np.random.RandomState(0)
    cells2 = grid.cells.copy()
    mask = rng.uniform(size=cells2.shape) < 0.3
    cells2[mask] = rng.randint(1, 32768, size=int(mask.sum())).astype(np.uint16)
    dg.update(cells2)
    g2 = synthetic.GridSpec(cells2, grid.resolution, grid.max_x, grid.max_y)
    scores, poses, _ = rt.MatchBatch(inits, scans, dg)
    for k in range(len(scans)):
        want = oracle.rt2d_match(_ogrid(oracle, g2), scans[k], inits[k], 0.1, math.radians(7.0),
                                 0.1, 0.1)
        assert np.float32(scores[k]) == np.float32(want["score"])
        np.testing.assert_array_equal(poses[k], want["pose"])
    dg.close()


def test_rt_public_score_candidates(oracle, sm):
    """The reference's public ScoreCandidates on caller-supplied discrete scans and
    candidates (real_time_correlative_scan_matcher_2d.h:75)."""
    grid, occ, pose, scan = worlds.small_world(414, size_cells=150, beams=241, max_range=5.0)
    og = _ogrid(oracle, grid)
    lin, ang = 0.1, math.radians(7.0)
    fe = oracle.frontend2d(og, scan, pose, False, lin, ang, rt_mode=True)
    sp = oracle.search_params(lin, ang, scan, grid.resolution)
    S = fe["num_scans"]
    rng = np.random.RandomState(4)
    cand = np.stack([rng.randint(0, S, 300), rng.randint(-2, 3, 300), rng.randint(-2, 3, 300)],
                    axis=1).astype(np.int32)
    num_ang = (S - 1) // 2
    want = oracle.rt2d_score_candidates(og, fe["discrete_scans"], sp["num_linear_perturbations"],
                                        num_ang, fe["step"], 0.1, 0.1, cand)
    rt = sm.RealTimeCorrelativeScanMatcher2D(sm.RealTimeCorrelativeScanMatcherOptions(lin, ang, 0.1, 0.1))
    got = rt.ScoreCandidates(grid, fe["discrete_scans"], num_ang, fe["step"], cand)
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
