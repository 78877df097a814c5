"""GPU parity tests of the 2D path: CUDA engine (through the C ABI) vs the CPU
oracle on the same seeded inputs.  Bit-exact for integers (cell indices, bounds,
uint8 grid values, sums, winning candidate) and for float scores (the device
evaluates the reference's float expressions with round-to-nearest intrinsics in
the same order; tolerance asserted: 0, far inside north_star's 1e-5).
"""
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


def _check_match(o_res, found, score, pose, stats, exact_pose=True):
    assert bool(found) == o_res["found"]
    if not o_res["found"]:
        return
    assert np.float32(score) == np.float32(o_res["score"]), (score, o_res["score"])
    assert abs(float(score) - float(o_res["score"])) <= 1e-5
    if exact_pose:
        assert (stats["best_scan_index"], stats["best_x_offset"], stats["best_y_offset"]) == (
            o_res["best_scan_index"], o_res["best_x_offset"], o_res["best_y_offset"])
        np.testing.assert_array_equal(np.asarray(pose), o_res["pose"])


# ---- K1: precomputation grid stack (fast...2d.cc:91-186) ---------------------
@pytest.mark.parametrize("nx,ny,depth,seed", [(250, 250, 4, 42), (4, 4, 4, 42), (37, 61, 7, 3),
                                              (1, 1, 3, 0), (1000, 1000, 7, 0)])
def test_stack_levels_bit_exact(oracle, sm, nx, ny, depth, seed):
    rng = np.random.RandomState(seed)
    if nx == 1000:
        grid, _ = synthetic.make_grid2d(seed, 1000)
    else:
        cells = rng.randint(0, 32768, size=(ny, nx)).astype(np.uint16)
        cells[rng.uniform(size=cells.shape) < 0.2] = 0
        cells[rng.uniform(size=cells.shape) < 0.1] |= 0x8000  # update marker must be masked
        grid = synthetic.GridSpec(cells, 0.05, 1.0, 2.0)
    m = sm.FastCorrelativeScanMatcher2D(grid, sm.FastCorrelativeScanMatcherOptions2D(1., 0.5, depth))
    for level in range(depth):
        want = oracle.precompute_grid2d(grid.cells, grid.min_cost, grid.max_cost, 1 << level)
        got = m.precomputation_grid(level)
        assert got.shape == want.shape
        np.testing.assert_array_equal(got, want)
    m.close()


# ---- K2: rotate + discretise + ShrinkToFit ----------------------------------
@pytest.mark.parametrize("seed,full", [(1, False), (2, False), (3, True), (4, True)])
def test_discretize_bit_exact(oracle, sm, seed, full):
    grid, occ, pose, scan = worlds.small_world(seed)
    rng = np.random.RandomState(seed)
    init = pose + np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-0.3, 0.3)])
    opts = sm.FastCorrelativeScanMatcherOptions2D(2.0, 0.4, 5)
    m = sm.FastCorrelativeScanMatcher2D(grid, opts)
    ds, bounds = m.discretize(scan, init, full_submap=full)
    want = oracle.frontend2d(_ogrid(oracle, grid), scan, init, full=full, lin=2.0, ang=0.4)
    assert ds.shape == want["discrete_scans"].shape
    np.testing.assert_array_equal(ds, want["discrete_scans"])
    np.testing.assert_array_equal(bounds, want["bounds"])
    m.close()


def test_discretize_reference_golden_cells(sm):
    """correlative_scan_matcher_test.cc:72-96: the 7 exact cells, on the device."""
    cloud = np.array([[0.025, 0.175, 0], [-0.025, 0.175, 0], [-0.075, 0.175, 0],
                      [-0.125, 0.175, 0], [-0.125, 0.125, 0], [-0.125, 0.075, 0],
                      [-0.125, 0.025, 0]], np.float32)
    grid = synthetic.GridSpec(np.zeros((6, 6), np.uint16), 0.05, 0.05, 0.25)
    m = sm.FastCorrelativeScanMatcher2D(grid, sm.FastCorrelativeScanMatcherOptions2D(0.0, 0.0, 1))
    ds, _ = m.discretize(cloud, (0.0, 0.0, 0.0))
    mid = ds.shape[0] // 2
    assert ds[mid].tolist() == [[1, 0], [1, 1], [1, 2], [1, 3], [2, 3], [3, 3], [4, 3]]
    m.close()


# ---- K3: ScoreCandidates (fast...2d.cc:314-333) ------------------------------
@pytest.mark.parametrize("seed", [5, 6])
def test_score_candidates_bit_exact(oracle, sm, seed):
    grid, occ, pose, scan = worlds.small_world(seed)
    depth = 6
    m = sm.FastCorrelativeScanMatcher2D(grid, sm.FastCorrelativeScanMatcherOptions2D(3.0, 0.5, depth))
    om = oracle.FastCorrelativeScanMatcher2D(_ogrid(oracle, grid), 3.0, 0.5, depth)
    fe = oracle.frontend2d(_ogrid(oracle, grid), scan, pose, lin=3.0, ang=0.5)
    ds = fe["discrete_scans"]
    rng = np.random.RandomState(seed)
    C = 4000
    cand = np.stack([rng.randint(0, ds.shape[0], C), rng.randint(-260, 260, C),
                     rng.randint(-260, 260, C)], axis=1).astype(np.int32)
    for level in range(depth):
        want_scores, want_sums = om.score_candidates(level, ds, cand)
        got_scores, got_sums = m.score_candidates(level, ds, cand)
        np.testing.assert_array_equal(got_sums, want_sums)
        np.testing.assert_array_equal(got_scores.view(np.uint32), want_scores.view(np.uint32))
    m.close()


# ---- full Match / MatchFullSubmap -------------------------------------------
@pytest.mark.parametrize("seed", range(8))
def test_match_local_window_parity(oracle, sm, seed):
    grid, occ, pose, scan = worlds.small_world(100 + seed)
    rng = np.random.RandomState(seed)
    init = pose + np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-0.3, 0.3)])
    depth = 3 + seed % 4
    lin, ang, min_score = 2.0, 0.5, 0.3
    m = sm.FastCorrelativeScanMatcher2D(grid, sm.FastCorrelativeScanMatcherOptions2D(lin, ang, depth))
    om = oracle.FastCorrelativeScanMatcher2D(_ogrid(oracle, grid), lin, ang, depth)
    want = om.match(init, scan, min_score)
    found, score, est = m.Match(init, scan, min_score)
    _check_match(want, found, score, est, m.last_stats)
    assert m.last_stats["num_scans"] == want["num_scans"]
    assert m.last_stats["lowest_resolution_candidates"] == want["lowest_resolution_candidates"]
    m.close()


@pytest.mark.parametrize("seed", range(4))
def test_match_full_submap_parity(oracle, sm, seed):
    grid, occ, pose, scan = worlds.small_world(200 + seed, size_cells=160, beams=121)
    depth = 5
    m = sm.FastCorrelativeScanMatcher2D(grid, sm.FastCorrelativeScanMatcherOptions2D(1.0, 0.3, depth))
    om = oracle.FastCorrelativeScanMatcher2D(_ogrid(oracle, grid), 1.0, 0.3, depth)
    want = om.match_full_submap(scan, 0.4)
    found, score, est = m.MatchFullSubmap(scan, 0.4)
    _check_match(want, found, score, est, m.last_stats)
    m.close()


def test_match_no_pose_above_min_score(oracle, sm):
    grid, occ, pose, scan = worlds.small_world(7)
    m = sm.FastCorrelativeScanMatcher2D(grid, sm.FastCorrelativeScanMatcherOptions2D(1.0, 0.3, 4))
    om = oracle.FastCorrelativeScanMatcher2D(_ogrid(oracle, grid), 1.0, 0.3, 4)
    want = om.match(pose, scan, 0.95)
    assert not want["found"]
    found, score, est = m.Match(pose, scan, 0.95)
    assert found is False and score is None and est is None
    m.close()


# Reference tests ported onto the device: sparse clouds on near-empty maps give
# MANY exactly tied candidates, so this also exercises the DFS-order tie rules.
FAST_CLOUD = np.array([[-2.5, 0.5, 0], [-2.0, 0.5, 0], [0.0, -0.5, 0], [0.5, -1.6, 0],
                       [2.5, 0.5, 0], [2.5, 1.7, 0]], np.float32)


def _transform(cloud, pose):
    c, s = math.cos(pose[2]), math.sin(pose[2])
    out = cloud.copy()
    out[:, 0] = c * cloud[:, 0] - s * cloud[:, 1] + pose[0]
    out[:, 1] = s * cloud[:, 0] + c * cloud[:, 1] + pose[1]
    return out.astype(np.float32)


def test_reference_correct_pose_test_on_device(oracle, sm):
    """fast_correlative_scan_matcher_2d_test.cc:144-192 on the engine, compared to
    the oracle bit for bit (ties included)."""
    rng = np.random.RandomState(42)
    tied = 0
    for _ in range(25):
        u = rng.uniform(-1.0, 1.0, 3)
        expected = (2.0 * u[0], 2.0 * u[1], 0.5 * u[2])
        og = worlds.insert_range_data(oracle, 200, 200, 0.05, 5.0, 5.0, expected[:2],
                                      _transform(FAST_CLOUD, expected))
        grid = synthetic.GridSpec(og.cells, 0.05, 5.0, 5.0)
        m = sm.FastCorrelativeScanMatcher2D(grid, sm.FastCorrelativeScanMatcherOptions2D(3.0, 1.0, 3))
        om = oracle.FastCorrelativeScanMatcher2D(og, 3.0, 1.0, 3)
        want = om.match((0.0, 0.0, 0.0), FAST_CLOUD, 0.1)
        found, score, est = m.Match((0.0, 0.0, 0.0), FAST_CLOUD, 0.1)
        _check_match(want, found, score, est, m.last_stats)
        tied += m.last_stats["leaves_tied"] > 1
        m.close()
    assert tied > 0, "expected the sparse-cloud fixture to produce tied optima"


def test_full_submap_reference_test_on_device(oracle, sm):
    """fast_correlative_scan_matcher_2d_test.cc:194-246, depth 6."""
    full_cloud = np.array([[-2.5, 0.5, 0], [-2.25, 0.5, 0], [0.0, 0.5, 0], [0.25, 1.6, 0],
                           [2.5, 0.5, 0], [2.0, 1.8, 0]], np.float32)
    rng = np.random.RandomState(42)
    for _ in range(6):
        u = rng.uniform(-1.0, 1.0, 6)
        pert = (10.0 * u[0], 10.0 * u[1], 1.6 * u[2])
        cloud = _transform(full_cloud, pert)
        true_pose = (2.0 * u[3], 2.0 * u[4], 0.5 * u[5])
        og = worlds.insert_range_data(oracle, 200, 200, 0.05, 5.0, 5.0, true_pose[:2],
                                      _transform(full_cloud, true_pose))
        grid = synthetic.GridSpec(og.cells, 0.05, 5.0, 5.0)
        m = sm.FastCorrelativeScanMatcher2D(grid, sm.FastCorrelativeScanMatcherOptions2D(3.0, 1.0, 6))
        om = oracle.FastCorrelativeScanMatcher2D(og, 3.0, 1.0, 6)
        want = om.match_full_submap(cloud, 0.1)
        found, score, est = m.MatchFullSubmap(cloud, 0.1)
        _check_match(want, found, score, est, m.last_stats)
        m.close()


# ---- batched form ------------------------------------------------------------
def test_match_batch_equals_single_calls(oracle, sm):
    grids = [worlds.small_world(300 + i)[0] for i in range(3)]
    worlds_ = [worlds.small_world(300 + i) for i in range(3)]
    opts = sm.FastCorrelativeScanMatcherOptions2D(1.5, 0.4, 5)
    matchers = [sm.FastCorrelativeScanMatcher2D(g, opts) for g in grids]
    clouds = [sm.DeviceCloud(w[3]) for w in worlds_]
    jobs = np.zeros(9, sm.JOB2D_DTYPE)
    rng = np.random.RandomState(0)
    k = 0
    for si in range(3):
        for ci in range(3):
            jobs[k]["stack_index"] = si
            jobs[k]["cloud_index"] = ci
            jobs[k]["initial_pose"] = worlds_[ci][2] + rng.uniform(-0.5, 0.5, 3) * [1, 1, 0.3]
            jobs[k]["min_score"] = 0.3
            k += 1
    res, stats = sm.match_batch(matchers, clouds, jobs, 1.5, 0.4)
    for k in range(9):
        om = oracle.FastCorrelativeScanMatcher2D(_ogrid(oracle, grids[jobs[k]["stack_index"]]),
                                                 1.5, 0.4, 5)
        want = om.match(jobs[k]["initial_pose"], worlds_[jobs[k]["cloud_index"]][3], 0.3)
        assert bool(res[k]["found"]) == want["found"]
        if want["found"]:
            assert np.float32(res[k]["score"]) == want["score"]
            np.testing.assert_array_equal(res[k]["pose_estimate"], want["pose"])
    assert stats["candidates_scored"] > 0
    for m in matchers:
        m.close()


# ---- BASELINE-size case through size-independent properties ------------------
def test_full_size_properties(oracle, sm):
    """1000x1000 @5cm, 1081 beams, depth 7 (BASELINE config 2 shape).  The oracle
    needs ~10 s per full-submap match at this size, so one case is compared
    directly and the rest through properties: the optimum of a scan taken inside
    the map scores above min_score, lands within one cell / one angular step of
    the truth, and re-scoring the returned candidate at level 0 reproduces the
    score exactly."""
    grid, occ = synthetic.make_grid2d(0, 1000)
    rng = np.random.RandomState(1)
    pose = synthetic.random_free_pose(occ, grid, rng)
    scan = synthetic.cast_scan(occ, grid, pose, seed=1)
    opts = sm.FastCorrelativeScanMatcherOptions2D(7.0, math.radians(30.0), 7)
    m = sm.FastCorrelativeScanMatcher2D(grid, opts)
    init = pose + np.array([1.5, -2.0, 0.2])
    found, score, est = m.Match(init, scan, 0.55)
    assert found
    om = oracle.FastCorrelativeScanMatcher2D(_ogrid(oracle, grid), 7.0, math.radians(30.0), 7)
    want = om.match(init, scan, 0.55)
    _check_match(want, found, score, est, m.last_stats)
    assert abs(est[0] - pose[0]) < 0.11 and abs(est[1] - pose[1]) < 0.11
    assert abs(est[2] - pose[2]) < 0.01
    # re-score the winner at full resolution
    ds, bounds = m.discretize(scan, init)
    st = m.last_stats
    sc, _ = m.score_candidates(0, ds, [[st["best_scan_index"], st["best_x_offset"],
                                         st["best_y_offset"]]])
    assert sc[0] == score
    # full-submap search on the device only: property checks
    found, score_f, est_f = m.MatchFullSubmap(scan, 0.5)
    # (a different window centre shifts the sub-cell alignment, so score_f need not
    # reach the local-window score; it must still clear min_score and find the pose)
    assert found and score_f > 0.5
    assert abs(est_f[0] - pose[0]) < 0.11 and abs(est_f[1] - pose[1]) < 0.11
    dth = (est_f[2] - pose[2] + math.pi) % (2 * math.pi) - math.pi
    assert abs(dth) < 0.01
    m.close()


# ---- RealTimeCorrelativeScanMatcher2D ---------------------------------------
@pytest.mark.parametrize("seed", range(4))
def test_rt_match_parity(oracle, sm, seed):
    grid, occ, pose, scan = worlds.small_world(400 + seed, size_cells=200, beams=361, max_range=6.0)
    rng = np.random.RandomState(seed)
    init = pose + np.array([rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05),
                            rng.uniform(-0.05, 0.05)])
    opts = sm.RealTimeCorrelativeScanMatcherOptions(0.1, math.radians(7.0), 0.1, 0.1)
    rt = sm.RealTimeCorrelativeScanMatcher2D(opts)
    score, est = rt.Match(init, scan, grid)
    want = oracle.rt2d_match(_ogrid(oracle, grid), scan, init, 0.1, math.radians(7.0), 0.1, 0.1)
    assert np.float32(score) == np.float32(want["score"])
    np.testing.assert_array_equal(est, want["pose"])
    assert rt.last_stats["candidates_scored"] == want["candidates_scored"]


@pytest.mark.parametrize("seed", range(3))
def test_rt_match_tsdf_parity(oracle, sm, seed):
    """TSDF grid type of the real-time matcher (real_time...2d.cc:38-59): ordered
    weighted float sums, bit-compared with the oracle on a synthetic TSDF."""
    grid, occ, pose, scan = worlds.small_world(500 + seed, size_cells=160, beams=241, max_range=5.0)
    rng = np.random.RandomState(seed)
    trunc, max_w = 0.3, 10.0
    # signed distance to the nearest wall cell (coarse, brute force on a small grid)
    ys, xs = np.nonzero(occ)
    yy, xx = np.mgrid[0:occ.shape[0], 0:occ.shape[1]]
    d = np.full(occ.shape, 1e9)
    for k in range(0, len(ys), 512):
        d = np.minimum(d, np.sqrt((yy[..., None] - ys[k:k + 512]) ** 2 +
                                  (xx[..., None] - xs[k:k + 512]) ** 2).min(axis=-1))
    tsd = np.clip(d * grid.resolution, -trunc, trunc).astype(np.float32)
    w = rng.uniform(0.0, max_w, occ.shape).astype(np.float32)
    tv = np.zeros(occ.shape, np.uint16)
    wv = np.zeros(occ.shape, np.uint16)
    known = rng.uniform(size=occ.shape) < 0.8
    for (y, x) in zip(*np.nonzero(known)):
        tv[y, x], wv[y, x] = oracle.tsdf_values(trunc, max_w, float(tsd[y, x]), float(w[y, x]))
    init = pose + np.array([0.04, -0.03, 0.02])
    opts = sm.RealTimeCorrelativeScanMatcherOptions(0.1, math.radians(5.0), 0.1, 0.1)
    rt = sm.RealTimeCorrelativeScanMatcher2D(opts)
    spec = sm.TSDF2DSpec(tv, wv, grid.resolution, grid.max_x, grid.max_y, trunc, max_w)
    score, est = rt.MatchTSDF(init, scan, spec)
    want = oracle.rt2d_match_tsdf(tv, wv, grid.resolution, grid.max_x, grid.max_y, trunc, max_w,
                                  scan, init, 0.1, math.radians(5.0), 0.1, 0.1)
    assert np.float32(score) == np.float32(want["score"])
    np.testing.assert_array_equal(est, want["pose"])
    assert 0.0 < score <= 1.0
