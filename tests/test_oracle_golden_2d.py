"""Pins the CPU oracle (oracle/) against the reference's own known-answer tests
for the 2D path.  Every test cites the reference test it ports (paths relative
to /root/reference/cartographer/).  CPU only.
"""
import math

import numpy as np
import pytest

from tests import worlds


# mapping/probability_values_test.cc + value_conversion_tables_test.cc
def test_probability_constants_and_tables(oracle):
    kmin, kmax = oracle.constant(0), oracle.constant(1)
    assert kmin == pytest.approx(0.1, abs=1e-7) and kmax == pytest.approx(0.9, abs=1e-7)
    # probability_values_test.cc:24-36: value <-> float round trip, unknown == 0
    t = oracle.value_to_cost_table(oracle.constant(2), oracle.constant(3))
    assert t.shape == (65536,)
    assert t[0] == np.float32(oracle.constant(3))             # unknown -> max cost
    assert t[1] == pytest.approx(oracle.constant(2), abs=1e-6)  # lower bound
    assert t[32767] == pytest.approx(oracle.constant(3), abs=1e-6)
    # value_conversion_tables_test.cc:45-70: entries with the update marker repeat
    assert np.array_equal(t[:32768], t[32768:])
    assert np.all(np.diff(t[1:32768]) > 0)
    for v in (1, 2, 100, 16000, 32766, 32767):
        assert oracle.correspondence_cost_to_value(float(t[v])) == v
    # probability_values_test.cc:38-60: clamping
    assert oracle.probability_to_value(0.0) == 1
    assert oracle.probability_to_value(1.0) == 32767


# mapping/2d/map_limits_test.cc + probability_grid_test.cc:151-181 (GetCellIndex)
def test_get_cell_index(oracle):
    # probability_grid_test.cc:151-181: MapLimits(2., (8, 14), CellLimits(14, 8))
    L = (2.0, 8.0, 14.0)
    assert oracle.get_cell_index(*L, 7.0, 13.0) == (0, 0)
    assert oracle.get_cell_index(*L, 7.0, -13.0) == (13, 0)
    assert oracle.get_cell_index(*L, -7.0, 13.0) == (0, 7)
    assert oracle.get_cell_index(*L, -7.0, -13.0) == (13, 7)
    # around the origin (:170-181)
    assert oracle.get_cell_index(*L, 0.5, 0.5) == (6, 3)
    assert oracle.get_cell_index(*L, 1.5, 1.5) == (6, 3)
    assert oracle.get_cell_index(*L, 0.5, -0.5) == (7, 3)
    assert oracle.get_cell_index(*L, -0.5, 0.5) == (6, 4)
    assert oracle.get_cell_index(*L, -0.5, -0.5) == (7, 4)


# internal/2d/scan_matching/correlative_scan_matcher_test.cc:26-40, 42-56
def test_search_parameters_and_candidate(oracle):
    c = oracle.candidate(4, 5, 0.03, 0.05, 3, 4, -5)
    assert c[0] == pytest.approx(0.25, abs=1e-9)
    assert c[1] == pytest.approx(-0.2, abs=1e-9)
    assert c[2] == pytest.approx(-0.06, abs=1e-9)


# correlative_scan_matcher_test.cc:58-70
def test_generate_rotated_scans(oracle):
    scans = oracle.generate_rotated_scans([[-1.0, 1.0, 0.0]], 1, math.pi / 2.0)
    assert scans.shape == (3, 1, 3)
    np.testing.assert_allclose(scans[0, 0, :2], [1.0, 1.0], atol=1e-6)
    np.testing.assert_allclose(scans[1, 0, :2], [-1.0, 1.0], atol=1e-6)
    np.testing.assert_allclose(scans[2, 0, :2], [-1.0, -1.0], atol=1e-6)


TEST_CLOUD_7 = np.array([[0.025, 0.175, 0], [-0.025, 0.175, 0], [-0.075, 0.175, 0],
                         [-0.125, 0.175, 0], [-0.125, 0.125, 0], [-0.125, 0.075, 0],
                         [-0.125, 0.025, 0]], np.float32)


# correlative_scan_matcher_test.cc:72-96 — 7 exact integer cell indices
def test_discretize_scans_exact(oracle):
    scans = oracle.generate_rotated_scans(TEST_CLOUD_7, 0, 0.0)
    d = oracle.discretize_scans(0.05, 0.05, 0.25, 6, 6, scans)
    assert d.shape == (1, 7, 2)
    assert d[0].tolist() == [[1, 0], [1, 1], [1, 2], [1, 3], [2, 3], [3, 3], [4, 3]]


def _naive_window_max(prob, width):
    """max over [x, x+w) x [y, y+w) of GetProbability (0.1 outside)."""
    ny, nx = prob.shape
    pad = np.full((ny + width, nx + width), np.float32(0.1), np.float32)
    pad[:ny, :nx] = prob
    out = np.full((ny, nx), -np.inf, np.float32)
    for dy in range(width):
        for dx in range(width):
            out = np.maximum(out, pad[dy:dy + ny, dx:dx + nx])
    return out


def _to_score(oracle, v):
    kmin, kmax = np.float32(oracle.constant(0)), np.float32(oracle.constant(1))
    return kmin + np.float32(v) * ((kmax - kmin) / np.float32(255.0))


def _random_u8_grid(oracle, nx, ny, x0, y0, seed):
    """fast_correlative_scan_matcher_2d_test.cc:40-55: cells set to ToScore(u8)."""
    rng = np.random.RandomState(seed)
    cells = np.zeros((ny, nx), np.uint16)
    vals = rng.randint(0, 256, size=(ny - y0, nx - x0))
    lut = np.array([oracle.correspondence_cost_to_value(float(np.float32(1.0) - _to_score(oracle, v)))
                    for v in range(256)], np.uint16)
    cells[y0:, x0:] = lut[vals]
    return cells


# fast_correlative_scan_matcher_2d_test.cc:37-77 (CorrectValues) and :79-117 (Tiny)
@pytest.mark.parametrize("nx,ny,x0,y0,widths", [(250, 250, 50, 50, (1, 2, 3, 8)),
                                                 (4, 4, 0, 0, (1, 2, 3, 8, 200))])
def test_precomputation_grid_matches_naive_max(oracle, nx, ny, x0, y0, widths):
    cells = _random_u8_grid(oracle, nx, ny, x0, y0, 42)
    table = oracle.value_to_cost_table(oracle.constant(2), oracle.constant(3))
    prob = (np.float32(1.0) - table[cells]).astype(np.float32)
    for w in widths:
        pg = oracle.precompute_grid2d(cells, oracle.constant(2), oracle.constant(3), w)
        assert pg.shape == (ny + w - 1, nx + w - 1)
        # GetValue(xy) reads local index xy - offset, offset = (-w+1, -w+1)
        got = _to_score(oracle, pg[w - 1:w - 1 + ny, w - 1:w - 1 + nx].astype(np.float32))
        np.testing.assert_allclose(got, _naive_window_max(prob, w), atol=1e-4)


FAST_CLOUD = np.array([[-2.5, 0.5, 0], [-2.0, 0.5, 0], [0.0, -0.5, 0], [0.5, -1.6, 0],
                       [2.5, 0.5, 0], [2.5, 1.7, 0]], np.float32)
FULL_CLOUD = np.array([[-2.5, 0.5, 0], [-2.25, 0.5, 0], [0.0, 0.5, 0], [0.25, 1.6, 0],
                       [2.5, 0.5, 0], [2.0, 1.8, 0]], np.float32)


def _mat(pose):
    c, s = math.cos(pose[2]), math.sin(pose[2])
    return np.array([[c, -s, pose[0]], [s, c, pose[1]], [0, 0, 1]], np.float32)


def is_nearly(a, b, eps):
    """transform/rigid_transform_test_helpers.h:42-46: Eigen isApprox on the
    homogeneous matrices: |A - B|_F^2 <= eps^2 * min(|A|_F^2, |B|_F^2)."""
    A, B = _mat(a), _mat(b)
    return float(((A - B) ** 2).sum()) <= eps * eps * min(float((A ** 2).sum()),
                                                         float((B ** 2).sum()))


def _transform(cloud, pose):
    c, s = math.cos(pose[2]), math.sin(pose[2])
    out = cloud.copy()
    out[:, 0] = c * cloud[:, 0] - s * cloud[:, 1] + pose[0]
    out[:, 1] = s * cloud[:, 0] + c * cloud[:, 1] + pose[1]
    return out.astype(np.float32)


# fast_correlative_scan_matcher_2d_test.cc:144-192 (CorrectPose): 50 random poses,
# 200x200 @0.05 grid built by inserting the cloud at the true pose, depth 3,
# window 3 m / 1 rad, min_score 0.1, pose within 0.03.
def test_fast_matcher_correct_pose(oracle):
    rng = np.random.RandomState(42)
    for _ in range(50):
        u = rng.uniform(-1.0, 1.0, 3)
        expected = (2.0 * u[0], 2.0 * u[1], 0.5 * u[2])
        grid = worlds.insert_range_data(oracle, 200, 200, 0.05, 5.0, 5.0, expected[:2],
                                        _transform(FAST_CLOUD, expected))
        m = oracle.FastCorrelativeScanMatcher2D(grid, 3.0, 1.0, 3)
        r = m.match((0.0, 0.0, 0.0), FAST_CLOUD, 0.1)
        assert r["found"] and r["score"] > 0.1
        assert is_nearly(expected, r["pose"], 0.03), (expected, r["pose"])


# fast_correlative_scan_matcher_2d_test.cc:194-246 (FullSubmapMatching), depth 6.
def test_fast_matcher_full_submap(oracle):
    rng = np.random.RandomState(42)
    for _ in range(20):
        u = rng.uniform(-1.0, 1.0, 6)
        pert = (10.0 * u[0], 10.0 * u[1], 1.6 * u[2])
        cloud = _transform(FULL_CLOUD, pert)
        true_pose = (2.0 * u[3], 2.0 * u[4], 0.5 * u[5])     # pose of the unperturbed cloud
        # expected = true_pose * pert^-1
        th = true_pose[2] - pert[2]
        c, s = math.cos(th), math.sin(th)
        ex = true_pose[0] - (c * pert[0] - s * pert[1])
        ey = true_pose[1] - (s * pert[0] + c * pert[1])
        grid = worlds.insert_range_data(oracle, 200, 200, 0.05, 5.0, 5.0, true_pose[:2],
                                        _transform(FULL_CLOUD, true_pose))
        m = oracle.FastCorrelativeScanMatcher2D(grid, 3.0, 1.0, 6)
        r = m.match_full_submap(cloud, 0.1)
        assert r["found"] and r["score"] > 0.1
        assert is_nearly((ex, ey, th), r["pose"], 0.03), ((ex, ey, th), r["pose"])


# real_time_correlative_scan_matcher_2d_test.cc:125-160 (ScorePerfect/Partial):
# 6x6 grid @0.05 max (0.05, 0.25); cloud inserted at identity.
def test_rt_score_candidates_bands(oracle):
    grid = worlds.insert_range_data(oracle, 6, 6, 0.05, 0.05, 0.25, (0.0, 0.0), TEST_CLOUD_7,
                                    grow=False)
    scans = oracle.generate_rotated_scans(TEST_CLOUD_7, 0, 0.0)
    d = oracle.discretize_scans(0.05, 0.05, 0.25, 6, 6, scans)
    perfect = oracle.rt2d_score_candidates(grid, d, 0, 0, 0.0, 0.0, 0.0, [[0, 0, 0]])
    # ScoreCandidatesProbabilityGrid expects 0.7 for the perfect candidate (:136)
    assert perfect[0] == pytest.approx(0.7, abs=1e-2)
    partial = oracle.rt2d_score_candidates(grid, d, 0, 0, 0.0, 0.0, 0.0, [[0, 0, 1]])
    # shifted by one cell: strictly less than perfect, more than all-unknown (:152-160)
    assert 0.1 < partial[0] < 0.7


def test_rt_match_recovers_offset(oracle):
    grid = worlds.insert_range_data(oracle, 6, 6, 0.05, 0.05, 0.25, (0.0, 0.0), TEST_CLOUD_7,
                                    grow=False)
    r = oracle.rt2d_match(grid, TEST_CLOUD_7, (0.0, 0.0, 0.0), 0.6, 0.16, 0.0, 0.0)
    assert r["score"] == pytest.approx(0.7, abs=1e-2)
    assert abs(r["pose"][0]) < 1e-9 and abs(r["pose"][1]) < 1e-9


# real_time_correlative_scan_matcher_2d_test.cc:146-160, 180-198 (TSDF grid type): a
# cloud lying exactly on the zero crossing scores ~1; shifted by one cell it scores
# between 1 - 4/(7*6) and 1 (3 of 7 points stay on the surface, truncation 0.3 = 6 cells).
def test_rt_tsdf_score_bands(oracle):
    trunc, max_w, res = 0.3, 10.0, 0.05
    nx = ny = 20
    max_x, max_y = 0.3, 0.5
    tv = np.zeros((ny, nx), np.uint16)
    wv = np.zeros((ny, nx), np.uint16)
    cells = [oracle.get_cell_index(res, max_x, max_y, float(p[0]), float(p[1]))
             for p in TEST_CLOUD_7]
    for cy in range(ny):
        for cx in range(nx):
            dist = min(math.hypot(cx - a, cy - b) for a, b in cells) * res
            if dist <= trunc:
                tv[cy, cx], wv[cy, cx] = oracle.tsdf_values(trunc, max_w, dist, 5.0)
    perfect = oracle.rt2d_match_tsdf(tv, wv, res, max_x, max_y, trunc, max_w, TEST_CLOUD_7,
                                     (0.0, 0.0, 0.0), 0.0, 0.0, 0.0, 0.0)
    assert perfect["score"] > 0.95 and perfect["score"] <= 1.0 + 1e-6
    shifted = TEST_CLOUD_7.copy()
    shifted[:, 0] -= 0.05  # one cell along index y
    part = oracle.rt2d_match_tsdf(tv, wv, res, max_x, max_y, trunc, max_w, shifted,
                                  (0.0, 0.0, 0.0), 0.0, 0.0, 0.0, 0.0)
    assert 1.0 - 4.0 / (7.0 * 6.0) - 1e-3 < part["score"] < 1.0
