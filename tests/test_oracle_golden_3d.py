"""Pins the CPU oracle's 3D path against the reference's own tests
(mapping/3d/hybrid_grid_test.cc, mapping/internal/3d/scan_matching/*_test.cc).  CPU only."""
import math

import numpy as np
import pytest

from benchmarks import synthetic
from tests import worlds3d


# mapping/3d/hybrid_grid_test.cc:106-124 (GetCellIndex): resolution 2.f
def test_hybrid_get_cell_index(oracle):
    assert oracle.hybrid_get_cell_index(2.0, (0.0, 0.0, 0.0)) == (0, 0, 0)
    assert oracle.hybrid_get_cell_index(2.0, (0.0, 26.0, 0.0)) == (0, 13, 0)
    assert oracle.hybrid_get_cell_index(2.0, (14.0, 0.0, 0.0)) == (7, 0, 0)
    assert oracle.hybrid_get_cell_index(2.0, (14.0, 26.0, 0.0)) == (7, 13, 0)
    # "Check around the origin." (:116-123)
    assert oracle.hybrid_get_cell_index(2.0, (8.5, 11.5, 0.5)) == (4, 6, 0)
    assert oracle.hybrid_get_cell_index(2.0, (7.5, 12.5, 1.5)) == (4, 6, 1)
    assert oracle.hybrid_get_cell_index(2.0, (6.5, 14.5, 2.5)) == (3, 7, 1)
    assert oracle.hybrid_get_cell_index(2.0, (5.5, 13.5, 3.5)) == (3, 7, 2)


def test_hybrid_grid_sparse_semantics(oracle):
    """hybrid_grid_test.cc:70-104: unset cells read as unknown, the cube grows on demand."""
    idx = np.array([[1, 2, 3], [-70, 0, 0], [600, -600, 5]], np.int32)
    vals = np.array([oracle.probability_to_value(p) for p in (0.6, 0.7, 0.8)], np.uint16)
    g = oracle.HybridGrid(1.0, idx, vals)
    assert g.grid_size() == 2048            # 64 << bits with |600| < grid_size / 2
    assert g.get_probability(1, 2, 3) == pytest.approx(0.6, abs=1e-4)
    assert g.get_probability(600, -600, 5) == pytest.approx(0.8, abs=1e-4)
    assert g.get_probability(0, 0, 0) == pytest.approx(0.1, abs=1e-6)      # unknown
    assert g.get_probability(100000, 0, 0) == pytest.approx(0.1, abs=1e-6)  # outside the cube


# precomputation_grid_3d_test.cc:31-77: 1000 random voxels in [-50, 49]^3, res 2,
# depths 0..3 with full-resolution PrecomputeGrid, vs naive 3D window max, tol 1e-2.
def test_precomputation_grid_3d_matches_naive(oracle):
    rng = np.random.RandomState(23847)
    idx = rng.randint(-50, 50, size=(1000, 3)).astype(np.int32)
    prob = rng.uniform(0.1, 0.9, 1000).astype(np.float32)
    # later SetProbability calls overwrite earlier ones at the same cell
    _, last = np.unique(idx[::-1], axis=0, return_index=True)
    keep = len(idx) - 1 - last
    idx, prob = idx[keep], prob[keep]
    vals = np.array([oracle.probability_to_value(float(p)) for p in prob], np.uint16)
    hi = oracle.HybridGrid(2.0, idx, vals)
    opts = dict(branch_and_bound_depth=4, full_resolution_depth=4, min_rotational_score=0.0,
                min_low_resolution_score=0.0, linear_xy_search_window=1.0,
                linear_z_search_window=1.0, angular_search_window=0.1)
    m = oracle.FastCorrelativeScanMatcher3D(hi, hi, np.zeros(10, np.float32), opts)
    dense = np.full((160, 160, 160), np.float32(0.1), np.float32)  # [z, y, x], offset 60
    table = np.float32(0.1) + (vals.astype(np.float32) - 1) * np.float32(0.8 / 32766.0)
    dense[idx[:, 2] + 60, idx[:, 1] + 60, idx[:, 0] + 60] = table
    for depth in range(4):
        w = 1 << depth
        lo, lvl = m.level(depth)
        for _ in range(100):
            x, y, z = rng.randint(-50, 50, 3)
            want = dense[z + 60:z + 60 + w, y + 60:y + 60 + w, x + 60:x + 60 + w].max()
            lx, ly, lz = x - lo[0], y - lo[1], z - lo[2]
            v = 0
            if 0 <= lx < lvl.shape[2] and 0 <= ly < lvl.shape[1] and 0 <= lz < lvl.shape[0]:
                v = int(lvl[lz, ly, lx])
            got = 0.1 + v * (0.8 / 255.0)
            assert abs(got - want) < 1e-2


# rotational_scan_matcher_test.cc:28-36
def test_rotational_only_same_histogram_is_score_one(oracle):
    h = np.array([1.0, 43.0, 0.5, 0.3123, 23.0, 42.0, 0.0], np.float32)
    s = oracle.rotational_match(h, h, 0.0, [0.0, 1.0])
    assert s[0] == pytest.approx(1.0, abs=1e-6)
    assert s[1] < 1.0


# rotational_scan_matcher_test.cc:38-67
def test_rotational_interpolates_as_expected(oracle):
    n = 10
    per_bucket = np.float32(math.pi / n)
    unit = lambda k: np.eye(n, dtype=np.float32)[k]
    t = np.float32(0.0)
    while t < 1.0:
        expected = t / math.hypot(t, 1 - t)
        s = oracle.rotational_match(unit(3), unit(2), 0.0, [t * per_bucket])
        assert s[0] == pytest.approx(expected, abs=1e-6)
        s = oracle.rotational_match(unit(3), unit(2), 0.0, [(2 - t) * per_bucket])
        assert s[0] == pytest.approx(expected, abs=1e-6)
        s = oracle.rotational_match(unit(3), unit(4), 0.0, [-t * per_bucket, (t - 2) * per_bucket])
        assert s[0] == pytest.approx(expected, abs=1e-6)
        assert s[1] == pytest.approx(expected, abs=1e-6)
        t = np.float32(t + np.float32(0.1))


# fast_correlative_scan_matcher_3d_test.cc:146-204: 12-point axis cloud, HybridGrid(0.05),
# depth 6 / full_resolution_depth 6, windows 0.8 / 0.8 / 0.3, min_score 0.1.
def test_fast_matcher_3d_correct_pose(oracle):
    rng = np.random.RandomState(42)
    for trial in range(20):
        expected = worlds3d.random_pose(rng)
        hi = worlds3d.insert_cloud(oracle, 0.05, expected)
        m = oracle.FastCorrelativeScanMatcher3D(hi, hi, np.zeros(10, np.float32),
                                                worlds3d.TEST_OPTIONS)
        node = worlds3d.node_data(worlds3d.AXIS_CLOUD)
        ident = [0, 0, 0, 1, 0, 0, 0]
        r = m.match(ident, ident, node, 0.1)
        assert r["found"] and r["score"] > 0.1
        assert r["rotational_score"] > 0.09 and r["low_resolution_score"] > 0.14
        assert worlds3d.is_nearly(expected, r["pose"], 0.05), (expected, r["pose"])
        # a low-resolution cloud far away must be rejected by the gate (:167-174)
        far = worlds3d.node_data(worlds3d.AXIS_CLOUD, low=np.array([[42.0, 42.0, 42.0]], np.float32))
        assert not m.match(ident, ident, far, 0.1)["found"]


def test_fast_matcher_3d_full_submap(oracle):
    rng = np.random.RandomState(42)
    expected = worlds3d.random_pose(rng)
    hi = worlds3d.insert_cloud(oracle, 0.05, expected)
    m = oracle.FastCorrelativeScanMatcher3D(hi, hi, np.zeros(10, np.float32), worlds3d.TEST_OPTIONS)
    node = worlds3d.node_data(worlds3d.AXIS_CLOUD)
    r = m.match_full_submap([1, 0, 0, 0], [1, 0, 0, 0], node, 0.1)
    assert r["found"] and r["score"] > 0.1
    assert worlds3d.is_nearly(expected, r["pose"], 0.05), (expected, r["pose"])
    far = worlds3d.node_data(worlds3d.AXIS_CLOUD, low=np.array([[42.0, 42.0, 42.0]], np.float32))
    assert not m.match_full_submap([1, 0, 0, 0], [1, 0, 0, 0], far, 0.1)["found"]


# ---- real_time_correlative_scan_matcher_3d_test.cc:35-136 -------------------------------
def _rt3d_fixture(oracle):
    """The reference fixture: 7 points, grid cells of expected_pose * point set to
    probability 1 (clamped to kMaxProbability), resolution 0.1; options 0.3 m / 1 deg /
    1e-1 / 1."""
    pts = np.array([[-3, 2, 0], [-4, 2, 0], [-5, 2, 0], [-6, 2, 0], [-6, 3, 1], [-6, 4, 2],
                    [-7, 3, 1]], np.float32)
    idx = [oracle.hybrid_get_cell_index(0.1, p + np.array([-1, 0, 0], np.float32)) for p in pts]
    v = oracle.probability_to_value(1.0)
    grid = oracle.HybridGrid(0.1, np.array(idx, np.int32), np.full(len(idx), v, np.uint16))
    return pts, grid, np.array(idx, np.int32), np.full(len(idx), v, np.uint16)


def _mat4(pose):
    t, (w, x, y, z) = pose[:3], pose[3:]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    M = np.eye(4)
    M[:3, :3] = R
    M[:3, 3] = t
    return M


def rt3d_reference_initial_poses():
    """PerfectEstimate, AlongX, AlongZ, AlongXYZ, RotationAroundX / Y / YZ.  Eigen::AngleAxisd
    with the un-normalised axis (0, 1, 1) yields the quaternion (cos a/2, sin a/2 * axis)."""
    a = 0.8 / 180.0 * math.pi
    c, s = math.cos(a / 2), math.sin(a / 2)
    return [[-1, 0, 0, 1, 0, 0, 0], [-0.8, 0, 0, 1, 0, 0, 0], [-1, 0, -0.2, 1, 0, 0, 0],
            [-0.9, -0.2, 0.2, 1, 0, 0, 0], [-1, 0, 0, c, s, 0, 0], [-1, 0, 0, c, 0, s, 0],
            [-1, 0, 0, c, 0, s, s]]


def is_nearly(pose, expected, eps):
    """transform::IsNearly = Eigen isApprox on the 4x4 matrices
    (transform/rigid_transform_test_helpers.h:42-46)."""
    A, B = _mat4(np.asarray(pose, float)), _mat4(np.asarray(expected, float))
    return np.linalg.norm(A - B) <= eps * min(np.linalg.norm(A), np.linalg.norm(B))


def test_rt3d_reference_test_cases(oracle):
    pts, grid, _, _ = _rt3d_fixture(oracle)
    for init in rt3d_reference_initial_poses():
        r = oracle.rt3d_match(grid, pts, init, 0.3, math.radians(1.0), 1e-1, 1.0)
        assert r["score"] > 0.0
        assert is_nearly(r["pose"], [-1, 0, 0, 1, 0, 0, 0], 1e-3), (init, r)
        # (2 * 3 + 1)^3 translations x (2 * 1 + 1)^3 rotations
        assert r["candidates_scored"] == 343 * 27
