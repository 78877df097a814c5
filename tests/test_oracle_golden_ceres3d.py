"""Pins the oracle's CeresScanMatcher3D restatement (oracle/oracle_ceres3d.{h,cc}) to what
the reference holds for it: interpolated_grid_test.cc and ceres_scan_matcher_3d_test.cc.
Ceres itself is absent (pinned by the reference at commit 58c5edae), so bit parity against
Ceres is unpinned — see the header of oracle_ceres3d.h.  CPU only."""
import math

import numpy as np
import pytest

from tests.test_oracle_golden_3d import is_nearly

POINTS = np.array([[-3, 2, 0], [-4, 2, 0], [-5, 2, 0], [-6, 2, 0], [-6, 3, 1], [-6, 4, 2],
                   [-7, 3, 1]], np.float32)


def _grid(oracle, resolution, points):
    idx = np.array([oracle.hybrid_get_cell_index(resolution, p) for p in points], np.int32)
    val = np.full(len(idx), oracle.probability_to_value(1.0), np.uint16)
    return oracle.HybridGrid(resolution, idx, val)


# interpolated_grid_test.cc:30-62 — the interpolation reproduces the grid at voxel centres
def test_interpolates_grid_points(oracle):
    grid = _grid(oracle, 0.1, POINTS)
    res = float(np.float32(0.1))
    checked = 0
    z = -1.0
    while z < 3.0:
        y = 1.0
        while y < 5.0:
            # (x walks the whole row in the reference; every 5th row here keeps the test fast)
            if checked % 5 == 0:
                x = -8.0
                while x < -2.0:
                    ci = oracle.hybrid_get_cell_index(0.1, [x, y, z])
                    assert oracle.interpolated_probability(grid, x, y, z) == pytest.approx(
                        grid.get_probability(*ci), abs=1e-6)
                    x += res
            checked += 1
            y += res
        z += res


# interpolated_grid_test.cc:64-88 — monotonic between neighbouring voxel centres along x
def test_monotonic_between_grid_points_in_x(oracle):
    grid = _grid(oracle, 0.1, POINTS)
    res = float(np.float32(0.1))
    step = res / 10.0
    seen = 0
    for p in POINTS:
        for dx in (-res, 0.0):
            x, y, z = float(p[0]) + dx, float(p[1]), float(p[2])
            a = grid.get_probability(*oracle.hybrid_get_cell_index(0.1, [x, y, z]))
            b = grid.get_probability(*oracle.hybrid_get_cell_index(0.1, [x + res, y, z]))
            if abs(b - a) < 1e-6:
                continue
            s = step
            while s < res - 2 * step:
                d = (oracle.interpolated_probability(grid, x + s + step, y, z) -
                     oracle.interpolated_probability(grid, x + s, y, z))
                assert (b - a) * d > 0.0
                s += step
            seen += 1
    assert seen >= 4


def test_interpolation_gradient_matches_differences(oracle):
    grid = _grid(oracle, 0.1, POINTS)
    rng = np.random.RandomState(2)
    for _ in range(50):
        p = POINTS[rng.randint(len(POINTS))] + rng.uniform(-0.09, 0.09, 3)
        f, g = oracle.interpolated_probability(grid, *p, gradient=True)
        h = 1e-7
        for k in range(3):
            d = np.zeros(3)
            d[k] = h
            fd = (oracle.interpolated_probability(grid, *(p + d)) -
                  oracle.interpolated_probability(grid, *(p - d))) / (2 * h)
            assert fd == pytest.approx(g[k], abs=2e-4)


def _fixture(oracle):
    """ceres_scan_matcher_3d_test.cc:36-75: resolution 1, cells of expected_pose * point set to
    probability 1; weights 1 / 0.01 / 0.1, non-monotonic steps, 10 iterations.  The fixture's
    intensity block is left out (the constraint builder passes no intensity grid)."""
    grid = _grid(oracle, 1.0, POINTS + np.array([-1, 0, 0], np.float32))
    opts = dict(occupied_space_weights=[1.0], translation_weight=0.01, rotation_weight=0.1,
                use_nonmonotonic_steps=True, max_num_iterations=10)
    return grid, opts


# ceres_scan_matcher_3d_test.cc:99-116: PerfectEstimate, AlongX, AlongZ, AlongXYZ
@pytest.mark.parametrize("start", [(-1.0, 0.0, 0.0), (-0.8, 0.0, 0.0), (-1.0, 0.0, -0.2),
                                   (-0.9, -0.2, 0.2)])
def test_ceres_scan_matcher_3d_known_answers(oracle, start):
    grid, opts = _fixture(oracle)
    init = [start[0], start[1], start[2], 1.0, 0.0, 0.0, 0.0]
    out = oracle.ceres3d_match([(POINTS, grid)], init[:3], init, **opts)
    assert out["final_cost"] == pytest.approx(0.0, abs=1e-2)
    assert is_nearly(out["pose"], [-1, 0, 0, 1, 0, 0, 0], 3e-2), out


# ceres_scan_matcher_3d_test.cc:118-131: FullPoseCorrection
def test_ceres_scan_matcher_3d_full_pose_correction(oracle):
    grid, opts = _fixture(oracle)
    a = 0.05
    c, s = math.cos(a), math.sin(a)
    # the cloud rotated by +0.05 rad about z (float, as TransformPointCloud does)
    cloud = np.stack([np.float32(c) * POINTS[:, 0] - np.float32(s) * POINTS[:, 1],
                      np.float32(s) * POINTS[:, 0] + np.float32(c) * POINTS[:, 1],
                      POINTS[:, 2]], 1).astype(np.float32)
    expected = [-1, 0, 0, math.cos(-a / 2), 0, 0, math.sin(-a / 2)]   # expected * rotation^-1
    init = [-0.95, -0.05, 0.05, math.cos(a / 2), math.sin(a / 2), 0.0, 0.0]   # about x
    out = oracle.ceres3d_match([(cloud, grid)], init[:3], init, **opts)
    assert out["final_cost"] == pytest.approx(0.0, abs=1e-2)
    assert is_nearly(out["pose"], expected, 3e-2), out


def test_jacobian_matches_differences_through_the_parameterisation(oracle):
    """Tangent-space Jacobian against central differences of r(x (+) delta)."""
    grid, _ = _fixture(oracle)
    lo = _grid(oracle, 2.0, POINTS + np.array([-1, 0, 0], np.float32))
    pose = np.array([-0.93, 0.04, 0.07, math.cos(0.1), 0.0, math.sin(0.1) * 0.6, math.sin(0.1) * 0.8])
    tq = [1.0, 0.0, 0.0, 0.0]
    pairs = [(POINTS, grid), (POINTS[:5], lo)]
    res, jac = oracle.ceres3d_evaluate(pairs, pose, [-1, 0, 0], tq)
    assert res.shape == (7 + 5 + 6,) and jac.shape == (18, 6)

    def plus(x, d):
        out = x.copy()
        out[:3] += d[:3]
        n = np.linalg.norm(d[3:])
        if n > 0:
            z = np.concatenate([[math.cos(n)], math.sin(n) / n * d[3:]])
            w = x[3:]
            out[3:] = [z[0] * w[0] - z[1] * w[1] - z[2] * w[2] - z[3] * w[3],
                       z[0] * w[1] + z[1] * w[0] + z[2] * w[3] - z[3] * w[2],
                       z[0] * w[2] - z[1] * w[3] + z[2] * w[0] + z[3] * w[1],
                       z[0] * w[3] + z[1] * w[2] - z[2] * w[1] + z[3] * w[0]]
        return out
    h = 1e-6
    for k in range(6):
        d = np.zeros(6)
        d[k] = h
        rp, _ = oracle.ceres3d_evaluate(pairs, plus(pose, d), [-1, 0, 0], tq, jacobian=False)
        rm, _ = oracle.ceres3d_evaluate(pairs, plus(pose, -d), [-1, 0, 0], tq, jacobian=False)
        assert np.allclose((rp - rm) / (2 * h), jac[:, k], rtol=0, atol=2e-5)
