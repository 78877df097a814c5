"""Pins the oracle's CeresScanMatcher2D restatement (oracle/oracle_ceres2d.{h,cc}) to the
known answers the reference holds for it.  Ceres itself is absent from the reference tree
and from this image (pinned by the reference at commit 58c5edae, bazel/repositories.bzl:
136-142), so these tests — the reference's own, with its tolerances — are the pin; bit parity
against Ceres is unpinned and DESIGN.md says so.  CPU only.
"""
import math

import numpy as np
import pytest


def _unknown_grid(oracle, nx, ny, resolution, max_x, max_y):
    return oracle.Grid2D(np.zeros((ny, nx), np.uint16), resolution, max_x, max_y)


# internal/2d/scan_matching/occupied_space_cost_function_2d_test.cc:31-51
def test_occupied_space_cost_function_smoke(oracle):
    grid = _unknown_grid(oracle, 2, 2, 1.0, 1.0, 1.0)
    res, jac = oracle.ceres2d_evaluate(grid, [[0.0, 0.0, 0.0]], [0.0, 0.0, 0.0], [0.0, 0.0], 0.0,
                                       occupied_space_weight=1.0)
    k_max_probability = np.float64(np.float32(1.0) - np.float32(0.1))
    assert res[0] == k_max_probability          # DoubleEq(kMaxProbability)
    assert np.all(jac[0] == 0.0)                # a constant field has no gradient
    res2, _ = oracle.ceres2d_evaluate(grid, [[0.0, 0.0, 0.0]], [0.0, 0.0, 0.0], [0.0, 0.0], 0.0,
                                      occupied_space_weight=1.0, jacobian=False)
    assert res2[0] == k_max_probability


def _rigid2_matrix(p):
    c, s = math.cos(p[2]), math.sin(p[2])
    return np.array([[c, -s, p[0]], [s, c, p[1]], [0.0, 0.0, 1.0]])


def is_nearly(pose, expected, epsilon):
    """transform::IsNearly (transform/rigid_transform_test_helpers.h:42-46) =
    Eigen isApprox of the homogeneous matrices: |a - b|_F <= epsilon * min(|a|_F, |b|_F)."""
    a, b = _rigid2_matrix(pose), _rigid2_matrix(expected)
    return np.linalg.norm(a - b) <= epsilon * min(np.linalg.norm(a), np.linalg.norm(b))


def _ceres_test_fixture(oracle):
    # ceres_scan_matcher_2d_test.cc:36-62: 20 x 20 cells of 1 m, max (10, 10); the cell that
    # holds (-3.5, 2.5) is occupied; one point at (-3, 2); weights 1 / 0.1 / 1.5, 50 iterations
    cells = np.zeros((20, 20), np.uint16)
    cx, cy = oracle.get_cell_index(1.0, 10.0, 10.0, -3.5, 2.5)
    cells[cy, cx] = oracle.correspondence_cost_to_value(
        float(np.float32(1.0) - np.float32(oracle.constant(1))))   # SetProbability(kMaxProbability)
    grid = oracle.Grid2D(cells, 1.0, 10.0, 10.0)
    cloud = np.array([[-3.0, 2.0, 0.0]], np.float32)
    opts = dict(occupied_space_weight=1.0, translation_weight=0.1, rotation_weight=1.5,
                use_nonmonotonic_steps=True, max_num_iterations=50)
    return grid, cloud, opts


# ceres_scan_matcher_2d_test.cc:64-97: testPerfectEstimate, testOptimizeAlongX/Y/XY
@pytest.mark.parametrize("start", [(-0.5, 0.5), (-0.3, 0.5), (-0.45, 0.3), (-0.3, 0.3)])
def test_ceres_scan_matcher_known_answers(oracle, start):
    grid, cloud, opts = _ceres_test_fixture(oracle)
    init = [start[0], start[1], 0.0]
    out = oracle.ceres2d_match(grid, cloud, init[:2], init, **opts)
    assert out["final_cost"] == pytest.approx(0.0, abs=1e-2)          # EXPECT_NEAR(0., final_cost, 1e-2)
    # EXPECT_THAT(pose, transform::IsNearly(Translation(-0.5, 0.5), 1e-2))
    assert is_nearly(out["pose"], (-0.5, 0.5, 0.0), 1e-2)
    assert out["final_cost"] <= out["initial_cost"]


def smooth_grid(oracle, nx=50, ny=40, resolution=0.05, max_x=1.3, max_y=0.9):
    """A smooth correspondence-cost field (two Gaussian wells), so difference quotients of the
    interpolant converge."""
    yy, xx = np.mgrid[0:ny, 0:nx]
    cost = 0.9 - 0.5 * np.exp(-((xx - 20) ** 2 + (yy - 15) ** 2) / 60.0) \
        - 0.3 * np.exp(-((xx - 35) ** 2 + (yy - 28) ** 2) / 90.0)
    cells = 1 + np.rint((cost - 0.1) / 0.8 * 32766.0)
    return oracle.Grid2D(cells.astype(np.uint16), resolution, max_x, max_y)


def test_jacobian_matches_finite_differences(oracle):
    """The dual-number derivative against central differences of the plain-double path.  The
    kPadding shift leaves 2^-23 of a cell of coordinate resolution, so the step is 1e-3 cells."""
    rng = np.random.RandomState(3)
    grid = smooth_grid(oracle)
    cloud = np.concatenate([rng.uniform(-0.5, 0.5, (64, 2)), np.zeros((64, 1))], 1).astype(np.float32)
    pose = np.array([0.21, -0.13, 0.3])
    res, jac = oracle.ceres2d_evaluate(grid, cloud, pose, [0.2, -0.1], 0.25)
    assert np.abs(jac[:64]).max() > 0.1
    h = 5e-5
    for k in range(3):
        d = np.zeros(3)
        d[k] = h
        rp, _ = oracle.ceres2d_evaluate(grid, cloud, pose + d, [0.2, -0.1], 0.25, jacobian=False)
        rm, _ = oracle.ceres2d_evaluate(grid, cloud, pose - d, [0.2, -0.1], 0.25, jacobian=False)
        fd = (rp - rm) / (2 * h)
        assert np.allclose(fd, jac[:, k], rtol=2e-3, atol=2e-3)
        assert np.allclose(fd[-3:], jac[-3:, k], atol=1e-8)


def test_refinement_descends_and_returns_the_best_iterate(oracle):
    """On a smooth field the refinement lowers the cost, stays near the start (the priors)
    and, under non-monotonic steps, returns the lowest-cost iterate it visited."""
    rng = np.random.RandomState(4)
    grid = smooth_grid(oracle)
    ang = rng.uniform(0, 2 * math.pi, 200)
    # a ring of returns around the first well's centre (cell (20, 15) -> world)
    # cells[row, column]: world x = max_x - (row + 0.5) * res, world y = max_y - (column + 0.5) * res
    cx, cy = 1.3 - (15 + 0.5) * 0.05, 0.9 - (20 + 0.5) * 0.05
    ring = np.stack([0.08 * np.cos(ang), 0.08 * np.sin(ang), np.zeros_like(ang)], 1).astype(np.float32)
    for nm in (True, False):
        init = np.array([cx + 0.06, cy - 0.05, 0.1])
        out = oracle.ceres2d_match(grid, ring, init[:2], init, use_nonmonotonic_steps=nm,
                                   max_num_iterations=30)
        assert out["final_cost"] < out["initial_cost"]
        assert out["num_successful_steps"] >= 1
        assert np.linalg.norm(out["pose"][:2] - np.array([cx, cy])) < np.linalg.norm(init[:2] - np.array([cx, cy]))
        res, _ = oracle.ceres2d_evaluate(grid, ring, out["pose"], init[:2], init[2], jacobian=True)
        assert 0.5 * float(res @ res) == pytest.approx(out["final_cost"], rel=1e-12)
