"""Device CeresScanMatcher2D (csm_ceres_match2d_batch / csm_ceres_evaluate2d) against the
oracle's restatement (oracle/oracle_ceres2d.cc) and against the reference's own known
answers (ceres_scan_matcher_2d_test.cc, occupied_space_cost_function_2d_test.cc).

Floating point: doubles on both sides, no FMA contraction.  Per-point residuals / Jacobian
rows are compared to 1e-12 (they are the same operations in the same order; the test hook
takes cos / sin from the host).  Solved poses are compared to 1e-7 — north_star's float
tolerance is 1e-5 — because the device sums the normal equations in block-tree order and
evaluates cos / sin with the device routines (<= 2 ulp)."""
import math

import numpy as np
import pytest

from benchmarks import synthetic
from cartographer_b200 import constraint_builder as cb
from cartographer_b200 import scan_matching as sm
from cartographer_b200._lib import CsmError
from tests.test_constraint_builder_cpu import OracleExecutor, _fill, _small_queue
from tests.test_oracle_golden_ceres2d import _ceres_test_fixture, is_nearly, smooth_grid

pytestmark = pytest.mark.gpu

POSE_TOL = 1e-7


def _same_solution(got_pose, got_sum, want):
    assert np.allclose(got_pose, want["pose"], rtol=0, atol=POSE_TOL), (got_pose, want["pose"])
    assert got_sum["initial_cost"] == pytest.approx(want["initial_cost"], rel=1e-12)
    assert got_sum["final_cost"] == pytest.approx(want["final_cost"], rel=1e-9)
    assert got_sum["iterations"] == want["iterations"]
    assert got_sum["num_successful_steps"] == want["num_successful_steps"]
    assert got_sum["termination"] == want["termination"]


# occupied_space_cost_function_2d_test.cc:31-51
def test_occupied_space_cost_function_smoke_on_device(oracle):
    grid = oracle.Grid2D(np.zeros((2, 2), np.uint16), 1.0, 1.0, 1.0)
    dev = sm.RealTimeGrid2D(grid)
    m = sm.CeresScanMatcher2D(sm.CeresScanMatcherOptions2D(occupied_space_weight=1.0))
    res, jac = m.Evaluate(dev, [[0.0, 0.0, 0.0]], [0.0, 0.0, 0.0], [0.0, 0.0], 0.0)
    assert res[0] == np.float64(np.float32(1.0) - np.float32(0.1))   # DoubleEq(kMaxProbability)
    assert np.all(jac[0] == 0.0)
    dev.close()


@pytest.mark.parametrize("theta", [0.0, 0.3, -2.4])
def test_residuals_and_jacobian_equal_the_oracle(oracle, theta):
    rng = np.random.RandomState(11)
    for grid in (smooth_grid(oracle),
                 oracle.Grid2D(rng.randint(0, 32768, size=(40, 50)).astype(np.uint16), 0.05, 1.3, 0.9)):
        dev = sm.RealTimeGrid2D(grid)
        # points inside, on the border of and outside the grid (kMaxCorrespondenceCost there)
        cloud = np.concatenate([rng.uniform(-1.6, 1.6, (300, 2)), np.zeros((300, 1))], 1).astype(np.float32)
        pose = np.array([0.21, -0.13, theta])
        m = sm.CeresScanMatcher2D(sm.CeresScanMatcherOptions2D())
        for with_jac in (True, False):
            got_r, got_j = m.Evaluate(dev, cloud, pose, [0.2, -0.1], 0.25, jacobian=with_jac)
            want_r, want_j = oracle.ceres2d_evaluate(grid, cloud, pose, [0.2, -0.1], 0.25,
                                                     jacobian=with_jac)
            assert np.allclose(got_r, want_r, rtol=0, atol=1e-12)
            if with_jac:
                assert np.allclose(got_j, want_j, rtol=0, atol=1e-10)
        dev.close()


# ceres_scan_matcher_2d_test.cc:64-97 on the device, and equal to the oracle's run
@pytest.mark.parametrize("start", [(-0.5, 0.5), (-0.3, 0.5), (-0.45, 0.3), (-0.3, 0.3)])
def test_reference_known_answers_on_device(oracle, start):
    grid, cloud, opts = _ceres_test_fixture(oracle)
    dev = sm.RealTimeGrid2D(grid)
    m = sm.CeresScanMatcher2D(sm.CeresScanMatcherOptions2D(**opts))
    init = [start[0], start[1], 0.0]
    pose, summary = m.Match(init[:2], init, cloud, dev)
    assert summary["final_cost"] == pytest.approx(0.0, abs=1e-2)
    assert is_nearly(pose, (-0.5, 0.5, 0.0), 1e-2)
    _same_solution(pose, summary, oracle.ceres2d_match(grid, cloud, init[:2], init, **opts))
    dev.close()


def _floor_plan_case(oracle, seed, size_cells, beams):
    grid, occ = synthetic.make_grid2d(seed, size_cells=size_cells)
    rng = np.random.RandomState(seed)
    pose = synthetic.random_free_pose(occ, grid, rng, margin_cells=10)
    scan = synthetic.cast_scan(occ, grid, pose, beams=beams, max_range=8.0, seed=seed)
    og = oracle.Grid2D(grid.cells, grid.resolution, grid.max_x, grid.max_y)
    return grid, og, np.asarray(pose, np.float64), scan


@pytest.mark.parametrize("nonmonotonic", [True, False])
def test_match_equals_oracle_on_floor_plans(oracle, nonmonotonic):
    opts = sm.CeresScanMatcherOptions2D(use_nonmonotonic_steps=nonmonotonic, max_num_iterations=20)
    m = sm.CeresScanMatcher2D(opts)
    for seed in (3, 4, 5):
        grid, og, pose, scan = _floor_plan_case(oracle, seed, 200, 181)
        dev = sm.RealTimeGrid2D(grid)
        rng = np.random.RandomState(100 + seed)
        for _ in range(3):
            init = pose + np.array([rng.uniform(-0.06, 0.06), rng.uniform(-0.06, 0.06),
                                    rng.uniform(-0.02, 0.02)])
            got_pose, got_sum = m.Match(init[:2], init, scan, dev)
            want = oracle.ceres2d_match(og, scan, init[:2], init, opts.occupied_space_weight,
                                        opts.translation_weight, opts.rotation_weight,
                                        nonmonotonic, opts.max_num_iterations)
            _same_solution(got_pose, got_sum, want)
            assert got_sum["final_cost"] <= got_sum["initial_cost"]
        dev.close()


def test_batch_over_several_submaps_equals_single_calls(oracle):
    """One launch over matches that belong to different submaps (csm_ceres_job2d::grid)."""
    m = sm.CeresScanMatcher2D()
    cases, devs = [], []
    for seed in (6, 7, 8):
        grid, og, pose, scan = _floor_plan_case(oracle, seed, 160, 121)
        dev = sm.RealTimeGrid2D(grid)
        devs.append(dev)
        for k in range(4):
            init = pose + np.array([0.03 * (k - 1), -0.02 * k, 0.005 * k])
            cases.append((dev, og, scan, init))
    poses, sums = m.MatchBatch([c[3][:2] for c in cases], [c[3] for c in cases],
                               [c[2] for c in cases], [c[0] for c in cases])
    assert m.last_stats["host_syncs"] == 1
    for (dev, og, scan, init), p, s in zip(cases, poses, sums):
        p1, s1 = m.Match(init[:2], init, scan, dev)
        assert np.array_equal(p, p1) and s == s1          # same kernel, same block order
        _same_solution(p, s, oracle.ceres2d_match(og, scan, init[:2], init))
    for d in devs:
        d.close()


def test_baseline_size_refinement_1000x1000_1081_beams(oracle):
    """The config-4 shape: a 1081-beam node scan refined in a 1000 x 1000 submap after the
    fast matcher found it (constraint_builder_2d.cc:229-249)."""
    grid, og, pose, scan = _floor_plan_case(oracle, 21, 1000, 1081)
    opt = sm.FastCorrelativeScanMatcherOptions2D(7.0, math.radians(30.0), 7)
    fast = sm.FastCorrelativeScanMatcher2D(grid, opt)
    init = pose + np.array([0.8, -0.6, 0.1])
    found, score, est = fast.Match(init, scan, 0.55)
    assert found
    dev = sm.RealTimeGrid2D(grid)
    m = sm.CeresScanMatcher2D()
    got_pose, got_sum = m.Match(est[:2], est, scan, dev)
    _same_solution(got_pose, got_sum, oracle.ceres2d_match(og, scan, est[:2], est))
    # a refinement, not a new search: the estimate moves by less than a cell
    assert np.linalg.norm(got_pose[:2] - np.asarray(est[:2])) < grid.resolution
    assert got_sum["final_cost"] < got_sum["initial_cost"]
    dev.close()
    fast.close()


def test_constraint_builder_with_refinement_equals_oracle():
    opts, submaps, clouds, poses = _small_queue()
    opts.ceres_scan_matcher_options = sm.CeresScanMatcherOptions2D()
    ref = cb.ConstraintBuilder2D(opts, executor=OracleExecutor(opts))
    _fill(ref, submaps, clouds, poses)
    want = ref.WhenDone(lambda r: None)
    dev = cb.ConstraintBuilder2D(opts)
    _fill(dev, submaps, clouds, poses)
    got = dev.WhenDone(lambda r: None)
    assert len(want) > 0 and len(got) == len(want)
    for a, b in zip(got, want):
        assert a.submap_id == b.submap_id and a.node_id == b.node_id
        assert np.float32(a.score) == np.float32(b.score)
        assert np.allclose(a.zbar_ij, b.zbar_ij, rtol=0, atol=POSE_TOL)
    assert dev.executor.stats["refined"] == len(want)
    for sid in list(dev.executor.matchers):
        dev.DeleteScanMatcher(sid)


def test_invalid_options_are_rejected(oracle):
    grid = oracle.Grid2D(np.zeros((4, 4), np.uint16), 1.0, 2.0, 2.0)
    dev = sm.RealTimeGrid2D(grid)
    for bad in (dict(occupied_space_weight=0.0), dict(translation_weight=-1.0),
                dict(rotation_weight=0.0), dict(max_num_iterations=0)):
        m = sm.CeresScanMatcher2D(sm.CeresScanMatcherOptions2D(**bad))
        with pytest.raises(CsmError):          # CHECK_GT(..., 0.) in the reference
            m.Match([0.0, 0.0], [0.0, 0.0, 0.0], [[0.0, 0.0, 0.0]], dev)
    dev.close()
