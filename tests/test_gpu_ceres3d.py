"""Device CeresScanMatcher3D (csm_ceres_match3d_batch / csm_ceres_evaluate3d) against the
oracle's restatement (oracle/oracle_ceres3d.cc) and the reference's own known answers
(ceres_scan_matcher_3d_test.cc).  Doubles on both sides, no FMA contraction: per-point
residuals and tangent-space Jacobian rows to 1e-12 (there is no transcendental in them),
solved poses to 1e-7 (block-tree sums; device cos / sin in the quaternion update).  The
BASELINE-size cases (32 k and 64 k points) are in tests/test_gpu_baseline_sizes.py."""
import math

import numpy as np
import pytest

from benchmarks import synthetic
from cartographer_b200 import scan_matching as sm
from cartographer_b200._lib import CsmError
from tests.test_gpu_parity_3d import _building_case
from tests.test_oracle_golden_3d import is_nearly
from tests.test_oracle_golden_ceres3d import POINTS

pytestmark = pytest.mark.gpu

POSE_TOL = 1e-7


def _spec(oracle, resolution, points):
    idx = np.array([oracle.hybrid_get_cell_index(resolution, p) for p in points], np.int32)
    val = np.full(len(idx), oracle.probability_to_value(1.0), np.uint16)
    return synthetic.HybridGridSpec(resolution, idx, val)


def _ogrid(oracle, spec):
    return oracle.HybridGrid(spec.resolution, spec.indices, spec.values)


def _same_solution(got_pose, got_sum, want):
    assert np.allclose(got_pose, want["pose"], rtol=0, atol=POSE_TOL), (got_pose, want["pose"])
    assert got_sum["initial_cost"] == pytest.approx(want["initial_cost"], rel=1e-12, abs=1e-15)
    assert got_sum["final_cost"] == pytest.approx(want["final_cost"], rel=1e-9, abs=1e-15)
    assert got_sum["iterations"] == want["iterations"]
    assert got_sum["num_successful_steps"] == want["num_successful_steps"]
    assert got_sum["termination"] == want["termination"]


REFERENCE_OPTIONS = dict(occupied_space_weights=[1.0], translation_weight=0.01,
                         rotation_weight=0.1, use_nonmonotonic_steps=True, max_num_iterations=10)


def _reference_matcher():
    return sm.CeresScanMatcher3D(sm.CeresScanMatcherOptions3D(
        occupied_space_weight_0=1.0, occupied_space_weight_1=1.0, translation_weight=0.01,
        rotation_weight=0.1, use_nonmonotonic_steps=True, max_num_iterations=10))


# ceres_scan_matcher_3d_test.cc:99-116 on the device (fixture without the intensity block)
@pytest.mark.parametrize("start", [(-1.0, 0.0, 0.0), (-0.8, 0.0, 0.0), (-1.0, 0.0, -0.2),
                                   (-0.9, -0.2, 0.2)])
def test_reference_known_answers_on_device(oracle, start):
    spec = _spec(oracle, 1.0, POINTS + np.array([-1, 0, 0], np.float32))
    dev = sm.DeviceHybridGrid(spec)
    init = [start[0], start[1], start[2], 1.0, 0.0, 0.0, 0.0]
    pose, summary = _reference_matcher().Match(init[:3], init, [(POINTS, dev)])
    assert summary["final_cost"] == pytest.approx(0.0, abs=1e-2)
    assert is_nearly(pose, [-1, 0, 0, 1, 0, 0, 0], 3e-2)
    _same_solution(pose, summary, oracle.ceres3d_match([(POINTS, _ogrid(oracle, spec))], init[:3],
                                                       init, **REFERENCE_OPTIONS))
    dev.close()


# ceres_scan_matcher_3d_test.cc:118-131
def test_reference_full_pose_correction_on_device(oracle):
    spec = _spec(oracle, 1.0, POINTS + np.array([-1, 0, 0], np.float32))
    dev = sm.DeviceHybridGrid(spec)
    a = 0.05
    c, s = math.cos(a), math.sin(a)
    cloud = np.stack([np.float32(c) * POINTS[:, 0] - np.float32(s) * POINTS[:, 1],
                      np.float32(s) * POINTS[:, 0] + np.float32(c) * POINTS[:, 1],
                      POINTS[:, 2]], 1).astype(np.float32)
    expected = [-1, 0, 0, math.cos(-a / 2), 0, 0, math.sin(-a / 2)]
    init = [-0.95, -0.05, 0.05, math.cos(a / 2), math.sin(a / 2), 0.0, 0.0]
    pose, summary = _reference_matcher().Match(init[:3], init, [(cloud, dev)])
    assert summary["final_cost"] == pytest.approx(0.0, abs=1e-2)
    assert is_nearly(pose, expected, 3e-2)
    _same_solution(pose, summary, oracle.ceres3d_match([(cloud, _ogrid(oracle, spec))], init[:3],
                                                       init, **REFERENCE_OPTIONS))
    dev.close()


def test_residuals_and_jacobian_equal_the_oracle(oracle):
    hi, lo, cloud, low, _, _, node_pose = _building_case(11)
    dhi, dlo = sm.DeviceHybridGrid(hi), sm.DeviceHybridGrid(lo)
    ohi, olo = _ogrid(oracle, hi), _ogrid(oracle, lo)
    m = sm.CeresScanMatcher3D()
    a = 0.07
    q = np.array([math.cos(a), math.sin(a) * 0.2, -math.sin(a) * 0.3, math.sin(a) * 0.933])
    pose = np.concatenate([node_pose[:3] + [0.04, -0.03, 0.02], q / np.linalg.norm(q)])
    target_q = node_pose[3:]
    for with_jac in (True, False):
        got_r, got_j = m.Evaluate([(cloud, dhi), (low, dlo)], pose, node_pose[:3], target_q,
                                  jacobian=with_jac)
        want_r, want_j = oracle.ceres3d_evaluate([(cloud, ohi), (low, olo)], pose, node_pose[:3],
                                                 target_q, jacobian=with_jac)
        assert got_r.shape == want_r.shape == (len(cloud) + len(low) + 6,)
        assert np.allclose(got_r, want_r, rtol=0, atol=1e-12)
        if with_jac:
            assert np.abs(want_j).max() > 0.01
            assert np.allclose(got_j, want_j, rtol=0, atol=1e-10)
    # one cloud only (occupied_space_weight_size == 1)
    got_r, _ = m.Evaluate([(cloud, dhi)], pose, node_pose[:3], target_q)
    want_r, _ = oracle.ceres3d_evaluate([(cloud, ohi)], pose, node_pose[:3], target_q)
    assert np.allclose(got_r, want_r, rtol=0, atol=1e-12)
    dhi.close()
    dlo.close()


@pytest.mark.parametrize("seed,nonmonotonic", [(11, False), (12, False), (13, True)])
def test_match_equals_oracle_in_buildings(oracle, seed, nonmonotonic):
    hi, lo, cloud, low, _, _, node_pose = _building_case(seed)
    dhi, dlo = sm.DeviceHybridGrid(hi), sm.DeviceHybridGrid(lo)
    ohi, olo = _ogrid(oracle, hi), _ogrid(oracle, lo)
    m = sm.CeresScanMatcher3D(sm.CeresScanMatcherOptions3D(use_nonmonotonic_steps=nonmonotonic,
                                                           max_num_iterations=12))
    rng = np.random.RandomState(seed)
    for _ in range(2):
        init = node_pose.copy()
        init[:3] += rng.uniform(-0.06, 0.06, 3)
        yaw = 2 * math.atan2(node_pose[6], node_pose[3]) + rng.uniform(-0.02, 0.02)
        init[3:] = [math.cos(yaw / 2), 0, 0, math.sin(yaw / 2)]
        pose, summary = m.Match(init[:3], init, [(cloud, dhi), (low, dlo)])
        want = oracle.ceres3d_match([(cloud, ohi), (low, olo)], init[:3], init,
                                    use_nonmonotonic_steps=nonmonotonic, max_num_iterations=12)
        _same_solution(pose, summary, want)
        assert summary["final_cost"] <= summary["initial_cost"]
        assert abs(np.linalg.norm(pose[3:]) - 1.0) < 1e-9      # Plus keeps the quaternion unit
    dhi.close()
    dlo.close()


def test_batch_over_several_submaps_equals_single_calls(oracle):
    m = sm.CeresScanMatcher3D()
    cases, keep = [], []
    for seed in (11, 12):
        hi, lo, cloud, low, _, _, node_pose = _building_case(seed)
        dhi, dlo = sm.DeviceHybridGrid(hi), sm.DeviceHybridGrid(lo)
        keep += [dhi, dlo]
        for k in range(3):
            init = node_pose.copy()
            init[:3] += [0.03 * (k - 1), -0.02 * k, 0.01 * k]
            cases.append((init, [(cloud, dhi), (low, dlo)],
                          [(cloud, _ogrid(oracle, hi)), (low, _ogrid(oracle, lo))]))
    poses, sums = m.MatchBatch([c[0][:3] for c in cases], [c[0] for c in cases],
                               [c[1] for c in cases])
    assert m.last_stats["host_syncs"] == 1
    for (init, pairs, opairs), p, s in zip(cases, poses, sums):
        p1, s1 = m.Match(init[:3], init, pairs)
        assert np.array_equal(p, p1) and s == s1
        _same_solution(p, s, oracle.ceres3d_match(opairs, init[:3], init))
    for g in keep:
        g.close()


def test_constraint_builder_3d_with_refinement(oracle):
    """ConstraintBuilder3D with ceres_scan_matcher_options_3d: the constraints are the
    oracle's fast matches refined by the oracle's restatement."""
    from cartographer_b200 import constraint_builder as cb
    od = dict(branch_and_bound_depth=6, full_resolution_depth=3, min_rotational_score=0.3,
              min_low_resolution_score=0.25, linear_xy_search_window=1.5,
              linear_z_search_window=0.5, angular_search_window=0.25)
    hi, lo, cloud, low, hn, hs, node_pose = _building_case(11)
    opts = cb.ConstraintBuilderOptions3D(sampling_ratio=1.0, max_constraint_distance=50.0,
                                         min_score=0.3, global_localization_min_score=0.3,
                                         ceres_scan_matcher_options_3d=sm.CeresScanMatcherOptions3D(),
                                         **od)
    ohi, olo = _ogrid(oracle, hi), _ogrid(oracle, lo)
    om = oracle.FastCorrelativeScanMatcher3D(ohi, olo, hs, od)
    b = cb.ConstraintBuilder3D(opts)
    sub = cb.Submap3D(hi, lo, hs, om.hi.grid_size())
    node = sm.TrajectoryNodeData3D(cloud, low, hn, (1.0, 0.0, 0.0, 0.0))
    ident = [0, 0, 0, 1, 0, 0, 0]
    init = node_pose.copy()
    init[:3] += [0.5, -0.4, 0.1]
    b.MaybeAddConstraint((0, 0), sub, (0, 0), node, init, ident)
    b.NotifyEndOfNode()
    got = b.WhenDone(lambda r: None)
    nd = dict(gravity_alignment=(1.0, 0.0, 0.0, 0.0), high_resolution_point_cloud=cloud,
              low_resolution_point_cloud=low, rotational_scan_matcher_histogram=hn)
    want = om.match(init, ident, nd, 0.3)
    assert want["found"] and len(got) == 1
    refined = oracle.ceres3d_match([(cloud, ohi), (low, olo)], want["pose"][:3], want["pose"])
    assert np.float32(got[0].score) == want["score"]
    assert np.allclose(got[0].zbar_ij, refined["pose"], rtol=0, atol=POSE_TOL)
    assert b.executor.stats["refined"] == 1
    b.DeleteScanMatcher((0, 0))


def test_invalid_options_are_rejected(oracle):
    spec = _spec(oracle, 1.0, POINTS)
    dev = sm.DeviceHybridGrid(spec)
    init = [0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0]
    for bad in (dict(occupied_space_weight_0=0.0), dict(translation_weight=0.0),
                dict(rotation_weight=-1.0), dict(max_num_iterations=0),
                dict(only_optimize_yaw=True)):
        m = sm.CeresScanMatcher3D(sm.CeresScanMatcherOptions3D(**bad))
        with pytest.raises(CsmError):
            m.Match(init[:3], init, [(POINTS, dev)])
    dev.close()
