"""GPU parity tests of the 3D path (FastCorrelativeScanMatcher3D) against the CPU
oracle: precomputation pyramid (incl. half-resolution depths), rotational scores,
discretised cells, and full Match / MatchFullSubmap results incl. the
low-resolution gate.  Integers and float scores are compared for equality."""
import math

import numpy as np
import pytest

from benchmarks import synthetic
from tests import worlds3d

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sm():
    from cartographer_b200 import scan_matching
    return scan_matching


def _ogrid(oracle, spec):
    return oracle.HybridGrid(spec.resolution, spec.indices, spec.values)


def _opts(sm, d):
    return sm.FastCorrelativeScanMatcherOptions3D(
        d["branch_and_bound_depth"], d["full_resolution_depth"], d["min_rotational_score"],
        d["min_low_resolution_score"], d["linear_xy_search_window"], d["linear_z_search_window"],
        d["angular_search_window"])


def _node(sm, d):
    return sm.TrajectoryNodeData3D(d["high_resolution_point_cloud"],
                                   d["low_resolution_point_cloud"],
                                   d["rotational_scan_matcher_histogram"], d["gravity_alignment"])


def _check(want, got, stats):
    assert (got is not None) == want["found"]
    if not want["found"]:
        return
    assert got["score"] == want["score"]
    assert (got["best_scan_index"],) + got["best_offset"] == (
        want["best_scan_index"], want["best_x"], want["best_y"], want["best_z"])
    np.testing.assert_array_equal(got["pose_estimate"], want["pose"])
    assert got["rotational_score"] == want["rotational_score"]
    assert got["low_resolution_score"] == want["low_resolution_score"]


# ---- K5: PrecomputationGridStack3D -------------------------------------------------
@pytest.mark.parametrize("depth,frd,seed", [(4, 4, 1), (6, 3, 2), (8, 3, 3), (5, 1, 4)])
def test_precomputation_stack_3d_bit_exact(oracle, sm, depth, frd, seed):
    rng = np.random.RandomState(seed)
    idx = np.unique(rng.randint(-40, 40, size=(3000, 3)).astype(np.int32), axis=0)
    vals = rng.randint(1, 32768, size=len(idx)).astype(np.uint16)
    spec = synthetic.HybridGridSpec(0.1, idx, vals)
    od = dict(worlds3d.TEST_OPTIONS, branch_and_bound_depth=depth, full_resolution_depth=frd)
    om = oracle.FastCorrelativeScanMatcher3D(_ogrid(oracle, spec), _ogrid(oracle, spec),
                                             np.zeros(10, np.float32), od)
    m = sm.FastCorrelativeScanMatcher3D(spec, spec, np.zeros(10, np.float32), _opts(sm, od))
    for d in range(depth):
        lo, want = om.level(d)
        # compare over a box that covers both the oracle's and the engine's bounding box
        elo, evol = m.precomputation_grid(d)
        blo = np.minimum(lo, elo) - 2
        bhi = np.maximum(lo + np.array(want.shape[::-1]), elo + np.array(evol.shape[::-1])) + 2
        dims = (bhi - blo).astype(np.int32)
        w = om.level_box(d, blo, dims)
        _, g = m.precomputation_grid(d, blo, dims)
        np.testing.assert_array_equal(g, w)
        assert w.any()
    m.close()


# ---- K7: rotational scan matcher ---------------------------------------------------
def test_rotational_match_bit_exact(oracle, sm):
    rng = np.random.RandomState(0)
    a = rng.uniform(0, 5, 120).astype(np.float32)
    b = rng.uniform(0, 5, 120).astype(np.float32)
    angles = np.linspace(-3.1, 3.1, 257).astype(np.float32)
    want = oracle.rotational_match(a, b, 0.37, angles)
    got = sm.rotational_match(a, b, 0.37, angles)
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    # the reference's own identities (rotational_scan_matcher_test.cc:28-36)
    h = np.array([1.0, 43.0, 0.5, 0.3123, 23.0, 42.0, 0.0], np.float32)
    s = sm.rotational_match(h, h, 0.0, [0.0, 1.0])
    assert abs(s[0] - 1.0) < 1e-6 and s[1] < 1.0


def _building_case(seed, size_m=12.0, rings=8, az=256):
    occ, cell, origin = synthetic.make_building(seed, size_m=size_m, height_m=3.0, cell=0.1)
    pts = synthetic.building_surface_points(occ, cell, origin)
    hi = synthetic.grid_from_points(pts, 0.1, seed)
    lo = synthetic.grid_from_points(pts, 0.45, seed + 1)
    rng = np.random.RandomState(seed)
    for _ in range(100):
        pose = np.array([rng.uniform(-size_m / 4, size_m / 4), rng.uniform(-size_m / 4, size_m / 4),
                         1.2, rng.uniform(-math.pi, math.pi)])
        c = np.floor((pose[:3] - origin) / cell).astype(int)
        if not occ[c[2] - 2:c[2] + 3, c[1] - 3:c[1] + 4, c[0] - 3:c[0] + 4].any():
            break
    cloud = synthetic.cast_lidar_3d(occ, cell, origin, pose, rings=rings, azimuths=az,
                                    max_range=8.0, seed=seed)
    low = synthetic.voxel_downsample(cloud, 0.45)
    hist_node = synthetic.rotational_histogram(cloud, 24)
    world = worlds3d.transform_points(
        np.array([pose[0], pose[1], pose[2], math.cos(pose[3] / 2), 0, 0, math.sin(pose[3] / 2)]),
        cloud)
    hist_submap = synthetic.rotational_histogram(world, 24)
    node_pose = np.array([pose[0], pose[1], pose[2], math.cos(pose[3] / 2), 0, 0,
                          math.sin(pose[3] / 2)])
    return hi, lo, cloud, low, hist_node, hist_submap, node_pose


# ---- discretisation -----------------------------------------------------------------
@pytest.mark.parametrize("seed,full", [(1, False), (2, True)])
def test_discretize_3d_bit_exact(oracle, sm, seed, full):
    hi, lo, cloud, low, hn, hs, node_pose = _building_case(seed)
    od = dict(branch_and_bound_depth=5, full_resolution_depth=3, min_rotational_score=0.3,
              min_low_resolution_score=0.3, linear_xy_search_window=1.0,
              linear_z_search_window=0.4, angular_search_window=0.2)
    node = dict(gravity_alignment=(1.0, 0.0, 0.0, 0.0), high_resolution_point_cloud=cloud,
                low_resolution_point_cloud=low, rotational_scan_matcher_histogram=hn)
    init = node_pose + np.array([0.3, -0.2, 0.05, 0, 0, 0, 0])
    sub = [0.1, -0.1, 0.0, math.cos(0.05), 0, 0, math.sin(0.05)]
    om = oracle.FastCorrelativeScanMatcher3D(_ogrid(oracle, hi), _ogrid(oracle, lo), hs, od)
    m = sm.FastCorrelativeScanMatcher3D(hi, lo, hs, _opts(sm, od),
                                        grid_size_in_voxels=om.hi.grid_size())
    wc, wp, wr = om.discrete_scans(full, init, sub, node)
    gc, gp, gr = m.discretize(full, init, sub, _node(sm, node))
    assert gc.shape == wc.shape and len(wc) > 0
    np.testing.assert_array_equal(gr.view(np.uint32), wr.view(np.uint32))
    np.testing.assert_array_equal(gp.view(np.uint32), wp.view(np.uint32))
    np.testing.assert_array_equal(gc, wc)
    m.close()


# ---- full matches --------------------------------------------------------------------
def test_reference_3d_tests_on_device(oracle, sm):
    """fast_correlative_scan_matcher_3d_test.cc:146-204 on the engine, compared to the
    oracle bit for bit (score, offsets, pose, rotational and low-resolution scores),
    including the low-resolution rejection."""
    rng = np.random.RandomState(42)
    ident = [0, 0, 0, 1, 0, 0, 0]
    for trial in range(10):
        expected = worlds3d.random_pose(rng)
        og = worlds3d.insert_cloud(oracle, 0.05, expected)
        om = oracle.FastCorrelativeScanMatcher3D(og, og, np.zeros(10, np.float32),
                                                 worlds3d.TEST_OPTIONS)
        m = sm.FastCorrelativeScanMatcher3D(og.spec, og.spec, np.zeros(10, np.float32),
                                            _opts(sm, worlds3d.TEST_OPTIONS),
                                            grid_size_in_voxels=og.grid_size())
        node = worlds3d.node_data(worlds3d.AXIS_CLOUD)
        _check(om.match(ident, ident, node, 0.1), m.Match(ident, ident, _node(sm, node), 0.1),
               m.last_stats)
        far = worlds3d.node_data(worlds3d.AXIS_CLOUD,
                                 low=np.array([[42.0, 42.0, 42.0]], np.float32))
        assert not om.match(ident, ident, far, 0.1)["found"]
        assert m.Match(ident, ident, _node(sm, far), 0.1) is None
        if trial == 0:
            _check(om.match_full_submap([1, 0, 0, 0], [1, 0, 0, 0], node, 0.1),
                   m.MatchFullSubmap([1, 0, 0, 0], [1, 0, 0, 0], _node(sm, node), 0.1),
                   m.last_stats)
        m.close()


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_match_3d_building_parity(oracle, sm, seed):
    hi, lo, cloud, low, hn, hs, node_pose = _building_case(seed)
    od = dict(branch_and_bound_depth=6, full_resolution_depth=3, min_rotational_score=0.3,
              min_low_resolution_score=0.25, linear_xy_search_window=1.5,
              linear_z_search_window=0.5, angular_search_window=0.25)
    node = dict(gravity_alignment=(1.0, 0.0, 0.0, 0.0), high_resolution_point_cloud=cloud,
                low_resolution_point_cloud=low, rotational_scan_matcher_histogram=hn)
    rng = np.random.RandomState(seed)
    init = node_pose.copy()
    init[:3] += rng.uniform(-0.8, 0.8, 3) * [1, 1, 0.3]
    yaw = 2 * math.atan2(node_pose[6], node_pose[3]) + rng.uniform(-0.15, 0.15)
    init[3:] = [math.cos(yaw / 2), 0, 0, math.sin(yaw / 2)]
    sub = [0, 0, 0, 1, 0, 0, 0]
    om = oracle.FastCorrelativeScanMatcher3D(_ogrid(oracle, hi), _ogrid(oracle, lo), hs, od)
    m = sm.FastCorrelativeScanMatcher3D(hi, lo, hs, _opts(sm, od),
                                        grid_size_in_voxels=om.hi.grid_size())
    want = om.match(init, sub, node, 0.3)
    got = m.Match(init, sub, _node(sm, node), 0.3)
    _check(want, got, m.last_stats)
    assert want["found"]
    assert m.last_stats["lowest_resolution_candidates"] == want["lowest_resolution_candidates"]
    m.close()


def test_constraint_builder_3d_on_device(oracle, sm):
    """constraint_builder_3d_test.cc:61-: MaybeAddConstraint x2 + MaybeAddGlobalConstraint
    per round produce INTER_SUBMAP constraints whose poses equal the oracle's matches."""
    from cartographer_b200 import constraint_builder as cb
    rng = np.random.RandomState(42)
    expected = worlds3d.random_pose(rng)
    og = worlds3d.insert_cloud(oracle, 0.05, expected)
    opts = cb.ConstraintBuilderOptions3D(sampling_ratio=1.0, max_constraint_distance=50.0,
                                         min_score=0.1, global_localization_min_score=0.1,
                                         **{k: v for k, v in worlds3d.TEST_OPTIONS.items()})
    b = cb.ConstraintBuilder3D(opts)
    sub = cb.Submap3D(og.spec, og.spec, np.zeros(10, np.float32), og.grid_size())
    node = _node(sm, worlds3d.node_data(worlds3d.AXIS_CLOUD))
    ident = [0, 0, 0, 1, 0, 0, 0]
    b.MaybeAddConstraint((0, 0), sub, (0, 0), node, ident, ident)
    b.MaybeAddConstraint((0, 0), sub, (0, 1), node, ident, ident)
    b.MaybeAddGlobalConstraint((0, 0), sub, (0, 2), node, [1, 0, 0, 0], [1, 0, 0, 0])
    b.NotifyEndOfNode()
    got = b.WhenDone(lambda r: None)
    assert len(got) == 3 and b.GetNumFinishedNodes() == 1
    om = oracle.FastCorrelativeScanMatcher3D(og, og, np.zeros(10, np.float32),
                                             worlds3d.TEST_OPTIONS)
    want = om.match(ident, ident, worlds3d.node_data(worlds3d.AXIS_CLOUD), 0.1)
    for c in got[:2]:
        assert c.tag == "INTER_SUBMAP" and np.float32(c.score) == want["score"]
        np.testing.assert_array_equal(np.array(c.zbar_ij), want["pose"])
    b.DeleteScanMatcher((0, 0))


def test_match_3d_batch_equals_single_calls(sm):
    """csm_match3d_batch (the ConstraintBuilder3D queue in one call, several matches in
    flight on separate streams) returns exactly what csm_match3d returns per job."""
    from cartographer_b200._lib import CsmError
    od = dict(branch_and_bound_depth=6, full_resolution_depth=3, min_rotational_score=0.3,
              min_low_resolution_score=0.25, linear_xy_search_window=1.5,
              linear_z_search_window=0.5, angular_search_window=0.25)
    matchers, nodes, poses = [], [], []
    for seed in (11, 12):
        hi, lo, cloud, low, hn, hs, node_pose = _building_case(seed)
        matchers.append(sm.FastCorrelativeScanMatcher3D(hi, lo, hs, _opts(sm, od)))
        nodes.append(sm.TrajectoryNodeData3D(cloud, low, hn, (1.0, 0.0, 0.0, 0.0)))
        poses.append(node_pose)
    sub = [0, 0, 0, 1, 0, 0, 0]
    jobs = []
    for mi in range(2):
        for ni in range(2):
            for rep in range(3):   # more jobs than distinct pairs: lanes get reused
                init = poses[ni].copy()
                init[0] += 0.1 * rep
                jobs.append((mi, ni, False, init, sub, 0.3))
    jobs.append((0, 0, True, [0, 0, 0] + list(poses[0][3:]), sub, 0.3))
    got, st = sm.match_batch3d(matchers, nodes, jobs, max_concurrency=4)
    assert len(got) == len(jobs) and st["candidates_scored"] > 0
    found = 0
    for (mi, ni, full, npose, spose, ms), g in zip(jobs, got):
        want, _ = matchers[mi].match_raw(full, npose, spose, nodes[ni], ms)
        assert (g is None) == (want is None)
        if g is not None:
            found += 1
            assert g["score"] == want["score"] and g["best_offset"] == want["best_offset"]
            assert g["best_scan_index"] == want["best_scan_index"]
            np.testing.assert_array_equal(g["pose_estimate"], want["pose_estimate"])
            assert g["low_resolution_score"] == want["low_resolution_score"]
    assert found >= 3
    assert sm.match_batch3d(matchers, nodes, [])[0] == []
    with pytest.raises(CsmError):
        sm.match_batch3d(matchers, nodes, [(5, 0, False, poses[0], sub, 0.3)])
    for m in matchers:
        m.close()
