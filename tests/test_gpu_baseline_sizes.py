"""Engine vs oracle, directly, at the sizes BASELINE.json quotes its configs on
(not only through size-independent properties):

  config 1  RealTimeCSM2D, 1081 beams, 200x200 @5 cm, +-0.1 m / +-7 deg
  config 2  FastCSM2D MatchFullSubmap, 1081 beams, 1000x1000 @5 cm, depth 7
  config 3  FastCSM3D Match, 16 rings x 2048 az (~32 k points)
  config 5  FastCSM3D Match, 64 rings x 1024 az (~64 k points)

(config 4's shape — local windows on a 1000x1000 submap — is
test_gpu_parity_2d.py::test_match_batch_at_baseline_size / test_full_size_properties.)
Everything is compared for equality: score bits, winning candidate, double pose.
Reference tests these mirror: fast_correlative_scan_matcher_2d_test.cc:194-246,
fast_correlative_scan_matcher_3d_test.cc:146-204,
real_time_correlative_scan_matcher_2d_test.cc:125-160.
"""
import math
import os

import numpy as np
import pytest

from benchmarks import synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sm():
    from cartographer_b200 import scan_matching
    return scan_matching


def _ogrid(oracle, g):
    return oracle.Grid2D(g.cells, g.resolution, g.max_x, g.max_y)


def test_full_submap_1000x1000_depth7_vs_oracle(oracle, sm):
    """BASELINE config 2, the bench workload: MatchFullSubmap of 1081-beam scans on the
    1000x1000 grid, depth 7, min_score 0.6.  The oracle takes ~10-20 s per match on one
    thread; the matches run on separate host threads."""
    grid, occ = synthetic.make_grid2d(0, 1000)
    rng = np.random.RandomState(17)
    scans = []
    for i in range(4):
        pose = synthetic.random_free_pose(occ, grid, rng)
        scans.append(synthetic.cast_scan(occ, grid, pose, seed=i))
    lin, ang, depth, min_score = 7.0, math.radians(30.0), 7, 0.6
    m = sm.FastCorrelativeScanMatcher2D(grid, sm.FastCorrelativeScanMatcherOptions2D(lin, ang, depth))
    clouds = [sm.DeviceCloud(s) for s in scans]
    jobs = np.zeros(len(scans), sm.JOB2D_DTYPE)
    for k in range(len(scans)):
        jobs[k]["cloud_index"] = k
        jobs[k]["full_submap"] = 1
        jobs[k]["min_score"] = min_score
    res, st = sm.match_batch([m], clouds, jobs, lin, ang)
    assert st["host_syncs"] <= 2, st   # the level loop is device-resident
    om = oracle.FastCorrelativeScanMatcher2D(_ogrid(oracle, grid), lin, ang, depth)
    threads = min(len(scans), os.cpu_count() or 1)
    _, found, scores, poses, cs = oracle.fast2d_batch(
        [om], [0] * len(scans), list(range(len(scans))), np.zeros((len(scans), 3)), scans, True,
        min_score, threads)
    assert found.sum() >= 1
    for k in range(len(scans)):
        assert bool(res[k]["found"]) == bool(found[k]), k
        if found[k]:
            assert np.float32(res[k]["score"]) == scores[k], (k, res[k]["score"], scores[k])
            np.testing.assert_array_equal(res[k]["pose_estimate"], poses[k])
    # single-call form against the same oracle answers
    f1, s1, p1 = m.MatchFullSubmap(scans[0], min_score)
    assert bool(f1) == bool(found[0])
    if f1:
        assert np.float32(s1) == scores[0]
        np.testing.assert_array_equal(p1, poses[0])
    # both implementations score every lowest-resolution candidate
    assert st["lowest_resolution_candidates"] > 0
    m.close()


def test_rt_1081_beams_200x200_vs_oracle(oracle, sm):
    """BASELINE config 1: 1081 beams, 200x200 crop, +-0.1 m / +-7 deg, weights 0.1 / 0.1."""
    big, occ = synthetic.make_grid2d(7, 1000)
    grid, occ2 = synthetic.crop_grid(big, occ, 400, 400, 200, 200)
    rng = np.random.RandomState(1)
    opts = sm.RealTimeCorrelativeScanMatcherOptions(0.1, math.radians(7.0), 0.1, 0.1)
    rt = sm.RealTimeCorrelativeScanMatcher2D(opts)
    og = _ogrid(oracle, grid)
    for i in range(6):
        pose = synthetic.random_free_pose(occ2, grid, rng, margin_cells=15)
        scan = synthetic.cast_scan(occ2, grid, pose, seed=i, max_range=30.0)
        assert len(scan) == 1081
        init = pose + rng.uniform(-1, 1, 3) * [0.05, 0.05, math.radians(3)]
        score, est = rt.Match(init, scan, grid)
        want = oracle.rt2d_match(og, scan, init, 0.1, math.radians(7.0), 0.1, 0.1)
        assert np.float32(score) == np.float32(want["score"])
        np.testing.assert_array_equal(est, want["pose"])
        assert rt.last_stats["candidates_scored"] == want["candidates_scored"]


def _config3d(sm, oracle, rings, az, n_nodes, seed):
    o3 = sm.FastCorrelativeScanMatcherOptions3D(min_rotational_score=0.45)
    od = dict(branch_and_bound_depth=8, full_resolution_depth=3, min_rotational_score=0.45,
              min_low_resolution_score=0.55, linear_xy_search_window=5.0,
              linear_z_search_window=1.0, angular_search_window=math.radians(15.0))
    hi, lo, sub_hist, world = synthetic.make_submap3d(seed, 40.0, rings, az, 20.0)
    rng = np.random.RandomState(500 + seed)
    nodes = [synthetic.make_node3d(world, rng, rings, az, 20.0, seed=7000 + k)
             for k in range(n_nodes)]
    m = sm.FastCorrelativeScanMatcher3D(hi, lo, sub_hist, o3)
    ohi = oracle.HybridGrid(hi.resolution, hi.indices, hi.values)
    olo = oracle.HybridGrid(lo.resolution, lo.indices, lo.values)
    om = oracle.FastCorrelativeScanMatcher3D(ohi, olo, sub_hist, od)
    ident = [0, 0, 0, 1, 0, 0, 0]
    rng = np.random.RandomState(9)
    n_found = 0
    # the post-match refinement at the same size (constraint_builder_3d.cc:265-275)
    dhi, dlo = sm.DeviceHybridGrid(hi), sm.DeviceHybridGrid(lo)
    ceres = sm.CeresScanMatcher3D()
    for n in nodes:
        init = n["pose"].copy()
        init[:3] += rng.uniform(-1, 1, 3) * [2.0, 2.0, 0.3]
        yaw = 2 * math.atan2(n["pose"][6], n["pose"][3]) + rng.uniform(-1, 1) * math.radians(8)
        init[3:] = [math.cos(yaw / 2), 0, 0, math.sin(yaw / 2)]
        got, gst = m.match_raw(False, init, ident,
                               sm.TrajectoryNodeData3D(n["cloud"], n["low"], n["hist"]), 0.55)
        node = dict(gravity_alignment=(1.0, 0.0, 0.0, 0.0), high_resolution_point_cloud=n["cloud"],
                    low_resolution_point_cloud=n["low"], rotational_scan_matcher_histogram=n["hist"])
        want = om.match(init, ident, node, 0.55)
        assert (got is not None) == want["found"]
        if want["found"]:
            n_found += 1
            assert got["score"] == want["score"]
            np.testing.assert_array_equal(got["pose_estimate"], want["pose"])
            assert got["rotational_score"] == want["rotational_score"]
            assert got["low_resolution_score"] == want["low_resolution_score"]
            est = got["pose_estimate"]
            rp, rs = ceres.Match(est[:3], est, [(n["cloud"], dhi), (n["low"], dlo)])
            rw = oracle.ceres3d_match([(n["cloud"], ohi), (n["low"], olo)], est[:3], est)
            assert np.allclose(rp, rw["pose"], rtol=0, atol=1e-7), (rp, rw["pose"])
            assert rs["iterations"] == rw["iterations"]
            assert rs["num_successful_steps"] == rw["num_successful_steps"]
            assert rs["termination"] == rw["termination"]
            assert rs["final_cost"] == pytest.approx(rw["final_cost"], rel=1e-9)
            assert rs["final_cost"] <= rs["initial_cost"]
    dhi.close()
    dlo.close()
    m.close()
    return n_found, len(nodes[0]["cloud"])


def test_fast3d_config3_cloud_vs_oracle(oracle, sm):
    """BASELINE config 3: 16 rings x 2048 az (~32 k points) vs 10 cm / 45 cm hybrid grids,
    depth 8 / full-resolution depth 3, 5 m / 1 m / 15 deg."""
    n_found, npts = _config3d(sm, oracle, 16, 2048, 2, 40)
    assert npts > 20000
    assert n_found >= 1


def test_fast3d_config5_node_vs_oracle(oracle, sm):
    """BASELINE config 5: 64 rings x 1024 az (~64 k points)."""
    n_found, npts = _config3d(sm, oracle, 64, 1024, 1, 41)
    assert npts > 40000
