"""RealTimeCorrelativeScanMatcher3D on the device vs the oracle: score bits and the
7-double pose compared for equality.  Reference:
real_time_correlative_scan_matcher_3d.cc:34-117, real_time_correlative_scan_matcher_3d_test.cc."""
import math

import numpy as np
import pytest

from benchmarks import synthetic
from tests.test_oracle_golden_3d import _rt3d_fixture, is_nearly, rt3d_reference_initial_poses

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sm():
    from cartographer_b200 import scan_matching
    return scan_matching


def test_rt3d_reference_tests_on_device(oracle, sm):
    """The reference's seven test cases: the device result passes the reference's own
    criterion and equals the oracle bit for bit."""
    pts, ogrid, idx, val = _rt3d_fixture(oracle)
    dg = sm.DeviceHybridGrid(synthetic.HybridGridSpec(0.1, idx, val))
    rt = sm.RealTimeCorrelativeScanMatcher3D(
        sm.RealTimeCorrelativeScanMatcherOptions(0.3, math.radians(1.0), 1e-1, 1.0))
    for init in rt3d_reference_initial_poses():
        score, pose = rt.Match(init, pts, dg)
        want = oracle.rt3d_match(ogrid, pts, init, 0.3, math.radians(1.0), 1e-1, 1.0)
        assert is_nearly(pose, [-1, 0, 0, 1, 0, 0, 0], 1e-3)
        assert np.float32(score) == want["score"]
        np.testing.assert_array_equal(pose, want["pose"])
        assert rt.last_stats["candidates_scored"] == want["candidates_scored"]
    dg.close()


def test_rt3d_building_scan_parity(oracle, sm):
    """A lidar scan inside the synthetic building against its 10 cm hybrid grid, default
    local-SLAM window (trajectory_builder_3d.lua: 0.15 m / 1 deg, weights 1e-1 / 1e-1):
    125 translations x (2A+1)^3 rotations, a few thousand points."""
    hi, lo, hist, world = synthetic.make_submap3d(44, 40.0, 16, 256, 20.0)
    rng = np.random.RandomState(3)
    node = synthetic.make_node3d(world, rng, 16, 256, 20.0, seed=9)
    cloud = node["cloud"][::3]
    ogrid = oracle.HybridGrid(hi.resolution, hi.indices, hi.values)
    dg = sm.DeviceHybridGrid(hi)
    rt = sm.RealTimeCorrelativeScanMatcher3D(
        sm.RealTimeCorrelativeScanMatcherOptions(0.15, math.radians(1.0), 1e-1, 1e-1))
    for k in range(2):
        init = node["pose"].copy()
        init[:3] += rng.uniform(-1, 1, 3) * [0.1, 0.1, 0.05]
        yaw = 2 * math.atan2(init[6], init[3]) + rng.uniform(-1, 1) * math.radians(0.5)
        init[3:] = [math.cos(yaw / 2), 0, 0, math.sin(yaw / 2)]
        score, pose = rt.Match(init, cloud, dg)
        want = oracle.rt3d_match(ogrid, cloud, init, 0.15, math.radians(1.0), 1e-1, 1e-1)
        assert np.float32(score) == want["score"], (score, want["score"])
        np.testing.assert_array_equal(pose, want["pose"])
        assert rt.last_stats["candidates_scored"] == want["candidates_scored"]
    dg.close()
