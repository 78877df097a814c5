"""The CUDA engine against the COMMITTED golden vectors (tests/golden/golden_v1.npz):
no oracle run is involved — inputs and expected outputs both come from the file."""
import os
import zlib

import numpy as np
import pytest

from benchmarks import synthetic
from tests import worlds3d

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")


def _bits(x):
    return int(np.array([np.float32(x)]).view(np.uint32)[0])


def test_engine_reproduces_golden_2d():
    from cartographer_b200 import scan_matching as sm
    d = np.load(GOLDEN)
    for name in [str(n) for n in d["names"] if str(n).startswith("fast2d")]:
        res, mx, my = d[name + "/limits"]
        full, lin, ang, depth, min_score = d[name + "/opts"]
        grid = synthetic.GridSpec(d[name + "/cells"], res, mx, my)
        m = sm.FastCorrelativeScanMatcher2D(
            grid, sm.FastCorrelativeScanMatcherOptions2D(lin, ang, int(depth)))
        cloud = d[name + "/cloud"]
        if full:
            found, score, pose = m.MatchFullSubmap(cloud, min_score)
        else:
            found, score, pose = m.Match(d[name + "/init"], cloud, min_score)
        assert int(found) == int(d[name + "/found"][0])
        assert _bits(score) == int(d[name + "/score_bits"][0])
        np.testing.assert_array_equal(pose, d[name + "/pose"])
        st = m.last_stats
        assert [st["best_scan_index"], st["best_x_offset"], st["best_y_offset"]] == \
            d[name + "/best"].tolist()
        assert st["lowest_resolution_candidates"] == int(d[name + "/lowest_resolution_candidates"][0])
        crc = [zlib.crc32(m.precomputation_grid(l).tobytes()) for l in range(int(depth))]
        assert crc == d[name + "/level_crc32"].tolist()
        m.close()


def test_engine_reproduces_golden_rt2d():
    from cartographer_b200 import scan_matching as sm
    d = np.load(GOLDEN)
    res, mx, my = d["rt2d/limits"]
    lin, ang, wt, wr = d["rt2d/opts"]
    grid = synthetic.GridSpec(d["rt2d/cells"], res, mx, my)
    rt = sm.RealTimeCorrelativeScanMatcher2D(
        sm.RealTimeCorrelativeScanMatcherOptions(lin, ang, wt, wr))
    score, pose = rt.Match(d["rt2d/init"], d["rt2d/cloud"], grid)
    assert score == float(d["rt2d/score"][0])
    np.testing.assert_array_equal(pose, d["rt2d/pose"])


def test_engine_reproduces_golden_3d():
    from cartographer_b200 import scan_matching as sm
    d = np.load(GOLDEN)
    ident = [0, 0, 0, 1, 0, 0, 0]
    o = worlds3d.TEST_OPTIONS
    opts = sm.FastCorrelativeScanMatcherOptions3D(
        o["branch_and_bound_depth"], o["full_resolution_depth"], o["min_rotational_score"],
        o["min_low_resolution_score"], o["linear_xy_search_window"],
        o["linear_z_search_window"], o["angular_search_window"])
    for name in [str(n) for n in d["names"] if str(n).startswith("fast3d")]:
        spec = synthetic.HybridGridSpec(float(d[name + "/resolution"][0]), d[name + "/indices"],
                                        d[name + "/values"])
        m = sm.FastCorrelativeScanMatcher3D(spec, spec, np.zeros(10, np.float32), opts,
                                            grid_size_in_voxels=int(d[name + "/grid_size"][0]))
        node = sm.TrajectoryNodeData3D(d[name + "/cloud"], d[name + "/cloud"],
                                       np.zeros(10, np.float32), (1.0, 0.0, 0.0, 0.0))
        got = m.Match(ident, ident, node, 0.1)
        assert (got is not None) == bool(d[name + "/found"][0])
        assert _bits(got["score"]) == int(d[name + "/score_bits"][0])
        np.testing.assert_array_equal(got["pose_estimate"], d[name + "/pose"])
        assert [got["best_scan_index"]] + list(got["best_offset"]) == d[name + "/best"].tolist()
        assert [_bits(got["rotational_score"]), _bits(got["low_resolution_score"])] == \
            d[name + "/gate"].tolist()
        m.close()
