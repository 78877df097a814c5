"""The committed golden vectors (tests/golden/golden_v1.npz, written by
tests/golden/make_golden.py) against a live run of the CPU oracle: the oracle cannot
drift without this failing.  CPU only."""
import os
import zlib

import numpy as np

from tests import worlds3d

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")


def _bits(x):
    return int(np.array([np.float32(x)]).view(np.uint32)[0])


def test_golden_file_is_self_contained():
    d = np.load(GOLDEN)
    names = [str(n) for n in d["names"]]
    assert len(names) >= 6 and "rt2d/score" in d.files
    for n in names:
        assert n + "/found" in d.files and n + "/score_bits" in d.files and n + "/pose" in d.files


def test_oracle_reproduces_golden_2d(oracle):
    d = np.load(GOLDEN)
    for name in [str(n) for n in d["names"] if str(n).startswith("fast2d")]:
        res, mx, my = d[name + "/limits"]
        full, lin, ang, depth, min_score = d[name + "/opts"]
        og = oracle.Grid2D(d[name + "/cells"], res, mx, my)
        om = oracle.FastCorrelativeScanMatcher2D(og, lin, ang, int(depth))
        cloud = d[name + "/cloud"]
        w = om.match_full_submap(cloud, min_score) if full else om.match(d[name + "/init"], cloud,
                                                                         min_score)
        assert int(w["found"]) == int(d[name + "/found"][0])
        assert _bits(w["score"]) == int(d[name + "/score_bits"][0])
        np.testing.assert_array_equal(w["pose"], d[name + "/pose"])
        assert [w["best_scan_index"], w["best_x_offset"], w["best_y_offset"]] == \
            d[name + "/best"].tolist()
        assert w["lowest_resolution_candidates"] == int(d[name + "/lowest_resolution_candidates"][0])
        crc = [zlib.crc32(om.level(l).tobytes()) for l in range(int(depth))]
        assert crc == d[name + "/level_crc32"].tolist()


def test_oracle_reproduces_golden_rt2d(oracle):
    d = np.load(GOLDEN)
    res, mx, my = d["rt2d/limits"]
    lin, ang, wt, wr = d["rt2d/opts"]
    og = oracle.Grid2D(d["rt2d/cells"], res, mx, my)
    w = oracle.rt2d_match(og, d["rt2d/cloud"], d["rt2d/init"], lin, ang, wt, wr)
    assert w["score"] == float(d["rt2d/score"][0])
    np.testing.assert_array_equal(w["pose"], d["rt2d/pose"])


def test_oracle_reproduces_golden_3d(oracle):
    d = np.load(GOLDEN)
    ident = [0, 0, 0, 1, 0, 0, 0]
    for name in [str(n) for n in d["names"] if str(n).startswith("fast3d")]:
        og = oracle.HybridGrid(float(d[name + "/resolution"][0]), d[name + "/indices"],
                               d[name + "/values"])
        om = oracle.FastCorrelativeScanMatcher3D(og, og, np.zeros(10, np.float32),
                                                 worlds3d.TEST_OPTIONS)
        w = om.match(ident, ident, worlds3d.node_data(d[name + "/cloud"]), 0.1)
        assert int(w["found"]) == int(d[name + "/found"][0])
        assert _bits(w["score"]) == int(d[name + "/score_bits"][0])
        np.testing.assert_array_equal(w["pose"], d[name + "/pose"])
        assert [w["best_scan_index"], w["best_x"], w["best_y"], w["best_z"]] == \
            d[name + "/best"].tolist()
        assert [_bits(w["rotational_score"]), _bits(w["low_resolution_score"])] == \
            d[name + "/gate"].tolist()
