"""Runs the C++ drop-in adapter (reference-shaped FastCorrelativeScanMatcher2D/3D,
RealTime...2D/3D and ConstraintBuilder2D/3D classes over the C ABI) on the device and
compares what the C++ classes returned — score bit patterns, poses to the last digit —
with the oracle on the same inputs (rebuilt here from the self-test's formulas)."""
import math
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "cartographer_b200", "adapter")


def _f(bits):
    return np.float32(struct.unpack("<f", struct.pack("<I", int(bits, 16)))[0])


def lround(x):
    """std::lround: halves away from zero."""
    x = float(x)
    return int(math.copysign(math.floor(abs(x) + 0.5), x))


def _selftest_inputs():
    """The grid and cloud of adapter_selftest.cc (an L-shaped wall of value 5000)."""
    n = 120
    cells = np.zeros((n, n), np.uint16)
    cloud = []
    for i in range(20, 100):
        cells[40, i] = 5000
        cells[i, 30] = 5000
    for i in range(20, 100, 2):
        cloud.append([np.float32(3.0 - (40 + 0.5) * 0.05), np.float32(3.0 - (i + 0.5) * 0.05), 0.0])
        cloud.append([np.float32(3.0 - (i + 0.5) * 0.05), np.float32(3.0 - (30 + 0.5) * 0.05), 0.0])
    return cells, np.array(cloud, np.float32)


@pytest.mark.gpu
def test_adapter_selftest_runs_on_device(oracle):
    exe = os.path.join(ADAPTER, "adapter_selftest")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", ADAPTER, "-s"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(out.stdout, out.stderr)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "link check only" not in out.stdout, "no CUDA device seen by the adapter"
    assert "ConstraintBuilder2D 1 callback(s), 2 constraints" in out.stdout
    assert "ConstraintBuilder3D" in out.stdout
    res = {}
    for ln in out.stdout.splitlines():
        if ln.startswith("RESULT "):
            parts = ln.split()
            res.setdefault(parts[1], []).append(parts[2:])

    cells, cloud = _selftest_inputs()
    og = oracle.Grid2D(cells, 0.05, 3.0, 3.0, np.float32(0.1), np.float32(0.9))
    om = oracle.FastCorrelativeScanMatcher2D(og, 1.0, 0.3, 4)
    # FastCorrelativeScanMatcher2D::Match / MatchFullSubmap through the C++ class
    want = om.match([0.2, -0.15, 0.05], cloud, 0.5)
    got = res["fast2d"][0]
    assert int(got[0]) == int(want["found"])
    assert _f(got[1]) == want["score"]
    np.testing.assert_array_equal([float(v) for v in got[2:5]], want["pose"])
    want = om.match_full_submap(cloud, 0.5)
    got = res["fast2d_full"][0]
    assert int(got[0]) == int(want["found"])
    if want["found"]:
        assert _f(got[1]) == want["score"]
        np.testing.assert_array_equal([float(v) for v in got[2:5]], want["pose"])
    # RealTimeCorrelativeScanMatcher2D::Match and ::ScoreCandidates
    ogrt = oracle.Grid2D(cells, 0.05, 3.0, 3.0)
    want = oracle.rt2d_match(ogrt, cloud, [0.06, -0.04, 0.02], 0.1, 0.1, 0.1, 0.1)
    got = res["rt2d"][0]
    assert _f(got[0]) == np.float32(want["score"])
    np.testing.assert_array_equal([float(v) for v in got[1:4]], want["pose"])
    scan = []
    for i in range(20, 100, 2):
        scan += [[i, 40], [30, i]]
    want = oracle.rt2d_score_candidates(ogrt, np.array([scan], np.int32), 2, 0, 0.01, 0.1, 0.1,
                                        np.array([[0, 0, 0], [0, 1, 0], [0, -2, 2]], np.int32))
    got = res["rt2d_candidates"][0]
    assert [_f(b) for b in got] == list(want)
    # ConstraintBuilder2D: zbar_ij = submap_pose^-1 * pose_estimate of the two found pairs
    sp = (0.1, -0.05, 0.02)

    def compose(a, b):
        c, s = math.cos(a[2]), math.sin(a[2])
        return (c * b[0] - s * b[1] + a[0], s * b[0] + c * b[1] + a[1], a[2] + b[2])

    def inverse(a):
        c, s = math.cos(-a[2]), math.sin(-a[2])
        return (-(c * a[0] - s * a[1]), -(s * a[0] + c * a[1]), -a[2])
    init = compose(sp, compose(inverse(sp), (0.2, -0.15, 0.05)))
    w_local = om.match(init, cloud, 0.5)
    w_global = om.match_full_submap(cloud, 0.4)
    assert w_local["found"] and w_global["found"]
    for got, want in zip(res["cb2d"], (w_local, w_global)):
        z = compose(inverse(sp), tuple(want["pose"]))
        np.testing.assert_allclose([float(v) for v in got[1:3]], z[:2], rtol=0, atol=1e-12)
        assert abs(float(got[3]) - z[2]) < 1e-12   # yaw went through a quaternion (Embed3D)
    # the same queue with the device refinement (constraint_builder_2d.cc:245-249), and
    # CeresScanMatcher2D::Match through the C++ class; doubles, tolerance as in
    # tests/test_gpu_ceres2d.py
    for got, want in zip(res["cb2d_refined"], (w_local, w_global)):
        r = oracle.ceres2d_match(ogrt, cloud, want["pose"][:2], want["pose"])
        z = compose(inverse(sp), tuple(r["pose"]))
        np.testing.assert_allclose([float(v) for v in got[1:4]], z, rtol=0, atol=1e-7)
    got = res["ceres2d"][0]
    want = oracle.ceres2d_match(ogrt, cloud, [0.03, -0.02], [0.03, -0.02, 0.01])
    np.testing.assert_allclose([float(v) for v in got[0:3]], want["pose"], rtol=0, atol=1e-7)
    assert float(got[3]) == pytest.approx(want["initial_cost"], rel=1e-12)
    assert float(got[4]) == pytest.approx(want["final_cost"], rel=1e-9)
    assert [int(got[5]), int(got[6])] == [want["iterations"], want["num_successful_steps"]]
    assert oracle.CERES_TERMINATION[int(got[7])] == want["termination"]
    # FastCorrelativeScanMatcher3D::Match
    idx, val, cloud3 = [], [], []
    tx, ty, tz = np.float32(0.2), np.float32(-0.15), np.float32(0.1)
    for axis in range(3):
        d = np.float32(4.0)
        while d <= 5.5:
            p = [np.float32(0), np.float32(0), np.float32(0)]
            p[axis] = d
            cloud3.append(p)
            idx.append([lround(np.float32(p[0] + tx) / np.float32(0.05)),
                        lround(np.float32(p[1] + ty) / np.float32(0.05)),
                        lround(np.float32(p[2] + tz) / np.float32(0.05))])
            val.append(24575)
            d = np.float32(d + np.float32(0.5))
    cloud3 = np.array(cloud3, np.float32)
    ohi = oracle.HybridGrid(0.05, np.array(idx, np.int32), np.array(val, np.uint16))
    olo = oracle.HybridGrid(0.05, np.array(idx, np.int32), np.array(val, np.uint16))
    hist = np.zeros(10, np.float32)
    od = dict(branch_and_bound_depth=6, full_resolution_depth=6, min_rotational_score=0.1,
              min_low_resolution_score=0.15, linear_xy_search_window=0.8,
              linear_z_search_window=0.8, angular_search_window=0.3)
    om3 = oracle.FastCorrelativeScanMatcher3D(ohi, olo, hist, od)
    node = dict(gravity_alignment=(1.0, 0.0, 0.0, 0.0), high_resolution_point_cloud=cloud3,
                low_resolution_point_cloud=cloud3, rotational_scan_matcher_histogram=hist)
    ident = [0, 0, 0, 1, 0, 0, 0]
    want = om3.match(ident, ident, node, 0.1)
    got = res["fast3d"][0]
    assert want["found"]
    assert _f(got[0]) == want["score"]
    np.testing.assert_array_equal([float(v) for v in got[1:8]], want["pose"])
    assert _f(got[8]) == want["rotational_score"] and _f(got[9]) == want["low_resolution_score"]
    # ConstraintBuilder3D with the device refinement (constraint_builder_3d.cc:265-275) and
    # CeresScanMatcher3D::Match through the C++ class (doubles, 1e-7)
    refined = oracle.ceres3d_match([(cloud3, ohi), (cloud3, olo)], want["pose"][:3], want["pose"])
    got = res["cb3d_refined"][0]
    np.testing.assert_allclose([float(v) for v in got[0:7]], refined["pose"], rtol=0, atol=1e-7)
    start = [0.22, -0.13, 0.08, 1, 0, 0, 0]
    direct = oracle.ceres3d_match([(cloud3, ohi)], start[:3], start, occupied_space_weights=[5.0])
    got = res["ceres3d"][0]
    np.testing.assert_allclose([float(v) for v in got[0:7]], direct["pose"], rtol=0, atol=1e-7)
    assert float(got[7]) == pytest.approx(direct["initial_cost"], rel=1e-12)
    assert float(got[8]) == pytest.approx(direct["final_cost"], rel=1e-9)
    assert [int(got[9]), int(got[10])] == [direct["iterations"], direct["num_successful_steps"]]
    assert oracle.CERES_TERMINATION[int(got[11])] == direct["termination"]
    # RealTimeCorrelativeScanMatcher3D::Match
    want = oracle.rt3d_match(ohi, cloud3, [0.25, -0.1, 0.05, 1, 0, 0, 0], 0.1, 0.01, 0.1, 1.0)
    got = res["rt3d"][0]
    assert _f(got[0]) == want["score"]
    np.testing.assert_array_equal([float(v) for v in got[1:8]], want["pose"])


def test_adapter_headers_compile_and_link():
    """CPU: the adapter builds against the stand-in headers and links the C ABI; without
    a device the binary reports that and exits 0 (no CPU fallback to exercise)."""
    subprocess.check_call(["make", "-C", ADAPTER, "-s"])
    out = subprocess.run([os.path.join(ADAPTER, "adapter_selftest")], capture_output=True,
                         text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    # the builder's gating / WhenDone cycle and the sampler sequence run without a device
    assert "host-only checks passed" in out.stdout
