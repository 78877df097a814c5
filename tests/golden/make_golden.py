#!/usr/bin/env python
"""Generates tests/golden/golden_v1.npz — self-contained input/output vectors of the
hot path, produced by the CPU oracle (oracle/), which is pinned against the reference's
own known-answer tests (tests/test_oracle_golden_*.py).

The reference is C++ with dependencies that are absent here (Eigen, Ceres, glog, abseil,
protobuf), so it cannot be executed to dump vectors; its own tests pin no exact
FastCorrelativeScanMatcher score or tie-break.  These fixtures freeze
the oracle's answers so that (a) the oracle cannot drift unnoticed
(tests/test_golden_fixtures_cpu.py) and (b) the CUDA engine is checked against committed
vectors, not only against a live oracle run (tests/test_gpu_golden.py).

    python tests/golden/make_golden.py        # rewrites golden_v1.npz

Inputs are stored inside the file (grids, clouds, poses, options), so the fixtures do not
depend on the synthetic generators staying unchanged.
"""
import math
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pyoracle as oracle  # noqa: E402
from tests import worlds, worlds3d     # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")


def bits(x):
    return np.array([np.float32(x)]).view(np.uint32)[0]


def cases_2d():
    """(name, grid record, cloud, initial pose, full, lin, ang, depth, min_score)."""
    out = []
    for seed, full, depth, lin, ang, min_score in [(3, False, 5, 2.0, 0.5, 0.3),
                                                   (4, True, 6, 0.0, 0.0, 0.4),
                                                   (5, False, 3, 1.0, 0.3, 0.2),
                                                   (6, True, 7, 0.0, 0.0, 0.5)]:
        grid, occ, pose, scan = worlds.small_world(seed, size_cells=160 + 20 * (seed % 3))
        rng = np.random.RandomState(100 + seed)
        init = np.array(pose) + rng.uniform(-1, 1, 3) * [0.6, 0.6, 0.2]
        out.append(("fast2d_%d" % seed, grid, scan, init, full, lin, ang, depth, min_score))
    return out


def main():
    data = {}
    names = []
    for name, grid, scan, init, full, lin, ang, depth, min_score in cases_2d():
        og = oracle.Grid2D(grid.cells, grid.resolution, grid.max_x, grid.max_y)
        om = oracle.FastCorrelativeScanMatcher2D(og, lin, ang, depth)
        w = om.match_full_submap(scan, min_score) if full else om.match(init, scan, min_score)
        data[name + "/cells"] = np.ascontiguousarray(grid.cells, np.uint16)
        data[name + "/limits"] = np.array([grid.resolution, grid.max_x, grid.max_y], np.float64)
        data[name + "/cloud"] = np.ascontiguousarray(scan, np.float32)
        data[name + "/init"] = np.asarray(init, np.float64)
        data[name + "/opts"] = np.array([float(full), lin, ang, depth, min_score], np.float64)
        data[name + "/found"] = np.array([int(w["found"])], np.int32)
        data[name + "/score_bits"] = np.array([bits(w["score"])], np.uint32)
        data[name + "/pose"] = np.asarray(w["pose"], np.float64)
        data[name + "/best"] = np.array([w["best_scan_index"], w["best_x_offset"],
                                         w["best_y_offset"]], np.int32)
        data[name + "/lowest_resolution_candidates"] = np.array(
            [w["lowest_resolution_candidates"]], np.int64)
        data[name + "/level_crc32"] = np.array(
            [zlib.crc32(om.level(l).tobytes()) for l in range(depth)], np.uint32)
        names.append(name)

    # real-time matcher (config 1 shape, small)
    grid, occ, pose, scan = worlds.small_world(9, size_cells=200)
    og = oracle.Grid2D(grid.cells, grid.resolution, grid.max_x, grid.max_y)
    init = np.array(pose) + [0.04, -0.03, 0.02]
    w = oracle.rt2d_match(og, scan, init, 0.1, math.radians(7.0), 0.1, 0.1)
    data["rt2d/cells"] = np.ascontiguousarray(grid.cells, np.uint16)
    data["rt2d/limits"] = np.array([grid.resolution, grid.max_x, grid.max_y], np.float64)
    data["rt2d/cloud"] = np.ascontiguousarray(scan, np.float32)
    data["rt2d/init"] = np.asarray(init, np.float64)
    data["rt2d/opts"] = np.array([0.1, math.radians(7.0), 0.1, 0.1], np.float64)
    data["rt2d/score"] = np.array([w["score"]], np.float64)
    data["rt2d/pose"] = np.asarray(w["pose"], np.float64)

    # 3D: the reference test's axis cloud inserted at two of its mt19937(42) poses
    rng = np.random.RandomState(42)
    for k in range(2):
        expected = worlds3d.random_pose(rng)
        og3 = worlds3d.insert_cloud(oracle, 0.05, expected)
        om3 = oracle.FastCorrelativeScanMatcher3D(og3, og3, np.zeros(10, np.float32),
                                                  worlds3d.TEST_OPTIONS)
        ident = [0, 0, 0, 1, 0, 0, 0]
        w = om3.match(ident, ident, worlds3d.node_data(worlds3d.AXIS_CLOUD), 0.1)
        name = "fast3d_%d" % k
        data[name + "/indices"] = np.ascontiguousarray(og3.spec.indices, np.int32)
        data[name + "/values"] = np.ascontiguousarray(og3.spec.values, np.uint16)
        data[name + "/resolution"] = np.array([0.05], np.float64)
        data[name + "/grid_size"] = np.array([og3.grid_size()], np.int32)
        data[name + "/cloud"] = worlds3d.AXIS_CLOUD
        data[name + "/found"] = np.array([int(w["found"])], np.int32)
        data[name + "/score_bits"] = np.array([bits(w["score"])], np.uint32)
        data[name + "/pose"] = np.asarray(w["pose"], np.float64)
        data[name + "/best"] = np.array([w["best_scan_index"], w["best_x"], w["best_y"],
                                         w["best_z"]], np.int32)
        data[name + "/gate"] = np.array([bits(w["rotational_score"]),
                                         bits(w["low_resolution_score"])], np.uint32)
        names.append(name)
    data["names"] = np.array(names)
    np.savez_compressed(OUT, **data)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(data), "arrays")


if __name__ == "__main__":
    main()
