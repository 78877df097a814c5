"""Host-side mirror of cartographer::mapping::scan_matching for the hot path,
over the C ABI (include/csm_abi.h).  Same class / method names, argument meaning
and "no match" convention as the reference; out-parameters become return values.

Reference: cartographer/mapping/internal/2d/scan_matching/
  fast_correlative_scan_matcher_2d.h:112-136, real_time_correlative_scan_matcher_2d.h:53-85.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import CsmStats, JOB2D_DTYPE, RESULT2D_DTYPE, check, lib, ptr


@dataclass
class FastCorrelativeScanMatcherOptions2D:
    """proto/scan_matching/fast_correlative_scan_matcher_options_2d.proto"""
    linear_search_window: float
    angular_search_window: float
    branch_and_bound_depth: int


@dataclass
class RealTimeCorrelativeScanMatcherOptions:
    """proto/scan_matching/real_time_correlative_scan_matcher_options.proto"""
    linear_search_window: float
    angular_search_window: float
    translation_delta_cost_weight: float
    rotation_delta_cost_weight: float


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError("point cloud must be N x 3 float32")
    return a


class DeviceCloud:
    """sensor::PointCloud resident on the device (csm_cloud)."""

    def __init__(self, xyz, device=0):
        xyz = _f32(xyz)
        self.num_points = len(xyz)
        self._h = C.c_void_p()
        check(lib().csm_cloud_create(ptr(xyz, C.c_float), C.c_int32(len(xyz)),
                                     C.c_int32(device), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib().csm_cloud_destroy(self._h)
            self._h = None

    __del__ = close


class FastCorrelativeScanMatcher2D:
    """FastCorrelativeScanMatcher2D(grid, options): builds the precomputation grid
    stack on the device (fast_correlative_scan_matcher_2d.cc:188-194).  `grid` is any
    record with cells (uint16 [num_y, num_x]), resolution, max_x, max_y, min_cost,
    max_cost — Grid2D's limits() and correspondence_cost_cells()."""

    def __init__(self, grid, options, device=0):
        self.options = options
        self.device = device
        cells = np.ascontiguousarray(grid.cells, dtype=np.uint16)
        self._h = C.c_void_p()
        check(lib().csm_stack2d_create(
            ptr(cells, C.c_uint16), C.c_int32(cells.shape[1]), C.c_int32(cells.shape[0]),
            C.c_double(grid.resolution), C.c_double(grid.max_x), C.c_double(grid.max_y),
            C.c_float(grid.min_cost), C.c_float(grid.max_cost),
            C.c_int32(options.branch_and_bound_depth), C.c_int32(device), C.byref(self._h)))
        self.last_stats = None

    def close(self):
        if getattr(self, "_h", None):
            lib().csm_stack2d_destroy(self._h)
            self._h = None

    __del__ = close

    def _match(self, xyz, initial_pose, full, min_score):
        xyz = _f32(xyz)
        ip = np.ascontiguousarray(initial_pose, dtype=np.float64)
        found = C.c_int32(0)
        score = C.c_float(0.0)
        pose = np.zeros(3, np.float64)
        stats = CsmStats()
        check(lib().csm_match2d(self._h, ptr(xyz, C.c_float), C.c_int32(len(xyz)),
                                ptr(ip, C.c_double), C.c_int32(int(full)),
                                C.c_double(self.options.linear_search_window),
                                C.c_double(self.options.angular_search_window),
                                C.c_float(min_score), C.byref(found), C.byref(score),
                                ptr(pose, C.c_double), C.byref(stats)))
        self.last_stats = stats.as_dict()
        if not found.value:
            return False, None, None
        return True, np.float32(score.value), pose

    def Match(self, initial_pose_estimate, point_cloud, min_score):
        """-> (found, score, pose_estimate[x, y, yaw]); (False, None, None) if no
        score above min_score (strict), fast_correlative_scan_matcher_2d.cc:198-208."""
        return self._match(point_cloud, initial_pose_estimate, False, min_score)

    def MatchFullSubmap(self, point_cloud, min_score):
        """fast_correlative_scan_matcher_2d.cc:210-225."""
        return self._match(point_cloud, (0.0, 0.0, 0.0), True, min_score)

    # -- "visible for testing" hooks ------------------------------------------------
    def precomputation_grid(self, level):
        nx, ny = C.c_int32(), C.c_int32()
        check(lib().csm_stack2d_read_level(self._h, C.c_int32(level), None, C.byref(nx),
                                           C.byref(ny)))
        out = np.empty((ny.value, nx.value), np.uint8)
        check(lib().csm_stack2d_read_level(self._h, C.c_int32(level), ptr(out, C.c_uint8),
                                           C.byref(nx), C.byref(ny)))
        return out

    def score_candidates(self, level, discrete_scans, candidates):
        ds = np.ascontiguousarray(discrete_scans, dtype=np.int32)
        cand = np.ascontiguousarray(candidates, dtype=np.int32).reshape(-1, 3)
        S, n, _ = ds.shape
        scores = np.empty(len(cand), np.float32)
        sums = np.empty(len(cand), np.int32)
        check(lib().csm_score_candidates2d(self._h, C.c_int32(level), ptr(ds, C.c_int32),
                                           C.c_int32(S), C.c_int32(n), ptr(cand, C.c_int32),
                                           C.c_int32(len(cand)), ptr(scores, C.c_float),
                                           ptr(sums, C.c_int32)))
        return scores, sums

    def discretize(self, point_cloud, initial_pose_estimate=(0.0, 0.0, 0.0), full_submap=False):
        xyz = _f32(point_cloud)
        ip = np.ascontiguousarray(initial_pose_estimate, dtype=np.float64)
        S = C.c_int32(0)
        args = (self._h, ptr(xyz, C.c_float), C.c_int32(len(xyz)), ptr(ip, C.c_double),
                C.c_int32(int(full_submap)), C.c_double(self.options.linear_search_window),
                C.c_double(self.options.angular_search_window), C.byref(S))
        check(lib().csm_discretize2d(*args, None, None))
        ds = np.empty((S.value, len(xyz), 2), np.int32)
        bounds = np.empty((S.value, 4), np.int32)
        check(lib().csm_discretize2d(*args, ptr(ds, C.c_int32), ptr(bounds, C.c_int32)))
        return ds, bounds


def match_batch(matchers, clouds, jobs, linear_search_window, angular_search_window):
    """Batched FastCorrelativeScanMatcher2D searches (csm_match2d_batch).

    matchers: list of FastCorrelativeScanMatcher2D; clouds: list of DeviceCloud;
    jobs: numpy array of JOB2D_DTYPE.  Returns (results[RESULT2D_DTYPE], stats dict).
    """
    jobs = np.ascontiguousarray(jobs, dtype=JOB2D_DTYPE)
    hs = (C.c_void_p * len(matchers))(*[m._h for m in matchers])
    cs = (C.c_void_p * len(clouds))(*[c._h for c in clouds])
    results = np.zeros(len(jobs), RESULT2D_DTYPE)
    stats = CsmStats()
    check(lib().csm_match2d_batch(hs, C.c_int32(len(matchers)), cs, C.c_int32(len(clouds)),
                                  jobs.ctypes.data_as(C.c_void_p), C.c_int32(len(jobs)),
                                  C.c_double(linear_search_window),
                                  C.c_double(angular_search_window),
                                  results.ctypes.data_as(C.c_void_p), C.byref(stats)))
    return results, stats.as_dict()


class RealTimeCorrelativeScanMatcher2D:
    """real_time_correlative_scan_matcher_2d.h:53-85."""

    def __init__(self, options, device=0):
        self.options = options
        self.device = device
        self.last_stats = None

    def Match(self, initial_pose_estimate, point_cloud, grid):
        """-> (score, pose_estimate), real_time_correlative_scan_matcher_2d.cc:117-149."""
        xyz = _f32(point_cloud)
        cells = np.ascontiguousarray(grid.cells, dtype=np.uint16)
        ip = np.ascontiguousarray(initial_pose_estimate, dtype=np.float64)
        pose = np.zeros(3, np.float64)
        score = C.c_double(0.0)
        stats = CsmStats()
        o = self.options
        check(lib().csm_rt_match2d(
            ptr(cells, C.c_uint16), C.c_int32(cells.shape[1]), C.c_int32(cells.shape[0]),
            C.c_double(grid.resolution), C.c_double(grid.max_x), C.c_double(grid.max_y),
            ptr(xyz, C.c_float), C.c_int32(len(xyz)), ptr(ip, C.c_double),
            C.c_double(o.linear_search_window), C.c_double(o.angular_search_window),
            C.c_double(o.translation_delta_cost_weight),
            C.c_double(o.rotation_delta_cost_weight), C.c_int32(self.device), C.byref(score),
            ptr(pose, C.c_double), C.byref(stats)))
        self.last_stats = stats.as_dict()
        return score.value, pose


def kernel_launch_count():
    return int(lib().csm_kernel_launch_count())


def device_count():
    n = C.c_int32(0)
    check(lib().csm_device_count(C.byref(n)))
    return n.value


__all__ = ["FastCorrelativeScanMatcherOptions2D", "RealTimeCorrelativeScanMatcherOptions",
           "FastCorrelativeScanMatcher2D", "RealTimeCorrelativeScanMatcher2D", "DeviceCloud",
           "match_batch", "kernel_launch_count", "device_count", "JOB2D_DTYPE",
           "RESULT2D_DTYPE", "_lib"]
