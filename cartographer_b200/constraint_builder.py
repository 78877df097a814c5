"""Host-side mirror of cartographer::mapping::constraints::ConstraintBuilder2D
(cartographer/mapping/internal/constraints/constraint_builder_2d.{h,cc}) for the
fast-correlative-scan-matching part of the queue.

Same public surface and gating as the reference —
  MaybeAddConstraint / MaybeAddGlobalConstraint / NotifyEndOfNode / WhenDone /
  GetNumFinishedNodes / DeleteScanMatcher
— but instead of one thread-pool task per (submap, node) pair the queued pairs are
drained as ONE batched device call (csm_match2d_batch) when WhenDone is invoked.
With torch.distributed initialised the queue is sharded submap-major across ranks
(rank r owns the submaps whose first-use order index % world == r; stacks are
only built on their owner) and the fixed-size constraint records are exchanged
with a single all_gather; every rank then reports the same Result in queue order.

By default constraints are the fast matcher's: `pose` is ComputeSubmapPose(submap)^-1 *
pose_estimate (constraint_builder_2d.cc:251-258), which is what the bit-exact parity
tests pin.  With ConstraintBuilderOptions.ceres_scan_matcher_options set, every found
match is first refined as constraint_builder_2d.cc:245-249 does — all of a drain's
matches in one csm_ceres_match2d_batch launch (scan_matching.CeresScanMatcher2D; Ceres
itself is not linked).
"""
import math
from dataclasses import dataclass, field

import numpy as np

RECORD_DTYPE = np.dtype([("submap_trajectory", "<i4"), ("submap_index", "<i4"),
                         ("node_trajectory", "<i4"), ("node_index", "<i4"),
                         ("found", "<u4"), ("score", "<f4"), ("pose", "<f8", (3,)),
                         ("pad", "<u4", (4,))], align=True)
assert RECORD_DTYPE.itemsize == 64


@dataclass
class ConstraintBuilderOptions:
    """constraints/proto/constraint_builder_options.proto (the fields on the path);
    defaults from configuration_files/pose_graph.lua:18-36."""
    sampling_ratio: float = 0.3
    max_constraint_distance: float = 15.0
    min_score: float = 0.55
    global_localization_min_score: float = 0.6
    loop_closure_translation_weight: float = 1.1e4
    loop_closure_rotation_weight: float = 1e5
    linear_search_window: float = 7.0
    angular_search_window: float = math.radians(30.0)
    branch_and_bound_depth: int = 7
    # scan_matching.CeresScanMatcherOptions2D, or None to leave matches unrefined
    ceres_scan_matcher_options: object = None


@dataclass
class Submap2D:
    """What the builder needs of a Submap2D: its grid and local pose (x, y, yaw)."""
    grid: object
    local_pose: tuple = (0.0, 0.0, 0.0)


@dataclass
class Constraint:
    """PoseGraphInterface::Constraint (mapping/pose_graph_interface.h:36-53)."""
    submap_id: tuple
    node_id: tuple
    zbar_ij: tuple            # (x, y, yaw) of submap i <- node j
    translation_weight: float
    rotation_weight: float
    tag: str = "INTER_SUBMAP"
    score: float = 0.0


class FixedRatioSampler:
    """common/fixed_ratio_sampler.cc:23-39."""

    def __init__(self, ratio):
        if not 0.0 <= ratio <= 1.0:
            raise ValueError("ratio must be in [0, 1]")
        self.ratio = ratio
        self.num_pulses = 0
        self.num_samples = 0

    def Pulse(self):
        self.num_pulses += 1
        if self.num_samples / self.num_pulses < self.ratio:
            self.num_samples += 1
            return True
        return False


def _compose(a, b):
    """transform::Rigid2d operator* (transform/rigid_transform.h:93-99)."""
    c, s = math.cos(a[2]), math.sin(a[2])
    return (c * b[0] - s * b[1] + a[0], s * b[0] + c * b[1] + a[1], a[2] + b[2])


def _inverse(a):
    """Rigid2::inverse (transform/rigid_transform.h:76-80)."""
    c, s = math.cos(-a[2]), math.sin(-a[2])
    return (-(c * a[0] - s * a[1]), -(s * a[0] + c * a[1]), -a[2])


@dataclass
class _Job:
    submap_id: tuple
    node_id: tuple
    cloud_key: int
    initial_pose: tuple
    full: bool
    min_score: float


class CudaExecutor:
    """Drains a list of jobs through csm_match2d_batch on one device.  Keeps one
    device stack per submap (DispatchScanMatcherConstruction,
    constraint_builder_2d.cc:165-186) and one device cloud per node scan."""

    def __init__(self, options, device=0):
        from . import scan_matching as sm
        self.sm = sm
        self.options = options
        self.device = device
        self.matchers = {}
        self.grids = {}     # submap id -> RealTimeGrid2D (only with the Ceres refinement on)
        self.clouds = {}
        self.stats = {"candidates_scored": 0, "device_ms": 0.0, "searched": 0, "refined": 0,
                      "refine_device_ms": 0.0}

    def ensure_matcher(self, submap_id, grid):
        if submap_id not in self.matchers:
            o = self.options
            self.matchers[submap_id] = self.sm.FastCorrelativeScanMatcher2D(
                grid, self.sm.FastCorrelativeScanMatcherOptions2D(
                    o.linear_search_window, o.angular_search_window, o.branch_and_bound_depth),
                device=self.device)

    def delete_matcher(self, submap_id):
        m = self.matchers.pop(submap_id, None)
        if m is not None:
            m.close()
        g = self.grids.pop(submap_id, None)
        if g is not None:
            g.close()

    def refine(self, jobs, submaps, clouds, out):
        """constraint_builder_2d.cc:245-249 for every found match of the drain, one launch."""
        sm = self.sm
        idx = [i for i, r in enumerate(out) if r[0]]
        if not idx:
            return out
        for i in idx:
            sid = jobs[i].submap_id
            if sid not in self.grids:
                self.grids[sid] = sm.RealTimeGrid2D(submaps[sid].grid, device=self.device)
        matcher = sm.CeresScanMatcher2D(self.options.ceres_scan_matcher_options)
        poses, _ = matcher.MatchBatch([out[i][2][:2] for i in idx], [out[i][2] for i in idx],
                                      [clouds[jobs[i].cloud_key] for i in idx],
                                      [self.grids[jobs[i].submap_id] for i in idx])
        self.stats["refined"] += len(idx)
        self.stats["refine_device_ms"] += matcher.last_stats["device_ms"]
        out = list(out)
        for i, p in zip(idx, poses):
            out[i] = (True, out[i][1], (float(p[0]), float(p[1]), float(p[2])))
        return out

    def run(self, jobs, submaps, clouds):
        """-> list of (found, score, pose_estimate) in job order."""
        if not jobs:
            return []
        sm = self.sm
        sub_ids = sorted({j.submap_id for j in jobs})
        for sid in sub_ids:
            self.ensure_matcher(sid, submaps[sid].grid)
        cloud_keys = sorted({j.cloud_key for j in jobs})
        for k in cloud_keys:
            if k not in self.clouds:
                self.clouds[k] = sm.DeviceCloud(clouds[k], device=self.device)
        sidx = {s: i for i, s in enumerate(sub_ids)}
        cidx = {k: i for i, k in enumerate(cloud_keys)}
        arr = np.zeros(len(jobs), sm.JOB2D_DTYPE)
        for i, j in enumerate(jobs):
            arr[i]["stack_index"] = sidx[j.submap_id]
            arr[i]["cloud_index"] = cidx[j.cloud_key]
            arr[i]["full_submap"] = int(j.full)
            arr[i]["initial_pose"] = j.initial_pose
            arr[i]["min_score"] = j.min_score
        res, st = sm.match_batch([self.matchers[s] for s in sub_ids],
                                 [self.clouds[k] for k in cloud_keys], arr,
                                 self.options.linear_search_window,
                                 self.options.angular_search_window)
        self.stats["candidates_scored"] += st["candidates_scored"]
        self.stats["device_ms"] += st["device_ms"]
        self.stats["searched"] += len(jobs)
        out = [(bool(r["found"]), float(r["score"]), tuple(r["pose_estimate"])) for r in res]
        if self.options.ceres_scan_matcher_options is not None:
            out = self.refine(jobs, submaps, clouds, out)
        return out

    def release_clouds(self):
        for c in self.clouds.values():
            c.close()
        self.clouds = {}


class ConstraintBuilder2D:
    """constraint_builder_2d.h:60-107.  `executor.run(jobs, submaps, clouds)` does the
    matching (CudaExecutor in production); `process_group` (torch.distributed) turns
    on submap-major sharding + the single all_gather of constraint records."""

    def __init__(self, options, executor=None, process_group=None, device=0, context=None):
        """`context` (scan_matching.MultiGpuContext): the library's own NCCL communicator;
        when given, the single all-gather runs inside libcsm_b200.so (csm_ctx_allgather ->
        ncclAllGather) and no torch.distributed group is needed."""
        self.options = options
        self.executor = executor if executor is not None else CudaExecutor(options, device)
        self.pg = process_group
        self.ctx = context
        self.device = device
        self._jobs = []
        self._submaps = {}
        self._clouds = {}
        self._submap_order = {}
        self._samplers = {}
        self._when_done = None
        self.num_started_nodes = 0
        self.num_finished_nodes = 0
        self.last_records = None

    # -- queue --------------------------------------------------------------------
    def _register(self, submap_id, submap, constant_data):
        self._submaps[submap_id] = submap
        self._submap_order.setdefault(submap_id, len(self._submap_order))
        key = id(constant_data)
        self._clouds[key] = constant_data
        return key

    def MaybeAddConstraint(self, submap_id, submap, node_id, constant_data,
                           initial_relative_pose):
        """constraint_builder_2d.cc:77-112.  constant_data is the node's
        filtered_gravity_aligned_point_cloud (N x 3 float32)."""
        if math.hypot(initial_relative_pose[0], initial_relative_pose[1]) > \
                self.options.max_constraint_distance:
            return
        sampler = self._samplers.setdefault(submap_id,
                                            FixedRatioSampler(self.options.sampling_ratio))
        if not sampler.Pulse():
            return
        key = self._register(submap_id, submap, constant_data)
        initial_pose = _compose(tuple(submap.local_pose), tuple(initial_relative_pose))  # :196-197
        self._jobs.append(_Job(submap_id, node_id, key, initial_pose, False,
                               self.options.min_score))

    def MaybeAddGlobalConstraint(self, submap_id, submap, node_id, constant_data):
        """constraint_builder_2d.cc:114-137."""
        key = self._register(submap_id, submap, constant_data)
        self._jobs.append(_Job(submap_id, node_id, key, (0.0, 0.0, 0.0), True,
                               self.options.global_localization_min_score))

    def NotifyEndOfNode(self):
        """constraint_builder_2d.cc:139-151 (the node finishes when WhenDone drains)."""
        self.num_started_nodes += 1

    def GetNumFinishedNodes(self):
        return self.num_finished_nodes

    def DeleteScanMatcher(self, submap_id):
        """constraint_builder_2d.cc:307-316."""
        if hasattr(self.executor, "delete_matcher"):
            self.executor.delete_matcher(submap_id)
        self._samplers.pop(submap_id, None)

    # -- drain ---------------------------------------------------------------------
    def _world(self):
        if self.ctx is not None:
            return self.ctx.rank, self.ctx.world_size
        if self.pg is None:
            return 0, 1
        import torch.distributed as dist
        return dist.get_rank(self.pg), dist.get_world_size(self.pg)

    def WhenDone(self, callback):
        """constraint_builder_2d.cc:153-163 + RunWhenDoneCallback :279-300: runs the
        queue, then calls callback(Result) with the found constraints in queue order."""
        if self._when_done is not None:
            raise RuntimeError("WhenDone called twice")  # CHECK(when_done_ == nullptr)
        self._when_done = callback
        rank, world = self._world()
        jobs = self._jobs
        owner = [self._submap_order[j.submap_id] % world for j in jobs]
        mine = [i for i, o in enumerate(owner) if o == rank]
        out = self.executor.run([jobs[i] for i in mine], self._submaps, self._clouds)
        local = np.zeros(len(mine), RECORD_DTYPE)
        for rec, i, (found, score, pose) in zip(local, mine, out):
            j = jobs[i]
            rec["submap_trajectory"], rec["submap_index"] = j.submap_id
            rec["node_trajectory"], rec["node_index"] = j.node_id
            rec["found"] = int(found)
            if found:
                rec["score"] = score
                # constraint_transform = ComputeSubmapPose(submap)^-1 * pose_estimate (:251-252)
                rec["pose"] = _compose(_inverse(tuple(self._submaps[j.submap_id].local_pose)), pose)
        records = self._allgather(local, mine, len(jobs), rank, world)
        self.last_records = records
        result = []
        for rec in records:
            if rec["found"]:
                result.append(Constraint(
                    (int(rec["submap_trajectory"]), int(rec["submap_index"])),
                    (int(rec["node_trajectory"]), int(rec["node_index"])),
                    tuple(float(v) for v in rec["pose"]),
                    self.options.loop_closure_translation_weight,
                    self.options.loop_closure_rotation_weight, "INTER_SUBMAP",
                    float(rec["score"])))
        self._jobs = []
        self._clouds = {}
        if hasattr(self.executor, "release_clouds"):
            self.executor.release_clouds()
        self.num_finished_nodes = self.num_started_nodes
        cb, self._when_done = self._when_done, None
        cb(result)
        return result

    def _allgather(self, local, mine, total, rank, world):
        """The path's single collective: fixed-size records, padded to the largest
        shard, gathered to every rank and re-ordered into queue order."""
        if world == 1:
            return local
        counts = [sum(1 for j in self._jobs if self._submap_order[j.submap_id] % world == r)
                  for r in range(world)]
        cap = max(1, max(counts))
        if self.ctx is not None:   # ncclAllGather inside the library
            send = np.zeros(cap, RECORD_DTYPE)
            send[:len(local)] = local
            parts = self.ctx.allgather(send.tobytes())
            allrec = np.stack([np.frombuffer(p, dtype=RECORD_DTYPE) for p in parts])
            out = np.zeros(total, RECORD_DTYPE)
            cursor = [0] * world
            for i, j in enumerate(self._jobs):
                r = self._submap_order[j.submap_id] % world
                out[i] = allrec[r, cursor[r]]
                cursor[r] += 1
            return out
        import torch
        import torch.distributed as dist
        backend = dist.get_backend(self.pg)
        dev = torch.device("cuda", self.device) if backend == "nccl" else torch.device("cpu")
        send = np.zeros(cap, RECORD_DTYPE)
        send[:len(local)] = local
        t_send = torch.from_numpy(send.view(np.uint8).reshape(-1).copy()).to(dev)
        t_recv = torch.empty(world * cap * RECORD_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(t_recv, t_send, group=self.pg)
        allrec = np.frombuffer(t_recv.cpu().numpy().tobytes(), dtype=RECORD_DTYPE).reshape(
            world, cap)
        out = np.zeros(total, RECORD_DTYPE)
        cursor = [0] * world
        for i, j in enumerate(self._jobs):
            r = self._submap_order[j.submap_id] % world
            out[i] = allrec[r, cursor[r]]
            cursor[r] += 1
        return out


# ===========================================================================
# ConstraintBuilder3D (mapping/internal/constraints/constraint_builder_3d.{h,cc})
# ===========================================================================
RECORD3D_DTYPE = np.dtype([("submap_trajectory", "<i4"), ("submap_index", "<i4"),
                           ("node_trajectory", "<i4"), ("node_index", "<i4"),
                           ("found", "<u4"), ("score", "<f4"), ("pose", "<f8", (7,)),
                           ("rotational_score", "<f4"), ("low_resolution_score", "<f4"),
                           ("pad", "<u4", (2,))], align=True)
assert RECORD3D_DTYPE.itemsize == 96


@dataclass
class ConstraintBuilderOptions3D:
    """constraint_builder_options.proto + fast_correlative_scan_matcher_options_3d;
    defaults from configuration_files/pose_graph.lua:18-48."""
    sampling_ratio: float = 0.3
    max_constraint_distance: float = 15.0
    min_score: float = 0.55
    global_localization_min_score: float = 0.6
    loop_closure_translation_weight: float = 1.1e4
    loop_closure_rotation_weight: float = 1e5
    branch_and_bound_depth: int = 8
    full_resolution_depth: int = 3
    min_rotational_score: float = 0.77
    min_low_resolution_score: float = 0.55
    linear_xy_search_window: float = 5.0
    linear_z_search_window: float = 1.0
    angular_search_window: float = math.radians(15.0)
    # scan_matching.CeresScanMatcherOptions3D, or None to leave matches unrefined
    ceres_scan_matcher_options_3d: object = None


@dataclass
class Submap3D:
    """What the builder needs of a Submap3D (constraint_builder_3d.cc:172-198):
    high / low resolution hybrid grids and the rotational scan matcher histogram."""
    high_resolution_hybrid_grid: object
    low_resolution_hybrid_grid: object
    rotational_scan_matcher_histogram: object
    grid_size_in_voxels: int = 0


@dataclass
class _Job3:
    submap_id: tuple
    node_id: tuple
    node_key: int
    node_pose: tuple
    submap_pose: tuple
    full: bool
    min_score: float


class CudaExecutor3D:
    """One csm_matcher3d per submap (DispatchScanMatcherConstruction,
    constraint_builder_3d.cc:172-198); jobs run through csm_match3d_batch."""

    def __init__(self, options, device=0, threads=8):
        from . import scan_matching as sm
        self.sm = sm
        self.options = options
        self.device = device
        self.threads = threads
        self.matchers = {}
        self.grids = {}    # submap id -> (high, low) DeviceHybridGrid, for the refinement
        self.stats = {"candidates_scored": 0, "device_ms": 0.0, "searched": 0, "refined": 0,
                      "refine_device_ms": 0.0}

    def delete_matcher(self, submap_id):
        m = self.matchers.pop(submap_id, None)
        if m is not None:
            m.close()
        for g in self.grids.pop(submap_id, ()):
            g.close()

    def refine(self, jobs, submaps, nodes, out):
        """constraint_builder_3d.cc:265-275 for every found match of the drain, one launch."""
        sm = self.sm
        idx = [i for i, r in enumerate(out) if r is not None]
        if not idx:
            return out
        pairs = []
        for i in idx:
            sid = jobs[i].submap_id
            if sid not in self.grids:
                sub = submaps[sid]
                self.grids[sid] = (sm.DeviceHybridGrid(sub.high_resolution_hybrid_grid, self.device),
                                   sm.DeviceHybridGrid(sub.low_resolution_hybrid_grid, self.device))
            node = nodes[jobs[i].node_key]
            hi, lo = self.grids[sid]
            pairs.append([(node.high_resolution_point_cloud, hi),
                          (node.low_resolution_point_cloud, lo)])
        matcher = sm.CeresScanMatcher3D(self.options.ceres_scan_matcher_options_3d)
        poses, _ = matcher.MatchBatch([out[i]["pose_estimate"][:3] for i in idx],
                                      [out[i]["pose_estimate"] for i in idx], pairs)
        self.stats["refined"] += len(idx)
        self.stats["refine_device_ms"] += matcher.last_stats["device_ms"]
        out = list(out)
        for i, p in zip(idx, poses):
            out[i] = dict(out[i], pose_estimate=np.array(p, np.float64))
        return out

    def run(self, jobs, submaps, nodes):
        """Matchers are built first (one per submap, like the reference's creation
        tasks); the whole queue then goes to the library in one call."""
        sm, o = self.sm, self.options
        for j in jobs:
            if j.submap_id not in self.matchers:
                sub = submaps[j.submap_id]
                self.matchers[j.submap_id] = sm.FastCorrelativeScanMatcher3D(
                    sub.high_resolution_hybrid_grid, sub.low_resolution_hybrid_grid,
                    sub.rotational_scan_matcher_histogram,
                    sm.FastCorrelativeScanMatcherOptions3D(
                        o.branch_and_bound_depth, o.full_resolution_depth,
                        o.min_rotational_score, o.min_low_resolution_score,
                        o.linear_xy_search_window, o.linear_z_search_window,
                        o.angular_search_window),
                    device=self.device, grid_size_in_voxels=sub.grid_size_in_voxels)

        # one csm_match3d_batch call: the library keeps `threads` matches in flight on
        # separate CUDA streams (the reference's pool threads call Match concurrently
        # too, constraint_builder_3d.cc:104-113)
        sub_ids = sorted({j.submap_id for j in jobs})
        node_keys = sorted({j.node_key for j in jobs})
        sub_index = {k: i for i, k in enumerate(sub_ids)}
        node_index = {k: i for i, k in enumerate(node_keys)}
        out, st = sm.match_batch3d(
            [self.matchers[k] for k in sub_ids], [nodes[k] for k in node_keys],
            [(sub_index[j.submap_id], node_index[j.node_key], j.full, j.node_pose,
              j.submap_pose, j.min_score) for j in jobs],
            max_concurrency=self.threads)
        self.stats["candidates_scored"] += st["candidates_scored"]
        self.stats["device_ms"] += st["device_ms"]
        self.stats["searched"] += len(jobs)
        if self.options.ceres_scan_matcher_options_3d is not None:
            out = self.refine(jobs, submaps, nodes, out)
        return out


class ConstraintBuilder3D:
    """constraint_builder_3d.h:59-114 over the 3D engine; same sharding / single
    all_gather scheme as ConstraintBuilder2D.  Constraint poses are the fast
    matcher's pose_estimate (submap <- node); with
    ConstraintBuilderOptions3D.ceres_scan_matcher_options_3d set they are first refined as
    constraint_builder_3d.cc:265-275 does (one csm_ceres_match3d_batch launch per drain)."""

    def __init__(self, options, executor=None, process_group=None, device=0):
        self.options = options
        self.executor = executor if executor is not None else CudaExecutor3D(options, device)
        self.pg = process_group
        self.device = device
        self._jobs, self._submaps, self._nodes = [], {}, {}
        self._submap_order, self._samplers = {}, {}
        self._when_done = None
        self.num_started_nodes = 0
        self.num_finished_nodes = 0
        self.last_records = None

    def _register(self, submap_id, submap, constant_data):
        self._submaps[submap_id] = submap
        self._submap_order.setdefault(submap_id, len(self._submap_order))
        key = id(constant_data)
        self._nodes[key] = constant_data
        return key

    def MaybeAddConstraint(self, submap_id, submap, node_id, constant_data, global_node_pose,
                           global_submap_pose):
        """constraint_builder_3d.cc:79-114; poses are [tx,ty,tz,qw,qx,qy,qz]."""
        d = np.asarray(global_node_pose[:3], float) - np.asarray(global_submap_pose[:3], float)
        if float(np.sqrt((d * d).sum())) > self.options.max_constraint_distance:
            return
        sampler = self._samplers.setdefault(submap_id,
                                            FixedRatioSampler(self.options.sampling_ratio))
        if not sampler.Pulse():
            return
        key = self._register(submap_id, submap, constant_data)
        self._jobs.append(_Job3(submap_id, node_id, key, tuple(global_node_pose),
                                tuple(global_submap_pose), False, self.options.min_score))

    def MaybeAddGlobalConstraint(self, submap_id, submap, node_id, constant_data,
                                 global_node_rotation, global_submap_rotation):
        """constraint_builder_3d.cc:116-142; rotations are [qw,qx,qy,qz]."""
        key = self._register(submap_id, submap, constant_data)
        self._jobs.append(_Job3(submap_id, node_id, key, (0, 0, 0) + tuple(global_node_rotation),
                                (0, 0, 0) + tuple(global_submap_rotation), True,
                                self.options.global_localization_min_score))

    def NotifyEndOfNode(self):
        self.num_started_nodes += 1

    def GetNumFinishedNodes(self):
        return self.num_finished_nodes

    def DeleteScanMatcher(self, submap_id):
        if hasattr(self.executor, "delete_matcher"):
            self.executor.delete_matcher(submap_id)
        self._samplers.pop(submap_id, None)

    def WhenDone(self, callback):
        if self._when_done is not None:
            raise RuntimeError("WhenDone called twice")
        self._when_done = callback
        rank, world = (0, 1)
        if self.pg is not None:
            import torch.distributed as dist
            rank, world = dist.get_rank(self.pg), dist.get_world_size(self.pg)
        jobs = self._jobs
        mine = [i for i, j in enumerate(jobs) if self._submap_order[j.submap_id] % world == rank]
        out = self.executor.run([jobs[i] for i in mine], self._submaps, self._nodes)
        local = np.zeros(len(mine), RECORD3D_DTYPE)
        for rec, i, r in zip(local, mine, out):
            j = jobs[i]
            rec["submap_trajectory"], rec["submap_index"] = j.submap_id
            rec["node_trajectory"], rec["node_index"] = j.node_id
            if r is not None:
                rec["found"] = 1
                rec["score"] = r["score"]
                rec["pose"] = r["pose_estimate"]
                rec["rotational_score"] = r["rotational_score"]
                rec["low_resolution_score"] = r["low_resolution_score"]
        records = _allgather_records(self, local, len(jobs), rank, world, RECORD3D_DTYPE)
        self.last_records = records
        result = [Constraint((int(r["submap_trajectory"]), int(r["submap_index"])),
                             (int(r["node_trajectory"]), int(r["node_index"])),
                             tuple(float(v) for v in r["pose"]),
                             self.options.loop_closure_translation_weight,
                             self.options.loop_closure_rotation_weight, "INTER_SUBMAP",
                             float(r["score"])) for r in records if r["found"]]
        self._jobs, self._nodes = [], {}
        self.num_finished_nodes = self.num_started_nodes
        cb, self._when_done = self._when_done, None
        cb(result)
        return result


def _allgather_records(builder, local, total, rank, world, dtype):
    """Shared single-collective exchange (see ConstraintBuilder2D._allgather)."""
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    owner = [builder._submap_order[j.submap_id] % world for j in builder._jobs]
    counts = [owner.count(r) for r in range(world)]
    cap = max(1, max(counts))
    backend = dist.get_backend(builder.pg)
    dev = torch.device("cuda", builder.device) if backend == "nccl" else torch.device("cpu")
    send = np.zeros(cap, dtype)
    send[:len(local)] = local
    t_send = torch.from_numpy(send.view(np.uint8).reshape(-1).copy()).to(dev)
    t_recv = torch.empty(world * cap * dtype.itemsize, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(t_recv, t_send, group=builder.pg)
    allrec = np.frombuffer(t_recv.cpu().numpy().tobytes(), dtype=dtype).reshape(world, cap)
    out = np.zeros(total, dtype)
    cursor = [0] * world
    for i, r in enumerate(owner):
        out[i] = allrec[r, cursor[r]]
        cursor[r] += 1
    return out
