"""ConstraintBuilder3D host logic on CPU: the reference's call cycle
(constraints/constraint_builder_3d_test.cc:61-) and the N>1 path — submap-major sharding
+ the single all_gather of 96-byte records — with world_size 2 over gloo.  The executor is
the oracle here (test infrastructure); on the GPU the same class drives csm_match3d_batch
(tests/test_gpu_parity_3d.py::test_constraint_builder_3d_on_device)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from cartographer_b200 import constraint_builder as cb
from tests import worlds3d


class OracleExecutor3D:
    def __init__(self, options):
        from oracle import pyoracle
        pyoracle.build()
        self.o = pyoracle
        self.options = options
        self.matchers = {}
        self.touched = set()

    def run(self, jobs, submaps, nodes):
        o, out = self.options, []
        od = dict(branch_and_bound_depth=o.branch_and_bound_depth,
                  full_resolution_depth=o.full_resolution_depth,
                  min_rotational_score=o.min_rotational_score,
                  min_low_resolution_score=o.min_low_resolution_score,
                  linear_xy_search_window=o.linear_xy_search_window,
                  linear_z_search_window=o.linear_z_search_window,
                  angular_search_window=o.angular_search_window)
        for j in jobs:
            self.touched.add(j.submap_id)
            if j.submap_id not in self.matchers:
                sub = submaps[j.submap_id]
                hi = self.o.HybridGrid(sub.high_resolution_hybrid_grid.resolution,
                                       sub.high_resolution_hybrid_grid.indices,
                                       sub.high_resolution_hybrid_grid.values)
                lo = self.o.HybridGrid(sub.low_resolution_hybrid_grid.resolution,
                                       sub.low_resolution_hybrid_grid.indices,
                                       sub.low_resolution_hybrid_grid.values)
                self.matchers[j.submap_id] = (self.o.FastCorrelativeScanMatcher3D(
                    hi, lo, sub.rotational_scan_matcher_histogram, od), hi, lo)
            m = self.matchers[j.submap_id][0]
            node = nodes[j.node_key]
            r = (m.match_full_submap(j.node_pose[3:], j.submap_pose[3:], node, j.min_score)
                 if j.full else m.match(j.node_pose, j.submap_pose, node, j.min_score))
            out.append(None if not r["found"] else dict(
                score=r["score"], pose_estimate=r["pose"], rotational_score=r["rotational_score"],
                low_resolution_score=r["low_resolution_score"]))
        return out

    def delete_matcher(self, submap_id):
        self.matchers.pop(submap_id, None)


def _queue():
    opts = cb.ConstraintBuilderOptions3D(sampling_ratio=1.0, max_constraint_distance=50.0,
                                         min_score=0.1, global_localization_min_score=0.1,
                                         **worlds3d.TEST_OPTIONS)
    rng = np.random.RandomState(42)
    submaps = {}
    for s in range(3):
        pose = worlds3d.random_pose(rng)
        spec = worlds3d.range_insert_3d(0.05, pose[:3], worlds3d.transform_points(
            pose, worlds3d.AXIS_CLOUD))
        submaps[(0, s)] = cb.Submap3D(spec, spec, np.zeros(10, np.float32))
    node = worlds3d.node_data(worlds3d.AXIS_CLOUD)
    return opts, submaps, node


def _fill(b, submaps, node):
    ident = [0, 0, 0, 1, 0, 0, 0]
    for n in range(2):
        for sid, sub in submaps.items():
            b.MaybeAddConstraint(sid, sub, (0, n), node, ident, ident)
        b.MaybeAddGlobalConstraint((0, n), submaps[(0, n)], (0, n), node, [1, 0, 0, 0],
                                   [1, 0, 0, 0])
        b.NotifyEndOfNode()


def test_3d_builder_cycle_on_the_oracle():
    opts, submaps, node = _queue()
    b = cb.ConstraintBuilder3D(opts, executor=OracleExecutor3D(opts))
    _fill(b, submaps, node)
    called = []
    res = b.WhenDone(called.append)
    assert len(called) == 1 and called[0] is res
    assert b.GetNumFinishedNodes() == 2 and len(res) >= 6
    assert all(c.tag == "INTER_SUBMAP" and len(c.zbar_ij) == 7 for c in res)
    # far pairs are gated (constraint_builder_3d.cc:85-88)
    b.MaybeAddConstraint((0, 0), submaps[(0, 0)], (0, 9), node, [100, 0, 0, 1, 0, 0, 0],
                         [0, 0, 0, 1, 0, 0, 0])
    assert len(b._jobs) == 0
    b.DeleteScanMatcher((0, 0))


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    opts, submaps, node = _queue()
    ex = OracleExecutor3D(opts)
    b = cb.ConstraintBuilder3D(opts, executor=ex, process_group=dist.group.WORLD)
    _fill(b, submaps, node)
    res = b.WhenDone(lambda r: None)
    ret[rank] = ([(c.submap_id, c.node_id, c.zbar_ij, c.score) for c in res], sorted(ex.touched),
                 b.last_records.tobytes())
    dist.destroy_process_group()


def test_3d_sharded_queue_world2_gloo():
    opts, submaps, node = _queue()
    single = cb.ConstraintBuilder3D(opts, executor=OracleExecutor3D(opts))
    _fill(single, submaps, node)
    want = [(c.submap_id, c.node_id, c.zbar_ij, c.score)
            for c in single.WhenDone(lambda r: None)]
    assert len(want) > 0
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret[0][0] == want and ret[1][0] == want   # full, identical Result on every rank
    assert ret[0][2] == ret[1][2]                    # byte-identical gathered records
    assert set(ret[0][1]).isdisjoint(ret[1][1])      # a matcher is only built on its owner
    assert sorted(ret[0][1] + ret[1][1]) == sorted(submaps.keys())
