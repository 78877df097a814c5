"""ConstraintBuilder2D host logic on CPU: gating / sampling / callback semantics
(ports of constraints/constraint_builder_2d_test.cc) and the N>1 path — submap-major
sharding + the single all_gather — with world_size 2 over gloo.  The executor is
the oracle here (test infrastructure); on the GPU the same class drives
csm_match2d_batch (tests/test_gpu_constraint_builder.py).
"""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cartographer_b200 import constraint_builder as cb
from benchmarks import synthetic


class OracleExecutor:
    """Runs jobs through the CPU oracle; records which submaps this rank touched."""

    def __init__(self, options):
        from oracle import pyoracle
        pyoracle.build()
        self.o = pyoracle
        self.options = options
        self.matchers = {}
        self.touched = set()

    def run(self, jobs, submaps, clouds):
        out = []
        opt = self.options
        for j in jobs:
            self.touched.add(j.submap_id)
            if j.submap_id not in self.matchers:
                g = submaps[j.submap_id].grid
                og = self.o.Grid2D(g.cells, g.resolution, g.max_x, g.max_y)
                self.matchers[j.submap_id] = self.o.FastCorrelativeScanMatcher2D(
                    og, opt.linear_search_window, opt.angular_search_window,
                    opt.branch_and_bound_depth)
            m = self.matchers[j.submap_id]
            r = (m.match_full_submap(clouds[j.cloud_key], j.min_score) if j.full else
                 m.match(j.initial_pose, clouds[j.cloud_key], j.min_score))
            pose = tuple(r["pose"])
            co = opt.ceres_scan_matcher_options
            if r["found"] and co is not None:   # constraint_builder_2d.cc:245-249
                og = self.matchers[j.submap_id].grid
                pose = tuple(self.o.ceres2d_match(
                    og, clouds[j.cloud_key], pose[:2], pose, co.occupied_space_weight,
                    co.translation_weight, co.rotation_weight, co.use_nonmonotonic_steps,
                    co.max_num_iterations)["pose"])
            out.append((r["found"], float(r["score"]), pose))
        return out

    def delete_matcher(self, submap_id):
        self.matchers.pop(submap_id, None)


def _options(**kw):
    base = dict(sampling_ratio=1.0, max_constraint_distance=10.0, min_score=0.0,
                global_localization_min_score=0.0, linear_search_window=0.5,
                angular_search_window=0.2, branch_and_bound_depth=3)
    base.update(kw)
    return cb.ConstraintBuilderOptions(**base)


def _empty_submap():
    # constraint_builder_2d_test.cc:70-112: 100 x 110 empty grid, one-point cloud,
    # min_score 0 so every search "finds".
    cells = np.zeros((110, 100), np.uint16)
    return cb.Submap2D(synthetic.GridSpec(cells, 0.1, 5.0, 5.5), (0.0, 0.0, 0.0))


def test_calls_back():
    """constraint_builder_2d_test.cc:58-68."""
    b = cb.ConstraintBuilder2D(_options(), executor=OracleExecutor(_options()))
    assert b.GetNumFinishedNodes() == 0
    b.NotifyEndOfNode()
    got = []
    b.WhenDone(lambda result: got.append(result))
    assert got == [[]]
    assert b.GetNumFinishedNodes() == 1


def test_finds_constraints():
    """constraint_builder_2d_test.cc:70-112: 3 INTER_SUBMAP constraints per round."""
    opts = _options()
    b = cb.ConstraintBuilder2D(opts, executor=OracleExecutor(opts))
    cloud = np.array([[0.1, 0.2, 0.3]], np.float32)
    submap = _empty_submap()
    expected_nodes = 0
    for i in range(2):
        assert b.GetNumFinishedNodes() == expected_nodes
        for j in range(2):
            b.MaybeAddConstraint((0, 0), submap, (0, 0), cloud, (0.0, 0.0, 0.0))
        b.MaybeAddGlobalConstraint((0, 0), submap, (0, 0), cloud)
        b.NotifyEndOfNode()
        got = []
        b.WhenDone(lambda result: got.append(result))
        expected_nodes += 1
        assert b.GetNumFinishedNodes() == expected_nodes
        assert len(got[0]) == 3
        for c in got[0]:
            assert c.tag == "INTER_SUBMAP"
            assert c.submap_id == (0, 0) and c.node_id == (0, 0)
            assert c.translation_weight == opts.loop_closure_translation_weight
        b.DeleteScanMatcher((0, 0))


def test_distance_gate_and_sampler():
    """constraint_builder_2d.cc:81-90 and common/fixed_ratio_sampler.cc:32-39."""
    opts = _options(sampling_ratio=0.3, max_constraint_distance=2.0)
    b = cb.ConstraintBuilder2D(opts, executor=OracleExecutor(opts))
    cloud = np.array([[0.1, 0.2, 0.0]], np.float32)
    submap = _empty_submap()
    b.MaybeAddConstraint((0, 0), submap, (0, 1), cloud, (3.0, 0.0, 0.0))  # too far: no pulse
    assert len(b._jobs) == 0
    kept = 0
    for n in range(100):
        before = len(b._jobs)
        b.MaybeAddConstraint((0, 0), submap, (0, n), cloud, (0.1, 0.0, 0.0))
        kept += len(b._jobs) - before
    assert kept == 30  # FixedRatioSampler(0.3) keeps exactly 30 of 100 pulses
    s = cb.FixedRatioSampler(0.3)
    assert [s.Pulse() for _ in range(10)] == [True, False, False, True, False, False, True,
                                              False, False, False]
    b.NotifyEndOfNode()
    b.WhenDone(lambda r: None)


def _small_queue():
    opts = _options(linear_search_window=1.0, angular_search_window=0.3,
                    branch_and_bound_depth=4, min_score=0.3, global_localization_min_score=0.3)
    submaps, clouds, poses = {}, [], []
    rng = np.random.RandomState(5)
    for s in range(4):
        grid, occ = synthetic.make_grid2d(500 + s, size_cells=120)
        submaps[(0, s)] = (cb.Submap2D(grid, (0.3 * s, -0.2 * s, 0.1 * s)), occ)
    for n in range(3):
        g, occ = submaps[(0, n)][0].grid, submaps[(0, n)][1]
        pose = synthetic.random_free_pose(occ, g, rng, margin_cells=10)
        clouds.append(synthetic.cast_scan(occ, g, pose, beams=91, max_range=5.0, seed=n))
        poses.append(pose)
    return opts, submaps, clouds, poses


def _fill(builder, submaps, clouds, poses):
    for n, cloud in enumerate(clouds):
        for sid, (sub, _) in submaps.items():
            rel = cb._compose(cb._inverse(tuple(sub.local_pose)), tuple(poses[n]))
            builder.MaybeAddConstraint(sid, sub, (0, n), cloud, rel)
        builder.MaybeAddGlobalConstraint((0, n % 4), submaps[(0, n % 4)][0], (0, n), cloud)
        builder.NotifyEndOfNode()


def _worker(rank, world, port, ret, refine=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    opts, submaps, clouds, poses = _small_queue()
    if refine:
        from cartographer_b200 import scan_matching as sm
        opts.ceres_scan_matcher_options = sm.CeresScanMatcherOptions2D()
    ex = OracleExecutor(opts)
    b = cb.ConstraintBuilder2D(opts, executor=ex, process_group=dist.group.WORLD)
    _fill(b, submaps, clouds, poses)
    res = b.WhenDone(lambda r: None)
    ret[rank] = ([(c.submap_id, c.node_id, c.zbar_ij, c.score) for c in res],
                 sorted(ex.touched), b.last_records.tobytes())
    dist.destroy_process_group()


@pytest.mark.parametrize("refine", [False, True])
def test_sharded_queue_world2_gloo(refine):
    """refine=True: the owner of a submap refines its found matches BEFORE the single
    all-gather, so every rank still ends with the single-process Result."""
    opts, submaps, clouds, poses = _small_queue()
    if refine:
        from cartographer_b200 import scan_matching as sm
        opts.ceres_scan_matcher_options = sm.CeresScanMatcherOptions2D()
    single = cb.ConstraintBuilder2D(opts, executor=OracleExecutor(opts))
    _fill(single, submaps, clouds, poses)
    want = single.WhenDone(lambda r: None)
    assert len(want) > 0
    want_list = [(c.submap_id, c.node_id, c.zbar_ij, c.score) for c in want]

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret, refine)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    # every rank holds the full, identical Result in queue order == single-process Result
    assert ret[0][0] == want_list and ret[1][0] == want_list
    assert ret[0][2] == ret[1][2]
    # submap-major sharding: disjoint owners, stacks only built where owned
    assert set(ret[0][1]).isdisjoint(ret[1][1])
    assert sorted(ret[0][1] + ret[1][1]) == sorted(submaps.keys())


def test_ceres_refinement_changes_only_the_poses():
    """With ceres_scan_matcher_options set the queue yields the same constraints (ids,
    scores, order) with refined poses (constraint_builder_2d.cc:245-258); host logic only —
    the executor is the CPU oracle."""
    from cartographer_b200 import scan_matching as sm
    opts, submaps, clouds, poses = _small_queue()
    plain = cb.ConstraintBuilder2D(opts, executor=OracleExecutor(opts))
    _fill(plain, submaps, clouds, poses)
    want = plain.WhenDone(lambda r: None)
    opts2, _, _, _ = _small_queue()
    opts2.ceres_scan_matcher_options = sm.CeresScanMatcherOptions2D()
    refined = cb.ConstraintBuilder2D(opts2, executor=OracleExecutor(opts2))
    _fill(refined, submaps, clouds, poses)
    got = refined.WhenDone(lambda r: None)
    assert len(got) == len(want) > 0
    moved = 0
    for a, b in zip(got, want):
        assert a.submap_id == b.submap_id and a.node_id == b.node_id and a.score == b.score
        d = np.abs(np.array(a.zbar_ij) - np.array(b.zbar_ij))
        assert d[:2].max() < 0.2 and d[2] < 0.1        # a refinement, not a new search
        moved += int(d.max() > 0)
    assert moved > 0
