#!/usr/bin/env python
"""Per-config measurements for BASELINE.md (configs 1, 3, 4, 5 of BASELINE.json;
config 2 is bench.py's headline).  Each config runs the engine through the C ABI
on one GPU, times it, runs a bounded sample of the same jobs through the CPU oracle
(all host threads for the batches, one thread for single matches), spot-checks
parity on that sample, and prints one JSON line.

  python benchmarks/run_configs.py [--configs 1,3,4,5] [--scale 1.0]

Sizes are the BASELINE configs scaled so that the whole script ends in minutes
(`--scale` multiplies the batch sizes); every line states the size it ran.
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cartographer_b200 import scan_matching as sm  # noqa: E402
from benchmarks import synthetic  # noqa: E402
from oracle import pyoracle as oracle  # noqa: E402  (baseline + spot checks only)

B2 = 1081 * 9 + 16
BRT = 1081 * 10 + 16


def sync():
    import torch
    torch.cuda.synchronize()


def cfg1(scale):
    """2D RealTimeCSM: 1081 beams vs 200x200 @5 cm, +-0.1 m / +-7 deg, weights 0.1/0.1."""
    big, occ = synthetic.make_grid2d(7, 1000)
    grid, occ2 = synthetic.crop_grid(big, occ, 400, 400, 200, 200)
    rng = np.random.RandomState(1)
    n = int(200 * scale)
    scans, inits = [], []
    for i in range(n):
        pose = synthetic.random_free_pose(occ2, grid, rng, margin_cells=15)
        scans.append(synthetic.cast_scan(occ2, grid, pose, seed=i, max_range=30.0))
        inits.append(pose + rng.uniform(-1, 1, 3) * [0.05, 0.05, math.radians(3)])
    opts = sm.RealTimeCorrelativeScanMatcherOptions(0.1, math.radians(7.0), 0.1, 0.1)
    rt = sm.RealTimeCorrelativeScanMatcher2D(opts)
    rt.Match(inits[0], scans[0], grid)
    sync()
    t0 = time.perf_counter()
    cand, res = 0, []
    for s, i in zip(scans, inits):
        res.append(rt.Match(i, s, grid))
        cand += rt.last_stats["candidates_scored"]
    sync()
    gpu_s = time.perf_counter() - t0
    og = oracle.Grid2D(grid.cells, grid.resolution, grid.max_x, grid.max_y)
    t0 = time.perf_counter()
    ccand, ok = 0, True
    for k in range(min(n, 50)):
        w = oracle.rt2d_match(og, scans[k], inits[k], 0.1, math.radians(7.0), 0.1, 0.1)
        ccand += w["candidates_scored"]
        ok &= (np.float32(res[k][0]) == np.float32(w["score"])) and np.array_equal(res[k][1], w["pose"])
    cpu_s = time.perf_counter() - t0
    return {"config": 1, "what": "RealTimeCSM2D 1081 beams vs 200x200@5cm, 0.1m/7deg", "matches": n,
            "gpu_cand_per_s": cand / gpu_s, "gpu_matches_per_s": n / gpu_s,
            "gpu_algorithmic_GBps": cand * BRT / gpu_s / 1e9,
            "cpu_cand_per_s_1thread": ccand / cpu_s, "cpu_matches_per_s_1thread": min(n, 50) / cpu_s,
            "parity_ok": bool(ok), "note": "through csm_rt_match2d with host buffers (grid + scan H2D per call)"}


def cfg4(scale):
    """ConstraintBuilder2D batch: submaps x nodes, local search 7 m / 30 deg / depth 7 / 0.55."""
    n_sub, n_node = max(2, int(12 * scale)), max(2, int(60 * scale))
    lin, ang, depth, min_score = 7.0, math.radians(30.0), 7, 0.55
    opts = sm.FastCorrelativeScanMatcherOptions2D(lin, ang, depth)
    worlds = [synthetic.make_grid2d(s, 1000) for s in range(n_sub)]
    rng = np.random.RandomState(3)
    t0 = time.perf_counter()
    matchers = [sm.FastCorrelativeScanMatcher2D(g, opts) for g, _ in worlds]
    sync()
    build_s = time.perf_counter() - t0
    # every node scan is cast in submap (node % n_sub)'s world; against the other
    # submaps the match is a (mostly unsuccessful) search, like a real queue
    scans, truths = [], []
    for nidx in range(n_node):
        g, occ = worlds[nidx % n_sub]
        pose = synthetic.random_free_pose(occ, g, rng)
        scans.append(synthetic.cast_scan(occ, g, pose, seed=1000 + nidx))
        truths.append(pose)
    clouds = [sm.DeviceCloud(s) for s in scans]
    jobs = np.zeros(n_sub * n_node, sm.JOB2D_DTYPE)
    k = 0
    for si in range(n_sub):
        for ni in range(n_node):
            jobs[k]["stack_index"] = si
            jobs[k]["cloud_index"] = ni
            jobs[k]["initial_pose"] = truths[ni] + rng.uniform(-1, 1, 3) * [3.0, 3.0, math.radians(15)]
            jobs[k]["min_score"] = min_score
            k += 1
    sm.match_batch(matchers, clouds, jobs, lin, ang)   # warm-up: sizes the device workspace
    sync()
    t0 = time.perf_counter()
    res, st = sm.match_batch(matchers, clouds, jobs, lin, ang)
    sync()
    gpu_s = time.perf_counter() - t0
    # per-kernel device time of one more batch (CUDA events inside the library)
    import ctypes as C
    from cartographer_b200._lib import lib
    lib().csm_profile_enable(1)
    sm.match_batch(matchers, clouds, jobs, lin, ang)
    buf = C.create_string_buffer(8192)
    lib().csm_profile_read(buf, 8192)
    lib().csm_profile_enable(0)
    prof = {}
    for ln in buf.value.decode().strip().splitlines():
        nm, n_l, t_ms, units = ln.split()
        prof[nm] = {"launches": int(n_l), "ms": round(float(t_ms), 4)}
    # CPU: bounded sample on all host threads
    threads = max(1, min(os.cpu_count() or 1, 64))
    sample = np.arange(0, len(jobs), max(1, len(jobs) // (threads * 4)))[:threads * 4]
    oms = {}
    t_build0 = time.perf_counter()
    for si in sorted({int(jobs[j]["stack_index"]) for j in sample}):
        g = worlds[si][0]
        oms[si] = oracle.FastCorrelativeScanMatcher2D(
            oracle.Grid2D(g.cells, g.resolution, g.max_x, g.max_y), lin, ang, depth)
    cpu_build_s = (time.perf_counter() - t_build0) / max(1, len(oms))
    keys = sorted(oms)
    secs, found, scores, poses, cs = oracle.fast2d_batch(
        [oms[k] for k in keys], [keys.index(int(jobs[j]["stack_index"])) for j in sample],
        [int(jobs[j]["cloud_index"]) for j in sample],
        np.array([jobs[j]["initial_pose"] for j in sample]), scans, False, min_score, threads)
    ok = True
    for i, j in enumerate(sample):
        ok &= bool(res[j]["found"]) == bool(found[i])
        if found[i]:
            ok &= np.float32(res[j]["score"]) == scores[i] and np.array_equal(res[j]["pose_estimate"], poses[i])
    return {"config": 4, "what": "ConstraintBuilder2D batch, local search 7m/30deg/depth7/min 0.55",
            "submaps": n_sub, "nodes": n_node, "jobs": len(jobs), "found": int(res["found"].sum()),
            "gpu_constraints_per_s": len(jobs) / gpu_s, "gpu_cand_per_s": st["candidates_scored"] / gpu_s,
            "gpu_algorithmic_GBps": st["candidates_scored"] * B2 / gpu_s / 1e9,
            "gpu_stack_build_ms_per_submap": 1e3 * build_s / n_sub,
            "gpu_device_ms": st["device_ms"], "gpu_wall_ms": 1e3 * gpu_s,
            "host_tie_resolves": st["host_tie_resolves"],
            "cpu_threads": threads, "cpu_sample_jobs": len(sample),
            "cpu_constraints_per_s": len(sample) / secs, "cpu_cand_per_s": float(cs.sum()) / secs,
            "cpu_stack_build_s_per_submap": cpu_build_s, "parity_ok": bool(ok),
            "gpu_kernels_one_batch": prof}


def cfg3_5(scale, which):
    """3D FastCSM: config 3 = 32 k-point VLP-16-like cloud, single matches; config 5 =
    64-ring cloud, submaps x nodes batch through ConstraintBuilder3D's executor."""
    if which == 3:
        rings, az, n_sub, n_node = 16, 2048, 1, max(2, int(6 * scale))
    else:
        rings, az, n_sub, n_node = 64, 1024, max(1, int(2 * scale)), max(2, int(12 * scale))
    # pose_graph.lua:40-48 defaults, except min_rotational_score: the synthetic
    # axis-aligned building yields two-peak histograms whose cosine at the TRUE yaw is
    # 0.5-0.9, so the default 0.77 would reject most true revisits before any scoring.
    o3 = sm.FastCorrelativeScanMatcherOptions3D(min_rotational_score=0.45)
    od = dict(branch_and_bound_depth=8, full_resolution_depth=3, min_rotational_score=0.45,
              min_low_resolution_score=0.55, linear_xy_search_window=5.0,
              linear_z_search_window=1.0, angular_search_window=math.radians(15.0))
    min_score = 0.55
    subs = []
    for s in range(n_sub):
        hi, lo, sub_hist, world = synthetic.make_submap3d(40 + s, 40.0, rings, az, 20.0)
        rng = np.random.RandomState(500 + s)
        nodes = [synthetic.make_node3d(world, rng, rings, az, 20.0, seed=7000 + 100 * s + k)
                 for k in range(n_node)]
        subs.append(dict(hi=hi, lo=lo, hist=sub_hist, nodes=nodes))
    t0 = time.perf_counter()
    ms = [sm.FastCorrelativeScanMatcher3D(s["hi"], s["lo"], s["hist"], o3) for s in subs]
    sync()
    build_s = time.perf_counter() - t0
    jobs = []
    rng = np.random.RandomState(9)
    for si, s in enumerate(subs):
        for n in s["nodes"]:
            init = n["pose"].copy()
            init[:3] += rng.uniform(-1, 1, 3) * [2.0, 2.0, 0.3]
            yaw = 2 * math.atan2(n["pose"][6], n["pose"][3]) + rng.uniform(-1, 1) * math.radians(8)
            init[3:] = [math.cos(yaw / 2), 0, 0, math.sin(yaw / 2)]
            jobs.append((si, n, init))
    ident = [0, 0, 0, 1, 0, 0, 0]
    mk = lambda n: sm.TrajectoryNodeData3D(n["cloud"], n["low"], n["hist"])
    ms[jobs[0][0]].Match(jobs[0][2], ident, mk(jobs[0][1]), min_score)
    sync()
    threads_gpu = 1 if which == 3 else 8   # config 5 = the ConstraintBuilder3D queue: 8 matches in flight
    if threads_gpu > 1:
        # one csm_match3d_batch call for the whole queue
        nodes_u, node_ix = [], {}
        for si, n, init in jobs:
            if id(n) not in node_ix:
                node_ix[id(n)] = len(nodes_u)
                nodes_u.append(mk(n))
        bjobs = [(si, node_ix[id(n)], False, init, ident, min_score) for si, n, init in jobs]
        sm.match_batch3d(ms, nodes_u, bjobs, max_concurrency=threads_gpu)   # warm-up: sizes every lane
        sync()
        t0 = time.perf_counter()
        results, bst = sm.match_batch3d(ms, nodes_u, bjobs, max_concurrency=threads_gpu)
        sync()
        gpu_s = time.perf_counter() - t0
        cand, dev_ms = bst["candidates_scored"], bst["device_ms"]
    else:
        t0 = time.perf_counter()
        pairs = [ms[si].match_raw(False, init, ident, mk(n), min_score) for si, n, init in jobs]
        sync()
        gpu_s = time.perf_counter() - t0
        results = [p[0] for p in pairs]
        cand = sum(p[1]["candidates_scored"] for p in pairs)
        dev_ms = sum(p[1]["device_ms"] for p in pairs)
    found = sum(r is not None for r in results)
    # CPU oracle on a bounded sample (single thread per match, sequential)
    sample = jobs[:max(1, min(len(jobs), 3))]
    if os.environ.get("CSM_SKIP_CPU"):   # A/B timing of kernel variants: no oracle leg
        sample = []
    ccand, ok = 0, True
    t_cpu = 0.0
    for ji, (si, n, init) in enumerate(sample):
        s = subs[si]
        ohi = oracle.HybridGrid(s["hi"].resolution, s["hi"].indices, s["hi"].values)
        olo = oracle.HybridGrid(s["lo"].resolution, s["lo"].indices, s["lo"].values)
        om = oracle.FastCorrelativeScanMatcher3D(ohi, olo, s["hist"], od)
        node = dict(gravity_alignment=(1.0, 0.0, 0.0, 0.0), high_resolution_point_cloud=n["cloud"],
                    low_resolution_point_cloud=n["low"], rotational_scan_matcher_histogram=n["hist"])
        t0 = time.perf_counter()
        w = om.match(init, ident, node, min_score)
        t_cpu += time.perf_counter() - t0
        ccand += w["candidates_scored"]
        g = results[ji]
        ok &= (g is not None) == w["found"]
        if w["found"]:
            ok &= g["score"] == w["score"] and np.array_equal(g["pose_estimate"], w["pose"])
    # per-kernel device time of ONE match that is found (CUDA events inside the library)
    import ctypes as C
    from cartographer_b200._lib import lib
    prof = {}
    for ji, r in enumerate(results):
        if r is not None:
            lib().csm_profile_enable(1)
            si, n, init = jobs[ji]
            ms[si].match_raw(False, init, ident, mk(n), min_score)
            buf = C.create_string_buffer(8192)
            lib().csm_profile_read(buf, 8192)
            lib().csm_profile_enable(0)
            for ln in buf.value.decode().strip().splitlines():
                nm, n_l, t_ms, units = ln.split()
                prof[nm] = {"launches": int(n_l), "ms": round(float(t_ms), 4)}
            break
    npts = int(np.mean([len(n["cloud"]) for _, n, _ in jobs]))
    b3 = npts * 13 + 20
    return {"config": which, "what": "FastCSM3D %d rings x %d az (~%d pts) vs HybridGrid@10cm + low-res@45cm, "
            "depth 8 / full-res 3, 5m/1m/15deg" % (rings, az, npts), "submaps": n_sub,
            "matches": len(jobs), "found": found, "gpu_matches_per_s": len(jobs) / gpu_s,
            "gpu_cand_per_s": cand / gpu_s, "gpu_algorithmic_GBps": cand * b3 / gpu_s / 1e9,
            "gpu_device_ms_per_match": dev_ms / len(jobs), "gpu_wall_ms_per_match": 1e3 * gpu_s / len(jobs),
            "gpu_matcher_build_ms": 1e3 * build_s / n_sub, "gpu_host_threads": threads_gpu,
            "kernels_of_one_found_match": prof,
            "cpu_sample_matches": len(sample),
            "cpu_matches_per_s_1thread": len(sample) / t_cpu if t_cpu else None,
            "cpu_cand_per_s_1thread": ccand / t_cpu if t_cpu else None, "parity_ok": bool(ok)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,3,4,5")
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    oracle.build()
    for c in [int(x) for x in args.configs.split(",")]:
        t0 = time.time()
        if c == 1:
            out = cfg1(args.scale)
        elif c == 4:
            out = cfg4(args.scale)
        else:
            out = cfg3_5(args.scale, c)
        out["script_seconds"] = time.time() - t0
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
