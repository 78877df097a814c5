#!/usr/bin/env python
"""bench.py — headline benchmark of the correlative scan-matching hot path.

Metric (BASELINE.json): candidate poses scored / s (+ loop-closure constraints / s).
Workload at every N: BASELINE config[1] — 2D FastCorrelativeScanMatcher
MatchFullSubmap, 1081-beam synthetic scans vs a 1000x1000 @5 cm ProbabilityGrid,
depth-7 PrecomputationGridStack.  A step = MATCHES_PER_STEP full-submap matches of
distinct scans against the rank's submap (one csm_match2d_batch call); at N > 1
every rank owns its own submap + scans (weak scaling, the ConstraintBuilder queue
sharded by submap) and the winning constraints are all-gathered over NCCL once per
step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`value`  : device-resident inputs (clouds + stack in HBM before the timed region).
`e2e`    : the same matches through csm_match2d with HOST clouds (H2D + D2H inside).
`--impl reference`: the CPU oracle (restated reference path; the real reference does
not build here) on the host cores, same metric / workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from benchmarks import synthetic  # noqa: E402

BYTES_PER_CANDIDATE = 1081 * (8 + 1) + 16  # N*(8+1)+16 @ N=1081
MATCHES_PER_STEP = 16
MIN_SCORE = 0.6          # pose_graph.lua:28 global_localization_min_score
DEPTH = 7                # pose_graph.lua:27
LIN, ANG = 7.0, math.radians(30.0)
WORKLOAD = "fast2d_MatchFullSubmap_1081beams_1000x1000_5cm_depth7"


def make_world(seed, num_scans):
    """Every rank owns a submap of the same synthetic floor plan (so that weak scaling
    compares like with like) and matches its own node scans (poses and noise seeded
    by `seed` = rank) against it."""
    grid, occ = synthetic.make_grid2d(0, 1000)
    rng = np.random.RandomState(seed * 1000 + 17)
    scans = []
    for i in range(num_scans):
        pose = synthetic.random_free_pose(occ, grid, rng)
        scans.append(synthetic.cast_scan(occ, grid, pose, seed=seed * 100000 + i))
    return grid, scans


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device = device
        self.rows = []
        self.stop_flag = False

    def run(self):
        try:
            p = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                  "--format=csv,noheader,nounits", "-lms", "100"],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.proc = p
        for line in p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])
            if self.stop_flag:
                break
        p.terminate()

    def summary(self):
        self.stop_flag = True
        time.sleep(0.15)
        if getattr(self, "proc", None):
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def cpu_reference_step(oracle, og_matcher, scans, threads):
    """One bounded CPU sample: `threads` full-submap matches, one per worker thread."""
    jobs = list(range(len(scans)))
    secs, found, scores, poses, cs = oracle.fast2d_batch(
        [og_matcher], [0] * len(jobs), jobs, np.zeros((len(jobs), 3)), scans, True, MIN_SCORE,
        threads)
    return secs, int(cs.sum()), len(jobs), int(found.sum())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as oracle
    oracle.build()
    threads = max(1, min(os.cpu_count() or 1, 64))
    per_step = threads  # one full-submap match per host thread per step (bounded sample)
    grid, scans = make_world(0, per_step)
    og = oracle.Grid2D(grid.cells, grid.resolution, grid.max_x, grid.max_y)
    t0 = time.perf_counter()
    om = oracle.FastCorrelativeScanMatcher2D(og, LIN, ANG, DEPTH)
    build_s = time.perf_counter() - t0
    tot_s, tot_c, tot_m = 0.0, 0, 0
    for it in range(args.warmup + args.steps):
        secs, cands, matches, _ = cpu_reference_step(oracle, om, scans, threads)
        if it >= args.warmup:
            tot_s += secs
            tot_c += cands
            tot_m += matches
    value = tot_c / tot_s
    line = {
        "impl": "reference", "metric": "candidate_poses_scored_per_sec", "value": value,
        "unit": "candidates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_s / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "matches_per_step": per_step,
                   "min_score": MIN_SCORE, "oracle_stack_build_s": build_s},
        "constraints_per_sec": tot_m / tot_s,
        "cpu_baseline": {"value": value, "unit": "candidates/s", "cores": threads,
                         "kind": "port",
                         "sample": "%d MatchFullSubmap per step, one per host thread, "
                                   "oracle/ (C++ restatement, -O3 -DNDEBUG, no -march)" % per_step},
        "e2e": {"value": value, "unit": "candidates/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    from cartographer_b200 import scan_matching as sm
    from cartographer_b200._lib import lib

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    total_steps = args.warmup + args.steps
    num_scans = total_steps * MATCHES_PER_STEP
    grid, scans = make_world(rank, num_scans + MATCHES_PER_STEP)
    opts = sm.FastCorrelativeScanMatcherOptions2D(LIN, ANG, DEPTH)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    matcher = sm.FastCorrelativeScanMatcher2D(grid, opts, device=local_rank)
    stack_build_ms = 1e3 * (time.perf_counter() - t0)
    clouds = [sm.DeviceCloud(s, device=local_rank) for s in scans]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def jobs_for(step):
        jobs = np.zeros(MATCHES_PER_STEP, sm.JOB2D_DTYPE)
        for b in range(MATCHES_PER_STEP):
            jobs[b]["stack_index"] = 0
            jobs[b]["cloud_index"] = step * MATCHES_PER_STEP + b
            jobs[b]["full_submap"] = 1
            jobs[b]["min_score"] = MIN_SCORE
        return jobs

    gathered = None
    if dist is not None:
        gathered = torch.empty(world * MATCHES_PER_STEP * sm.RESULT2D_DTYPE.itemsize,
                               dtype=torch.uint8, device=dev)

    def allgather_results(res):
        """The path's only collective: every rank ends up with all constraints."""
        if dist is None:
            return res
        mine = torch.from_numpy(res.view(np.uint8).reshape(-1).copy()).to(dev, non_blocking=False)
        dist.all_gather_into_tensor(gathered, mine)
        return gathered

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def finish():
        # end of a step on this rank: the allgather (a collective, so it already waits
        # for every rank's results) has completed and the device is idle
        torch.cuda.synchronize()

    # ---- device-resident leg (value) -------------------------------------------
    # No poller runs during the timed legs (an nvidia-smi loop contends on the driver
    # lock with kernel launches); the clocks are sampled in a separate pass that repeats
    # the same steps right after the timed ones (see `clocks.note`).
    launches0 = sm.kernel_launch_count()
    step_s, cand, found = [], 0, 0
    dev_ms = 0.0
    host_syncs = 0
    results_by_step = []
    for it in range(total_steps):
        flush.zero_()
        barrier()
        if it == args.warmup:
            launches0 = sm.kernel_launch_count()
        t0 = time.perf_counter()
        res, st = sm.match_batch([matcher], clouds, jobs_for(it), LIN, ANG)
        allgather_results(res)
        finish()
        dt = time.perf_counter() - t0
        results_by_step.append(res.copy())
        if it >= args.warmup:
            step_s.append(dt)
            cand += st["candidates_scored"]
            found += int(res["found"].sum())
            dev_ms += st["device_ms"]
            host_syncs = max(host_syncs, st["host_syncs"])
    launches = sm.kernel_launch_count() - launches0
    elapsed = float(sum(step_s))
    t = torch.tensor([elapsed, float(cand), float(found)], dtype=torch.float64, device=dev)
    if dist is not None:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        elapsed, cand_all, found_all = float(tmax[0]), float(tsum[1]), float(tsum[2])
    else:
        cand_all, found_all = float(cand), float(found)
    matches_all = world * args.steps * MATCHES_PER_STEP
    value = cand_all / elapsed

    # ---- end-to-end leg (host clouds through csm_match2d) ------------------------
    e2e_s, e2e_c = [], 0
    for it in range(total_steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        # the step's scans start in HOST memory: upload them (csm_cloud_create = H2D),
        # run the batch, read the results back, release the device copies
        step_clouds = [sm.DeviceCloud(scans[it * MATCHES_PER_STEP + b], device=local_rank)
                       for b in range(MATCHES_PER_STEP)]
        jobs = jobs_for(0)
        res, st_e = sm.match_batch([matcher], step_clouds, jobs, LIN, ANG)
        c_step = st_e["candidates_scored"]
        for c in step_clouds:
            c.close()
        allgather_results(res)
        finish()
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            e2e_s.append(dt)
            e2e_c += c_step
    te = torch.tensor([float(sum(e2e_s)), float(e2e_c)], dtype=torch.float64, device=dev)
    if dist is not None:
        a = te.clone()
        dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = te.clone()
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        e2e_value = float(b[1]) / float(a[0])
    else:
        e2e_value = float(te[1]) / float(te[0])
    h2d = MATCHES_PER_STEP * 1081 * 12
    d2h = MATCHES_PER_STEP * sm.RESULT2D_DTYPE.itemsize

    # ---- clocks under load (separate pass over the same steps, rank 0's GPU) -------
    clocks = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    if rank == 0:
        sampler = ClockSampler(local_rank)
        sampler.start()
        t_end = time.perf_counter() + 1.5
        it = 0
        while time.perf_counter() < t_end or it < args.steps:
            sm.match_batch([matcher], clouds, jobs_for(args.warmup + it % max(1, args.steps)), LIN, ANG)
            it += 1
        clocks = sampler.summary()
        clocks["note"] = ("sampled with nvidia-smi -lms 100 while the same steps ran again right "
                          "after the timed region (no poller inside the timed legs)")
    barrier()

    # ---- roofline of the dominant kernel (CUDA events on the engine's stream) ----
    import ctypes as C
    roofline = None
    if rank == 0:
        lib().csm_profile_enable(1)
        sm.match_batch([matcher], clouds, jobs_for(args.warmup), LIN, ANG)
        buf = C.create_string_buffer(8192)
        lib().csm_profile_read(buf, 8192)
        lib().csm_profile_enable(0)
        kernels = {}
        for ln in buf.value.decode().strip().splitlines():
            nm, n_l, ms, units = ln.split()
            kernels[nm] = {"launches": int(n_l), "ms": float(ms), "units": float(units)}
        tot_ms = sum(k["ms"] for k in kernels.values())
        top = max(kernels, key=lambda k: kernels[k]["ms"])
        k = kernels[top]
        peak, how = measured_peak_gbs()
        ach = k["units"] * BYTES_PER_CANDIDATE / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else 0.0
        traffic = None
        try:  # DRAM bytes per launch of this kernel from the committed ncu capture
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                traffic = json.load(f).get(top)
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak,
                    "peak_source": how, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "traffic_note": "ncu dram bytes of the largest launch of this kernel; the "
                                    "working set is L2-resident, so DRAM traffic << algorithmic bytes",
                    "launches": k["launches"],
                    "avg_launch_ms": k["ms"] / max(1, k["launches"]),
                    "share_of_step": k["ms"] / tot_ms if tot_ms else None,
                    "bytes_per_candidate": BYTES_PER_CANDIDATE,
                    "kernels": {n: {"ms": round(v["ms"], 4), "launches": v["launches"],
                                    "candidates": v["units"]} for n, v in kernels.items()}}

    # ---- CPU baseline (oracle on the host cores, bounded sample) -----------------
    cpu = None
    parity_checked = parity_failed = 0
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import pyoracle as oracle
        oracle.build()
        threads = max(1, min(os.cpu_count() or 1, 64))
        og = oracle.Grid2D(grid.cells, grid.resolution, grid.max_x, grid.max_y)
        om = oracle.FastCorrelativeScanMatcher2D(og, LIN, ANG, DEPTH)
        sample = scans[:threads] if len(scans) >= threads else (scans * threads)[:threads]
        secs, found_c, scores_c, poses_c, cs_c = oracle.fast2d_batch(
            [om], [0] * len(sample), list(range(len(sample))), np.zeros((len(sample), 3)), sample,
            True, MIN_SCORE, threads)
        c_cpu, m_cpu = int(cs_c.sum()), len(sample)
        # the same scans went through the engine in the timed steps: compare bit for bit
        for k in range(min(len(sample), len(scans), total_steps * MATCHES_PER_STEP)):
            g = results_by_step[k // MATCHES_PER_STEP][k % MATCHES_PER_STEP]
            ok = bool(g["found"]) == bool(found_c[k])
            if ok and found_c[k]:
                ok = (np.float32(g["score"]) == scores_c[k] and
                      np.array_equal(g["pose_estimate"], poses_c[k]))
            parity_checked += 1
            parity_failed += 0 if ok else 1
        cpu = {"value": c_cpu / secs, "unit": "candidates/s", "cores": threads, "kind": "port",
               "constraints_per_sec": m_cpu / secs,
               "sample": "%d MatchFullSubmap (one per host thread) of the same workload, "
                         "%.1f s wall" % (m_cpu, secs)}

    if rank == 0:
        line = {
            "metric": "candidate_poses_scored_per_sec", "value": value, "unit": "candidates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * elapsed / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "matches_per_step_per_gpu": MATCHES_PER_STEP,
                       "world": "one submap of the same synthetic floor plan per rank, per-rank node scans",
                       "min_score": MIN_SCORE, "l2": "flushed between steps (256 MB write)",
                       "stack_build_ms": stack_build_ms, "found": found_all,
                       "parallelism": "submap-sharded x%d" % world},
            "constraints_per_sec": matches_all / elapsed,
            "device_ms_per_step": dev_ms / max(1, args.steps),
            "e2e": {"value": e2e_value, "unit": "candidates/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "host_syncs_per_batch": host_syncs, "clocks": clocks,
            "parity_checked": parity_checked, "parity_failed": parity_failed,
            "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
