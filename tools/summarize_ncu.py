#!/usr/bin/env python
"""Summarise ncu output into small text files for profiles/.
  launch list : python tools/summarize_ncu.py launches <launches.csv>
  full capture: python tools/summarize_ncu.py full <file.ncu-rep>
  roofline    : python tools/summarize_ncu.py roofline <file.ncu-rep> <bench.json> <out.json>
                per-step totals (first captured step) of the counters bench.py's roofline
                block needs: L1 data-pipe LSU wavefronts, L2 bytes, DRAM bytes, time.
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui].strip(), 1e-6)
        n = r[ki].split("(")[0]
        agg[n][0] += 1
        agg[n][1] += v * scale
    tot = sum(v[1] for v in agg.values())
    print("%-64s %8s %12s %7s" % ("kernel", "launches", "total ms", "share"))
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-64s %8d %12.4f %6.1f%%" % (n[:64], c, t, 100 * t / tot))
    print("%-64s %8d %12.4f" % ("TOTAL", sum(v[0] for v in agg.values()), tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[h.index("Kernel Name")][:100])
        for k in KEYS:
            if k in h:
                print("  %-82s %s %s" % (k, r[h.index(k)], units[h.index(k)]))
        print()


def roofline(path, bench_json, out_path):
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]

    def col(r, k, default=0.0):
        if k not in h:
            return default
        v = r[h.index(k)].replace(",", "")
        try:
            x = float(v)
        except ValueError:
            return default
        u = units[h.index(k)].strip()
        return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "ns": 1e-6, "s": 1e3,
                    "ms": 1.0}.get(u, 1.0)

    # launches in capture order; a step starts at every lowest-resolution pass
    def short(full_name):
        n = full_name.split("(")[0].replace("void ", "")
        base = n.split("<")[0]
        if base == "k_rt_match" and "<" in n:   # the library profiles the three forms by name
            form = n.split("<")[1].split(",")[0].strip()
            base += {"0": "_tma", "1": "_gather", "2": "_tsdf"}.get(form, "")
        return base
    launches_ = [(short(r[h.index("Kernel Name")]), r) for r in rows[2:]]
    starts = [i for i, (n, _) in enumerate(launches_) if n.startswith("k_score_top")]
    lo = starts[0] if starts else 0
    hi = starts[1] if len(starts) > 1 else len(launches_)
    bench = json.loads(open(bench_json).read().strip().splitlines()[-1])
    live = bench.get("roofline", {}).get("kernels", {})
    res = {}
    for name, r in launches_[lo:hi]:
        e = res.setdefault(name, {"launches_per_step": 0, "time_ms": 0.0,
                                  "l1_lsu_wavefronts_per_step": 0.0, "lts_bytes_per_step": 0.0,
                                  "dram_bytes_per_step": 0.0, "_pct_w": 0.0})
        t = col(r, "gpu__time_duration.sum")
        e["launches_per_step"] += 1
        e["time_ms"] += t
        # one data-pipe wavefront per clock per SM is the peak, so
        # wavefronts = pct / 100 * elapsed SM cycles * 148 (the raw .sum counter is not in --set full)
        e["l1_lsu_wavefronts_per_step"] += (
            col(r, "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed") / 100.0 *
            col(r, "sm__cycles_elapsed.avg") * 148)
        e["issue_pct_w"] = e.get("issue_pct_w", 0.0) + t * col(
            r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
        e["inst_per_step"] = e.get("inst_per_step", 0.0) + col(r, "smsp__inst_executed.sum")
        e["lts_bytes_per_step"] += 32.0 * col(r, "lts__t_sectors.sum")
        e["dram_bytes_per_step"] += col(r, "dram__bytes_read.sum") + col(r, "dram__bytes_write.sum")
        e["_pct_w"] += t * col(r, "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed")
    for name, e in res.items():
        e["l1_lsu_pct"] = e.pop("_pct_w") / e["time_ms"] if e["time_ms"] else None
        e["issue_active_pct"] = e.pop("issue_pct_w", 0.0) / e["time_ms"] if e["time_ms"] else None
        e["units_per_step"] = live.get(name, {}).get("candidates")
        e["source"] = ("ncu --set full --clock-control none, first captured step of "
                       "`python bench.py --steps 1 --warmup 1` (%s); units_per_step from the "
                       "same session's bench line" % path.split("/")[-1])
    import os
    merged = {}
    if os.path.exists(out_path):   # several captures (one per workload) share the file
        merged = json.load(open(out_path))
    merged.update(res)
    json.dump(merged, open(out_path, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "roofline":
        roofline(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        {"launches": launches, "full": full}[cmd](sys.argv[2])
