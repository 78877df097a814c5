#!/usr/bin/env python
"""CPU study (no product code): how many node evaluations an exact branch and bound needs
on the config-2 workload if nodes also bound GROUPS of neighbouring rotations.

Today's tree (the reference's, fast_correlative_scan_matcher_2d.cc:335-378): every rotated
scan is a separate root set; a node (scan k, block 2^h) is bounded with precomputation
level h.  Joint tree: a node (2^g rotations around k_c, block 2^h), g <= h-1, is bounded
with level h+1 read at the centre rotation's cells shifted by -D, D = 2^(g-1) + 1: between
neighbouring rotations a point moves less than one cell (that is how the angular step is
chosen, correlative_scan_matcher_2d.cc:39-45), so over the group its cell stays within +-D
of the centre rotation's cell and the window 2^h + 2D <= 2^(h+1) covers block and sweep.

Both counts below assume the bound S* is known (the engine's dives deliver it), i.e. they
count the nodes ANY exact search must evaluate in that tree: all nodes whose parent's
bound is >= S*.  Output: evaluations per level for both trees.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench                      # noqa: E402  (workload generator)
from oracle import pyoracle as oracle  # noqa: E402


def level_grid(cells, h):
    """PrecomputationGrid2D of width 2^h as int array with its offset (wide grid)."""
    w = 1 << h
    pg = oracle.precompute_grid2d(cells, oracle.constant(2), oracle.constant(3), w)
    return pg.astype(np.int32), w - 1   # value(x, y) = pg[y + off, x + off], 0 outside


def score(pg, off, pts, xo, yo):
    """sum over points of GetValue(p + (xo, yo)); pts: (n, 2) cells; xo, yo: (m,) offsets."""
    wy, wx = pg.shape
    out = np.zeros(len(xo), np.int64)
    for a in range(0, len(xo), 4096):
        x = pts[None, :, 0] + xo[a:a + 4096, None] + off
        y = pts[None, :, 1] + yo[a:a + 4096, None] + off
        ok = (x >= 0) & (x < wx) & (y >= 0) & (y < wy)
        v = pg[np.clip(y, 0, wy - 1), np.clip(x, 0, wx - 1)]
        out[a:a + 4096] = (v * ok).sum(1)
    return out


def study(grid, cloud, init, full, lin, ang, depth, min_score, schedules, verbose=True):
    og = oracle.Grid2D(grid.cells, grid.resolution, grid.max_x, grid.max_y)
    om = oracle.FastCorrelativeScanMatcher2D(og, lin, ang, depth)
    want = om.match_full_submap(cloud, min_score) if full else om.match(init, cloud, min_score)
    if not want["found"]:
        return None
    n = len(cloud)
    # integer threshold equivalent to score >= S*
    fe = oracle.frontend2d(og, cloud, init, full=full, lin=lin, ang=ang)
    ds = fe["discrete_scans"].astype(np.int64)     # (S, n, 2)
    bounds = fe["bounds"].astype(np.int64)         # (S, 4) min_x max_x min_y max_y
    S = len(ds)
    levels = [level_grid(grid.cells, h) for h in range(depth + 1)]   # one more level than the stack
    # find S* as an integer sum: best leaf sum over the whole window is what the oracle found
    k, bx, by = want["best_scan_index"], want["best_x_offset"], want["best_y_offset"]
    s_star = int(score(levels[0][0], levels[0][1], ds[k], np.array([bx]), np.array([by]))[0])
    if verbose:
        print("match: scans %d, points %d, S* sum %d (score %.4f), found %s" %
              (S, n, s_star, want["score"], want["found"]))
    t0 = time.time()

    # ---- today's tree: per scan, top lattice at h = depth-1, expand nodes with bound >= S* ----
    top = depth - 1
    cur = []   # (scan, xo, yo) arrays per scan
    evals_now = {}
    tot = 0
    front = []
    for k in range(S):
        mnx, mxx, mny, mxy = bounds[k]
        xs = np.arange(mnx, mxx + 1, 1 << top)
        ys = np.arange(mny, mxy + 1, 1 << top)
        X, Y = np.meshgrid(xs, ys, indexing="ij")
        xo, yo = X.ravel(), Y.ravel()
        sc = score(levels[top][0], levels[top][1], ds[k], xo, yo)
        tot += len(xo)
        keep = sc >= s_star
        front.append((np.full(keep.sum(), k), xo[keep], yo[keep]))
    evals_now[top] = tot
    fk = np.concatenate([f[0] for f in front]); fx = np.concatenate([f[1] for f in front])
    fy = np.concatenate([f[2] for f in front])
    for h in range(top, 0, -1):
        half = 1 << (h - 1)
        nk, nx, ny = [], [], []
        cnt = 0
        for k in np.unique(fk):
            m = fk == k
            px, py = fx[m], fy[m]
            mnx, mxx, mny, mxy = bounds[k]
            for dx in (0, half):
                for dy in (0, half):
                    cx, cy = px + dx, py + dy
                    ok = (cx <= mxx) & (cy <= mxy)
                    cx, cy = cx[ok], cy[ok]
                    sc = score(levels[h - 1][0], levels[h - 1][1], ds[k], cx, cy)
                    cnt += len(cx)
                    keep = sc >= s_star
                    nk.append(np.full(keep.sum(), k)); nx.append(cx[keep]); ny.append(cy[keep])
        evals_now[h - 1] = cnt
        fk, fx, fy = np.concatenate(nk), np.concatenate(nx), np.concatenate(ny)
    leaves_now = set(zip(fk.tolist(), fx.tolist(), fy.tolist()))
    if verbose:
        print("today's tree   : evaluations per level (top..0):",
              [evals_now[h] for h in range(top, -1, -1)], "total", sum(evals_now.values()),
              "optimal leaves", len(fk), "[%.0f s]" % (time.time() - t0))
    # the oracle's answer is one of the optimal leaves
    assert (want["best_scan_index"], want["best_x_offset"], want["best_y_offset"]) in leaves_now

    # ---- joint trees: a schedule of (g, h) node types, g = log2(rotations per node) ----
    def run_schedule(schedule):
        """schedule: [(g0, h0), (g1, h1), ...] ending in (0, 0); each step keeps or lowers
        g and h; a node (g, h) with g >= 1 is bounded with level h+1 at shift -D."""
        evals = {}
        g, h = schedule[0]
        assert h == top
        cnt = 0
        fr_k, fr_x, fr_y = [], [], []
        for k_lo in range(0, S, 1 << g):
            k_hi = min(S, k_lo + (1 << g))
            if g == 0:
                kc, D, lvl = k_lo, 0, h
            else:
                kc, D, lvl = min(S - 1, k_lo + (1 << (g - 1))), (1 << (g - 1)) + 1, h + 1
            mnx, mxx = bounds[k_lo:k_hi, 0].min(), bounds[k_lo:k_hi, 1].max()
            mny, mxy = bounds[k_lo:k_hi, 2].min(), bounds[k_lo:k_hi, 3].max()
            xs = np.arange(mnx, mxx + 1, 1 << h)
            ys = np.arange(mny, mxy + 1, 1 << h)
            X, Y = np.meshgrid(xs, ys, indexing="ij")
            xo, yo = X.ravel(), Y.ravel()
            sc = score(levels[lvl][0], levels[lvl][1], ds[kc], xo - D, yo - D)
            cnt += len(xo)
            keep = sc >= s_star
            fr_k.append(np.full(keep.sum(), k_lo)); fr_x.append(xo[keep]); fr_y.append(yo[keep])
        evals[(g, h)] = cnt
        fk, fx, fy = np.concatenate(fr_k), np.concatenate(fr_x), np.concatenate(fr_y)
        for (g2, h2) in schedule[1:]:
            assert g2 <= g and h2 <= h and (g2 == 0 or 2 * ((1 << (g2 - 1)) + 1) <= (1 << h2))
            sub_step = 1 << g2
            n_sub = 1 << (g - g2)
            dxy = [0] if h2 == h else list(range(0, 1 << h, 1 << h2))
            nk, nx, ny = [], [], []
            cnt = 0
            for k_lo in np.unique(fk):
                m = fk == k_lo
                px, py = fx[m], fy[m]
                for t in range(n_sub):
                    ks = k_lo + t * sub_step
                    if ks >= S:
                        continue
                    k_hi = min(S, ks + sub_step)
                    if g2 == 0:
                        kc, D, lvl = ks, 0, h2
                    else:
                        kc, D, lvl = (min(S - 1, ks + (1 << (g2 - 1))), (1 << (g2 - 1)) + 1,
                                      h2 + 1)
                    mxx = bounds[ks:k_hi, 1].max(); mxy = bounds[ks:k_hi, 3].max()
                    mnx = bounds[ks:k_hi, 0].min(); mny = bounds[ks:k_hi, 2].min()
                    for dx in dxy:
                        for dy in dxy:
                            cx, cy = px + dx, py + dy
                            ok = (cx <= mxx) & (cy <= mxy)
                            ok &= (cx + (1 << h2) > mnx) & (cy + (1 << h2) > mny)
                            if g2 == 0 and h2 == 0:
                                ok &= (cx >= mnx) & (cy >= mny)
                            cx, cy = cx[ok], cy[ok]
                            sc = score(levels[lvl][0], levels[lvl][1], ds[kc], cx - D, cy - D)
                            cnt += len(cx)
                            keep = sc >= s_star
                            nk.append(np.full(keep.sum(), ks)); nx.append(cx[keep]); ny.append(cy[keep])
            evals[(g2, h2)] = evals.get((g2, h2), 0) + cnt
            fk, fx, fy = np.concatenate(nk), np.concatenate(nx), np.concatenate(ny)
            g, h = g2, h2
        return evals, set(zip(fk.tolist(), fx.tolist(), fy.tolist()))

    out = {}
    for name, sch in schedules.items():
        t0 = time.time()
        ev, leaves = run_schedule(sch)
        out[name] = (sum(ev.values()), leaves == leaves_now)
        if verbose:
            print("%-26s total %8d  same leaves %s  per (g,h): %s  [%.0f s]" %
                  (name, sum(ev.values()), leaves == leaves_now, ev, time.time() - t0))
    return out


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--check":
        # exactness check on many small worlds (local and full-submap searches, depth 5)
        from tests import worlds
        sch = {"reference order": [(0, 4), (0, 3), (0, 2), (0, 1), (0, 0)],
               "joint": [(3, 4), (2, 3), (1, 2), (0, 1), (0, 0)],
               "theta first": [(3, 4), (1, 4), (0, 4), (0, 3), (0, 2), (0, 1), (0, 0)]}
        bad = 0
        for seed in range(int(sys.argv[2])):
            grid, occ, pose, scan = worlds.small_world(seed, size_cells=120 + 10 * (seed % 5))
            rng = np.random.RandomState(seed)
            full = seed % 3 == 0
            init = np.array(pose) + rng.uniform(-1, 1, 3) * [0.5, 0.5, 0.2]
            r = study(grid, scan, init, full, 1.5, 0.4, 5, 0.2, sch, verbose=False)
            if r is None:
                continue
            ok = all(v[1] for v in r.values())
            bad += not ok
            print("seed %3d %s %s" % (seed, "full " if full else "local",
                                      {k: v[0] for k, v in r.items()}), "OK" if ok else "MISMATCH")
        print("mismatches:", bad)
        return
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    grid, scans = bench.make_world(seed, 1)
    schedules = {
        "reference order": [(0, 6), (0, 5), (0, 4), (0, 3), (0, 2), (0, 1), (0, 0)],
        "joint, g-1/h-1": [(5, 6), (4, 5), (3, 4), (2, 3), (1, 2), (0, 1), (0, 0)],
        "theta first at the top": [(5, 6), (2, 6), (0, 6), (0, 5), (0, 4), (0, 3), (0, 2), (0, 1), (0, 0)],
    }
    study(grid, scans[0], (0, 0, 0), True, bench.LIN, bench.ANG, bench.DEPTH, bench.MIN_SCORE,
          schedules)


if __name__ == "__main__":
    main()
