"""Small fixture builders for the tests (CPU, numpy)."""
import numpy as np

from benchmarks import synthetic


def insert_range_data(oracle, nx, ny, res, max_x, max_y, origin_xy, returns_xyz, grow=True,
                      hit_probability=0.7, miss_probability=0.4):
    """Stand-in for ProbabilityGridRangeDataInserter2D::Insert on an empty grid
    (mapping/2d/probability_grid_range_data_inserter_2d.cc:52-133 with
    hit 0.7 / miss 0.4 as in the reference tests): every hit cell gets the hit
    probability, every still-unknown cell crossed by a ray gets the miss
    probability (hits have priority, one update per cell).  The ray traversal is
    a plain fine-stepped walk, not the reference's subpixel mask — the fixture
    only has to be a plausible map; map *writing* is out of scope.
    """
    del grow
    cells = np.zeros((ny, nx), np.uint16)
    hit_v = oracle.correspondence_cost_to_value(1.0 - hit_probability)
    miss_v = oracle.correspondence_cost_to_value(1.0 - miss_probability)
    grid = oracle.Grid2D(cells, res, max_x, max_y)
    pts = np.asarray(returns_xyz, np.float32)
    hits = [oracle.get_cell_index(res, max_x, max_y, float(p[0]), float(p[1])) for p in pts]
    for (cx, cy) in hits:
        if 0 <= cx < nx and 0 <= cy < ny:
            cells[cy, cx] = hit_v
    o = np.asarray(origin_xy, np.float64)
    for p, (hx, hy) in zip(pts, hits):
        d = np.asarray(p[:2], np.float64) - o
        n = max(2, int(np.linalg.norm(d) / (res * 0.1)))
        for t in np.linspace(0.0, 1.0, n, endpoint=False):
            q = o + t * d
            cx, cy = oracle.get_cell_index(res, max_x, max_y, float(q[0]), float(q[1]))
            if (cx, cy) != (hx, hy) and 0 <= cx < nx and 0 <= cy < ny and cells[cy, cx] == 0:
                cells[cy, cx] = miss_v
    grid.cells = np.ascontiguousarray(cells)
    return grid


def small_world(seed, size_cells=200, beams=181, max_range=8.0):
    """A small floor plan + one scan from a random free pose (for fast parity cases)."""
    rng = np.random.RandomState(seed)
    grid, occ = synthetic.make_grid2d(seed, size_cells=size_cells)
    pose = synthetic.random_free_pose(occ, grid, rng, margin_cells=10)
    scan = synthetic.cast_scan(occ, grid, pose, beams=beams, max_range=max_range, seed=seed)
    return grid, occ, pose, scan
