"""Executable specification of the two device layouts of a precomputation level
(cartographer_b200/csrc/engine2d.cuh: StackDev::dec4 and StackDev::win), in numpy.
The CUDA kernels index these arrays with exactly the formulas below; the GPU parity
tests prove the kernels, these tests prove the formulas (every candidate / child value
they deliver equals PrecomputationGrid2D::GetValue, fast_correlative_scan_matcher_2d.h:
56-71: 0 outside the wide grid).  CPU only."""
import numpy as np
import pytest


def _get_value(lvl, x, y):
    wy, wx = lvl.shape
    return int(lvl[y, x]) if 0 <= x < wx and 0 <= y < wy else 0


def _build_dec4(lvl, h):
    """D[((ay*s+ax)*jd + J)*ids + I] = lvl[s*J+ay, s*I+ax]; four copies, copy k shifted
    left by k bytes, index 0 at byte 16 of every copy (k_stack_decimate4)."""
    wy, wx = lvl.shape
    s = 1 << h
    id_, jd = (wx + s - 1) // s, (wy + s - 1) // s
    ids = (id_ + 3 + 3) // 4 * 4          # >= 3 zero bytes after every row
    total = s * s * jd * ids
    D = np.zeros(total, np.int64)
    for ay in range(s):
        for ax in range(s):
            for J in range(jd):
                for I in range(id_):
                    x, y = s * I + ax, s * J + ay
                    if x < wx and y < wy:
                        D[((ay * s + ax) * jd + J) * ids + I] = lvl[y, x]
    lpad = (total + 32 + 15) // 16 * 16
    dec4 = np.zeros(4 * lpad, np.int64)
    for k in range(4):
        for u in range(lpad):
            t = u - 16 + k
            if 0 <= t < total:
                dec4[k * lpad + u] = D[t]
    return dec4, lpad, id_, jd, ids


@pytest.mark.parametrize("h,wx,wy", [(2, 19, 17), (2, 20, 20), (3, 41, 37), (1, 9, 9), (3, 67, 70)])
def test_tile_pass_indexing(h, wx, wy):
    """k_score_top_tile: lane f-1.. of the tile of one scan point, copy k = qx & 3,
    lattice cell = (row - qy, column - qx); summed over points it equals the
    candidate-by-candidate definition (ScoreCandidates, fast...2d.cc:314-333)."""
    rng = np.random.RandomState(h * 100 + wx)
    lvl = rng.randint(0, 256, (wy, wx))
    dec4, lpad, id_, jd, ids = _build_dec4(lvl, h)
    s, s1 = 1 << h, (1 << h) - 1
    nxc, nyc = int(rng.randint(1, 14)), int(rng.randint(1, 14))
    min_x, min_y = int(rng.randint(-40, 10)), int(rng.randint(-40, 10))
    pts = rng.randint(-3 * s, wx + 3 * s, (40, 2))
    want = np.zeros((nxc, nyc), np.int64)
    for i in range(nxc):
        for j in range(nyc):
            want[i, j] = sum(_get_value(lvl, c[0] + min_x + s1 + i * s, c[1] + min_y + s1 + j * s)
                             for c in pts)
    rw, qr = ids // 4, (nxc + 3) // 4
    W = jd * rw + 1
    lat = np.zeros((nyc, qr * 4), np.int64)
    for c in pts:
        bx, by = c[0] + min_x + s1, c[1] + min_y + s1
        qx, qy = bx >> h, by >> h
        off = ((((by & s1) << h) | (bx & s1)) * jd) * ids
        k = qx & 3
        for f in range(-1, W - 1):
            addr = 16 + off + k * lpad + 4 * f          # dec = dec4 + 16
            word = dec4[addr:addr + 4]
            r = (f + rw) // rw - 1
            c0 = ((f - r * rw) << 2) + k
            if c0 + 3 >= ids:                            # the word continues in the next row
                r, c0 = r + 1, c0 - ids
            j, i0 = r - qy, c0 - qx
            if c0 < id_ and 0 <= r < jd and 0 <= j < nyc and 0 <= i0 < qr * 4:
                assert i0 % 4 == 0
                lat[j, i0:i0 + 4] += word
    np.testing.assert_array_equal(lat[:, :nxc].T, want)


@pytest.mark.parametrize("h,nx,ny", [(1, 7, 9), (2, 10, 13), (3, 20, 17), (3, 8, 8), (4, 30, 40)])
def test_child_window_indexing(h, nx, ny):
    """StackDev::win: one word = the four children (level h-1) of a level-h node for one
    scan point (k_stack_window / ScoreChildren / k_expand_lattice)."""
    rng = np.random.RandomState(h * 10 + nx)
    S, s = 1 << h, 1 << (h - 1)
    S1 = S - 1
    wx, wy = nx + s - 1, ny + s - 1                      # wide limits of level h-1
    lvl = rng.randint(1, 256, (wy, wx))
    ids, jd = (wx + S - 1) // S + 1, (wy + S - 1) // S + 1
    win = np.zeros((S * S * jd * ids, 4), np.int64)
    for u in range(len(win)):
        I = u % ids - 1
        r = u // ids
        J = r % jd - 1
        r //= jd
        ax, ay = r % S, r // S
        x, y = S * I + ax, S * J + ay
        win[u] = [_get_value(lvl, x, y), _get_value(lvl, x + s, y),
                  _get_value(lvl, x, y + s), _get_value(lvl, x + s, y + s)]
    for _ in range(400):
        px = int(rng.randint(-3 * S, nx + 3 * S))       # scan point + node offset
        py = int(rng.randint(-3 * S, ny + 3 * S))
        lx, ly = px + s - 1, py + s - 1                  # wide index of child (0, 0)
        Qx, Qy = (lx >> h) + 1, (ly >> h) + 1
        word = np.zeros(4, np.int64)
        if 0 <= Qx < ids and 0 <= Qy < jd:
            word = win[((((ly & S1) << h) | (lx & S1)) * jd + Qy) * ids + Qx]
        want = [_get_value(lvl, lx, ly), _get_value(lvl, lx + s, ly),
                _get_value(lvl, lx, ly + s), _get_value(lvl, lx + s, ly + s)]
        assert list(word) == want
