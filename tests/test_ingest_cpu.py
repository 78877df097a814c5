"""Grid ingest from Cartographer's serialized forms: the hand-written
protobuf / pbstream decoding in libcsm_b200.so against messages encoded here byte by byte
from the reference's .proto definitions (mapping/proto/grid_2d.proto:23-42, map_limits.proto,
cell_limits_2d.proto, serialization.proto, submap.proto; framing io/proto_stream.cc:27-110).
Host-only entry points: no GPU needed."""
import ctypes as C
import gzip
import struct

import numpy as np
import pytest

from tests import protowire as pw


@pytest.fixture(scope="module")
def lib():
    from cartographer_b200 import _lib
    return _lib.lib()


class Info(C.Structure):
    _fields_ = [("num_x_cells", C.c_int32), ("num_y_cells", C.c_int32),
                ("resolution", C.c_double), ("max_x", C.c_double), ("max_y", C.c_double),
                ("min_correspondence_cost", C.c_float), ("max_correspondence_cost", C.c_float),
                ("is_tsdf", C.c_int32), ("reserved", C.c_int32)]


def _decode(lib, blob, want_cells):
    info = Info()
    buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
    cells = np.zeros(max(1, want_cells), np.uint16)
    st = lib.csm_grid2d_proto_decode(buf, C.c_int64(len(blob)), C.byref(info),
                                     cells.ctypes.data_as(C.POINTER(C.c_uint16)),
                                     C.c_int64(len(cells)))
    return st, info, cells


def test_grid2d_proto_decode(lib):
    rng = np.random.RandomState(0)
    cells = rng.randint(0, 65536, size=(37, 53)).astype(np.uint16)
    blob = pw.grid2d(cells, 0.05, 12.5, -3.25, np.float32(0.1), np.float32(0.9))
    st, info, got = _decode(lib, blob, cells.size)
    assert st == 0
    assert (info.num_x_cells, info.num_y_cells) == (53, 37)
    assert (info.resolution, info.max_x, info.max_y) == (0.05, 12.5, -3.25)
    assert np.float32(info.min_correspondence_cost) == np.float32(0.1)
    assert np.float32(info.max_correspondence_cost) == np.float32(0.9)
    assert info.is_tsdf == 0
    np.testing.assert_array_equal(got.reshape(37, 53), cells)
    # un-packed repeated field encoding is legal protobuf too
    blob2 = pw.grid2d(cells[:2, :3], 0.05, 1.0, 2.0, np.float32(0.1), np.float32(0.9), packed=False)
    st, info, got = _decode(lib, blob2, 6)
    assert st == 0
    np.testing.assert_array_equal(got.reshape(2, 3), cells[:2, :3])


def test_grid2d_proto_legacy_bounds_and_errors(lib):
    cells = np.arange(6, dtype=np.uint16).reshape(2, 3)
    # legacy protos without the two cost fields get the default bounds (grid_2d.cc:24-44)
    st, info, _ = _decode(lib, pw.grid2d(cells, 0.05, 1.0, 2.0, None, None), 6)
    assert st == 0
    k_min_p = np.float32(0.1)
    k_max_p = np.float32(1.0) - k_min_p
    assert np.float32(info.min_correspondence_cost) == np.float32(1.0) - k_max_p
    assert np.float32(info.max_correspondence_cost) == np.float32(1.0) - k_min_p
    # cell count must match the limits; truncated messages are rejected with a status code
    bad = pw.grid2d(cells, 0.05, 1.0, 2.0, np.float32(0.1), np.float32(0.9), lie_about_cells=True)
    assert _decode(lib, bad, 6)[0] == 1
    good = pw.grid2d(cells, 0.05, 1.0, 2.0, np.float32(0.1), np.float32(0.9))
    assert _decode(lib, good[:-3], 6)[0] == 1
    assert b"Grid2D" in lib.csm_last_error_string() or b"cell" in lib.csm_last_error_string()


def test_pbstream_walk_counts_submaps(lib, tmp_path):
    rng = np.random.RandomState(1)
    grids = [rng.randint(0, 32768, size=(20 + k, 30)).astype(np.uint16) for k in range(3)]
    path = tmp_path / "map.pbstream"
    pw.write_pbstream(str(path), [(0, k, pw.grid2d(g, 0.05, 1.0, 2.0, np.float32(0.1),
                                                   np.float32(0.9)))
                                  for k, g in enumerate(grids)])
    n = C.c_int32(-1)
    st = lib.csm_pbstream_load_stacks2d(str(path).encode(), 7, 0, 0, None, None, C.byref(n))
    assert st == 0 and n.value == 3
    # not a proto stream
    other = tmp_path / "junk.bin"
    other.write_bytes(b"\x00" * 64)
    assert lib.csm_pbstream_load_stacks2d(str(other).encode(), 7, 0, 0, None, None,
                                          C.byref(n)) == 1
