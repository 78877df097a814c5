"""Minimal protobuf wire-format WRITER for the tests of the ingest path: encodes the few
Cartographer messages byte by byte from the field numbers in the reference's .proto files
(grid_2d.proto:23-42, map_limits.proto:22-26, cell_limits_2d.proto:19-22,
transform.proto Vector2d, hybrid_grid.proto:19-28, submap.proto:24-41,
serialization.proto Submap / SerializedData / SerializationHeader) and the pbstream framing
of io/proto_stream.cc:27-60."""
import gzip
import struct


def varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def key(field, wire):
    return varint((field << 3) | wire)


def f_varint(field, v):
    return key(field, 0) + varint(v)


def f_double(field, v):
    return key(field, 1) + struct.pack("<d", v)


def f_float(field, v):
    return key(field, 5) + struct.pack("<f", v)


def f_bytes(field, payload):
    return key(field, 2) + varint(len(payload)) + payload


def zigzag(v):
    return (v << 1) ^ (v >> 63)


def grid2d(cells, resolution, max_x, max_y, min_cost, max_cost, packed=True,
           lie_about_cells=False):
    ny, nx = cells.shape
    limits = (f_double(1, resolution) + f_bytes(2, f_double(1, max_x) + f_double(2, max_y)) +
              f_bytes(3, f_varint(1, nx + (1 if lie_about_cells else 0)) + f_varint(2, ny)))
    flat = [int(c) for c in cells.reshape(-1)]
    if packed:
        body = f_bytes(2, b"".join(varint(c) for c in flat))
    else:
        body = b"".join(f_varint(2, c) for c in flat)
    msg = f_bytes(1, limits) + body
    msg += f_bytes(3, f_varint(1, nx - 1) + f_varint(2, ny - 1))   # known_cells_box (ignored)
    msg += f_bytes(4, b"")                                          # probability_grid_2d {}
    if min_cost is not None:
        msg += f_float(6, float(min_cost)) + f_float(7, float(max_cost))
    return msg


def hybrid_grid(resolution, indices, values):
    xs = b"".join(varint(zigzag(int(i[0]))) for i in indices)
    ys = b"".join(varint(zigzag(int(i[1]))) for i in indices)
    zs = b"".join(varint(zigzag(int(i[2]))) for i in indices)
    vs = b"".join(varint(int(v)) for v in values)
    return f_float(1, resolution) + f_bytes(3, xs) + f_bytes(4, ys) + f_bytes(5, zs) + f_bytes(6, vs)


def write_pbstream(path, submaps):
    """submaps: [(trajectory_id, submap_index, serialized Grid2D)]"""
    def chunk(payload):
        z = gzip.compress(payload, compresslevel=1)
        return struct.pack("<Q", len(z)) + z
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", 0x7b1d1f7b5bf501db))
        f.write(chunk(f_varint(1, 2)))                     # SerializationHeader{format_version}
        f.write(chunk(f_bytes(1, b"")))                    # SerializedData{pose_graph {}}
        for traj, index, grid in submaps:
            submap2d = f_varint(2, 90) + f_varint(3, 1) + f_bytes(4, grid)
            submap = f_bytes(1, f_varint(1, traj) + f_varint(2, index)) + f_bytes(2, submap2d)
            f.write(chunk(f_bytes(3, submap)))             # SerializedData{submap = 3}
