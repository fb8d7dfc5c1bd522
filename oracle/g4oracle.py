"""ctypes binding of the CPU oracle (oracle/libg4oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module; the product package gridfour_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libg4oracle.so")

CODEC_HUFFMAN, CODEC_DEFLATE, CODEC_FLOAT, CODEC_CANON_HUFFMAN, CODEC_LSOP12 = 0, 1, 2, 3, 4
PRED_DIFFERENCING, PRED_LINEAR, PRED_TRIANGLE, PRED_DIFF_NULLS = 1, 2, 3, 4
INT4_NULL_CODE = -(2**31)
TERRAIN_SEED = 0x9E3779B97F4A7C15


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    srcs.append(os.path.join(_HERE, "..", "include", "g4terrain.h"))
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libg4oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.g4o_huffman_encode.restype = C.c_long
        L.g4o_canon_encode.restype = C.c_long
        L.g4o_codec_encode_i32.restype = C.c_long
        L.g4o_lsop12_encode_opts.restype = C.c_long
        L.g4o_lsop08_encode.restype = C.c_long
        L.g4o_codec_encode_f32.restype = C.c_long
        L.g4o_master_encode_i32.restype = C.c_long
        L.g4o_master_encode_f32.restype = C.c_long
        L.g4o_encode_grid.restype = C.c_long
        L.g4o_decode_grid.restype = C.c_long
        L.g4o_crc32c.restype = C.c_uint32
        L.g4o_java_round.argtypes = [C.c_float]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _u8(b):
    return np.frombuffer(bytes(b), dtype=np.uint8) if not isinstance(b, np.ndarray) else np.ascontiguousarray(b, dtype=np.uint8)


def m32_encode(values):
    v = _i32(values).ravel()
    out = np.zeros(6 * max(1, v.size), np.uint8)
    n = lib().g4o_m32_encode(_p(v), C.c_int(v.size), _p(out))
    return out[:n].tobytes()


def m32_decode(data, maxn=None):
    b = _u8(data)
    maxn = b.size if maxn is None else maxn
    out = np.zeros(max(1, maxn), np.int32)
    n = lib().g4o_m32_decode(_p(b), C.c_int(b.size), _p(out), C.c_int(maxn))
    if n < 0:
        raise ValueError("malformed M32 stream")
    return out[:n].copy()


def predictor_encode(model, tile):
    t = _i32(tile)
    nr, nc = t.shape
    out = np.zeros(6 * t.size, np.uint8)
    seed = C.c_int32(0)
    n = lib().g4o_predictor_encode(model, nr, nc, _p(t), _p(out), C.byref(seed))
    return n, seed.value, out[: max(n, 0)].tobytes()


def predictor_encode_int(model, tile):
    t = _i32(tile)
    nr, nc = t.shape
    out = np.zeros(t.size, np.int32)
    seed = C.c_int32(0)
    n = lib().g4o_predictor_encode_int(model, nr, nc, _p(t), _p(out), C.byref(seed))
    return n, seed.value, out[: max(n, 0)].copy()


def predictor_decode(model, seed, nr, nc, data):
    b = _u8(data)
    out = np.zeros((nr, nc), np.int32)
    rc = lib().g4o_predictor_decode(model, C.c_int32(seed), nr, nc, _p(b), C.c_int(b.size), _p(out))
    if rc:
        raise ValueError("predictor decode failed")
    return out


def predictor_decode_int(model, seed, nr, nc, residuals):
    r = _i32(residuals).ravel()
    out = np.zeros((nr, nc), np.int32)
    rc = lib().g4o_predictor_decode_int(model, C.c_int32(seed), nr, nc, _p(r), C.c_int(r.size), _p(out))
    if rc:
        raise ValueError("predictor decode failed")
    return out


def huffman_encode(symbols):
    s = _u8(symbols)
    out = np.zeros(s.size * 32 + 1024, np.uint8)
    nbits = lib().g4o_huffman_encode(_p(s), C.c_int(s.size), _p(out), C.c_long(out.size))
    if nbits < 0:
        raise ValueError("huffman encode failed")
    return out[: (nbits + 7) // 8].tobytes(), nbits


def huffman_decode(data, nsym):
    b = _u8(data)
    out = np.zeros(max(1, nsym), np.uint8)
    pos = C.c_long(0)
    rc = lib().g4o_huffman_decode(_p(b), C.c_long(b.size), C.c_int(nsym), _p(out), C.byref(pos))
    if rc:
        raise ValueError("huffman decode failed")
    return out[:nsym].tobytes(), pos.value


def huffman_code_lengths(symbols):
    s = _u8(symbols)
    out = np.zeros(256, np.int32)
    n = lib().g4o_huffman_code_lengths(_p(s), C.c_int(s.size), _p(out))
    return n, out


def canon_encode(text):
    t = _i32(text).ravel()
    out = np.zeros(t.size * 8 + 4096, np.uint8)
    nbits = lib().g4o_canon_encode(_p(t), C.c_int(t.size), _p(out), C.c_long(out.size))
    if nbits < 0:
        raise ValueError("canonical encode failed (%d)" % nbits)
    return out[: (nbits + 7) // 8].tobytes(), nbits


def canon_decode(data, nsym):
    b = _u8(data)
    out = np.zeros(max(1, nsym), np.int32)
    pos = C.c_long(0)
    rc = lib().g4o_canon_decode(_p(b), C.c_long(b.size), C.c_int(nsym), _p(out), C.byref(pos))
    if rc:
        raise ValueError("canonical decode failed")
    return out[:nsym].copy(), pos.value


def canon_tree_lengths(counts):
    c = np.ascontiguousarray(counts, dtype=np.int32)
    out = np.zeros(c.size, np.int32)
    rc = lib().g4o_canon_tree_lengths(_p(c), C.c_int(c.size), _p(out))
    if rc < 0:
        raise ValueError("tree build failed")
    return out, bool(rc)


def length_encode(code_lengths):
    c = np.ascontiguousarray(code_lengths, dtype=np.int32)
    codes = np.zeros(c.size, np.int32)
    runs = np.zeros(c.size, np.int32)
    n = lib().g4o_length_encode(C.c_int(c.size), _p(c), _p(codes), _p(runs))
    return codes[:n].copy(), runs[:n].copy()


def lsop12_coefficients(tile):
    t = _i32(tile)
    nr, nc = t.shape
    ud = np.zeros(12, np.float64)
    ok = lib().g4o_lsop12_coefficients(nr, nc, _p(t), _p(ud))
    return ud if ok else None


def lsop08_encode(codec_index, tile):
    """LsEncoder08.encode (legacy 8-coefficient codec); None where the reference returns null / throws on a singular matrix."""
    t = _i32(tile)
    nr, nc = t.shape
    cap = t.size * 8 + 4096
    out = np.zeros(cap, np.uint8)
    n = lib().g4o_lsop08_encode(codec_index, nr, nc, _p(t), _p(out), C.c_long(cap))
    if n == -1:
        return None
    if n < 0:
        raise ValueError("oracle LSOP08 encode failed (%d)" % n)
    return out[:n].tobytes()


def lsop08_decode(nr, nc, packing):
    b = _u8(packing)
    out = np.zeros((nr, nc), np.int32)
    if lib().g4o_lsop08_decode(nr, nc, _p(b), C.c_long(b.size), _p(out)):
        raise IOError("oracle LSOP08 decode failed")
    return out


def lsop12_residual_streams(tile):
    """(seed, 12 float32 coefficients, initializer M32 bytes, interior M32 bytes) of LsOptimalPredictor12.encode, or None."""
    t = _i32(tile)
    nr, nc = t.shape
    seed = C.c_int32(0)
    u = np.zeros(12, np.float32)
    a, b = np.zeros(t.size * 6 + 64, np.uint8), np.zeros(t.size * 6 + 64, np.uint8)
    na, nb = C.c_long(0), C.c_long(0)
    ok = lib().g4o_lsop12_residual_streams(nr, nc, _p(t), C.byref(seed), _p(u), _p(a), C.byref(na), _p(b), C.byref(nb))
    if ok != 1:
        return None
    return seed.value, u, a[: na.value].tobytes(), b[: nb.value].tobytes()


def huffman_decode_at(data, nsym, bitpos):
    """Legacy Huffman decode of nsym symbols starting at bit `bitpos`; returns (symbols, bit position after the stream)."""
    b = _u8(data)
    out = np.zeros(max(1, nsym), np.uint8)
    pos = C.c_long(bitpos)
    rc = lib().g4o_huffman_decode_at(_p(b), C.c_long(b.size), C.c_int(nsym), _p(out), C.byref(pos))
    if rc:
        raise ValueError("huffman decode failed")
    return out[:nsym].tobytes(), pos.value


def java_round(x):
    return lib().g4o_java_round(C.c_float(x))


def crc32c(data):
    b = _u8(data)
    return lib().g4o_crc32c(_p(b), C.c_long(b.size))


def codec_encode_i32(codec, codec_index, tile, cap=None):
    """Returns (packing bytes or None when the codec declines, predictor code)."""
    t = _i32(tile)
    nr, nc = t.shape
    cap = cap or (t.size * 8 + 4096)
    out = np.zeros(cap, np.uint8)
    pred = C.c_int(0)
    n = lib().g4o_codec_encode_i32(codec, codec_index, nr, nc, _p(t), _p(out), C.c_long(cap), C.byref(pred))
    if n == -1:
        return None, 0
    if n < 0:
        raise ValueError("oracle encode failed (%d)" % n)
    return out[:n].tobytes(), pred.value


def lsop12_encode(codec_index, tile, deflate=True, checksum=False):
    t = _i32(tile)
    nr, nc = t.shape
    cap = t.size * 8 + 4096
    out = np.zeros(cap, np.uint8)
    n = lib().g4o_lsop12_encode_opts(codec_index, nr, nc, _p(t), _p(out), C.c_long(cap), int(deflate), int(checksum))
    if n == -1:
        return None
    if n < 0:
        raise ValueError("oracle encode failed (%d)" % n)
    return out[:n].tobytes()


def codec_decode_i32(codec, nr, nc, packing):
    b = _u8(packing)
    out = np.zeros((nr, nc), np.int32)
    rc = lib().g4o_codec_decode_i32(codec, nr, nc, _p(b), C.c_long(b.size), _p(out))
    if rc == 1:
        return None
    if rc:
        raise IOError("oracle decode failed")
    return out


def codec_encode_f32(codec_index, tile):
    t = np.ascontiguousarray(tile, dtype=np.float32)
    nr, nc = t.shape
    cap = t.size * 8 + 4096
    out = np.zeros(cap, np.uint8)
    n = lib().g4o_codec_encode_f32(codec_index, nr, nc, _p(t), _p(out), C.c_long(cap))
    if n < 0:
        raise ValueError("oracle float encode failed")
    return out[:n].tobytes()


def codec_decode_f32(nr, nc, packing):
    b = _u8(packing)
    out = np.zeros((nr, nc), np.float32)
    rc = lib().g4o_codec_decode_f32(nr, nc, _p(b), C.c_long(b.size), _p(out))
    if rc:
        raise IOError("oracle float decode failed")
    return out


def master_encode_i32(codec_ids, tile):
    t = _i32(tile)
    nr, nc = t.shape
    ids = np.ascontiguousarray(codec_ids, dtype=np.int32)
    cap = t.size * 4 + 64
    out = np.zeros(cap, np.uint8)
    n = lib().g4o_master_encode_i32(_p(ids), C.c_int(ids.size), nr, nc, _p(t), _p(out), C.c_long(cap))
    if n < 0:
        raise ValueError("oracle master encode failed")
    return out[:n].tobytes()


def master_decode_i32(codec_ids, nr, nc, payload):
    b = _u8(payload)
    ids = np.ascontiguousarray(codec_ids, dtype=np.int32)
    out = np.zeros((nr, nc), np.int32)
    rc = lib().g4o_master_decode_i32(_p(ids), C.c_int(ids.size), nr, nc, _p(b), C.c_long(b.size), _p(out))
    if rc:
        raise IOError("oracle master decode failed")
    return out


def master_encode_f32(codec_ids, tile):
    t = np.ascontiguousarray(tile, dtype=np.float32)
    nr, nc = t.shape
    ids = np.ascontiguousarray(codec_ids, dtype=np.int32)
    cap = t.size * 4 + 64
    out = np.zeros(cap, np.uint8)
    n = lib().g4o_master_encode_f32(_p(ids), C.c_int(ids.size), nr, nc, _p(t), _p(out), C.c_long(cap))
    if n < 0:
        raise ValueError("oracle master encode failed")
    return out[:n].tobytes()


def master_decode_f32(codec_ids, nr, nc, payload):
    b = _u8(payload)
    ids = np.ascontiguousarray(codec_ids, dtype=np.int32)
    out = np.zeros((nr, nc), np.float32)
    rc = lib().g4o_master_decode_f32(_p(ids), C.c_int(ids.size), nr, nc, _p(b), C.c_long(b.size), _p(out))
    if rc:
        raise IOError("oracle master decode failed")
    return out


def encode_grid(codec_ids, grid, tile_rows, tile_cols, n_threads=1):
    """CodecMaster rule over every tile of `grid` (int32 or float32).  Returns (arena, slot_bytes, lens)."""
    g = np.ascontiguousarray(grid)
    is_float = g.dtype == np.float32
    assert g.dtype in (np.int32, np.float32)
    ids = np.ascontiguousarray(codec_ids, dtype=np.int32)
    n_tiles = (g.shape[0] // tile_rows) * (g.shape[1] // tile_cols)
    slot = tile_rows * tile_cols * 4
    arena = np.zeros(n_tiles * slot, np.uint8)
    lens = np.zeros(n_tiles, np.uint32)
    rc = lib().g4o_encode_grid(_p(ids), C.c_int(ids.size), int(is_float), _p(g), C.c_long(g.shape[0]), C.c_long(g.shape[1]),
                               tile_rows, tile_cols, n_threads, _p(arena), C.c_long(slot), _p(lens))
    if rc:
        raise ValueError("oracle grid encode failed at tile %d" % (-rc - 1))
    return arena, slot, lens


def decode_grid(codec_ids, arena, offsets, lens, grid_rows, grid_cols, tile_rows, tile_cols, dtype=np.int32, n_threads=1):
    ids = np.ascontiguousarray(codec_ids, dtype=np.int32)
    a = np.ascontiguousarray(arena, dtype=np.uint8)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    ln = np.ascontiguousarray(lens, dtype=np.uint32)
    out = np.zeros((grid_rows, grid_cols), dtype)
    rc = lib().g4o_decode_grid(_p(ids), C.c_int(ids.size), int(dtype == np.float32), _p(a), _p(off), _p(ln),
                               C.c_long(grid_rows), C.c_long(grid_cols), tile_rows, tile_cols, n_threads, _p(out))
    if rc:
        raise IOError("oracle grid decode failed at tile %d" % (-rc - 1))
    return out


def hardware_threads():
    return lib().g4o_hardware_threads()


def terrain_i32(row0, col0, nr, nc, seed=TERRAIN_SEED, n_threads=1):
    out = np.zeros((nr, nc), np.int32)
    lib().g4o_terrain_i32(C.c_uint64(seed), C.c_long(row0), C.c_long(col0), C.c_long(nr), C.c_long(nc), _p(out), n_threads)
    return out


def terrain_f32(row0, col0, nr, nc, seed=TERRAIN_SEED, n_threads=1):
    out = np.zeros((nr, nc), np.float32)
    lib().g4o_terrain_f32(C.c_uint64(seed), C.c_long(row0), C.c_long(col0), C.c_long(nr), C.c_long(nc), _p(out), n_threads)
    return out


# ---- ICompressionDecoder.analyze (test infrastructure) ------------------------------------------------------------------
def analyze_m32(codec, nr, nc, packing):
    """What CodecHuffman.analyze (compress/CodecHuffman.java:172-199) or CodecDeflate.analyze (compress/CodecDeflate.java:71-106)
    hands to CodecStats (compress/CodecStats.java:84-131) for one tile, restated with numpy:
    (predictor, nBytes, nSymbols, nBitsOverhead, nM32, observed, entropy, successor-pair counts[65536])."""
    import math
    import zlib

    p = bytes(packing)
    n_m32 = int.from_bytes(p[6:10], "little")
    overhead = 0
    if codec == CODEC_HUFFMAN:
        m32, _pos = huffman_decode_at(p[10:], n_m32, 0)
        bits = np.unpackbits(np.frombuffer(p[10:], np.uint8), bitorder="little")
        n_leaf = int(p[10]) + 1
        if n_leaf == 1:
            overhead = 8 + 1 + 8                      # HuffmanDecoder.decodeTree :67-77
        else:
            pos, leaves = 8, 0
            while leaves < n_leaf:                    # pre-order: branch '0', leaf '1' + 8-bit symbol (:80-159)
                if bits[pos]:
                    pos += 9
                    leaves += 1
                else:
                    pos += 1
            overhead = pos
    elif codec == CODEC_DEFLATE:
        m32 = zlib.decompress(p[10:])
        assert len(m32) == n_m32
    else:
        raise ValueError("analyze_m32: CodecHuffman or CodecDeflate")
    b = np.frombuffer(m32, np.uint8)
    counts = np.bincount(b, minlength=256)
    d = float(n_m32)
    s = 0.0
    for i in range(256):                              # CodecStats.addCountsForM32 :118-125, same order
        if counts[i] > 0:
            pr = counts[i] / d
            s += pr * math.log(pr) / math.log(2.0)
    pairs = np.zeros(65536, np.uint64)
    if n_m32 >= 2:
        np.add.at(pairs, (b[:-1].astype(np.int64) << 8) | b[1:].astype(np.int64), 1)
    return p[1], len(p) - 10, nr * nc, overhead, n_m32, int((counts > 0).sum()), -s, pairs
