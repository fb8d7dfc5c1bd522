"""-m gpu: encode side of CodecDeflate, CodecFloat and the LSOP12 Deflate alternative.

The GPU encoder replays zlib's deflate_slow (g4_deflate_enc.cuh) and the oracle calls the system zlib at the same
level, so the packings are expected to be byte-identical; north_star only demands that they decode through the
reference and keep the codec choice, which is checked independently (oracle decode + stock zlib inflate)."""
import zlib

import numpy as np
import pytest

from gpu_common import first_diff, parity_grids

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g4():
    import gridfour_b200

    return gridfour_b200


def test_deflate_encode_matches_oracle(g4, oracle):
    enc = g4.CodecDeflate()
    for name, grid in parity_grids(oracle).items():
        want, info = oracle.codec_encode_i32(oracle.CODEC_DEFLATE, 1, grid)
        got = enc.encode(1, grid.shape[0], grid.shape[1], grid)
        assert got is not None, name
        # independent checks: stock zlib inflates the stream; the oracle's decoder reproduces the tile
        m32 = zlib.decompress(got[10:])
        assert len(m32) == int.from_bytes(got[6:10], "little"), name
        back = oracle.codec_decode_i32(oracle.CODEC_DEFLATE, grid.shape[0], grid.shape[1], got)
        assert np.array_equal(back, grid), "%s: %s" % (name, first_diff(back, grid))
        assert got[1] == want[1], "%s: predictor %d vs oracle %d" % (name, got[1], want[1])
        assert got == want, "%s: %s" % (name, first_diff(got, want))


def test_deflate_all_null_tile_declines(g4):
    nulls = np.full((8, 8), -(2 ** 31), np.int32)
    assert g4.CodecDeflate().encode(0, 8, 8, nulls) is None  # CodecDeflate.java:167-169


def test_float_encode_matches_oracle(g4, oracle):
    rng = np.random.default_rng(4)
    enc = g4.CodecFloat()
    tiles = [oracle.terrain_f32(0, 0, 30, 40), oracle.terrain_f32(100, 200, 120, 120),
             rng.integers(0, 2 ** 32, (16, 16), dtype=np.uint64).astype(np.uint32).view(np.float32),
             np.zeros((3, 5), np.float32), oracle.terrain_f32(9, 9, 37, 41),
             np.array([[np.nan, -0.0, np.inf], [1e-40, -np.inf, 3.5]], np.float32)]
    for t in tiles:
        want = oracle.codec_encode_f32(2, t)
        got = enc.encodeFloats(2, t.shape[0], t.shape[1], t)
        assert got is not None
        back = oracle.codec_decode_f32(t.shape[0], t.shape[1], got)
        assert np.array_equal(back.view(np.uint32), t.view(np.uint32)), first_diff(back.view(np.uint32), t.view(np.uint32))
        off = 2
        for plane in range(5):  # every plane is a stock zlib stream of the right size
            n = int.from_bytes(got[off:off + 4], "little")
            raw = zlib.decompress(got[off + 4:off + 4 + n])
            assert len(raw) == ((t.size + 7) // 8 if plane == 0 else t.size)
            off += 4 + n
        assert off == len(got)
        assert got == want, first_diff(got, want)
        out = g4.CodecFloat().decodeFloats(t.shape[0], t.shape[1], got)
        assert np.array_equal(out.view(np.uint32), t.view(np.uint32))


def test_float_level9_parser_long_planes_match_oracle(g4, oracle):
    """CodecFloat's Deflater(9) streams go through the warp-per-stream parser (deflate_lazy_kernel: chains of up to 4096
    candidates searched 32 at a time, only where zlib searches).  Planes of 43,200 bytes and more: chains that reach the
    level's limit, matches of 258 next to literal stretches, candidates beyond MAX_DIST (32,506) and an all-equal plane;
    a batch so that several warps of a CTA run different streams.  Byte-identical to the oracle's zlib."""
    rng = np.random.default_rng(11)
    smooth = oracle.terrain_f32(40, 60, 180, 240)
    steps = np.floor(oracle.terrain_f32(0, 0, 180, 240) / 7.0).astype(np.float32)  # plateaus: long runs in every plane
    few = rng.choice(np.array([1.5, -2.25, 1024.0, 3.0e-3], np.float32), (200, 300))  # 60,000 cells, four byte patterns: deep chains
    flat = np.full((180, 240), 12.5, np.float32)
    enc = g4.CodecFloat()
    for name, t in (("smooth", smooth), ("steps", steps), ("few", few), ("flat", flat)):
        want = oracle.codec_encode_f32(0, t)
        got = enc.encodeFloats(0, t.shape[0], t.shape[1], t)
        assert got is not None and got == want, "%s: %s" % (name, first_diff(got, want))
        out = g4.CodecFloat().decodeFloats(t.shape[0], t.shape[1], got)
        assert np.array_equal(out.view(np.uint32), t.view(np.uint32)), name
    grid = oracle.terrain_f32(0, 0, 2 * 180, 3 * 240)
    spec = g4.CodecSpecification(default=False)
    spec.addCompressionCodec("GvrsFloat", g4.CodecFloat, g4.CodecFloat)
    master = g4.CodecMaster(spec)
    batch = master.encodeTiles(grid, 180, 240)
    for t in range(6):
        tr, tc = divmod(t, 3)
        tile = np.ascontiguousarray(grid[tr * 180:(tr + 1) * 180, tc * 240:(tc + 1) * 240])
        want = oracle.codec_encode_f32(0, tile)
        assert batch.payload(t) == want, "tile %d: %s" % (t, first_diff(batch.payload(t), want))
    assert np.array_equal(master.decodeTiles(batch).view(np.uint32), grid.view(np.uint32))


def test_lsop_deflate_alternative_matches_oracle(g4, oracle):
    """Repetitive tiles make LsEncoder12 prefer the two zlib streams (type 1); terrain keeps canonical Huffman (type 2)."""
    r, c = np.mgrid[0:64, 0:64]
    grids = [((r % (3 + k)) * 1000 + (c % (5 + k)) * 37 + (r // 16) * 5 + (r * c) % (2 + k)).astype(np.int32) for k in range(6)]
    grids += [oracle.terrain_i32(0, 0, 90, 120), oracle.terrain_i32(7000, 3000, 180, 240)]
    types = set()
    for g in grids:
        want = oracle.lsop12_encode(0, g)
        got = g4.LsEncoder12().encode(0, g.shape[0], g.shape[1], g)
        if want is None:
            assert got is None
            continue
        assert got is not None
        types.add(got[1] & 0x0F)
        back = oracle.codec_decode_i32(oracle.CODEC_LSOP12, g.shape[0], g.shape[1], got)
        assert np.array_equal(back, g), first_diff(back, g)
        assert got == want, first_diff(got, want)
        out = g4.LsDecoder12().decode(g.shape[0], g.shape[1], got)
        assert np.array_equal(out, g)
    assert types == {1, 2}


def test_batched_best_of_matches_codec_master(g4, oracle):
    """Config-1/3 codec lists through encodeTiles: per-tile bytes, codec choice and bits/sample equal the oracle's
    CodecMaster.encodeSingleThread (ties -> lowest index) + TileElementInt raw fallback."""
    grid = oracle.terrain_i32(0, 0, 2 * 90, 3 * 120)
    grid[0:90, 0:120] = np.random.default_rng(1).integers(-(2 ** 31), 2 ** 31, (90, 120), dtype=np.int64).astype(np.int32)  # raw tile
    for names, ids in ((("GvrsHuffman", "GvrsDeflate"), [0, 1]), (("GvrsHuffman", "GvrsDeflate", "LSOP12"), [0, 1, 4])):
        spec = g4.CodecSpecification(default=False)
        std = {"GvrsHuffman": (g4.CodecHuffman,), "GvrsDeflate": (g4.CodecDeflate,), "LSOP12": (g4.LsEncoder12, g4.LsDecoder12)}
        for nme in names:
            spec.addCompressionCodec(nme, *std[nme])
        master = g4.CodecMaster(spec)
        batch = master.encodeTiles(grid, 90, 120)
        for t in range(6):
            tr, tc = divmod(t, 3)
            tile = grid[tr * 90:(tr + 1) * 90, tc * 120:(tc + 1) * 120]
            want = oracle.master_encode_i32(ids, tile)
            assert batch.payload(t) == want, "tile %d codecs %s: %s" % (t, names, first_diff(batch.payload(t), want))
        assert np.array_equal(master.decodeTiles(batch), grid)


def test_deflate_stream_length_edges_match_oracle(g4, oracle):
    """Streams on both sides of the staged encoder's 65,535-byte limit (sort / match / decide / emit kernels below it, the
    one-thread-per-stream replay above), several DEFLATE blocks per stream (> 16,383 symbols), long runs (every position
    in one hash bucket, matches of 258) and incompressible data (stored blocks): byte-identical to the oracle's zlib."""
    rng = np.random.default_rng(5)
    enc = g4.CodecDeflate()
    tiles = {}
    # 150x150 = 22,500 cells: one-byte residuals (22.5 KB, two blocks), two-byte (45 KB), three-byte (67.5 KB > limit)
    tiles["1byte"] = rng.integers(-60, 60, (150, 150)).cumsum(axis=1).astype(np.int32)
    tiles["2byte"] = (rng.integers(-15000, 15000, (150, 150))).astype(np.int32)
    tiles["3byte"] = (rng.integers(-2000000, 2000000, (150, 150))).astype(np.int32)
    # 104x105 = 10,920 cells of 6-byte codes -> 65,520 bytes for one predictor, a few more or fewer for the others
    tiles["near_limit"] = np.where(rng.random((104, 105)) < 0.5, 2 ** 31 - 7, -(2 ** 31) + 9).astype(np.int32)
    tiles["zeros"] = np.zeros((200, 200), np.int32)
    tiles["period7"] = (np.arange(180 * 240) % 7).reshape(180, 240).astype(np.int32)
    tiles["noise"] = rng.integers(-(2 ** 31), 2 ** 31, (90, 120), dtype=np.int64).astype(np.int32)
    for name, grid in tiles.items():
        got = enc.encode(1, grid.shape[0], grid.shape[1], grid)
        want, _pred = oracle.codec_encode_i32(oracle.CODEC_DEFLATE, 1, grid)
        if want is None:
            assert got is None, name
            continue
        assert got is not None and bytes(got) == bytes(want), "%s: %s" % (name, first_diff(got, want))
        out = g4.CodecDeflate().decode(grid.shape[0], grid.shape[1], got)
        assert np.array_equal(out, grid), name
