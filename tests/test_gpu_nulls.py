"""-m gpu: tiles that contain INT4_NULL_CODE (PredictorModelDifferencingWithNulls) through
CodecHuffman, CodecDeflate and CodecCanonHuffman: packings byte-identical to the oracle, decode bit-exact."""
import numpy as np
import pytest

from gpu_common import first_diff

pytestmark = pytest.mark.gpu
NULL = -(2 ** 31)


@pytest.fixture(scope="module")
def g4():
    import gridfour_b200

    return gridfour_b200


def null_tiles(oracle):
    rng = np.random.default_rng(11)
    tiles = {}
    t = oracle.terrain_i32(500, 700, 90, 120).copy()
    t[rng.random(t.shape) < 0.05] = NULL
    tiles["sparse_nulls"] = t
    t = oracle.terrain_i32(0, 0, 61, 77).copy()
    t[:, 0] = NULL  # every row starts with a null: each row restarts from the seed
    t[10:20, 30:50] = NULL
    tiles["null_first_column"] = t
    t = oracle.terrain_i32(3000, 100, 45, 60).copy()
    t[0, 0] = NULL
    t[-1, -1] = NULL
    tiles["corners"] = t
    t = np.full((20, 30), NULL, np.int32)
    t[7, 11] = 1234  # a single valid cell
    tiles["single_valid"] = t
    t = rng.integers(-(2 ** 31) + 1, 2 ** 31, (24, 40), dtype=np.int64).astype(np.int32)  # overflowing deltas
    t[rng.random(t.shape) < 0.3] = NULL
    tiles["noise_nulls"] = t
    t = oracle.terrain_i32(9, 9, 180, 240).copy()
    t[60:120, 80:160] = NULL  # a void, as in SRTM
    tiles["void_180x240"] = t
    return tiles


@pytest.mark.parametrize("codec", ["CodecHuffman", "CodecDeflate", "CodecCanonHuffman"])
def test_null_tiles_encode_and_decode(g4, oracle, codec):
    cls = getattr(g4, codec)
    oid = {"CodecHuffman": oracle.CODEC_HUFFMAN, "CodecDeflate": oracle.CODEC_DEFLATE,
           "CodecCanonHuffman": oracle.CODEC_CANON_HUFFMAN}[codec]
    for name, tile in null_tiles(oracle).items():
        want, _ = oracle.codec_encode_i32(oid, 1, tile)
        got = cls().encode(1, tile.shape[0], tile.shape[1], tile)
        assert got is not None, "%s %s" % (codec, name)
        assert got[1] == 4, "%s %s: predictor code %d" % (codec, name, got[1])
        assert got == want, "%s %s: %s" % (codec, name, first_diff(got, want))
        out = cls().decode(tile.shape[0], tile.shape[1], want)
        assert np.array_equal(out, tile), "%s %s: %s" % (codec, name, first_diff(out, tile))


def test_all_null_tile_is_declined_by_every_int_codec(g4):
    t = np.full((8, 8), NULL, np.int32)
    for cls in (g4.CodecHuffman, g4.CodecDeflate, g4.CodecCanonHuffman):
        assert cls().encode(0, 8, 8, t) is None


def test_batched_band_with_null_tiles(g4, oracle):
    grid = oracle.terrain_i32(0, 0, 2 * 90, 2 * 120).copy()
    grid[20:70, 130:200] = NULL
    spec = g4.CodecSpecification(default=False)
    spec.addCompressionCodec("GvrsHuffman", g4.CodecHuffman)
    spec.addCompressionCodec("GvrsDeflate", g4.CodecDeflate)
    master = g4.CodecMaster(spec)
    batch = master.encodeTiles(grid, 90, 120)
    for t in range(4):
        tr, tc = divmod(t, 2)
        tile = grid[tr * 90:(tr + 1) * 90, tc * 120:(tc + 1) * 120]
        want = oracle.master_encode_i32([0, 1], tile)
        assert batch.payload(t) == want, "tile %d: %s" % (t, first_diff(batch.payload(t), want))
    assert np.array_equal(master.decodeTiles(batch), grid)
