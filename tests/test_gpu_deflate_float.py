"""-m gpu: the hand-written RFC1950/1951 decoder behind CodecDeflate, CodecFloat and LSOP12 type 1, checked against
the reference's own fixtures (streams written by java.util.zip) and against oracle/zlib streams."""
import json
import os
import zlib

import numpy as np
import pytest

from gpu_common import first_diff, parity_grids

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "fixtures.json")))["samples"]


@pytest.fixture(scope="module")
def g4():
    import gridfour_b200

    return gridfour_b200


def expected_tile(sample, tile_index):
    s = GOLD[sample]
    tcols = s["grid_cols"] // s["tile_cols"]
    tr, tc = divmod(tile_index, tcols)
    r = np.arange(s["tile_rows"])[:, None] + tr * s["tile_rows"]
    c = np.arange(s["tile_cols"])[None, :] + tc * s["tile_cols"]
    return (r * s["grid_cols"] + c - 1).astype(np.int64)


@pytest.mark.parametrize("sample", ["Sample04_ShortComp", "Sample05_IntComp", "Sample07_ICFComp"])
def test_reference_deflate_fixtures_decode_on_gpu(g4, sample):
    s = GOLD[sample]
    dec = g4.CodecDeflate()
    for k, hexs in s["tiles"].items():
        out = dec.decode(s["tile_rows"], s["tile_cols"], bytes.fromhex(hexs))
        assert np.array_equal(out, expected_tile(sample, int(k))), "%s tile %s" % (sample, k)


def test_reference_float_fixture_decodes_on_gpu(g4):
    s = GOLD["Sample06_FltComp"]
    dec = g4.CodecFloat()
    for k, hexs in s["tiles"].items():
        out = dec.decodeFloats(s["tile_rows"], s["tile_cols"], bytes.fromhex(hexs))
        exp = expected_tile("Sample06_FltComp", int(k)).astype(np.float32)
        assert np.array_equal(out.view(np.uint32), exp.view(np.uint32)), "tile %s" % k


def test_deflate_decode_bit_exact(g4, oracle):
    dec = g4.CodecDeflate()
    for name, grid in parity_grids(oracle).items():
        packing, _ = oracle.codec_encode_i32(oracle.CODEC_DEFLATE, 1, grid)
        out = dec.decode(grid.shape[0], grid.shape[1], packing)
        assert np.array_equal(out, grid), "%s: %s" % (name, first_diff(out, grid))


@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_inflate_block_types(g4, oracle, level):
    """Stored (level 0), fixed and dynamic blocks, multi-block streams: build CodecDeflate packings by hand."""
    dec = g4.CodecDeflate()
    for name in ("terrain180x240", "noise", "steps", "const"):
        grid = parity_grids(oracle)[name]
        for pred in (1, 3):
            n, seed, m32 = oracle.predictor_encode(pred, grid)
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED):
                co = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
                z = co.compress(m32) + co.flush()
                packing = bytes([1, pred]) + int(seed).to_bytes(4, "little", signed=True) + int(n).to_bytes(4, "little") + z
                out = dec.decode(grid.shape[0], grid.shape[1], packing)
                assert np.array_equal(out, grid), "%s pred %d level %d strat %d: %s" % (name, pred, level, strategy,
                                                                                      first_diff(out, grid))


def test_inflate_rejects_corrupt_streams(g4, oracle):
    dec = g4.CodecDeflate()
    grid = parity_grids(oracle)["terrain90x120"]
    packing, _ = oracle.codec_encode_i32(oracle.CODEC_DEFLATE, 1, grid)
    bad = bytearray(packing)
    bad[-1] ^= 0x55  # Adler-32 trailer
    with pytest.raises(IOError):
        dec.decode(90, 120, bytes(bad))
    bad = bytearray(packing)
    bad[10] = 0x79  # zlib header check bits
    with pytest.raises(IOError):
        dec.decode(90, 120, bytes(bad))
    with pytest.raises(IOError):
        dec.decode(90, 120, packing[: len(packing) // 2])


def test_float_decode_bit_exact(g4, oracle):
    rng = np.random.default_rng(4)
    dec = g4.CodecFloat()
    tiles = [oracle.terrain_f32(0, 0, 30, 40), oracle.terrain_f32(100, 200, 120, 120),
             rng.integers(0, 2 ** 32, (16, 16), dtype=np.uint64).astype(np.uint32).view(np.float32),
             np.zeros((3, 5), np.float32), oracle.terrain_f32(9, 9, 37, 41)]
    for t in tiles:
        p = oracle.codec_encode_f32(2, t)
        out = dec.decodeFloats(t.shape[0], t.shape[1], p)
        assert np.array_equal(out.view(np.uint32), t.view(np.uint32)), first_diff(out.view(np.uint32), t.view(np.uint32))


def test_lsop_type1_two_zlib_streams(g4, oracle):
    """A repetitive tile makes the reference pick Deflate for the LSOP residuals (LsEncoder12.java:180-218)."""
    r, c = np.mgrid[0:64, 0:64]
    found = 0
    for k in range(6):
        rep = ((r % (3 + k)) * 1000 + (c % (5 + k)) * 37 + (r // 16) * 5 + (r * c) % (2 + k)).astype(np.int32)
        p = oracle.lsop12_encode(0, rep)
        if p is None or (p[1] & 0x0F) != 1:
            continue
        found += 1
        out = g4.LsDecoder12().decode(64, 64, p)
        assert np.array_equal(out, rep), first_diff(out, rep)
    assert found >= 1
