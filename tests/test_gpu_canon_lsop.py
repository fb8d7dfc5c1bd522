"""-m gpu: CodecCanonHuffman and LSOP12 CUDA paths vs the CPU oracle and the reference's LSOP fixture."""
import json
import math
import os

import numpy as np
import pytest

from gpu_common import first_diff, parity_grids

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "fixtures.json")))["samples"]


@pytest.fixture(scope="module")
def g4():
    import gridfour_b200

    return gridfour_b200


def test_canon_decode_bit_exact(g4, oracle):
    codec = g4.CodecCanonHuffman()
    for name, grid in parity_grids(oracle).items():
        packing, _ = oracle.codec_encode_i32(oracle.CODEC_CANON_HUFFMAN, 3, grid)
        out = codec.decode(grid.shape[0], grid.shape[1], packing)
        assert np.array_equal(out, grid), "%s: %s" % (name, first_diff(out, grid))


def test_canon_decode_escape_ranges(g4, oracle):
    """Residual magnitudes that exercise every escape form (2/4/6-bit, 8/16/24-bit)."""
    codec = g4.CodecCanonHuffman()
    rng = np.random.default_rng(5)
    for lim in (300, 1500, 7000, 30000, 4_000_000, 2 ** 30):
        grid = rng.integers(-lim, lim, (40, 52), dtype=np.int64).cumsum(axis=1).astype(np.int32)
        packing, _ = oracle.codec_encode_i32(oracle.CODEC_CANON_HUFFMAN, 0, grid)
        if packing is None:
            continue
        out = codec.decode(40, 52, packing)
        assert np.array_equal(out, grid), "lim %d: %s" % (lim, first_diff(out, grid))


def test_lsop_reference_fixture_decodes_on_gpu(g4):
    """Sample14_LSOP.gvrs: legacy header + legacy Huffman + M32, written by the reference's Java LsEncoder."""
    s = GOLD["Sample14_LSOP"]
    payload = bytes.fromhex(s["tiles"]["0"])
    out = g4.LsDecoder12().decode(101, 101, payload)
    exp = np.zeros((101, 101), np.int32)
    for r in range(101):
        for c in range(101):
            exp[r, c] = math.floor(math.sin(c / 100.0 * math.pi) * math.sin(r / 100.0 * math.pi) * 1000.0 + 0.5)
    assert np.array_equal(out, exp), first_diff(out, exp)


def test_lsop_decode_bit_exact(g4, oracle):
    dec = g4.LsDecoder12()
    n = 0
    for name, grid in parity_grids(oracle).items():
        for kwargs in ({"deflate": False}, {"deflate": False, "checksum": True}):
            packing = oracle.lsop12_encode(2, grid, **kwargs)
            if packing is None:
                continue
            out = dec.decode(grid.shape[0], grid.shape[1], packing)
            assert np.array_equal(out, grid), "%s %s: %s" % (name, kwargs, first_diff(out, grid))
            n += 1
    assert n >= 10


def test_lsop_decode_many_shapes(g4, oracle):
    dec = g4.LsDecoder12()
    for (R, C) in [(6, 6), (7, 9), (33, 40), (34, 37), (35, 100), (64, 64), (66, 130), (120, 120), (200, 50)]:
        grid = oracle.terrain_i32(R * 7, C * 3, R, C)
        packing = oracle.lsop12_encode(0, grid, deflate=False)
        if packing is None:
            continue
        out = dec.decode(R, C, packing)
        assert np.array_equal(out, grid), "%dx%d: %s" % (R, C, first_diff(out, grid))


def test_batch_decode_mixed_codecs(g4, oracle):
    """decodeTiles over payloads the ORACLE encoded with best-of [Huffman, CanonHuffman, LSOP12 (canonical only)]."""
    spec = g4.CodecSpecification(default=False)
    spec.addCompressionCodec("GvrsHuffman", g4.CodecHuffman)
    spec.addCompressionCodec("GvrsCanonicalHuffman", g4.CodecCanonHuffman)
    spec.addCompressionCodec("LSOP12", g4.LsEncoder12, g4.LsDecoder12)
    master = g4.CodecMaster(spec)
    R, C, TR, TC = 90, 120, 3, 4
    grid = oracle.terrain_i32(2000, 1000, TR * R, TC * C)
    grid[0:R, C:2 * C] = 7  # uniform tile (t=1) -> canonical 6-byte shortcut
    payloads = []
    for t in range(TR * TC):
        tr, tc = divmod(t, TC)
        tile = grid[tr * R:(tr + 1) * R, tc * C:(tc + 1) * C]
        cands = [oracle.codec_encode_i32(0, 0, tile)[0], oracle.codec_encode_i32(3, 1, tile)[0],
                 oracle.lsop12_encode(2, tile, deflate=False)]
        cands = [c for c in cands if c is not None]
        best = min(cands, key=len)
        payloads.append(best if t % 3 else cands[t % len(cands)])  # mix codecs regardless of size
    offsets, arena, pos = [], bytearray(), 0
    for p in payloads:
        offsets.append(pos)
        arena += p + b"\0" * ((-len(p)) % 8)
        pos = len(arena)
    band = master._band(grid.shape, np.int32, R, C)
    batch = g4.TileBatch(np.frombuffer(bytes(arena), np.uint8), np.array(offsets, np.uint64),
                         np.array([len(p) for p in payloads], np.uint32), None, None, None, len(arena), band)
    out = master.decodeTiles(batch)
    assert np.array_equal(out, grid), first_diff(out, grid)
    used = sorted({p[0] for p in payloads})
    assert used == [0, 1, 2], used
    assert len(payloads[1]) == 6


def test_canon_encode_streams_byte_exact(g4, oracle):
    codec = g4.CodecCanonHuffman()
    for name, grid in parity_grids(oracle).items():
        exp, pred = oracle.codec_encode_i32(oracle.CODEC_CANON_HUFFMAN, 3, grid)
        got = codec.encode(3, grid.shape[0], grid.shape[1], grid)
        assert got is not None, name
        assert got == exp, "%s (pred gpu %d, oracle %d): %s" % (name, codec.lastPredictor, pred, first_diff(got, exp))


def test_canon_encode_package_merge_path(g4, oracle):
    """A tile whose residual histogram is Fibonacci-like pushes the Huffman depth past 15 (PackageMerge.java)."""
    fib = [1, 1]
    while len(fib) < 24:
        fib.append(fib[-1] + fib[-2])
    vals = np.repeat(np.arange(23) - 11, fib[1:24]).astype(np.int64)
    np.random.default_rng(0).shuffle(vals)
    n = 200 * 300
    res = np.zeros(n, np.int64)
    res[: vals.size] = vals[: n]
    grid = res.reshape(200, 300).cumsum(axis=1).astype(np.int32)
    exp, pred = oracle.codec_encode_i32(oracle.CODEC_CANON_HUFFMAN, 0, grid)
    got = g4.CodecCanonHuffman().encode(0, 200, 300, grid)
    assert got == exp, first_diff(got, exp)


def test_lsop_coefficients_and_streams_byte_exact(g4, oracle):
    enc = g4.LsEncoder12()
    checked = 0
    for name, grid in parity_grids(oracle).items():
        exp = oracle.lsop12_encode(1, grid)
        got = enc.encode(1, grid.shape[0], grid.shape[1], grid)
        if exp is None:
            assert got is None, name
            continue
        assert got is not None, name
        # 12 float32 coefficients at bytes 7..55: bit-identical (north_star allows 1 ulp)
        assert got[7:55] == exp[7:55], "%s coefficients differ" % name
        assert got == exp, "%s: %s" % (name, first_diff(got, exp))
        checked += 1
    assert checked >= 8


def test_batch_best_of_three_matches_oracle(g4, oracle):
    """encodeTiles with [GvrsHuffman, GvrsCanonicalHuffman, LSOP12]: per-tile choice and bytes equal the oracle's
    CodecMaster rule (LSOP12 with its default Deflate alternative)."""
    spec = g4.CodecSpecification(default=False)
    spec.addCompressionCodec("GvrsHuffman", g4.CodecHuffman)
    spec.addCompressionCodec("GvrsCanonicalHuffman", g4.CodecCanonHuffman)
    spec.addCompressionCodec("LSOP12", g4.LsEncoder12, g4.LsDecoder12)
    master = g4.CodecMaster(spec)
    R, C, TR, TC = 90, 120, 3, 3
    grid = oracle.terrain_i32(4000, 8000, TR * R, TC * C)
    grid[0:R, 0:C] = -5
    batch = master.encodeTiles(grid, R, C)
    for t in range(TR * TC):
        tr, tc = divmod(t, TC)
        tile = grid[tr * R:(tr + 1) * R, tc * C:(tc + 1) * C]
        cands = [oracle.codec_encode_i32(0, 0, tile)[0], oracle.codec_encode_i32(3, 1, tile)[0],
                 oracle.lsop12_encode(2, tile)]
        best = None
        for c in cands:
            if c is not None and (best is None or len(c) < len(best)):
                best = c
        assert batch.payload(t) == best, "tile %d: %s" % (t, first_diff(batch.payload(t), best))
    out = master.decodeTiles(batch)
    assert np.array_equal(out, grid)


def test_lsop_value_checksum_is_verified_on_the_gpu(g4, oracle):
    """LsDecoder12.java:153-158 recomputes the CRC-32C of the decoded values and prints on a mismatch.  The GPU does the
    same check per tile: a matching checksum is silent, a wrong one still delivers the values and is reported
    (status G4_CHECKSUM_MISMATCH -> ValueChecksumWarning), per tile in a batch."""
    import warnings

    tr, tc = 180, 240
    grid = oracle.terrain_i32(700, 900, tr, 3 * tc)
    tiles = [np.ascontiguousarray(grid[:, k * tc:(k + 1) * tc]) for k in range(3)]
    for deflate in (False, True):
        packs = [oracle.lsop12_encode(0, t, deflate=deflate, checksum=True) for t in tiles]
        assert all(p[1] & 0x80 for p in packs)
        dec = g4.LsDecoder12()
        with warnings.catch_warnings():
            warnings.simplefilter("error")
            assert np.array_equal(dec.decode(tr, tc, packs[0]), tiles[0])
        # the stored value is the reference's: CRC-32C of the little-endian samples
        def cks_pos(p):  # revised header (LsHeader.java:220-245): the two M32 lengths are present unless the body is canonical
            return 55 if (p[1] & 0x0F) == 2 else 63

        stored = int.from_bytes(packs[0][cks_pos(packs[0]):cks_pos(packs[0]) + 4], "little")
        assert stored == oracle.crc32c(tiles[0].astype("<i4").tobytes())
        bad = bytearray(packs[1])
        bad[cks_pos(packs[1])] ^= 0x01
        with pytest.warns(g4.ValueChecksumWarning):
            out = dec.decode(tr, tc, bytes(bad))
        assert np.array_equal(out, tiles[1])
        # batched: only the damaged tile is flagged
        spec = g4.CodecSpecification(default=False)
        spec.addCompressionCodec("LSOP12", g4.LsEncoder12, g4.LsDecoder12)
        master = g4.CodecMaster(spec)
        arena = bytearray()
        offsets, lens = [], []
        for p in (packs[0], bytes(bad), packs[2]):
            arena += bytes((-len(arena)) & 7)
            offsets.append(len(arena))
            lens.append(len(p))
            arena += p
        band = master._band((tr, 3 * tc), np.int32, tr, tc)
        b = g4.TileBatch(np.frombuffer(bytes(arena) + bytes(16), np.uint8), np.array(offsets, np.uint64), np.array(lens, np.uint32), None, None,
                         None, len(arena), band)
        with pytest.warns(g4.ValueChecksumWarning):
            got = master.decodeTiles(b)
        assert np.array_equal(got, grid)
        assert master.lastStatus.tolist() == [0, 2, 0]
