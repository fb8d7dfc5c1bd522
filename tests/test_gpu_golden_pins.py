"""-m gpu: the encode-side pins of tests/test_oracle_golden.py through the CUDA path (C ABI via the host mirror).

The reference's sample files were written by its Java encoders from documented closed-form grids
(core/src/test/resources/org/gridfour/gvrs/SampleFiles/README.txt:8-55), so they fix encoder output: the GPU encoders
must reproduce the stored packings (CodecDeflate), the stored coefficients (LSOP12) and the stored planes (CodecFloat)."""
import json
import math
import os
import zlib

import numpy as np
import pytest

from test_oracle_golden import CODEC_IDS, GOLD, expected_tile, float_planes, sample14_grid, sample14_parts

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g4():
    import gridfour_b200

    return gridfour_b200


@pytest.mark.parametrize("sample", ["Sample04_ShortComp", "Sample05_IntComp", "Sample07_ICFComp"])
def test_gpu_deflate_encoder_reproduces_the_jdk_written_packings(g4, sample):
    s = GOLD[sample]
    enc = g4.CodecDeflate()
    for k, hexs in s["tiles"].items():
        want = bytes.fromhex(hexs)
        tile = expected_tile(sample, int(k)).astype(np.int32)
        got = enc.encode(want[0], s["tile_rows"], s["tile_cols"], tile)
        assert got == want, "tile %s: %d vs %d bytes" % (k, len(got or b""), len(want))


def test_gpu_lsop_coefficients_equal_the_jvm_written_ones(g4, oracle):
    """Sample14's header holds the coefficients the JVM computed; the GPU fit (FP64 moments, one-thread JAMA-order LU)
    must give the same 48 bytes, the same seed, and a packing the oracle decodes to the grid."""
    _p, seed, coef, _ni, _nn, _body = sample14_parts()
    grid = sample14_grid()
    got = g4.LsEncoder12().encode(0, 101, 101, grid)
    assert got is not None and got[1] & 0x40  # revised header (LsHeader.java:220-245)
    assert got[2] == 12
    assert int.from_bytes(got[3:7], "little", signed=True) == seed
    assert bytes(got[7:55]) == coef.tobytes()
    assert np.array_equal(oracle.codec_decode_i32(oracle.CODEC_LSOP12, 101, 101, got), grid)
    assert np.array_equal(g4.LsDecoder12().decode(101, 101, got), grid)
    # and the reference-written packing itself (legacy header, legacy Huffman) decodes on the GPU
    assert np.array_equal(g4.LsDecoder12().decode(101, 101, _p), grid)


def test_gpu_float_encoder_carries_the_documented_planes(g4, oracle):
    s = GOLD["Sample06_FltComp"]
    enc = g4.CodecFloat()
    for k, hexs in s["tiles"].items():
        ref = bytes.fromhex(hexs)
        tile = expected_tile("Sample06_FltComp", int(k)).astype(np.float32)
        want = float_planes(tile)
        got = enc.encodeFloats(ref[0], s["tile_rows"], s["tile_cols"], tile)
        assert got == oracle.codec_encode_f32(ref[0], tile)
        off = 2
        for i in range(5):
            n = int.from_bytes(got[off:off + 4], "little")
            assert got[off + 4:off + 6] == b"\x78\xda"  # Deflater(9), CodecFloat.java:271
            assert zlib.decompress(got[off + 4:off + 4 + n]) == want[i], "tile %s plane %d" % (k, i)
            off += 4 + n
        assert off == len(got)
        assert np.array_equal(enc.decodeFloats(s["tile_rows"], s["tile_cols"], ref).view(np.uint32), tile.view(np.uint32))


def test_gpu_huffman_encoder_on_the_sample14_grid(g4, oracle):
    """H1/H2 tie-breaking on the pinned oracle: CodecHuffman of the Sample14 grid, GPU == oracle byte for byte."""
    grid = sample14_grid()
    want, _ = oracle.codec_encode_i32(oracle.CODEC_HUFFMAN, 0, grid)
    assert g4.CodecHuffman().encode(0, 101, 101, grid) == want


# ---- config-5 tile shapes against the oracle (they were round-trip-only before) -----------------------------------------
SHAPES5 = [(60, 60), (120, 120), (128, 128), (256, 256), (512, 512)]


def _tiles5(oracle, shape, rng):
    r, c = shape
    t = oracle.terrain_i32(9000, 17000, r, c)
    yield "terrain", t
    yield "terrain+escapes", (t + (rng.random(shape) < 0.001) * rng.integers(-3000, 3000, shape)).astype(np.int32)
    yield "terrain+many escapes", (t + (rng.random(shape) < 0.05) * rng.integers(-3000, 3000, shape)).astype(np.int32)
    yield "beyond 2^21", (t.astype(np.int64) * 400).astype(np.int32)
    big = t.copy()
    big[r // 2, c // 2] = 2 ** 31 - 1
    yield "one huge spike", big


@pytest.mark.parametrize("shape", SHAPES5)
def test_config5_shapes_match_the_oracle(g4, oracle, shape):
    rng = np.random.default_rng(5)
    r, c = shape
    codecs = [("GvrsHuffman", g4.CodecHuffman, g4.CodecHuffman, oracle.CODEC_HUFFMAN),
              ("GvrsCanonicalHuffman", g4.CodecCanonHuffman, g4.CodecCanonHuffman, oracle.CODEC_CANON_HUFFMAN),
              ("LSOP12", g4.LsEncoder12, g4.LsDecoder12, oracle.CODEC_LSOP12)]
    if r * c <= 128 * 128:
        codecs.append(("GvrsDeflate", g4.CodecDeflate, g4.CodecDeflate, oracle.CODEC_DEFLATE))
    for name, tile in _tiles5(oracle, shape, rng):
        for cname, enc, dec, oid in codecs:
            tag = "%s %s %s" % (cname, shape, name)
            want, _ = oracle.codec_encode_i32(oid, 1, tile)
            got = enc().encode(1, r, c, tile)
            assert (got is None) == (want is None), tag
            if want is None:
                continue
            assert got == want, tag
            assert np.array_equal(dec().decode(r, c, want), tile), tag


def test_lsop_fast_path_exception_and_range_tiles_in_a_band(g4, oracle):
    """A band that mixes plain tiles with tiles the fast LSOP path has to treat specially: residuals that are no byte
    (exception list), more of them than the list holds, and values beyond the 2^21 range of the fast arithmetic (both
    handed to the general kernels).  Every payload equals the oracle's and the band decodes to the input."""
    rng = np.random.default_rng(11)
    tr, tc = 180, 240
    base = oracle.terrain_i32(2000, 0, 2 * tr, 6 * tc).copy()
    kinds = ["plain", "few", "many", "range", "spike", "plain", "few", "plain", "many", "plain", "range", "few"]
    for t, kind in enumerate(kinds):
        r0, c0 = (t // 6) * tr, (t % 6) * tc
        v = base[r0:r0 + tr, c0:c0 + tc]
        if kind == "few":
            v += ((rng.random((tr, tc)) < 0.0005) * rng.integers(-5000, 5000, (tr, tc))).astype(np.int32)
        elif kind == "many":
            v += ((rng.random((tr, tc)) < 0.02) * rng.integers(-5000, 5000, (tr, tc))).astype(np.int32)
        elif kind == "range":
            v *= 500
        elif kind == "spike":
            v[90, 100] = -(2 ** 31) + 5
    spec = g4.CodecSpecification(default=False)
    spec.addCompressionCodec("LSOP12", g4.LsEncoder12, g4.LsDecoder12)
    master = g4.CodecMaster(spec)
    batch = master.encodeTiles(base, tr, tc)
    for t in range(len(kinds)):
        r0, c0 = (t // 6) * tr, (t % 6) * tc
        want = oracle.master_encode_i32([oracle.CODEC_LSOP12], base[r0:r0 + tr, c0:c0 + tc])
        assert batch.payload(t) == want, "tile %d (%s)" % (t, kinds[t])
    out = master.decodeTiles(batch)
    bad = np.nonzero(out != base)
    assert bad[0].size == 0, "first mismatch at %s" % ((bad[0][0], bad[1][0]),)
