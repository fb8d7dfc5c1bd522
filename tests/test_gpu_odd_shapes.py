"""-m gpu: degenerate and odd tile shapes (2x2 ... 300x2, 65x63) and extreme sample values (five-byte M32
codes, overflowing deltas, constants) through every codec: the packing is the oracle's byte for byte (or both decline),
and both packings decode to the input.  The reference's own tests use round tiles only; these are the shapes its
predictors special-case (first row / first column / rows shorter than the stencil)."""
import numpy as np
import pytest

from gpu_common import first_diff

pytestmark = pytest.mark.gpu

SHAPES = [(2, 2), (2, 3), (3, 2), (3, 3), (4, 5), (5, 6), (6, 5), (6, 6), (7, 13), (13, 7), (16, 16), (31, 33), (65, 63), (2, 300),
          (300, 2), (6, 257), (257, 6)]


@pytest.fixture(scope="module")
def g4():
    import gridfour_b200

    return gridfour_b200


def _int_tiles(oracle, shape, rng):
    r, c = shape
    yield "terrain", oracle.terrain_i32(123, 457, r, c)
    yield "constant", np.full(shape, -7, np.int32)
    yield "ramp", (np.arange(r * c, dtype=np.int64).reshape(shape) * 3 - 1000).astype(np.int32)
    yield "small_noise", rng.integers(-3, 4, shape).astype(np.int32)
    yield "byte_noise", rng.integers(-300, 300, shape).astype(np.int32)
    yield "wide_noise", rng.integers(-(2 ** 31) + 1, 2 ** 31, shape, dtype=np.int64).astype(np.int32)  # overflowing deltas
    t = oracle.terrain_i32(5, 5, r, c).copy()
    t.flat[rng.integers(0, r * c)] = 2 ** 31 - 1
    t.flat[rng.integers(0, r * c)] = -(2 ** 31) + 1
    yield "spikes", t


@pytest.mark.parametrize("codec", ["CodecHuffman", "CodecDeflate", "CodecCanonHuffman", "CodecLsop12"])
def test_int_codecs_on_odd_shapes(g4, oracle, codec):
    enc_cls, dec_cls = (g4.LsEncoder12, g4.LsDecoder12) if codec == "CodecLsop12" else (getattr(g4, codec),) * 2
    oid = {"CodecHuffman": oracle.CODEC_HUFFMAN, "CodecDeflate": oracle.CODEC_DEFLATE,
           "CodecCanonHuffman": oracle.CODEC_CANON_HUFFMAN, "CodecLsop12": oracle.CODEC_LSOP12}[codec]
    rng = np.random.default_rng(2024)
    n_packed = 0
    for shape in SHAPES:
        for name, tile in _int_tiles(oracle, shape, rng):
            tag = "%s %s %s" % (codec, shape, name)
            want, _ = oracle.codec_encode_i32(oid, 2, tile)
            got = enc_cls().encode(2, shape[0], shape[1], tile)
            if want is None:
                assert got is None, tag + ": the oracle declines, the GPU packs"
                continue
            assert got is not None, tag + ": the GPU declines, the oracle packs"
            assert got == want, tag + ": " + first_diff(got, want)
            out = dec_cls().decode(shape[0], shape[1], want)
            assert np.array_equal(out, tile), tag + ": " + first_diff(out, tile)
            n_packed += 1
    assert n_packed > 40


def test_float_codec_on_odd_shapes(g4, oracle):
    rng = np.random.default_rng(7)
    for shape in SHAPES:
        r, c = shape
        tiles = {"terrain": oracle.terrain_f32(11, 13, r, c), "constant": np.full(shape, 1.5, np.float32),
                 "noise": rng.standard_normal(shape).astype(np.float32),
                 "specials": rng.choice(np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, 3.4e38], np.float32), shape)}
        for name, tile in tiles.items():
            tag = "CodecFloat %s %s" % (shape, name)
            want = oracle.codec_encode_f32(2, tile)
            got = g4.CodecFloat().encodeFloats(2, r, c, tile)
            assert got == want, tag + ": " + first_diff(got, want)
            out = g4.CodecFloat().decodeFloats(r, c, want)
            assert np.array_equal(out.view(np.uint32), tile.view(np.uint32)), tag


def test_tiles_thinner_than_two_cells_are_declined(g4):
    """The Linear predictor reads columns 0 and 1 of every row (PredictorModelLinear.java:120-150), so the reference
    has no defined result for a one-column tile.  The per-tile ENCODE entry declines such tiles like a reference codec
    that returns null (CodecMaster then stores the tile raw: ICompressionEncoder.java:61); the band entry points and
    decode refuse the shape (g4codec.h, G4_ERR_UNSUPPORTED)."""
    for shape in ((1, 8), (8, 1), (1, 1)):
        tile = np.zeros(shape, np.int32)
        assert g4.CodecHuffman().encode(0, shape[0], shape[1], tile) is None
        assert g4.LsEncoder12().encode(0, shape[0], shape[1], tile) is None
        assert g4.CodecMaster().encode(shape[0], shape[1], tile) is None  # every codec declines -> the caller stores raw
        with pytest.raises(Exception):
            g4.CodecMaster().encodeTiles(tile, shape[0], shape[1])
        with pytest.raises(Exception):
            g4.CodecHuffman().decode(shape[0], shape[1], bytes(16))


def test_canonical_inconsistent_escape_range_follows_the_reference(g4, oracle):
    """CanonicalHuffman counts residuals in [-8388608, -8333609] as 16-bit escapes but writes them as 24-bit escapes
    (CanonicalHuffman.java:258 against :395).  The packing the reference produces is valid whenever the symbol of -1 has a
    code; the GPU encoder must produce exactly that packing (and the same predictor choice, which goes by real lengths)."""
    rng = np.random.default_rng(99)
    base = oracle.terrain_i32(40, 40, 32, 48)
    hit = 0
    for k, residual in enumerate((-8388608, -8333609, -8360000, -8350001)):
        tile = base.copy()
        r, c = 5 + 3 * k, 7 + 5 * k
        tile[r, c] = tile[r, c - 1] + residual          # a differencing residual inside the range
        tile[r + 1, c + 2] += int(rng.integers(-40000, -33000))
        want, pred = oracle.codec_encode_i32(oracle.CODEC_CANON_HUFFMAN, 3, tile)
        got = g4.CodecCanonHuffman().encode(3, 32, 48, tile)
        assert want is not None and got is not None, k
        assert got == want, "case %d: %s" % (k, first_diff(got, want))
        assert np.array_equal(g4.CodecCanonHuffman().decode(32, 48, got), tile), k
        assert np.array_equal(oracle.codec_decode_i32(oracle.CODEC_CANON_HUFFMAN, 32, 48, got), tile), k
        text, _ = oracle.canon_decode(want[6:], 32 * 48 - 1)
        hit += int(np.any((text >= -8388608) & (text <= -8333609)))
    assert hit >= 1                                      # at least one winning predictor really carried such a residual
