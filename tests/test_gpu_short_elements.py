"""-m gpu: TileElementShort semantics around the integer codecs (gvrs/TileElementShort.java:211-248,
gvrs/TileElement.java:85-93): samples widened to int with fill value -> INT4_NULL_CODE, payloads identical to encoding the
widened tile, raw form = 2 bytes per sample rounded up to a multiple of 4 and chosen when nothing beats it, decoded nulls come
back as SHORT_NULL_CODE."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NULL = -(2 ** 31)


def test_short_band_matches_widened_int_encoding(oracle):
    import gridfour_b200 as g4

    rng = np.random.default_rng(5)
    tr, tc = 45, 61  # odd cell count: the raw form is padded to a multiple of 4 bytes
    grid = (oracle.terrain_i32(0, 0, 2 * tr, 3 * tc) // 2).astype(np.int16)
    fill = -9999
    grid[10:20, 5:30] = fill                                        # a void inside tile 0 -> nulls (predictor 4)
    grid[tr:2 * tr, tc:2 * tc] = rng.integers(-32768, 32768, (tr, tc)).astype(np.int16)  # noise tile -> raw (2 bytes/sample)
    spec = g4.CodecSpecification(default=False)
    spec.addCompressionCodec("GvrsHuffman", g4.CodecHuffman)
    spec.addCompressionCodec("GvrsDeflate", g4.CodecDeflate)
    master = g4.CodecMaster(spec)
    batch = master.encodeTiles(grid, tr, tc, fillValue=fill)
    std = (2 * tr * tc + 3) & ~3
    for t in range(6):
        r, c = divmod(t, 3)
        tile = grid[r * tr:(r + 1) * tr, c * tc:(c + 1) * tc]
        wide = tile.astype(np.int32)
        wide[tile == fill] = NULL
        want = oracle.master_encode_i32([0, 1], wide)  # CodecMaster over the widened tile, raw when >= 4n
        if want is None or len(want) >= std or len(want) == 4 * tr * tc:
            exp = tile.astype("<i2").tobytes().ljust(std, b"\0")
        else:
            exp = want
        assert batch.payload(t) == exp, "tile %d" % t
    assert int(batch.lens[4]) == std and int(batch.codec[4]) == 255
    out = master.decodeTiles(batch)
    exp = grid.copy()
    exp[grid == fill] = -32768  # TileElementShort.decode maps null to SHORT_NULL_CODE, not to the fill value
    assert out.dtype == np.int16 and np.array_equal(out, exp)
