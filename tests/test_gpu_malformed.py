"""-m gpu: malformed packings must come back as FormatError (the reference throws IOException or reads unchecked; the GPU
path is bounds-checked) and must never take the process down.  Covers the staged fast paths (LSOP12 head/text kernels,
CodecCanonHuffman) and the legacy Huffman decoder."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def mutations(p, rng, n=40):
    """Truncations, bit flips and byte stomps of a valid packing (byte 0 = codec index is left alone)."""
    p = bytearray(p)
    out = [bytes(p[:k]) for k in (1, 5, 6, 7, 12, 20, 55, 56, 63, 64, len(p) // 3, len(p) // 2, len(p) - 1) if 0 < k < len(p)]
    for _ in range(n):
        q = bytearray(p)
        k = int(rng.integers(1, len(q)))
        mode = int(rng.integers(0, 3))
        if mode == 0:
            q[k] ^= 1 << int(rng.integers(0, 8))
        elif mode == 1:
            q[k] = int(rng.integers(0, 256))
        else:
            q[k:k + 8] = bytes(int(x) for x in rng.integers(0, 256, min(8, len(q) - k)))
        out.append(bytes(q))
    return out


@pytest.mark.parametrize("codec", ["LSOP12", "CodecCanonHuffman", "CodecHuffman", "CodecDeflate"])
def test_mutated_packings_decode_or_fail_cleanly(oracle, codec):
    import gridfour_b200 as g4

    rng = np.random.default_rng(99)
    tile = oracle.terrain_i32(0, 0, 180, 240)
    if codec == "LSOP12":
        good = oracle.lsop12_encode(0, tile)
        dec = g4.LsDecoder12()
    else:
        oid = {"CodecCanonHuffman": oracle.CODEC_CANON_HUFFMAN, "CodecHuffman": oracle.CODEC_HUFFMAN,
               "CodecDeflate": oracle.CODEC_DEFLATE}[codec]
        good, _ = oracle.codec_encode_i32(oid, 0, tile)
        dec = getattr(g4, codec)()
    assert np.array_equal(dec.decode(180, 240, good), tile)
    failed = 0
    for bad in mutations(good, rng):
        try:
            out = dec.decode(180, 240, bad)
            assert out is None or out.shape == (180, 240)  # a mutation may still be a valid stream of other values
        except (g4.FormatError, g4.G4Error):
            failed += 1
    assert failed >= 5
    # the context survives: a good packing still decodes afterwards
    assert np.array_equal(dec.decode(180, 240, good), tile)
