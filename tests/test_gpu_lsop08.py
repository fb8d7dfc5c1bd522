"""-m gpu: LsDecoder08 (the legacy 8-coefficient LSOP codec, decode only) on the GPU vs the CPU oracle.

Reference: lsop/LsDecoder08.java:65-163.  The reference holds no LSOP08 fixture and no longer registers the codec
(lsop/LsCodecUtility.java:73), so the packings come from the oracle's restatement of LsEncoder08 (oracle/g4o_lsop08.cpp):
parity unpinned by the reference, GPU vs oracle bit-exact.
"""
import numpy as np
import pytest

from gpu_common import first_diff, parity_grids

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g4():
    import gridfour_b200

    return gridfour_b200


def _packings(oracle):
    out = []
    for name, grid in parity_grids(oracle).items():
        p = oracle.lsop08_encode(3, grid)
        if p is not None:
            out.append((name, grid, p))
    return out


def test_lsop08_decode_bit_exact(g4, oracle):
    dec = g4.LsDecoder08()
    cases = _packings(oracle)
    kinds = set()
    for name, grid, p in cases:
        want = oracle.lsop08_decode(grid.shape[0], grid.shape[1], p)
        assert np.array_equal(want, grid), name  # the oracle's own round trip
        kinds.add(p[1] & 0x0F)
        got = dec.decode(grid.shape[0], grid.shape[1], p)
        assert np.array_equal(got, grid), "%s: %s" % (name, first_diff(got, grid))
    assert len(cases) >= 8
    assert kinds == {0, 1}, "both body types (legacy Huffman, two zlib streams) must be exercised: %s" % kinds


def test_lsop08_in_a_batch_and_encode_declines(g4, oracle):
    """A codec list naming LSOP08: tiles written by the (oracle's) legacy encoder decode in the batched call; the GPU
    encoder of the codec declines, so encodeTiles never picks it."""
    tr, tc = 45, 60
    grid = oracle.terrain_i32(500, 900, 2 * tr, 3 * tc)
    spec = g4.CodecSpecification(default=False)
    spec.addCompressionCodec("GvrsHuffman", g4.CodecHuffman)
    spec.addCompressionCodec("LSOP08", g4.LsEncoder08, g4.LsDecoder08)
    master = g4.CodecMaster(spec)
    assert g4.LsEncoder08().encode(1, tr, tc, grid[:tr, :tc]) is None
    batch = master.encodeTiles(grid, tr, tc)
    assert not (np.asarray(batch.codec) == 1).any()
    # splice oracle-written LSOP08 packings over the batch
    payloads = []
    for t in range(6):
        r, c = divmod(t, 3)
        sub = np.ascontiguousarray(grid[r * tr:(r + 1) * tr, c * tc:(c + 1) * tc])
        p = oracle.lsop08_encode(1, sub) if t % 2 == 0 else batch.payload(t)
        assert p is not None
        payloads.append(p)
    offsets, lens, pos = [], [], 0
    for p in payloads:
        offsets.append(pos)
        lens.append(len(p))
        pos += (len(p) + 7) & ~7
    arena = np.zeros(pos + 64, np.uint8)
    for o, p in zip(offsets, payloads):
        arena[o:o + len(p)] = np.frombuffer(p, np.uint8)
    b2 = g4.TileBatch(arena, np.asarray(offsets, np.uint64), np.asarray(lens, np.uint32), None, None, None, pos, batch.band)
    out = master.decodeTiles(b2)
    assert np.array_equal(out, grid), first_diff(out, grid)


def test_lsop08_malformed(g4, oracle):
    grid = oracle.terrain_i32(0, 0, 40, 50)
    p = bytearray(oracle.lsop08_encode(0, grid))
    dec = g4.LsDecoder08()
    with pytest.raises(g4.FormatError):
        dec.decode(40, 50, bytes(p[:20]))
    q = bytearray(p)
    q[2] = 12  # a 12-coefficient header handed to the 8-coefficient decoder
    with pytest.raises(g4.FormatError):
        dec.decode(40, 50, bytes(q))
