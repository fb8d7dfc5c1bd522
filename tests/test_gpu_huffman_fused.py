"""-m gpu: the fused fast path of CodecHuffman decode (g4_huff2.cuh) against the oracle.

The fast path takes Triangle-predicted tiles whose M32 codes are all one byte long and hands every other tile to the
general kernel through a defer list; both routes, and the hand-over between them inside one batch, must give the oracle's
values.  Reference: compress/CodecHuffman.java:133-153, HuffmanDecoder.java:65-187, PredictorModelTriangle.java:160-198.
"""
import numpy as np
import pytest

from gpu_common import first_diff

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g4():
    import gridfour_b200

    return gridfour_b200


def _packing(oracle, pred, grid, index=0):
    """A CodecHuffman packing of `grid` with predictor `pred` forced (the encoder itself picks the smallest)."""
    n, seed, m32 = oracle.predictor_encode(pred, grid)
    text, _ = oracle.huffman_encode(np.frombuffer(m32, np.uint8))
    return bytes([index, pred]) + int(seed).to_bytes(4, "little", signed=True) + int(n).to_bytes(4, "little") + text, n


def _batch(g4, payloads, band):
    offsets, lens, pos = [], [], 0
    for p in payloads:
        offsets.append(pos)
        lens.append(len(p))
        pos += (len(p) + 7) & ~7
    arena = np.zeros(pos + 64, np.uint8)
    for o, p in zip(offsets, payloads):
        arena[o:o + len(p)] = np.frombuffer(p, np.uint8)
    return g4.TileBatch(arena, np.asarray(offsets, np.uint64), np.asarray(lens, np.uint32), None, None, None, pos, band)


@pytest.mark.parametrize("shape", [(90, 120), (180, 240), (33, 44), (64, 260), (40, 516), (70, 62), (37, 1030)])
def test_triangle_one_byte_tiles(g4, oracle, shape):
    """Fast-path tiles of several shapes: widths below / above 256 and 512, not multiples of 4, rows not multiples of 32."""
    r, c = shape
    rng = np.random.default_rng(r * 1000 + c)
    # a smooth field with small second differences: every Triangle residual fits one M32 byte
    grid = (np.add.outer(np.arange(r) * 3, np.arange(c) * 2) + rng.integers(-20, 21, (r, c))).astype(np.int32)
    p, n = _packing(oracle, 3, grid)
    assert n == r * c - 1, "test data must be all one-byte codes"
    want = oracle.codec_decode_i32(oracle.CODEC_HUFFMAN, r, c, p)
    assert np.array_equal(want, grid)
    got = g4.CodecHuffman().decode(r, c, p)
    assert np.array_equal(got, grid), first_diff(got, grid)


def test_int_min_code_and_wraparound(g4, oracle):
    """The one-byte code 0x80 is INT_MIN (CodecM32.java:313-324); sums wrap modulo 2^32 like Java ints."""
    r, c = 48, 64
    grid = np.zeros((r, c), np.int64)
    grid[:, :] = np.add.outer(np.arange(r), np.arange(c))
    grid = grid.astype(np.int32)
    n, seed, m32 = oracle.predictor_encode(3, grid)
    m = bytearray(m32)
    assert n == r * c - 1
    for k in (5, 700, 2999):  # plant INT_MIN residuals: the decoded field changes, but both decoders must agree
        m[k] = 0x80
    text, _ = oracle.huffman_encode(np.frombuffer(bytes(m), np.uint8))
    p = bytes([0, 3]) + int(seed).to_bytes(4, "little", signed=True) + int(n).to_bytes(4, "little") + text
    want = oracle.codec_decode_i32(oracle.CODEC_HUFFMAN, r, c, p)
    got = g4.CodecHuffman().decode(r, c, p)
    assert np.array_equal(got, want), first_diff(got, want)


def test_fused_and_deferred_tiles_in_one_batch(g4, oracle):
    """Triangle / one-byte tiles (fast path) beside Differencing, Linear, multi-byte, single-symbol and malformed tiles."""
    tr, tc = 60, 80
    rng = np.random.default_rng(3)
    smooth = lambda k: (np.add.outer(np.arange(tr) * (k + 1), np.arange(tc)) + rng.integers(-9, 10, (tr, tc))).astype(np.int32)
    tiles, payloads, expect_bad = [], [], []
    for k in range(12):
        g = smooth(k)
        if k % 6 == 3:
            g = rng.integers(-40000, 40000, (tr, tc)).astype(np.int32)  # multi-byte M32 codes
        if k % 6 == 4:
            g = np.full((tr, tc), 7 * k, np.int32)  # constant tile: single-symbol tree
        pred = 3 if k % 6 in (0, 3, 4, 5) else (1 if k % 6 == 1 else 2)
        p, _ = _packing(oracle, pred, g)
        bad = k == 11
        if bad:
            q = bytearray(p)
            q[10] = 0xFF  # 256 leaves claimed, then a tree that cannot hold them
            q[11:40] = bytes(29)
            p = bytes(q)
        tiles.append(g)
        payloads.append(p)
        expect_bad.append(bad)
    spec = g4.CodecSpecification(default=False)
    spec.addCompressionCodec("GvrsHuffman", g4.CodecHuffman)
    master = g4.CodecMaster(spec)
    band = master._band((3 * tr, 4 * tc), np.int32, tr, tc)
    b = _batch(g4, payloads, band)
    out = np.zeros((3 * tr, 4 * tc), np.int32)
    try:
        master.decodeTiles(b, out=out)
    except g4.FormatError:
        pass
    st = np.asarray(master.lastStatus)
    for t, (g, bad) in enumerate(zip(tiles, expect_bad)):
        rr, cc = divmod(t, 4)
        if bad:
            try:
                oracle.codec_decode_i32(oracle.CODEC_HUFFMAN, tr, tc, payloads[t])
                oracle_rejects = False
            except Exception:
                oracle_rejects = True
            assert oracle_rejects and st[t] < 0, "tile %d: a malformed tree must be a format error (status %d)" % (t, st[t])
        else:
            assert st[t] == 0, (t, st[t])
            got = out[rr * tr:(rr + 1) * tr, cc * tc:(cc + 1) * tc]
            assert np.array_equal(got, g), "tile %d: %s" % (t, first_diff(got, g))


def test_deep_tree_goes_to_the_general_kernel(g4, oracle):
    """A Fibonacci-like symbol histogram gives a legacy Huffman tree deeper than the tree kernel's 62-level path masks."""
    r, c = 96, 128
    n = r * c - 1
    # residual bytes with counts 1, 1, 2, 3, 5, ... (as far as n allows) -> a maximally skewed tree
    counts, a, b = [], 1, 1
    while sum(counts) + a <= n and len(counts) < 40:
        counts.append(a)
        a, b = b, a + b
    sym = np.concatenate([np.full(k, i - 20, np.int64) for i, k in enumerate(counts)])
    res = np.zeros(n, np.int64)
    res[: sym.size] = sym[:n]
    rng = np.random.default_rng(9)
    rng.shuffle(res)
    # Triangle field -> grid (2-D prefix sum in stream order): build the grid the oracle's predictor maps to these residuals
    F = np.zeros((r, c), np.int64)
    F[0, 1:] = res[: c - 1]
    F[1:, 0] = res[c - 1: c - 1 + r - 1]
    F[1:, 1:] = res[c + r - 2:].reshape(r - 1, c - 1)
    grid = F.cumsum(axis=0).cumsum(axis=1).astype(np.int32)
    p, nn = _packing(oracle, 3, grid)
    assert nn == n
    got = g4.CodecHuffman().decode(r, c, p)
    assert np.array_equal(got, grid), first_diff(got, grid)


@pytest.mark.parametrize("shape,nspikes", [((90, 120), 1), ((90, 120), 40), ((180, 240), 7), ((180, 240), 900), ((64, 260), 12),
                                           ((48, 64), 3000)])
def test_triangle_tiles_with_multi_byte_codes(g4, oracle, shape, nspikes):
    """Spikes make some Triangle residuals need 2..6 M32 bytes: the fast path keeps one byte per residual plus an exception
    list (up to its capacity), beyond that -- or when the code bytes outgrow its buffer -- the tile goes to the general kernel."""
    r, c = shape
    rng = np.random.default_rng(r + c + nspikes)
    grid = (np.add.outer(np.arange(r) * 2, np.arange(c) * 3) + rng.integers(-15, 16, (r, c))).astype(np.int64)
    rr = rng.integers(0, r, nspikes)
    cc = rng.integers(0, c, nspikes)
    mag = rng.choice([200, 300, 20000, 3000000, 400000000, -2 ** 31 + 5], nspikes)
    grid[rr, cc] += mag
    grid = grid.astype(np.int32)  # wraps like Java ints
    p, n = _packing(oracle, 3, grid)
    assert n > r * c - 1
    want = oracle.codec_decode_i32(oracle.CODEC_HUFFMAN, r, c, p)
    assert np.array_equal(want, grid)
    got = g4.CodecHuffman().decode(r, c, p)
    assert np.array_equal(got, grid), first_diff(got, grid)


def test_multi_byte_stream_that_ends_inside_a_code(g4, oracle):
    """nM32 one byte short of a multi-byte code's end: the reference reads past its array (an exception); here a format error."""
    r, c = 40, 52
    grid = np.add.outer(np.arange(r), np.arange(c)).astype(np.int32)
    grid[r - 1, c - 1] += 100000  # the last residual takes several bytes
    n, seed, m32 = oracle.predictor_encode(3, grid)
    m = np.frombuffer(m32, np.uint8)[:-1]  # drop the final byte of the last code
    text, _ = oracle.huffman_encode(m)
    p = bytes([0, 3]) + int(seed).to_bytes(4, "little", signed=True) + int(n - 1).to_bytes(4, "little") + text
    with pytest.raises(IOError):
        g4.CodecHuffman().decode(r, c, p)
