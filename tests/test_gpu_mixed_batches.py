"""-m gpu: batched encodeTiles / decodeTiles over rasters whose tiles differ in kind (terrain, noise of several widths,
constants, nulls, voids, spikes, full-range samples), for several tile shapes and codec lists in different orders.  Every
payload is the oracle's best-of-list packing for that tile (CodecMaster.encode: first codec wins ties, raw when nothing
beats 4 bytes per sample) and the batch decodes to the input."""
import numpy as np
import pytest

from gpu_common import first_diff

pytestmark = pytest.mark.gpu
NULL = -(2 ** 31)

STD = {"GvrsHuffman": ("CodecHuffman", "CodecHuffman", 0), "GvrsDeflate": ("CodecDeflate", "CodecDeflate", 1),
       "GvrsCanonicalHuffman": ("CodecCanonHuffman", "CodecCanonHuffman", 3), "LSOP12": ("LsEncoder12", "LsDecoder12", 4)}


@pytest.fixture(scope="module")
def g4():
    import gridfour_b200

    return gridfour_b200


def _tile(oracle, rng, kind, r, c, k):
    t = oracle.terrain_i32(100 * k, 37 * k, r, c).copy()
    if kind == "terrain":
        return t
    if kind == "constant":
        return np.full((r, c), int(rng.integers(-5000, 5000)), np.int32)
    if kind == "all_null":
        return np.full((r, c), NULL, np.int32)
    if kind == "small_noise":
        return rng.integers(-4, 5, (r, c)).astype(np.int32)
    if kind == "wide_noise":
        return rng.integers(-(2 ** 31) + 1, 2 ** 31, (r, c), dtype=np.int64).astype(np.int32)
    if kind == "mid_noise":
        return rng.integers(-(2 ** 23), 2 ** 23, (r, c)).astype(np.int32)     # residuals around the 16/24-bit escape edge
    if kind == "sparse_nulls":
        t[rng.random((r, c)) < 0.03] = NULL
        return t
    if kind == "void":
        t[r // 3:2 * r // 3, c // 4:3 * c // 4] = NULL
        return t
    if kind == "spikes":
        for _ in range(3):
            t[rng.integers(0, r), rng.integers(0, c)] = int(rng.integers(-(2 ** 31) + 1, 2 ** 31))
        return t
    raise AssertionError(kind)


KINDS = ["terrain", "constant", "all_null", "small_noise", "wide_noise", "mid_noise", "sparse_nulls", "void", "spikes"]


@pytest.mark.parametrize("shape", [(8, 8), (17, 23), (32, 48), (60, 60), (90, 120)])
def test_mixed_batches_match_oracle_selection(g4, oracle, shape):
    r, c = shape
    rng = np.random.default_rng(1000 + r)
    lists = [["GvrsHuffman", "GvrsDeflate"], ["GvrsDeflate", "GvrsHuffman", "LSOP12"], ["LSOP12", "GvrsCanonicalHuffman"],
             ["GvrsCanonicalHuffman", "GvrsHuffman", "GvrsDeflate", "LSOP12"]]
    td, ta = 3, 6
    kinds = [KINDS[i % len(KINDS)] for i in rng.permutation(td * ta)]
    grid = np.zeros((td * r, ta * c), np.int32)
    for t, kind in enumerate(kinds):
        tr, tc = divmod(t, ta)
        grid[tr * r:(tr + 1) * r, tc * c:(tc + 1) * c] = _tile(oracle, rng, kind, r, c, t)
    for names in lists:
        spec = g4.CodecSpecification(default=False)
        for nme in names:
            spec.addCompressionCodec(nme, getattr(g4, STD[nme][0]), getattr(g4, STD[nme][1]))
        ids = [STD[nme][2] for nme in names]
        master = g4.CodecMaster(spec)
        batch = master.encodeTiles(grid, r, c)
        for t, kind in enumerate(kinds):
            tr, tc = divmod(t, ta)
            tile = np.ascontiguousarray(grid[tr * r:(tr + 1) * r, tc * c:(tc + 1) * c])
            want = oracle.master_encode_i32(ids, tile)
            got = batch.payload(t)
            tag = "%s tile %d (%s) codecs %s" % (shape, t, kind, names)
            assert got == want, tag + ": " + first_diff(got, want)      # the oracle returns the raw tile when nothing wins
        assert np.array_equal(master.decodeTiles(batch), grid), "%s %s" % (shape, names)
