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


def _ftile(oracle, rng, kind, r, c, k):
    t = oracle.terrain_f32(50 * k, 31 * k, r, c).copy()
    if kind == "terrain":
        return t
    if kind == "constant":
        return np.full((r, c), 2.5, np.float32)
    if kind == "nan_fill":
        return np.full((r, c), np.nan, np.float32)
    if kind == "noise":
        return rng.standard_normal((r, c)).astype(np.float32)
    if kind == "bits":
        return rng.integers(0, 2 ** 32, (r, c), dtype=np.uint64).astype(np.uint32).view(np.float32)   # every bit pattern
    if kind == "specials":
        return rng.choice(np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-40, 3.4e38], np.float32), (r, c))
    if kind == "nan_void":
        t[r // 4:r // 2, c // 4:c // 2] = np.nan
        return t
    raise AssertionError(kind)


FKINDS = ["terrain", "constant", "nan_fill", "noise", "bits", "specials", "nan_void"]


@pytest.mark.parametrize("shape", [(8, 8), (17, 23), (60, 60), (90, 120)])
def test_mixed_float_batches_match_oracle_selection(g4, oracle, shape):
    r, c = shape
    rng = np.random.default_rng(2000 + r)
    td, ta = 2, 7
    kinds = [FKINDS[i % len(FKINDS)] for i in rng.permutation(td * ta)]
    grid = np.zeros((td * r, ta * c), np.float32)
    for t, kind in enumerate(kinds):
        tr, tc = divmod(t, ta)
        grid[tr * r:(tr + 1) * r, tc * c:(tc + 1) * c] = _ftile(oracle, rng, kind, r, c, t)
    for names, ids in ((["GvrsFloat"], [2]), (["GvrsHuffman", "GvrsFloat", "LSOP12"], [0, 2, 4])):
        spec = g4.CodecSpecification(default=False)
        for nme in names:
            cls = {"GvrsFloat": ("CodecFloat", "CodecFloat")}.get(nme) or STD[nme][:2]
            spec.addCompressionCodec(nme, getattr(g4, cls[0]), getattr(g4, cls[1]))
        master = g4.CodecMaster(spec)
        batch = master.encodeTiles(grid, r, c)
        for t, kind in enumerate(kinds):
            tr, tc = divmod(t, ta)
            tile = np.ascontiguousarray(grid[tr * r:(tr + 1) * r, tc * c:(tc + 1) * c])
            want = oracle.master_encode_f32(ids, tile)
            got = batch.payload(t)
            assert got == want, "%s tile %d (%s) codecs %s: %s" % (shape, t, kind, names, first_diff(got, want))
        out = master.decodeTiles(batch)
        assert np.array_equal(out.view(np.uint32), grid.view(np.uint32)), "%s %s" % (shape, names)


def test_mixed_short_batches(g4, oracle):
    """Short elements: the fill value is coded as null, every other 16-bit value (extremes included) survives; the fill
    value itself comes back the way the reference gives it back."""
    rng = np.random.default_rng(31)
    r, c, td, ta = 30, 40, 3, 4
    for fill in (-32768, 0, 32767, -9999):
        grid = np.zeros((td * r, ta * c), np.int16)
        for t in range(td * ta):
            tr, tc = divmod(t, ta)
            kind = t % 6
            if kind == 0:
                tile = (oracle.terrain_i32(7 * t, 3 * t, r, c) % 30000).astype(np.int16)
            elif kind == 1:
                tile = rng.integers(-32768, 32768, (r, c)).astype(np.int16)
            elif kind == 2:
                tile = np.full((r, c), fill, np.int16)
            elif kind == 3:
                tile = np.full((r, c), 1234, np.int16)
            elif kind == 4:
                tile = rng.integers(-3, 4, (r, c)).astype(np.int16)
                tile[rng.random((r, c)) < 0.1] = fill
            else:
                tile = rng.choice(np.array([-32768, -32767, -1, 0, 1, 32766, 32767], np.int16), (r, c))
            grid[tr * r:(tr + 1) * r, tc * c:(tc + 1) * c] = tile
        master = g4.CodecMaster()
        batch = master.encodeTiles(grid, r, c, fillValue=fill)
        out = master.decodeTiles(batch)
        assert out.dtype == np.int16
        # TileElementShort.decode (gvrs/TileElementShort.java:232-247): a compressed tile gives the null code back as
        # SHORT_NULL_CODE (-32768) whatever the element's fill value is; a raw tile gives the stored shorts back.
        want = grid.copy()
        n_raw = 0
        for t in range(td * ta):
            tr, tc = divmod(t, ta)
            if int(batch.lens[t]) == (2 * r * c + 3) // 4 * 4:
                n_raw += 1
                continue
            w = want[tr * r:(tr + 1) * r, tc * c:(tc + 1) * c]
            w[w == fill] = -32768
        assert 0 < n_raw < td * ta
        assert np.array_equal(out, want), fill


@pytest.mark.parametrize("shape,tiles", [((90, 121), (3, 3)), ((45, 62), (4, 5)), ((180, 243), (2, 3)), ((127, 509), (2, 2)),
                                          ((512, 512), (1, 2)), ((1024, 1024), (1, 1)), ((1000, 1047), (1, 1))])
def test_unaligned_and_large_tiles(g4, oracle, shape, tiles):
    """Tile widths that are not multiples of four (tiles start at unaligned addresses: the run sinks and the deferred LSOP
    path instead of the packed 16-byte stores) and tiles up to the library's limit of 2^20 samples (staging larger than
    the default shared-memory carve-out): packings are the oracle's, batches decode to the input."""
    r, c = shape
    td, ta = tiles
    grid = oracle.terrain_i32(321, 123, td * r, ta * c, n_threads=8)
    names = ["GvrsHuffman", "GvrsDeflate", "LSOP12"]
    spec = g4.CodecSpecification(default=False)
    for nme in names:
        spec.addCompressionCodec(nme, getattr(g4, STD[nme][0]), getattr(g4, STD[nme][1]))
    ids = [STD[nme][2] for nme in names]
    master = g4.CodecMaster(spec)
    batch = master.encodeTiles(grid, r, c)
    for t in range(td * ta):
        tr, tc = divmod(t, ta)
        tile = np.ascontiguousarray(grid[tr * r:(tr + 1) * r, tc * c:(tc + 1) * c])
        want = oracle.master_encode_i32(ids, tile)
        got = batch.payload(t)
        assert got == want, "%s tile %d: %s" % (shape, t, first_diff(got, want))
    assert np.array_equal(master.decodeTiles(batch), grid)
    # every codec on its own, so that the ones that lose the selection are decoded too
    for nme in names:
        s1 = g4.CodecSpecification(default=False)
        s1.addCompressionCodec(nme, getattr(g4, STD[nme][0]), getattr(g4, STD[nme][1]))
        m1 = g4.CodecMaster(s1)
        assert np.array_equal(m1.decodeTiles(m1.encodeTiles(grid, r, c)), grid), nme


@pytest.mark.parametrize("space", ["host", "device"])
@pytest.mark.parametrize("dtype", [np.int32, np.int16, np.float32])
def test_band_inside_a_wider_raster(g4, oracle, space, dtype):
    """g4_band_desc.grid_pitch larger than the band's width (a band cut out of a wider raster, the way a tile-cache window
    sits in a page): the same payloads as for the compact band, and the decode writes nothing outside the band."""
    import ctypes as C

    from gridfour_b200 import _lib
    from gridfour_b200._lib import G4_MEM_DEVICE, G4_MEM_HOST

    r, c, td, ta, pitch, col0 = 30, 44, 3, 4, 4 * 44 + 37, 5       # col0 = 5: band rows start at unaligned addresses
    if dtype == np.float32:
        compact = oracle.terrain_f32(10, 20, td * r, ta * c)
    else:
        compact = (oracle.terrain_i32(10, 20, td * r, ta * c) % 20000).astype(dtype)
    wide = np.full((td * r, pitch), 77, dtype)
    wide[:, col0:col0 + ta * c] = compact
    master = g4.CodecMaster()
    ref = master.encodeTiles(compact, r, c)
    L = _lib.lib()
    cl = master.spec.native_list()
    band = master._band(compact.shape, dtype, r, c, pitch=pitch)
    nT = td * ta
    cap = int(L.g4_encode_arena_bound(C.byref(band)))
    total = C.c_uint64(0)
    item = np.dtype(dtype).itemsize
    if space == "host":
        arena = np.empty(cap, np.uint8)
        offsets, lens = np.empty(nT, np.uint64), np.empty(nT, np.uint32)
        codec, pred, status = np.empty(nT, np.uint8), np.empty(nT, np.uint8), np.empty(nT, np.int32)
        ptr = lambda a: a.ctypes.data
        mem, src = G4_MEM_HOST, wide
        grid_ptr = wide.ctypes.data + col0 * item
    else:
        import torch

        dev = torch.device("cuda:0")
        arena = torch.empty(cap, dtype=torch.uint8, device=dev)
        offsets, lens = torch.empty(nT, dtype=torch.int64, device=dev), torch.empty(nT, dtype=torch.int32, device=dev)
        codec, pred = torch.empty(nT, dtype=torch.uint8, device=dev), torch.empty(nT, dtype=torch.uint8, device=dev)
        status = torch.empty(nT, dtype=torch.int32, device=dev)
        ptr = lambda a: a.data_ptr()
        mem, src = G4_MEM_DEVICE, torch.from_numpy(wide).to(dev)
        grid_ptr = src.data_ptr() + col0 * item
    st = L.g4_encode_tiles(master._context()._h, C.byref(cl), C.byref(band), mem, grid_ptr, ptr(arena), cap, ptr(offsets), ptr(lens),
                           ptr(codec), ptr(pred), ptr(status), C.byref(total))
    assert st == 0
    to_np = (lambda a: a) if space == "host" else (lambda a: a.cpu().numpy())
    a_np, o_np, l_np = to_np(arena), to_np(offsets).astype(np.uint64), to_np(lens).astype(np.uint32)
    for t in range(nT):
        got = a_np[int(o_np[t]):int(o_np[t]) + int(l_np[t])].tobytes()
        assert got == ref.payload(t), "tile %d: %s" % (t, first_diff(got, ref.payload(t)))
    # decode into a fresh wide raster: the band is the input, everything around it keeps its marker
    if space == "host":
        out = np.full((td * r, pitch), 55, dtype)
        out_ptr = out.ctypes.data + col0 * item
    else:
        out_t = torch.full((td * r, pitch), 55, dtype=src.dtype, device=dev)
        out_ptr = out_t.data_ptr() + col0 * item
    st = L.g4_decode_tiles(master._context()._h, C.byref(cl), C.byref(band), mem, ptr(arena), ptr(offsets), ptr(lens), out_ptr, ptr(status))
    assert st == 0
    if space == "device":
        master._context().synchronize()
        out = out_t.cpu().numpy()
    assert np.array_equal(out[:, col0:col0 + ta * c].view(np.uint8), compact.view(np.uint8))
    outside = np.ones(out.shape, bool)
    outside[:, col0:col0 + ta * c] = False
    assert np.all(out[outside] == 55)
