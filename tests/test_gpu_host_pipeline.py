"""-m gpu: the host-buffer (G4_MEM_HOST) decode path pipelines payload H2D, kernels and raster D2H over chunks of
tile rows once the raster exceeds 32 MB; the result must equal the single-shot path and the input."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_host_decode_pipeline_matches_input():
    import gridfour_b200 as g4
    from oracle import g4oracle

    rows, cols, tr, tc = 33 * 90, 24 * 120, 90, 120  # 34.2 MB raster, 792 tiles, 33 tile rows -> 16 uneven chunks
    grid = g4oracle.terrain_i32(0, 0, rows, cols, n_threads=8)
    grid[5 * 90:6 * 90, 2 * 120:3 * 120] = np.random.default_rng(3).integers(-(2 ** 31), 2 ** 31, (90, 120), dtype=np.int64).astype(np.int32)
    spec = g4.CodecSpecification(default=False)
    spec.addCompressionCodec("GvrsHuffman", g4.CodecHuffman)
    spec.addCompressionCodec("LSOP12", g4.LsEncoder12, g4.LsDecoder12)
    master = g4.CodecMaster(spec)
    batch = master.encodeTiles(grid, tr, tc)  # host path
    assert int((batch.codec == 255).sum()) == 1  # the noise tile is stored raw
    out = master.decodeTiles(batch)
    assert np.array_equal(out, grid)
    # tiles out of order in the arena (descending offsets): the call falls back to one chunk and still decodes
    order = np.argsort(-batch.offsets.astype(np.int64))
    arena2 = np.zeros_like(batch.arena)
    off2 = np.zeros_like(batch.offsets)
    pos = 0
    for t in order:
        n = int(batch.lens[t])
        arena2[pos:pos + n] = batch.arena[int(batch.offsets[t]):int(batch.offsets[t]) + n]
        off2[t] = pos
        pos += (n + 7) & ~7
    b2 = g4.TileBatch(arena2, off2, batch.lens, None, None, None, pos, batch.band)
    assert np.array_equal(master.decodeTiles(b2), grid)


def test_host_decode_pipeline_into_a_wider_raster():
    """The chunked path with grid_pitch wider than the band: every chunk copies back only the band's columns."""
    import ctypes as C

    import gridfour_b200 as g4
    from gridfour_b200 import _lib
    from gridfour_b200._lib import G4_MEM_HOST
    from oracle import g4oracle

    rows, cols, tr, tc, pitch, col0 = 33 * 90, 24 * 120, 90, 120, 24 * 120 + 50, 13
    grid = g4oracle.terrain_i32(50, 60, rows, cols, n_threads=8)
    master = g4.CodecMaster()
    batch = master.encodeTiles(grid, tr, tc)
    band = master._band(grid.shape, np.int32, tr, tc, pitch=pitch)
    out = np.full((rows, pitch), 55, np.int32)
    status = np.empty(33 * 24, np.int32)
    cl = master.spec.native_list()
    st = _lib.lib().g4_decode_tiles(master._context()._h, C.byref(cl), C.byref(band), G4_MEM_HOST, batch.arena.ctypes.data,
                                    batch.offsets.ctypes.data, batch.lens.ctypes.data, out.ctypes.data + 4 * col0, status.ctypes.data)
    assert st == 0
    assert np.array_equal(out[:, col0:col0 + cols], grid)
    assert np.all(out[:, :col0] == 55) and np.all(out[:, col0 + cols:] == 55)
