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
