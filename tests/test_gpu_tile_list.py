"""-m gpu: tile-LIST calls (g4_encode_tile_list / g4_decode_tile_list) and the batched tile cache built on them.

Reference call pattern: gvrs/RasterTileCache.java:253-294 (flush: whatever tiles are dirty, each in its own int[]),
:339-426 (readTile on a miss), gvrs/GvrsElement.java:298-404 (readBlock).  The payloads must equal what the oracle's
CodecMaster gives for the same tiles, wherever the tiles live."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CODECS = ["GvrsHuffman", "GvrsDeflate", "LSOP12"]


@pytest.fixture(scope="module")
def g4():
    import gridfour_b200

    return gridfour_b200


def _master(g4, names=CODECS):
    std = {"GvrsHuffman": (g4.CodecHuffman, g4.CodecHuffman), "GvrsDeflate": (g4.CodecDeflate, g4.CodecDeflate),
           "LSOP12": (g4.LsEncoder12, g4.LsDecoder12), "GvrsFloat": (g4.CodecFloat, g4.CodecFloat)}
    spec = g4.CodecSpecification(default=False)
    for n in names:
        spec.addCompressionCodec(n, *std[n])
    return g4.CodecMaster(spec)


def _scattered(oracle, rng, tr, tc, n, dtype=np.int32):
    """n tiles placed in one flat base array: some in their own contiguous buffers (pitch = tc), some inside wider
    rasters (pitch > tc), at offsets of every alignment class."""
    tiles, refs, pos = [], [], 3
    flat = np.zeros(n * (tr * (tc + 40) + 64) + 64, dtype)
    for k in range(n):
        t = oracle.terrain_i32(100 * k, 37 * k, tr, tc) if dtype == np.int32 else oracle.terrain_f32(100 * k, 37 * k, tr, tc)
        pitch = tc if k % 3 == 0 else tc + (8 if k % 3 == 1 else 13)
        pos += [0, 1, 4, 8][k % 4]
        view = flat[pos:pos + tr * pitch].reshape(tr, pitch)
        view[:, :tc] = t
        view[:, tc:] = -77  # neighbours that must not be touched or read
        tiles.append(t)
        refs.append((pos, pitch))
        pos += tr * pitch + 5
    return flat, tiles, refs


@pytest.mark.parametrize("shape", [(90, 120), (180, 240), (64, 50)])
def test_tile_list_host_matches_oracle(g4, oracle, shape):
    rng = np.random.default_rng(1)
    tr, tc = shape
    master = _master(g4)
    ids = [oracle.CODEC_HUFFMAN, oracle.CODEC_DEFLATE, oracle.CODEC_LSOP12]
    flat, tiles, refs = _scattered(oracle, rng, tr, tc, 7)
    batch = master.encodeTileList(flat, refs, tr, tc)
    for k, t in enumerate(tiles):
        assert batch.payload(k) == oracle.master_encode_i32(ids, t), "tile %d" % k
    # decode into ANOTHER scattering (reversed order, different pitches)
    refs2, pos = [], 11
    for k, (_off, pitch) in reversed(list(enumerate(refs))):
        refs2.append((pos, pitch + 2))
        pos += tr * (pitch + 2) + 3 * k + 1
    refs2 = refs2[::-1]   # tile k of the batch goes to refs2[k]: the tiles end up in reverse order in memory
    out = np.full(pos + 100, -5, np.int32)
    master.decodeTileList(batch.arena, batch.offsets, batch.lens, out, refs2, tr, tc)
    touched = np.zeros(out.size, bool)
    for k, (off, pitch) in enumerate(refs2):
        v = out[off:off + tr * pitch].reshape(tr, pitch)
        assert np.array_equal(v[:, :tc], tiles[k]), "tile %d" % k
        for r in range(tr):
            touched[off + r * pitch:off + r * pitch + tc] = True
    assert np.all(out[~touched] == -5), "cells outside the tiles were written"


def test_tile_list_device_matches_oracle(g4, oracle):
    import torch

    rng = np.random.default_rng(2)
    tr, tc = 180, 240
    master = _master(g4, ["LSOP12", "GvrsHuffman"])
    ids = [oracle.CODEC_LSOP12, oracle.CODEC_HUFFMAN]
    flat, tiles, refs = _scattered(oracle, rng, tr, tc, 9)
    # aligned placement as well (the vectorised LSOP12 path): tiles back to back in a tile cache's slot array
    slots = np.stack(tiles)
    for base, rf in ((flat, refs), (slots.reshape(-1), [(k * tr * tc, tc) for k in range(len(tiles))])):
        d = torch.from_numpy(base).cuda()
        batch = master.encodeTileList(d, rf, tr, tc)
        arena, off, ln = batch.arena.cpu().numpy(), batch.offsets.cpu().numpy(), batch.lens.cpu().numpy()
        for k, t in enumerate(tiles):
            assert arena[off[k]:off[k] + ln[k]].tobytes() == oracle.master_encode_i32(ids, t), "tile %d" % k
        out = torch.full_like(d, -9)
        master.decodeTileList(batch.arena, batch.offsets, batch.lens, out, rf, tr, tc)
        o = out.cpu().numpy()
        for k, (o0, pitch) in enumerate(rf):
            assert np.array_equal(o[o0:o0 + tr * pitch].reshape(tr, pitch)[:, :tc], tiles[k]), "tile %d" % k


def test_tile_list_float(g4, oracle):
    rng = np.random.default_rng(3)
    master = _master(g4, ["GvrsFloat"])
    flat, tiles, refs = _scattered(oracle, rng, 120, 120, 5, np.float32)
    batch = master.encodeTileList(flat, refs, 120, 120)
    for k, t in enumerate(tiles):
        assert batch.payload(k) == oracle.master_encode_f32([oracle.CODEC_FLOAT], t)
    out = np.zeros_like(flat)
    master.decodeTileList(batch.arena, batch.offsets, batch.lens, out, refs, 120, 120)
    for k, (off, pitch) in enumerate(refs):
        assert np.array_equal(out[off:off + 120 * pitch].reshape(120, pitch)[:, :120].view(np.uint32), tiles[k].view(np.uint32))


def test_tile_cache_read_block_and_flush(g4, oracle):
    """RasterTileCache: block reads decode every missing tile of the window in ONE call; writes mark tiles dirty and
    flush() encodes them in ONE call; what it encodes equals the oracle's CodecMaster packing of the modified tiles."""
    from gridfour_b200 import gvrs

    tr, tc = 90, 120
    grid = oracle.terrain_i32(3000, 5000, 6 * tr, 5 * tc)
    master = _master(g4)
    ctx = master._context()
    batch = master.encodeTiles(grid, tr, tc)
    spec = gvrs.GvrsSpec(grid.shape[0], grid.shape[1], tr, tc, [gvrs.ElementSpec.integer("z")], CODECS, checksum=True)
    w = gvrs.GvrsWriter(spec)
    keep = [t for t in range(30) if t != 11]           # tile 11 is absent from the file
    w.add_tile_records(ctx, batch.arena, np.asarray(batch.offsets)[keep], np.asarray(batch.lens)[keep], tile_index=keep)
    img = gvrs.GvrsImage.parse(w.finish(ctx))
    want = grid.copy()
    want[2 * tr:3 * tr, 1 * tc:2 * tc] = -2147483648
    cache = g4.RasterTileCache(img, master, max_tiles=8)
    # a block across six tiles: one batched decode
    blk = cache.readBlock(100, 130, 170, 300)
    assert np.array_equal(blk, want[100:270, 130:430]) and cache.decode_calls == 1
    # cells of resident tiles: no further decode
    assert cache.readValue(120, 140) == want[120, 140] and cache.decode_calls == 1
    # a block larger than the cache is read window by window
    assert np.array_equal(cache.readBlock(0, 0, 6 * tr, 5 * tc), want)
    assert cache.decode_calls <= 1 + 4
    # writes: three tiles become dirty, flush() = one batched encode whose payloads are the oracle's
    new = want.copy()
    new[85:95, 115:250] += 17
    cache.writeBlock(85, 115, new[85:95, 115:250])
    n_enc = cache.encode_calls
    tiles, b2 = cache.flush()
    assert cache.encode_calls == n_enc + 1 and cache.flush() is None
    assert tiles == [0, 1, 2, 5, 6, 7]
    ids = [oracle.CODEC_HUFFMAN, oracle.CODEC_DEFLATE, oracle.CODEC_LSOP12]
    for k, t in enumerate(tiles):
        r, c = divmod(t, 5)
        assert b2.payload(k) == oracle.master_encode_i32(ids, new[r * tr:(r + 1) * tr, c * tc:(c + 1) * tc]), "tile %d" % t
    assert np.array_equal(cache.readBlock(0, 0, 6 * tr, 5 * tc), new)


def test_tile_cache_on_reference_sample_files(g4, oracle):
    """The cache over the reference's own multi-tile sample files: every block read equals read_raster."""
    from gridfour_b200 import gvrs
    from gvrs_common import sample_files

    std = {"GvrsHuffman": (g4.CodecHuffman, g4.CodecHuffman), "GvrsDeflate": (g4.CodecDeflate, g4.CodecDeflate),
           "LSOP12": (g4.LsEncoder12, g4.LsDecoder12), "GvrsFloat": (g4.CodecFloat, g4.CodecFloat),
           "GvrsCanonicalHuffman": (g4.CodecCanonHuffman, g4.CodecCanonHuffman)}
    for name, data in sample_files().items():
        img = gvrs.GvrsImage.parse(data)
        spec = g4.CodecSpecification(default=False)
        for c in img.spec.codecs:
            spec.addCompressionCodec(c, *std[c])
        master = g4.CodecMaster(spec)
        for e, el in enumerate(img.spec.elements):
            if el.type_code == gvrs.ELEM_SHORT:
                continue
            full = img.read_raster(master, element=e, crop=True)
            cache = g4.RasterTileCache(img, master, element=e, max_tiles=4)
            nr, nc = full.shape
            got = cache.readBlock(0, 0, nr, nc)
            assert np.array_equal(got.view(np.uint32), full.view(np.uint32)), name
            r0, c0 = nr // 3, nc // 4
            assert np.array_equal(cache.readBlock(r0, c0, nr - r0, nc - c0).view(np.uint32), full[r0:, c0:].view(np.uint32)), name
