"""-m gpu: the file layer either side of the codec path through the C ABI:
g4_crc32c / g4_pack_tile_records / g4_unpack_tile_records + g4_decode_tiles with a file image as the arena.
Golden vectors: the reference's own sample files (tests/golden/gvrs_samples.json); oracle: oracle/g4oracle (CRC-32C,
codecs)."""
import struct

import numpy as np
import pytest

from gvrs_common import rebuild, sample_files, tile_record_parts

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g4():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gridfour_b200

    return gridfour_b200


def test_gpu_crc32c_matches_the_reference_files_and_the_oracle(g4, oracle):
    from gridfour_b200 import gvrs

    ctx = g4.Context.default()
    checked = 0
    for name, image in sample_files().items():
        img = gvrs.GvrsImage.parse(image)
        if img.spec.checksum:
            checked += img.verify_checksums(ctx)
    assert checked >= 120
    assert gvrs.crc32c(ctx, b"123456789") == 0xE3069283
    rng = np.random.default_rng(11)
    data = rng.integers(0, 256, 200000, dtype=np.uint8)
    off = rng.integers(0, 150000, 300).astype(np.uint64)
    size = rng.integers(0, 40000, 300).astype(np.uint32)
    size[:20] = np.arange(20)  # empty and tiny ranges, every alignment of the start
    off[:40] = np.arange(40)
    got = gvrs.crc32c_ranges(ctx, data, off, size)
    raw = data.tobytes()
    for o, s, c in zip(off, size, got):
        assert int(c) == oracle.crc32c(raw[int(o):int(o) + int(s)]), (int(o), int(s))


def test_gpu_rebuild_reproduces_every_reference_file_byte_for_byte(g4):
    """Header, metadata, directories from the model; tile records framed and every checksum computed on the GPU."""
    from gridfour_b200 import gvrs

    ctx = g4.Context.default()
    for name, image in sample_files().items():
        img = gvrs.GvrsImage.parse(image)
        again = rebuild(img, ctx, gpu_tiles=True)
        if img.spec.checksum:
            assert again == image, name
        else:
            assert len(again) == len(image), name


def _master_for(g4, codecs):
    spec = g4.CodecSpecification(default=False)
    classes = {"GvrsHuffman": (g4.CodecHuffman,), "GvrsDeflate": (g4.CodecDeflate,), "GvrsFloat": (g4.CodecFloat,),
               "GvrsCanonicalHuffman": (g4.CodecCanonHuffman,), "LSOP12": (g4.LsEncoder12, g4.LsDecoder12)}
    for c in codecs:
        spec.addCompressionCodec(c, *classes[c])
    return g4.CodecMaster(spec)


ORACLE_IDS = {"GvrsHuffman": 0, "GvrsDeflate": 1, "GvrsFloat": 2, "GvrsCanonicalHuffman": 3, "LSOP12": 4}


def test_gpu_decodes_reference_files_straight_from_the_image(g4, oracle):
    """Every sample file, every element: tile records located and checked on the GPU, tiles decoded with the image as the
    arena; compared with the oracle's decode of the same payloads (raw tiles: the little-endian samples themselves)."""
    from gridfour_b200 import gvrs

    n_files = 0
    for name, image in sample_files().items():
        img = gvrs.GvrsImage.parse(image)
        s = img.spec
        master = _master_for(g4, s.codecs)
        ids = [ORACLE_IDS[c] for c in s.codecs]
        n = s.tile_rows * s.tile_cols
        directory = img.tile_directory()
        for k, e in enumerate(s.elements):
          got = img.read_raster(master, element=k)
          assert got.shape == (s.tiles_down * s.tile_rows, s.tiles_across * s.tile_cols)
          for t in range(s.tiles_down * s.tiles_across):
            tr, tc = divmod(t, s.tiles_across)
            tile = got[tr * s.tile_rows:(tr + 1) * s.tile_rows, tc * s.tile_cols:(tc + 1) * s.tile_cols]
            if t not in directory:
                want = np.full((s.tile_rows, s.tile_cols), e.fill_value, dtype=tile.dtype)
            else:
                pos = directory[t] + 4
                for _ in range(k):
                    pos += 4 + struct.unpack_from("<i", image, pos)[0]
                ln = struct.unpack_from("<i", image, pos)[0]
                payload = image[pos + 4:pos + 4 + ln]
                if ln == e.standard_size(n):
                    want = np.frombuffer(payload[:n * tile.dtype.itemsize], dtype=tile.dtype).reshape(s.tile_rows, s.tile_cols)
                elif e.type_code == gvrs.ELEM_FLOAT:
                    want = oracle.master_decode_f32(ids, s.tile_rows, s.tile_cols, payload)
                else:
                    want = oracle.master_decode_i32(ids, s.tile_rows, s.tile_cols, payload)
                    if e.type_code == gvrs.ELEM_SHORT:
                        want = want.astype(np.int16)
            assert np.array_equal(tile.view(np.uint8), np.ascontiguousarray(want).view(np.uint8)), (name, t)
        n_files += 1
    assert n_files == 17


def test_gpu_written_file_round_trip(g4, oracle):
    """encodeTiles -> GPU-framed tile records -> file image -> parse -> checksum check -> decode from the image."""
    from gridfour_b200 import gvrs

    grid = oracle.terrain_i32(3000, 5000, 6 * 90, 5 * 120)
    codecs = ["GvrsHuffman", "GvrsDeflate", "LSOP12"]
    master = _master_for(g4, codecs)
    ctx = master._context()
    batch = master.encodeTiles(grid, 90, 120)
    spec = gvrs.GvrsSpec(grid.shape[0], grid.shape[1], 90, 120, [gvrs.ElementSpec.integer("z")], codecs, checksum=True)
    w = gvrs.GvrsWriter(spec, uuid=bytes(range(16)), time_modified=1700000000000)
    for name, rid, typ, content, desc in gvrs.codec_metadata(codecs):
        w.add_metadata(name, rid, typ, content, desc)
    w.add_tile_records(ctx, batch.arena, batch.offsets, batch.lens)
    image = w.finish(ctx)
    img = gvrs.GvrsImage.parse(image)
    assert img.verify_checksums(ctx) == len(img.records) + 1
    # the tile records are what the oracle's framing gives: [size][2,0,0,0][tileIndex][len][payload] pad [crc]
    d = img.tile_directory()
    assert sorted(d) == list(range(30))
    for t, pos in d.items():
        size = struct.unpack_from("<i", image, pos - 8)[0]
        ti, ln = struct.unpack_from("<ii", image, pos)
        assert ti == t and ln == int(batch.lens[t]) and size == ((ln + 20 + 7) & ~7)
        assert image[pos + 8:pos + 8 + ln] == batch.payload(t)
        assert image[pos - 4:pos] == bytes([2, 0, 0, 0])
        assert struct.unpack_from("<I", image, pos - 8 + size - 4)[0] == oracle.crc32c(image[pos - 8:pos - 8 + size - 4])
    assert np.array_equal(img.read_raster(master), grid)
    # a flipped payload bit is caught by the record checksum
    bad = bytearray(image)
    bad[d[7] + 40] ^= 0x10
    with pytest.raises(IOError):
        gvrs.GvrsImage.parse(bytes(bad)).read_raster(master)
    # a tile that the directory does not list reads as fill values
    w2 = gvrs.GvrsWriter(spec)
    keep = [t for t in range(30) if t != 11]
    w2.add_tile_records(ctx, batch.arena, np.asarray(batch.offsets)[keep], np.asarray(batch.lens)[keep], tile_index=keep)
    out = gvrs.GvrsImage.parse(w2.finish(ctx)).read_raster(master)
    want = grid.copy()
    want[2 * 90:3 * 90, 1 * 120:2 * 120] = -2147483648
    assert np.array_equal(out, want)


def test_gpu_record_errors_are_reported_per_tile(g4):
    from gridfour_b200 import gvrs

    ctx = g4.Context.default()
    payload = bytes(range(40))
    recs, pos = gvrs.pack_tile_records(ctx, payload + bytes(8), [0], [40], 64, True)
    image = bytearray(64) + recs
    off, lens, status = gvrs.unpack_tile_records(ctx, bytes(image), pos, True)
    assert int(status[0]) == 0 and int(lens[0]) == 40 and bytes(image[int(off[0]):int(off[0]) + 40]) == payload
    for mutate in (lambda b: b.__setitem__(64 + 4, 1),                      # not a tile record
                   lambda b: struct.pack_into("<i", b, 64, 4096),           # size beyond the image
                   lambda b: struct.pack_into("<i", b, 64 + 12, 1 << 20),   # payload length beyond the record
                   lambda b: b.__setitem__(64 + 30, b[64 + 30] ^ 1)):       # checksum
        b = bytearray(image)
        mutate(b)
        _, _, st = gvrs.unpack_tile_records(ctx, bytes(b), pos, True)
        assert int(st[0]) == -2
    _, _, st = gvrs.unpack_tile_records(ctx, bytes(image), np.array([0, 3, 1 << 40], dtype=np.uint64), True)
    assert [int(x) for x in st] == [1, -2, -2]


def test_partial_edge_tiles_round_trip(g4, oracle):
    """A raster whose dimensions are not multiples of the tile size: the edge tiles are full tiles whose outside cells
    hold the fill value (here the integer null code, so the edge tiles go through the nulls predictor); every packing is
    the oracle's, and the cropped read is the input."""
    from gridfour_b200 import gvrs

    INT_NULL = -2147483648
    codecs = ["GvrsHuffman", "GvrsDeflate", "LSOP12"]
    master = _master_for(g4, codecs)
    ids = [ORACLE_IDS[c] for c in codecs]
    for fill in (INT_NULL, -9999):
        grid = oracle.terrain_i32(100, 200, 500, 700)
        spec = gvrs.GvrsSpec(500, 700, 90, 120, [gvrs.ElementSpec.integer("z", fill_value=fill)], codecs, checksum=True)
        assert (spec.tiles_down, spec.tiles_across) == (6, 6)
        w = gvrs.GvrsWriter(spec, uuid=bytes(16), time_modified=1)
        batch = w.add_raster(master, grid)
        image = w.finish(master._context())
        full = np.full((540, 720), fill, dtype=np.int32)
        full[:500, :700] = grid
        n_null_pred = 0
        for t in range(36):
            tr, tc = divmod(t, 6)
            tile = np.ascontiguousarray(full[tr * 90:(tr + 1) * 90, tc * 120:(tc + 1) * 120])
            want = oracle.master_encode_i32(ids, tile)
            got = batch.payload(t)
            assert got == want, t              # (the oracle returns the raw tile when nothing compresses)
            if len(got) < 4 * 90 * 120:
                n_null_pred += got[1] == 4
        if fill == INT_NULL:
            assert n_null_pred >= 11          # the 11 edge tiles can only be predicted by the nulls model
        img = gvrs.GvrsImage.parse(image)
        assert img.verify_checksums(master._context()) == len(img.records) + 1
        assert np.array_equal(img.read_raster(master, crop=True), grid)
        assert np.array_equal(img.read_raster(master), full)
