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


def test_corrupt_tile_directory_is_a_format_error(oracle):
    """g4_decode_tiles_bounded: a directory entry that points past the arena, or a payload longer than the raw tile, is
    reported per tile (FormatError) and never dereferenced; the other tiles of the batch still decode."""
    import torch

    import gridfour_b200 as g4

    grid = oracle.terrain_i32(0, 0, 2 * 90, 3 * 120)
    spec = g4.CodecSpecification(default=False)
    spec.addCompressionCodec("GvrsHuffman", g4.CodecHuffman)
    spec.addCompressionCodec("LSOP12", g4.LsEncoder12, g4.LsDecoder12)
    master = g4.CodecMaster(spec)
    for device in (False, True):
        batch = master.encodeTiles(torch.from_numpy(grid).cuda() if device else grid, 90, 120)
        offsets, lens = batch.offsets.clone() if device else batch.offsets.copy(), batch.lens.clone() if device else batch.lens.copy()
        offsets[1] = int(batch.arena.numel() if device else batch.arena.size) + (1 << 40)  # far outside
        lens[4] = 4 * 90 * 120 + 8                                                       # longer than a raw tile
        bad = g4.TileBatch(batch.arena, offsets, lens, None, None, None, batch.total_bytes, batch.band)
        with pytest.raises(g4.FormatError):
            master.decodeTiles(bad)
        st = master.lastStatus.cpu().numpy() if device else master.lastStatus
        assert st[1] != 0 and st[4] != 0 and not st[[0, 2, 3, 5]].any()
        out = master.decodeTiles(batch)
        assert np.array_equal(out.cpu().numpy() if device else out, grid)


def test_device_tensors_from_pending_torch_work_are_ordered(oracle):
    """The context launches on its own stream: a CUDA tensor still being produced on torch's current stream must be
    complete before the codec kernels read it (g4_context_order_stream), and torch must see the finished result."""
    import torch

    import gridfour_b200 as g4

    base = torch.from_numpy(oracle.terrain_i32(0, 0, 4 * 180, 8 * 240)).cuda()
    spec = g4.CodecSpecification(default=False)
    spec.addCompressionCodec("LSOP12", g4.LsEncoder12, g4.LsDecoder12)
    master = g4.CodecMaster(spec)
    ref = master.encodeTiles(base + 7, 180, 240)
    torch.cuda.synchronize()
    want_lens = ref.lens.clone()
    side = torch.cuda.Stream()
    for _ in range(5):
        with torch.cuda.stream(side):
            big = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
            for _k in range(20):
                big.normal_()            # keeps `side` busy for a while
            grid = base + 7              # queued behind the busy work, on `side`
            batch = master.encodeTiles(grid, 180, 240)   # must wait for `grid`
            out = master.decodeTiles(batch)
            ok = torch.equal(out, grid)  # queued on `side` again: must see the decoder's output
        assert torch.equal(batch.lens, want_lens)
        assert bool(ok)


def test_contexts_on_two_devices_in_one_process(oracle):
    """Function attributes (dynamic shared memory opt-in) are per device: a second GPU in the same process must work."""
    import torch

    import gridfour_b200 as g4

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    tile = oracle.terrain_i32(0, 0, 180, 240)
    want = oracle.lsop12_encode(0, tile)
    for dev in (0, 1, 0):
        ctx = g4.Context(dev)
        assert g4.LsEncoder12(ctx).encode(0, 180, 240, tile) == want
        assert np.array_equal(g4.LsDecoder12(ctx).decode(180, 240, want), tile)
        assert np.array_equal(g4.CodecHuffman(ctx).decode(180, 240, g4.CodecHuffman(ctx).encode(0, 180, 240, tile)), tile)
        ctx.close()


def test_pipelined_device_calls(oracle):
    """g4_context_set_async: device batches are only enqueued; the per-tile status (device memory) is valid after the
    context's stream has run, a malformed tile shows there instead of in the return code."""
    import torch

    import gridfour_b200 as g4

    tr, tc = 60, 80
    grid = torch.from_numpy(oracle.terrain_i32(300, 700, 2 * tr, 3 * tc)).cuda()
    spec = g4.CodecSpecification(default=False)
    spec.addCompressionCodec("GvrsHuffman", g4.CodecHuffman)
    spec.addCompressionCodec("LSOP12", g4.LsEncoder12, g4.LsDecoder12)
    master = g4.CodecMaster(spec)
    batch = master.encodeTiles(grid, tr, tc)
    ctx = master._context()
    out = torch.zeros_like(grid)
    ctx.set_async(True)
    try:
        for _ in range(3):  # three batches in flight on the context's stream
            master.decodeTiles(batch, out=out)
        ctx.synchronize()
        torch.cuda.synchronize()
        assert int((master.lastStatus != 0).sum()) == 0 and torch.equal(out, grid)
        # a corrupted tile: the call itself still returns (nothing was awaited), the status tensor carries the error
        bad = batch.arena.clone()
        o = int(batch.offsets[4])
        bad[o + 1] = 77  # unknown predictor / header byte
        b2 = g4.TileBatch(bad, batch.offsets, batch.lens, batch.codec, batch.predictor, batch.status, batch.total_bytes, batch.band)
        master.decodeTiles(b2, out=out)
        ctx.synchronize()
        torch.cuda.synchronize()
        st = master.lastStatus.cpu().numpy()
        assert st[4] < 0 and (np.delete(st, 4) == 0).all(), st
    finally:
        ctx.set_async(False)
    with pytest.raises(IOError):
        master.decodeTiles(b2, out=out)  # awaited again: the first failing tile's status is the call's result
