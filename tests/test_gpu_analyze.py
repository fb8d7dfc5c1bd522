"""-m gpu: ICompressionDecoder.analyze / reportAnalysisData (compress/ICompressionDecoder.java:79-92) for the M32-based
codecs.  GPU statistics (g4_analyze_tiles) against a numpy restatement of CodecHuffman.analyze, CodecDeflate.analyze and
CodecStats (oracle.analyze_m32): integer quantities exactly, the FP64 entropy within 1e-12 relative (Math.log and the
device's log may differ in the last place)."""
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g4():
    import gridfour_b200

    return gridfour_b200


def _tiles(oracle):
    rng = np.random.default_rng(4)
    t = [oracle.terrain_i32(300 * k, 500 * k, 90, 120) for k in range(5)]
    t.append(np.full((90, 120), 7, np.int32))                                   # one distinct symbol
    t.append((np.arange(90 * 120, dtype=np.int32).reshape(90, 120) * 3))         # ramp
    t.append(rng.integers(-40000, 40000, (90, 120)).astype(np.int32))           # multi-byte M32 codes
    return t


@pytest.mark.parametrize("name", ["CodecHuffman", "CodecDeflate"])
def test_analyze_matches_codec_stats(g4, oracle, name):
    oid = oracle.CODEC_HUFFMAN if name == "CodecHuffman" else oracle.CODEC_DEFLATE
    codec = getattr(g4, name)()
    want_rows = {}
    pairs = np.zeros((5, 65536), np.uint64)
    n = 0
    for k, tile in enumerate(_tiles(oracle)):
        packing, _ = oracle.codec_encode_i32(oid, 1, tile)
        if packing is None:
            continue
        codec.analyze(90, 120, packing)
        pred, n_bytes, n_sym, overhead, n_m32, observed, entropy, pr = oracle.analyze_m32(oid, 90, 120, packing)
        row = want_rows.setdefault(pred, [0, 0, 0, 0, 0, 0, 0.0, 0])
        for j, v in enumerate((1, n_bytes, n_sym, overhead, n_m32, observed, entropy, 1)):
            row[j] += v
        pairs[pred] += pr
        n += 1
    assert n >= 7
    for pred, row in want_rows.items():
        st = codec.codecStats[pred]
        assert (st.nTilesCounted, st.nBytesTotal, st.nSymbolsTotal, st.nBitsOverheadTotal, st.sumLengthM32, st.sumObservedM32) == tuple(row[:6])
        assert st.nM32Counted == row[7]
        assert abs(st.sumEntropyM32 - row[6]) <= 1e-12 * max(1.0, abs(row[6]))
        assert np.array_equal(st.sB, pairs[pred])
    total = codec.codecStats[-1]
    assert total.nTilesCounted == n and np.array_equal(total.sB, pairs.sum(axis=0))
    assert total.getH2() > 0.0
    out = io.StringIO()
    codec.reportAnalysisData(out, 2 * n)
    text = out.getvalue().splitlines()
    assert text[0].startswith("Gridfour_Huffman" if name == "CodecHuffman" else "Gridfour_Deflate")
    assert ("bits in tree" in text[1]) == (name == "CodecHuffman")
    assert any(line.strip().startswith("All Predictors") and "(50.0 %)" in line for line in text)
    codec.clearAnalysisData()
    out = io.StringIO()
    codec.reportAnalysisData(out, 10)
    assert "Tiles Compressed:  0" in out.getvalue()


def test_analyze_tiles_batch(g4, oracle):
    """A best-of batch (Huffman + Deflate + LSOP12): per-tile records for the M32 codecs, G4_DECLINED for the others."""
    grid = oracle.terrain_i32(0, 0, 2 * 90, 4 * 120).copy()
    grid[:90, :120] = np.random.default_rng(0).integers(-2 ** 31, 2 ** 31 - 1, (90, 120), dtype=np.int64).astype(np.int32)  # -> raw
    spec = g4.CodecSpecification(default=False)
    spec.addCompressionCodec("GvrsHuffman", g4.CodecHuffman)
    spec.addCompressionCodec("GvrsDeflate", g4.CodecDeflate)
    master = g4.CodecMaster(spec)
    batch = master.encodeTiles(grid, 90, 120)
    ts = g4.analyze_tiles(master._context(), spec.native_list(), batch.band, batch.arena, batch.offsets, batch.lens)
    assert ts[0].status == 1 and ts[0].codec_kind == -1          # the raw tile
    ids = [oracle.CODEC_HUFFMAN, oracle.CODEC_DEFLATE]
    for t in range(1, 8):
        p = batch.payload(t)
        want = oracle.analyze_m32(ids[p[0]], 90, 120, p)
        got = ts[t]
        assert got.status == 0 and got.codec_kind == ids[p[0]]
        assert (got.predictor, got.n_bytes, got.n_symbols, got.n_bits_overhead, got.n_m32, got.observed) == want[:6]
        assert abs(got.entropy - want[6]) <= 1e-12 * max(1.0, abs(want[6]))
