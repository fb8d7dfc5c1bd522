"""Host side of ICompressionDecoder.analyze / reportAnalysisData: compress/CodecStats.java:42-300 as a mirror that is FED by
the GPU (g4_analyze_tiles: per-tile byte counts, M32 length, distinct symbols, first-order entropy, successor-pair counts),
plus the report tables of CodecHuffman.reportAnalysisData (:202-234) and CodecDeflate.reportAnalysisData (:231-260).
Paths under /root/reference/core/src/main/java/org/gridfour/compress/.  Nothing here touches sample data: the sums below
are the reference's bookkeeping (one addition per tile, in tile order)."""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import G4_MEM_HOST, G4_OK, BandDesc, CodecList, check

PREDICTOR_NAMES = ["None", "Differencing", "Linear", "Triangle", "DifferencingWithNulls"]  # PredictorModelType.values()


class TileStats(C.Structure):
    """g4_tile_stats"""
    _fields_ = [("codec_kind", C.c_int32), ("predictor", C.c_int32), ("n_bytes", C.c_uint32), ("n_symbols", C.c_uint32),
                ("n_bits_overhead", C.c_uint32), ("n_m32", C.c_uint32), ("observed", C.c_uint32), ("status", C.c_int32),
                ("entropy", C.c_double)]


class CodecStats:
    """compress/CodecStats.java: the same fields, the same accessors."""

    def __init__(self, name):
        self.name = name
        self.nTilesCounted = self.nBytesTotal = self.nSymbolsTotal = self.nBitsOverheadTotal = 0
        self.nM32Counted = self.sumLengthM32 = self.sumObservedM32 = 0
        self.sumEntropyM32 = 0.0
        self.sB = np.zeros(65536, np.uint64)  # successor pairs (prior << 8 | value); sA[value] = column sum

    def getLabel(self):
        return self.name

    def addToCounts(self, nBytesForTile, nSymbolsInTile, nBitsOverhead):
        self.nTilesCounted += 1
        self.nBytesTotal += int(nBytesForTile)
        self.nSymbolsTotal += int(nSymbolsInTile)
        self.nBitsOverheadTotal += int(nBitsOverhead)

    def addTile(self, ts):
        """CodecHuffman.analyze / CodecDeflate.analyze for one tile, from the GPU's record (g4_tile_stats)."""
        self.addToCounts(ts.n_bytes, ts.n_symbols, ts.n_bits_overhead)
        if ts.n_m32 > 0:  # addCountsForM32 (:100-131)
            self.nM32Counted += 1
            self.sumLengthM32 += int(ts.n_m32)
            self.sumObservedM32 += int(ts.observed)
            self.sumEntropyM32 += float(ts.entropy)

    def getH2(self):  # :133-165
        sB = self.sB.reshape(256, 256).astype(np.float64)
        sA = sB.sum(axis=0)
        k = sA.sum()
        if k == 0:
            return 0.0
        h2 = 0.0
        for i in range(256):
            if sA[i] > 0:
                row = sB[i]
                n = row.sum()
                pj = row[row > 0] / n
                h2 += (sA[i] / k) * float((pj * np.log(pj)).sum())
        return -h2

    def getEntropy(self):
        return self.sumEntropyM32 / self.nM32Counted if self.nM32Counted else 0.0

    def clear(self):
        self.nTilesCounted = self.nBytesTotal = self.nSymbolsTotal = self.nBitsOverheadTotal = 0

    def getBitsPerSymbol(self):
        return 8.0 * self.nBytesTotal / self.nSymbolsTotal if self.nSymbolsTotal else 0.0

    def getTileCount(self):
        return self.nTilesCounted

    def getAverageMCodeLength(self):
        return self.sumLengthM32 / self.nM32Counted if self.nM32Counted else 0.0

    def getAverageObservedMCodes(self):
        return self.sumObservedM32 / self.nTilesCounted if self.nTilesCounted else 0.0

    def getAverageOverhead(self):
        return self.nBitsOverheadTotal / self.nTilesCounted if self.nTilesCounted else 0.0

    def getAverageLength(self):
        return self.nBytesTotal / self.nTilesCounted if self.nTilesCounted else 0.0


def analyze_tiles(context, codec_list, band, arena, offsets, lens, pairs=None):
    """g4_analyze_tiles over host buffers -> the list of g4_tile_stats records (one per tile)."""
    a = np.frombuffer(arena, dtype=np.uint8) if not isinstance(arena, np.ndarray) else np.ascontiguousarray(arena)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    ln = np.ascontiguousarray(lens, dtype=np.uint32)
    n = len(off)
    out = (TileStats * n)()
    check(_lib.lib().g4_analyze_tiles(context._h, C.byref(codec_list), C.byref(band), G4_MEM_HOST, a.ctypes.data, int(a.size), off.ctypes.data,
                                      ln.ctypes.data, out, None if pairs is None else pairs.ctypes.data), "g4_analyze_tiles")
    return out


class AnalysisMixin:
    """analyze / reportAnalysisData / clearAnalysisData of the M32-based codecs (CodecHuffman, CodecDeflate)."""

    _report_title = ""
    _with_tree = False

    def _stats(self):
        if getattr(self, "codecStats", None) is None:
            self.codecStats = [CodecStats(n) for n in PREDICTOR_NAMES] + [CodecStats("All Predictors")]
            self._pairs = np.zeros((2, 5, 65536), np.uint64)
        return self.codecStats

    def analyze(self, nRows, nColumns, packing):
        from .codecs import CodecMaster

        stats = self._stats()
        cl = CodecList()
        cl.n_codecs = int(packing[0]) + 1
        for k in range(cl.n_codecs):
            cl.codec_ids[k] = self.codec_id
        band = CodecMaster._band((nRows, nColumns), np.int32, nRows, nColumns)
        b = np.frombuffer(bytes(packing) + bytes(16), dtype=np.uint8)
        ts = analyze_tiles(self._context(), cl, band, b, [0], [len(packing)], self._pairs)[0]
        if ts.status < 0:
            check(ts.status, "g4_analyze_tiles")
        if ts.status != G4_OK:
            return
        stats[ts.predictor].addTile(ts)
        stats[-1].addTile(ts)
        kind = 1 if self.codec_id == _lib.G4_CODEC_DEFLATE else 0
        for p in range(5):
            stats[p].sB = self._pairs[kind, p]
        stats[-1].sB = self._pairs[kind].sum(axis=0)

    def reportAnalysisData(self, ps, nTilesInRaster):
        ps.write("%s                               Compressed Output    |       Predictor Residuals\n" % self._report_title)
        if getattr(self, "codecStats", None) is None or nTilesInRaster == 0:
            ps.write("   Tiles Compressed:  0\n")
            return
        ps.write("  Predictor                Times Used        bits/sym    bits/tile  |  m32 avg-len   avg-unique  entropy%s\n"
                 % (" | bits in tree" if self._with_tree else ""))
        for st in self.codecStats:
            if st.getLabel().lower() == "none":
                continue
            line = "   %-20.20s %8d (%4.1f %%)     %5.2f  %12.1f   | %10.1f      %6.1f    %6.2f" % (
                st.getLabel(), st.getTileCount(), 100.0 * st.getTileCount() / nTilesInRaster, st.getBitsPerSymbol(),
                st.getAverageLength() * 8, st.getAverageMCodeLength(), st.getAverageObservedMCodes(), st.getEntropy())
            if self._with_tree:
                line += "   | %6.1f" % st.getAverageOverhead()
            ps.write(line + "\n")

    def clearAnalysisData(self):
        self.codecStats = None
