"""Host-side mirror of the predictor models (compress/IPredictorModel.java:42-173 and its four implementations under
/root/reference/core/src/main/java/org/gridfour/compress/): same method names, argument meaning and return values.
Every call runs the CUDA kernels of csrc/g4_predictor.cu through the C ABI (g4_predictor_encode / _decode / _encode_int /
_decode_int); nothing is computed in Python.

Like the reference's models these objects are stateful (encode stores the seed that getSeed() returns,
PredictorModelDifferencing.java:104) and therefore not shareable between threads."""
import ctypes as C
from enum import Enum

import numpy as np

from . import _lib
from ._lib import G4_DECLINED, G4_OK, check
from .codecs import Context


class PredictorModelType(Enum):
    """compress/PredictorModelType.java:42-103: the code a packing stores for its predictor."""
    NONE = 0
    Differencing = 1
    Linear = 2
    Triangle = 3
    DifferencingWithNulls = 4

    def getCodeValue(self):
        return self.value

    @staticmethod
    def valueOf(code):
        try:
            return PredictorModelType(int(code))
        except ValueError:
            return PredictorModelType.NONE  # PredictorModelType.java:99-100


class IPredictorModel:
    _type = PredictorModelType.NONE

    def __init__(self, context=None):
        self._ctx = context
        self.encodedSeed = 0

    def _context(self):
        return self._ctx or Context.default()

    def getPredictorType(self):
        return self._type

    def isNullDataSupported(self):
        return self._type is PredictorModelType.DifferencingWithNulls

    def getSeed(self):
        return self.encodedSeed

    # int encode(int nRows, int nColumns, int[] values, byte[] encoding) -> number of M32 bytes, -1 on failure
    def encode(self, nRows, nColumns, values, encoding):
        v = np.ascontiguousarray(values, dtype=np.int32).reshape(nRows, nColumns)
        out = np.frombuffer(encoding, dtype=np.uint8) if not isinstance(encoding, np.ndarray) else encoding
        seed, n = C.c_int32(0), C.c_size_t(0)
        st = _lib.lib().g4_predictor_encode(self._context()._h, self._type.value, nRows, nColumns, v.ctypes.data, C.byref(seed),
                                            out.ctypes.data, out.size, C.byref(n))
        if st == G4_DECLINED:
            return -1
        check(st, "g4_predictor_encode")
        self.encodedSeed = seed.value
        return int(n.value)

    # void decode(int seed, int nRows, int nColumns, byte[] encoding, int offset, int length, int[] output)
    def decode(self, seed, nRows, nColumns, encoding, offset, length, output):
        b = np.frombuffer(bytes(encoding[offset:offset + length]), dtype=np.uint8)
        out = np.asarray(output)
        assert out.dtype == np.int32 and out.size >= nRows * nColumns and out.flags["C_CONTIGUOUS"]
        check(_lib.lib().g4_predictor_decode(self._context()._h, self._type.value, int(seed), nRows, nColumns, b.ctypes.data, b.size,
                                             out.ctypes.data), "g4_predictor_decode")

    # int encodeInt(int nRows, int nColumns, int[] values, int[] encoding) -> number of residuals
    def encodeInt(self, nRows, nColumns, values, encoding):
        v = np.ascontiguousarray(values, dtype=np.int32).reshape(nRows, nColumns)
        out = np.asarray(encoding)
        assert out.dtype == np.int32 and out.flags["C_CONTIGUOUS"]
        seed, n = C.c_int32(0), C.c_size_t(0)
        st = _lib.lib().g4_predictor_encode_int(self._context()._h, self._type.value, nRows, nColumns, v.ctypes.data, C.byref(seed),
                                                out.ctypes.data, out.size, C.byref(n))
        if st == G4_DECLINED:
            return -1
        check(st, "g4_predictor_encode_int")
        self.encodedSeed = seed.value
        return int(n.value)

    # void decodeInt(int seed, int nRows, int nColumns, int[] encoding, int offset, int length, int[] output)
    def decodeInt(self, seed, nRows, nColumns, encoding, offset, length, output):
        # (the reference's Linear and Triangle models ignore `offset`, PredictorModelLinear.java:196 / Triangle :195;
        # callers only ever pass 0)
        r = np.ascontiguousarray(np.asarray(encoding, dtype=np.int32)[offset:offset + length])
        out = np.asarray(output)
        assert out.dtype == np.int32 and out.size >= nRows * nColumns and out.flags["C_CONTIGUOUS"]
        check(_lib.lib().g4_predictor_decode_int(self._context()._h, self._type.value, int(seed), nRows, nColumns, r.ctypes.data, r.size,
                                                 out.ctypes.data), "g4_predictor_decode_int")


class PredictorModelDifferencing(IPredictorModel):
    _type = PredictorModelType.Differencing


class PredictorModelLinear(IPredictorModel):
    _type = PredictorModelType.Linear


class PredictorModelTriangle(IPredictorModel):
    _type = PredictorModelType.Triangle


class PredictorModelDifferencingWithNulls(IPredictorModel):
    _type = PredictorModelType.DifferencingWithNulls
