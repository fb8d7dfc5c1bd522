// CPU ORACLE (test infrastructure only) -- CodecM32 and the four predictor models.
// Restates C/compress/CodecM32.java and C/compress/PredictorModel{Differencing,Linear,Triangle,
// DifferencingWithNulls}.java   (C/ = /root/reference/core/src/main/java/org/gridfour/).
// Java `int` wrap-around arithmetic is expressed with uint32_t; `long` with int64_t.
#include "g4oracle.h"
#include <cmath>

namespace g4o {

static inline int32_t wsub(int32_t a, int32_t b) { return int32_t(uint32_t(a) - uint32_t(b)); }
static inline int32_t wadd(int32_t a, int32_t b) { return int32_t(uint32_t(a) + uint32_t(b)); }

// CodecM32.java:257-311
void M32Writer::encode(int32_t value) {
  const int loMask = 0x7f, hiBit = 0x80;
  int32_t absValue;
  if (value < 0) {
    if (value == INT32_MIN) { buf[off++] = 0x80; return; }
    if (value > -127) { buf[off++] = uint8_t(value); return; }
    buf[off++] = uint8_t(-127);
    absValue = -value;
  } else {
    if (value < 127) { buf[off++] = uint8_t(value); return; }
    buf[off++] = 127;
    absValue = value;
  }
  if (absValue <= 254) {
    buf[off++] = uint8_t(absValue - 127);
  } else if (absValue <= 16638) {
    int d = absValue - 255;
    buf[off++] = uint8_t(((d >> 7) & loMask) | hiBit);
    buf[off++] = uint8_t(d & loMask);
  } else if (absValue <= 2113790) {
    int d = absValue - 16639;
    buf[off++] = uint8_t(((d >> 14) & loMask) | hiBit);
    buf[off++] = uint8_t(((d >> 7) & loMask) | hiBit);
    buf[off++] = uint8_t(d & loMask);
  } else if (absValue <= 270549246) {
    int d = absValue - 2113791;
    buf[off++] = uint8_t(((d >> 21) & loMask) | hiBit);
    buf[off++] = uint8_t(((d >> 14) & loMask) | hiBit);
    buf[off++] = uint8_t(((d >> 7) & loMask) | hiBit);
    buf[off++] = uint8_t(d & loMask);
  } else {
    int d = absValue - 270549247;
    buf[off++] = uint8_t(((d >> 28) & loMask) | hiBit);
    buf[off++] = uint8_t(((d >> 21) & loMask) | hiBit);
    buf[off++] = uint8_t(((d >> 14) & loMask) | hiBit);
    buf[off++] = uint8_t(((d >> 7) & loMask) | hiBit);
    buf[off++] = uint8_t(d & loMask);
  }
}

// CodecM32.java:313-356.  The reference does no bounds checking; the oracle throws instead of
// reading past `limit` so malformed inputs are detectable in tests.
int32_t M32Reader::decode() {
  static const int32_t segmentBaseValue[5] = {127, 255, 16639, 2113791, 270549247};
  if (off >= limit) throw std::out_of_range("M32 read past end");
  int symbol = int8_t(buf[off++]);
  if (symbol == -128) return INT32_MIN;
  if (-127 < symbol && symbol < 127) return symbol;
  uint32_t delta = 0;
  for (int i = 0; i < 5; i++) {
    if (off >= limit) throw std::out_of_range("M32 read past end");
    int sample = int8_t(buf[off++]);
    delta = (delta << 7) | uint32_t(sample & 0x7f);
    if ((sample & 0x80) == 0) {
      if (symbol == -127) delta = uint32_t(0) - delta - uint32_t(segmentBaseValue[i]);
      else delta += uint32_t(segmentBaseValue[i]);
      break;
    }
  }
  return int32_t(delta);
}

// A sink abstraction lets one traversal serve both encode() (M32 bytes) and encodeInt() (ints),
// which are written out twice, line for line, in the reference.
struct Sink {
  M32Writer* m32 = nullptr;
  int32_t* ints = nullptr;
  int k = 0;
  void put(int32_t r) {
    if (m32) m32->encode(r);
    if (ints) ints[k] = r;
    k++;
  }
  int result() const { return m32 ? int(m32->off) : k; }
};
struct Source {
  M32Reader* m32 = nullptr;
  const int32_t* ints = nullptr;
  size_t n = 0, k = 0;
  int32_t get() {
    if (m32) return m32->decode();
    if (k >= n) throw std::out_of_range("residual read past end");
    return ints[k++];
  }
};

// PredictorModelDifferencing.java:112-142 (encode), :170-200 (encodeInt)
static int enc_differencing(int nRows, int nCols, const int32_t* v, Sink& s, int32_t* seed) {
  *seed = v[0];
  int32_t prior = v[0];
  for (int i = 1; i < nCols; i++) { int32_t t = v[i]; s.put(wsub(t, prior)); prior = t; }
  for (int r = 1; r < nRows; r++) {
    int idx = r * nCols;
    prior = v[idx - nCols];
    for (int i = 0; i < nCols; i++) { int32_t t = v[idx++]; s.put(wsub(t, prior)); prior = t; }
  }
  return s.result();
}
// PredictorModelDifferencing.java:145-167, :203-225
static void dec_differencing(int32_t seed, int nRows, int nCols, Source& s, int32_t* out) {
  out[0] = seed;
  int32_t prior = seed;
  for (int i = 1; i < nCols; i++) { prior = wadd(prior, s.get()); out[i] = prior; }
  for (int r = 1; r < nRows; r++) {
    int idx = r * nCols;
    prior = out[idx - nCols];
    for (int c = 0; c < nCols; c++) { prior = wadd(prior, s.get()); out[idx++] = prior; }
  }
}

// PredictorModelLinear.java:104-143, :146-185.  No guard for nCols<2 in the reference (reads values[1]).
static int enc_linear(int nRows, int nCols, const int32_t* v, Sink& s, int32_t* seed) {
  *seed = v[0];
  int64_t prior = v[0];
  int64_t delta = int64_t(v[1]) - prior;
  s.put(int32_t(delta));
  for (int r = 1; r < nRows; r++) {
    int idx = r * nCols;
    int64_t test = v[idx];
    delta = test - prior;
    s.put(int32_t(delta));
    prior = test;
    test = v[idx + 1];
    delta = test - prior;
    s.put(int32_t(delta));
  }
  for (int r = 0; r < nRows; r++) {
    int idx = r * nCols;
    int64_t a = v[idx], b = v[idx + 1];
    for (int c = 2; c < nCols; c++) {
      int32_t cv = v[idx + c];
      int32_t prediction = int32_t(2 * b - a);
      s.put(wsub(cv, prediction));
      a = b;
      b = cv;
    }
  }
  return s.result();
}
// PredictorModelLinear.java:66-101, :188-223
static void dec_linear(int32_t seed, int nRows, int nCols, Source& s, int32_t* out) {
  int64_t prior = seed;
  out[0] = seed;
  out[1] = int32_t(int64_t(s.get()) + prior);
  for (int r = 1; r < nRows; r++) {
    int idx = r * nCols;
    int64_t test = int64_t(s.get()) + prior;
    out[idx] = int32_t(test);
    prior = test;
    out[idx + 1] = int32_t(int64_t(s.get()) + test);
  }
  for (int r = 0; r < nRows; r++) {
    int idx = r * nCols;
    int64_t a = out[idx], b = out[idx + 1];
    for (int c = 2; c < nCols; c++) {
      int32_t residual = s.get();
      int32_t prediction = int32_t(2 * b - a);
      int32_t cv = wadd(prediction, residual);
      a = b;
      b = cv;
      out[idx + c] = cv;
    }
  }
}

// PredictorModelTriangle.java:101-145, :148-186
static int enc_triangle(int nRows, int nCols, const int32_t* v, Sink& s, int32_t* seed) {
  if (nRows < 2 || nCols < 2) return -1;
  *seed = v[0];
  int64_t prior = v[0];
  for (int i = 1; i < nCols; i++) { int64_t t = v[i]; s.put(int32_t(t - prior)); prior = t; }
  prior = v[0];
  for (int i = 1; i < nRows; i++) { int64_t t = v[i * nCols]; s.put(int32_t(t - prior)); prior = t; }
  for (int r = 1; r < nRows; r++) {
    int k1 = r * nCols, k0 = k1 - nCols;
    for (int i = 1; i < nCols; i++) {
      int64_t za = v[k0++], zb = v[k1++], zc = v[k0];
      int32_t prediction = int32_t(zc + zb - za);
      s.put(wsub(v[k1], prediction));
    }
  }
  return s.result();
}
// PredictorModelTriangle.java:62-98, :189-217
static void dec_triangle(int32_t seed, int nRows, int nCols, Source& s, int32_t* out) {
  out[0] = seed;
  int32_t prior = seed;
  for (int i = 1; i < nCols; i++) { prior = wadd(prior, s.get()); out[i] = prior; }
  prior = seed;
  for (int i = 1; i < nRows; i++) { prior = wadd(prior, s.get()); out[i * nCols] = prior; }
  for (int r = 1; r < nRows; r++) {
    int k1 = r * nCols, k0 = k1 - nCols;
    for (int i = 1; i < nCols; i++) {
      int64_t za = out[k0++], zb = out[k1++], zc = out[k0];
      int32_t prediction = int32_t(zb + zc - za);
      out[k1] = wadd(prediction, s.get());
    }
  }
}

// PredictorModelDifferencingWithNulls.java:66-134, :169-237
static int enc_diff_nulls(int nRows, int nCols, const int32_t* v, Sink& s, int32_t* seed) {
  int64_t sumStart = 0;
  int nStart = 0;
  bool nullFlag = true;
  for (int r = 0; r < nRows; r++) {
    int ro = r * nCols;
    for (int c = 0; c < nCols; c++) {
      int32_t t = v[ro + c];
      if (t == INT4_NULL_CODE) nullFlag = true;
      else { if (nullFlag) { sumStart += t; nStart++; } nullFlag = false; }
    }
    nullFlag = v[ro] == INT4_NULL_CODE;
  }
  if (nStart == 0) return 0;
  double avgStart = double(sumStart) / nStart;
  int32_t encodedSeed = int32_t(std::floor(avgStart + 0.5));
  *seed = encodedSeed;
  int64_t prior = encodedSeed;
  nullFlag = false;
  for (int r = 0; r < nRows; r++) {
    int idx = r * nCols;
    for (int c = 0; c < nCols; c++) {
      int32_t t = v[idx++];
      if (t == INT4_NULL_CODE) { nullFlag = true; s.put(INT4_NULL_CODE); }
      else {
        if (nullFlag) { prior = encodedSeed; nullFlag = false; }
        int64_t delta = int64_t(t) - prior;
        s.put(int32_t(delta));
        prior = t;
      }
    }
    prior = v[r * nCols];
    nullFlag = prior == INT4_NULL_CODE;
  }
  return s.result();
}
// PredictorModelDifferencingWithNulls.java:137-166, :240-269
static void dec_diff_nulls(int32_t seed, int nRows, int nCols, Source& s, int32_t* out) {
  int32_t prior = seed;
  bool nullFlag = true;
  for (int r = 0; r < nRows; r++) {
    int idx = r * nCols;
    for (int c = 0; c < nCols; c++) {
      int32_t t = s.get();
      if (t == INT4_NULL_CODE) { nullFlag = true; out[idx++] = INT4_NULL_CODE; }
      else {
        if (nullFlag) { nullFlag = false; prior = seed; }
        prior = wadd(prior, t);
        out[idx++] = prior;
      }
    }
    prior = out[r * nCols];
    nullFlag = prior == INT4_NULL_CODE;
  }
}

static int enc_dispatch(int model, int nRows, int nCols, const int32_t* v, Sink& s, int32_t* seed) {
  switch (model) {
    case PRED_DIFFERENCING: return enc_differencing(nRows, nCols, v, s, seed);
    case PRED_LINEAR: return enc_linear(nRows, nCols, v, s, seed);
    case PRED_TRIANGLE: return enc_triangle(nRows, nCols, v, s, seed);
    case PRED_DIFF_NULLS: return enc_diff_nulls(nRows, nCols, v, s, seed);
    default: throw std::runtime_error("Unknown PredictorCorrector type");
  }
}
static void dec_dispatch(int model, int32_t seed, int nRows, int nCols, Source& s, int32_t* out) {
  switch (model) {
    case PRED_DIFFERENCING: dec_differencing(seed, nRows, nCols, s, out); break;
    case PRED_LINEAR: dec_linear(seed, nRows, nCols, s, out); break;
    case PRED_TRIANGLE: dec_triangle(seed, nRows, nCols, s, out); break;
    case PRED_DIFF_NULLS: dec_diff_nulls(seed, nRows, nCols, s, out); break;
    default: throw std::runtime_error("Unknown PredictorCorrector type");  // CodecHuffman.java:166-167
  }
}

int predictor_encode_int(int model, int nRows, int nCols, const int32_t* values, int32_t* out, int32_t* seed) {
  Sink s; s.ints = out;
  return enc_dispatch(model, nRows, nCols, values, s, seed);
}
int predictor_encode(int model, int nRows, int nCols, const int32_t* values, uint8_t* out, int32_t* seed) {
  M32Writer w(out);
  Sink s; s.m32 = &w;
  return enc_dispatch(model, nRows, nCols, values, s, seed);
}
void predictor_decode_int(int model, int32_t seed, int nRows, int nCols, const int32_t* enc, size_t nEnc, int32_t* out) {
  Source s; s.ints = enc; s.n = nEnc;
  dec_dispatch(model, seed, nRows, nCols, s, out);
}
void predictor_decode(int model, int32_t seed, int nRows, int nCols, const uint8_t* enc, size_t nEnc, int32_t* out) {
  M32Reader r(enc, nEnc);
  Source s; s.m32 = &r;
  dec_dispatch(model, seed, nRows, nCols, s, out);
}

}  // namespace g4o
