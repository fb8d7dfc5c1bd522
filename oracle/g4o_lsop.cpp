// CPU ORACLE (test infrastructure only) -- LSOP12 (Lewis-Smith optimal predictor, 12 coefficients).
// Restates C/lsop/{LsEncoder12,LsOptimalPredictor12,LsDecoder12,LsHeader}.java and
// C/util/jama/LUDecomposition.java:70-135,253-286  (C/ = /root/reference/core/src/main/java/org/gridfour/).
// MUST be compiled with -ffp-contract=off (Java never fuses multiply-add).
#include "g4oracle.h"
#include <cmath>

namespace g4o {

// StrictMath.round(float) as specified since JDK 7u/8 (java.lang.Math.round(float)):
// floor(a + 1/2) computed on the bit pattern, NaN -> 0, saturating.
int32_t java_round_float(float a) {
  int32_t intBits;
  std::memcpy(&intBits, &a, 4);
  int biasedExp = (intBits & 0x7F800000) >> 23;
  int shift = (24 - 2 + 127) - biasedExp;
  if ((shift & -32) == 0) {
    int32_t r = (intBits & 0x007FFFFF) | 0x00800000;
    if (intBits < 0) r = -r;
    return ((r >> shift) + 1) >> 1;
  }
  // (int) a : NaN -> 0, saturate
  if (a != a) return 0;
  if (a >= 2147483648.0f) return INT32_MAX;
  if (a <= -2147483648.0f) return INT32_MIN;
  return int32_t(a);
}

uint32_t crc32c(const uint8_t* p, size_t n) {  // C/util/GridfourCRC32C.java:160-167 (Castagnoli, reflected)
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0x82f63b78u : c >> 1;
      table[i] = c;
    }
    init = true;
  }
  uint32_t crc = 0xffffffffu;
  for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
  return crc ^ 0xffffffffu;
}

// JAMA LUDecomposition ctor (:70-135) + solve for one right-hand side (:253-286), n = 13.
static bool lu_solve13(double LU[13][13], double X[13]) {
  const int n = 13, m = 13;
  int piv[13];
  for (int i = 0; i < m; i++) piv[i] = i;
  double LUcolj[13];
  for (int j = 0; j < n; j++) {
    for (int i = 0; i < m; i++) LUcolj[i] = LU[i][j];
    for (int i = 0; i < m; i++) {
      double* LUrowi = LU[i];
      int kmax = i < j ? i : j;
      double s = 0.0;
      for (int k = 0; k < kmax; k++) s += LUrowi[k] * LUcolj[k];
      LUrowi[j] = LUcolj[i] -= s;
    }
    int p = j;
    for (int i = j + 1; i < m; i++)
      if (std::fabs(LUcolj[i]) > std::fabs(LUcolj[p])) p = i;
    if (p != j) {
      for (int k = 0; k < n; k++) { double t = LU[p][k]; LU[p][k] = LU[j][k]; LU[j][k] = t; }
      int k = piv[p]; piv[p] = piv[j]; piv[j] = k;
    }
    if (LU[j][j] != 0.0)
      for (int i = j + 1; i < m; i++) LU[i][j] /= LU[j][j];
  }
  for (int j = 0; j < n; j++) if (LU[j][j] == 0) return false;  // isNonsingular :148-154
  double B[13];
  for (int i = 0; i < n; i++) B[i] = X[piv[i]];
  for (int i = 0; i < n; i++) X[i] = B[i];
  for (int k = 0; k < n; k++)
    for (int i = k + 1; i < n; i++) X[i] -= X[k] * LU[i][k];
  for (int k = n - 1; k >= 0; k--) {
    X[k] /= LU[k][k];
    for (int i = 0; i < k; i++) X[i] -= X[k] * LU[i][k];
  }
  return true;
}

// LsOptimalPredictor12.computeCoefficients (:311-383)
bool lsop12_coefficients(int nRows, int nCols, const int32_t* values, double ud[12]) {
  if (nRows < 6 || nCols < 6) return false;
  double z[13], s[13] = {0};
  static thread_local double c[13][13];
  for (int i = 0; i < 13; i++) for (int j = 0; j < 13; j++) c[i][j] = 0;
  for (int r = 2; r < nRows; r++) {
    for (int col = 2; col < nCols - 2; col++) {
      int index = r * nCols + col;
      z[0] = values[index];
      z[1] = values[index - 1];
      z[2] = values[index - nCols - 1];
      z[3] = values[index - nCols];
      z[4] = values[index - nCols + 1];
      z[5] = values[index - nCols + 2];
      z[6] = values[index - 2];
      z[7] = values[index - nCols - 2];
      z[8] = values[index - 2 * nCols - 2];
      z[9] = values[index - 2 * nCols - 1];
      z[10] = values[index - 2 * nCols];
      z[11] = values[index - 2 * nCols + 1];
      z[12] = values[index - 2 * nCols + 2];
      for (int i = 0; i < 13; i++) s[i] += z[i];
      for (int i = 0; i < 13; i++)
        for (int j = i; j < 13; j++) c[i][j] += z[i] * z[j];
    }
  }
  for (int i = 1; i < 13; i++) for (int j = 0; j < i; j++) c[i][j] = c[j][i];
  double m[13][13];
  for (int i = 0; i < 13; i++) for (int j = 0; j < 13; j++) m[i][j] = 0;
  for (int i = 1; i < 13; i++) {
    for (int j = 1; j < 13; j++) m[i - 1][j - 1] = c[i][j];
    m[i - 1][12] = s[i];
  }
  for (int j = 1; j < 13; j++) m[12][j - 1] = s[j];
  double b[13];
  for (int i = 1; i < 13; i++) b[i - 1] = c[0][i];
  b[12] = s[0];
  if (!lu_solve13(m, b)) return false;
  for (int i = 0; i < 12; i++) ud[i] = b[i];
  return true;
}

namespace {
struct LsResult {
  int32_t seed;
  float u[12];
  std::vector<int32_t> initInt, interiorInt;
  std::vector<uint8_t> initCodes, interiorCodes;
};

// the 12-tap float32 stencil, evaluated strictly left to right (LsOptimalPredictor12.java:256-267)
inline float stencil(const float* u, const int32_t* v, int index, int nCols) {
  float p = u[0] * float(v[index - 1])
          + u[1] * float(v[index - nCols - 1])
          + u[2] * float(v[index - nCols])
          + u[3] * float(v[index - nCols + 1])
          + u[4] * float(v[index - nCols + 2])
          + u[5] * float(v[index - 2])
          + u[6] * float(v[index - nCols - 2])
          + u[7] * float(v[index - 2 * nCols - 2])
          + u[8] * float(v[index - 2 * nCols - 1])
          + u[9] * float(v[index - 2 * nCols])
          + u[10] * float(v[index - 2 * nCols + 1])
          + u[11] * float(v[index - 2 * nCols + 2]);
  return p;
}

// LsOptimalPredictor12.encode (:109-292)
bool ls_predict(int nRows, int nCols, const int32_t* v, LsResult& R) {
  if (nRows < 6 || nCols < 6) return false;
  int n = nRows * 4 + nCols * 2 - 9;
  R.initInt.clear(); R.initInt.reserve(n);
  R.seed = v[0];
  int64_t prior = v[0];
  for (int i = 1; i < nCols; i++) { int64_t t = v[i]; R.initInt.push_back(int32_t(t - prior)); prior = t; }
  prior = v[0];
  for (int i = 1; i < nRows; i++) { int64_t t = v[i * nCols]; R.initInt.push_back(int32_t(t - prior)); prior = t; }
  for (int i = 1; i < nCols; i++) {
    int idx = nCols + i;
    R.initInt.push_back(int32_t(int64_t(v[idx]) - ((int64_t(v[idx - 1]) + v[idx - nCols]) - v[idx - nCols - 1])));
  }
  for (int i = 2; i < nRows; i++) {
    int idx = i * nCols + 1;
    R.initInt.push_back(int32_t(int64_t(v[idx]) - ((int64_t(v[idx - 1]) + v[idx - nCols]) - v[idx - nCols - 1])));
  }
  for (int i = 2; i < nRows; i++) {
    int idx = i * nCols + nCols - 2;
    R.initInt.push_back(int32_t(int64_t(v[idx]) - ((int64_t(v[idx - 1]) + v[idx - nCols]) - v[idx - nCols - 1])));
    idx++;
    R.initInt.push_back(int32_t(int64_t(v[idx]) - ((int64_t(v[idx - 1]) + v[idx - nCols]) - v[idx - nCols - 1])));
  }
  double ud[12];
  if (!lsop12_coefficients(nRows, nCols, v, ud)) return false;
  for (int i = 0; i < 12; i++) R.u[i] = float(ud[i]);
  int nInt = (nRows - 2) * (nCols - 4);
  R.interiorInt.clear(); R.interiorInt.reserve(nInt);
  for (int r = 2; r < nRows; r++) {
    for (int c = 2; c < nCols - 2; c++) {
      int index = r * nCols + c;
      float p = stencil(R.u, v, index, nCols);
      int32_t estimate = java_round_float(p);
      R.interiorInt.push_back(int32_t(uint32_t(v[index]) - uint32_t(estimate)));
    }
  }
  R.initCodes.resize(R.initInt.size() * M32_MAX_BYTES_PER_VALUE);
  { M32Writer w(R.initCodes.data()); for (int32_t x : R.initInt) w.encode(x); R.initCodes.resize(w.off); }
  R.interiorCodes.resize(R.interiorInt.size() * M32_MAX_BYTES_PER_VALUE);
  { M32Writer w(R.interiorCodes.data()); for (int32_t x : R.interiorInt) w.encode(x); R.interiorCodes.resize(w.off); }
  return true;
}

void put_i32(std::vector<uint8_t>& b, uint32_t v) { for (int i = 0; i < 4; i++) b.push_back(uint8_t(v >> (8 * i))); }
uint32_t get_u32(const uint8_t* p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24); }

// LsHeader.packHeader (LsHeader.java:210-265)
std::vector<uint8_t> pack_header(int codecIndex, const LsResult& R, int type, bool cks, uint32_t checksum) {
  std::vector<uint8_t> h;
  h.push_back(uint8_t(codecIndex));
  h.push_back(uint8_t(type | 0x40 | (cks ? 0x80 : 0)));
  h.push_back(12);
  put_i32(h, uint32_t(R.seed));
  for (int i = 0; i < 12; i++) { uint32_t bits; std::memcpy(&bits, &R.u[i], 4); put_i32(h, bits); }
  if (type != 2) { put_i32(h, uint32_t(R.initCodes.size())); put_i32(h, uint32_t(R.interiorCodes.size())); }
  if (cks) put_i32(h, checksum);
  return h;
}

uint32_t value_checksum(int nRows, int nCols, const int32_t* v) {  // LsHeader.java:391-406
  size_t n = size_t(nRows) * nCols;
  std::vector<uint8_t> b(n * 4);
  for (size_t i = 0; i < n; i++) for (int k = 0; k < 4; k++) b[4 * i + k] = uint8_t(uint32_t(v[i]) >> (8 * k));
  return crc32c(b.data(), b.size());
}
}  // namespace

// LsEncoder12.encode (LsEncoder12.java:122-219)
bool codec_lsop12_encode(int codecIndex, int nRows, int nCols, const int32_t* v, std::vector<uint8_t>& out,
                         bool deflateEnabled, bool checksumEnabled, EncodeInfo* info) {
  LsResult R;
  if (!ls_predict(nRows, nCols, v, R)) return false;
  uint32_t checksum = checksumEnabled ? value_checksum(nRows, nCols, v) : 0;
  std::vector<uint8_t> header = pack_header(codecIndex, R, 2, checksumEnabled, checksum);
  BitOut canonStore;
  canon_encode(canonStore, int(R.initInt.size()), R.initInt.data());
  canon_encode(canonStore, int(R.interiorInt.size()), R.interiorInt.data());
  std::vector<uint8_t> canon = canonStore.text();
  size_t canonLength = canon.size();
  out = header;
  out.insert(out.end(), canon.begin(), canon.end());
  if (info) { info->predictor = 2; info->seed = R.seed; }
  if (!deflateEnabled) return true;
  std::vector<uint8_t> insidePack(R.interiorCodes.size() + 128);
  int insideN = zlib_deflate_capped(6, R.interiorCodes.data(), R.interiorCodes.size(), insidePack.data(), insidePack.size());
  if (insideN <= 0 || size_t(insideN) >= canonLength) return true;
  std::vector<uint8_t> initPack(R.initCodes.size() + 128);
  int initN = zlib_deflate_capped(6, R.initCodes.data(), R.initCodes.size(), initPack.data(), initPack.size());
  if (initN <= 0 || size_t(initN) + size_t(insideN) >= canonLength) return true;
  header = pack_header(codecIndex, R, 1, checksumEnabled, checksum);
  out = header;
  out.insert(out.end(), initPack.begin(), initPack.begin() + initN);
  out.insert(out.end(), insidePack.begin(), insidePack.begin() + insideN);
  if (info) info->predictor = 1;
  return true;
}

// LsDecoder12.decode (LsDecoder12.java:94-160) with LsHeader ctor (LsHeader.java:104-189)
void codec_lsop12_decode(int nRows, int nCols, const uint8_t* packing, size_t len, int32_t* values) {
  if (len < 3) throw std::runtime_error("LSOP packing too short");
  size_t off = 1;
  int nCoef, type;
  int32_t seed;
  float u[12] = {0};
  uint32_t nInitCodes = 0, nInteriorCodes = 0;
  bool cks = false;
  auto need = [&](size_t k) { if (off + k > len) throw std::runtime_error("LSOP header truncated"); };
  auto read_u = [&](int n) {
    for (int i = 0; i < n; i++) {
      need(4);
      uint32_t bits = get_u32(packing + off); off += 4;
      if (i < 12) std::memcpy(&u[i], &bits, 4);
    }
  };
  if ((packing[1] & 0x40) == 0) {  // legacy layout
    nCoef = int8_t(packing[off++]);
    need(4); seed = int32_t(get_u32(packing + off)); off += 4;
    if (nCoef != 12) throw std::runtime_error("LSOP12 decoder given a non-12-coefficient header");
    read_u(nCoef);
    need(9);
    nInitCodes = get_u32(packing + off); off += 4;
    nInteriorCodes = get_u32(packing + off); off += 4;
    type = packing[off] & 0x0f;
    cks = (packing[off] & 0x80) != 0;
    off++;
    if (cks) { need(4); off += 4; }
  } else {
    type = packing[off] & 0x0f;
    cks = (packing[off] & 0x80) != 0;
    off++;
    nCoef = int8_t(packing[off++]);
    need(4); seed = int32_t(get_u32(packing + off)); off += 4;
    if (nCoef != 12) throw std::runtime_error("LSOP12 decoder given a non-12-coefficient header");
    read_u(nCoef);
    if (type != 2) {
      need(8);
      nInitCodes = get_u32(packing + off); off += 4;
      nInteriorCodes = get_u32(packing + off); off += 4;
    }
    if (cks) { need(4); off += 4; }
  }
  const size_t headerSize = off;
  if (nRows < 6 || nCols < 6) throw std::runtime_error("LSOP tile too small");
  const int nInit = nRows * 4 + nCols * 2 - 9;
  const int nInterior = (nRows - 2) * (nCols - 4);
  std::vector<int32_t> initInt(nInit, 0), interiorInt(nInterior, 0);
  if (type == 2) {
    BitIn in(packing + headerSize, len - headerSize);
    canon_decode(in, nInit, initInt.data());
    canon_decode(in, nInterior, interiorInt.data());
  } else {
    std::vector<uint8_t> initCodes(nInitCodes), interiorCodes(nInteriorCodes);
    if (type == 0) {
      BitIn in(packing + headerSize, len - headerSize);
      huffman_decode(in, int(nInitCodes), initCodes.data());
      huffman_decode(in, int(nInteriorCodes), interiorCodes.data());
    } else {
      size_t consumed = 0;
      int t = zlib_inflate(packing + headerSize, len - headerSize, initCodes.data(), initCodes.size(), &consumed);
      if (t < 0) throw std::runtime_error("zlib data error");
      if (uint32_t(t) < nInitCodes) throw std::runtime_error("Format mismatch, unable to read initializer codes");
      size_t o2 = headerSize + consumed;
      t = zlib_inflate(packing + o2, len - o2, interiorCodes.data(), interiorCodes.size(), &consumed);
      if (t < 0) throw std::runtime_error("zlib data error");
      if (uint32_t(t) < nInteriorCodes) throw std::runtime_error("Format mismatch, unable to read interior codes");
    }
    // the M32 flavour interleaves reads from the initializer stream; decode both to ints first, which
    // yields the identical value sequence (each stream is consumed strictly in order).
    M32Reader ri(initCodes.data(), initCodes.size());
    for (int i = 0; i < nInit; i++) initInt[i] = ri.decode();
    M32Reader rn(interiorCodes.data(), interiorCodes.size());
    for (int i = 0; i < nInterior; i++) interiorInt[i] = rn.decode();
  }
  // unpackInitializers (:204-241)
  int k = 0;
  values[0] = seed;
  int32_t vv = seed;
  for (int i = 1; i < nCols; i++) { vv = int32_t(uint32_t(vv) + uint32_t(initInt[k++])); values[i] = vv; }
  vv = seed;
  for (int i = 1; i < nRows; i++) { vv = int32_t(uint32_t(vv) + uint32_t(initInt[k++])); values[i * nCols] = vv; }
  for (int i = 1; i < nCols; i++) {
    int idx = nCols + i;
    values[idx] = int32_t(int64_t(initInt[k++]) + ((int64_t(values[idx - 1]) + values[idx - nCols]) - values[idx - nCols - 1]));
  }
  for (int i = 2; i < nRows; i++) {
    int idx = i * nCols + 1;
    values[idx] = int32_t(int64_t(initInt[k++]) + ((int64_t(values[idx - 1]) + values[idx - nCols]) - values[idx - nCols - 1]));
  }
  // unpackInterior (:353-470)
  int ki = 0;
  for (int r = 2; r < nRows; r++) {
    for (int c = 2; c < nCols - 2; c++) {
      int index = r * nCols + c;
      float p = stencil(u, values, index, nCols);
      int32_t estimate = java_round_float(p);
      values[index] = int32_t(uint32_t(estimate) + uint32_t(interiorInt[ki++]));
    }
    int idx = r * nCols + nCols - 2;
    values[idx] = int32_t(int64_t(initInt[k++]) + ((int64_t(values[idx - 1]) + values[idx - nCols]) - values[idx - nCols - 1]));
    idx++;
    values[idx] = int32_t(int64_t(initInt[k++]) + ((int64_t(values[idx - 1]) + values[idx - nCols]) - values[idx - nCols - 1]));
  }
}

}  // namespace g4o

// Test infrastructure: the predictor half of LsOptimalPredictor12.encode (:109-292) on its own -- seed, the twelve float32
// coefficients and the two M32-coded residual streams (LsOptimalPredictorResult.java:44-76).  Returns false where encode
// returns null.  Buffers must hold 6 bytes per residual.
namespace g4o {
bool lsop12_residual_streams(int nRows, int nCols, const int32_t* v, int32_t* seed, float u[12], uint8_t* initCodes, long* nInit,
                             uint8_t* interiorCodes, long* nInterior) {
  LsResult R;
  if (!ls_predict(nRows, nCols, v, R)) return false;
  *seed = R.seed;
  for (int i = 0; i < 12; i++) u[i] = R.u[i];
  std::copy(R.initCodes.begin(), R.initCodes.end(), initCodes);
  std::copy(R.interiorCodes.begin(), R.interiorCodes.end(), interiorCodes);
  *nInit = long(R.initCodes.size());
  *nInterior = long(R.interiorCodes.size());
  return true;
}
}  // namespace g4o
