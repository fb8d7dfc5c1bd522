// g4o_lsop08.cpp -- CPU oracle (TEST INFRASTRUCTURE, never linked into the product): the 8-coefficient Lewis-Smith
// optimal predictor, the reference's legacy "LSOP08" codec.  Line-faithful restatement of
//   lsop/LsOptimalPredictor08.java:58-245   (initializers, FP64 normal equations, float32 stencil, (int)(p + 0.5f))
//   lsop/LsEncoder08.java:66-137            (header, Deflate(6) of both M32 streams, legacy Huffman alternative)
//   lsop/LsDecoder08.java:65-163            (decode, unpackInitializers, unpackInterior)
//   lsop/LsHeader.java:104-265, util/jama/LUDecomposition.java:70-135,253-286
// under /root/reference/core/src/main/java/org/gridfour/.  The reference does not register this codec any more
// (lsop/LsCodecUtility.java:73 is commented out) and holds no fixture for it: PARITY UNPINNED by the reference; the
// restatement is pinned only by its own round trip and by sharing M32 / Huffman / zlib / LU code paths with the pinned
// LSOP12 restatement.
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <vector>
#include "g4oracle.h"

namespace g4o {

namespace {

// JAMA LUDecomposition + solve for one right-hand side, n = 9 (same operation order as the 13 x 13 form in g4o_lsop.cpp)
bool lu_solve9(double LU[9][9], double X[9]) {
  const int n = 9;
  int piv[9];
  for (int i = 0; i < n; i++) piv[i] = i;
  double col[9];
  for (int j = 0; j < n; j++) {
    for (int i = 0; i < n; i++) col[i] = LU[i][j];
    for (int i = 0; i < n; i++) {
      int kmax = i < j ? i : j;
      double s = 0.0;
      for (int k = 0; k < kmax; k++) s += LU[i][k] * col[k];
      LU[i][j] = col[i] -= s;
    }
    int p = j;
    for (int i = j + 1; i < n; i++)
      if (std::fabs(col[i]) > std::fabs(col[p])) p = i;
    if (p != j) {
      for (int k = 0; k < n; k++) { double t = LU[p][k]; LU[p][k] = LU[j][k]; LU[j][k] = t; }
      int k = piv[p]; piv[p] = piv[j]; piv[j] = k;
    }
    if (LU[j][j] != 0.0)
      for (int i = j + 1; i < n; i++) LU[i][j] /= LU[j][j];
  }
  for (int j = 0; j < n; j++) if (LU[j][j] == 0) return false;  // lud.solve throws "Matrix is singular."
  double B[9];
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

// (int)(p + 0.5f): float addition, then Java's narrowing conversion (toward zero, saturating, NaN -> 0)
inline int32_t java_f2i(float a) {
  if (a != a) return 0;
  if (a >= 2147483648.0f) return INT32_MAX;
  if (a <= -2147483648.0f) return INT32_MIN;
  return int32_t(a);
}

inline float stencil8(const float* u, const int32_t* v, int index, int nCols) {  // LsOptimalPredictor08.java:140-148
  float p = u[0] * float(v[index - 1])
          + u[1] * float(v[index - nCols - 1])
          + u[2] * float(v[index - nCols])
          + u[3] * float(v[index - 2])
          + u[4] * float(v[index - nCols - 2])
          + u[5] * float(v[index - 2 * nCols - 2])
          + u[6] * float(v[index - 2 * nCols - 1])
          + u[7] * float(v[index - 2 * nCols]);
  return p;
}

void put32(std::vector<uint8_t>& b, uint32_t v) { for (int i = 0; i < 4; i++) b.push_back(uint8_t(v >> (8 * i))); }
uint32_t get32(const uint8_t* p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24); }

}  // namespace

// LsOptimalPredictor08.computeCoefficients (:183-244).  A singular matrix makes the reference THROW (no catch in
// LsOptimalPredictor08.encode); the oracle reports it like the null return.
bool lsop08_coefficients(int nRows, int nCols, const int32_t* values, double ud[8]) {
  if (nRows < 4 || nCols < 4) return false;
  double z[9], s[9] = {0};
  double c[9][9] = {{0}};
  for (int r = 2; r < nRows; r++) {
    for (int col = 2; col < nCols; col++) {
      int index = r * nCols + col;
      z[0] = values[index];
      z[1] = values[index - 1];
      z[2] = values[index - nCols - 1];
      z[3] = values[index - nCols];
      z[4] = values[index - 2];
      z[5] = values[index - nCols - 2];
      z[6] = values[index - 2 * nCols - 2];
      z[7] = values[index - 2 * nCols - 1];
      z[8] = values[index - 2 * nCols];
      for (int i = 0; i < 9; i++) s[i] += z[i];
      for (int i = 0; i < 9; i++)
        for (int j = i; j < 9; j++) c[i][j] += z[i] * z[j];
    }
  }
  for (int i = 1; i < 9; i++)
    for (int j = 0; j < i; j++) c[i][j] = c[j][i];
  double m[9][9] = {{0}};
  for (int i = 1; i < 9; i++) {
    for (int j = 1; j < 9; j++) m[i - 1][j - 1] = c[i][j];
    m[i - 1][8] = s[i];
  }
  for (int j = 1; j < 9; j++) m[8][j - 1] = s[j];
  double b[9];
  for (int i = 1; i < 9; i++) b[i - 1] = c[0][i];
  b[8] = s[0];
  if (!lu_solve9(m, b)) return false;
  for (int i = 0; i < 8; i++) ud[i] = b[i];
  return true;
}

// LsEncoder08.encode (:66-137) over LsOptimalPredictor08.encode (:58-181)
bool codec_lsop08_encode(int codecIndex, int nRows, int nCols, const int32_t* v, std::vector<uint8_t>& out) {
  if (nRows < 4 || nCols < 4) return false;
  std::vector<uint8_t> initCodes(size_t(nCols + nRows) * 2 * M32_MAX_BYTES_PER_VALUE);
  M32Writer wi(initCodes.data());
  const int32_t seed = v[0];
  int64_t prior = seed;
  for (int i = 1; i < nCols; i++) { int64_t t = v[i]; wi.encode(int32_t(t - prior)); prior = t; }
  prior = v[0];
  for (int i = 0; i < nCols; i++) { int64_t t = v[i + nCols]; wi.encode(int32_t(t - prior)); prior = t; }
  for (int r = 2; r < nRows; r++) {
    int index = r * nCols;
    prior = v[index - nCols];
    for (int i = 0; i < 2; i++) { int64_t t = v[index++]; wi.encode(int32_t(t - prior)); prior = t; }
  }
  initCodes.resize(wi.off);
  double ud[8];
  if (!lsop08_coefficients(nRows, nCols, v, ud)) return false;
  float u[8];
  for (int i = 0; i < 8; i++) u[i] = float(ud[i]);
  std::vector<uint8_t> interiorCodes(size_t(nRows - 2) * size_t(nCols - 2) * M32_MAX_BYTES_PER_VALUE);
  M32Writer wn(interiorCodes.data());
  for (int r = 2; r < nRows; r++)
    for (int c = 2; c < nCols; c++) {
      int index = r * nCols + c;
      float p = stencil8(u, v, index, nCols);
      wn.encode(int32_t(uint32_t(v[index]) - uint32_t(java_f2i(p + 0.5f))));
    }
  interiorCodes.resize(wn.off);
  auto header = [&](int type) {  // LsHeader.packHeader (:210-265), revised layout
    std::vector<uint8_t> h;
    h.push_back(uint8_t(codecIndex));
    h.push_back(uint8_t(type | 0x40));
    h.push_back(8);
    put32(h, uint32_t(seed));
    for (int i = 0; i < 8; i++) { uint32_t bits; std::memcpy(&bits, &u[i], 4); put32(h, bits); }
    put32(h, uint32_t(initCodes.size()));
    put32(h, uint32_t(interiorCodes.size()));
    return h;
  };
  std::vector<uint8_t> initPack(initCodes.size() + 128), insidePack(interiorCodes.size() + 128);
  int initN = zlib_deflate_capped(6, initCodes.data(), initCodes.size(), initPack.data(), initPack.size());
  if (initN <= 0) return false;
  int insideN = zlib_deflate_capped(6, interiorCodes.data(), interiorCodes.size(), insidePack.data(), insidePack.size());
  if (insideN <= 0) return false;
  out = header(1);
  out.insert(out.end(), initPack.begin(), initPack.begin() + initN);
  out.insert(out.end(), insidePack.begin(), insidePack.begin() + insideN);
  BitOut store;
  huffman_encode(store, int(initCodes.size()), initCodes.data());
  huffman_encode(store, int(interiorCodes.size()), interiorCodes.data());
  std::vector<uint8_t> huff = store.text();
  if (int(huff.size()) < initN + insideN) {
    out = header(0);
    out[out.size() - 1] = 0;  // sic: LsEncoder08.java:122 clears the last header byte (the top byte of nInteriorCodes)
    out.insert(out.end(), huff.begin(), huff.end());
  }
  return true;
}

// LsDecoder08.decode (:65-113), unpackInitializers (:115-135), unpackInterior (:137-163)
void codec_lsop08_decode(int nRows, int nCols, const uint8_t* packing, size_t len, int32_t* values) {
  if (len < 3) throw std::runtime_error("LSOP packing too short");
  size_t off = 1;
  auto need = [&](size_t k) { if (off + k > len) throw std::runtime_error("LSOP header truncated"); };
  int type = 0, nCoef;
  bool cks = false;
  int32_t seed;
  float u[8] = {0};
  uint32_t nInitCodes = 0, nInteriorCodes = 0;
  const bool legacy = (packing[1] & 0x40) == 0;
  if (!legacy) { type = packing[off] & 0x0f; cks = (packing[off] & 0x80) != 0; off++; }
  need(1);
  nCoef = int8_t(packing[off++]);
  if (nCoef != 8) throw std::runtime_error("LSOP08 decoder given a header that does not carry 8 coefficients");
  need(4); seed = int32_t(get32(packing + off)); off += 4;
  for (int i = 0; i < 8; i++) { need(4); uint32_t bits = get32(packing + off); off += 4; std::memcpy(&u[i], &bits, 4); }
  if (legacy || type != 2) {
    need(8);
    nInitCodes = get32(packing + off); off += 4;
    nInteriorCodes = get32(packing + off); off += 4;
  }
  if (legacy) { need(1); type = packing[off] & 0x0f; cks = (packing[off] & 0x80) != 0; off++; }
  if (cks) { need(4); off += 4; }
  const size_t headerSize = off;
  if (nRows < 2 || nCols < 2) throw std::runtime_error("LSOP tile too small");
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
  M32Reader ri(initCodes.data(), initCodes.size());
  values[0] = seed;
  int32_t vv = seed;
  for (int i = 1; i < nCols; i++) { vv = int32_t(uint32_t(vv) + uint32_t(ri.decode())); values[i] = vv; }
  vv = seed;
  for (int i = 0; i < nCols; i++) { vv = int32_t(uint32_t(vv) + uint32_t(ri.decode())); values[nCols + i] = vv; }
  for (int r = 2; r < nRows; r++) {
    int o = r * nCols;
    values[o] = int32_t(uint32_t(values[o - nCols]) + uint32_t(ri.decode()));
    values[o + 1] = int32_t(uint32_t(values[o]) + uint32_t(ri.decode()));
  }
  M32Reader rn(interiorCodes.data(), interiorCodes.size());
  for (int r = 2; r < nRows; r++)
    for (int c = 2; c < nCols; c++) {
      int index = r * nCols + c;
      float p = stencil8(u, values, index, nCols);
      values[index] = int32_t(uint32_t(java_f2i(p + 0.5f)) + uint32_t(rn.decode()));
    }
}

}  // namespace g4o
