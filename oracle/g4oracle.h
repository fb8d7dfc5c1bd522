// g4oracle.h -- CPU ORACLE for the GVRS tile-codec path.  TEST INFRASTRUCTURE ONLY.
//
// This directory is a plain C++17 restatement of the reference's (gwlucastrig/gridfour, pure Java)
// tile-codec algorithms.  It exists so that the CUDA product path in gridfour_b200/ can be checked
// bit-for-bit.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load it; the product (gridfour_b200/, include/) never links or calls anything here.
//
// Parity pinning: the decode side is pinned by the reference's own binary fixtures
// (core/src/test/resources/org/gridfour/gvrs/SampleFiles/Sample04/05/06/07/14, extracted to
// tests/golden/ by tests/golden/make_golden.py) and by the CodecM32Test size table
// (core/src/test/java/org/gridfour/compress/CodecM32Test.java:93-110).  The encode side of
// HuffmanEncoder / CanonicalHuffman / LsEncoder12 has NO golden vector in the reference
// ("parity unpinned" for encoder output bytes): it is argued from line-by-line restatement plus
// round trips through the independently restated decoders.  java.util.zip is replaced by the
// system zlib (same algorithm family; the JDK version is unpinned by the reference).
//
// Paths below are relative to /root/reference/core/src/main/java/org/gridfour/ ("C/").
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <vector>
#include <stdexcept>

namespace g4o {

constexpr int32_t INT4_NULL_CODE = INT32_MIN;  // C/util/GridfourConstants.java:61

// ---------------------------------------------------------------------------------------------
// BitOutputStore / BitInputStore   (C/io/BitOutputStore.java:205-301, C/io/BitInputStore.java:95-220)
// LSB-first bit order, little-endian bytes.
// ---------------------------------------------------------------------------------------------
struct BitOut {
  std::vector<uint8_t> bytes;
  uint64_t scratch = 0;
  int64_t nBits = 0;
  void appendBit(int v) {  // BitOutputStore.java:205-215
    int k = int(nBits & 63);
    if (v) scratch |= (uint64_t(1) << k);
    nBits++;
    if (k == 63) flushWord();
  }
  void appendBits(int n, uint32_t value) {  // BitOutputStore.java:224-264
    if (n < 1 || n > 32) throw std::invalid_argument("appendBits range");
    uint64_t v = uint64_t(value) & (n == 32 ? 0xffffffffull : ((uint64_t(1) << n) - 1));
    int k = int(nBits & 63);
    int nFree = 64 - k;
    if (nFree < n) {
      scratch |= (v << k);  // low part (upper bits fall off)
      nBits += nFree;
      flushWord();
      scratch = v >> nFree;
      nBits += n - nFree;
    } else {
      scratch |= (v << k);
      nBits += n;
      if (k + n == 64) flushWord();
    }
  }
  void flushWord() {
    for (int i = 0; i < 8; i++) bytes.push_back(uint8_t(scratch >> (8 * i)));
    scratch = 0;
  }
  int64_t lengthBits() const { return nBits; }
  int64_t lengthBytes() const { return (nBits + 7) / 8; }  // BitOutputStore.java:299-301
  std::vector<uint8_t> text() const {                       // BitOutputStore.java:271-288
    std::vector<uint8_t> b = bytes;
    int k = int(nBits & 63);
    int nb = (k + 7) / 8;
    uint64_t s = scratch;
    for (int i = 0; i < nb; i++) { b.push_back(uint8_t(s)); s >>= 8; }
    return b;
  }
};

struct BitIn {
  const uint8_t* p;
  int64_t nBits;   // limit in bits
  int64_t iBit = 0;
  BitIn(const uint8_t* data, size_t lenBytes) : p(data), nBits(int64_t(lenBytes) * 8) {}
  int getBit() {  // BitInputStore.java:112-125
    if (iBit >= nBits) throw std::out_of_range("Attempt to read past end of data");
    int b = (p[iBit >> 3] >> (iBit & 7)) & 1;
    iBit++;
    return b;
  }
  uint32_t getBits(int n) {  // BitInputStore.java:134-173
    if (n < 1 || n > 32) throw std::invalid_argument("getBits range");
    if (iBit + n > nBits) throw std::out_of_range("Attempt to read past end of data");
    uint32_t v = 0;
    for (int i = 0; i < n; i++) {
      v |= uint32_t((p[iBit >> 3] >> (iBit & 7)) & 1) << i;
      iBit++;
    }
    return v;
  }
  int64_t position() const { return iBit; }
};

// ---------------------------------------------------------------------------------------------
// CodecM32  (C/compress/CodecM32.java:257-356)
// ---------------------------------------------------------------------------------------------
struct M32Writer {
  uint8_t* buf;
  size_t off = 0;
  explicit M32Writer(uint8_t* b) : buf(b) {}
  void encode(int32_t value);  // CodecM32.java:257-311
};
struct M32Reader {
  const uint8_t* buf;
  size_t off = 0;
  size_t limit;
  M32Reader(const uint8_t* b, size_t n) : buf(b), limit(n) {}
  int32_t decode();  // CodecM32.java:327-356 (bounds-checked here; the reference is unchecked)
};
constexpr int M32_MAX_BYTES_PER_VALUE = 6;  // CodecM32.java:180

// ---------------------------------------------------------------------------------------------
// Predictors (C/compress/PredictorModel*.java).  model codes: PredictorModelType.java:46-63
// ---------------------------------------------------------------------------------------------
enum Predictor { PRED_NONE = 0, PRED_DIFFERENCING = 1, PRED_LINEAR = 2, PRED_TRIANGLE = 3, PRED_DIFF_NULLS = 4 };

// Returns number of residual ints (or -1 / 0 on the reference's failure paths); seed via *seed.
int predictor_encode_int(int model, int nRows, int nCols, const int32_t* values, int32_t* out, int32_t* seed);
void predictor_decode_int(int model, int32_t seed, int nRows, int nCols, const int32_t* enc, size_t nEnc, int32_t* out);
// Byte (M32) flavour.  Returns number of M32 bytes (or -1 / 0).
int predictor_encode(int model, int nRows, int nCols, const int32_t* values, uint8_t* out, int32_t* seed);
void predictor_decode(int model, int32_t seed, int nRows, int nCols, const uint8_t* enc, size_t nEnc, int32_t* out);

// ---------------------------------------------------------------------------------------------
// Legacy Huffman (C/compress/HuffmanEncoder.java:124-305, HuffmanDecoder.java:65-187)
// ---------------------------------------------------------------------------------------------
void huffman_encode(BitOut& out, int nSymbols, const uint8_t* symbols);
void huffman_decode(BitIn& in, int nSymbols, uint8_t* symbols);
// code lengths only (for tests of the tie-breaking); returns number of distinct symbols
int huffman_code_lengths(int nSymbols, const uint8_t* symbols, int lengths[256]);

// ---------------------------------------------------------------------------------------------
// Canonical Huffman (C/compress/canonicalHuffman/*.java)
// ---------------------------------------------------------------------------------------------
void canon_encode(BitOut& out, int nSymbols, const int32_t* text);        // CanonicalHuffman.java:177-283
bool canon_decode(BitIn& in, int nSymbolsInText, int32_t* text);          // CanonicalHuffman.java:441-519
// exposed for unit tests
void canon_tree_lengths(const int* counts, int nSymbols, int* lengths, bool* limited);  // TreeBuilder.java:75-188
void package_merge(int maxLen, const int* sortedCounts, int n, int* nBits);             // PackageMerge.java:91-175
int length_encode(int n, const int* codeLen, int* codes, int* runs);                    // LengthEncoder.java:86-167

// ---------------------------------------------------------------------------------------------
// LSOP12 (C/lsop/*.java, C/util/jama/LUDecomposition.java)
// ---------------------------------------------------------------------------------------------
bool lsop12_coefficients(int nRows, int nCols, const int32_t* values, double ud[12]);
// LSOP08 (legacy 8-coefficient codec): lsop/LsOptimalPredictor08.java, LsEncoder08.java, LsDecoder08.java (g4o_lsop08.cpp)
bool lsop08_coefficients(int nRows, int nCols, const int32_t* values, double ud[8]);
bool codec_lsop08_encode(int codecIndex, int nRows, int nCols, const int32_t* v, std::vector<uint8_t>& out);
void codec_lsop08_decode(int nRows, int nCols, const uint8_t* packing, size_t len, int32_t* values);
bool lsop12_residual_streams(int nRows, int nCols, const int32_t* v, int32_t* seed, float u[12], uint8_t* initCodes, long* nInit,
                             uint8_t* interiorCodes, long* nInterior);  // LsOptimalPredictor12.java:109-292 (test infrastructure)  // LsOptimalPredictor12.java:311-383
int32_t java_round_float(float a);                                                     // StrictMath.round(float)
uint32_t crc32c(const uint8_t* p, size_t n);                                           // C/util/GridfourCRC32C.java

// ---------------------------------------------------------------------------------------------
// Codecs.  encode returns false for Java `null` (declined).  decode throws std::runtime_error
// where the reference throws IOException.
// ---------------------------------------------------------------------------------------------
enum CodecId { CODEC_HUFFMAN = 0, CODEC_DEFLATE = 1, CODEC_FLOAT = 2, CODEC_CANON_HUFFMAN = 3, CODEC_LSOP12 = 4 };

struct EncodeInfo {  // diagnostics for parity tests (what the GPU must also report)
  int predictor = 0;      // predictor code of the winner (LSOP: compression type)
  int32_t seed = 0;
};

bool codec_huffman_encode(int codecIndex, int nRows, int nCols, const int32_t* v, std::vector<uint8_t>& out, EncodeInfo* info = nullptr);
void codec_huffman_decode(int nRows, int nCols, const uint8_t* packing, size_t len, int32_t* out);
bool codec_deflate_encode(int codecIndex, int nRows, int nCols, const int32_t* v, std::vector<uint8_t>& out, EncodeInfo* info = nullptr);
bool codec_deflate_decode(int nRows, int nCols, const uint8_t* packing, size_t len, int32_t* out);  // false == Java null
bool codec_canon_encode(int codecIndex, int nRows, int nCols, const int32_t* v, std::vector<uint8_t>& out, EncodeInfo* info = nullptr);
void codec_canon_decode(int nRows, int nCols, const uint8_t* packing, size_t len, int32_t* out);
bool codec_float_encode(int codecIndex, int nRows, int nCols, const float* v, std::vector<uint8_t>& out);
void codec_float_decode(int nRows, int nCols, const uint8_t* packing, size_t len, float* out);
bool codec_lsop12_encode(int codecIndex, int nRows, int nCols, const int32_t* v, std::vector<uint8_t>& out,
                         bool deflateEnabled = true, bool checksum = false, EncodeInfo* info = nullptr);
void codec_lsop12_decode(int nRows, int nCols, const uint8_t* packing, size_t len, int32_t* out);

// zlib stand-in for java.util.zip (Deflater(level), finish(), one deflate() call into a capped buffer)
int zlib_deflate_capped(int level, const uint8_t* in, size_t n, uint8_t* out, size_t cap);
// returns bytes produced (<0 on data error); *consumed = input bytes read (Inflater.getBytesRead)
int zlib_inflate(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* consumed);

// CodecMaster.encodeSingleThread (C/gvrs/CodecMaster.java:150-169) + TileElementInt.encode raw
// fallback (C/gvrs/TileElementInt.java:196-206).  codecIds[k] is the codec at list position k.
// Returns the element payload (raw little-endian when no codec beats 4*n bytes).
void master_encode_i32(const int* codecIds, int nCodecs, int nRows, int nCols, const int32_t* v, std::vector<uint8_t>& out);
void master_decode_i32(const int* codecIds, int nCodecs, int nRows, int nCols, const uint8_t* payload, size_t len, int32_t* out);
void master_encode_f32(const int* codecIds, int nCodecs, int nRows, int nCols, const float* v, std::vector<uint8_t>& out);
void master_decode_f32(const int* codecIds, int nCodecs, int nRows, int nCols, const uint8_t* payload, size_t len, float* out);

}  // namespace g4o
