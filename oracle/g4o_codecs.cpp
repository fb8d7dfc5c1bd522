// CPU ORACLE (test infrastructure only) -- the codec classes and the CodecMaster selection rule.
// Restates C/compress/{CodecHuffman,CodecDeflate,CodecFloat}.java, C/compress/canonicalHuffman/
// CodecCanonHuffman.java, C/gvrs/CodecMaster.java:150-203, C/gvrs/TileElementInt.java:196-219 and
// TileElementFloat.java:209-232 (C/ = /root/reference/core/src/main/java/org/gridfour/).
// java.util.zip is replaced by the system zlib (deflateInit(level), one Z_FINISH call, capped output).
#include "g4oracle.h"
#include <zlib.h>

namespace g4o {

int zlib_deflate_capped(int level, const uint8_t* in, size_t n, uint8_t* out, size_t cap) {
  z_stream s;
  std::memset(&s, 0, sizeof(s));
  if (deflateInit(&s, level) != Z_OK) return -1;
  s.next_in = const_cast<Bytef*>(in);
  s.avail_in = uInt(n);
  s.next_out = out;
  s.avail_out = uInt(cap);
  deflate(&s, Z_FINISH);  // Deflater.finish() + one deflate() call: a non-fitting stream is truncated
  int produced = int(cap - s.avail_out);
  deflateEnd(&s);
  return produced;
}

int zlib_inflate(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* consumed) {
  z_stream s;
  std::memset(&s, 0, sizeof(s));
  if (inflateInit(&s) != Z_OK) return -1;
  s.next_in = const_cast<Bytef*>(in);
  s.avail_in = uInt(n);
  s.next_out = out;
  s.avail_out = uInt(cap);
  int rc = inflate(&s, Z_SYNC_FLUSH);  // Inflater.inflate(buf): one call, stops at stream end or full buffer
  int produced = int(cap - s.avail_out);
  if (consumed) *consumed = size_t(s.total_in);
  inflateEnd(&s);
  if (rc == Z_DATA_ERROR || rc == Z_NEED_DICT || rc == Z_MEM_ERROR) return -1;  // DataFormatException
  return produced;
}

namespace {
void scan_nulls(size_t n, const int32_t* v, bool* containsNull, bool* containsValid) {
  *containsNull = false; *containsValid = false;
  for (size_t i = 0; i < n; i++) {
    if (v[i] == INT4_NULL_CODE) *containsNull = true; else *containsValid = true;
  }
}
const int kModels[4] = {PRED_DIFFERENCING, PRED_LINEAR, PRED_TRIANGLE, PRED_DIFF_NULLS};
inline bool model_supports_nulls(int m) { return m == PRED_DIFF_NULLS; }
inline uint32_t rd32(const uint8_t* p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24); }
}  // namespace

// CodecHuffman.encode (CodecHuffman.java:70-119) + compress (:121-130)
bool codec_huffman_encode(int codecIndex, int nRows, int nCols, const int32_t* v, std::vector<uint8_t>& out, EncodeInfo* info) {
  size_t n = size_t(nRows) * nCols;
  bool cn, cv;
  scan_nulls(n, v, &cn, &cv);
  if (!cv) return false;
  std::vector<uint8_t> mCode(n * M32_MAX_BYTES_PER_VALUE);
  int64_t resultLength = INT32_MAX;
  bool have = false;
  for (int m : kModels) {
    if (cn != model_supports_nulls(m)) continue;
    int32_t seed = 0;
    int mLen = predictor_encode(m, nRows, nCols, v, mCode.data(), &seed);
    if (mLen > 0) {
      BitOut store;
      store.appendBits(8, uint32_t(codecIndex));
      store.appendBits(8, uint32_t(m));
      store.appendBits(32, uint32_t(seed));
      store.appendBits(32, uint32_t(mLen));
      huffman_encode(store, mLen, mCode.data());
      int64_t testLength = store.lengthBytes();
      if (testLength < resultLength) {
        resultLength = testLength;
        out = store.text();
        have = true;
        if (info) { info->predictor = m; info->seed = seed; }
      }
    }
  }
  return have;
}

// CodecHuffman.decode (:133-153)
void codec_huffman_decode(int nRows, int nCols, const uint8_t* packing, size_t len, int32_t* out) {
  if (len < 10) throw std::runtime_error("packing too short");
  int model = int8_t(packing[1]);
  if (model < 1 || model > 4) throw std::runtime_error("Unknown PredictorCorrector type");
  int32_t seed = int32_t(rd32(packing + 2));
  int32_t nM32 = int32_t(rd32(packing + 6));
  if (nM32 < 0) throw std::runtime_error("negative M32 count");  // Java: NegativeArraySizeException
  std::vector<uint8_t> codes(size_t(nM32) + 1);
  BitIn in(packing + 10, len - 10);
  huffman_decode(in, nM32, codes.data());
  predictor_decode(model, seed, nRows, nCols, codes.data(), size_t(nM32), out);
}

// CodecDeflate.encode (CodecDeflate.java:157-202) + compress (:204-228)
bool codec_deflate_encode(int codecIndex, int nRows, int nCols, const int32_t* v, std::vector<uint8_t>& out, EncodeInfo* info) {
  size_t n = size_t(nRows) * nCols;
  bool cn, cv;
  scan_nulls(n, v, &cn, &cv);
  if (!cv) return false;
  std::vector<uint8_t> mCode(n * M32_MAX_BYTES_PER_VALUE);
  int64_t resultLength = INT32_MAX;
  bool have = false;
  for (int m : kModels) {
    if (cn != model_supports_nulls(m)) continue;
    int32_t seed = 0;
    int mLen = predictor_encode(m, nRows, nCols, v, mCode.data(), &seed);
    if (mLen > 0) {
      std::vector<uint8_t> res(size_t(mLen) + 128);
      int dN = zlib_deflate_capped(6, mCode.data(), size_t(mLen), res.data() + 10, res.size() - 10);
      if (dN <= 0) continue;
      res[0] = uint8_t(codecIndex);
      res[1] = uint8_t(m);
      for (int k = 0; k < 4; k++) res[2 + k] = uint8_t(uint32_t(seed) >> (8 * k));
      for (int k = 0; k < 4; k++) res[6 + k] = uint8_t(uint32_t(mLen) >> (8 * k));
      res.resize(size_t(dN) + 10);
      if (int64_t(res.size()) < resultLength) {
        resultLength = int64_t(res.size());
        out = res;
        have = true;
        if (info) { info->predictor = m; info->seed = seed; }
      }
    }
  }
  return have;
}

// CodecDeflate.decode (:109-154)
bool codec_deflate_decode(int nRows, int nCols, const uint8_t* packing, size_t len, int32_t* out) {
  if (len < 10) throw std::runtime_error("packing too short");
  int model = int8_t(packing[1]);
  if (model < 1 || model > 4) throw std::runtime_error("Unknown PredictorCorrector type");
  int32_t seed = int32_t(rd32(packing + 2));
  int32_t nM32 = int32_t(rd32(packing + 6));
  if (nM32 < 0) throw std::runtime_error("negative M32 count");
  std::vector<uint8_t> codes(size_t(nM32) + 1);
  int t = zlib_inflate(packing + 10, len - 10, codes.data(), size_t(nM32), nullptr);
  if (t < 0) throw std::runtime_error("zlib data error");
  if (t > 0) {
    predictor_decode(model, seed, nRows, nCols, codes.data(), size_t(nM32), out);
    return true;
  }
  return false;
}

// CodecCanonHuffman.encode (CodecCanonHuffman.java:79-142) + compress (:144-159)
bool codec_canon_encode(int codecIndex, int nRows, int nCols, const int32_t* v, std::vector<uint8_t>& out, EncodeInfo* info) {
  size_t n = size_t(nRows) * nCols;
  bool cn, cv;
  scan_nulls(n, v, &cn, &cv);
  if (!cv) return false;
  bool uniform = true;
  for (size_t i = 1; i < n; i++) if (v[i] != v[0]) { uniform = false; break; }
  if (uniform) {
    out.assign(6, 0);
    out[0] = uint8_t(codecIndex);
    out[1] = 0;
    for (int k = 0; k < 4; k++) out[2 + k] = uint8_t(uint32_t(v[0]) >> (8 * k));
    if (info) { info->predictor = 0; info->seed = v[0]; }
    return true;
  }
  int64_t resultLength = INT32_MAX;
  bool have = false;
  std::vector<int32_t> residuals(n);
  for (int m : kModels) {
    if (cn != model_supports_nulls(m)) continue;
    int32_t seed = 0;
    int nRes = predictor_encode_int(m, nRows, nCols, v, residuals.data(), &seed);
    // the reference passes nRes straight to CanonicalHuffman.encode, which throws for nRes <= 0
    if (nRes <= 0) throw std::invalid_argument("Empty or null data input data");
    BitOut body;
    canon_encode(body, nRes, residuals.data());
    std::vector<uint8_t> b = body.text();
    std::vector<uint8_t> res(6);
    res[0] = uint8_t(codecIndex);
    res[1] = uint8_t(m);
    for (int k = 0; k < 4; k++) res[2 + k] = uint8_t(uint32_t(seed) >> (8 * k));
    res.insert(res.end(), b.begin(), b.end());
    if (int64_t(res.size()) < resultLength) {
      resultLength = int64_t(res.size());
      out = res;
      have = true;
      if (info) { info->predictor = m; info->seed = seed; }
    }
  }
  return have;
}

// CodecCanonHuffman.decode (:162-195)
void codec_canon_decode(int nRows, int nCols, const uint8_t* packing, size_t len, int32_t* out) {
  if (len < 6) throw std::runtime_error("packing too short");
  size_t n = size_t(nRows) * nCols;
  int model = int8_t(packing[1]);
  int32_t seed = int32_t(rd32(packing + 2));
  if (model == 0 && len == 6) { for (size_t i = 0; i < n; i++) out[i] = seed; return; }
  if (model < 1 || model > 4) throw std::runtime_error("Unknown PredictorCorrector type");
  std::vector<int32_t> residuals(n, 0);
  BitIn in(packing + 6, len - 6);
  canon_decode(in, int(n), residuals.data());
  predictor_decode_int(model, seed, nRows, nCols, residuals.data(), n, out);
}

// CodecFloat (CodecFloat.java:300-458)
namespace {
void encode_deltas(uint8_t* s, int nRows, int nCols) {  // :300-313
  int prior0 = 0, k = 0;
  for (int r = 0; r < nRows; r++) {
    int prior = prior0;
    prior0 = int8_t(s[k]);
    for (int c = 0; c < nCols; c++) { int test = int8_t(s[k]); s[k++] = uint8_t(test - prior); prior = test; }
  }
}
void decode_deltas(uint8_t* s, int nRows, int nCols) {  // :315-325
  int prior = 0, k = 0;
  for (int r = 0; r < nRows; r++) {
    for (int c = 0; c < nCols; c++) { prior += int8_t(s[k]); s[k++] = uint8_t(prior); }
    prior = int8_t(s[r * nCols]);
  }
}
std::vector<uint8_t> do_deflate9(const std::vector<uint8_t>& in) {  // :268-283
  std::vector<uint8_t> r(in.size() + 128);
  int d = zlib_deflate_capped(9, in.data(), in.size(), r.data(), r.size());
  if (d <= 0) throw std::runtime_error("Deflate failed");
  r.resize(size_t(d));
  return r;
}
}  // namespace

bool codec_float_encode(int codecIndex, int nRows, int nCols, const float* values, std::vector<uint8_t>& out) {
  size_t n = size_t(nRows) * nCols;
  std::vector<uint32_t> c(n);
  std::memcpy(c.data(), values, n * 4);
  BitOut bSign;
  for (size_t i = 0; i < n; i++) bSign.appendBit(int((c[i] >> 31) & 1));
  std::vector<uint8_t> comp[5];
  comp[0] = do_deflate9(bSign.text());
  std::vector<uint8_t> scratch(n);
  for (size_t i = 0; i < n; i++) scratch[i] = uint8_t((c[i] >> 23) & 0xff);
  comp[1] = do_deflate9(scratch);
  for (size_t i = 0; i < n; i++) scratch[i] = uint8_t((c[i] >> 16) & 0x7f);
  encode_deltas(scratch.data(), nRows, nCols);
  comp[2] = do_deflate9(scratch);
  for (size_t i = 0; i < n; i++) scratch[i] = uint8_t((c[i] >> 8) & 0xff);
  encode_deltas(scratch.data(), nRows, nCols);
  comp[3] = do_deflate9(scratch);
  for (size_t i = 0; i < n; i++) scratch[i] = uint8_t(c[i] & 0xff);
  encode_deltas(scratch.data(), nRows, nCols);
  comp[4] = do_deflate9(scratch);
  out.clear();
  out.push_back(uint8_t(codecIndex));
  out.push_back(0);
  for (int k = 0; k < 5; k++) {
    uint32_t L = uint32_t(comp[k].size());
    for (int b = 0; b < 4; b++) out.push_back(uint8_t(L >> (8 * b)));
    out.insert(out.end(), comp[k].begin(), comp[k].end());
  }
  return true;
}

void codec_float_decode(int nRows, int nCols, const uint8_t* packing, size_t len, float* out) {
  size_t n = size_t(nRows) * nCols;
  std::vector<uint8_t> scratch(n + 8);
  std::vector<uint32_t> raw(n, 0);
  size_t nSign = (n + 7) / 8;
  size_t off = 2;
  auto section = [&](size_t want) {
    if (off + 4 > len) throw std::runtime_error("CodecFloat packing truncated");
    uint32_t L = rd32(packing + off);
    off += 4;
    if (off + L > len) throw std::runtime_error("CodecFloat packing truncated");
    int t = zlib_inflate(packing + off, L, scratch.data(), want, nullptr);
    if (t < 0) throw std::runtime_error("Inflate failed");
    off += L;
  };
  section(nSign);
  for (size_t i = 0; i < n; i++) raw[i] = uint32_t((scratch[i >> 3] >> (i & 7)) & 1) << 31;
  section(n);
  for (size_t i = 0; i < n; i++) raw[i] |= uint32_t(scratch[i]) << 23;
  section(n);
  decode_deltas(scratch.data(), nRows, nCols);
  for (size_t i = 0; i < n; i++) raw[i] |= uint32_t(scratch[i] & 0x7f) << 16;
  section(n);
  decode_deltas(scratch.data(), nRows, nCols);
  for (size_t i = 0; i < n; i++) raw[i] |= uint32_t(scratch[i]) << 8;
  section(n);
  decode_deltas(scratch.data(), nRows, nCols);
  for (size_t i = 0; i < n; i++) raw[i] |= uint32_t(scratch[i]);
  std::memcpy(out, raw.data(), n * 4);
}

// ---------------------------------------------------------------------------------------------
// CodecMaster.encodeSingleThread (CodecMaster.java:150-169) + TileElementInt/Float raw fallback
// ---------------------------------------------------------------------------------------------
static bool implements_int(int id) { return id != CODEC_FLOAT; }
static bool implements_float(int id) { return id == CODEC_FLOAT; }

void master_encode_i32(const int* codecIds, int nCodecs, int nRows, int nCols, const int32_t* v, std::vector<uint8_t>& out) {
  size_t n = size_t(nRows) * nCols;
  std::vector<uint8_t> best, test;
  int64_t bestLen = INT32_MAX;
  bool have = false;
  for (int k = 0; k < nCodecs; k++) {
    if (!implements_int(codecIds[k])) continue;
    bool ok = false;
    switch (codecIds[k]) {
      case CODEC_HUFFMAN: ok = codec_huffman_encode(k, nRows, nCols, v, test); break;
      case CODEC_DEFLATE: ok = codec_deflate_encode(k, nRows, nCols, v, test); break;
      case CODEC_CANON_HUFFMAN: ok = codec_canon_encode(k, nRows, nCols, v, test); break;
      case CODEC_LSOP12: ok = codec_lsop12_encode(k, nRows, nCols, v, test); break;
      default: throw std::invalid_argument("unknown codec id");
    }
    if (ok && int64_t(test.size()) < bestLen) { best = test; bestLen = int64_t(test.size()); have = true; }
  }
  if (!have || best.size() >= n * 4) {  // TileElementInt.java:198-204
    out.resize(n * 4);
    for (size_t i = 0; i < n; i++) for (int b = 0; b < 4; b++) out[4 * i + b] = uint8_t(uint32_t(v[i]) >> (8 * b));
  } else {
    out = best;
  }
}

void master_decode_i32(const int* codecIds, int nCodecs, int nRows, int nCols, const uint8_t* payload, size_t len, int32_t* out) {
  size_t n = size_t(nRows) * nCols;
  if (len == n * 4) {  // TileElementInt.java:210-215
    for (size_t i = 0; i < n; i++) out[i] = int32_t(rd32(payload + 4 * i));
    return;
  }
  if (len < 1) throw std::runtime_error("empty packing");
  int index = payload[0];
  if (index >= nCodecs) throw std::runtime_error("Invalid compression-type code");  // CodecMaster.java:197-199
  switch (codecIds[index]) {
    case CODEC_HUFFMAN: codec_huffman_decode(nRows, nCols, payload, len, out); break;
    case CODEC_DEFLATE:
      if (!codec_deflate_decode(nRows, nCols, payload, len, out)) throw std::runtime_error("deflate decode returned null");
      break;
    case CODEC_CANON_HUFFMAN: codec_canon_decode(nRows, nCols, payload, len, out); break;
    case CODEC_LSOP12: codec_lsop12_decode(nRows, nCols, payload, len, out); break;
    default: throw std::runtime_error("codec does not decode integers");
  }
}

void master_encode_f32(const int* codecIds, int nCodecs, int nRows, int nCols, const float* v, std::vector<uint8_t>& out) {
  size_t n = size_t(nRows) * nCols;
  std::vector<uint8_t> best, test;
  int64_t bestLen = INT32_MAX;
  bool have = false;
  for (int k = 0; k < nCodecs; k++) {  // CodecMaster.encode(float) mirrors the int rule
    if (!implements_float(codecIds[k])) continue;
    if (codec_float_encode(k, nRows, nCols, v, test) && int64_t(test.size()) < bestLen) {
      best = test; bestLen = int64_t(test.size()); have = true;
    }
  }
  if (!have || best.size() >= n * 4) {
    out.resize(n * 4);
    std::memcpy(out.data(), v, n * 4);
  } else {
    out = best;
  }
}

void master_decode_f32(const int* codecIds, int nCodecs, int nRows, int nCols, const uint8_t* payload, size_t len, float* out) {
  size_t n = size_t(nRows) * nCols;
  if (len == n * 4) { std::memcpy(out, payload, n * 4); return; }
  if (len < 2) throw std::runtime_error("empty packing");
  int index = payload[0];
  if (index >= nCodecs) throw std::runtime_error("Invalid compression-type code");
  if (codecIds[index] != CODEC_FLOAT) throw std::runtime_error("codec does not decode floats");
  codec_float_decode(nRows, nCols, payload, len, out);
}

}  // namespace g4o
