// CPU ORACLE (test infrastructure only) -- flat C entry points for ctypes (tests/, bench.py cpu_baseline).
// Java exceptions restated as C++ exceptions are mapped to negative status codes here.
#include "g4oracle.h"
#include "../include/g4terrain.h"
#include <thread>
#include <atomic>
#include <algorithm>

using namespace g4o;

#define G4O_TRY try {
#define G4O_CATCH(errval) } catch (const std::exception&) { return (errval); }

extern "C" {

int g4o_m32_encode(const int32_t* v, int n, uint8_t* out) {
  M32Writer w(out);
  for (int i = 0; i < n; i++) w.encode(v[i]);
  return int(w.off);
}
int g4o_m32_decode(const uint8_t* b, int nb, int32_t* out, int maxn) {
  G4O_TRY
  M32Reader r(b, size_t(nb));
  int k = 0;
  while (r.off < r.limit && k < maxn) out[k++] = r.decode();
  return k;
  G4O_CATCH(-1)
}
int g4o_predictor_encode(int model, int nr, int nc, const int32_t* v, uint8_t* out, int32_t* seed) {
  G4O_TRY return predictor_encode(model, nr, nc, v, out, seed); G4O_CATCH(-2)
}
int g4o_predictor_encode_int(int model, int nr, int nc, const int32_t* v, int32_t* out, int32_t* seed) {
  G4O_TRY return predictor_encode_int(model, nr, nc, v, out, seed); G4O_CATCH(-2)
}
int g4o_predictor_decode(int model, int32_t seed, int nr, int nc, const uint8_t* enc, int n, int32_t* out) {
  G4O_TRY predictor_decode(model, seed, nr, nc, enc, size_t(n), out); return 0; G4O_CATCH(-1)
}
int g4o_predictor_decode_int(int model, int32_t seed, int nr, int nc, const int32_t* enc, int n, int32_t* out) {
  G4O_TRY predictor_decode_int(model, seed, nr, nc, enc, size_t(n), out); return 0; G4O_CATCH(-1)
}

long g4o_huffman_encode(const uint8_t* sym, int n, uint8_t* out, long cap) {
  G4O_TRY
  BitOut o;
  huffman_encode(o, n, sym);
  std::vector<uint8_t> t = o.text();
  if (long(t.size()) > cap) return -2;
  std::memcpy(out, t.data(), t.size());
  return long(o.lengthBits());
  G4O_CATCH(-1)
}
int g4o_huffman_decode(const uint8_t* in, long nbytes, int nsym, uint8_t* out, long* bitpos) {
  G4O_TRY
  BitIn b(in, size_t(nbytes));
  huffman_decode(b, nsym, out);
  if (bitpos) *bitpos = long(b.position());
  return 0;
  G4O_CATCH(-1)
}
int g4o_huffman_code_lengths(const uint8_t* sym, int n, int* lengths) {
  G4O_TRY return huffman_code_lengths(n, sym, lengths); G4O_CATCH(-1)
}
long g4o_canon_encode(const int32_t* text, int n, uint8_t* out, long cap) {
  G4O_TRY
  BitOut o;
  canon_encode(o, n, text);
  std::vector<uint8_t> t = o.text();
  if (long(t.size()) > cap) return -2;
  std::memcpy(out, t.data(), t.size());
  return long(o.lengthBits());
  G4O_CATCH(-1)
}
int g4o_canon_decode(const uint8_t* in, long nbytes, int nsym, int32_t* out, long* bitpos) {
  G4O_TRY
  BitIn b(in, size_t(nbytes));
  canon_decode(b, nsym, out);
  if (bitpos) *bitpos = long(b.position());
  return 0;
  G4O_CATCH(-1)
}
int g4o_canon_tree_lengths(const int* counts, int n, int* lengths) {
  G4O_TRY
  bool limited = false;
  canon_tree_lengths(counts, n, lengths, &limited);
  return limited ? 1 : 0;
  G4O_CATCH(-1)
}
int g4o_package_merge(int maxLen, const int* sortedCounts, int n, int* nBits) {
  G4O_TRY package_merge(maxLen, sortedCounts, n, nBits); return 0; G4O_CATCH(-1)
}
int g4o_length_encode(int n, const int* codeLen, int* codes, int* runs) {
  G4O_TRY return length_encode(n, codeLen, codes, runs); G4O_CATCH(-1)
}
int g4o_lsop12_coefficients(int nr, int nc, const int32_t* v, double* ud) {
  return lsop12_coefficients(nr, nc, v, ud) ? 1 : 0;
}
int g4o_lsop12_residual_streams(int nr, int nc, const int32_t* v, int32_t* seed, float* u, uint8_t* initCodes, long* nInit,
                                uint8_t* interiorCodes, long* nInterior) {
  G4O_TRY return lsop12_residual_streams(nr, nc, v, seed, u, initCodes, nInit, interiorCodes, nInterior) ? 1 : 0; G4O_CATCH(-1)
}
// Huffman decode that starts at bit *bitpos of `in` (two streams back to back in one bit store, LsDecoder12.java:119-124)
int g4o_huffman_decode_at(const uint8_t* in, long nbytes, int nsym, uint8_t* out, long* bitpos) {
  G4O_TRY
  BitIn b(in, size_t(nbytes));
  b.iBit = *bitpos;
  huffman_decode(b, nsym, out);
  *bitpos = long(b.position());
  return 0;
  G4O_CATCH(-1)
}
// LSOP08: returns the packing length, -1 when the codec declines (null / singular matrix), -2 cap too small, -3 exception
long g4o_lsop08_encode(int codecIndex, int nr, int nc, const int32_t* v, uint8_t* out, long cap) {
  G4O_TRY
  std::vector<uint8_t> p;
  if (!codec_lsop08_encode(codecIndex, nr, nc, v, p)) return -1;
  if (long(p.size()) > cap) return -2;
  std::copy(p.begin(), p.end(), out);
  return long(p.size());
  G4O_CATCH(-3)
}
int g4o_lsop08_decode(int nr, int nc, const uint8_t* packing, long len, int32_t* out) {
  G4O_TRY
  codec_lsop08_decode(nr, nc, packing, size_t(len), out);
  return 0;
  G4O_CATCH(-1)
}
int32_t g4o_java_round(float a) { return java_round_float(a); }
uint32_t g4o_crc32c(const uint8_t* p, long n) { return crc32c(p, size_t(n)); }

// returns packing length, -1 when the codec declines (Java null), -2 cap too small, -3 exception
long g4o_codec_encode_i32(int codec, int codecIndex, int nr, int nc, const int32_t* v, uint8_t* out, long cap, int* predictor) {
  // the reference's Linear predictor reads columns 0 and 1 of every row (PredictorModelLinear.java:120-150): one-column
  // tiles are outside its domain (ArrayIndexOutOfBoundsException there), refused here
  if (nr < 1 || nc < 2) return -3;
  G4O_TRY
  std::vector<uint8_t> p;
  EncodeInfo info;
  bool ok = false;
  switch (codec) {
    case CODEC_HUFFMAN: ok = codec_huffman_encode(codecIndex, nr, nc, v, p, &info); break;
    case CODEC_DEFLATE: ok = codec_deflate_encode(codecIndex, nr, nc, v, p, &info); break;
    case CODEC_CANON_HUFFMAN: ok = codec_canon_encode(codecIndex, nr, nc, v, p, &info); break;
    case CODEC_LSOP12: ok = codec_lsop12_encode(codecIndex, nr, nc, v, p, true, false, &info); break;
    default: return -3;
  }
  if (!ok) return -1;
  if (long(p.size()) > cap) return -2;
  std::memcpy(out, p.data(), p.size());
  if (predictor) *predictor = info.predictor;
  return long(p.size());
  G4O_CATCH(-3)
}
long g4o_lsop12_encode_opts(int codecIndex, int nr, int nc, const int32_t* v, uint8_t* out, long cap, int deflateEnabled, int checksum) {
  if (nr < 1 || nc < 2) return -3;
  G4O_TRY
  std::vector<uint8_t> p;
  if (!codec_lsop12_encode(codecIndex, nr, nc, v, p, deflateEnabled != 0, checksum != 0, nullptr)) return -1;
  if (long(p.size()) > cap) return -2;
  std::memcpy(out, p.data(), p.size());
  return long(p.size());
  G4O_CATCH(-3)
}
// 0 ok, 1 decoder returned null, -1 IOException-class error
int g4o_codec_decode_i32(int codec, int nr, int nc, const uint8_t* p, long len, int32_t* out) {
  G4O_TRY
  switch (codec) {
    case CODEC_HUFFMAN: codec_huffman_decode(nr, nc, p, size_t(len), out); return 0;
    case CODEC_DEFLATE: return codec_deflate_decode(nr, nc, p, size_t(len), out) ? 0 : 1;
    case CODEC_CANON_HUFFMAN: codec_canon_decode(nr, nc, p, size_t(len), out); return 0;
    case CODEC_LSOP12: codec_lsop12_decode(nr, nc, p, size_t(len), out); return 0;
    default: return -1;
  }
  G4O_CATCH(-1)
}
long g4o_codec_encode_f32(int codecIndex, int nr, int nc, const float* v, uint8_t* out, long cap) {
  G4O_TRY
  std::vector<uint8_t> p;
  if (!codec_float_encode(codecIndex, nr, nc, v, p)) return -1;
  if (long(p.size()) > cap) return -2;
  std::memcpy(out, p.data(), p.size());
  return long(p.size());
  G4O_CATCH(-3)
}
int g4o_codec_decode_f32(int nr, int nc, const uint8_t* p, long len, float* out) {
  G4O_TRY codec_float_decode(nr, nc, p, size_t(len), out); return 0; G4O_CATCH(-1)
}
long g4o_master_encode_i32(const int* ids, int nIds, int nr, int nc, const int32_t* v, uint8_t* out, long cap) {
  if (nr < 1 || nc < 2) return -3;
  G4O_TRY
  std::vector<uint8_t> p;
  master_encode_i32(ids, nIds, nr, nc, v, p);
  if (long(p.size()) > cap) return -2;
  std::memcpy(out, p.data(), p.size());
  return long(p.size());
  G4O_CATCH(-3)
}
int g4o_master_decode_i32(const int* ids, int nIds, int nr, int nc, const uint8_t* p, long len, int32_t* out) {
  G4O_TRY master_decode_i32(ids, nIds, nr, nc, p, size_t(len), out); return 0; G4O_CATCH(-1)
}
long g4o_master_encode_f32(const int* ids, int nIds, int nr, int nc, const float* v, uint8_t* out, long cap) {
  G4O_TRY
  std::vector<uint8_t> p;
  master_encode_f32(ids, nIds, nr, nc, v, p);
  if (long(p.size()) > cap) return -2;
  std::memcpy(out, p.data(), p.size());
  return long(p.size());
  G4O_CATCH(-3)
}
int g4o_master_decode_f32(const int* ids, int nIds, int nr, int nc, const uint8_t* p, long len, float* out) {
  G4O_TRY master_decode_f32(ids, nIds, nr, nc, p, size_t(len), out); return 0; G4O_CATCH(-1)
}

// ---------------------------------------------------------------------------------------------
// Tile-pool batch drivers (the CPU baseline: BASELINE.md section 2).  The grid is row-major with
// `gridCols` columns; tile t = (tr, tc) covers rows [tr*tileRows, ...), cols [tc*tileCols, ...).
// Each worker thread owns private codec state (reference codec objects are stateful).
// arena: nTiles slots of slotBytes; lens[t] = payload length.  is_float selects the f32 master.
// Returns 0, or -(index+1) of the first failing tile.
// ---------------------------------------------------------------------------------------------
static void gather_tile(const uint32_t* grid, long gridCols, int tr, int tc, int tileRows, int tileCols, uint32_t* tile) {
  for (int r = 0; r < tileRows; r++)
    std::memcpy(tile + size_t(r) * tileCols, grid + (size_t(tr) * tileRows + r) * gridCols + size_t(tc) * tileCols, size_t(tileCols) * 4);
}
static void scatter_tile(uint32_t* grid, long gridCols, int tr, int tc, int tileRows, int tileCols, const uint32_t* tile) {
  for (int r = 0; r < tileRows; r++)
    std::memcpy(grid + (size_t(tr) * tileRows + r) * gridCols + size_t(tc) * tileCols, tile + size_t(r) * tileCols, size_t(tileCols) * 4);
}

long g4o_encode_grid(const int* ids, int nIds, int is_float, const void* grid, long gridRows, long gridCols,
                     int tileRows, int tileCols, int nThreads, uint8_t* arena, long slotBytes, uint32_t* lens) {
  const long tRows = gridRows / tileRows, tCols = gridCols / tileCols, nTiles = tRows * tCols;
  std::atomic<long> next(0), firstErr(0);
  auto work = [&]() {
    std::vector<uint32_t> tile(size_t(tileRows) * tileCols);
    std::vector<uint8_t> p;
    for (;;) {
      long t = next.fetch_add(1);
      if (t >= nTiles) break;
      try {
        gather_tile(static_cast<const uint32_t*>(grid), gridCols, int(t / tCols), int(t % tCols), tileRows, tileCols, tile.data());
        if (is_float) master_encode_f32(ids, nIds, tileRows, tileCols, reinterpret_cast<const float*>(tile.data()), p);
        else master_encode_i32(ids, nIds, tileRows, tileCols, reinterpret_cast<const int32_t*>(tile.data()), p);
        if (long(p.size()) > slotBytes) throw std::runtime_error("slot too small");
        std::memcpy(arena + t * slotBytes, p.data(), p.size());
        lens[t] = uint32_t(p.size());
      } catch (const std::exception&) {
        long z = 0;
        firstErr.compare_exchange_strong(z, -(t + 1));
      }
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nThreads; i++) th.emplace_back(work);
  work();
  for (auto& x : th) x.join();
  return firstErr.load();
}

long g4o_decode_grid(const int* ids, int nIds, int is_float, const uint8_t* arena, const uint64_t* offsets, const uint32_t* lens,
                     long gridRows, long gridCols, int tileRows, int tileCols, int nThreads, void* grid) {
  const long tRows = gridRows / tileRows, tCols = gridCols / tileCols, nTiles = tRows * tCols;
  std::atomic<long> next(0), firstErr(0);
  auto work = [&]() {
    std::vector<uint32_t> tile(size_t(tileRows) * tileCols);
    for (;;) {
      long t = next.fetch_add(1);
      if (t >= nTiles) break;
      try {
        if (is_float) master_decode_f32(ids, nIds, tileRows, tileCols, arena + offsets[t], lens[t], reinterpret_cast<float*>(tile.data()));
        else master_decode_i32(ids, nIds, tileRows, tileCols, arena + offsets[t], lens[t], reinterpret_cast<int32_t*>(tile.data()));
        scatter_tile(static_cast<uint32_t*>(grid), gridCols, int(t / tCols), int(t % tCols), tileRows, tileCols, tile.data());
      } catch (const std::exception&) {
        long z = 0;
        firstErr.compare_exchange_strong(z, -(t + 1));
      }
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nThreads; i++) th.emplace_back(work);
  work();
  for (auto& x : th) x.join();
  return firstErr.load();
}

int g4o_hardware_threads() { return int(std::thread::hardware_concurrency()); }

// synthetic terrain (include/g4terrain.h), rows [row0,row0+nr) x cols [col0,col0+nc)
void g4o_terrain_i32(uint64_t seed, long row0, long col0, long nr, long nc, int32_t* out, int nThreads) {
  auto work = [&](long r0, long r1) {
    for (long r = r0; r < r1; r++)
      for (long c = 0; c < nc; c++) out[r * nc + c] = g4_terrain_m(seed, row0 + r, col0 + c);
  };
  std::vector<std::thread> th;
  nThreads = std::max(1, nThreads);
  long per = (nr + nThreads - 1) / nThreads;
  for (int i = 0; i < nThreads; i++) {
    long a = i * per, b = std::min(nr, a + per);
    if (a < b) th.emplace_back(work, a, b);
  }
  for (auto& x : th) x.join();
}
void g4o_terrain_f32(uint64_t seed, long row0, long col0, long nr, long nc, float* out, int nThreads) {
  auto work = [&](long r0, long r1) {
    for (long r = r0; r < r1; r++)
      for (long c = 0; c < nc; c++) out[r * nc + c] = float(g4_terrain_dm(seed, row0 + r, col0 + c)) * 0.1f;
  };
  std::vector<std::thread> th;
  nThreads = std::max(1, nThreads);
  long per = (nr + nThreads - 1) / nThreads;
  for (int i = 0; i < nThreads; i++) {
    long a = i * per, b = std::min(nr, a + per);
    if (a < b) th.emplace_back(work, a, b);
  }
  for (auto& x : th) x.join();
}

}  // extern "C"
