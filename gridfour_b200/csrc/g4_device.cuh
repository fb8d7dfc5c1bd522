// g4_device.cuh -- device-side building blocks shared by the sm_100a tile-codec kernels.
//
// Reference semantics (paths relative to /root/reference/core/src/main/java/org/gridfour/):
//   M32 varint                compress/CodecM32.java:257-356
//   predictor stream orders   compress/PredictorModel{Differencing,Linear,Triangle}.java
//   LSB-first bit streams     io/BitOutputStore.java:205-301, io/BitInputStore.java:95-220
// All integer arithmetic is uint32 wrap-around == Java int.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/g4codec.h"

namespace g4 {

constexpr int kThreads = 256;  // threads per CTA for all tile kernels (8 warps)
constexpr int kWarps = kThreads / 32;
constexpr int32_t kNull = INT32_MIN;  // util/GridfourConstants.java:61
// residual stream orders beyond the predictor codes (used with stream_to_cell)
constexpr int kStreamLsopInit = 5;
constexpr int kStreamLsopInterior = 6;
constexpr int kStreamLsop8Init = 7;      // LSOP08: row 0 | row 1 (all columns) | columns 0,1 of every row from 2
constexpr int kStreamLsop8Interior = 8;  // LSOP08: rows 2.., columns 2..C-1

// ------------------------------------------------------------------------------------------------
// Tile view: a tile is a strided window of the row-major raster held in HBM.
// ------------------------------------------------------------------------------------------------
struct TileView {
  int32_t* base;   // cell (0,0)
  int64_t pitch;   // samples per raster row
  int R, C;
  __device__ __forceinline__ int32_t* row(int r) const { return base + int64_t(r) * pitch; }
  __device__ __forceinline__ int32_t& at(int r, int c) const { return base[int64_t(r) * pitch + c]; }
};

// A band as the kernels see it: the public descriptor plus, for tile-LIST calls (g4_encode_tile_list /
// g4_decode_tile_list), one {offset, pitch} pair per tile -- scattered tiles that share nothing but their size, e.g.
// the int[] buffers of a tile cache (gvrs/RasterTileCache.java:253-294) or a window inside a user's block.
struct BandEx : g4_band_desc {
  const int64_t* tileOffset = nullptr;  // samples from the raster base to tile t's cell (0,0); null: tile grid of the band
  const int64_t* tilePitch = nullptr;   // samples per row of the raster tile t lives in; null: grid_pitch
  BandEx() = default;
  BandEx(const g4_band_desc& b) : g4_band_desc(b) {}
};

__device__ __forceinline__ TileView tile_view(const BandEx& b, void* grid, int t) {
  TileView v;
  v.R = b.tile_rows;
  v.C = b.tile_cols;
  if (b.tileOffset) {
    v.base = static_cast<int32_t*>(grid) + b.tileOffset[t];
    v.pitch = b.tilePitch ? b.tilePitch[t] : b.grid_pitch;
    return v;
  }
  int tr = t / b.tiles_across, tc = t - tr * b.tiles_across;
  v.base = static_cast<int32_t*>(grid) + int64_t(tr) * b.tile_rows * b.grid_pitch + int64_t(tc) * b.tile_cols;
  v.pitch = b.grid_pitch;
  return v;
}

// ------------------------------------------------------------------------------------------------
// Block-wide exclusive scan of one uint32 per thread (kThreads threads).  `sm` = kWarps+1 words.
// Returns the exclusive prefix; *total receives the block sum.  Contains two __syncthreads().
// ------------------------------------------------------------------------------------------------
template <int NT = kThreads>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t x, uint32_t* sm, uint32_t* total) {
  constexpr int kWarps = NT / 32;  // sm holds NT/32 + 1 words
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = x;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += y;
  }
  __syncthreads();  // protect sm from the previous use
  if (lane == 31) sm[warp] = inc;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kWarps; w++) {
    uint32_t s = sm[w];
    if (w < warp) base += s;
    tot += s;
  }
  *total = tot;
  return base + inc - x;
}

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t x) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  return x;
}

// ------------------------------------------------------------------------------------------------
// M32 (CodecM32.java:257-311).  Returns the byte count; bytes packed little-end-first into *packed
// (byte j of the code at bits [8j, 8j+8)).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int m32_encode(int32_t value, uint64_t* packed) {
  if (value == INT32_MIN) { *packed = 0x80; return 1; }
  if (value > -127 && value < 127) { *packed = uint64_t(uint8_t(value)); return 1; }
  uint32_t a = value < 0 ? uint32_t(-value) : uint32_t(value);
  uint64_t p = value < 0 ? 0x81u : 0x7fu;
  if (a <= 254u) {
    p |= uint64_t(a - 127u) << 8;
    *packed = p; return 2;
  } else if (a <= 16638u) {
    uint32_t d = a - 255u;
    p |= uint64_t(((d >> 7) & 0x7f) | 0x80) << 8;
    p |= uint64_t(d & 0x7f) << 16;
    *packed = p; return 3;
  } else if (a <= 2113790u) {
    uint32_t d = a - 16639u;
    p |= uint64_t(((d >> 14) & 0x7f) | 0x80) << 8;
    p |= uint64_t(((d >> 7) & 0x7f) | 0x80) << 16;
    p |= uint64_t(d & 0x7f) << 24;
    *packed = p; return 4;
  } else if (a <= 270549246u) {
    uint32_t d = a - 2113791u;
    p |= uint64_t(((d >> 21) & 0x7f) | 0x80) << 8;
    p |= uint64_t(((d >> 14) & 0x7f) | 0x80) << 16;
    p |= uint64_t(((d >> 7) & 0x7f) | 0x80) << 24;
    p |= uint64_t(d & 0x7f) << 32;
    *packed = p; return 5;
  } else {
    uint32_t d = a - 270549247u;
    p |= uint64_t(((d >> 28) & 0x7f) | 0x80) << 8;
    p |= uint64_t(((d >> 21) & 0x7f) | 0x80) << 16;
    p |= uint64_t(((d >> 14) & 0x7f) | 0x80) << 24;
    p |= uint64_t(((d >> 7) & 0x7f) | 0x80) << 32;
    p |= uint64_t(d & 0x7f) << 40;
    *packed = p; return 6;
  }
}

__device__ __forceinline__ int m32_length(int32_t value) {
  if (value == INT32_MIN) return 1;
  if (value > -127 && value < 127) return 1;
  uint32_t a = value < 0 ? uint32_t(-value) : uint32_t(value);
  return a <= 254u ? 2 : a <= 16638u ? 3 : a <= 2113790u ? 4 : a <= 270549246u ? 5 : 6;
}

// M32 decode of the code that starts at buf[p] (CodecM32.java:327-356).  n = stream length; returns the
// code length in bytes, or 0 when the code runs past the end / has more than 5 payload bytes (the
// reference reads unchecked; the GPU path reports G4_ERR_FORMAT instead).
__device__ __forceinline__ int m32_decode_at(const uint8_t* buf, uint32_t p, uint32_t n, int32_t* out) {
  int symbol = int8_t(buf[p]);
  if (symbol == -128) { *out = INT32_MIN; return 1; }
  if (symbol > -127 && symbol < 127) { *out = symbol; return 1; }
  const uint32_t base[5] = {127u, 255u, 16639u, 2113791u, 270549247u};
  uint32_t delta = 0;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    if (p + 1 + i >= n) return 0;
    uint32_t s = buf[p + 1 + i];
    delta = (delta << 7) | (s & 0x7f);
    if ((s & 0x80) == 0) {
      *out = int32_t(symbol == -127 ? (0u - delta - base[i]) : (delta + base[i]));
      return i + 2;
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Predictor residual stream order and residuals.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stream_to_cell(int pred, int k, int R, int C, int* r, int* c) {
  if (pred == G4_PRED_DIFFERENCING) {
    int idx = k + 1;
    int rr = idx / C;
    *r = rr; *c = idx - rr * C;
  } else if (pred == G4_PRED_LINEAR) {
    if (k == 0) { *r = 0; *c = 1; }
    else if (k < 2 * R - 1) { int j = k - 1; *r = 1 + (j >> 1); *c = j & 1; }
    else { int j = k - (2 * R - 1); int rr = j / (C - 2); *r = rr; *c = 2 + j - rr * (C - 2); }
  } else if (pred == G4_PRED_TRIANGLE) {
    if (k < C - 1) { *r = 0; *c = k + 1; }
    else if (k < C + R - 2) { *r = k - (C - 1) + 1; *c = 0; }
    else { int j = k - (C + R - 2); int rr = j / (C - 1); *r = 1 + rr; *c = 1 + j - rr * (C - 1); }
  } else if (pred == kStreamLsopInit) {
    // LsOptimalPredictor12.java:143-209: row 0 | column 0 | row 1 | column 1 (rows 2..) | last two columns per row 2..
    if (k < C - 1) { *r = 0; *c = k + 1; return; }
    k -= C - 1;
    if (k < R - 1) { *r = k + 1; *c = 0; return; }
    k -= R - 1;
    if (k < C - 1) { *r = 1; *c = k + 1; return; }
    k -= C - 1;
    if (k < R - 2) { *r = k + 2; *c = 1; return; }
    k -= R - 2;
    *r = 2 + (k >> 1);
    *c = C - 2 + (k & 1);
  } else if (pred == kStreamLsop8Init) {
    // LsOptimalPredictor08.java:70-104: row 0 (C-1) | row 1 (C) | for r = 2..R-1: (r,0), (r,1)
    if (k < C - 1) { *r = 0; *c = k + 1; return; }
    k -= C - 1;
    if (k < C) { *r = 1; *c = k; return; }
    k -= C;
    *r = 2 + (k >> 1);
    *c = k & 1;
  } else if (pred == kStreamLsop8Interior) {
    int rr = k / (C - 2);
    *r = 2 + rr; *c = 2 + k - rr * (C - 2);
  } else if (pred == kStreamLsopInterior) {
    // LsOptimalPredictor12.java:254-282: rows 2.., columns 2..C-3, row-major
    int rr = k / (C - 4);
    *r = 2 + rr; *c = 2 + k - rr * (C - 4);
  } else {  // DifferencingWithNulls: one residual per cell
    int rr = k / C;
    *r = rr; *c = k - rr * C;
  }
}

// Residual of cell (r,c) != (0,0) for the three null-free predictors, read from the raster.
// PredictorModelDifferencing.java:112-142, PredictorModelLinear.java:104-143, PredictorModelTriangle.java:101-145
__device__ __forceinline__ int32_t residual_at(int pred, const TileView& t, int r, int c) {
  const int32_t* row = t.row(r);
  uint32_t v = uint32_t(row[c]);
  if (pred == G4_PRED_DIFFERENCING) {
    uint32_t prior = c > 0 ? uint32_t(row[c - 1]) : uint32_t(row[-t.pitch]);
    return int32_t(v - prior);
  } else if (pred == G4_PRED_LINEAR) {
    if (c == 0) return int32_t(v - uint32_t(row[-t.pitch]));
    if (c == 1) return int32_t(v - uint32_t(row[0]));
    return int32_t(v - (2u * uint32_t(row[c - 1]) - uint32_t(row[c - 2])));
  } else {
    if (r == 0) return int32_t(v - uint32_t(row[c - 1]));
    if (c == 0) return int32_t(v - uint32_t(row[-t.pitch]));
    const int32_t* up = row - t.pitch;
    return int32_t(v - (uint32_t(row[c - 1]) + uint32_t(up[c]) - uint32_t(up[c - 1])));
  }
}

// ------------------------------------------------------------------------------------------------
// LSB-first bit reader over a byte range that may start at any address.  Reads are done as aligned
// 32-bit words and clamped to the range, so nothing outside [begin, begin+len) is ever dereferenced
// beyond its containing aligned words ... the clamp substitutes zeros past the last word.
// ------------------------------------------------------------------------------------------------
struct BitSrc {
  const uint32_t* words;  // aligned base
  uint32_t bit0;          // bit offset of stream bit 0 relative to words[0]
  uint32_t lastWord;      // index of the last word that may be read
  uint32_t nBits;         // stream length in bits
  __device__ __forceinline__ void init(const uint8_t* begin, uint32_t lenBytes) {
    uintptr_t a = reinterpret_cast<uintptr_t>(begin);
    words = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    bit0 = uint32_t(a & 3) * 8u;
    nBits = lenBytes * 8u;
    lastWord = lenBytes ? uint32_t((bit0 + nBits - 1) >> 5) : 0;
  }
  // 32 bits starting at stream bit position `pos` (bits past the end read as zero)
  __device__ __forceinline__ uint32_t peek32(uint32_t pos) const {
    uint32_t a = bit0 + pos;
    uint32_t w = a >> 5;
    uint32_t lo = w <= lastWord ? __ldg(words + w) : 0u;
    uint32_t hi = w + 1 <= lastWord ? __ldg(words + w + 1) : 0u;
    uint32_t v = __funnelshift_r(lo, hi, a & 31);
    return v;
  }
  __device__ __forceinline__ uint32_t bits(uint32_t pos, int n) const {  // n in 1..32
    uint32_t v = peek32(pos);
    return n == 32 ? v : (v & ((1u << n) - 1u));
  }
};

// 64-bit register bit buffer over a global-memory BitSrc (clamped word reads); >= 32 valid bits after every skip().
// The word after the buffered ones is fetched one refill AHEAD (`ahead`), so that the load's latency is not on the
// decoding loop's dependent chain.
struct GlobalCursor {
  const BitSrc* src;
  uint32_t pos, next, ahead;
  uint64_t buf;
  int avail;
  __device__ __forceinline__ uint32_t word(uint32_t i) const { return i <= src->lastWord ? __ldg(src->words + i) : 0u; }
  __device__ __forceinline__ void init(const BitSrc& s, uint32_t p) {
    src = &s;
    pos = p;
    const uint32_t a = s.bit0 + p, i = a >> 5, sh = a & 31;
    buf = ((uint64_t(word(i + 1)) << 32) | word(i)) >> sh;
    avail = 64 - int(sh);
    next = i + 2;
    ahead = word(next);
  }
  __device__ __forceinline__ uint32_t peek() const { return uint32_t(buf); }
  __device__ __forceinline__ void skip(uint32_t n) {
    buf >>= n;
    avail -= int(n);
    pos += n;
    if (avail < 32) {
      buf |= uint64_t(ahead) << avail;
      avail += 32;
      ahead = word(++next);
    }
  }
};

// Single-thread LSB-first bit writer into a word buffer (used for headers / trees).
struct BitSink {
  uint32_t* words;
  uint32_t pos;
  __device__ __forceinline__ void put(uint32_t v, int n) {  // n in 1..32, v < 2^n
    uint32_t w = pos >> 5, o = pos & 31;
    words[w] |= v << o;
    if (o + n > 32) words[w + 1] |= v >> (32 - o);
    pos += n;
  }
};

__device__ __forceinline__ uint32_t load_le32(const uint8_t* p) {
  return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24);
}

}  // namespace g4
