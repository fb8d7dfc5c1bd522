// g4_lsop_common.cuh -- pieces shared by the LSOP12 decode kernels (g4_lsop.cu: general path and encoder;
// g4_lsop_fast.cu: the byte hand-over / TMA-fed wavefront path).
//
// Reference (under /root/reference/core/src/main/java/org/gridfour/): lsop/LsHeader.java:104-189,
// compress/canonicalHuffman/CanonicalHuffman.java:441-519, CanonHuffTreeDecoder.java:68-177, LengthEncoder.java:197-236.
#pragma once
#include "g4_device.cuh"
#include "g4_canon.cuh"
#include "g4_canon_fast.cuh"

namespace g4 {

// StrictMath.round(float): floor(a + 1/2) evaluated on the bit pattern (java.lang.Math.round(float), JDK >= 8):
//   shift = 149 - biasedExp;  0 <= shift < 32 ? ((+-significand >> shift) + 1) >> 1 : (int) a
// For shift >= 32 (|a| < 2^-9) the cast gives 0, and so does the shifted form with the shift clamped to 31
// (the 24-bit significand shifts out completely: (0 + 1) >> 1 == 0, (-1 + 1) >> 1 == 0), which leaves one rare
// branch for shift < 0 (|a| >= 2^24, infinities, NaN: saturating cast, NaN -> 0).
// Integer form on purpose: floorf + __float2int_rd (FRND.FLOOR / F2I.FLOOR) measured 10x slower for the whole wavefront
// kernel on sm_100a (15.3 ms instead of 1.5 ms), see profiles/README.md.
__device__ __forceinline__ int32_t java_round(float a) {
  const int32_t bits = __float_as_int(a);
  const int shift = 149 - ((bits >> 23) & 0xff);
  if (shift < 0) return __float2int_rz(a);
  int32_t r = (bits & 0x007FFFFF) | 0x00800000;
  r = bits < 0 ? -r : r;
  return ((r >> min(shift, 31)) + 1) >> 1;
}

struct LsHeaderInfo {
  int type;          // 0 legacy Huffman, 1 Deflate, 2 canonical Huffman
  int32_t seed;
  uint32_t nInitCodes, nInteriorCodes;
  uint32_t headerSize;
  bool ok;
  bool hasChecksum;        // LsHeader.valueChecksumIncluded (bit 7 of the type byte)
  uint32_t valueChecksum;  // CRC-32C of the tile's values as little-endian int32 (LsHeader.computeChecksum :391-406)
};

// LsHeader(byte[],int) (LsHeader.java:104-189).  Coefficients are written to coef[0..11].
__device__ inline LsHeaderInfo parse_ls_header(const uint8_t* p, uint32_t len, float* coef, bool writeCoef, int nCoef = 12) {
  LsHeaderInfo h;
  h.ok = false;
  h.type = -1;
  h.seed = 0;
  h.nInitCodes = h.nInteriorCodes = 0;
  h.headerSize = 0;
  h.hasChecksum = false;
  h.valueChecksum = 0;
  if (len < uint32_t(3 + 4 + 4 * nCoef)) return h;
  uint32_t off = 1;
  bool legacy = (p[1] & 0x40) == 0;
  bool cks;
  if (legacy) {
    if (p[off++] != nCoef) return h;
  } else {
    h.type = p[off] & 0x0f;
    cks = (p[off] & 0x80) != 0;
    off++;
    if (p[off++] != nCoef) return h;
  }
  h.seed = int32_t(load_le32(p + off));
  off += 4;
  for (int i = 0; i < nCoef; i++) {
    if (writeCoef) coef[i] = __uint_as_float(load_le32(p + off));
    off += 4;
  }
  if (legacy) {
    if (off + 9 > len) return h;
    h.nInitCodes = load_le32(p + off);
    h.nInteriorCodes = load_le32(p + off + 4);
    off += 8;
    h.type = p[off] & 0x0f;
    cks = (p[off] & 0x80) != 0;
    off++;
  } else if (h.type != 2) {
    if (off + 8 > len) return h;
    h.nInitCodes = load_le32(p + off);
    h.nInteriorCodes = load_le32(p + off + 4);
    off += 8;
  }
  if (cks) {
    if (off + 4 > len) return h;
    h.hasChecksum = true;
    h.valueChecksum = load_le32(p + off);
    off += 4;
  }
  if (off > len || h.type < 0 || h.type > 2) return h;
  h.headerSize = off;
  h.ok = true;
  return h;
}

// per-tile record handed from the head kernels to the text kernels: [0..3] interior text start (absolute bit, 0 = tile not
// on the fast path), [8..267] code lengths, and (fast path) the decoding tables built from them: [272..791] sorted[260],
// [792..825] firstCode[17], [826..859] count[17], [860..893] offset[17] (uint16 each)
constexpr int kLsopMetaBytes = 896;
constexpr int kLsopMetaSorted = 272, kLsopMetaFirst = 792, kLsopMetaCount = 826, kLsopMetaOffset = 860;

struct CanonWarpShared {
  uint8_t lens[kCanonSymbols + 4];
  uint16_t sorted[kCanonSymbols];
  uint16_t firstCode[17], count[17], offset[17];
  uint16_t ctFirst[17], ctCount[17], ctOffset[17], ctSorted[20];
  uint16_t lut8[256];   // text code: sym | len << 9 | special; 0 = longer than 8 bits
  uint16_t ctLut[256];  // code-table code: sym | len << 8
  uint32_t cnt32[17], next[17];
  uint32_t textStart;
  int error;
};

// Warp-cooperative version of canon_fast_parse_header over a global-memory bit source.  All 32 lanes call.
__device__ inline void canon_warp_parse_header(CanonWarpShared& W, const BitSrc& src, uint32_t startBit) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  if (lane == 0) {
    W.error = 0;
    uint32_t pos = startBit + 1;  // reserved bit
    int k = 0, prior = 0;
    while (k < 20 && !W.error) {
      if (pos + 5 > src.nBits) { W.error = 1; break; }
      int index = int(src.bits(pos, 5));
      pos += 5;
      int n = 1, val = index;
      if (index <= 15) prior = index;
      else if (index == 16) { n = int(src.bits(pos, 2)) + 3; pos += 2; val = prior; }
      else if (index == 17) { n = int(src.bits(pos, 3)) + 3; pos += 3; val = 0; prior = 0; }
      else if (index == 18) { n = int(src.bits(pos, 7)) + 11; pos += 7; val = 0; prior = 0; }
      else continue;  // reference ignores other values
      if (k + n > 20) { W.error = 1; break; }
      for (int i = 0; i < n; i++) W.lens[k++] = uint8_t(val);
    }
    if (!W.error && !canon_build_tables(W.lens, 20, W.ctFirst, W.ctCount, W.ctOffset, W.ctSorted)) W.error = 1;
    W.textStart = pos;
  }
  __syncwarp();
  if (W.error) return;
  for (int e = lane; e < 256; e += 32) {  // 8-bit LUT of the code-table code
    uint32_t v = __brev(uint32_t(e));
    uint16_t entry = 0;
    for (int len = 1; len <= 8; len++) {
      uint32_t d = (v >> (32 - len)) - W.ctFirst[len];
      if (d < W.ctCount[len]) { entry = uint16_t(W.ctSorted[W.ctOffset[len] + d] | (len << 8)); break; }
    }
    W.ctLut[e] = entry;
  }
  __syncwarp();
  if (lane == 0) {  // the 260 text code lengths: serial by nature (variable-length codes)
    GlobalCursor cur;
    cur.init(src, W.textStart);
    int prior = 0;
    for (int i = 0; i < kCanonSymbols; i++) W.lens[i] = 0;
    for (int i = 0; i < kCanonSymbols; i++) {
      if (cur.pos >= src.nBits) { W.error = 1; break; }
      int test;
      uint32_t e = W.ctLut[cur.peek() & 0xffu];
      if (e) { test = int(e & 0xffu); cur.skip(e >> 8); }
      else {
        uint32_t p = cur.pos;
        test = canon_slow_symbol(W.ctFirst, W.ctCount, W.ctOffset, W.ctSorted, src, &p, 9);
        if (test >= 0) cur.init(src, p);
      }
      if (test < 0) { W.error = 1; break; }
      if (test <= 15) { W.lens[i] = uint8_t(test); prior = test; }
      else {
        int n, val = 0;
        if (test == 16) { n = int(cur.peek() & 3u) + 3; cur.skip(2); val = prior; }
        else if (test == 17) { n = int(cur.peek() & 7u) + 3; cur.skip(3); prior = 0; }
        else if (test == 18) { n = int(cur.peek() & 127u) + 11; cur.skip(7); prior = 0; }
        else continue;  // the code table's own end-of-text symbol: leaves a zero length
        if (i + n > kCanonSymbols) { W.error = 1; break; }
        for (int j = 0; j < n; j++) W.lens[i + j] = uint8_t(val);
        i += n - 1;
      }
    }
    if (W.lens[kSymEot] == 0) W.error = 1;
    W.textStart = cur.pos;
  }
  if (lane < 17) W.cnt32[lane] = 0;
  __syncwarp();
  if (W.error) return;
  // tables of the text code, in parallel
  for (int i = lane; i < kCanonSymbols; i += 32) {
    int l = W.lens[i];
    if (l) atomicAdd(&W.cnt32[l], 1u);
  }
  __syncwarp();
  if (lane == 0) {
    uint32_t code = 0, off = 0;
    W.count[0] = 0; W.firstCode[0] = 0; W.offset[0] = 0; W.next[0] = 0;
    W.count[16] = 0; W.firstCode[16] = 0; W.offset[16] = 0; W.next[16] = 0;
    for (int l = 1; l <= 15; l++) {
      uint32_t c = W.cnt32[l];
      W.firstCode[l] = uint16_t(code);
      W.offset[l] = uint16_t(off);
      W.count[l] = uint16_t(c);
      W.next[l] = off;
      if (code + c > (1u << l)) W.error = 1;  // over-subscribed
      code = (code + c) << 1;
      off += c;
    }
  }
  __syncwarp();
  if (W.error) return;
  for (int i0 = 0; i0 < kCanonSymbols; i0 += 32) {  // sorted[]: symbols by (length, symbol)
    const int i = i0 + lane;
    const int l = i < kCanonSymbols ? W.lens[i] : 0;
    const uint32_t same = __match_any_sync(0xffffffffu, l);
    if (l) {
      const uint32_t rank = __popc(same & ((1u << lane) - 1u));
      W.sorted[W.next[l] + rank] = uint16_t(i);
    }
    __syncwarp();
    if (l && (same & ((1u << lane) - 1u)) == 0) W.next[l] += __popc(same);
    __syncwarp();
  }
  for (int e = lane; e < 256; e += 32) {  // 8-bit LUT of the text code
    uint32_t v = __brev(uint32_t(e));
    uint16_t entry = 0;
    for (int len = 1; len <= 8; len++) {
      uint32_t d = (v >> (32 - len)) - W.firstCode[len];
      if (d < W.count[len]) {
        uint32_t sym = W.sorted[W.offset[len] + d];
        entry = uint16_t(sym | (uint32_t(len) << 9) | (sym >= 256 ? kFastSpecial : 0u));
        break;
      }
    }
    W.lut8[e] = entry;
  }
  __syncwarp();
}

// One symbol for the warp-level decoder: 8-bit LUT, canonical arithmetic for longer codes.  Returns the symbol and the
// position after its code, or -1.
__device__ __forceinline__ int canon_warp_symbol(const CanonWarpShared& W, const BitSrc& src, GlobalCursor& cur) {
  const uint32_t e = W.lut8[cur.peek() & 0xffu];
  if (e) {
    cur.skip((e >> 9) & 15u);
    return int(e & 0x1ffu);
  }
  uint32_t p = cur.pos;
  const int sym = canon_slow_symbol(W.firstCode, W.count, W.offset, W.sorted, src, &p, 9);
  if (sym >= 0) cur.init(src, p);
  return sym;
}

// Counting decode of one lane's sub-sequence (same rules as canon_fast_count): from `start` to the first value
// boundary at or after `limit` (<= src.nBits).  flag: 1 = end of text consumed, 2 = invalid code / ran past the data.
__device__ inline void canon_warp_count(const CanonWarpShared& W, const BitSrc& src, uint32_t start, uint32_t limit, uint32_t* endOut,
                                        uint32_t* cntOut, int* flagOut) {
  GlobalCursor cur;
  cur.init(src, start);
  uint32_t c = 0, end;
  int flag = 0;
  for (;;) {
    const uint32_t p0 = cur.pos;
    const int sym = canon_warp_symbol(W, src, cur);
    if (sym < 0) { flag = 2; end = p0; break; }
    if (sym == kSymEsc2 || sym == kSymEsc8) {
      cur.skip(sym == kSymEsc2 ? 2u : 8u);
      if (cur.pos > src.nBits) { flag = 2; end = p0; break; }
      continue;
    }
    if (p0 >= limit) { end = p0; break; }
    if (sym == kSymEot) { flag = 1; end = cur.pos; break; }
    c++;
  }
  *endOut = end;
  *cntOut = c;
  *flagOut = flag;
}

// Warp-level version of the self-synchronising sub-sequence decoder (g4_canon_fast.cuh) for SHORT texts: 32 lanes,
// one sub-sequence of kWarpSubBits bits per lane and region; regions follow each other until the end-of-text code is
// found.  emit(valueIndex, value) is called once per value by the lane that decoded it.  All 32 lanes call.
constexpr uint32_t kWarpSubBits = 192;  // > 84 bits, the longest value (code + three escapes), so a sub-sequence never overshoots the next one
template <class Emit>
__device__ inline bool canon_warp_decode_text(const CanonWarpShared& W, const BitSrc& src, uint32_t T0, uint32_t maxValues, Emit emit,
                                              uint32_t* endBit, uint32_t* nValues) {
  const int lane = threadIdx.x & 31;
  uint32_t kBase = 0, regionStart = T0;
  for (;;) {
    if (regionStart >= src.nBits) return false;  // data ended before end-of-text
    uint32_t regionEnd = regionStart + 32u * kWarpSubBits;
    if (regionEnd > src.nBits) regionEnd = src.nBits;
    const uint32_t lo = regionStart + uint32_t(lane) * kWarpSubBits;
    const bool has = lo < regionEnd;
    uint32_t limit = lo + kWarpSubBits;
    if (limit > regionEnd) limit = regionEnd;
    uint32_t start = lane == 0 ? regionStart : 0xffffffffu, end = regionEnd, cnt = 0;
    int flag = 0;
    if (has) {  // pass 0: only the end matters, start 48 bits before the limit and rely on self-synchronisation
      uint32_t from = lo;
      if (lane > 0 && limit - lo > 48u) from = limit - 48u;
      canon_warp_count(W, src, from, limit, &end, &cnt, &flag);
    }
    for (int pass = 0; pass < 34; pass++) {  // lane i starts where lane i-1 ended; converges in <= 32 passes
      uint32_t ns = __shfl_up_sync(0xffffffffu, end, 1);
      if (lane == 0) ns = regionStart;
      const bool ch = has && ns != start;
      if (ch) {
        start = ns;
        canon_warp_count(W, src, start, limit, &end, &cnt, &flag);
      }
      if (__ballot_sync(0xffffffffu, ch) == 0) break;
    }
    const uint32_t flagged = __ballot_sync(0xffffffffu, has && flag != 0);
    const int fe = flagged ? __ffs(flagged) - 1 : 32;
    if (fe < 32 && __shfl_sync(0xffffffffu, flag, fe) == 2) return false;
    const bool mine = has && lane <= fe;
    const uint32_t c = mine ? cnt : 0u;
    const uint32_t inc = warp_inclusive_scan(c);
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    if (kBase + total > maxValues) return false;
    bool bad = false;
    if (mine) {  // write pass: decode again, assembling escapes into values
      uint32_t k = kBase + inc - c;
      GlobalCursor cur;
      cur.init(src, start);
      bool have = false;
      uint32_t v = 0;
      for (;;) {
        const uint32_t p0 = cur.pos;
        const int sym = canon_warp_symbol(W, src, cur);
        if (sym < 0) { bad = true; break; }
        if (sym == kSymEsc2 || sym == kSymEsc8) {
          if (!have) { bad = true; break; }  // an escape with nothing to extend
          const uint32_t nb = sym == kSymEsc2 ? 2u : 8u;
          v = (v << nb) | (cur.peek() & ((1u << nb) - 1u));
          cur.skip(nb);
          continue;
        }
        if (p0 >= limit || sym == kSymEot) break;
        if (have) emit(k++, int32_t(v));
        have = true;
        v = sym == kSymNull ? uint32_t(INT32_MIN) : uint32_t(sym - 128);
      }
      if (have) emit(k, int32_t(v));
    }
    if (__ballot_sync(0xffffffffu, bad)) return false;
    kBase += total;
    if (fe < 32) {
      *endBit = __shfl_sync(0xffffffffu, end, fe);
      *nValues = kBase;
      return true;
    }
    const uint32_t hasMask = __ballot_sync(0xffffffffu, has);
    regionStart = __shfl_sync(0xffffffffu, end, 31 - __clz(hasMask));  // the last sub-sequence's end
  }
}

}  // namespace g4
