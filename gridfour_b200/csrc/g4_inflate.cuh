// g4_inflate.cuh -- hand-written RFC 1950 / RFC 1951 decoder (zlib-wrapped DEFLATE), one warp per stream.
//
// Replaces java.util.zip.Inflater at the reference's call sites (paths under
// /root/reference/core/src/main/java/org/gridfour/): compress/CodecDeflate.java:139-149,
// compress/CodecFloat.java:285-298, lsop/LsDecoder12.java:126-145.  `new Inflater()` expects the zlib wrapper
// and verifies the Adler-32 trailer once the final block has been consumed; so does this decoder.
//
// A GROUP of G lanes (G = 8 or 32, template parameter) owns a stream: the group's first lane walks the bit stream
// (table-driven: 10-bit root table for literal/length codes, 8-bit for distances, canonical arithmetic for longer codes);
// all G lanes build the tables and perform the LZ77 copies.  With G = 8 a warp inflates FOUR streams side by side: the four
// walking lanes execute the symbol loop in the same instructions, which is where the time goes (the loop is a chain of
// dependent one-lane instructions), at the price of narrower copies and table builds.  Groups only ever synchronise among
// themselves (__syncwarp / __shfl_sync with the group's lane mask).
#pragma once
#include "g4_device.cuh"
#include "g4_canon.cuh"  // canon_build_tables / canon_slow_symbol (DEFLATE uses the same canonical convention)

namespace g4 {

constexpr int kInfLitBits = 10;
constexpr int kInfDistBits = 8;

struct InflateWarpShared {
  uint16_t litLut[1 << kInfLitBits];   // sym | len<<9 ; 0 -> slow path
  uint16_t distLut[1 << kInfDistBits]; // sym | len<<9
  uint16_t litSorted[288], distSorted[32];
  uint16_t litFirst[17], litCount[17], litOffset[17];
  uint16_t distFirst[17], distCount[17], distOffset[17];
  uint8_t lens[320];                   // code lengths being assembled (288 + 32)
  int ok;
};

enum { kInfOk = 0, kInfDataError = 1, kInfTruncated = 2 };

// RFC 1951 section 3.2.5 / 3.2.7 tables
static __constant__ uint16_t kInfLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115,
                                                131, 163, 195, 227, 258};
static __constant__ uint8_t kInfLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static __constant__ uint16_t kInfDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025,
                                                 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static __constant__ uint8_t kInfDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12,
                                                 13, 13};
static __constant__ uint8_t kInfClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

template <int G>
__device__ inline void inflate_build_lut(const uint16_t* first, const uint16_t* count, const uint16_t* offset,
                                         const uint16_t* sorted, uint16_t* lut, int lutBits, uint32_t gmask) {
  const int sub = threadIdx.x & (G - 1);
  for (int e = sub; e < (1 << lutBits); e += G) {
    uint32_t v = __brev(uint32_t(e));
    uint16_t entry = 0;
    for (int len = 1; len <= lutBits; len++) {
      uint32_t code = v >> (32 - len);
      uint32_t d = code - first[len];
      if (d < count[len]) { entry = uint16_t(sorted[offset[len] + d] | (len << 9)); break; }
    }
    lut[e] = entry;
  }
  __syncwarp(gmask);
}

__device__ __forceinline__ int inflate_symbol(const uint16_t* lut, int lutBits, const uint16_t* first, const uint16_t* count,
                                              const uint16_t* offset, const uint16_t* sorted, const BitSrc& src, uint32_t* pos) {
  uint32_t v = src.peek32(*pos);
  uint32_t e = lut[v & ((1u << lutBits) - 1)];
  if (e) { *pos += e >> 9; return int(e & 0x1ffu); }
  return canon_slow_symbol(first, count, offset, sorted, src, pos, lutBits + 1);
}

// Adler-32 of out[0..n) by one group of G lanes.
template <int G>
__device__ inline uint32_t adler32_group(const uint8_t* out, uint32_t n, uint32_t gmask) {
  const int lane = threadIdx.x & (G - 1);
  // s1 = 1 + sum b_i ; s2 = n + sum (n - i) * b_i   (mod 65521), i = 0..n-1
  unsigned long long a = 0, b = 0;
  for (uint32_t i = lane; i < n; i += G) {
    uint32_t x = out[i];
    a += x;
    b += (unsigned long long)(n - i) * x;  // n <= 6.3e6: no 64-bit overflow
  }
  a %= 65521ull;
  b %= 65521ull;
#pragma unroll
  for (int d = G / 2; d >= 1; d >>= 1) {
    a += __shfl_xor_sync(gmask, a, d);
    b += __shfl_xor_sync(gmask, b, d);
  }
  uint32_t s1 = uint32_t((1ull + a) % 65521ull);
  uint32_t s2 = uint32_t((uint64_t(n % 65521u) + b) % 65521ull);
  return (s2 << 16) | s1;
}

// Inflates one zlib stream with a group of G lanes (every lane of the group calls with the group's arguments).  Output stops at `cap` bytes (Inflater.inflate(byte[]) with an
// exactly sized buffer); the end-of-block code and the trailer are still consumed when they follow directly,
// as zlib does.  *produced / *consumed as Inflater.inflate()'s return value / getBytesRead().
template <int G>
__device__ inline int inflate_group(InflateWarpShared& S, const uint8_t* in, uint32_t inLen, uint8_t* out, uint32_t cap,
                                    uint32_t* produced, uint32_t* consumed) {
  const int lane = threadIdx.x & (G - 1);  // lane inside the group
  const uint32_t gmask = G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
  *produced = 0;
  *consumed = 0;
  if (inLen < 2) return kInfTruncated;
  {
    uint32_t cmf = in[0], flg = in[1];
    if ((cmf & 0x0f) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31u != 0 || (flg & 0x20)) return kInfDataError;
  }
  BitSrc src;
  src.init(in + 2, inLen - 2);
  uint32_t pos = 0, op = 0;
  int status = kInfOk;
  bool last = false;
  bool stopped = false;  // output buffer full before the final end-of-block code
  while (!last && status == kInfOk) {
    if (pos + 3 > src.nBits) { status = kInfTruncated; break; }
    last = src.bits(pos, 1) != 0;
    const uint32_t type = src.bits(pos + 1, 2);
    pos += 3;
    if (type == 0) {  // stored
      pos = (pos + 7u) & ~7u;
      if (pos + 32 > src.nBits) { status = kInfTruncated; break; }
      uint32_t len = src.bits(pos, 16), nlen = src.bits(pos + 16, 16);
      pos += 32;
      if ((len ^ 0xffffu) != nlen) { status = kInfDataError; break; }
      if (pos + len * 8 > src.nBits) { status = kInfTruncated; break; }
      uint32_t take = len < cap - op ? len : cap - op;
      const uint8_t* sp = in + 2 + (pos >> 3);
      for (uint32_t i = lane; i < take; i += G) out[op + i] = sp[i];
      __syncwarp(gmask);
      op += take;
      pos += take * 8;
      if (take < len) { stopped = true; break; }  // output full
      continue;
    }
    if (type == 3) { status = kInfDataError; break; }
    // ---- code lengths -----------------------------------------------------------------------------------
    int nLit = 288, nDist = 30;
    if (type == 1) {
      for (int i = lane; i < 288; i += G) S.lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
      for (int i = lane; i < 32; i += G) S.lens[288 + i] = i < 30 ? 5 : 0;
      __syncwarp(gmask);
    } else {
      int ok = 1;
      if (lane == 0) {
        if (pos + 14 > src.nBits) ok = 0;
        nLit = int(src.bits(pos, 5)) + 257;
        nDist = int(src.bits(pos + 5, 5)) + 1;
        int nCl = int(src.bits(pos + 10, 4)) + 4;
        pos += 14;
        if (nLit > 286 || nDist > 30) ok = 0;
        uint8_t cl[19];
        for (int i = 0; i < 19; i++) cl[i] = 0;
        for (int i = 0; i < nCl; i++) { cl[kInfClOrder[i]] = uint8_t(src.bits(pos, 3)); pos += 3; }
        uint16_t fc[17], cn[17], of[17], so[19];
        if (ok && !canon_build_tables(cl, 19, fc, cn, of, so)) ok = 0;
        int minLen = 1;
        while (minLen < 7 && cn[minLen] == 0) minLen++;
        int i = 0, prev = 0;
        const int total = nLit + nDist;
        while (ok && i < total) {
          if (pos >= src.nBits) { ok = 0; break; }
          int sym = canon_slow_symbol(fc, cn, of, so, src, &pos, minLen);
          if (sym < 0) { ok = 0; break; }
          if (sym < 16) { S.lens[i++] = uint8_t(sym); prev = sym; continue; }
          int rep, val = 0;
          if (sym == 16) { if (i == 0) { ok = 0; break; } rep = 3 + int(src.bits(pos, 2)); pos += 2; val = prev; }
          else if (sym == 17) { rep = 3 + int(src.bits(pos, 3)); pos += 3; prev = 0; }
          else { rep = 11 + int(src.bits(pos, 7)); pos += 7; prev = 0; }
          if (i + rep > total) { ok = 0; break; }
          for (int k = 0; k < rep; k++) S.lens[i++] = uint8_t(val);
        }
        if (ok) {
          // move the distance lengths to their fixed offset (288) and clear the gaps
          uint8_t tmp[32];
          for (int k = 0; k < 32; k++) tmp[k] = k < nDist ? S.lens[nLit + k] : 0;
          for (int k = nLit; k < 288; k++) S.lens[k] = 0;
          for (int k = 0; k < 32; k++) S.lens[288 + k] = tmp[k];
          if (S.lens[256] == 0) ok = 0;  // no end-of-block code
        }
        S.ok = ok;
      }
      __syncwarp(gmask);
      pos = __shfl_sync(gmask, pos, 0, G);
      if (!S.ok) { status = kInfDataError; break; }
      __syncwarp(gmask);  // every lane has read S.ok before lane 0 rewrites it below
    }
    if (lane == 0) {
      int ok = canon_build_tables(S.lens, 288, S.litFirst, S.litCount, S.litOffset, S.litSorted) ? 1 : 0;
      bool anyDist = false;
      for (int k = 0; k < 32; k++) anyDist |= S.lens[288 + k] != 0;
      if (anyDist) { if (!canon_build_tables(S.lens + 288, 32, S.distFirst, S.distCount, S.distOffset, S.distSorted)) ok = 0; }
      else for (int l = 0; l <= 16; l++) { S.distFirst[l] = 0; S.distCount[l] = 0; S.distOffset[l] = 0; }
      S.ok = ok;
    }
    __syncwarp(gmask);
    if (!S.ok) { status = kInfDataError; break; }
    inflate_build_lut<G>(S.litFirst, S.litCount, S.litOffset, S.litSorted, S.litLut, kInfLitBits, gmask);
    inflate_build_lut<G>(S.distFirst, S.distCount, S.distOffset, S.distSorted, S.distLut, kInfDistBits, gmask);
    // ---- symbols ------------------------------------------------------------------------------------------
    bool full = false;
    // 64-bit register bit buffer of the walking lane: one LUT read per symbol, a global word load every 32 consumed bits.
    // It lives across the matches of a block (re-initialising it after every match put two dependent global loads on
    // the critical path of every match).
    GlobalCursor cur;
    if (lane == 0) cur.init(src, pos);
    for (;;) {
      // lane 0 decodes literals until it meets a match, the end of block, or an error
      int ev = 0;  // 1 = match, 2 = end of block, 3 = error/truncated, 4 = output full
      uint32_t mlen = 0, mdist = 0;
      if (lane == 0) {
        for (;;) {
          // Two or three literals per turn while the stream, the bit buffer and the output have room for them: both codes come
          // from the LUT (at most kInfLitBits bits each, the buffer holds 32 or more), one skip for the pair, no
          // per-symbol limit checks.  Anything else -- a long code, a length code, the end of block, the last bits
          // of the stream -- leaves the loop for the general code below.
          while (cur.pos + 32u <= src.nBits && op + 2u <= cap) {
            const uint32_t b = cur.peek();
            const uint32_t e1 = S.litLut[b & ((1u << kInfLitBits) - 1)];
            const uint32_t l1 = e1 >> 9;
            if (e1 == 0u || (e1 & 0x100u)) break;
            const uint32_t e2 = S.litLut[(b >> l1) & ((1u << kInfLitBits) - 1)];
            if (e2 == 0u || (e2 & 0x100u)) {
              out[op++] = uint8_t(e1);
              cur.skip(l1);
              break;
            }
            const uint32_t l2 = l1 + (e2 >> 9);
            const uint32_t e3 = S.litLut[(b >> l2) & ((1u << kInfLitBits) - 1)];
            out[op] = uint8_t(e1);
            out[op + 1] = uint8_t(e2);
            if (e3 == 0u || (e3 & 0x100u) || op + 3u > cap) {
              op += 2;
              cur.skip(l2);
              continue;
            }
            out[op + 2] = uint8_t(e3);
            op += 3;
            cur.skip(l2 + (e3 >> 9));
          }
          if (cur.pos >= src.nBits) { ev = 3; break; }
          const uint32_t p0 = cur.pos;
          int sym;
          {
            const uint32_t e = S.litLut[cur.peek() & ((1u << kInfLitBits) - 1)];
            if (e) { cur.skip(e >> 9); sym = int(e & 0x1ffu); }
            else {
              uint32_t p = cur.pos;
              sym = canon_slow_symbol(S.litFirst, S.litCount, S.litOffset, S.litSorted, src, &p, kInfLitBits + 1);
              if (sym >= 0) cur.init(src, p);
            }
          }
          if (sym < 0) { ev = 3; break; }
          if (sym < 256) {
            if (op >= cap) { cur.pos = p0; ev = 4; break; }
            out[op++] = uint8_t(sym);
            continue;
          }
          if (sym == 256) { ev = 2; break; }
          if (sym > 285) { ev = 3; break; }
          if (op >= cap) { cur.pos = p0; ev = 4; break; }
          int li = sym - 257;
          const uint32_t le = kInfLenExtra[li];
          mlen = kInfLenBase[li] + (le ? (cur.peek() & ((1u << le) - 1u)) : 0u);
          cur.skip(le);
          int ds;
          {
            const uint32_t e = S.distLut[cur.peek() & ((1u << kInfDistBits) - 1)];
            if (e) { cur.skip(e >> 9); ds = int(e & 0x1ffu); }
            else {
              uint32_t p = cur.pos;
              ds = canon_slow_symbol(S.distFirst, S.distCount, S.distOffset, S.distSorted, src, &p, kInfDistBits + 1);
              if (ds >= 0) cur.init(src, p);
            }
          }
          if (ds < 0 || ds > 29) { ev = 3; break; }
          const uint32_t de = kInfDistExtra[ds];
          mdist = kInfDistBase[ds] + (de ? (cur.peek() & ((1u << de) - 1u)) : 0u);
          cur.skip(de);
          if (mdist > op) { ev = 3; break; }  // distance too far back
          ev = 1;
          break;
        }
        pos = cur.pos;
      }
      ev = __shfl_sync(gmask, ev, 0, G);
      pos = __shfl_sync(gmask, pos, 0, G);
      op = __shfl_sync(gmask, op, 0, G);
      if (ev == 1) {
        mlen = __shfl_sync(gmask, mlen, 0, G);
        mdist = __shfl_sync(gmask, mdist, 0, G);
        uint32_t take = mlen < cap - op ? mlen : cap - op;
        __syncwarp(gmask);
        // overlapping copies repeat the last `mdist` bytes: source index wraps modulo the distance
        if (mdist >= take) {
          for (uint32_t i = lane; i < take; i += G) out[op + i] = out[op - mdist + i];
        } else {
          for (uint32_t i = lane; i < take; i += G) out[op + i] = out[op - mdist + (i % mdist)];
        }
        __syncwarp(gmask);
        op += take;
        if (take < mlen) { full = true; break; }
        continue;
      }
      if (ev == 2) break;
      if (ev == 4) { full = true; break; }
      status = pos >= src.nBits ? kInfTruncated : kInfDataError;
      break;
    }
    if (full) { stopped = true; break; }
  }
  __syncwarp(gmask);
  *produced = op;
  uint32_t usedBytes = 2 + ((pos + 7) >> 3);
  if (status == kInfOk && last && !stopped) {
    // the final block was consumed: verify the Adler-32 trailer (big-endian) like Inflater does.  A missing
    // trailer is not an error for Inflater.inflate() (it just waits for more input), so it is not one here.
    uint32_t tp = (pos + 7) >> 3;
    if (2 + tp + 4 <= inLen) {
      const uint8_t* tr = in + 2 + tp;
      uint32_t want = (uint32_t(tr[0]) << 24) | (uint32_t(tr[1]) << 16) | (uint32_t(tr[2]) << 8) | uint32_t(tr[3]);
      uint32_t got = adler32_group<G>(out, op, gmask);
      if (want != got) status = kInfDataError;
      usedBytes += 4;
    } else {
      usedBytes = inLen;
    }
  }
  *consumed = usedBytes;
  return status;
}

// One warp, one stream (the callers that inflate inside a larger kernel).
__device__ inline int inflate_warp(InflateWarpShared& S, const uint8_t* in, uint32_t inLen, uint8_t* out, uint32_t cap,
                                   uint32_t* produced, uint32_t* consumed) {
  return inflate_group<32>(S, in, inLen, out, cap, produced, consumed);
}

}  // namespace g4
