// g4_canon_fast.cuh -- canonical Huffman text decoder, fast path: the packing is staged in shared memory and every
// thread decodes from a 64-bit register bit buffer.
//
// Same stream format and the same self-synchronising sub-sequence scheme as g4_canon.cuh (reference:
// compress/canonicalHuffman/CanonicalHuffman.java:441-519, CanonHuffTreeDecoder.java:68-177, LengthEncoder.java:197-236
// under /root/reference/core/src/main/java/org/gridfour/).  What changes is where the bits come from and how often
// they are touched:
//   * the whole packing (typically ~19 KB for a 180x240 tile) is copied once, coalesced, into shared memory,
//   * a symbol costs one shared-memory LUT read and two 64-bit shifts; the buffer is refilled a word at a time,
//   * the serial table parse uses an 8-bit LUT for the 20-symbol code-table code,
//   * pass 0 only needs each sub-sequence's END, so it starts a fixed distance before the limit instead of decoding the
//     whole sub-sequence (a wrong guess is repaired by the later passes; the result never depends on it),
//   * values leave through a sink that is handed runs of consecutive value indices (begin / put / end), so the caller
//     can keep a running cell address instead of dividing per value.
// Packings that do not fit the staging buffer fall back to g4_canon.cuh.
#pragma once
#include "g4_canon.cuh"

namespace g4 {

constexpr int kFastStageWords = 7168;  // 28 KB of packing (5.3 bits/sample for a 180x240 tile)
constexpr int kFastMaxSub = 1280;
#ifndef G4_FAST_SUBBITS
#define G4_FAST_SUBBITS 160
#endif
#ifndef G4_FAST_LOOKBACK
#define G4_FAST_LOOKBACK 48
#endif
constexpr uint32_t kFastSubBits = G4_FAST_SUBBITS;    // target sub-sequence size
constexpr uint32_t kFastLookback = G4_FAST_LOOKBACK;  // pass 0 starts this many bits before the limit
constexpr int kFastLutBits = 11;
constexpr uint32_t kFastSpecial = 0x8000u;  // LUT flag: symbol >= 256 (null, escapes, end of text)

struct CanonFastShared {
  uint32_t mlut[1 << kFastLutBits];   // up to 3 plain values per lookup: s1 | s2 << 8 | s3 << 16 | bits << 24 | n << 28
  uint16_t lut[1 << kFastLutBits];    // sym | len << 9 | special; 0 = code longer than the LUT
  uint16_t sorted[kCanonSymbols];
  uint16_t firstCode[17], count[17], offset[17];
  uint8_t lens[kCanonSymbols + 4];
  uint16_t ctLut[256];                // code-table code: sym | len << 8; 0 = longer than 8 bits
  uint32_t endpos[kFastMaxSub];
  uint32_t startv[kFastMaxSub];       // start position endpos[i] was computed from
  uint16_t cnt[kFastMaxSub];
  uint8_t eot[kFastMaxSub];
  uint32_t scan[33];
  uint32_t textStart, endBit;
  int error, changed, firstEot;
  // staged packing, word 0 = packing bytes 0..3.  LAST member: a kernel may allocate more dynamic shared memory than
  // sizeof(CanonFastShared) and stage packings of up to canon_fast_stage_words(allocated bytes) words.
  uint32_t sw[kFastStageWords + 8];
};
__host__ __device__ constexpr size_t canon_fast_smem_bytes(uint32_t stageWords) {
  return sizeof(CanonFastShared) + (stageWords > uint32_t(kFastStageWords) ? size_t(stageWords - kFastStageWords) * 4 : 0);
}

// Bit source over the staged words (absolute bit positions inside the packing).
struct SmemBitSrc {
  const uint32_t* w;
  uint32_t nBits;
  __device__ __forceinline__ uint32_t peek32(uint32_t pos) const {
    uint32_t i = pos >> 5;
    return __funnelshift_r(w[i], w[i + 1], pos & 31);  // the staging buffer is zero padded past the data
  }
  __device__ __forceinline__ uint32_t bits(uint32_t pos, int n) const {
    uint32_t v = peek32(pos);
    return n == 32 ? v : (v & ((1u << n) - 1u));
  }
};

// Register bit window over S.sw: two staged words (64 bits) of which `used` < 32 are consumed, so peek() always has 32
// valid bits.  The position is kept RELATIVE to a limit (rem = limit - position), which is what the decoding loops
// compare against; pos(limit) recovers the absolute position.  The staged words are addressed through the
// shared-memory struct itself so that the loads stay LDS with an immediate base.
struct BitCursor {
  uint32_t lo, hi, used, next;
  int rem;
  __device__ __forceinline__ void init(const CanonFastShared& S, uint32_t p, uint32_t limit) {
    const uint32_t i = p >> 5;
    lo = S.sw[i];
    hi = S.sw[i + 1];
    used = p & 31u;
    next = i + 2;
    rem = int(limit) - int(p);
  }
  __device__ __forceinline__ uint32_t peek() const { return __funnelshift_r(lo, hi, used); }
  __device__ __forceinline__ uint32_t pos(uint32_t limit) const { return uint32_t(int(limit) - rem); }
  __device__ __forceinline__ void skip(const CanonFastShared& S, uint32_t n) {  // n <= 32
    used += n;
    rem -= int(n);
    if (used >= 32u) {
      lo = hi;
      hi = S.sw[next++];
      used -= 32u;
    }
  }
};

// Copies the packing into S.sw (coalesced 32-bit loads; `packing` is 4-byte aligned) and zero pads.  All threads call.
__device__ inline void canon_fast_stage(CanonFastShared& S, const uint8_t* packing, uint32_t len) {
  const uint32_t nWords = (len + 3) >> 2;
  const uint32_t* g = reinterpret_cast<const uint32_t*>(packing);
  for (uint32_t i = threadIdx.x; i < nWords; i += kThreads) S.sw[i] = __ldg(g + i);
  if (threadIdx.x < 8) S.sw[nWords + threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x == 0 && (len & 3)) S.sw[nWords - 1] &= (1u << (8 * (len & 3))) - 1u;  // bytes past the packing read as zero
  __syncthreads();
}

// LengthEncoder.readEncodedLengths (:197-236) + CanonHuffTreeDecoder.decodeTree (:131-177).  One thread.
__device__ inline void canon_fast_parse_header(CanonFastShared& S, const SmemBitSrc& src, uint32_t startBit) {
  S.error = 0;
  uint32_t pos = startBit + 1;  // reserved bit
  uint8_t ctLens[20];
  {
    int k = 0, prior = 0;
    while (k < 20) {
      if (pos + 5 > src.nBits) { S.error = 1; return; }
      int index = int(src.bits(pos, 5));
      pos += 5;
      int n = 1, val = index;
      if (index <= 15) prior = index;
      else if (index == 16) { n = int(src.bits(pos, 2)) + 3; pos += 2; val = prior; }
      else if (index == 17) { n = int(src.bits(pos, 3)) + 3; pos += 3; val = 0; prior = 0; }
      else if (index == 18) { n = int(src.bits(pos, 7)) + 11; pos += 7; val = 0; prior = 0; }
      else continue;  // reference ignores other values
      if (k + n > 20) { S.error = 1; return; }
      for (int i = 0; i < n; i++) ctLens[k++] = uint8_t(val);
    }
  }
  uint16_t fc[17], cn[17], of[17], so[20];
  if (!canon_build_tables(ctLens, 20, fc, cn, of, so)) { S.error = 1; return; }
  int minLen = 1;
  while (minLen < 15 && cn[minLen] == 0) minLen++;
  // 8-bit LUT of the code-table code: a code of length l (MSB first in the LSB-first stream) owns every index whose
  // low l bits are the bit-reversed code
  for (int i = 0; i < 256; i++) S.ctLut[i] = 0;
  for (int l = 1; l <= 8; l++)
    for (int d = 0; d < cn[l]; d++) {
      uint32_t rev = __brev(uint32_t(fc[l] + d)) >> (32 - l);
      uint16_t entry = uint16_t(so[of[l] + d] | (l << 8));
      for (uint32_t hi = 0; hi < (1u << (8 - l)); hi++) S.ctLut[rev | (hi << l)] = entry;
    }
  int prior = 0;
  for (int i = 0; i < kCanonSymbols; i++) S.lens[i] = 0;
  for (int i = 0; i < kCanonSymbols; i++) {
    if (pos >= src.nBits) { S.error = 1; return; }
    int test;
    uint32_t e = S.ctLut[src.peek32(pos) & 0xffu];
    if (e) { test = int(e & 0xffu); pos += e >> 8; }
    else test = canon_slow_symbol(fc, cn, of, so, src, &pos, minLen > 9 ? minLen : 9);
    if (test < 0) { S.error = 1; return; }
    if (test <= 15) { S.lens[i] = uint8_t(test); prior = test; }
    else {
      int n, val = 0;
      if (test == 16) { n = int(src.bits(pos, 2)) + 3; pos += 2; val = prior; }
      else if (test == 17) { n = int(src.bits(pos, 3)) + 3; pos += 3; prior = 0; }
      else if (test == 18) { n = int(src.bits(pos, 7)) + 11; pos += 7; prior = 0; }
      else continue;  // the code table's own end-of-text symbol: leaves a zero length
      if (i + n > kCanonSymbols) { S.error = 1; return; }
      for (int j = 0; j < n; j++) S.lens[i + j] = uint8_t(val);
      i += n - 1;
    }
  }
  if (!canon_build_tables(S.lens, kCanonSymbols, S.firstCode, S.count, S.offset, S.sorted)) { S.error = 1; return; }
  if (S.lens[kSymEot] == 0) { S.error = 1; return; }
  S.textStart = pos;
}

// The uncommon symbols: special symbols (null, escapes, end of text) and codes longer than the LUT.  `e` is the LUT
// entry read at position p0.  Returns the symbol (0..259) and the position after its code, or -1 for an invalid code.
__device__ __forceinline__ int canon_fast_rare_symbol(const CanonFastShared& S, uint32_t e, uint32_t p0, uint32_t nBits, uint32_t* after) {
  if (e) {
    *after = p0 + ((e >> 9) & 15u);
    return int(e & 0x1ffu);
  }
  SmemBitSrc src{S.sw, nBits};
  uint32_t p = p0;
  int sym = canon_slow_symbol(S.firstCode, S.count, S.offset, S.sorted, src, &p, kFastLutBits + 1);
  *after = p;
  return sym;
}

// Counting decode of one sub-sequence: from `start` to the first value boundary at or after `limit` (<= nBits).
// flag: 1 = end of text consumed, 2 = invalid code / ran past the data.
__device__ __forceinline__ void canon_fast_count(const CanonFastShared& S, uint32_t nBits, uint32_t start, uint32_t limit,
                                                 uint32_t* endOut, uint32_t* cntOut, int* flagOut) {
  BitCursor cur;
  cur.init(S, start, limit);
  uint32_t c = 0, end;
  int flag = 0;
  for (;;) {
    if (cur.rem >= kFastLutBits) {
      // the whole 11-bit window lies before the limit: every value coded inside it is consumed (up to 3 per lookup)
      const uint32_t m = S.mlut[cur.peek() & ((1u << kFastLutBits) - 1u)];
      if (m >> 28) {
        cur.skip(S, (m >> 24) & 15u);
        c += m >> 28;
        continue;
      }
    }
    const uint32_t e = S.lut[cur.peek() & ((1u << kFastLutBits) - 1u)];
    if (e - 1u < 0x7fffu) {  // LUT hit on a plain value: e = sym | len << 9
      if (cur.rem <= 0) { end = cur.pos(limit); break; }
      cur.skip(S, e >> 9);
      c++;
      continue;
    }
    const uint32_t p0 = cur.pos(limit);
    uint32_t after;
    const int sym = canon_fast_rare_symbol(S, e, p0, nBits, &after);
    if (sym < 0) { flag = 2; end = p0; break; }
    if (sym == kSymEsc2 || sym == kSymEsc8) {
      after += sym == kSymEsc2 ? 2u : 8u;
      if (after > nBits) { flag = 2; end = p0; break; }
    } else {
      if (p0 >= limit) { end = p0; break; }
      if (sym == kSymEot) { flag = 1; end = after; break; }
      c++;  // long-code value or null symbol
    }
    cur.init(S, after, limit);
  }
  *endOut = end;
  *cntOut = c;
  *flagOut = flag;
}

// Parallel construction of firstCode / count / offset / sorted from S.lens (CanonHuffTreeDecoder.java:68-129 builds
// the equivalent tree).  Sets S.error for an over-subscribed or empty code.  All threads call.
template <int NT = kThreads>
__device__ inline void canon_fast_tables_cta(CanonFastShared& S) {
  __shared__ uint32_t cnt32[17];
  const int tid = threadIdx.x;
  if (tid < 17) cnt32[tid] = 0;
  __syncthreads();
  for (int i = tid; i < kCanonSymbols; i += NT) {
    int l = S.lens[i];
    if (l > 15) S.error = 1;
    else if (l) atomicAdd(&cnt32[l], 1u);
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t code = 0, off = 0;
    S.count[0] = 0; S.firstCode[0] = 0; S.offset[0] = 0;
    S.count[16] = 0; S.firstCode[16] = 0; S.offset[16] = 0;
    for (int l = 1; l <= 15; l++) {
      uint32_t c = cnt32[l];
      S.firstCode[l] = uint16_t(code);
      S.offset[l] = uint16_t(off);
      S.count[l] = uint16_t(c);
      if (code + c > (1u << l)) S.error = 1;  // over-subscribed
      code = (code + c) << 1;
      off += c;
    }
    if (off == 0 || S.lens[kSymEot] == 0) S.error = 1;
  }
  __syncthreads();
  if (tid < 32) {  // sorted[]: symbols by (length, symbol); one warp, 32 symbols per round, ranks by __match_any_sync
    __shared__ uint32_t nextSlot[17];
    if (tid < 17) nextSlot[tid] = S.offset[tid];
    __syncwarp();
    for (int i0 = 0; i0 < kCanonSymbols; i0 += 32) {
      const int i = i0 + tid;
      int l = i < kCanonSymbols ? S.lens[i] : 0;
      if (l > 15) l = 0;
      const uint32_t same = __match_any_sync(0xffffffffu, l);
      const uint32_t before = same & ((1u << tid) - 1u);
      if (l) S.sorted[nextSlot[l] + __popc(before)] = uint16_t(i);
      __syncwarp();
      if (l && before == 0u) nextSlot[l] += __popc(same);
      __syncwarp();
    }
  }
  __syncthreads();
}

// 11-bit lookup table: thread per prefix, canonical arithmetic on the bit-reversed prefix.  All threads call.
template <int NT = kThreads>
__device__ inline void canon_fast_build_lut(CanonFastShared& S) {
  for (int e = threadIdx.x; e < (1 << kFastLutBits); e += NT) {
    uint32_t v = __brev(uint32_t(e));
    uint16_t entry = 0;
    for (int len = 1; len <= kFastLutBits; len++) {
      uint32_t code = v >> (32 - len);
      uint32_t d = code - S.firstCode[len];
      if (d < S.count[len]) {
        uint32_t sym = S.sorted[S.offset[len] + d];
        entry = uint16_t(sym | (uint32_t(len) << 9) | (sym >= 256 ? kFastSpecial : 0u));
        break;
      }
    }
    S.lut[e] = entry;
  }
  __syncthreads();
  // multi-symbol table: the plain values whose codes lie completely inside the 11-bit window (at most 3)
  for (int e = threadIdx.x; e < (1 << kFastLutBits); e += NT) {
    uint32_t used = 0, n = 0, syms = 0;
    while (n < 3) {
      const uint32_t x = S.lut[(uint32_t(e) >> used) & ((1u << kFastLutBits) - 1u)];
      const uint32_t len = x >> 9;
      if (x - 1u >= 0x7fffu || used + len > uint32_t(kFastLutBits)) break;  // long code, special symbol, or cut by the window
      syms |= (x & 0xffu) << (8 * n);
      used += len;
      n++;
    }
    S.mlut[e] = syms | (used << 24) | (n << 28);
  }
  __syncthreads();
}

// Decodes the TEXT of one canonical stream staged in S.sw; tables and LUT are ready and the text starts at bit T0.
// *endBit are absolute bit positions inside the packing, nBits = 8 * packing length.  The sink receives runs:
// begin(firstValueIndex), put(value) ..., end().  hintBits bounds the region searched first (0 = everything).
// All threads call.
template <class Sink, int NT = kThreads, uint32_t SUBBITS = kFastSubBits>
__device__ bool canon_fast_decode_text(CanonFastShared& S, uint32_t nBits, const uint32_t T0, uint32_t maxValues, uint32_t hintBits,
                                       Sink sink, uint32_t* endBit, uint32_t* nValues) {
  constexpr int kThreads = NT;                       // shadows the global: every loop below strides by the CTA size
  constexpr int kFastRounds = kFastMaxSub / NT;      // sub-sequences per thread (5 with 256 threads, 2 with 512)
  const int tid = threadIdx.x;
  uint32_t regionEnd = nBits;
  if (hintBits && T0 + hintBits < regionEnd) regionEnd = T0 + hintBits;
  for (;;) {  // region growth until the end-of-text code is inside the region
    const uint32_t avail = regionEnd - T0;
    // sub-sequence size: about kFastSubBits, adjusted so that the sub-sequences fill whole rounds of kThreads threads
    uint32_t rounds = (avail / SUBBITS + kThreads - 1) / kThreads;
    if (rounds < 1u) rounds = 1u;
    if (rounds > uint32_t(kFastRounds)) rounds = kFastRounds;
    uint32_t B = (avail + rounds * kThreads - 1) / (rounds * kThreads);
    if (B < 96u) B = 96u;
    const int nSub = int((avail + B - 1) / B);
    // pass 0: only the END of every sub-sequence matters here, so start kFastLookback bits before the limit and rely on
    // self-synchronisation; sub-sequence 0 starts at the true text start.
#pragma unroll 1
    for (int i = tid; i < nSub; i += kThreads) {
      uint32_t limit = T0 + uint32_t(i + 1) * B;
      if (limit > regionEnd) limit = regionEnd;
      uint32_t from = T0 + uint32_t(i) * B;
      if (i > 0 && limit - from > kFastLookback) from = limit - kFastLookback;
      uint32_t e, c;
      int f;
      canon_fast_count(S, nBits, from, limit, &e, &c, &f);
      S.endpos[i] = e;
      S.cnt[i] = uint16_t(c);
      S.eot[i] = uint8_t(f);
      S.startv[i] = i > 0 ? 0xffffffffu : T0;  // forces the exact decode of every later sub-sequence in the first pass below
    }
    // synchronisation passes: sub-sequence i must start where i-1 ended.  Reads of endpos[i-1] may see this pass's or
    // the previous pass's value (both are candidates); the loop ends only after a pass in which nothing was rewritten,
    // and in that pass every start was compared against final values.
    volatile uint32_t* vend = S.endpos;
    __syncthreads();
    for (int pass = 0; pass <= nSub; pass++) {
      bool any = false;
#pragma unroll 1
      for (int i = tid; i < nSub; i += kThreads) {
        if (i == 0) continue;
        const uint32_t ns = vend[i - 1];
        if (ns != S.startv[i]) {
          S.startv[i] = ns;
          uint32_t limit = T0 + uint32_t(i + 1) * B;
          if (limit > regionEnd) limit = regionEnd;
          uint32_t e, c;
          int f;
          canon_fast_count(S, nBits, ns, limit, &e, &c, &f);
          vend[i] = e;
          S.cnt[i] = uint16_t(c);
          S.eot[i] = uint8_t(f);
          any = true;
        }
      }
      if (!__syncthreads_or(any ? 1 : 0)) break;  // one barrier per pass: it also orders this pass's writes before the next pass's reads
    }
    __syncthreads();
    if (tid == 0) S.firstEot = nSub;
    __syncthreads();
#pragma unroll 1
    for (int i = tid; i < nSub; i += kThreads)
      if (S.eot[i]) atomicMin(&S.firstEot, i);
    __syncthreads();
    const int fe = S.firstEot;
    if (fe < nSub && S.eot[fe] == 2) return false;  // invalid code, or the data ended before end-of-text
    if (fe == nSub) {                               // no end-of-text inside the region: widen it
      if (regionEnd >= nBits) return false;
      uint32_t grown = (regionEnd - T0) * 4u;
      regionEnd = (grown > nBits - T0) ? nBits : T0 + grown;
      __syncthreads();
      continue;
    }
    // value offsets: thread tid owns sub-sequences tid*kFastRounds .. +kFastRounds-1 for the scan
    uint32_t mySum = 0;
#pragma unroll
    for (int j = 0; j < kFastRounds; j++) {
      int i = tid * kFastRounds + j;
      mySum += (i <= fe && i < nSub) ? S.cnt[i] : 0u;
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan<NT>(mySum, S.scan, &total);
    if (total > maxValues) return false;
    if (tid == 0) S.endBit = S.endpos[fe];
    __syncthreads();
    uint32_t* offv = S.endpos;  // end positions are no longer needed: the array now holds the first value index of every sub-sequence
    {
      uint32_t run = ex;
#pragma unroll
      for (int j = 0; j < kFastRounds; j++) {
        int i = tid * kFastRounds + j;
        if (i < nSub) {
          offv[i] = run;
          run += (i <= fe) ? S.cnt[i] : 0u;
        }
      }
    }
    __syncthreads();
    // write pass: decode again, assembling escapes into values.  Every thread takes `rounds` CONSECUTIVE sub-sequences
    // as one run (they are contiguous in the stream and in the value order), so the sink sees one long run per thread
    // instead of one short run per sub-sequence.
    bool bad = false;
    const int last = fe < nSub - 1 ? fe : nSub - 1;
    const int i0 = tid * int(rounds);
    if constexpr (Sink::kPacked) {
      // Packed form of the write pass: the sink takes the (up to three) symbol bytes of a lookup in one call and keeps
      // them in a register byte queue, so the common path has no per-value work at all.
      if (i0 <= last) {
        const int i1 = i0 + int(rounds) - 1 < last ? i0 + int(rounds) - 1 : last;
        uint32_t limit = T0 + uint32_t(i1 + 1) * B;
        if (limit > regionEnd) limit = regionEnd;
        BitCursor cur;
        cur.init(S, S.startv[i0], limit);
        sink.begin(offv[i0]);
        bool have = false;
        for (;;) {
          if (cur.rem >= kFastLutBits) {
            const uint32_t m = S.mlut[cur.peek() & ((1u << kFastLutBits) - 1u)];
            const uint32_t n = m >> 28;
            if (n) {
              cur.skip(S, (m >> 24) & 15u);
              sink.push(m & 0xffffffu, int(n));
              have = true;
              continue;
            }
          }
          const uint32_t e = S.lut[cur.peek() & ((1u << kFastLutBits) - 1u)];
          if (e - 1u < 0x7fffu) {  // plain value
            if (cur.rem <= 0) break;
            cur.skip(S, e >> 9);
            sink.push(e & 0xffu, 1);
            have = true;
            continue;
          }
          const uint32_t p0 = cur.pos(limit);
          uint32_t after;
          const int sym = canon_fast_rare_symbol(S, e, p0, nBits, &after);
          if (sym < 0) { bad = true; break; }
          if (sym == kSymEsc2 || sym == kSymEsc8) {
            if (!have) { bad = true; break; }  // an escape with nothing to extend (CanonicalHuffman.java:495-504 would index -1)
            const int nb = sym == kSymEsc2 ? 2 : 8;
            if (after + nb > nBits) { bad = true; break; }
            SmemBitSrc src{S.sw, nBits};
            sink.amend(nb, src.bits(after, nb));
            after += nb;
          } else {
            if (p0 >= limit) break;
            if (sym == kSymEot) break;
            if (sym == kSymNull) sink.put_rare(INT32_MIN);
            else sink.push(uint32_t(sym), 1);  // a byte value with a code longer than the LUT
            have = true;
          }
          cur.init(S, after, limit);
        }
        sink.end();
      }
    } else if (i0 <= last) {
      const int i1 = i0 + int(rounds) - 1 < last ? i0 + int(rounds) - 1 : last;
      uint32_t limit = T0 + uint32_t(i1 + 1) * B;
      if (limit > regionEnd) limit = regionEnd;
      BitCursor cur;
      cur.init(S, S.startv[i0], limit);
      sink.begin(offv[i0]);
      bool have = false;
      uint32_t v = 0;
      for (;;) {
        const uint32_t p0 = cur.pos(limit);
        if (cur.rem >= kFastLutBits) {
          const uint32_t m = S.mlut[cur.peek() & ((1u << kFastLutBits) - 1u)];
          const uint32_t n = m >> 28;
          if (n) {
            cur.skip(S, (m >> 24) & 15u);
            if (have) sink.put(int32_t(v));
            have = true;
            if (n == 1) v = (m & 0xffu) - 128u;
            else {
              sink.put(int32_t((m & 0xffu) - 128u));
              if (n == 2) v = ((m >> 8) & 0xffu) - 128u;
              else {
                sink.put(int32_t(((m >> 8) & 0xffu) - 128u));
                v = ((m >> 16) & 0xffu) - 128u;
              }
            }
            continue;
          }
        }
        const uint32_t e = S.lut[cur.peek() & ((1u << kFastLutBits) - 1u)];
        if (e - 1u < 0x7fffu) {  // plain value
          if (p0 >= limit) break;
          cur.skip(S, e >> 9);
          if (have) sink.put(int32_t(v));
          have = true;
          v = (e & 0xffu) - 128u;
          continue;
        }
        uint32_t after;
        const int sym = canon_fast_rare_symbol(S, e, p0, nBits, &after);
        if (sym < 0) { bad = true; break; }
        if (sym == kSymEsc2 || sym == kSymEsc8) {
          if (!have) { bad = true; break; }  // an escape with nothing to extend (CanonicalHuffman.java:495-504 would index -1)
          const int nb = sym == kSymEsc2 ? 2 : 8;
          if (after + nb > nBits) { bad = true; break; }
          SmemBitSrc src{S.sw, nBits};
          v = (v << nb) | src.bits(after, nb);
          after += nb;
        } else {
          if (p0 >= limit) break;
          if (sym == kSymEot) break;
          if (have) sink.put(int32_t(v));
          have = true;
          v = sym == kSymNull ? uint32_t(INT32_MIN) : uint32_t(sym - 128);
        }
        cur.init(S, after, limit);
      }
      if (have) sink.put(int32_t(v));
      sink.end();
    }
    if (__syncthreads_or(bad ? 1 : 0)) return false;
    *endBit = S.endBit;
    *nValues = total;
    return true;
  }
}

// Run sink for any residual stream order: one stream_to_cell per value (begin(first value index), put(value) ..., end()).
struct CellRunSink {
  static constexpr bool kPacked = false;
  TileView t;
  int order;
  uint32_t k;
  __device__ __forceinline__ void begin(uint32_t k0) { k = k0; }
  __device__ __forceinline__ void put(int32_t v) {
    int r, c;
    stream_to_cell(order, int(k++), t.R, t.C, &r, &c);
    t.at(r, c) = v;
  }
  __device__ __forceinline__ void end() {}
};

// Header + text of one canonical stream staged in S.sw (startBit absolute).  All threads call.
template <class Sink>
__device__ bool canon_fast_decode_stream(CanonFastShared& S, uint32_t nBits, uint32_t startBit, uint32_t maxValues, uint32_t hintBits,
                                         Sink sink, uint32_t* endBit, uint32_t* nValues) {
  __syncthreads();
  if (threadIdx.x == 0) {
    SmemBitSrc src{S.sw, nBits};
    canon_fast_parse_header(S, src, startBit);
  }
  __syncthreads();
  if (S.error) return false;
  canon_fast_build_lut(S);
  return canon_fast_decode_text(S, nBits, S.textStart, maxValues, hintBits, sink, endBit, nValues);
}

}  // namespace g4
