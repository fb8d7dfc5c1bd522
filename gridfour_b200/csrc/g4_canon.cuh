// g4_canon.cuh -- canonical Huffman coder over the 260-symbol integer alphabet, CTA-cooperative.
//
// Reference (under /root/reference/core/src/main/java/org/gridfour/compress/canonicalHuffman/):
//   CanonicalHuffman.java:177-283 (encode), :352-418 (countSymbols), :441-519 (decode, decodeText)
//   TreeBuilder.java:75-318, PackageMerge.java:91-175, LengthEncoder.java:86-236, HuffmanCodeBits.java:47-74
//   CanonHuffTreeDecoder.java:68-177
// Stream: 1 reserved bit, 20 raw 5-bit code-table lengths (run coded), 260 text lengths coded with the
// 20-symbol table, text (codes MSB-first inside the LSB-first bit stream; escapes 258 +2 raw bits,
// 257 +8 raw bits), end-of-text code 259.
#pragma once
#include "g4_device.cuh"

namespace g4 {

constexpr int kCanonSymbols = 260;
constexpr int kSymNull = 256, kSymEsc8 = 257, kSymEsc2 = 258, kSymEot = 259;
constexpr int kCanonLutBits = 11;
constexpr int kCanonMaxSub = 2048;
constexpr int kCanonSubPerThread = kCanonMaxSub / kThreads;

struct CanonDecShared {
  uint16_t lut[1 << kCanonLutBits];  // sym | len<<9 (len 1..11); 0 = needs the slow path
  uint16_t sorted[kCanonSymbols];    // symbols ordered by (length, symbol)
  uint16_t firstCode[17], count[17], offset[17];
  uint8_t lens[kCanonSymbols + 4];
  uint32_t endpos[kCanonMaxSub];
  uint16_t cnt[kCanonMaxSub];
  uint8_t eot[kCanonMaxSub];
  uint32_t off[kCanonMaxSub];
  uint32_t scan[kWarps + 1];
  uint32_t textStart;
  int error;
  int changed;
  int firstEot;
};

// ---- serial table parse (one thread) ------------------------------------------------------------------
// canonical decode of one symbol by arithmetic on (firstCode,count,offset); codes are MSB-first
template <class Src>
__device__ inline int canon_slow_symbol(const uint16_t* firstCode, const uint16_t* count, const uint16_t* offset,
                                        const uint16_t* sorted, const Src& src, uint32_t* pos, int minLen) {
  uint32_t v = __brev(src.peek32(*pos));  // first stream bit is now the MSB
  for (int len = minLen; len <= 15; len++) {
    uint32_t code = v >> (32 - len);
    uint32_t d = code - firstCode[len];
    if (d < count[len]) { *pos += len; return sorted[offset[len] + d]; }
  }
  return -1;
}

// Builds firstCode/count/offset/sorted for `nSym` lengths.  Returns false for an unusable code.
__device__ inline bool canon_build_tables(const uint8_t* lens, int nSym, uint16_t* firstCode, uint16_t* count, uint16_t* offset,
                                          uint16_t* sorted) {
  for (int l = 0; l <= 16; l++) count[l] = 0;
  int used = 0;
  for (int i = 0; i < nSym; i++)
    if (lens[i]) { count[lens[i]]++; used++; }
  if (used == 0) return false;
  uint32_t code = 0, off = 0;
  for (int l = 1; l <= 15; l++) {
    firstCode[l] = uint16_t(code);
    offset[l] = uint16_t(off);
    if (code + count[l] > (1u << l)) return false;  // over-subscribed
    code = (code + count[l]) << 1;
    off += count[l];
  }
  uint16_t next[17];
  for (int l = 0; l <= 16; l++) next[l] = offset[l];
  for (int i = 0; i < nSym; i++)
    if (lens[i]) sorted[next[lens[i]]++] = uint16_t(i);
  return true;
}

// LengthEncoder.readEncodedLengths (:197-236) + CanonHuffTreeDecoder.decodeTree (:131-177).  One thread.
__device__ inline void canon_parse_header(CanonDecShared& S, const BitSrc& src, uint32_t startBit) {
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
  int prior = 0;
  for (int i = 0; i < kCanonSymbols; i++) S.lens[i] = 0;
  for (int i = 0; i < kCanonSymbols; i++) {
    if (pos >= src.nBits) { S.error = 1; return; }
    int test = canon_slow_symbol(fc, cn, of, so, src, &pos, minLen);
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

__device__ __forceinline__ int canon_decode_symbol(const CanonDecShared& S, const BitSrc& src, uint32_t* pos) {
  uint32_t v = src.peek32(*pos);
  uint32_t e = S.lut[v & ((1u << kCanonLutBits) - 1)];
  if (e) { *pos += e >> 9; return int(e & 0x1ffu); }
  return canon_slow_symbol(S.firstCode, S.count, S.offset, S.sorted, src, pos, kCanonLutBits + 1);
}

// One sub-sequence, counting only.  Stops at the first value boundary at or after `limit`, at end-of-text,
// or when the data runs out.  flag: 1 = EOT consumed, 2 = invalid code / ran out.
__device__ __forceinline__ void canon_count_subseq(const CanonDecShared& S, const BitSrc& src, uint32_t start, uint32_t limit,
                                                   uint32_t* endOut, uint32_t* cntOut, int* flagOut) {
  uint32_t pos = start, c = 0;
  int flag = 0;
  for (;;) {
    uint32_t p0 = pos;
    if (p0 >= src.nBits) { flag = 2; break; }
    int sym = canon_decode_symbol(S, src, &pos);
    if (sym < 0) { flag = 2; pos = p0; break; }
    bool esc = sym == kSymEsc2 || sym == kSymEsc8;
    if (p0 >= limit && !esc) { pos = p0; break; }
    if (sym == kSymEot) { flag = 1; break; }
    if (esc) pos += sym == kSymEsc2 ? 2u : 8u;
    else c++;
  }
  *endOut = pos;
  *cntOut = c;
  *flagOut = flag;
}

// Decodes one canonical stream that starts at bit `startBit` of `src`.  sink(valueIndex, value) is called
// once per decoded value (any thread).  hintBits bounds the region searched first (0 = everything).
// On success *endBit = bit after the end-of-text code and *nValues = number of values.  All threads call.
template <class Sink>
__device__ bool canon_decode_stream(CanonDecShared& S, const BitSrc& src, uint32_t startBit, uint32_t maxValues, uint32_t hintBits,
                                    Sink sink, uint32_t* endBit, uint32_t* nValues) {
  const int tid = threadIdx.x;
  __syncthreads();
  if (tid == 0) canon_parse_header(S, src, startBit);
  __syncthreads();
  if (S.error) return false;
  // lookup table: thread per prefix, canonical arithmetic on the bit-reversed prefix
  for (int e = tid; e < (1 << kCanonLutBits); e += kThreads) {
    uint32_t v = __brev(uint32_t(e));
    uint16_t entry = 0;
    for (int len = 1; len <= kCanonLutBits; len++) {
      uint32_t code = v >> (32 - len);
      uint32_t d = code - S.firstCode[len];
      if (d < S.count[len]) { entry = uint16_t(S.sorted[S.offset[len] + d] | (len << 9)); break; }
    }
    S.lut[e] = entry;
  }
  __syncthreads();
  const uint32_t T0 = S.textStart;
  uint32_t regionEnd = src.nBits;
  if (hintBits && T0 + hintBits < regionEnd) regionEnd = T0 + hintBits;
  for (;;) {  // region doubling until the end-of-text code is inside the region
    const uint32_t avail = regionEnd - T0;
    uint32_t B = (avail + kCanonMaxSub - 1) / kCanonMaxSub;
    B = (B + 31u) & ~31u;
    if (B < 128u) B = 128u;
    const int nSub = int((avail + B - 1) / B);
    uint32_t myStart[kCanonSubPerThread];
#pragma unroll
    for (int j = 0; j < kCanonSubPerThread; j++) {
      int i = tid + j * kThreads;
      myStart[j] = T0 + uint32_t(i) * B;
      if (i < nSub) {
        uint32_t limit = T0 + uint32_t(i + 1) * B;
        if (limit > regionEnd) limit = regionEnd;
        uint32_t e, c;
        int f;
        canon_count_subseq(S, src, myStart[j], limit, &e, &c, &f);
        S.endpos[i] = e;
        S.cnt[i] = uint16_t(c);
        S.eot[i] = uint8_t(f);
      }
    }
    for (int pass = 0; pass < nSub; pass++) {
      __syncthreads();
      if (tid == 0) S.changed = 0;
      uint32_t ns[kCanonSubPerThread];
#pragma unroll
      for (int j = 0; j < kCanonSubPerThread; j++) {
        int i = tid + j * kThreads;
        ns[j] = (i > 0 && i < nSub) ? S.endpos[i - 1] : myStart[j];
      }
      __syncthreads();
      bool any = false;
#pragma unroll
      for (int j = 0; j < kCanonSubPerThread; j++) {
        int i = tid + j * kThreads;
        if (i < nSub && ns[j] != myStart[j]) {
          myStart[j] = ns[j];
          uint32_t limit = T0 + uint32_t(i + 1) * B;
          if (limit > regionEnd) limit = regionEnd;
          uint32_t e, c;
          int f;
          canon_count_subseq(S, src, myStart[j], limit, &e, &c, &f);
          S.endpos[i] = e;
          S.cnt[i] = uint16_t(c);
          S.eot[i] = uint8_t(f);
          any = true;
        }
      }
      if (any) S.changed = 1;
      __syncthreads();
      if (!S.changed) break;
    }
    // first sub-sequence of the synchronised chain that reached end-of-text (or failed)
    __syncthreads();
    if (tid == 0) S.firstEot = nSub;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kCanonSubPerThread; j++) {
      int i = tid + j * kThreads;
      if (i < nSub && S.eot[i]) atomicMin(&S.firstEot, i);
    }
    __syncthreads();
    const int fe = S.firstEot;
    if (fe < nSub && S.eot[fe] == 2) return false;  // invalid code, or the data ended before end-of-text
    if (fe == nSub) {                               // no end-of-text inside the region: widen it
      if (regionEnd >= src.nBits) return false;
      uint32_t grown = (regionEnd - T0) * 4u;
      regionEnd = (grown > src.nBits - T0) ? src.nBits : T0 + grown;
      __syncthreads();
      continue;
    }
    // value offsets (contiguous ownership for the scan)
    uint32_t local[kCanonSubPerThread];
    uint32_t mySum = 0;
#pragma unroll
    for (int j = 0; j < kCanonSubPerThread; j++) {
      int i = tid * kCanonSubPerThread + j;
      local[j] = (i <= fe) ? S.cnt[i] : 0u;
      mySum += local[j];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(mySum, S.scan, &total);
    if (total > maxValues) return false;
    {
      uint32_t run = ex;
#pragma unroll
      for (int j = 0; j < kCanonSubPerThread; j++) {
        int i = tid * kCanonSubPerThread + j;
        if (i < nSub) S.off[i] = run;
        run += local[j];
      }
    }
    __syncthreads();
    // write pass: decode again, assembling escapes into values
    bool bad = false;
#pragma unroll
    for (int j = 0; j < kCanonSubPerThread; j++) {
      int i = tid + j * kThreads;
      if (i <= fe && i < nSub) {
        uint32_t pos = myStart[j];
        uint32_t limit = T0 + uint32_t(i + 1) * B;
        if (limit > regionEnd) limit = regionEnd;
        uint32_t o = S.off[i];
        bool have = false;
        uint32_t cur = 0;
        for (;;) {
          uint32_t p0 = pos;
          if (p0 >= src.nBits) { bad = true; break; }
          int sym = canon_decode_symbol(S, src, &pos);
          if (sym < 0) { bad = true; break; }
          bool esc = sym == kSymEsc2 || sym == kSymEsc8;
          if (p0 >= limit && !esc) break;
          if (sym == kSymEot) break;
          if (esc) {
            if (!have) { bad = true; break; }  // an escape with nothing to extend (CanonicalHuffman.java:495-504 would index -1)
            if (sym == kSymEsc2) { cur = (cur << 2) | src.bits(pos, 2); pos += 2; }
            else { cur = (cur << 8) | src.bits(pos, 8); pos += 8; }
          } else {
            if (have) sink(o++, int32_t(cur));
            have = true;
            cur = sym == kSymNull ? uint32_t(INT32_MIN) : uint32_t(sym - 128);
          }
        }
        if (have) sink(o, int32_t(cur));
      }
    }
    if (__syncthreads_or(bad ? 1 : 0)) return false;
    *endBit = S.endpos[fe];
    *nValues = total;
    return true;
  }
}

}  // namespace g4
