// g4_huffdec.cuh -- legacy Huffman stream decoder (tree parse, lookup table, self-synchronising parallel
// sub-sequence decode).  Reference: compress/HuffmanDecoder.java:65-187
// (under /root/reference/core/src/main/java/org/gridfour/).  Used by CodecHuffman decode and by the LSOP12
// legacy-Huffman variant (lsop/LsDecoder12.java:119-124).
#pragma once
#include "g4_device.cuh"

namespace g4 {


constexpr int kLutBits = 11;
constexpr int kMaxSub = 2048;                   // sub-sequences per tile (shared arrays)
constexpr int kSubPerThread = kMaxSub / kThreads;

struct HuffDecShared {
  uint16_t lut[1 << kLutBits];  // bit15: node reference (low 9 bits = node); else sym | len<<9
  uint16_t kid[512][2];
  int16_t leafSym[512];         // -1 for branch nodes
  uint32_t endpos[kMaxSub];
  uint16_t cnt[kMaxSub];
  uint32_t scan[kWarps + 1];
  uint32_t treeBits;
  int nLeaf;
  int single;                   // single-symbol stream: the symbol, else -1
  int error;
  int changed;
};

// HuffmanDecoder.decodeTree (HuffmanDecoder.java:65-161), bounds-checked.  One thread.
__device__ inline void parse_tree(HuffDecShared& S, const BitSrc& src, uint32_t startBit) {
  S.error = 0;
  S.single = -1;
  int L = int(src.bits(startBit, 8)) + 1;
  S.nLeaf = L;
  uint32_t pos = startBit + 8;
  if (src.bits(pos, 1)) {
    S.single = int(src.bits(pos + 1, 8));
    S.treeBits = startBit + 17;
    if (startBit + 17 > src.nBits) S.error = 1;
    return;
  }
  pos = startBit + 9;
  uint16_t stack[260];
  uint8_t slot[512];
  int nodes = 1, sp = 1, leaves = 0;
  stack[0] = 0;
  slot[0] = 0;
  S.leafSym[0] = -1;
  GlobalCursor cur;  // register bit buffer: the tree is read one and eight bits at a time
  cur.init(src, pos);
  while (leaves < L) {
    if (sp == 0 || nodes >= 511 || cur.pos + 9 > src.nBits + 8) { S.error = 1; return; }
    int parent = stack[sp - 1];
    uint32_t bit = cur.peek() & 1u;
    cur.skip(1);
    int id = nodes++;
    S.kid[parent][slot[parent]++] = uint16_t(id);
    if (bit) {
      S.leafSym[id] = int16_t(cur.peek() & 0xffu);
      cur.skip(8);
      leaves++;
      while (sp > 0 && slot[stack[sp - 1]] == 2) sp--;
    } else {
      if (sp >= 258) { S.error = 1; return; }
      S.leafSym[id] = -1;
      slot[id] = 0;
      stack[sp++] = uint16_t(id);
    }
  }
  pos = cur.pos;
  if (sp != 0 || pos > src.nBits) { S.error = 1; return; }  // incomplete tree / truncated stream
  S.treeBits = pos;
}

// Decode one symbol at *pos.
__device__ __forceinline__ int decode_symbol(const HuffDecShared& S, const BitSrc& src, uint32_t* pos) {
  uint32_t v = src.peek32(*pos);
  uint32_t e = S.lut[v & ((1u << kLutBits) - 1)];
  if (!(e & 0x8000u)) {
    *pos += (e >> 9) & 15u;
    return int(e & 0xffu);
  }
  int n = int(e & 0x1ffu);
  int consumed = kLutBits;
  uint32_t p = *pos;
  while (S.leafSym[n] < 0) {
    if (consumed == 32) { p += 32; v = src.peek32(p); consumed = 0; }
    n = S.kid[n][(v >> consumed) & 1u];
    consumed++;
  }
  *pos = p + consumed;
  return S.leafSym[n];
}

__device__ __forceinline__ void decode_subseq(const HuffDecShared& S, const BitSrc& src, uint32_t start, uint32_t limit,
                                              uint32_t* endOut, uint32_t* cntOut) {
  uint32_t pos = start, c = 0;
  while (pos < limit) { decode_symbol(S, src, &pos); c++; }
  *endOut = pos;
  *cntOut = c;
}

// Decodes the legacy Huffman text of `nSym` symbols (tree first) into
// `out` (memory visible to the CTA); the tree starts at bit `startBit` of `src`.  Returns the bit position just after the last symbol through
// *endBit.  Shared by CodecHuffman decode and the LSOP12 legacy-Huffman variant.  All threads call.
__device__ inline bool huffman_decode_stream(HuffDecShared& S, const BitSrc& src, uint32_t startBit, uint32_t nSym, uint8_t* out,
                                             uint32_t* endBit) {
  const int tid = threadIdx.x;
  __syncthreads();
  if (tid == 0) parse_tree(S, src, startBit);
  __syncthreads();
  if (S.error) return false;
  if (S.single >= 0) {
    for (uint32_t i = tid; i < nSym; i += kThreads) out[i] = uint8_t(S.single);
    *endBit = S.treeBits;
    __syncthreads();
    return true;
  }
  // lookup table: every thread resolves a slice of the 2^11 prefixes by walking the tree
  for (int e = tid; e < (1 << kLutBits); e += kThreads) {
    int n = 0;
    uint16_t entry = 0;
    int d = 0;
    for (; d < kLutBits; d++) {
      n = S.kid[n][(e >> d) & 1];
      if (S.leafSym[n] >= 0) { entry = uint16_t(S.leafSym[n] | ((d + 1) << 9)); break; }
    }
    if (d == kLutBits) entry = uint16_t(0x8000u | n);
    S.lut[e] = entry;
  }
  __syncthreads();
  const uint32_t T0 = S.treeBits;
  const uint32_t avail = src.nBits - T0;
  uint32_t B = (avail + kMaxSub - 1) / kMaxSub;
  B = (B + 31u) & ~31u;
  if (B < 128u) B = 128u;
  const int nSub = int((avail + B - 1) / B);
  // pass 0: speculative decode of every sub-sequence from its nominal start
  uint32_t myStart[kSubPerThread];
#pragma unroll
  for (int j = 0; j < kSubPerThread; j++) {
    int i = tid + j * kThreads;
    myStart[j] = T0 + uint32_t(i) * B;
    if (i < nSub) {
      uint32_t limit = T0 + uint32_t(i + 1) * B;
      if (limit > src.nBits) limit = src.nBits;
      uint32_t e, c;
      decode_subseq(S, src, myStart[j], limit, &e, &c);
      S.endpos[i] = e;
      S.cnt[i] = uint16_t(c);
    }
  }
  // synchronisation passes: sub-sequence i must start where i-1 ended
  for (int pass = 0; pass < nSub; pass++) {
    __syncthreads();
    if (tid == 0) S.changed = 0;
    uint32_t ns[kSubPerThread];
#pragma unroll
    for (int j = 0; j < kSubPerThread; j++) {
      int i = tid + j * kThreads;
      ns[j] = (i > 0 && i < nSub) ? S.endpos[i - 1] : myStart[j];
    }
    __syncthreads();
    bool any = false;
#pragma unroll
    for (int j = 0; j < kSubPerThread; j++) {
      int i = tid + j * kThreads;
      if (i < nSub && ns[j] != myStart[j]) {
        myStart[j] = ns[j];
        uint32_t limit = T0 + uint32_t(i + 1) * B;
        if (limit > src.nBits) limit = src.nBits;
        uint32_t e, c;
        decode_subseq(S, src, myStart[j], limit, &e, &c);
        S.endpos[i] = e;
        S.cnt[i] = uint16_t(c);
        any = true;
      }
    }
    if (any) S.changed = 1;
    __syncthreads();
    if (!S.changed) break;
  }
  // symbol offsets of the sub-sequences
  uint32_t local[kSubPerThread];
  uint32_t mySum = 0;
  // thread tid owns sub-sequences tid*kSubPerThread .. +kSubPerThread-1 for the scan (contiguous)
#pragma unroll
  for (int j = 0; j < kSubPerThread; j++) {
    int i = tid * kSubPerThread + j;
    local[j] = i < nSub ? S.cnt[i] : 0u;
    mySum += local[j];
  }
  uint32_t total;
  uint32_t ex = block_exclusive_scan(mySum, S.scan, &total);
  if (total < nSym) return false;  // text shorter than the header claims
  __syncthreads();
  __shared__ uint32_t sOff[kMaxSub];  // first output symbol index of every sub-sequence
  {
    uint32_t run = ex;
#pragma unroll
    for (int j = 0; j < kSubPerThread; j++) {
      int i = tid * kSubPerThread + j;
      if (i < nSub) sOff[i] = run;
      run += local[j];
    }
  }
  __syncthreads();
  // write pass
  uint32_t lastEnd = 0;
#pragma unroll
  for (int j = 0; j < kSubPerThread; j++) {
    int i = tid + j * kThreads;
    if (i < nSub) {
      uint32_t pos = myStart[j];
      uint32_t limit = T0 + uint32_t(i + 1) * B;
      if (limit > src.nBits) limit = src.nBits;
      uint32_t o = sOff[i];
      while (pos < limit && o < nSym) {
        int s = decode_symbol(S, src, &pos);
        out[o++] = uint8_t(s);
        if (o == nSym) lastEnd = pos;  // this thread decoded the final symbol
      }
    }
  }
  if (lastEnd) S.scan[kWarps] = lastEnd;
  __syncthreads();
  *endBit = S.scan[kWarps];
  return true;
}


}  // namespace g4
