// g4_huff_fast.cuh -- legacy Huffman text decoder, fast path: the packing is staged in shared memory and every thread
// decodes from a 64-bit register bit buffer, up to three symbols per table lookup.
//
// Same stream format and the same self-synchronising sub-sequence scheme as g4_huffdec.cuh (reference:
// compress/HuffmanDecoder.java:65-187, compress/CodecHuffman.java:133-153 under
// /root/reference/core/src/main/java/org/gridfour/); the machinery mirrors g4_canon_fast.cuh.  The symbols are the
// bytes of the M32 code of the residuals, so the write pass is a byte stream: a register byte queue that leaves eight
// aligned bytes at a time.  Packings that do not fit the staging buffer use g4_huffdec.cuh.
#pragma once
#include "g4_huffdec.cuh"

namespace g4 {

constexpr int kHfMaxSub = 1280;
constexpr int kHfRounds = kHfMaxSub / kThreads;
constexpr uint32_t kHfSubBits = 160;
constexpr uint32_t kHfLookback = 64;
constexpr int kHfStageWordsMin = 2048;  // 8 KB

struct HuffFastShared {
  uint32_t mlut[1 << kLutBits];  // up to 3 symbols per lookup: s1 | s2 << 8 | s3 << 16 | bits << 24 | n << 28
  uint16_t lut[1 << kLutBits];   // sym | len << 9; bit 15: code longer than the table, low 9 bits = tree node reached
  uint16_t kid[512][2];
  int16_t leafSym[512];
  uint32_t endpos[kHfMaxSub];
  uint32_t startv[kHfMaxSub];
  uint16_t cnt[kHfMaxSub];
  uint32_t scan[kWarps + 1];
  uint32_t treeBits, endBit;
  int nLeaf, single, error, changed;
  uint32_t sw[kHfStageWordsMin + 8];  // staged packing from its 8-byte aligned start; LAST member, may be allocated longer
};
__host__ __device__ constexpr size_t huff_fast_smem_bytes(uint32_t stageWords) {
  return sizeof(HuffFastShared) + (stageWords > uint32_t(kHfStageWordsMin) ? size_t(stageWords - kHfStageWordsMin) * 4 : 0);
}

struct HfBits {  // random access into the staged words (absolute staged bit positions)
  const uint32_t* w;
  __device__ __forceinline__ uint32_t peek32(uint32_t pos) const { return __funnelshift_r(w[pos >> 5], w[(pos >> 5) + 1], pos & 31); }
  __device__ __forceinline__ uint32_t bits(uint32_t pos, int n) const { return peek32(pos) & ((1u << n) - 1u); }
};

struct HfCursor {  // 64-bit register bit buffer; at least 32 valid bits after every skip()
  uint32_t pos, next;
  uint64_t buf;
  int avail;
  __device__ __forceinline__ void init(const HuffFastShared& S, uint32_t p) {
    pos = p;
    const uint32_t i = p >> 5, s = p & 31;
    buf = ((uint64_t(S.sw[i + 1]) << 32) | S.sw[i]) >> s;
    avail = 64 - int(s);
    next = i + 2;
  }
  __device__ __forceinline__ uint32_t peek() const { return uint32_t(buf); }
  __device__ __forceinline__ void skip(const HuffFastShared& S, uint32_t n) {
    buf >>= n;
    avail -= int(n);
    pos += n;
    if (avail < 32) {
      buf |= uint64_t(S.sw[next++]) << avail;
      avail += 32;
    }
  }
};

// HuffmanDecoder.decodeTree (:65-161) over the staged words, bounds-checked.  One thread.  nBits = end of the packing.
__device__ inline void hf_parse_tree(HuffFastShared& S, uint32_t startBit, uint32_t nBits) {
  HfBits src{S.sw};
  S.error = 0;
  S.single = -1;
  if (startBit + 17 > nBits) { S.error = 1; return; }
  const int L = int(src.bits(startBit, 8)) + 1;
  S.nLeaf = L;
  uint32_t pos = startBit + 8;
  if (src.bits(pos, 1)) {
    S.single = int(src.bits(pos + 1, 8));
    S.treeBits = startBit + 17;
    return;
  }
  pos = startBit + 9;
  uint16_t stack[260];
  uint8_t slot[512];
  int nodes = 1, sp = 1, leaves = 0;
  stack[0] = 0;
  slot[0] = 0;
  S.leafSym[0] = -1;
  HfCursor cur;  // register bit buffer: the tree is read one and eight bits at a time while 255 threads wait
  cur.init(S, pos);
  while (leaves < L) {
    if (sp == 0 || nodes >= 511 || cur.pos + 9 > nBits + 8) { S.error = 1; return; }
    const int parent = stack[sp - 1];
    const uint32_t bit = cur.peek() & 1u;
    cur.skip(S, 1);
    const int id = nodes++;
    S.kid[parent][slot[parent]++] = uint16_t(id);
    if (bit) {
      S.leafSym[id] = int16_t(cur.peek() & 0xffu);
      slot[id] = 2;
      cur.skip(S, 8);
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
  if (sp != 0 || pos > nBits) { S.error = 1; return; }
  S.treeBits = pos;
}

// One symbol whose code is longer than the table: walk the tree from the node the table entry names.
__device__ __forceinline__ int hf_long_symbol(const HuffFastShared& S, uint32_t e, uint32_t p0, uint32_t* after) {
  HfBits src{S.sw};
  int n = int(e & 0x1ffu);
  uint32_t p = p0 + uint32_t(kLutBits);
  while (S.leafSym[n] < 0) {
    n = S.kid[n][src.bits(p, 1)];
    p++;
  }
  *after = p;
  return S.leafSym[n];
}

// Counting decode of one sub-sequence: from `start` to the first symbol boundary at or after `limit`.
__device__ __forceinline__ void hf_count(const HuffFastShared& S, uint32_t start, uint32_t limit, uint32_t* endOut, uint32_t* cntOut) {
  HfCursor cur;
  cur.init(S, start);
  uint32_t c = 0;
  for (;;) {
    const uint32_t p0 = cur.pos;
    if (p0 >= limit) break;
    if (p0 + uint32_t(kLutBits) <= limit) {
      const uint32_t m = S.mlut[cur.peek() & ((1u << kLutBits) - 1u)];
      if (m >> 28) {
        cur.skip(S, (m >> 24) & 15u);
        c += m >> 28;
        continue;
      }
    }
    const uint32_t e = S.lut[cur.peek() & ((1u << kLutBits) - 1u)];
    if (!(e & 0x8000u)) {
      cur.skip(S, (e >> 9) & 15u);
      c++;
      continue;
    }
    uint32_t after;
    hf_long_symbol(S, e, p0, &after);
    c++;
    cur.init(S, after);
  }
  *endOut = cur.pos;
  *cntOut = c;
}

// Byte sink of the write pass: bytes go out eight at a time once the output index is 8-byte aligned.
struct HfByteSink {
  uint8_t* base;
  uint32_t o;     // next output index
  int cnt, need;  // queued bytes; bytes that complete the current aligned group of eight
  uint64_t lo;
  uint32_t hi;
  __device__ __forceinline__ void begin(uint8_t* b, uint32_t o0) {
    base = b;
    o = o0;
    cnt = 0;
    lo = 0;
    hi = 0;
    need = 8 - int(o0 & 7u);
  }
  __device__ __forceinline__ void drop(int k) {
    const int sh = 8 * k;  // k = 1..10; hi holds at most two bytes
    if (k >= 8) { lo = uint64_t(hi) >> (sh - 64); hi = 0; }
    else if (k > 0) { lo = (lo >> sh) | (uint64_t(hi) << (64 - sh)); hi = sh < 16 ? hi >> sh : 0u; }
    cnt -= k;
  }
  __device__ __forceinline__ void write_bytes(int k) {
    uint64_t x = lo;
    for (int i = 0; i < k && i < 8; i++) { base[o + i] = uint8_t(x); x >>= 8; }
    if (k > 8) base[o + 8] = uint8_t(hi);
    if (k > 9) base[o + 9] = uint8_t(hi >> 8);
    drop(k);
    o += uint32_t(k);
  }
  __device__ __forceinline__ void push(uint32_t bytes, int n) {
    const int sh = cnt * 8;  // cnt <= 7
    lo |= uint64_t(bytes) << sh;
    if (sh > 40) hi |= bytes >> (64 - sh);
    cnt += n;
    while (cnt >= need) {
      if (need == 8) {
        *reinterpret_cast<uint2*>(base + o) = make_uint2(uint32_t(lo), uint32_t(lo >> 32));
        lo = hi;
        hi = 0;
        cnt -= 8;
        o += 8;
      } else write_bytes(need);
      need = 8;
    }
  }
  __device__ __forceinline__ void end() { write_bytes(cnt); }
};

// Decodes the legacy Huffman stream (tree at staged bit `startBit`, then the text of nSym symbols) into out[0..nSym);
// nBits = staged bit position of the end of the packing.  out must be 8-byte aligned and writable up to nSym + 8.
// All threads call.  *endBit = staged bit position just after the last symbol.
__device__ inline bool huff_fast_decode_stream(HuffFastShared& S, uint32_t startBit, uint32_t nBits, uint32_t nSym, uint8_t* out,
                                               uint32_t* endBit) {
  const int tid = threadIdx.x;
  __syncthreads();
  if (tid == 0) hf_parse_tree(S, startBit, nBits);
  __syncthreads();
  if (S.error) return false;
  if (S.single >= 0) {
    for (uint32_t i = tid; i < nSym; i += kThreads) out[i] = uint8_t(S.single);
    *endBit = S.treeBits;
    __syncthreads();
    return true;
  }
  for (int e = tid; e < (1 << kLutBits); e += kThreads) {
    int n = 0, d = 0;
    uint16_t entry = 0;
    for (; d < kLutBits; d++) {
      n = S.kid[n][(e >> d) & 1];
      if (S.leafSym[n] >= 0) { entry = uint16_t(S.leafSym[n] | ((d + 1) << 9)); break; }
    }
    if (d == kLutBits) entry = uint16_t(0x8000u | n);
    S.lut[e] = entry;
  }
  __syncthreads();
  for (int e = tid; e < (1 << kLutBits); e += kThreads) {
    uint32_t used = 0, n = 0, syms = 0;
    while (n < 3) {
      const uint32_t x = S.lut[(uint32_t(e) >> used) & ((1u << kLutBits) - 1u)];
      const uint32_t len = (x >> 9) & 15u;
      if ((x & 0x8000u) || used + len > uint32_t(kLutBits)) break;
      syms |= (x & 0xffu) << (8 * n);
      used += len;
      n++;
    }
    S.mlut[e] = syms | (used << 24) | (n << 28);
  }
  __syncthreads();
  const uint32_t T0 = S.treeBits;
  if (T0 > nBits) return false;
  const uint32_t avail = nBits - T0;
  uint32_t rounds = (avail / kHfSubBits + kThreads - 1) / kThreads;
  if (rounds < 1u) rounds = 1u;
  if (rounds > uint32_t(kHfRounds)) rounds = kHfRounds;
  uint32_t B = (avail + rounds * kThreads - 1) / (rounds * kThreads);
  if (B < 96u) B = 96u;
  const int nSub = int((avail + B - 1) / B);
  if (nSub == 0) return nSym == 0;
  // pass 0: only the END of every sub-sequence matters, so start a little before the limit and rely on self-synchronisation
#pragma unroll 1
  for (int i = tid; i < nSub; i += kThreads) {
    uint32_t limit = T0 + uint32_t(i + 1) * B;
    if (limit > nBits) limit = nBits;
    uint32_t from = T0 + uint32_t(i) * B;
    if (i > 0 && limit - from > kHfLookback) from = limit - kHfLookback;
    uint32_t e, c;
    hf_count(S, from, limit, &e, &c);
    S.endpos[i] = e;
    S.cnt[i] = uint16_t(c);
    S.startv[i] = i > 0 ? 0xffffffffu : T0;
  }
  // Chaotic relaxation on purpose: a thread may read vend[i-1] in the same pass in which its owner rewrites it (racecheck
  // reports this read/write pair).  Either value is a 32-bit end position that was valid at some point, the writer raises
  // `changed`, and the loop only ends after a pass in which nobody wrote -- every read of that pass saw final values.
  volatile uint32_t* vend = S.endpos;
  for (int pass = 0; pass <= nSub; pass++) {
    __syncthreads();
    if (tid == 0) S.changed = 0;
    __syncthreads();
    bool any = false;
#pragma unroll 1
    for (int i = tid; i < nSub; i += kThreads) {
      if (i == 0) continue;
      const uint32_t ns = vend[i - 1];
      if (ns != S.startv[i]) {
        S.startv[i] = ns;
        uint32_t limit = T0 + uint32_t(i + 1) * B;
        if (limit > nBits) limit = nBits;
        uint32_t e, c;
        hf_count(S, ns, limit, &e, &c);
        vend[i] = e;
        S.cnt[i] = uint16_t(c);
        any = true;
      }
    }
    if (any) S.changed = 1;
    __syncthreads();
    if (!S.changed) break;
  }
  __syncthreads();
  // symbol offsets: thread tid owns sub-sequences tid*kHfRounds .. +kHfRounds-1 for the scan
  uint32_t mySum = 0;
#pragma unroll
  for (int j = 0; j < kHfRounds; j++) {
    const int i = tid * kHfRounds + j;
    mySum += i < nSub ? S.cnt[i] : 0u;
  }
  uint32_t total;
  const uint32_t ex = block_exclusive_scan(mySum, S.scan, &total);
  if (total < nSym) return false;  // text shorter than the header claims
  __syncthreads();
  uint32_t* offv = S.endpos;  // end positions are no longer needed
  {
    uint32_t run = ex;
#pragma unroll
    for (int j = 0; j < kHfRounds; j++) {
      const int i = tid * kHfRounds + j;
      if (i < nSub) {
        offv[i] = run;
        run += S.cnt[i];
      }
    }
  }
  if (tid == 0) S.endBit = 0;
  __syncthreads();
  // write pass: every thread takes `rounds` consecutive sub-sequences as one run
  const int i0 = tid * int(rounds);
  if (i0 < nSub && offv[i0] < nSym) {
    const int i1 = i0 + int(rounds) - 1 < nSub - 1 ? i0 + int(rounds) - 1 : nSub - 1;
    uint32_t limit = T0 + uint32_t(i1 + 1) * B;
    if (limit > nBits) limit = nBits;
    const uint32_t maxCount = nSym - offv[i0];  // the text may be followed by padding bits that decode to symbols
    uint32_t produced = 0;
    HfCursor cur;
    cur.init(S, S.startv[i0]);
    HfByteSink sink;
    sink.begin(out, offv[i0]);
    for (;;) {
      const uint32_t p0 = cur.pos;
      if (p0 >= limit || produced >= maxCount) break;
      if (p0 + uint32_t(kLutBits) <= limit && produced + 3u <= maxCount) {
        const uint32_t m = S.mlut[cur.peek() & ((1u << kLutBits) - 1u)];
        const uint32_t n = m >> 28;
        if (n) {
          cur.skip(S, (m >> 24) & 15u);
          sink.push(m & 0xffffffu, int(n));
          produced += n;
          continue;
        }
      }
      const uint32_t e = S.lut[cur.peek() & ((1u << kLutBits) - 1u)];
      if (!(e & 0x8000u)) {
        cur.skip(S, (e >> 9) & 15u);
        sink.push(e & 0xffu, 1);
      } else {
        uint32_t after;
        const int sym = hf_long_symbol(S, e, p0, &after);
        sink.push(uint32_t(sym), 1);
        cur.init(S, after);
      }
      produced++;
    }
    sink.end();
    if (produced == maxCount) S.endBit = cur.pos;  // this thread decoded the final symbol
  }
  __syncthreads();
  *endBit = S.endBit;
  return true;
}

}  // namespace g4
