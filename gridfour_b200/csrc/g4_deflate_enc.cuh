// g4_deflate_enc.cuh -- hand-written RFC 1950 / RFC 1951 ENCODER, one thread per stream.
//
// Stands in for java.util.zip.Deflater(level).finish() + one deflate() call at the reference's call sites
// (compress/CodecDeflate.java:204-213 level 6, compress/CodecFloat.java:268-283 level 9,
//  lsop/LsEncoder12.java:180-196 level 6; paths under /root/reference/core/src/main/java/org/gridfour/).
// The JDK delegates to zlib; zlib is NOT part of /root/reference, so this file restates zlib's published
// algorithm (deflate.c "deflate_slow" lazy matching with the level's good/lazy/nice/chain limits, 15-bit
// 3-byte hash chains over a 32 KiB window, a block every 16383 symbols, trees.c Huffman construction with
// the depth tie-break and bit-length overflow repair, stored/fixed/dynamic choice by cost).  Streams need
// not be byte-identical to zlib's -- they must inflate with any RFC 1950 decoder and be close enough in
// size that the reference's codec selection is reproduced -- but every step follows the same rule zlib uses.
//
// Hash-chain matching is serial by nature (each decision depends on the previous match), so parallelism
// comes from streams: thousands of (tile, predictor) streams are in flight, one per thread, with their hash
// tables and symbol buffers in HBM scratch.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

// The encoder is plain integer code; it also compiles for the host so that tests/test_deflate_host.py can
// pin it byte for byte against the system zlib without a GPU (the product only ever runs it on the device).
#define G4_HD __host__ __device__

namespace g4 {

G4_HD __forceinline__ int g4_clz32(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __clz(int(x));
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
G4_HD __forceinline__ uint32_t g4_brev32(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __brev(x);
#else
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
  x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
  return (x >> 16) | (x << 16);
#endif
}

constexpr int kDefWSize = 32768, kDefWMask = 32767, kDefHashMask = 32767;
constexpr int kDefMinMatch = 3, kDefMaxMatch = 258;
constexpr int kDefMinLookahead = kDefMaxMatch + kDefMinMatch + 1;
constexpr int kDefMaxDist = kDefWSize - kDefMinLookahead;
constexpr int kDefTooFar = 4096;
constexpr int kDefLitBufSize = 16384;  // memLevel 8
constexpr int kDefLCodes = 286, kDefDCodes = 30, kDefBlCodes = 19, kDefHeapSize = 2 * kDefLCodes + 1;

struct DeflateLevel {
  int goodLength, maxLazy, niceLength, maxChain;
  uint32_t zlibHeader;
};
G4_HD __forceinline__ DeflateLevel deflate_level(int level) {
  DeflateLevel L;
  if (level >= 9) { L.goodLength = 32; L.maxLazy = 258; L.niceLength = 258; L.maxChain = 4096; L.zlibHeader = 0x78DAu; }
  else { L.goodLength = 8; L.maxLazy = 16; L.niceLength = 128; L.maxChain = 128; L.zlibHeader = 0x789Cu; }
  return L;
}

struct CtData {
  uint16_t fc;  // frequency or code
  uint16_t dl;  // dad or length
};

// Huffman construction state of one block (trees.c): small enough for shared memory.
struct DeflateTrees {
  CtData ltree[kDefHeapSize], dtree[2 * kDefDCodes + 1], bltree[2 * kDefBlCodes + 1];
  int16_t heap[kDefHeapSize];
  // Order key of the node in heap[k], k = 1 .. heapLen: frequency << 8 | depth.  trees.c's smaller(n, m) -- "freq[n] <
  // freq[m], or equal and depth[n] <= depth[m]" -- is key(n) <= key(m), so a sift-down step reads two adjacent keys
  // instead of chasing heap[] -> tree[].Freq -> depth[] for each child (the tree build is a chain of dependent loads on
  // ONE thread, and the emit kernel's other 255 wait for it).
  uint32_t hkey[kDefHeapSize];
};

// Per-stream work area in HBM (one-thread-per-stream encoder).
struct DeflateWork {
  uint16_t symDist[kDefLitBufSize];
  uint8_t symLc[kDefLitBufSize];
  DeflateTrees T;
  uint32_t head[kDefWSize];
  uint32_t prev[kDefWSize];
};

struct DeflateOut {
  uint8_t* out;
  uint32_t cap, pos;
  uint64_t bits;
  int nbits;
  bool overflow;
  G4_HD __forceinline__ void put_byte(uint32_t b) {
    if (pos < cap) out[pos] = uint8_t(b); else overflow = true;
    pos++;
  }
  G4_HD __forceinline__ void send_bits(uint32_t v, int n) {  // LSB first
    bits |= uint64_t(v) << nbits;
    nbits += n;
    while (nbits >= 8) { put_byte(uint32_t(bits) & 0xffu); bits >>= 8; nbits -= 8; }
  }
  G4_HD __forceinline__ void windup() {
    if (nbits > 0) put_byte(uint32_t(bits) & 0xffu);
    bits = 0;
    nbits = 0;
  }
};

G4_HD __forceinline__ uint32_t def_bi_reverse(uint32_t code, int len) { return g4_brev32(code) >> (32 - len); }

// length/distance code arithmetic (equivalent to zlib's _length_code / _dist_code / base / extra tables)
G4_HD __forceinline__ int def_length_code(int lc) {  // lc = length - 3
  if (lc < 8) return lc;
  if (lc == 255) return 28;
  int msb = 31 - g4_clz32(uint32_t(lc));
  int extra = msb - 2;
  return 4 * extra + 4 + ((lc >> extra) & 3);
}
G4_HD __forceinline__ int def_length_extra(int code) { return (code < 8 || code == 28) ? 0 : (code >> 2) - 1; }
G4_HD __forceinline__ int def_length_base(int code) { return code < 8 ? code : code == 28 ? 255 : (4 + (code & 3)) << ((code >> 2) - 1); }
G4_HD __forceinline__ int def_dist_code(int d) {  // d = distance - 1
  if (d < 4) return d;
  int msb = 31 - g4_clz32(uint32_t(d));
  return 2 * msb + ((d >> (msb - 1)) & 1);
}
G4_HD __forceinline__ int def_dist_extra(int code) { return code < 4 ? 0 : (code >> 1) - 1; }
G4_HD __forceinline__ int def_dist_base(int code) { return code < 4 ? code : (2 + (code & 1)) << ((code >> 1) - 1); }
G4_HD __forceinline__ int def_static_llen(int n) { return n < 144 ? 8 : n < 256 ? 9 : n < 280 ? 7 : 8; }
G4_HD __forceinline__ uint32_t def_static_lcode(int n) {
  uint32_t c = n < 144 ? 0x30u + n : n < 256 ? 0x190u + (n - 144) : n < 280 ? uint32_t(n - 256) : 0xC0u + (n - 280);
  return def_bi_reverse(c, def_static_llen(n));
}

struct DeflateTreeDesc {
  CtData* tree;
  int elems, maxLength, extraBase;  // extraBase: first code with extra bits (257 lit/len, 0 dist, 0 bl)
  int kind;                         // 0 literal/length, 1 distance, 2 bit-length
  int maxCode;
};

G4_HD __forceinline__ int def_extra_bits(const DeflateTreeDesc& d, int n) {
  if (d.kind == 0) return n >= 257 ? def_length_extra(n - 257) : 0;
  if (d.kind == 1) return def_dist_extra(n);
  return n == 16 ? 2 : n == 17 ? 3 : n == 18 ? 7 : 0;
}
G4_HD __forceinline__ int def_static_len(const DeflateTreeDesc& d, int n) { return d.kind == 0 ? def_static_llen(n) : d.kind == 1 ? 5 : 0; }

struct DeflateState {
  DeflateWork* W;    // symbol buffer (and hash chains) of the serial encoder; unused by the tree code
  DeflateTrees* T;
  unsigned long long optLen, staticLen;
  int heapLen, heapMax;
  uint16_t blCount[16];
};

G4_HD inline void def_pqdownheap(DeflateState& s, const CtData*, int k) {
  int16_t* heap = s.T->heap;
  uint32_t* hkey = s.T->hkey;
  const int v = heap[k];
  const uint32_t vk = hkey[k];
  int j = k << 1;
  while (j <= s.heapLen) {
    uint32_t kj = hkey[j];
    if (j < s.heapLen) {
      const uint32_t kj1 = hkey[j + 1];
      if (kj1 <= kj) { j++; kj = kj1; }  // smaller(heap[j + 1], heap[j])
    }
    if (vk <= kj) break;  // smaller(v, heap[j])
    heap[k] = heap[j];
    hkey[k] = kj;
    k = j;
    j <<= 1;
  }
  heap[k] = int16_t(v);
  hkey[k] = vk;
}

// trees.c gen_bitlen
G4_HD inline void def_gen_bitlen(DeflateState& s, DeflateTreeDesc& d) {
  CtData* tree = d.tree;
  int16_t* heap = s.T->heap;
  const int maxLength = d.maxLength;
  for (int b = 0; b <= 15; b++) s.blCount[b] = 0;
  tree[heap[s.heapMax]].dl = 0;
  int overflow = 0;
  int h;
  for (h = s.heapMax + 1; h < kDefHeapSize; h++) {
    int n = heap[h];
    int bits = tree[tree[n].dl].dl + 1;
    if (bits > maxLength) { bits = maxLength; overflow++; }
    tree[n].dl = uint16_t(bits);
    if (n > d.maxCode) continue;
    s.blCount[bits]++;
    int xbits = def_extra_bits(d, n);
    unsigned f = tree[n].fc;
    s.optLen += (unsigned long long)f * unsigned(bits + xbits);
    if (d.kind != 2) s.staticLen += (unsigned long long)f * unsigned(def_static_len(d, n) + xbits);
  }
  if (overflow == 0) return;
  do {
    int bits = maxLength - 1;
    while (s.blCount[bits] == 0) bits--;
    s.blCount[bits]--;
    s.blCount[bits + 1] += 2;
    s.blCount[maxLength]--;
    overflow -= 2;
  } while (overflow > 0);
  for (int bits = maxLength; bits != 0; bits--) {
    int n = s.blCount[bits];
    while (n != 0) {
      int m = heap[--h];
      if (m > d.maxCode) continue;
      if (tree[m].dl != unsigned(bits)) {
        s.optLen += (unsigned long long)(long long)(bits - int(tree[m].dl)) * tree[m].fc;
        tree[m].dl = uint16_t(bits);
      }
      n--;
    }
  }
}

// trees.c build_tree (+ gen_codes)
G4_HD inline void def_build_tree(DeflateState& s, DeflateTreeDesc& d) {
  CtData* tree = d.tree;
  int16_t* heap = s.T->heap;
  uint32_t* hkey = s.T->hkey;
  const int elems = d.elems;
  int maxCode = -1;
  s.heapLen = 0;
  s.heapMax = kDefHeapSize;
  for (int n = 0; n < elems; n++) {
    const uint32_t f = tree[n].fc;
    if (f != 0) { heap[++s.heapLen] = int16_t(maxCode = n); hkey[s.heapLen] = f << 8; }  // depth 0
    else tree[n].dl = 0;
  }
  while (s.heapLen < 2) {
    int node = heap[++s.heapLen] = int16_t(maxCode < 2 ? ++maxCode : 0);
    tree[node].fc = 1;
    hkey[s.heapLen] = 1u << 8;
    s.optLen--;
    if (d.kind != 2) s.staticLen -= unsigned(def_static_len(d, node));
  }
  d.maxCode = maxCode;
  for (int n = s.heapLen / 2; n >= 1; n--) def_pqdownheap(s, tree, n);
  int node = elems;
  do {
    const int n = heap[1];
    const uint32_t kn = hkey[1];
    heap[1] = heap[s.heapLen];
    hkey[1] = hkey[s.heapLen];
    s.heapLen--;
    def_pqdownheap(s, tree, 1);
    const int m = heap[1];
    const uint32_t km = hkey[1];
    heap[--s.heapMax] = int16_t(n);
    heap[--s.heapMax] = int16_t(m);
    const uint32_t f = ((kn >> 8) + (km >> 8)) & 0xffffu;  // Freq is a ush
    const uint32_t dn = kn & 0xffu, dm = km & 0xffu;
    tree[node].fc = uint16_t(f);
    tree[n].dl = tree[m].dl = uint16_t(node);
    heap[1] = int16_t(node++);
    hkey[1] = (f << 8) | (((dn >= dm ? dn : dm) + 1u) & 0xffu);  // depth is a uch
    def_pqdownheap(s, tree, 1);
  } while (s.heapLen >= 2);
  heap[--s.heapMax] = heap[1];
  def_gen_bitlen(s, d);
  // gen_codes
  uint16_t nextCode[16];
  unsigned code = 0;
  for (int b = 1; b <= 15; b++) { code = (code + s.blCount[b - 1]) << 1; nextCode[b] = uint16_t(code); }
  for (int n = 0; n <= maxCode; n++) {
    int len = tree[n].dl;
    if (len == 0) continue;
    tree[n].fc = uint16_t(def_bi_reverse(nextCode[len]++, len));
  }
}

// trees.c scan_tree: run-length statistics of a code-length sequence into the bit-length tree
G4_HD inline void def_scan_tree(DeflateState& s, CtData* tree, int maxCode) {
  CtData* bl = s.T->bltree;
  int prevlen = -1, nextlen = tree[0].dl, count = 0, maxCount = 7, minCount = 4;
  if (nextlen == 0) { maxCount = 138; minCount = 3; }
  tree[maxCode + 1].dl = 0xffff;  // guard
  for (int n = 0; n <= maxCode; n++) {
    int curlen = nextlen;
    nextlen = tree[n + 1].dl;
    if (++count < maxCount && curlen == nextlen) continue;
    else if (count < minCount) bl[curlen].fc += uint16_t(count);
    else if (curlen != 0) { if (curlen != prevlen) bl[curlen].fc++; bl[16].fc++; }
    else if (count <= 10) bl[17].fc++;
    else bl[18].fc++;
    count = 0;
    prevlen = curlen;
    if (nextlen == 0) { maxCount = 138; minCount = 3; }
    else if (curlen == nextlen) { maxCount = 6; minCount = 3; }
    else { maxCount = 7; minCount = 4; }
  }
}

// trees.c send_tree
template <class Out>
G4_HD inline void def_send_tree(DeflateState& s, Out& o, CtData* tree, int maxCode) {
  const CtData* bl = s.T->bltree;
  int prevlen = -1, nextlen = tree[0].dl, count = 0, maxCount = 7, minCount = 4;
  if (nextlen == 0) { maxCount = 138; minCount = 3; }
  for (int n = 0; n <= maxCode; n++) {
    int curlen = nextlen;
    nextlen = tree[n + 1].dl;
    if (++count < maxCount && curlen == nextlen) continue;
    else if (count < minCount) { do { o.send_bits(bl[curlen].fc, bl[curlen].dl); } while (--count != 0); }
    else if (curlen != 0) {
      if (curlen != prevlen) { o.send_bits(bl[curlen].fc, bl[curlen].dl); count--; }
      o.send_bits(bl[16].fc, bl[16].dl);
      o.send_bits(uint32_t(count - 3), 2);
    } else if (count <= 10) { o.send_bits(bl[17].fc, bl[17].dl); o.send_bits(uint32_t(count - 3), 3); }
    else { o.send_bits(bl[18].fc, bl[18].dl); o.send_bits(uint32_t(count - 11), 7); }
    count = 0;
    prevlen = curlen;
    if (nextlen == 0) { maxCount = 138; minCount = 3; }
    else if (curlen == nextlen) { maxCount = 6; minCount = 3; }
    else { maxCount = 7; minCount = 4; }
  }
}

G4_HD inline void def_init_block(DeflateState& s) {
  DeflateTrees* W = s.T;
  for (int n = 0; n < kDefLCodes; n++) W->ltree[n].fc = 0;
  for (int n = 0; n < kDefDCodes; n++) W->dtree[n].fc = 0;
  for (int n = 0; n < kDefBlCodes; n++) W->bltree[n].fc = 0;
  W->ltree[256].fc = 1;
  s.optLen = s.staticLen = 0;
}

// trees.c compress_block
G4_HD inline void def_compress_block(DeflateState& s, DeflateOut& o, int nSym, bool useStatic) {
  const DeflateTrees* W = s.T;
  for (int i = 0; i < nSym; i++) {
    unsigned dist = s.W->symDist[i];
    int lc = s.W->symLc[i];
    if (dist == 0) {
      if (useStatic) o.send_bits(def_static_lcode(lc), def_static_llen(lc));
      else o.send_bits(W->ltree[lc].fc, W->ltree[lc].dl);
    } else {
      int code = def_length_code(lc);
      int ls = code + 257;
      if (useStatic) o.send_bits(def_static_lcode(ls), def_static_llen(ls));
      else o.send_bits(W->ltree[ls].fc, W->ltree[ls].dl);
      int extra = def_length_extra(code);
      if (extra) o.send_bits(uint32_t(lc - def_length_base(code)), extra);
      dist--;
      code = def_dist_code(int(dist));
      if (useStatic) o.send_bits(def_bi_reverse(uint32_t(code), 5), 5);
      else o.send_bits(W->dtree[code].fc, W->dtree[code].dl);
      extra = def_dist_extra(code);
      if (extra) o.send_bits(uint32_t(int(dist) - def_dist_base(code)), extra);
    }
  }
  if (useStatic) o.send_bits(def_static_lcode(256), 7);
  else o.send_bits(W->ltree[256].fc, W->ltree[256].dl);
}

// trees.c _tr_flush_block, first half: Huffman trees of the block and the stored / fixed / dynamic choice.
struct DeflateBlockPlan {
  int type;  // 0 stored, 1 fixed, 2 dynamic
  int lMaxCode, dMaxCode, maxBlIndex;
};
G4_HD __forceinline__ int def_bl_order(int i) {
  const uint8_t blOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  return blOrder[i];
}
G4_HD inline DeflateBlockPlan def_plan_block(DeflateState& s, uint32_t storedLen, bool storedOk) {
  DeflateTrees* W = s.T;
  DeflateTreeDesc ld{W->ltree, kDefLCodes, 15, 257, 0, 0};
  DeflateTreeDesc dd{W->dtree, kDefDCodes, 15, 0, 1, 0};
  DeflateTreeDesc bd{W->bltree, kDefBlCodes, 7, 0, 2, 0};
  def_build_tree(s, ld);
  def_build_tree(s, dd);
  // build_bl_tree
  def_scan_tree(s, W->ltree, ld.maxCode);
  def_scan_tree(s, W->dtree, dd.maxCode);
  def_build_tree(s, bd);
  int maxBlIndex;
  for (maxBlIndex = kDefBlCodes - 1; maxBlIndex >= 3; maxBlIndex--)
    if (W->bltree[def_bl_order(maxBlIndex)].dl != 0) break;
  s.optLen += 3ull * (maxBlIndex + 1) + 5 + 5 + 4;
  unsigned long long optLenb = (s.optLen + 3 + 7) >> 3;
  unsigned long long staticLenb = (s.staticLen + 3 + 7) >> 3;
  if (staticLenb <= optLenb) optLenb = staticLenb;
  DeflateBlockPlan P;
  P.lMaxCode = ld.maxCode;
  P.dMaxCode = dd.maxCode;
  P.maxBlIndex = maxBlIndex;
  P.type = (storedLen + 4ull <= optLenb && storedOk) ? 0 : staticLenb == optLenb ? 1 : 2;
  return P;
}
// header of a dynamic block after the 3 type bits (trees.c send_all_trees)
template <class Out>
G4_HD inline void def_send_all_trees(DeflateState& s, Out& o, const DeflateBlockPlan& P) {
  DeflateTrees* W = s.T;
  o.send_bits(uint32_t(P.lMaxCode + 1 - 257), 5);
  o.send_bits(uint32_t(P.dMaxCode + 1 - 1), 5);
  o.send_bits(uint32_t(P.maxBlIndex + 1 - 4), 4);
  for (int rank = 0; rank <= P.maxBlIndex; rank++) o.send_bits(W->bltree[def_bl_order(rank)].dl, 3);
  def_send_tree(s, o, W->ltree, P.lMaxCode);
  def_send_tree(s, o, W->dtree, P.dMaxCode);
}

// trees.c _tr_flush_block
G4_HD inline void def_flush_block(DeflateState& s, DeflateOut& o, const uint8_t* buf, uint32_t storedLen, int nSym, bool last,
                                  bool storedOk) {
  const DeflateBlockPlan P = def_plan_block(s, storedLen, storedOk);
  if (P.type == 0) {
    o.send_bits(uint32_t(0 << 1) + (last ? 1u : 0u), 3);
    o.windup();
    o.put_byte(storedLen & 0xff);
    o.put_byte((storedLen >> 8) & 0xff);
    o.put_byte((~storedLen) & 0xff);
    o.put_byte(((~storedLen) >> 8) & 0xff);
    for (uint32_t i = 0; i < storedLen; i++) o.put_byte(buf[i]);
  } else if (P.type == 1) {
    o.send_bits((1u << 1) + (last ? 1u : 0u), 3);
    def_compress_block(s, o, nSym, true);
  } else {
    o.send_bits((2u << 1) + (last ? 1u : 0u), 3);
    def_send_all_trees(s, o, P);
    def_compress_block(s, o, nSym, false);
  }
  def_init_block(s);
  if (last) o.windup();
}

G4_HD __forceinline__ uint32_t def_hash3(const uint8_t* p) {
  return ((uint32_t(p[0]) << 10) ^ (uint32_t(p[1]) << 5) ^ uint32_t(p[2])) & kDefHashMask;
}

// Compresses in[0..n) as one zlib stream into out[0..cap).  Returns the number of bytes written; a stream that
// does not fit is truncated at `cap` exactly as the reference's single deflate() call into a fixed buffer
// truncates it (CodecDeflate.java:209-213).  The caller guarantees in[] is readable up to in[n + 7] (bytes past
// n never become part of a match, but the candidate pre-filter may look at in[n] and in[n + 1] as zlib's does).
G4_HD inline uint32_t deflate_stream(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t cap, DeflateWork* W, int level) {
  const DeflateLevel L = deflate_level(level);
  DeflateState s;
  s.W = W;
  s.T = &W->T;
  DeflateOut o{out, cap, 0, 0, 0, false};
  o.put_byte(L.zlibHeader >> 8);
  o.put_byte(L.zlibHeader & 0xff);
  for (int i = 0; i < kDefWSize; i++) W->head[i] = 0;
  def_init_block(s);
  uint32_t strstart = 0, blockStart = 0, matchStart = 0, prevMatch = 0;
  int matchLength = kDefMinMatch - 1, prevLength = kDefMinMatch - 1;
  bool matchAvailable = false;
  int nSym = 0;
  // zlib's 64 KiB sliding window, tracked in absolute stream positions: winBase = position of window[0],
  // winFill = end of the data fill_window() has loaded.  They only matter for (a) `lookahead`, which is what is
  // loaded, not what remains, and (b) whether a stored block is still possible (block start inside the window).
  uint32_t winBase = 0;
  uint32_t winFill = n < 2u * kDefWSize ? n : 2u * kDefWSize;
  auto insert_string = [&](uint32_t pos) -> uint32_t {
    uint32_t h = def_hash3(in + pos);
    uint32_t hh = W->head[h];
    W->prev[pos & kDefWMask] = hh;
    W->head[h] = pos;
    return hh;
  };
  auto tally_lit = [&](uint32_t c) -> bool {
    W->symDist[nSym] = 0;
    W->symLc[nSym] = uint8_t(c);
    nSym++;
    W->T.ltree[c].fc++;
    return nSym == kDefLitBufSize - 1;
  };
  auto tally_dist = [&](uint32_t dist, uint32_t lc) -> bool {
    W->symDist[nSym] = uint16_t(dist);
    W->symLc[nSym] = uint8_t(lc);
    nSym++;
    dist--;
    W->T.ltree[def_length_code(int(lc)) + 257].fc++;
    W->T.dtree[def_dist_code(int(dist))].fc++;
    return nSym == kDefLitBufSize - 1;
  };
  auto flush = [&](bool last) {
    // _tr_flush_block gets buf == NULL when the block started before the current window (block_start < 0)
    const bool storedOk = blockStart >= winBase;
    def_flush_block(s, o, in + blockStart, strstart - blockStart, nSym, last, storedOk);
    blockStart = strstart;
    nSym = 0;
  };
  for (;;) {
    if (winFill - strstart < uint32_t(kDefMinLookahead) && winFill < n) {  // fill_window(): slide by 32 KiB, load more
      if (strstart - winBase >= uint32_t(kDefWSize + kDefMaxDist)) {
        winBase += kDefWSize;
        // positions below the new base leave the hash chains; the entry that lands on relative 0 reads as NIL
        // (both are beyond MAX_DIST of every later strstart, so the chain walk's limit test already excludes them)
      }
      winFill = n - winBase < 2u * kDefWSize ? n : winBase + 2u * kDefWSize;
    }
    const uint32_t lookahead = winFill - strstart;
    if (lookahead == 0) break;
    uint32_t hashHead = 0;
    if (lookahead >= uint32_t(kDefMinMatch)) hashHead = insert_string(strstart);
    prevLength = matchLength;
    prevMatch = matchStart;
    matchLength = kDefMinMatch - 1;
    if (hashHead > winBase && prevLength < L.maxLazy && strstart - hashHead <= uint32_t(kDefMaxDist)) {
      // longest_match
      int chainLength = L.maxChain;
      int bestLen = prevLength;
      int niceMatch = L.niceLength;
      uint32_t limit = strstart > uint32_t(kDefMaxDist) ? strstart - uint32_t(kDefMaxDist) : 0u;
      if (limit < winBase) limit = winBase;
      if (prevLength >= L.goodLength) chainLength >>= 2;
      if (uint32_t(niceMatch) > lookahead) niceMatch = int(lookahead);
      // zlib compares up to MAX_MATCH bytes into whatever follows the data and clamps afterwards; stopping at
      // `lookahead` is equivalent (a candidate that reaches it also reaches nice_match and ends the walk)
      const int maxLen = lookahead < uint32_t(kDefMaxMatch) ? int(lookahead) : kDefMaxMatch;
      const uint8_t* scan = in + strstart;
      uint8_t scanEnd1 = scan[bestLen - 1], scanEnd = scan[bestLen];
      uint32_t cur = hashHead;
      // The next chain link is loaded BEFORE the candidate is examined: both loads depend only on `cur`, and a GPU
      // thread does not speculate past the candidate's branches, so this halves the dependent memory latency per link.
      for (;;) {
        const uint32_t nextLink = W->prev[cur & kDefWMask];
        const uint8_t* match = in + cur;
        if (match[bestLen] == scanEnd && match[bestLen - 1] == scanEnd1 && match[0] == scan[0] && match[1] == scan[1]) {
          int len = 2;
          while (len < maxLen && scan[len] == match[len]) len++;
          if (len > bestLen) {
            matchStart = cur;
            bestLen = len;
            if (len >= niceMatch) break;
            scanEnd1 = scan[bestLen - 1];
            scanEnd = scan[bestLen];
          }
        }
        cur = nextLink;
        if (!(cur > limit && --chainLength != 0)) break;
      }
      matchLength = uint32_t(bestLen) <= lookahead ? bestLen : int(lookahead);
      if (matchLength <= 5 && (matchLength == kDefMinMatch && strstart - matchStart > uint32_t(kDefTooFar))) matchLength = kDefMinMatch - 1;
    }
    if (prevLength >= kDefMinMatch && matchLength <= prevLength) {
      const uint32_t maxInsert = strstart + lookahead - kDefMinMatch;
      bool bflush = tally_dist(strstart - 1 - prevMatch, uint32_t(prevLength - kDefMinMatch));
      int pl = prevLength - 2;
      do {
        if (++strstart <= maxInsert) insert_string(strstart);
      } while (--pl != 0);
      matchAvailable = false;
      matchLength = kDefMinMatch - 1;
      strstart++;
      if (bflush) flush(false);
    } else if (matchAvailable) {
      bool bflush = tally_lit(in[strstart - 1]);
      if (bflush) flush(false);
      strstart++;
    } else {
      matchAvailable = true;
      strstart++;
    }
  }
  if (matchAvailable) tally_lit(in[strstart - 1]);
  flush(true);
  // Adler-32 trailer, big-endian
  uint32_t s1 = 1, s2 = 0;
  for (uint32_t i = 0; i < n;) {
    uint32_t chunk = n - i < 5552u ? n - i : 5552u;
    for (uint32_t k = 0; k < chunk; k++) { s1 += in[i + k]; s2 += s1; }
    s1 %= 65521u;
    s2 %= 65521u;
    i += chunk;
  }
  o.put_byte(s2 >> 8);
  o.put_byte(s2 & 0xff);
  o.put_byte(s1 >> 8);
  o.put_byte(s1 & 0xff);
  return o.overflow ? cap : o.pos;
}

// =====================================================================================================================
// Staged form of the same encoder for streams of at most kDefStagedMax bytes (no window slide: zlib loads the whole
// input at once and positions fit 16 bits).  deflate_slow inserts EVERY position p <= n-3 into the hash chains, in
// order, whether or not it searches there, so the chain of p -- "earlier positions with the same 3-byte hash, nearest
// first" -- is a property of the input alone.  That splits the work into
//   sort    positions ordered by (hash, position): the chain of the position in slot i is slots i-1, i-2, ... of its
//           bucket (rank[i] = number of earlier positions in the bucket),
//   match   for EVERY position, in parallel, longest_match() over that list: once with the level's full chain length
//           and once with a quarter of it (the walk zlib does when prev_length >= good_match), started from
//           best_len = 2.  Starting from 2 instead of prev_length changes nothing that the caller can observe: the
//           walk visits the same candidates and stops at the same nice_match candidate, so its result is the
//           first-in-chain longest candidate, which zlib's walk also ends with whenever that is longer than prev_length
//           -- and when it is not, both forms leave match_length <= prev_length and the previous match is emitted,
//   emit    the serial lazy-evaluation loop (one thread per stream) reads the table instead of walking chains, and
//           produces the same symbols, blocks and bytes as deflate_stream().
// =====================================================================================================================
constexpr uint32_t kDefStagedMax = 65535;

// Table entry: match length (9 bits, 2 = none) | distance << 9.
G4_HD __forceinline__ uint32_t def_pack_match(int len, uint32_t dist) { return uint32_t(len) | (dist << 9); }

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t def_load32(const uint8_t* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
  return __funnelshift_r(q[0], q[1], uint32_t(a & 3u) * 8u);
}
#endif

// longest_match for the position in `slot` of the stream's sorted position list.  x = full chain, y = quarter chain.
G4_HD inline uint2 def_find_match(const uint8_t* in, uint32_t n, const DeflateLevel& L, const uint16_t* sorted, uint32_t slot,
                                  uint32_t rank) {
  const uint32_t p = sorted[slot];
  const uint32_t lookahead = n - p;
  const int maxLen = lookahead < uint32_t(kDefMaxMatch) ? int(lookahead) : kDefMaxMatch;
  const int niceMatch = uint32_t(L.niceLength) > lookahead ? int(lookahead) : L.niceLength;
  const uint32_t limit = p > uint32_t(kDefMaxDist) ? p - uint32_t(kDefMaxDist) : 0u;
  const uint32_t nCand = rank < uint32_t(L.maxChain) ? rank : uint32_t(L.maxChain);
  const uint32_t quarter = uint32_t(L.maxChain) >> 2;
  const uint8_t* scan = in + p;
  int bestLen = kDefMinMatch - 1;
  uint32_t bestStart = 0;
  uint8_t scanEnd1 = scan[bestLen - 1], scanEnd = scan[bestLen];
  const uint8_t s0 = scan[0], s1 = scan[1];
  uint32_t quarterEntry = 0;
  bool haveQuarter = false;
  for (uint32_t k = 1; k <= nCand; k++) {
    if (k - 1 == quarter) { quarterEntry = def_pack_match(bestLen, p - bestStart); haveQuarter = true; }
    const uint32_t cur = sorted[slot - k];
    // the first candidate may sit at distance MAX_DIST exactly (deflate_slow tests strstart - hash_head <= MAX_DIST),
    // later ones must be nearer (longest_match continues while cur_match > limit); position 0 doubles as NIL
    if (k == 1 ? (cur == 0 || p - cur > uint32_t(kDefMaxDist)) : cur <= limit) break;
    const uint8_t* match = in + cur;
    if (match[bestLen] == scanEnd && match[bestLen - 1] == scanEnd1 && match[0] == s0 && match[1] == s1) {
      int len = 2;
#ifdef __CUDA_ARCH__
      // four bytes per step: unaligned 32-bit reads assembled from aligned words (in[] is padded past n)
      bool differ = false;
      while (len + 4 <= maxLen) {
        const uint32_t x = def_load32(scan + len) ^ def_load32(match + len);
        if (x) { len += (__ffs(int(x)) - 1) >> 3; differ = true; break; }
        len += 4;
      }
      if (!differ)
#endif
      while (len < maxLen && scan[len] == match[len]) len++;
      if (len > bestLen) {
        bestStart = cur;
        bestLen = len;
        if (len >= niceMatch) break;
        scanEnd1 = scan[bestLen - 1];
        scanEnd = scan[bestLen];
      }
    }
  }
  const uint32_t full = def_pack_match(bestLen, p - bestStart);
  return make_uint2(full, haveQuarter ? quarterEntry : full);
}

// deflate_stream() with longest_match replaced by the table (table[p] for every p <= n-3).  n <= kDefStagedMax.
G4_HD inline uint32_t deflate_stream_table(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t cap, DeflateWork* W, int level,
                                           const uint2* table) {
  const DeflateLevel L = deflate_level(level);
  DeflateState s;
  s.W = W;
  s.T = &W->T;
  DeflateOut o{out, cap, 0, 0, 0, false};
  o.put_byte(L.zlibHeader >> 8);
  o.put_byte(L.zlibHeader & 0xff);
  def_init_block(s);
  uint32_t strstart = 0, blockStart = 0, matchStart = 0, prevMatch = 0;
  int matchLength = kDefMinMatch - 1, prevLength = kDefMinMatch - 1;
  bool matchAvailable = false;
  int nSym = 0;
  auto tally_lit = [&](uint32_t c) -> bool {
    W->symDist[nSym] = 0;
    W->symLc[nSym] = uint8_t(c);
    nSym++;
    W->T.ltree[c].fc++;
    return nSym == kDefLitBufSize - 1;
  };
  auto tally_dist = [&](uint32_t dist, uint32_t lc) -> bool {
    W->symDist[nSym] = uint16_t(dist);
    W->symLc[nSym] = uint8_t(lc);
    nSym++;
    dist--;
    W->T.ltree[def_length_code(int(lc)) + 257].fc++;
    W->T.dtree[def_dist_code(int(dist))].fc++;
    return nSym == kDefLitBufSize - 1;
  };
  auto flush = [&](bool last) {
    def_flush_block(s, o, in + blockStart, strstart - blockStart, nSym, last, true);
    blockStart = strstart;
    nSym = 0;
  };
  while (strstart < n) {
    const uint32_t lookahead = n - strstart;
    prevLength = matchLength;
    prevMatch = matchStart;
    matchLength = kDefMinMatch - 1;
    if (lookahead >= uint32_t(kDefMinMatch) && prevLength < L.maxLazy) {
      const uint2 e = table[strstart];
      const uint32_t w = prevLength >= L.goodLength ? e.y : e.x;
      const int len = int(w & 0x1ffu);
      if (len > prevLength) {
        matchLength = len;
        matchStart = strstart - (w >> 9);
        if (matchLength == kDefMinMatch && strstart - matchStart > uint32_t(kDefTooFar)) matchLength = kDefMinMatch - 1;
      }
    }
    if (prevLength >= kDefMinMatch && matchLength <= prevLength) {
      const bool bflush = tally_dist(strstart - 1 - prevMatch, uint32_t(prevLength - kDefMinMatch));
      strstart += uint32_t(prevLength - 1);
      matchAvailable = false;
      matchLength = kDefMinMatch - 1;
      if (bflush) flush(false);
    } else if (matchAvailable) {
      const bool bflush = tally_lit(in[strstart - 1]);
      if (bflush) flush(false);
      strstart++;
    } else {
      matchAvailable = true;
      strstart++;
    }
  }
  if (matchAvailable) tally_lit(in[strstart - 1]);
  flush(true);
  uint32_t s1 = 1, s2 = 0;
  for (uint32_t i = 0; i < n;) {
    uint32_t chunk = n - i < 5552u ? n - i : 5552u;
    for (uint32_t k = 0; k < chunk; k++) { s1 += in[i + k]; s2 += s1; }
    s1 %= 65521u;
    s2 %= 65521u;
    i += chunk;
  }
  o.put_byte(s2 >> 8);
  o.put_byte(s2 & 0xff);
  o.put_byte(s1 >> 8);
  o.put_byte(s1 & 0xff);
  return o.overflow ? cap : o.pos;
}

// ---- staged form, split for the device: decisions (one thread per stream) / block emission (one CTA per stream) --------
constexpr int kDefMaxBlocks = 8;  // <= 65535 symbols in blocks of 16383, plus a possibly empty final block
struct DeflateBlocks {
  uint32_t nBlocks;
  uint32_t symEnd[kDefMaxBlocks];  // symbols [symEnd[b-1], symEnd[b]) belong to block b
  uint32_t posEnd[kDefMaxBlocks];  // input bytes [posEnd[b-1], posEnd[b]) are what a stored block b would carry
};

// The lazy-evaluation loop of deflate_stream_table() alone: writes the symbol list (dist 0 = literal) and the block
// boundaries zlib's symbol buffer (16383 entries) imposes.
G4_HD inline void deflate_decide_table(const uint8_t* in, uint32_t n, int level, const uint2* table, uint16_t* symDist, uint8_t* symLc,
                                       DeflateBlocks* B) {
  const DeflateLevel L = deflate_level(level);
  uint32_t strstart = 0, matchStart = 0, prevMatch = 0, k = 0, inBlock = 0, nBlocks = 0;
  int matchLength = kDefMinMatch - 1, prevLength = kDefMinMatch - 1;
  bool matchAvailable = false;
  auto tally = [&](uint32_t dist, uint32_t lc) -> bool {
    symDist[k] = uint16_t(dist);
    symLc[k] = uint8_t(lc);
    k++;
    return ++inBlock == uint32_t(kDefLitBufSize - 1);
  };
  auto flush = [&]() {
    B->symEnd[nBlocks] = k;
    B->posEnd[nBlocks] = strstart;
    nBlocks++;
    inBlock = 0;
  };
  while (strstart < n) {
    const uint32_t lookahead = n - strstart;
    prevLength = matchLength;
    prevMatch = matchStart;
    matchLength = kDefMinMatch - 1;
    if (lookahead >= uint32_t(kDefMinMatch) && prevLength < L.maxLazy) {
      const uint2 e = table[strstart];
      const uint32_t w = prevLength >= L.goodLength ? e.y : e.x;
      const int len = int(w & 0x1ffu);
      if (len > prevLength) {
        matchLength = len;
        matchStart = strstart - (w >> 9);
        if (matchLength == kDefMinMatch && strstart - matchStart > uint32_t(kDefTooFar)) matchLength = kDefMinMatch - 1;
      }
    }
    if (prevLength >= kDefMinMatch && matchLength <= prevLength) {
      const bool bflush = tally(strstart - 1 - prevMatch, uint32_t(prevLength - kDefMinMatch));
      strstart += uint32_t(prevLength - 1);
      matchAvailable = false;
      matchLength = kDefMinMatch - 1;
      if (bflush) flush();
    } else if (matchAvailable) {
      const bool bflush = tally(0, in[strstart - 1]);
      if (bflush) flush();
      strstart++;
    } else {
      matchAvailable = true;
      strstart++;
    }
  }
  if (matchAvailable) tally(0, in[strstart - 1]);
  flush();
  B->nBlocks = nBlocks;
}

// Serial emission of the blocks deflate_decide_table() produced (host harness; the device emits them CTA-parallel in
// g4_deflate_encode.cu with the same plan / header / code functions).
inline uint32_t deflate_emit_blocks_serial(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t cap, DeflateWork* W, int level,
                                           const uint16_t* symDist, const uint8_t* symLc, const DeflateBlocks& B) {
  const DeflateLevel L = deflate_level(level);
  DeflateState s;
  s.W = W;
  s.T = &W->T;
  DeflateOut o{out, cap, 0, 0, 0, false};
  o.put_byte(L.zlibHeader >> 8);
  o.put_byte(L.zlibHeader & 0xff);
  uint32_t sym0 = 0, pos0 = 0;
  for (uint32_t b = 0; b < B.nBlocks; b++) {
    def_init_block(s);
    const uint32_t nSym = B.symEnd[b] - sym0;
    for (uint32_t i = 0; i < nSym; i++) {
      const uint32_t dist = symDist[sym0 + i], lc = symLc[sym0 + i];
      W->symDist[i] = uint16_t(dist);
      W->symLc[i] = uint8_t(lc);
      if (dist == 0) W->T.ltree[lc].fc++;
      else { W->T.ltree[def_length_code(int(lc)) + 257].fc++; W->T.dtree[def_dist_code(int(dist - 1))].fc++; }
    }
    def_flush_block(s, o, in + pos0, B.posEnd[b] - pos0, int(nSym), b + 1 == B.nBlocks, true);
    sym0 = B.symEnd[b];
    pos0 = B.posEnd[b];
  }
  uint32_t s1 = 1, s2 = 0;
  for (uint32_t i = 0; i < n; i++) { s1 = (s1 + in[i]) % 65521u; s2 = (s2 + s1) % 65521u; }
  o.put_byte(s2 >> 8);
  o.put_byte(s2 & 0xff);
  o.put_byte(s1 >> 8);
  o.put_byte(s1 & 0xff);
  return o.overflow ? cap : o.pos;
}

// Serial construction of the sorted position list (host harness / reference for the warp-cooperative device version).
inline void def_sort_positions_host(const uint8_t* in, uint32_t n, uint16_t* sorted, uint16_t* rank) {
  if (n < 3) return;
  const uint32_t nPos = n - 2;
  static uint32_t cnt[kDefWSize + 1];
  for (int i = 0; i <= kDefWSize; i++) cnt[i] = 0;
  for (uint32_t p = 0; p < nPos; p++) cnt[def_hash3(in + p) + 1]++;
  for (int i = 0; i < kDefWSize; i++) cnt[i + 1] += cnt[i];
  static uint32_t start[kDefWSize];
  for (int i = 0; i < kDefWSize; i++) start[i] = cnt[i];
  for (uint32_t p = 0; p < nPos; p++) {
    const uint32_t h = def_hash3(in + p);
    const uint32_t slot = cnt[h]++;
    sorted[slot] = uint16_t(p);
    rank[slot] = uint16_t(slot - start[h]);
  }
}

}  // namespace g4
