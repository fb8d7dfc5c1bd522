// g4_canon_enc.cuh -- canonical Huffman ENCODER (260-symbol integer alphabet), CTA-cooperative.
//
// Reference (under /root/reference/core/src/main/java/org/gridfour/compress/canonicalHuffman/):
//   CanonicalHuffman.java:177-343 (encode, buildCodeLengthTree), :352-418 (countSymbols)
//   TreeBuilder.java:75-301 (sort by count asc / symbol DESC, list-insertion merge, depth walk, canonical codes)
//   PackageMerge.java:91-175 (length limit 15), LengthEncoder.java:86-195, HuffmanCodeBits.java:47-74
// The stream produced here is byte-identical to the reference's for the same input values.
#pragma once
#include "g4_bitpack.cuh"
#include "g4_canon.cuh"

namespace g4 {

constexpr int kCtSymbols = 20;  // 19 length symbols + the code table's own end-of-text

struct CanonEncShared {
  uint32_t hist[kCanonSymbols + 4];
  uint32_t skey[kCanonSymbols + 4];   // (count << 9) | (511 - symbol), ascending == count asc, symbol desc
  uint16_t order[kCanonSymbols + 4];  // symbols in sorted order (only the first nUsed are meaningful)
  uint32_t bcount[kCanonSymbols];
  uint16_t bqueue[kCanonSymbols];
  uint16_t left[kCanonSymbols], right[kCanonSymbols];
  uint8_t depth[2 * kCanonSymbols + 8];
  uint8_t len[kCanonSymbols + 4];
  uint16_t rcode[kCanonSymbols + 4];  // bit-reversed canonical code: append `len` bits LSB-first
  uint8_t tcodes[kCanonSymbols + 4], truns[kCanonSymbols + 4];
  uint8_t ctLen[kCtSymbols + 4];
  uint16_t ctRcode[kCtSymbols + 4];
  uint8_t ccodes[kCtSymbols + 4], cruns[kCtSymbols + 4];
  int nT, nC, nUsed;
  uint32_t headerBits;
  uint32_t rawBits;
  unsigned long long totalBits;
  uint32_t scan[kWarps + 1];
};

// Symbols and raw escape bits of one value (CanonicalHuffman.java:211-274 / :352-418).
// Returns the base symbol; *nEsc2 / *nEsc8 = number of 2-bit / 8-bit escapes that follow.
__device__ __forceinline__ int canon_classify(int32_t s, int* nEsc2, int* nEsc8) {
  *nEsc2 = 0;
  *nEsc8 = 0;
  if (s >= -128 && s <= 127) return s + 128;
  if (s >= -512 && s <= 511) { *nEsc2 = 1; return (s >> 2) + 128; }
  if (s >= -2048 && s <= 2047) { *nEsc2 = 2; return (s >> 4) + 128; }
  if (s >= -8192 && s <= 8191) { *nEsc2 = 3; return (s >> 6) + 128; }
  if (s >= -32768 && s <= 32767) { *nEsc8 = 1; return (s >> 8) + 128; }
  if (s == INT32_MIN) return kSymNull;
  if (s >= -8388608 && s <= 8388607) { *nEsc8 = 2; return (s >> 16) + 128; }
  *nEsc8 = 3;
  return (s >> 24) + 128;
}
// The reference's encode() tests -8333608 where countSymbols() tests -8388608 (CanonicalHuffman.java:258 vs
// :395): values in [-8388608, -8333609] are COUNTED as 16-bit escapes (symbol (s>>16)+128, two byte escapes) but WRITTEN
// as 24-bit escapes (symbol (s>>24)+128 = 127, three byte escapes).  The text still decodes as long as symbol 127 has a
// code (it is the symbol of the value -1), so the GPU path does the same: the histogram follows countSymbols, the text
// pass follows encode (canon_classify_emit), and canon_reconcile_escapes puts the size right and declines the stream
// when symbol 127 has no code (the reference would write a text that cannot be decoded).
__device__ __forceinline__ bool canon_value_hits_reference_bug(int32_t s) { return s >= -8388608 && s <= -8333609; }
__device__ __forceinline__ int canon_classify_emit(int32_t s, int* nEsc2, int* nEsc8) {
  if (canon_value_hits_reference_bug(s)) { *nEsc2 = 0; *nEsc8 = 3; return (s >> 24) + 128; }
  return canon_classify(s, nEsc2, nEsc8);
}

// ---- serial pieces (one thread) ----------------------------------------------------------------------
// Huffman code lengths for the symbols listed in order[0..k) (ascending keys).  TreeBuilder.java:132-178.
// scratch: global memory for the package-merge fallback (>= 15*2*k*2 + 2*2*k*4 bytes).
__device__ inline void canon_tree_lengths_serial(CanonEncShared& S, const uint32_t* counts, int k, uint8_t* lenOut,
                                                 uint8_t* pmScratch) {
  // list-insertion merge with the reference's tie rule (same as the legacy encoder)
  int li = 0, bi = 0, bt = 0, nb = 0;
  for (int m = 0; m < k - 1; m++) {
    uint16_t node[2];
    uint32_t cnt[2];
    for (int q = 0; q < 2; q++) {
      bool takeBranch = (bi < bt) && (li >= k || S.bcount[S.bqueue[bi]] <= counts[S.order[li]]);
      if (takeBranch) { uint16_t id = S.bqueue[bi++]; node[q] = uint16_t(kCanonSymbols + id); cnt[q] = S.bcount[id]; }
      else { node[q] = uint16_t(li); cnt[q] = counts[S.order[li]]; li++; }  // leaves are referenced by sorted position
    }
    int id = nb++;
    uint32_t s = cnt[0] + cnt[1];
    S.bcount[id] = s;
    S.left[id] = node[0];
    S.right[id] = node[1];
    int pos = bt;
    while (pos > bi && S.bcount[S.bqueue[pos - 1]] >= s) { S.bqueue[pos] = S.bqueue[pos - 1]; pos--; }
    S.bqueue[pos] = uint16_t(id);
    bt++;
  }
  int maxLen = 0;
  S.depth[kCanonSymbols + (k - 2)] = 0;
  for (int id = k - 2; id >= 0; id--) {
    uint8_t d = S.depth[kCanonSymbols + id] + 1;
    S.depth[S.left[id]] = d;
    S.depth[S.right[id]] = d;
    if (S.left[id] < kCanonSymbols && d > maxLen) maxLen = d;
    if (S.right[id] < kCanonSymbols && d > maxLen) maxLen = d;
  }
  if (maxLen <= 15) {
    for (int i = 0; i < k; i++) lenOut[S.order[i]] = S.depth[i];
    return;
  }
  // PackageMerge.merge(15, sortNodes) (PackageMerge.java:91-175).  Entry "symbol" = position in the sorted
  // array, so the (count, index) sort is the identity.  Level arrays live in global scratch.
  const int B = k;
  int16_t* kind = reinterpret_cast<int16_t*>(pmScratch);                // [15][2B]: base index or -1 (package) or -2 (unset)
  uint32_t* cntA = reinterpret_cast<uint32_t*>(pmScratch + 15 * 2 * B * 2 + 16);
  uint32_t* cntB = cntA + 2 * B;
  int levelLen[15];
  levelLen[0] = B;
  for (int i = 0; i < B; i++) { kind[i] = int16_t(i); cntA[i] = counts[S.order[i]]; }
  uint32_t* prev = cntA;
  uint32_t* cur = cntB;
  for (int d = 1; d < 15; d++) {
    int16_t* kd = kind + d * 2 * B;
    int nPrev = levelLen[d - 1];
    int nPair = nPrev / 2;
    int mLen = B + nPair;
    for (int i = 0; i < mLen; i++) kd[i] = -2;
    int kk = 0, iBase = 0;
    for (int iPair = 0; iPair < nPair; iPair++) {
      uint32_t pc = prev[2 * iPair] + prev[2 * iPair + 1];
      while (iBase < B && counts[S.order[iBase]] <= pc) { kd[kk] = int16_t(iBase); cur[kk] = counts[S.order[iBase]]; kk++; iBase++; }
      kd[kk] = -1;
      cur[kk] = pc;
      kk++;
    }
    uint32_t lastPair = prev[2 * (nPair - 1)] + prev[2 * (nPair - 1) + 1];
    if (counts[S.order[B - 1]] > lastPair) { kd[mLen - 1] = int16_t(B - 1); cur[mLen - 1] = counts[S.order[B - 1]]; }  // :145-147
    levelLen[d] = mLen;
    uint32_t* t = prev; prev = cur; cur = t;
  }
  for (int i = 0; i < B; i++) S.depth[i] = 0;
  int nn = 2 * B - 2;
  for (int e = 14; e >= 0; e--) {
    const int16_t* kd = kind + e * 2 * B;
    int nMerged = 0;
    for (int i = 0; i < nn; i++) {
      if (kd[i] == -1) nMerged++;
      else if (kd[i] >= 0) S.depth[kd[i]]++;
    }
    nn = nMerged * 2;
  }
  for (int i = 0; i < B; i++) lenOut[S.order[i]] = S.depth[i];
}

// Canonical codes in (length, symbol) order (TreeBuilder.java:283-301), stored bit-reversed.
__device__ inline void canon_assign_codes(const uint8_t* len, int nSym, uint16_t* rcode) {
  uint16_t cnt[17], next[17];
  for (int l = 0; l <= 16; l++) cnt[l] = 0;
  for (int i = 0; i < nSym; i++) cnt[len[i]]++;
  uint32_t code = 0;
  cnt[0] = 0;
  for (int l = 1; l <= 15; l++) { next[l] = uint16_t(code); code = (code + cnt[l]) << 1; }
  for (int i = 0; i < nSym; i++) {
    int l = len[i];
    rcode[i] = l ? uint16_t(__brev(uint32_t(next[l]++)) >> (32 - l)) : 0;
  }
}

// LengthEncoder.encodeLengths (LengthEncoder.java:86-167)
__device__ inline int canon_length_encode(int n, const uint8_t* codeLen, uint8_t* codes, uint8_t* runs) {
  int prior = -1, nc = 0, i;
  for (int ic = 0; ic < n; ic++) {
    if (codeLen[ic] == 0) {
      prior = 0;
      for (i = ic + 1; i < n; i++) if (codeLen[i] != 0) break;
      int nZero = i - ic;
      if (nZero == 1) { codes[nc] = 0; runs[nc++] = 0; }
      else if (nZero == 2) { codes[nc] = 0; runs[nc++] = 0; codes[nc] = 0; runs[nc++] = 0; ic++; }
      else if (nZero <= 10) { codes[nc] = 17; runs[nc++] = uint8_t(nZero - 3); ic = i - 1; }
      else {
        if (nZero > 138) nZero = 138;
        codes[nc] = 18; runs[nc++] = uint8_t(nZero - 11);
        ic += nZero - 1;
      }
    } else if (codeLen[ic] == prior) {
      for (i = ic + 1; i < n; i++) if (codeLen[i] != prior) break;
      int nPrior = i - ic;
      if (nPrior == 1) { codes[nc] = uint8_t(prior); runs[nc++] = 0; }
      else if (nPrior == 2) { codes[nc] = uint8_t(prior); runs[nc++] = 0; codes[nc] = uint8_t(prior); runs[nc++] = 0; ic = i - 1; }
      else {
        if (nPrior > 6) nPrior = 6;
        codes[nc] = 16; runs[nc++] = uint8_t(nPrior - 3);
        ic += nPrior - 1;
      }
    } else {
      prior = codeLen[ic];
      codes[nc] = uint8_t(prior); runs[nc++] = 0;
    }
  }
  return nc;
}

__device__ __forceinline__ int canon_run_extra_bits(int code) { return code == 16 ? 2 : code == 17 ? 3 : code == 18 ? 7 : 0; }

// Sorts the used symbols of `counts[0..nSym)` into S.order (count asc, symbol desc).  All threads call.
__device__ inline int canon_sort_symbols(CanonEncShared& S, const uint32_t* counts, int nSym) {
  const int tid = threadIdx.x;
  __syncthreads();
  for (int i = tid; i < nSym; i += kThreads) S.skey[i] = counts[i] ? ((counts[i] << 9) | uint32_t(511 - i)) : 0xffffffffu;
  __syncthreads();
  for (int i = tid; i < nSym; i += kThreads) {
    uint32_t key = S.skey[i];
    if (key != 0xffffffffu) {
      int rank = 0;
      for (int j = 0; j < nSym; j++) rank += S.skey[j] < key;
      S.order[rank] = uint16_t(i);
    }
  }
  int used = 0;
  for (int i = tid; i < nSym; i += kThreads) used += counts[i] != 0;
  __shared__ int sUsed;
  if (tid == 0) sUsed = 0;
  __syncthreads();
  if (used) atomicAdd(&sUsed, used);
  __syncthreads();
  return sUsed;
}

// Everything after the histogram: code lengths, codes, table coding, sizes.  S.hist must be complete
// (including hist[kSymEot] = 1) and S.rawBits set.  All threads call; results are in S.
__device__ inline void canon_build_code(CanonEncShared& S, uint8_t* pmScratch) {
  const int tid = threadIdx.x;
  int used = canon_sort_symbols(S, S.hist, kCanonSymbols);
  if (tid == 0) {
    S.nUsed = used;
    for (int i = 0; i < kCanonSymbols; i++) S.len[i] = 0;
    canon_tree_lengths_serial(S, S.hist, used, S.len, pmScratch);
    canon_assign_codes(S.len, kCanonSymbols, S.rcode);
    S.nT = canon_length_encode(kCanonSymbols, S.len, S.tcodes, S.truns);
  }
  __syncthreads();
  // code-table tree over the 19 length symbols + end-of-text (CanonicalHuffman.java:285-301)
  __shared__ uint32_t ctCounts[kCtSymbols + 4];
  if (tid < kCtSymbols) ctCounts[tid] = tid == kCtSymbols - 1 ? 1u : 0u;
  __syncthreads();
  if (tid == 0)
    for (int i = 0; i < S.nT; i++) ctCounts[S.tcodes[i]]++;
  __syncthreads();
  int ctUsed = canon_sort_symbols(S, ctCounts, kCtSymbols);
  if (tid == 0) {
    for (int i = 0; i < kCtSymbols; i++) S.ctLen[i] = 0;
    canon_tree_lengths_serial(S, ctCounts, ctUsed, S.ctLen, pmScratch);
    canon_assign_codes(S.ctLen, kCtSymbols, S.ctRcode);
    S.nC = canon_length_encode(kCtSymbols, S.ctLen, S.ccodes, S.cruns);
    uint32_t hb = 1;
    for (int i = 0; i < S.nC; i++) hb += 5 + canon_run_extra_bits(S.ccodes[i]);
    for (int i = 0; i < S.nT; i++) hb += S.ctLen[S.tcodes[i]] + canon_run_extra_bits(S.tcodes[i]);
    S.headerBits = hb;
    unsigned long long tb = hb + S.rawBits;
    for (int i = 0; i < kCanonSymbols; i++) tb += (unsigned long long)(S.hist[i]) * S.len[i];
    S.totalBits = tb;
  }
  __syncthreads();
}

// Histogram of one value stream.  get(k) returns value k.  Returns false if a value hits the reference's
// inconsistent escape range.  All threads call.
template <class Get>
__device__ inline bool canon_histogram(CanonEncShared& S, Get get, uint32_t N) {
  const int tid = threadIdx.x;
  __syncthreads();
  for (int i = tid; i < kCanonSymbols + 4; i += kThreads) S.hist[i] = 0;
  if (tid == 0) S.rawBits = 0;
  __syncthreads();
  uint32_t raw = 0;
  bool bug = false;
  for (uint32_t k = tid; k < N; k += kThreads) {
    int32_t v = get(k);
    if (canon_value_hits_reference_bug(v)) bug = true;
    int e2, e8;
    int sym = canon_classify(v, &e2, &e8);
    atomicAdd(&S.hist[sym], 1u);
    if (e2) { atomicAdd(&S.hist[kSymEsc2], uint32_t(e2)); raw += 2u * e2; }
    if (e8) { atomicAdd(&S.hist[kSymEsc8], uint32_t(e8)); raw += 8u * e8; }
  }
  if (raw) atomicAdd(&S.rawBits, raw);
  if (tid == 0) S.hist[kSymEot] = 1;
  return __syncthreads_or(bug ? 1 : 0) == 0;
}

// For a stream whose histogram pass returned false (values in the inconsistent escape range), after canon_build_code:
// S.totalBits becomes what the text pass will write.  Returns false when the stream cannot be written.  All threads call.
template <class Get>
__device__ inline bool canon_reconcile_escapes(CanonEncShared& S, Get get, uint32_t N) {
  long long delta = 0;
  bool dead = false;
  for (uint32_t k = threadIdx.x; k < N; k += kThreads) {
    const int32_t v = get(k);
    if (!canon_value_hits_reference_bug(v)) continue;
    int e2, e8;
    const int counted = canon_classify(v, &e2, &e8), written = (v >> 24) + 128;
    if (S.len[written] == 0) dead = true;
    delta += (int(S.len[written]) + 3 * (int(S.len[kSymEsc8]) + 8)) - (int(S.len[counted]) + 2 * (int(S.len[kSymEsc8]) + 8));
  }
  if (delta) atomicAdd(&S.totalBits, (unsigned long long)delta);
  return __syncthreads_or(dead ? 1 : 0) == 0;
}

// Writes tables + text + end-of-text for a stream whose code has been built in S.  All threads call.
template <class Get>
__device__ inline void canon_emit_stream(CanonEncShared& S, BitWindow& W, BitOut& o, Get get, uint32_t N) {
  const int tid = threadIdx.x;
  // tables, serially (<= ~6.3 kbit)
  bitwin_reserve(W, o, S.headerBits + 64);
  __syncthreads();
  if (tid == 0) {
    WinSink sink{W.win, o.bitPos - o.gbase * 32u};
    sink.put(0, 1);  // reserved
    for (int i = 0; i < S.nC; i++) {  // LengthEncoder.writeEncodedLengths (:169-195)
      sink.put(S.ccodes[i], 5);
      int x = canon_run_extra_bits(S.ccodes[i]);
      if (x) sink.put(S.cruns[i], x);
    }
    for (int i = 0; i < S.nT; i++) {  // CanonicalHuffman.java:323-342
      int code = S.tcodes[i];
      sink.put(S.ctRcode[code], S.ctLen[code]);
      int x = canon_run_extra_bits(code);
      if (x) sink.put(S.truns[i], x);
    }
  }
  o.bitPos += S.headerBits;
  __syncthreads();
  // text: kIpt consecutive values per thread, <= 84 bits each
  constexpr int kIpt = 4;
  for (uint32_t k0 = 0; k0 < N; k0 += kThreads * kIpt) {
    int32_t v[kIpt];
    uint32_t myBits = 0;
    int nv = 0;
#pragma unroll
    for (int j = 0; j < kIpt; j++) {
      uint32_t k = k0 + tid * kIpt + j;
      if (k < N) {
        v[j] = get(k);
        int e2, e8;
        int sym = canon_classify_emit(v[j], &e2, &e8);
        myBits += S.len[sym] + e2 * (S.len[kSymEsc2] + 2) + e8 * (S.len[kSymEsc8] + 8);
        nv = j + 1;
      }
    }
    uint32_t chunkBits;
    uint32_t ex = block_exclusive_scan(myBits, S.scan, &chunkBits);
    bitwin_reserve(W, o, chunkBits);  // 1024 values * 84 bits always fit an empty window
    ThreadBits tb;
    tb.begin(W, o, o.bitPos + ex);
#pragma unroll
    for (int j = 0; j < kIpt; j++) {
      if (j < nv) {
        int32_t s = v[j];
        int e2, e8;
        int sym = canon_classify_emit(s, &e2, &e8);
        tb.put(S.rcode[sym], S.len[sym]);
        for (int q = e2 - 1; q >= 0; q--) { tb.put(S.rcode[kSymEsc2], S.len[kSymEsc2]); tb.put((uint32_t(s) >> (2 * q)) & 3u, 2); }
        for (int q = e8 - 1; q >= 0; q--) { tb.put(S.rcode[kSymEsc8], S.len[kSymEsc8]); tb.put((uint32_t(s) >> (8 * q)) & 0xffu, 8); }
      }
    }
    tb.end();
    o.bitPos += chunkBits;
    __syncthreads();
  }
  // end of text
  bitwin_reserve(W, o, 32);
  __syncthreads();
  if (tid == 0) {
    WinSink sink{W.win, o.bitPos - o.gbase * 32u};
    sink.put(S.rcode[kSymEot], S.len[kSymEot]);
  }
  o.bitPos += S.len[kSymEot];
  __syncthreads();
}

}  // namespace g4
