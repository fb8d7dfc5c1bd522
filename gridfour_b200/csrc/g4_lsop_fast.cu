// g4_lsop_fast.cu -- LSOP12 decode, fast path for sm_100a: byte hand-over + TMA-fed wavefront.
//
// Reference (under /root/reference/core/src/main/java/org/gridfour/lsop/): LsDecoder12.java:94-470 (decode,
// unpackInitializers, unpackInterior), LsHeader.java:104-189; canonical Huffman text:
// compress/canonicalHuffman/CanonicalHuffman.java:441-519.
//
// Three kernels over the canonical-Huffman (type 2) LSOP tiles of a band whose rows are 16-byte aligned:
//   H  lsop2_head_kernel   one warp per tile.  LSOP header, both code tables, the initializer stream (about 4R+2C
//      values) decoded into SHARED memory, where all initializer scans run: rows 0 and 1 leave as whole coalesced
//      rows; for every other row a 16-byte side record {v[r][0], v[r][1], D2[r], D1[r]} is written, D2/D1 being the
//      column prefix sums that turn the Triangle predictor of the last two columns (LsDecoder12.java:459-468) into
//      v[r][C-2] = v[r][C-3] + D2[r], v[r][C-1] = v[r][C-2] + D1[r] -- no dependence on the row above is left.
//   T  lsop2_text_kernel   one 512-thread CTA per tile.  The packing is staged into shared memory by ONE bulk-async
//      copy (cp.async.bulk + mbarrier) while the tables and the multi-symbol LUT are built; the interior residuals
//      leave as ONE BYTE each (symbol = residual + 128) into a shared-memory image of the tile's residual scratch,
//      which goes to HBM with one bulk-async store.  Residuals that do not fit a byte (escapes, nulls) are written as
//      byte 0 plus an entry in a small per-tile exception list.
//   W  lsop2_wave_kernel   one warp per tile.  The causal 12-tap float32 stencil as a wavefront (lane = row, one
//      4-column block behind the lane above).  The residual scratch is laid out per LANE (all rows a lane will ever
//      own, back to back) with a lane pitch L = 4 (mod 16): a tensor map whose row stride is L - 4 then hands the
//      warp a 32 x 64-byte box in which every lane finds ITS next sixteen blocks at the same offset -- the 4-byte
//      skew between lanes is absorbed by the descriptor.  One TMA box load per sixteen iterations, double buffered,
//      completion by mbarrier; values leave as whole 32-byte sectors (256-bit stores).
//
// Arithmetic of one cell (bit-identical to the reference for |p| < 2^21, checked per cell):
//   p   = u1*z1 + ... + u12*z12          strictly left to right, float32, no FMA (-fmad=false)
//   t1  = RD(p + 1.5*2^22)               multiples of 1/2: 1.5*2^22 + floor(2p)/2, exactly
//   t2  = RD(t1 + (1.5*2^22 + 0.5))      integers:         1.5*2^23 + floor(p + 1/2)  == StrictMath.round(p)
//   val = residual + (bits(t2) - bits(1.5*2^23))
//   float(val) = t2 + (float(residual) - 1.5*2^23)      exact, |val| < 2^24; the second operand does not depend on p
// which is 12 FMUL + 11 FADD + 3 FADD per cell on the dependent chain (the Java formula needs 23) and one integer add;
// the two round-down adds were verified against Math.round for every float32 in (-2^21, 2^21).
// A tile that leaves that range (or holds too many exceptions) is handed to the general kernels of g4_lsop.cu.
#include <cuda.h>
#include <cstdio>
#include <type_traits>
#include "g4_kernels.h"
#include "g4_predict.cuh"
#include "g4_lsop_common.cuh"

namespace g4 {

namespace {

constexpr int kTextThreads = 512;
constexpr uint32_t kTextSubBits = 320;        // target sub-sequence size of the text kernel: one sub-sequence per thread for a 180x240 tile
constexpr int kExcWords = 128;               // per tile: [0] count, [2+2i] interior index, [3+2i] value
constexpr int kExcCap = (kExcWords - 2) / 2;  // 63 exceptions; more -> general path
// Staged form of the text decode (lsop_text_decode below), on unless built with -DG4_TEXT_STAGED=0.
#ifndef G4_TEXT_STAGED
#define G4_TEXT_STAGED 1
#endif
constexpr bool kTextStaged = G4_TEXT_STAGED != 0;
constexpr int kStageWordsPerThread = 22;  // registers that carry a thread's staged symbol bytes across the barrier
constexpr int kSpillWords = 8;            // words behind every slot in a per-CTA global scratch (stays in L2): the slots hold only
                                          // about 7 % more than the average sub-sequence, so the longer ones spill their tail
constexpr int kResidGuard = 256;             // bytes in front of the first tile of the residual scratch
constexpr float kMagicHalf = 6291456.0f;     // 1.5 * 2^22
constexpr float kMagicHalfUp = 6291456.5f;   // 1.5 * 2^22 + 1/2
constexpr float kMagicInt = 12582912.0f;     // 1.5 * 2^23, bits 0x4B400000
constexpr float kRange = 2097152.0f;         // 2^21

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
}

// stream geometry shared by the three kernels (LsopFastGeom in g4_kernels.h):
//   lane l in [2, nLanes) owns rows l, l + rpg, l + 2 rpg, ...  (group g = row / rpg ... row = g * rpg + l)
//   word q of a lane's stream: q = 0 is a virtual block (columns 0,1 of the lane's first row); q = 1 + g * nB + b holds
//   the four residual bytes of columns 4b+2 .. 4b+5 of row g * rpg + l (b < nB - 1; block nB - 1 is the wrap block: the
//   row's last two columns and the next row's first two, no residual bytes)
__device__ __forceinline__ uint32_t lane_row_offset(const LsopFastGeom& g, int row) {
  const int gi = (row - 2) / g.rpg, l = 2 + (row - 2) - gi * g.rpg;
  return uint32_t(l) * uint32_t(g.laneBytes) + 4u * uint32_t(1 + gi * g.nB);
}

// ---- kernel H -----------------------------------------------------------------------------------------------------
// p[i] <- carry + p[0] + ... + p[i] (mod 2^32), i < n; one warp, shared memory.  Returns the last value.
__device__ inline uint32_t warp_scan_inplace(int32_t* p, int n, uint32_t carry, int stride = 1) {
  const int lane = threadIdx.x & 31;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    const uint32_t x = i < n ? uint32_t(p[i * stride]) : 0u;
    const uint32_t inc = warp_inclusive_scan(x);
    if (i < n) p[i * stride] = int32_t(carry + inc);
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
  __syncwarp();
  return carry;
}

__global__ void __launch_bounds__(kThreads, 3) lsop2_head_kernel(LsopFastArgs A, uint32_t maxPackBytes, int listBegin, int listEnd) {
  extern __shared__ __align__(16) unsigned char headSmem[];
  const DecodeArgs& a = A.a;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int R = A.g.R, C = A.g.C;
  const uint32_t nInit = uint32_t(4 * R + 2 * C - 9);
  const size_t perWarp = ((sizeof(CanonWarpShared) + 15) & ~size_t(15)) + ((size_t(nInit) * 4 + 15) & ~size_t(15));
  CanonWarpShared& W = *reinterpret_cast<CanonWarpShared*>(headSmem + perWarp * warp);
  int32_t* init = reinterpret_cast<int32_t*>(headSmem + perWarp * warp + ((sizeof(CanonWarpShared) + 15) & ~size_t(15)));
  const int li = listBegin + blockIdx.x * kWarps + warp;
  if (li >= listEnd || li >= *a.listCount) return;
  const int tIdx = a.list[li];
  uint8_t* m = A.meta + size_t(tIdx) * kLsopMetaBytes;
  if (lane == 0) {
    *reinterpret_cast<uint32_t*>(m) = 0;
    A.exc[size_t(tIdx) * kExcWords] = 0;
  }
  const TileView t = tile_view(a.band, a.grid, tIdx);
  const uint8_t* packing = a.arena + a.offsets[tIdx];
  const uint32_t len = a.lens[tIdx];
  LsHeaderInfo h = parse_ls_header(packing, len, A.coef + size_t(tIdx) * 12, lane == 0);
  if (lane == 0 && a.lsopCks && h.ok && h.hasChecksum) { a.lsopCks[2 * tIdx] = 1u; a.lsopCks[2 * tIdx + 1] = h.valueChecksum; }
  if (!h.ok) {
    if (lane == 0) a.status[tIdx] = G4_ERR_FORMAT;
    return;
  }
  if (h.type != 2 || len + 96u > maxPackBytes) {  // legacy Huffman / Deflate bodies, oversized packings: general kernels
    if (lane == 0) A.defer[atomicAdd(A.deferCount, 1)] = tIdx;
    return;
  }
  BitSrc src;
  src.init(packing, len);  // absolute bit positions inside the packing
  canon_warp_parse_header(W, src, h.headerSize * 8u);
  int status = W.error ? G4_ERR_FORMAT : G4_OK;
  uint32_t endBit = 0;
  if (status == G4_OK) {
    uint32_t eb = 0, nv = 0;
    auto emit = [&](uint32_t k, int32_t v) { init[k] = v; };
    const bool ok = canon_warp_decode_text(W, src, W.textStart, nInit, emit, &eb, &nv) && nv == nInit;
    __syncwarp();
    if (lane == 0) {
      if (!ok) W.error = 1;
      W.textStart = eb;
    }
    __syncwarp();
    if (W.error) status = G4_ERR_FORMAT;
    endBit = W.textStart;
  }
  if (status == G4_OK) {
    canon_warp_parse_header(W, src, endBit);  // interior stream: tables only, exported for kernel T
    if (W.error) status = G4_ERR_FORMAT;
    else {
      for (int i = lane; i < kCanonSymbols; i += 32) m[8 + i] = W.lens[i];
      if (lane == 0) *reinterpret_cast<uint32_t*>(m) = W.textStart;
    }
  }
  if (status == G4_OK) {
    // LsDecoder12.unpackInitializers (:204-241) on the stream-ordered values in shared memory
    int32_t* sA = init;                      // row 0, columns 1..C-1
    int32_t* sB = sA + (C - 1);              // column 0, rows 1..R-1
    int32_t* sC = sB + (R - 1);              // row 1, columns 1..C-1
    int32_t* sD = sC + (C - 1);              // column 1, rows 2..R-1
    int32_t* sE = sD + (R - 2);              // (r, C-2), (r, C-1) for r = 2..R-1
    const uint32_t seed = uint32_t(h.seed);
    warp_scan_inplace(sA, C - 1, seed);      // v[0][c]
    warp_scan_inplace(sB, R - 1, seed);      // v[r][0]
    // row 1: T[c] = v[1][c] - v[0][c], T[c] = T[c-1] + residual(1,c)
    warp_scan_inplace(sC, C - 1, uint32_t(sB[0]) - seed);
    for (int i = lane; i < C - 1; i += 32) sC[i] = int32_t(uint32_t(sC[i]) + uint32_t(sA[i]));  // v[1][c]
    __syncwarp();
    // column 1: U[r] = v[r][1] - v[r][0], U[r] = U[r-1] + residual(r,1)
    warp_scan_inplace(sD, R - 2, uint32_t(sC[0]) - uint32_t(sB[0]));
    // last two columns: D2[r] = v[r][C-2] - v[r][C-3] = D2[r-1] + residual(r,C-2); D1 the same one column further
    warp_scan_inplace(sE, R - 2, uint32_t(sC[C - 3]) - uint32_t(sC[C - 4]), 2);
    warp_scan_inplace(sE + 1, R - 2, uint32_t(sC[C - 2]) - uint32_t(sC[C - 3]), 2);
    int32_t* row0 = t.row(0);
    int32_t* row1 = t.row(1);
    for (int c = lane; c < C; c += 32) {
      row0[c] = c ? sA[c - 1] : int32_t(seed);
      row1[c] = c ? sC[c - 1] : sB[0];
    }
    int4* side = A.side + size_t(tIdx) * R;
    for (int r = lane; r < R; r += 32) {
      int4 s;
      if (r == 0) s = make_int4(int32_t(seed), sA[0], 0, 0);
      else if (r == 1) s = make_int4(sB[0], sC[0], 0, 0);
      else s = make_int4(sB[r - 1], int32_t(uint32_t(sB[r - 1]) + uint32_t(sD[r - 2])), sE[2 * (r - 2)], sE[2 * (r - 2) + 1]);
      side[r] = s;
    }
  }
  if (lane == 0) a.status[tIdx] = status;
}

// ---- kernel T -----------------------------------------------------------------------------------------------------
// Sink of the text decoder's packed write pass: one byte per value into the shared-memory image of the tile's residual
// scratch.  A run starts at any byte; its head goes out byte by byte up to the next word, the rest as aligned words
// (every row of the image is a whole number of words).
struct ByteTileSink {
  static constexpr bool kPacked = true;
  uint8_t* tile;       // shared memory, tileBytes
  uint32_t* exc;       // this tile's exception list (global)
  int w;               // C - 4, bytes per row
  int L, nB4, nLanes, rpg;  // lane pitch, 4 * nB, lanes in use, rows per group (LsopFastGeom)
  uint32_t addr;       // word-aligned image offset of queue byte 0
  int left;            // bytes from addr to the end of the row
  uint32_t rowOff;     // image offset of the current row
  int lane, rowBase;   // the current row = rowBase + lane, lane in [2, nLanes)
  int head;            // dummy bytes at the front of the queue (run head inside a word), 0 after the first store
  int cnt;             // queued bytes, dummies included; < 4 between calls
  uint64_t q;
  int excK, excSlot;   // the last exception this thread recorded
  int32_t excV;

  __device__ __forceinline__ void init(uint8_t* img, uint32_t* e, const LsopFastGeom& g) {
    tile = img;
    exc = e;
    w = g.C - 4;
    L = g.laneBytes;
    nB4 = 4 * g.nB;
    nLanes = g.nLanes;
    rpg = g.rpg;
  }
  __device__ __forceinline__ void begin(uint32_t k0) {
    const int rr = int(k0) / w, cc = int(k0) - rr * w;
    const int gi = rr / rpg;
    lane = 2 + rr - gi * rpg;
    rowBase = gi * rpg;
    rowOff = uint32_t(lane) * uint32_t(L) + 4u + uint32_t(gi) * uint32_t(nB4);
    const uint32_t a0 = rowOff + uint32_t(cc);
    head = int(a0 & 3u);
    addr = a0 & ~3u;
    left = w - (cc & ~3);
    cnt = head;
    q = 0;
    excK = -1;
    excSlot = 0;
    excV = 0;
  }
  __device__ __forceinline__ int cur_k() const {  // interior index of the value that will be queued next
    return (rowBase + lane - 2) * w + int(addr - rowOff) + cnt;
  }
  __device__ __forceinline__ void next_row() {
    rowOff += uint32_t(L);
    if (++lane == nLanes) {  // the next group's first row: back to lane 2, one row period further
      lane = 2;
      rowBase += rpg;
      rowOff += uint32_t(nB4) - uint32_t(rpg) * uint32_t(L);
    }
    addr = rowOff;
    left = w;
  }
  __device__ __forceinline__ void store_word() {  // cnt >= 4
    if (head) {
      for (int i = head; i < 4; i++) tile[addr + i] = uint8_t(q >> (8 * i));
      head = 0;
    } else *reinterpret_cast<uint32_t*>(tile + addr) = uint32_t(q);
    q >>= 32;
    cnt -= 4;
    addr += 4;
    left -= 4;
    if (__builtin_expect(left == 0, 0)) next_row();
  }
  __device__ __forceinline__ void push(uint32_t bytes, int n) {  // n = 1..3 symbol bytes, first value in the low byte
    q |= uint64_t(bytes) << (8 * cnt);
    cnt += n;
    if (cnt >= 4) store_word();
  }
  __device__ __forceinline__ int add_exception(int k, int32_t v) {
    const uint32_t slot = atomicAdd(exc, 1u);
    if (slot < uint32_t(kExcCap)) {
      exc[2 + 2 * slot] = uint32_t(k);
      exc[3 + 2 * slot] = uint32_t(v);
    }
    excK = k;
    excV = v;
    excSlot = int(slot);
    return int(slot);
  }
  __device__ __forceinline__ void put_rare(int32_t v) {  // a value that is no byte (null)
    add_exception(cur_k(), v);
    push(0u, 1);
  }
  // an escape extends the value before it (CanonicalHuffman.java:495-504): v = (v << nb) | bits
  __device__ __forceinline__ void amend(int nb, uint32_t bits) {
    const int kp = cur_k() - 1;
    if (kp == excK) {
      excV = int32_t((uint32_t(excV) << nb) | bits);
      if (excSlot < kExcCap) exc[3 + 2 * excSlot] = uint32_t(excV);
      return;
    }
    uint32_t b;
    if (cnt > head) {  // still queued
      const int sh = 8 * (cnt - 1);
      b = uint32_t(q >> sh) & 0xffu;
      q &= ~(0xffull << sh);
    } else {  // already in the image: the byte in front of addr + cnt, or the last byte of the previous row
      // (cnt == head: with head != 0 nothing of this run was stored yet, which `have` in the caller excludes)
      uint32_t p = addr + uint32_t(cnt);
      if (p > rowOff) p -= 1;
      else {
        uint32_t prevOff = rowOff - uint32_t(L);
        if (lane == 2) prevOff = rowOff - uint32_t(nB4) + uint32_t(rpg - 1) * uint32_t(L);
        p = prevOff + uint32_t(w) - 1;
      }
      b = tile[p];
      tile[p] = 0;
    }
    add_exception(kp, int32_t(((b - 128u) << nb) | bits));
  }
  __device__ __forceinline__ void end() {
    for (int i = head; i < cnt; i++) tile[addr + i] = uint8_t(q >> (8 * i));
  }
};

// ---- staged text decode -------------------------------------------------------------------------------------------
// canon_fast_decode_text decodes every bit at least twice: once to COUNT the values of each sub-sequence (their output
// positions are a prefix sum of the counts) and once to write them.  Here the counting pass also keeps the symbol
// bytes it sees: every sub-sequence has a small slot in the tile image itself (the image is not written before the
// prefix sum is known, so it is free), and after the prefix sum each thread takes its slot into registers, the CTA
// synchronises, and the bytes go to their place in the image -- a copy instead of a second decode.  A sub-sequence that
// outgrows its slot (the slots hold about 10 % more than the average) is decoded a second time like before; one that
// holds a value that is no byte (escape, null) is walked once more by text_exceptions_sub for the exception list.
constexpr int kSubRare = 4;      // S.eot[] flag: the sub-sequence holds an escape or a null
constexpr int kSubOverflow = 8;  // S.eot[] flag: more values than the slot holds

// Like canon_fast_count; additionally stores the symbol byte of every counted value to slot[0 .. slotWords) (shared memory).
__device__ __forceinline__ void text_stage_sub(const CanonFastShared& S, uint32_t nBits, uint32_t start, uint32_t limit, uint32_t* slot,
                                               uint32_t slotWords, uint32_t* spill, uint32_t* endOut, uint32_t* cntOut, int* flagOut) {
  BitCursor cur;
  cur.init(S, start, limit);
  uint32_t c = 0, end;
  int flag = 0;
  uint64_t q = 0;
  uint32_t qc = 0, w = 0;
  for (;;) {
    if (cur.rem >= kFastLutBits) {
      const uint32_t m = S.mlut[cur.peek() & ((1u << kFastLutBits) - 1u)];
      const uint32_t n = m >> 28;
      if (n) {
        cur.skip(S, (m >> 24) & 15u);
        c += n;
        q |= uint64_t(m & 0xffffffu) << (8 * qc);
        qc += n;
        if (qc >= 4) {
          if (w < slotWords) slot[w] = uint32_t(q);
          else if (w - slotWords < uint32_t(kSpillWords)) spill[w - slotWords] = uint32_t(q);
          w++;
          q >>= 32;
          qc -= 4;
        }
        continue;
      }
    }
    const uint32_t e = S.lut[cur.peek() & ((1u << kFastLutBits) - 1u)];
    uint32_t byte;
    if (e - 1u < 0x7fffu) {  // LUT hit on a plain value: e = sym | len << 9
      if (cur.rem <= 0) { end = cur.pos(limit); break; }
      cur.skip(S, e >> 9);
      byte = e & 0xffu;
    } else {
      const uint32_t p0 = cur.pos(limit);
      uint32_t after;
      const int sym = canon_fast_rare_symbol(S, e, p0, nBits, &after);
      if (sym < 0) { flag = 2; end = p0; break; }
      if (sym == kSymEsc2 || sym == kSymEsc8) {
        after += sym == kSymEsc2 ? 2u : 8u;
        if (after > nBits) { flag = 2; end = p0; break; }
        flag |= kSubRare;
        cur.init(S, after, limit);
        continue;
      }
      if (p0 >= limit) { end = p0; break; }
      if (sym == kSymEot) { flag |= 1; end = after; break; }
      if (sym == kSymNull) { flag |= kSubRare; byte = 0u; }
      else byte = uint32_t(sym);  // a byte value with a code longer than the LUT
      cur.init(S, after, limit);
    }
    c++;
    q |= uint64_t(byte) << (8 * qc);
    qc += 1;
    if (qc >= 4) {
      if (w < slotWords) slot[w] = uint32_t(q);
      else if (w - slotWords < uint32_t(kSpillWords)) spill[w - slotWords] = uint32_t(q);
      w++;
      q >>= 32;
      qc -= 4;
    }
  }
  if (qc) {
    if (w < slotWords) slot[w] = uint32_t(q);
    else if (w - slotWords < uint32_t(kSpillWords)) spill[w - slotWords] = uint32_t(q);
    w++;
  }
  if (w > slotWords + uint32_t(kSpillWords)) flag |= kSubOverflow;
  *endOut = end;
  *cntOut = c;
  *flagOut = flag;
}

// The write pass of canon_fast_decode_text for ONE sub-sequence (values from `start` up to the first value at or after
// `limit`), for the sub-sequences that did not fit their slot.  Returns false for a malformed stream.
__device__ __forceinline__ bool text_write_sub(const CanonFastShared& S, uint32_t nBits, uint32_t start, uint32_t limit, uint32_t k0,
                                            ByteTileSink& sink) {
  BitCursor cur;
  cur.init(S, start, limit);
  sink.begin(k0);
  bool have = false, ok = true;
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
    if (sym < 0) { ok = false; break; }
    if (sym == kSymEsc2 || sym == kSymEsc8) {
      const int nb = sym == kSymEsc2 ? 2 : 8;
      if (!have || after + nb > nBits) { ok = false; break; }
      SmemBitSrc src{S.sw, nBits};
      sink.amend(nb, src.bits(after, nb));
      after += nb;
    } else {
      if (p0 >= limit || sym == kSymEot) break;
      if (sym == kSymNull) sink.put_rare(INT32_MIN);
      else sink.push(uint32_t(sym), 1);  // a byte value with a code longer than the LUT
      have = true;
    }
    cur.init(S, after, limit);
  }
  sink.end();
  return ok;
}

// Exception-list entries of one sub-sequence whose values start at interior index k0 (the bytes are in the image).
__device__ __noinline__ void text_exceptions_sub(const CanonFastShared& S, uint32_t nBits, uint32_t start, uint32_t limit, uint32_t k0,
                                                 uint8_t* tile, uint32_t* exc, int w, int L, int nB4, int rpg) {
  auto image_offset = [&](uint32_t k) {
    const int rr = int(k) / w, cc = int(k) - rr * w, gi = rr / rpg, l = 2 + rr - gi * rpg;
    return uint32_t(l) * uint32_t(L) + 4u + uint32_t(gi) * uint32_t(nB4) + uint32_t(cc);
  };
  auto add = [&](uint32_t k, int32_t v) {
    const uint32_t slot = atomicAdd(exc, 1u);
    if (slot < uint32_t(kExcCap)) {
      exc[2 + 2 * slot] = k;
      exc[3 + 2 * slot] = uint32_t(v);
    }
  };
  SmemBitSrc src{S.sw, nBits};
  uint32_t pos = start, k = k0, pk = 0;
  bool pending = false;
  int32_t pv = 0;
  for (;;) {
    const uint32_t e = S.lut[src.peek32(pos) & ((1u << kFastLutBits) - 1u)];
    int sym;
    uint32_t after;
    if (e - 1u < 0x7fffu) { sym = int(e & 0xffu); after = pos + (e >> 9); }
    else {
      sym = canon_fast_rare_symbol(S, e, pos, nBits, &after);
      if (sym < 0) break;
    }
    if (sym == kSymEsc2 || sym == kSymEsc8) {  // extends the value before it (CanonicalHuffman.java:495-504): v = (v << nb) | bits
      const int nb = sym == kSymEsc2 ? 2 : 8;
      if (after + nb > nBits || k == k0) break;
      if (!pending) {
        pk = k - 1;
        const uint32_t o = image_offset(pk);
        pv = int32_t(tile[o]) - 128;
        tile[o] = 0;
        pending = true;
      }
      pv = int32_t((uint32_t(pv) << nb) | src.bits(after, nb));
      after += nb;
    } else {
      if (pending) { add(pk, pv); pending = false; }
      if (pos >= limit || sym == kSymEot) break;
      if (sym == kSymNull) add(k, INT32_MIN);
      k++;
    }
    pos = after;
  }
  if (pending) add(pk, pv);
}

// Decodes the interior text (tables and LUT ready, text at bit T0) into the tile image.  All kTextThreads threads call.
// Returns 0 = done, 1 = malformed stream, 2 = the tile does not suit the staged form (caller uses canon_fast_decode_text).
__device__ int lsop_text_decode(CanonFastShared& S, uint32_t nBits, const uint32_t T0, uint32_t nInterior, uint32_t imageBytes,
                                ByteTileSink sink, uint32_t* spillArea, uint32_t lookback) {
  constexpr int NT = kTextThreads;
  constexpr int kRounds = kFastMaxSub / NT;
  const int tid = threadIdx.x;
  const uint32_t avail = nBits - T0;
  uint32_t rounds = (avail / kTextSubBits + NT - 1) / NT;
  if (rounds < 1u) rounds = 1u;
  if (rounds > uint32_t(kRounds)) rounds = kRounds;
  uint32_t B = (avail + rounds * NT - 1) / (rounds * NT);
  if (B < 96u) B = 96u;
  const int nSub = int((avail + B - 1) / B);
  // slots: the image cut into nSub equal pieces, capped by what a thread can carry in registers
  uint32_t slotWords = (imageBytes / uint32_t(nSub)) >> 2;
  if (slotWords > uint32_t(kStageWordsPerThread) / rounds) slotWords = uint32_t(kStageWordsPerThread) / rounds;
  // not worth it unless slot + spill hold the average sub-sequence with room to spare (else most of them decode twice)
  if (uint64_t(slotWords + uint32_t(kSpillWords)) * 4u * uint64_t(nSub) * 3u < uint64_t(nInterior) * 4u) return 2;
  uint32_t* const image32 = reinterpret_cast<uint32_t*>(sink.tile);
  if (tid == 0) S.firstEot = nSub;
  // pass 0: only the END of every sub-sequence matters here, so start kFastLookback bits before the limit and rely on
  // self-synchronisation (sub-sequence 0 as well: every thread does the same amount)
#pragma unroll 1
  for (int i = tid; i < nSub; i += NT) {
    uint32_t limit = T0 + uint32_t(i + 1) * B;
    if (limit > nBits) limit = nBits;
    uint32_t from = T0 + uint32_t(i) * B;
    if (limit - from > lookback) from = limit - lookback;
    uint32_t e, c;
    int f;
    canon_fast_count(S, nBits, from, limit, &e, &c, &f);
    S.endpos[i] = e;
    S.startv[i] = 0xffffffffu;  // forces the exact decode of every sub-sequence in the first pass below
  }
  // synchronisation passes: sub-sequence i must start where i-1 ended (see canon_fast_decode_text)
  volatile uint32_t* vend = S.endpos;
  __syncthreads();
  for (int pass = 0; pass <= nSub; pass++) {
    bool any = false;
#pragma unroll 1
    for (int i = tid; i < nSub; i += NT) {
      const uint32_t ns = i ? vend[i - 1] : T0;
      if (ns != S.startv[i]) {
        S.startv[i] = ns;
        uint32_t limit = T0 + uint32_t(i + 1) * B;
        if (limit > nBits) limit = nBits;
        uint32_t e, c;
        int f;
        text_stage_sub(S, nBits, ns, limit, image32 + size_t(i) * slotWords, slotWords, spillArea + size_t(i) * kSpillWords, &e, &c, &f);
        vend[i] = e;
        S.cnt[i] = uint16_t(c);
        S.eot[i] = uint8_t(f);
        any = true;
      }
    }
    if (!__syncthreads_or(any ? 1 : 0)) break;
  }
#pragma unroll 1
  for (int i = tid; i < nSub; i += NT)
    if (S.eot[i] & 3) atomicMin(&S.firstEot, i);
  __syncthreads();
  const int fe = S.firstEot;
  if (fe == nSub || (S.eot[fe] & 3) == 2) return 1;  // no end-of-text, an invalid code, or the data ended early
  // value offsets: thread tid owns sub-sequences tid*kRounds .. +kRounds-1 for the scan
  uint32_t mySum = 0;
#pragma unroll
  for (int j = 0; j < kRounds; j++) {
    const int i = tid * kRounds + j;
    mySum += (i <= fe && i < nSub) ? S.cnt[i] : 0u;
  }
  uint32_t total;
  const uint32_t ex = block_exclusive_scan<NT>(mySum, S.scan, &total);
  if (total != nInterior) return 1;
  uint32_t* offv = S.endpos;  // end positions are no longer needed: the array now holds the first value index of every sub-sequence
  {
    uint32_t run = ex;
#pragma unroll
    for (int j = 0; j < kRounds; j++) {
      const int i = tid * kRounds + j;
      if (i < nSub) {
        offv[i] = run;
        run += (i <= fe) ? S.cnt[i] : 0u;
      }
    }
  }
  // slots -> registers (every thread: sub-sequences tid and tid + NT), barrier, registers -> their place in the image
  uint32_t r[kStageWordsPerThread];
#pragma unroll
  for (int j = 0; j < kStageWordsPerThread; j++) r[j] = 0;
  {
    const uint32_t perSub = uint32_t(kStageWordsPerThread) / rounds;
    for (uint32_t s = 0; s < rounds; s++) {
      const int i = tid + int(s) * NT;
      if (i <= fe && !(S.eot[i] & kSubOverflow)) {
        const uint32_t* slot = image32 + size_t(i) * slotWords;
#pragma unroll
        for (int j = 0; j < kStageWordsPerThread; j++)
          if (uint32_t(j) >= s * perSub && uint32_t(j) < s * perSub + slotWords) r[j] = slot[uint32_t(j) - s * perSub];
      }
    }
  }
  __syncthreads();
  bool bad = false;
  {
    const uint32_t perSub = uint32_t(kStageWordsPerThread) / rounds;
    for (uint32_t s = 0; s < rounds; s++) {
      const int i = tid + int(s) * NT;
      if (i > fe) continue;
      const uint32_t n = S.cnt[i];
      const int flags = S.eot[i];
      uint32_t limit = T0 + uint32_t(i + 1) * B;
      if (limit > nBits) limit = nBits;
      if (flags & kSubOverflow) {
        if (!text_write_sub(S, nBits, S.startv[i], limit, offv[i], sink)) bad = true;
        continue;
      }
      if (n == 0) continue;
      sink.begin(offv[i]);
#pragma unroll
      for (int j = 0; j < kStageWordsPerThread; j++) {
        const uint32_t at = (uint32_t(j) - s * perSub) * 4u;  // byte position of word j inside this sub-sequence's slot
        if (uint32_t(j) >= s * perSub && uint32_t(j) < s * perSub + slotWords && at < n) {
          const uint32_t m = n - at;
          sink.push(m >= 4u ? r[j] : (r[j] & ((1u << (8 * m)) - 1u)), m >= 4u ? 4 : int(m));
        }
      }
      if (n > slotWords * 4u) {  // the tail this thread spilled in the staging pass
        const uint32_t* sp = spillArea + size_t(i) * kSpillWords;
        for (uint32_t at = slotWords * 4u; at < n; at += 4u) {
          const uint32_t wv = *sp++, m = n - at;
          sink.push(m >= 4u ? wv : (wv & ((1u << (8 * m)) - 1u)), m >= 4u ? 4 : int(m));
        }
      }
      sink.end();
      if (flags & kSubRare) text_exceptions_sub(S, nBits, S.startv[i], limit, offv[i], sink.tile, sink.exc, sink.w, sink.L, sink.nB4, sink.rpg);
    }
  }
  return __syncthreads_or(bad ? 1 : 0) ? 1 : 0;
}

__global__ void __launch_bounds__(kTextThreads, 2) lsop2_text_kernel(LsopFastArgs A, uint32_t stageWords, int listBegin, int listEnd) {
  extern __shared__ __align__(128) unsigned char textSmem[];
  CanonFastShared& F = *reinterpret_cast<CanonFastShared*>(textSmem);
  uint8_t* tileImg = textSmem + ((canon_fast_smem_bytes(stageWords) + 127) & ~size_t(127));
  __shared__ int sTile;
  __shared__ __align__(8) uint64_t sBar;
  const DecodeArgs& a = A.a;
  const int tid = threadIdx.x;
  const uint32_t bar = smem_u32(&sBar);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  uint32_t parity = 0;
  for (;;) {
    if (tid == 0) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous tile's image has left shared memory
      sTile = listBegin + atomicAdd(a.counter, 1);
    }
    __syncthreads();
    const int li = sTile;
    if (li >= listEnd || li >= *a.listCount) break;
    const int tIdx = a.list[li];
    const uint8_t* m = A.meta + size_t(tIdx) * kLsopMetaBytes;
    const uint32_t T0 = *reinterpret_cast<const uint32_t*>(m);
    if (T0 == 0) { __syncthreads(); continue; }  // deferred or rejected by kernel H
    const uint8_t* packing = a.arena + a.offsets[tIdx];
    const uint32_t len = a.lens[tIdx];
    // stage: the packing from its 16-byte line start; whole 16-byte pieces by ONE bulk-async copy, the tail by bytes
    const uint32_t delta = uint32_t(reinterpret_cast<uintptr_t>(packing) & 15u);
    const uint8_t* src16 = packing - delta;
    const uint32_t span = len + delta, nBulk = span & ~15u;
    if (tid == 0 && nBulk) {
      mbar_expect_tx(bar, nBulk);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(F.sw)), "l"(src16),
                   "r"(nBulk), "r"(bar)
                   : "memory");
    }
    if (tid < 64) {  // tail bytes and zero padding (the decoder reads whole words, up to 8 words past the data)
      const uint32_t i = nBulk + uint32_t(tid);
      reinterpret_cast<uint8_t*>(F.sw)[i] = i < span ? src16[i] : uint8_t(0);
    }
    for (int i = tid; i < kCanonSymbols; i += kTextThreads) F.lens[i] = m[8 + i];
    if (tid == 0) F.error = 0;
    __syncthreads();
    canon_fast_tables_cta<kTextThreads>(F);
    bool ok = F.error == 0;
    if (ok) canon_fast_build_lut<kTextThreads>(F);
    if (nBulk) mbar_wait(bar, parity);
    parity ^= nBulk ? 1u : 0u;
    __syncthreads();
    int rc = ok ? 0 : 1;
    if (ok) {
      const uint32_t nInterior = uint32_t(A.g.R - 2) * uint32_t(A.g.C - 4);
      ByteTileSink sink;
      sink.init(tileImg, A.exc + size_t(tIdx) * kExcWords, A.g);
      if (kTextStaged)
        rc = lsop_text_decode(F, span * 8u, T0 + 8u * delta, nInterior, uint32_t(A.g.tileBytes), sink,
                              reinterpret_cast<uint32_t*>(A.textStage) + size_t(blockIdx.x) * (kFastMaxSub * kSpillWords), A.textLookback);
      else rc = 2;
      if (rc == 2) {  // (uniform) the two-pass form
        uint32_t endBit = 0, nv = 0;
        rc = (canon_fast_decode_text<ByteTileSink, kTextThreads, kTextSubBits>(F, span * 8u, T0 + 8u * delta, nInterior, 0u, sink, &endBit, &nv) &&
              nv == nInterior) ? 0 : 1;
      }
      ok = rc == 0;
    }
    __syncthreads();
    if (tid == 0) {
      if (rc == 1) a.status[tIdx] = G4_ERR_FORMAT;
      else if (A.exc[size_t(tIdx) * kExcWords] > uint32_t(kExcCap)) {  // general kernels
        A.exc[size_t(tIdx) * kExcWords] = uint32_t(kExcCap) + 1u;                  // (the wavefront kernel skips the tile)
        A.defer[atomicAdd(A.deferCount, 1)] = tIdx;
      }
      else {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        uint8_t* dst = A.resid + kResidGuard + size_t(tIdx) * size_t(A.g.tilePitch);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(tileImg)), "r"(uint32_t(A.g.tileBytes))
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- kernel W -----------------------------------------------------------------------------------------------------
#ifndef G4_WAVE_CTAS
#define G4_WAVE_CTAS 2
#endif
constexpr int kWaveStageBytes = 2048;  // one TMA box: 32 lanes x 64 bytes = sixteen iterations

// The value of an exceptional cell (byte 0 in the scratch): its list entry, or -128 when the byte was genuine.
__device__ __noinline__ int32_t wave_exception(const uint32_t* exc, int n, int k) {
  for (int i = 0; i < n; i++)
    if (int(exc[2 + 2 * i]) == k) return int32_t(exc[3 + 2 * i]);
  return -128;
}

template <bool WIDE>
__global__ void __launch_bounds__(kThreads, G4_WAVE_CTAS)
    lsop2_wave_kernel(const __grid_constant__ CUtensorMap tmap, LsopFastArgs A, int listBegin, int listEnd) {
  extern __shared__ __align__(1024) unsigned char waveSmem[];
  const DecodeArgs& a = A.a;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int li = listBegin + blockIdx.x * kWarps + warp;
  if (li >= listEnd || li >= *a.listCount) return;
  const int tIdx = a.list[li];
  if (a.status[tIdx] != G4_OK) return;
  if (*reinterpret_cast<const uint32_t*>(A.meta + size_t(tIdx) * kLsopMetaBytes) == 0u) return;  // general path
  const uint32_t* exc = A.exc + size_t(tIdx) * kExcWords;
  const int nExc = int(exc[0]);
  if (nExc > kExcCap) return;  // deferred by kernel T
  const LsopFastGeom& G = A.g;
  const int R = G.R, C = G.C, nB = G.nB, nLanes = G.nLanes, rpg = G.rpg;
  const TileView t = tile_view(a.band, a.grid, tIdx);
  unsigned char* ring = waveSmem + size_t(warp) * (2 * kWaveStageBytes);
  float4* rowbuf = reinterpret_cast<float4*>(waveSmem + size_t(kWarps) * (2 * kWaveStageBytes)) + size_t(warp) * (2 * nB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(waveSmem + size_t(kWarps) * (2 * kWaveStageBytes) + size_t(kWarps) * (2 * nB) * sizeof(float4)) + 2 * warp;
  const uint32_t bar0 = smem_u32(bars), ring0 = smem_u32(ring);
  const bool feeder = lane < 2;
  const bool computing = lane >= 2 && lane < nLanes;
  const bool writer = lane >= nLanes - 2 && lane < nLanes;  // its rows become the next group's feeder rows
  if (lane == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncwarp();
  auto issue = [&](int chunk) {  // lane 0: box of iterations 16 chunk .. 16 chunk + 15 into stage chunk & 1
    const uint32_t bar = bar0 + 8u * (chunk & 1), dst = ring0 + uint32_t(kWaveStageBytes) * (chunk & 1);
    mbar_expect_tx(bar, kWaveStageBytes);
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(64 * chunk), "r"(0), "r"(tIdx), "r"(bar)
                 : "memory");
  };
  if (lane == 0) issue(0);

  float u[12];
#pragma unroll
  for (int i = 0; i < 12; i++) u[i] = computing ? A.coef[size_t(tIdx) * 12 + i] : 0.f;
  const float bias = feeder ? kMagicInt : 2.f * kMagicInt + 128.f;  // x - bias == float(residual) - 1.5*2^23 (feeder: value - 1.5*2^23)
  const int4* side = A.side + size_t(tIdx) * R;
  uint32_t badBits = 0;  // nonzero: some cell left the range of the fast arithmetic
  const int32_t lim = int32_t(kRange);
  // feeder stream of group 0: rows 0 and 1 (written by kernel H), block b = columns 4b+2 .. 4b+5; the wrap block holds
  // the rows' last two columns and columns 0,1 of the feeder's next row (rows rpg, rpg+1)
  for (int idx = lane; idx < 2 * nB; idx += 32) {
    const int f = idx >= nB ? 1 : 0, b = idx - f * nB;
    const int32_t* rowf = t.row(f);
    int32_t v0, v1, v2, v3;
    if (b < nB - 1) {
      const int2 lo = *reinterpret_cast<const int2*>(rowf + 4 * b + 2), hi = *reinterpret_cast<const int2*>(rowf + 4 * b + 4);
      v0 = lo.x; v1 = lo.y; v2 = hi.x; v3 = hi.y;
    } else {
      const int2 lo = *reinterpret_cast<const int2*>(rowf + C - 2);
      v0 = lo.x; v1 = lo.y;
      const int4 s = rpg + f < R ? side[rpg + f] : make_int4(0, 0, 0, 0);
      v2 = s.x; v3 = s.y;
    }
    if (v0 <= -lim || v0 >= lim || v1 <= -lim || v1 >= lim || v2 <= -lim || v2 >= lim || v3 <= -lim || v3 >= lim) badBits = 1u;
    rowbuf[idx] = make_float4(float(v0), float(v1), float(v2), float(v3));
  }
  // columns 0,1 of every lane's first row (the virtual block at stream word 0)
  int32_t curD2 = 0, curD1 = 0;
  int4 sideNext = make_int4(0, 0, 0, 0);
  if (lane < nLanes && lane < R) sideNext = side[lane];
  __syncwarp();

  float av[8], bv[8];  // rows r-1 / r-2, columns c-2 .. c+5 relative to the block's first column
#pragma unroll
  for (int i = 0; i < 8; i++) { av[i] = 0.f; bv[i] = 0.f; }
  float f1 = 0.f, f2 = 0.f;        // own row, columns c-1 and c-2
  int32_t pi2 = 0, pi3 = 0;        // own outputs 2,3 of the previous block
  int4 sqPrev = make_int4(0, 0, 0, 0);
  int32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;      // wrap-block values (set in the wrap block, read only there)
  float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
  const uint32_t wrapB16 = uint32_t(nB - 1) * 16u;
  uint32_t b16 = lane < nLanes ? uint32_t(nB - 1 - lane) * 16u : 0u;  // 16 x block inside the row; every lane starts towards its virtual wrap block
  int row = lane - rpg;                       // row of the current stream position (virtual row before the first)
  uint32_t validM = 0u;                       // 0xFF800000 while `row` is a tile row this lane computes
  uint32_t started = 0u;                      // past the virtual block in front of the lane's first row
  char* rowp = reinterpret_cast<char*>(t.base + int64_t(row) * t.pitch);  // dereferenced only while valid
  const int64_t groupStep = int64_t(rpg) * t.pitch * 4;
  const int wInterior = C - 4;
  const uint32_t ringLane = ring0 + 64u * uint32_t(lane);
  const uint32_t swz16 = uint32_t((lane >> 1) & 3) * 16u;
  const uint32_t rbRead = smem_u32(rowbuf) + uint32_t(lane & 1) * uint32_t(nB) * 16u;
  const uint32_t rbWrite = smem_u32(rowbuf) + uint32_t(lane == nLanes - 1 ? nB : 0) * 16u;
  const uint32_t feederM = feeder ? 1u : 0u, writerM = writer ? 1u : 0u;
  const int nChunks = G.nIter >> 4;

  // one iteration: the block of four cells at 16-byte block offset b16 of the lane's current row
  auto step = [&](const uint32_t rw) {
    // ---- the rare blocks of a row, one branch: the block after the wrap (next row; also fetches the side record of the
    // row after it), the wrap block itself (the row's last two columns -- Triangle predictor folded
    // into D2/D1 by kernel H -- and columns 0,1 of the lane's next row; feeders take theirs from the stream except
    // before their first row)
    bool fix = false;
    uint32_t badM = validM;
    if (b16 >= wrapB16) {
      if (b16 > wrapB16) {
        b16 = 0;
        row += rpg;
        rowp += groupStep;
        curD2 = sideNext.z;
        curD1 = sideNext.w;
        validM = (computing && row < R) ? 0xFF800000u : 0u;
        badM = validM;
        started = 1u;
        // the side record of the row after this one: needed in this row's wrap block, a whole row of iterations from now
        const int nr = row + rpg;
        sideNext = (computing && nr < R) ? side[nr] : make_int4(0, 0, 0, 0);
      } else {
        badM = 0u;  // the stencil results of the wrap block are discarded
        if (!feeder || !started) {
          fix = true;
          w0 = int32_t(uint32_t(pi3) + uint32_t(curD2));
          w1 = int32_t(uint32_t(w0) + uint32_t(curD1));
          w2 = sideNext.x;
          w3 = sideNext.y;
          g0 = float(w0); g1 = float(w1); g2 = float(w2); g3 = float(w3);
          const bool cur = validM != 0u && (w0 <= -lim || w0 >= lim || w1 <= -lim || w1 >= lim);
          const bool nxt = computing && row + rpg < R && (w2 <= -lim || w2 >= lim || w3 <= -lim || w3 >= lim);
          if (cur || nxt) badBits = 1u;
        }
      }
    }
    // ---- inputs of the four positions: residual bytes (biased floats) or the feeder's finished values
    float x0 = __uint_as_float(__byte_perm(rw, 0x4B400000u, 0x7640));
    float x1 = __uint_as_float(__byte_perm(rw, 0x4B400000u, 0x7641));
    float x2 = __uint_as_float(__byte_perm(rw, 0x4B400000u, 0x7642));
    float x3 = __uint_as_float(__byte_perm(rw, 0x4B400000u, 0x7643));
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; @p ld.shared.v4.f32 {%0, %1, %2, %3}, [%5]; }"
                 : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3)
                 : "r"(feederM), "r"(rbRead + b16));
#pragma unroll
    for (int i = 0; i < 4; i++) { av[i] = av[i + 4]; bv[i] = bv[i + 4]; }
    int32_t out[4];
    float fout[4];
    uint32_t acc = 0;  // exponent bits of t1 that differ from those of [2^22, 2^23): nonzero <=> p outside [-2^21, 2^21) or NaN
    auto cells = [&](auto genTag, const int32_t* ex) {
      constexpr bool GEN = decltype(genTag)::value;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        av[4 + j] = __shfl_up_sync(0xffffffffu, f2, 1);
        bv[4 + j] = __shfl_up_sync(0xffffffffu, av[j], 1);
        // LsDecoder12.java:424-438 -- evaluated left to right in float32, no fused multiply-add
        float p = u[0] * f1;
        p = p + u[1] * av[j + 1];
        p = p + u[2] * av[j + 2];
        p = p + u[3] * av[j + 3];
        p = p + u[4] * av[j + 4];
        p = p + u[5] * f2;
        p = p + u[6] * av[j];
        p = p + u[7] * bv[j];
        p = p + u[8] * bv[j + 1];
        p = p + u[9] * bv[j + 2];
        p = p + u[10] * bv[j + 3];
        p = p + u[11] * bv[j + 4];
        const float t1 = __fadd_rd(p, kMagicHalf);
        acc |= __float_as_uint(t1) ^ 0x4A800000u;
        const float t2 = __fadd_rd(t1, kMagicHalfUp);  // 1.5*2^23 + StrictMath.round(p)
        const float xj = j == 0 ? x0 : j == 1 ? x1 : j == 2 ? x2 : x3;
        float fv = t2 + (xj - bias);
        int32_t iv;
        if constexpr (GEN) {
          iv = int32_t(uint32_t(__float_as_int(t2)) - 0x4B400000u + uint32_t(ex[j]));
          if (!feeder) fv = float(iv);
        } else iv = int32_t(uint32_t(__float_as_int(t2)) + __float_as_uint(xj) - 0x96800080u);  // round(p) + byte - 128
        if (fix) {
          fv = j == 0 ? g0 : j == 1 ? g1 : j == 2 ? g2 : g3;
          iv = j == 0 ? w0 : j == 1 ? w1 : j == 2 ? w2 : w3;
        }
        out[j] = iv;
        fout[j] = fv;
        f2 = f1;
        f1 = fv;
      }
    };
    // ---- exceptions: a zero byte in a tile that has an exception list -> the general form of the block
    bool general = false;
    if (nExc > 0) {
      const bool z = ((rw - 0x01010101u) & ~rw & 0x80808080u) != 0u;
      general = __any_sync(0xffffffffu, z && badM != 0u);
    }
    if (general) {
      int32_t ex[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t byte = (rw >> (8 * j)) & 0xffu;
        ex[j] = int32_t(byte) - 128;
        if (byte == 0u && badM != 0u) ex[j] = wave_exception(exc, nExc, (row - 2) * wInterior + int(b16 >> 2) + j);
      }
      cells(std::true_type{}, ex);
    } else cells(std::false_type{}, nullptr);
    badBits |= acc & badM;
    // ---- stores: columns 4b .. 4b+3 of the row (two from the previous block); whole sectors when WIDE
    const int4 sq = make_int4(pi2, pi3, out[0], out[1]);
    if (WIDE) {
      asm volatile("{ .reg .pred p; setp.ne.b32 p, %9, 0; @p st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8}; }" ::"l"(rowp + b16 - 16), "r"(sqPrev.x),
                   "r"(sqPrev.y), "r"(sqPrev.z), "r"(sqPrev.w), "r"(sq.x), "r"(sq.y), "r"(sq.z), "r"(sq.w), "r"(validM & (b16 << 19) & 0x00800000u)
                   : "memory");
      sqPrev = sq;
    } else if (validM) *reinterpret_cast<int4*>(rowp + b16) = sq;
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %0, 0; @p st.shared.v4.f32 [%1], {%2, %3, %4, %5}; }" ::"r"(writerM & started), "r"(rbWrite + b16),
                 "f"(fout[0]), "f"(fout[1]), "f"(fout[2]), "f"(fout[3])
                 : "memory");
    pi2 = out[2];
    pi3 = out[3];
    b16 += 16u;
    __syncwarp();
  };

  for (int chunk = 0; chunk < nChunks; chunk++) {
    if (lane == 0 && chunk + 1 < nChunks) issue(chunk + 1);
    mbar_wait(bar0 + 8u * (chunk & 1), uint32_t(chunk >> 1) & 1u);
    const uint32_t stage = ringLane + uint32_t(kWaveStageBytes) * (chunk & 1);
#pragma unroll 1
    for (uint32_t quad = 0; quad < 4; quad++) {  // sixteen residual bytes = four iterations per shared-memory read
      uint32_t r0, r1, r2, r3;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(stage + ((quad * 16u) ^ swz16)));
      step(r0);
      step(r1);
      step(r2);
      step(r3);
    }
  }
  if (__any_sync(0xffffffffu, badBits != 0u)) {  // outside the fast arithmetic's range: the general kernels redo the tile
    if (lane == 0) A.defer[atomicAdd(A.deferCount, 1)] = tIdx;
  }
}

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<TensorMapEncodeFn>(p);
  }();
  return fn;
}

size_t head_smem_bytes(const LsopFastGeom& g) {
  const size_t nInit = size_t(4 * g.R + 2 * g.C - 9);
  return size_t(kWarps) * (((sizeof(CanonWarpShared) + 15) & ~size_t(15)) + ((nInit * 4 + 15) & ~size_t(15))) + 16;
}
uint32_t text_stage_words(const LsopFastGeom& g) {
  // 5 bits per sample of the tile, at least the default 28 KB
  uint32_t w = uint32_t((uint64_t(g.R) * uint64_t(g.C) * 5 / 8 + 3) / 4);
  if (w < uint32_t(kFastStageWords)) w = kFastStageWords;
  return (w + 255u) & ~255u;
}
size_t text_smem_bytes(const LsopFastGeom& g) {
  return ((canon_fast_smem_bytes(text_stage_words(g)) + 127) & ~size_t(127)) + size_t(g.tileBytes);
}
size_t wave_smem_bytes(const LsopFastGeom& g) {
  return size_t(kWarps) * (2 * kWaveStageBytes) + size_t(kWarps) * (2 * g.nB) * sizeof(float4) + size_t(kWarps) * 16;
}

}  // namespace

bool lsop_fast_geometry(const g4_band_desc& band, const void* grid, LsopFastGeom* out) {
  LsopFastGeom g{};
  g.R = band.tile_rows;
  g.C = band.tile_cols;
  if (g.R < 6 || g.C < 16 || (g.C & 3) != 0 || (band.grid_pitch & 3) != 0 || (reinterpret_cast<uintptr_t>(grid) & 15) != 0) return false;
  g.nB = g.C / 4;
  g.nLanes = g.nB < 32 ? g.nB : 32;
  g.rpg = g.nLanes - 2;
  g.nGroups = (g.R - 2 + g.rpg - 1) / g.rpg;
  const int words = 1 + g.nGroups * g.nB;
  g.laneBytes = ((4 * words + 15) & ~15) + 4;  // == 4 (mod 16): the tensor map's row stride laneBytes - 4 is a multiple of 16
  g.tileBytes = 32 * g.laneBytes;
  g.tilePitch = ((g.tileBytes + g.laneBytes - 5) / (g.laneBytes - 4)) * (g.laneBytes - 4);  // every tensor stride a multiple of the one before
  g.nIter = (words + g.nLanes + 15) & ~15;
  g.wide = ((g.C & 7) == 0 && (band.grid_pitch & 7) == 0 && (reinterpret_cast<uintptr_t>(grid) & 31) == 0) ? 1 : 0;
  if (text_smem_bytes(g) > 224u * 1024u || head_smem_bytes(g) > 224u * 1024u || wave_smem_bytes(g) > 224u * 1024u) return false;
  if (!tensor_map_encoder()) return false;
  *out = g;
  return true;
}
size_t lsop_fast_side_bytes(const LsopFastGeom& g, int nTiles) { return size_t(nTiles) * size_t(g.R) * sizeof(int4); }
size_t lsop_fast_exc_bytes(int nTiles) { return size_t(nTiles) * kExcWords * sizeof(uint32_t); }
size_t lsop_fast_stage_bytes(int smCount) { return size_t(smCount) * 2 * kFastMaxSub * kSpillWords * sizeof(uint32_t); }  // text kernel: <= 2 CTAs per SM
size_t lsop_fast_resid_bytes(const LsopFastGeom& g, int nTiles) { return size_t(kResidGuard) + size_t(nTiles) * size_t(g.tilePitch) + size_t(g.tileBytes) + 4096; }

// Kernels H, T and W over the list positions [0, nTilesUpper).  Tiles the fast path cannot take are appended to
// A.defer / A.deferCount for the general kernels, which the caller launches AFTER this returns.
cudaError_t launch_lsop_decode_fast(const LsopFastArgs& A, int nTilesUpper, int smCount, int* textCounter, cudaStream_t s, int* launches) {
  const LsopFastGeom& g = A.g;
  // dynamic shared memory opt-in: per device and per geometry, so simply set before every launch set (host-side only)
  {
    cudaError_t e = cudaFuncSetAttribute(lsop2_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(head_smem_bytes(g)));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lsop2_text_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(text_smem_bytes(g)));
    if (e == cudaSuccess) {
      if (g.wide) e = cudaFuncSetAttribute(lsop2_wave_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(wave_smem_bytes(g)));
      else e = cudaFuncSetAttribute(lsop2_wave_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(wave_smem_bytes(g)));
    }
    if (e != cudaSuccess) return e;
  }
  // residual scratch as a 3-D tensor {byte x, lane, tile}: lane stride laneBytes - 4, so that box row l starts 4 l
  // bytes EARLIER in lane l's stream -- at iteration `it` every lane needs word it - l of its stream
  alignas(64) CUtensorMap tmap;
  {
    cuuint64_t dims[3] = {cuuint64_t(g.laneBytes) + 256, 32, cuuint64_t(A.a.band.tiles_down) * cuuint64_t(A.a.band.tiles_across)};
    cuuint64_t strides[2] = {cuuint64_t(g.laneBytes - 4), cuuint64_t(g.tilePitch)};
    cuuint32_t box[3] = {64, 32, 1};
    cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = tensor_map_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, A.resid + kResidGuard, dims, strides, box, es,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      fprintf(stderr, "g4: cuTensorMapEncodeTiled failed (%d), laneBytes %d tilePitch %d\n", int(r), g.laneBytes, g.tilePitch);
      return cudaErrorNotSupported;
    }
  }
  const uint32_t stageWords = text_stage_words(g);
  const int nCtasWarp = (nTilesUpper + kWarps - 1) / kWarps;
  lsop2_head_kernel<<<nCtasWarp, kThreads, head_smem_bytes(g), s>>>(A, stageWords * 4u, 0, nTilesUpper);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  LsopFastArgs T = A;
  T.a.counter = textCounter;
  const size_t textSmem = text_smem_bytes(g);
  int perSm = int((227u * 1024u) / (textSmem + 1024));
  if (perSm < 1) perSm = 1;
  if (perSm > 2) perSm = 2;
  int ctas = smCount * perSm;
  if (ctas > nTilesUpper) ctas = nTilesUpper;
  lsop2_text_kernel<<<ctas, kTextThreads, textSmem, s>>>(T, stageWords, 0, nTilesUpper);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (g.wide) lsop2_wave_kernel<true><<<nCtasWarp, kThreads, wave_smem_bytes(g), s>>>(tmap, A, 0, nTilesUpper);
  else lsop2_wave_kernel<false><<<nCtasWarp, kThreads, wave_smem_bytes(g), s>>>(tmap, A, 0, nTilesUpper);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (launches) *launches += 3;
  return cudaSuccess;
}

}  // namespace g4
