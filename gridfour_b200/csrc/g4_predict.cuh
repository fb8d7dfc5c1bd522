// g4_predict.cuh -- CTA-cooperative M32 parse and inverse predictors (decode side).
//
// The reference decodes a tile with one serial loop: M32 parse + running sums
// (compress/PredictorModelDifferencing.java:145-167, PredictorModelLinear.java:66-101,
//  PredictorModelTriangle.java:62-98).  Here:
//   * M32 code starts are found with a prefix scan over the 2-state START/CONT byte automaton,
//   * residuals are scattered to their cells of the output raster (HBM, L2-resident while the tile
//     is being worked on),
//   * the predictors are inverted in place with warp-shuffle prefix scans: 1-D row scans for
//     Differencing, a double scan for Linear, row scans + column sums (2-D inclusive scan) for Triangle.
#pragma once
#include "g4_device.cuh"

namespace g4 {

// ---- M32 byte automaton: state S (expect code start) / C (inside a multi-byte code) -------------
// A byte maps {S,C} -> {S,C}; the map is packed in 2 bits: bit0 = f(S), bit1 = f(C) (1 == C).
//   b <= 0x7E : reset   (S->S, C->S)    0b00
//   b == 0x7F : swap    (S->C, C->S)    0b01
//   b == 0x81 : set     (S->C, C->C)    0b11
//   otherwise : identity(S->S, C->C)    0b10   (0x80 = INT_MIN marker in S / payload in C; >=0x82)
__device__ __forceinline__ uint32_t m32_byte_map(uint32_t b) {
  return b <= 0x7Eu ? 0u : b == 0x7Fu ? 1u : b == 0x81u ? 3u : 2u;
}
__device__ __forceinline__ uint32_t m32_apply(uint32_t f, uint32_t s) { return (f >> s) & 1u; }
// compose: first f then g
__device__ __forceinline__ uint32_t m32_compose(uint32_t f, uint32_t g) {
  return m32_apply(g, m32_apply(f, 0)) | (m32_apply(g, m32_apply(f, 1)) << 1);
}

// Parses `n` M32 bytes (in memory visible to the whole CTA) into residuals and stores residual k at the
// cell given by the predictor's stream order.  `expect` = number of residuals the predictor needs.
// Returns true when the stream is well formed.  sm: kWarps+1 words of shared scratch.
__device__ inline bool m32_parse_to_cells(const uint8_t* buf, uint32_t n, int pred, const TileView& t, uint32_t expect,
                                          uint32_t* sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kPer = 16;
  uint32_t carryState = 0, valueBase = 0;
  bool ok = true;
  for (uint32_t chunk0 = 0; chunk0 < n; chunk0 += kThreads * kPer) {
    const uint32_t p0 = chunk0 + tid * kPer;
    uint8_t b[kPer];
    if (p0 < n) {
      uint4 q = *reinterpret_cast<const uint4*>(buf + p0);  // scratch slots are 16-byte aligned and padded
      uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int i = 0; i < kPer; i++) b[i] = uint8_t(w[i >> 2] >> (8 * (i & 3)));
    } else {
#pragma unroll
      for (int i = 0; i < kPer; i++) b[i] = 0;
    }
    const int nMine = p0 < n ? (n - p0 < uint32_t(kPer) ? int(n - p0) : kPer) : 0;
    uint32_t f = 2u;  // identity
#pragma unroll
    for (int i = 0; i < kPer; i++)
      if (i < nMine) f = m32_compose(f, m32_byte_map(b[i]));
    // inclusive scan of maps across the warp, then across warps
    uint32_t inc = f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc = m32_compose(y, inc);
    }
    __syncthreads();
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    uint32_t pre = 2u;  // map of all bytes of earlier warps
    uint32_t all = 2u;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
      uint32_t s = sm[w];
      if (w < warp) pre = m32_compose(pre, s);
      all = m32_compose(all, s);
    }
    uint32_t excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) excl = 2u;
    uint32_t state = m32_apply(m32_compose(pre, excl), carryState);
    // walk my bytes: mark starts
    uint32_t startMask = 0;
    uint32_t s = state;
#pragma unroll
    for (int i = 0; i < kPer; i++)
      if (i < nMine) {
        if (s == 0) startMask |= 1u << i;
        s = m32_apply(m32_byte_map(b[i]), s);
      }
    uint32_t tot;
    uint32_t vi = valueBase + block_exclusive_scan(__popc(startMask), sm, &tot);
    while (startMask) {
      int i = __ffs(startMask) - 1;
      startMask &= startMask - 1;
      int32_t val;
      int len = m32_decode_at(buf, p0 + i, n, &val);
      if (len == 0 || vi >= expect) { ok = false; }
      else {
        int r, c;
        stream_to_cell(pred, int(vi), t.R, t.C, &r, &c);
        t.at(r, c) = val;
      }
      vi++;
    }
    carryState = m32_apply(all, carryState);
    valueBase += tot;
  }
  if (valueBase != expect || carryState != 0) ok = false;
  return __syncthreads_and(ok ? 1 : 0) != 0;
}

// ---- inverse predictors, in place on the raster -------------------------------------------------

// Column 0: cell (0,0) holds the seed, cells (r,0) r>=1 hold v[r][0]-v[r-1][0].  After: values.
__device__ inline void column0_scan(const TileView& t, uint32_t* sm) {
  uint32_t carry = uint32_t(t.at(0, 0));
  for (int r0 = 1; r0 < t.R; r0 += kThreads) {
    int r = r0 + threadIdx.x;
    uint32_t x = r < t.R ? uint32_t(t.at(r, 0)) : 0u;
    uint32_t tot;
    uint32_t ex = block_exclusive_scan(x, sm, &tot);
    if (r < t.R) t.at(r, 0) = int32_t(carry + ex + x);
    carry += tot;
  }
  __syncthreads();
}

// One warp turns one row into its inclusive prefix sums (mode 0), or applies the Linear double scan
// (mode 1: row[0] = value, row[1] = first difference, row[j>=2] = second differences).
__device__ inline void row_scan_warp(int32_t* row, int C, int mode) {
  const int lane = threadIdx.x & 31;
  constexpr int kPer = 8;
  const bool vec = ((reinterpret_cast<uintptr_t>(row) & 15) == 0) && ((C & 3) == 0);
  uint32_t carry1 = 0, carry2 = 0;  // running sums carried across 256-element segments
  const uint32_t v0 = uint32_t(row[0]);
  for (int c0 = 0; c0 < C; c0 += 32 * kPer) {
    const int c = c0 + lane * kPer;
    uint32_t e[kPer];
    if (vec && c + kPer <= C) {
      int4 a = *reinterpret_cast<const int4*>(row + c);
      int4 b = *reinterpret_cast<const int4*>(row + c + 4);
      e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w; e[4] = b.x; e[5] = b.y; e[6] = b.z; e[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < kPer; i++) e[i] = (c + i < C) ? uint32_t(row[c + i]) : 0u;
    }
    if (mode == 1 && c == 0) e[0] = 0u;  // the value itself is not part of the difference chain
    // first scan
    uint32_t loc = 0;
#pragma unroll
    for (int i = 0; i < kPer; i++) { loc += e[i]; e[i] = loc; }
    uint32_t inc = warp_inclusive_scan(loc);
    uint32_t base = carry1 + inc - loc;
#pragma unroll
    for (int i = 0; i < kPer; i++) e[i] += base;
    carry1 += __shfl_sync(0xffffffffu, inc, 31);
    if (mode == 1) {  // second scan over the first differences
      uint32_t loc2 = 0;
#pragma unroll
      for (int i = 0; i < kPer; i++) { loc2 += e[i]; e[i] = loc2; }
      uint32_t inc2 = warp_inclusive_scan(loc2);
      uint32_t base2 = carry2 + inc2 - loc2 + v0;
#pragma unroll
      for (int i = 0; i < kPer; i++) e[i] += base2;
      carry2 += __shfl_sync(0xffffffffu, inc2, 31);
    }
    if (vec && c + kPer <= C) {
      *reinterpret_cast<int4*>(row + c) = make_int4(e[0], e[1], e[2], e[3]);
      *reinterpret_cast<int4*>(row + c + 4) = make_int4(e[4], e[5], e[6], e[7]);
    } else {
#pragma unroll
      for (int i = 0; i < kPer; i++)
        if (c + i < C) row[c + i] = int32_t(e[i]);
    }
  }
}

// Column sums for Triangle: v[r][c] = sum_{i<=r} S[i][c]; one thread per column, coalesced across threads.
__device__ inline void column_sums(const TileView& t) {
  constexpr int kU = 8;
  for (int c = threadIdx.x; c < t.C; c += kThreads) {
    uint32_t run = 0;
    int r = 0;
    for (; r + kU <= t.R; r += kU) {
      uint32_t x[kU];
#pragma unroll
      for (int i = 0; i < kU; i++) x[i] = uint32_t(t.at(r + i, c));
#pragma unroll
      for (int i = 0; i < kU; i++) { run += x[i]; t.at(r + i, c) = int32_t(run); }
    }
    for (; r < t.R; r++) { run += uint32_t(t.at(r, c)); t.at(r, c) = int32_t(run); }
  }
}

// Inverts predictor `pred` (1,2,3) in place.  Precondition: cell (0,0) = seed, every other cell holds
// its residual.  All kThreads threads of the CTA must call.
__device__ inline void predictor_inverse(int pred, const TileView& t, uint32_t* sm) {
  const int warp = threadIdx.x >> 5;
  if (pred == G4_PRED_DIFFERENCING) {
    column0_scan(t, sm);
    for (int r = warp; r < t.R; r += kWarps) row_scan_warp(t.row(r), t.C, 0);
  } else if (pred == G4_PRED_LINEAR) {
    column0_scan(t, sm);
    for (int r = warp; r < t.R; r += kWarps) row_scan_warp(t.row(r), t.C, 1);
  } else {  // Triangle: 2-D inclusive prefix sum of the residual field
    for (int r = warp; r < t.R; r += kWarps) row_scan_warp(t.row(r), t.C, 0);
    __syncthreads();
    column_sums(t);
  }
  __syncthreads();
}

// ---- PredictorModelDifferencingWithNulls (compress/PredictorModelDifferencingWithNulls.java:66-269) -------------------
// The predecessor of cell (r,c) is the previous cell of the row, or for c == 0 the first cell of the previous row;
// when that predecessor is null (or for cell (0,0)) the prediction restarts from the seed.  Null cells are coded as
// INT_MIN.  All of it is pointwise on the ORIGINAL values, so the encode side needs no scan.

// Seed = floor(mean of every valid value whose predecessor is null (or missing) + 0.5), evaluated in FP64 like the
// reference (:79-106).  *nStart = number of such values (0 -> the model declines).  All threads call.
__device__ inline int32_t nulls_seed(const TileView& t, int* nStart) {
  __shared__ unsigned long long sSum;
  __shared__ unsigned int sCnt;
  __syncthreads();
  if (threadIdx.x == 0) { sSum = 0; sCnt = 0; }
  __syncthreads();
  long long sum = 0;
  unsigned int cnt = 0;
  const int n = t.R * t.C;
  for (int i = threadIdx.x; i < n; i += kThreads) {
    const int r = i / t.C, c = i - r * t.C;
    const int32_t v = t.at(r, c);
    if (v == kNull) continue;
    const bool starts = c > 0 ? t.at(r, c - 1) == kNull : (r == 0 ? true : t.at(r - 1, 0) == kNull);
    if (starts) { sum += v; cnt++; }
  }
  if (cnt) { atomicAdd(&sSum, (unsigned long long)sum); atomicAdd(&sCnt, cnt); }
  __syncthreads();
  *nStart = int(sCnt);
  if (sCnt == 0) return 0;
  const double avg = double((long long)sSum) / double(sCnt);
  return int32_t(floor(avg + 0.5));  // values are ints, so the mean is inside the int range
}

// Residual of cell (r,c) (one per cell, INT_MIN for a null cell).
__device__ __forceinline__ int32_t residual_nulls_at(const TileView& t, int r, int c, int32_t seed) {
  const int32_t v = t.at(r, c);
  if (v == kNull) return kNull;
  int32_t prior = seed;
  if (c > 0) { int32_t p = t.at(r, c - 1); if (p != kNull) prior = p; }
  else if (r > 0) { int32_t p = t.at(r - 1, 0); if (p != kNull) prior = p; }
  return int32_t(uint32_t(v) - uint32_t(prior));  // (int)(long delta) wraps
}

// Inverse, in place: every cell holds its residual.  Tiles with nulls are the rare path, so this is the plain
// restatement: column 0 serially, then one thread per row.  All threads call.
__device__ inline void predictor_inverse_nulls(const TileView& t, int32_t seed) {
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t prior = seed;
    for (int r = 0; r < t.R; r++) {
      const int32_t res = t.at(r, 0);
      if (res == kNull) prior = seed;  // a null first cell: the next row restarts from the seed
      else { prior = int32_t(uint32_t(prior) + uint32_t(res)); t.at(r, 0) = prior; }
    }
  }
  __syncthreads();
  for (int r = threadIdx.x; r < t.R; r += kThreads) {
    int32_t* row = t.row(r);
    int32_t prior = row[0] == kNull ? seed : row[0];
    for (int c = 1; c < t.C; c++) {
      const int32_t res = row[c];
      if (res == kNull) prior = seed;
      else { prior = int32_t(uint32_t(prior) + uint32_t(res)); row[c] = prior; }
    }
  }
  __syncthreads();
}

}  // namespace g4
