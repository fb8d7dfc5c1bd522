// g4_lsop.cu -- LSOP12 (Lewis-Smith optimal predictor, 12 coefficients) on sm_100a.
//
// Reference (under /root/reference/core/src/main/java/org/gridfour/lsop/):
//   LsDecoder12.java:94-470 (decode, unpackInitializers, unpackInterior), LsHeader.java:104-265,
//   LsOptimalPredictor12.java:109-383, LsEncoder12.java:122-219
//
// Decode is two kernels over the LSOP tiles of a band:
//   A  one CTA per tile: header parse, entropy stage (canonical Huffman type 2, legacy Huffman + M32 type 0)
//      with residuals scattered to their raster cells, then rows 0/1 and columns 0/1 by prefix scans;
//   B  one WARP per tile: the causal 12-tap float32 stencil is nonlinear (rounding), so it cannot be a scan.
//      Cells with equal c+3r are independent: lane l owns row r0+l and runs three columns behind lane l-1.
//      Row r-1/r-2 operands travel between lanes by warp shuffles; the two last columns of every row
//      (Triangle predictor) ride the same schedule.  Float arithmetic is strictly left to right, no FMA.
#include "g4_kernels.h"
#include "g4_predict.cuh"
#include "g4_huffdec.cuh"
#include "g4_canon.cuh"
#include "g4_canon_enc.cuh"
#include "g4_inflate.cuh"
#include "g4_m32stream.cuh"

namespace g4 {

namespace {

union LsopDecShared {
  HuffDecShared h;
  CanonDecShared c;
  InflateWarpShared inf;
};

// StrictMath.round(float): floor(a + 1/2) evaluated on the bit pattern (java.lang.Math.round(float), JDK >= 8)
__device__ __forceinline__ int32_t java_round(float a) {
  int32_t bits = __float_as_int(a);
  int biasedExp = (bits & 0x7F800000) >> 23;
  int shift = (24 - 2 + 127) - biasedExp;
  if ((shift & -32) == 0) {
    int32_t r = (bits & 0x007FFFFF) | 0x00800000;
    if (bits < 0) r = -r;
    return ((r >> shift) + 1) >> 1;
  }
  return __float2int_rz(a);  // (int) a: saturating, NaN -> 0
}

struct LsHeaderInfo {
  int type;          // 0 legacy Huffman, 1 Deflate, 2 canonical Huffman
  int32_t seed;
  uint32_t nInitCodes, nInteriorCodes;
  uint32_t headerSize;
  bool ok;
};

// LsHeader(byte[],int) (LsHeader.java:104-189).  Coefficients are written to coef[0..11].
__device__ inline LsHeaderInfo parse_ls_header(const uint8_t* p, uint32_t len, float* coef, bool writeCoef) {
  LsHeaderInfo h;
  h.ok = false;
  h.type = -1;
  h.seed = 0;
  h.nInitCodes = h.nInteriorCodes = 0;
  h.headerSize = 0;
  if (len < 3 + 4 + 48) return h;
  uint32_t off = 1;
  bool legacy = (p[1] & 0x40) == 0;
  bool cks;
  if (legacy) {
    if (p[off++] != 12) return h;
  } else {
    h.type = p[off] & 0x0f;
    cks = (p[off] & 0x80) != 0;
    off++;
    if (p[off++] != 12) return h;
  }
  h.seed = int32_t(load_le32(p + off));
  off += 4;
  for (int i = 0; i < 12; i++) {
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
  if (cks) off += 4;
  if (off > len || h.type < 0 || h.type > 2) return h;
  h.headerSize = off;
  h.ok = true;
  return h;
}

struct CellSink {
  TileView t;
  int order;
  __device__ __forceinline__ void operator()(uint32_t k, int32_t v) const {
    int r, c;
    stream_to_cell(order, int(k), t.R, t.C, &r, &c);
    t.at(r, c) = v;
  }
};

}  // namespace

// ---- kernel A ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) lsop_decode_entropy_kernel(DecodeArgs a, float* coefOut) {
  __shared__ LsopDecShared S;
  __shared__ int sTile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (;;) {
    __syncthreads();
    if (tid == 0) sTile = atomicAdd(a.counter, 1);
    __syncthreads();
    const int li = sTile;
    if (li >= *a.listCount) break;
    const int tIdx = a.list[li];
    const TileView t = tile_view(a.band, a.grid, tIdx);
    const int R = t.R, C = t.C;
    const uint8_t* packing = a.arena + a.offsets[tIdx];
    const uint32_t len = a.lens[tIdx];
    int status = G4_OK;
    LsHeaderInfo h = parse_ls_header(packing, len, coefOut + size_t(tIdx) * 12, tid == 0);
    const uint32_t nInit = uint32_t(4 * R + 2 * C - 9);
    const uint32_t nInterior = uint32_t(R - 2) * uint32_t(C - 4);
    if (!h.ok || R < 6 || C < 6) status = G4_ERR_FORMAT;
    if (status == G4_OK) {
      BitSrc src;
      src.init(packing + h.headerSize, len - h.headerSize);
      if (h.type == 2) {
        uint32_t endBit = 0, nv = 0;
        CellSink s1{t, kStreamLsopInit};
        CellSink s2{t, kStreamLsopInterior};
        if (!canon_decode_stream(S.c, src, 0, nInit, 32768u, s1, &endBit, &nv) || nv != nInit) status = G4_ERR_FORMAT;
        else if (!canon_decode_stream(S.c, src, endBit, nInterior, 0u, s2, &endBit, &nv) || nv != nInterior) status = G4_ERR_FORMAT;
      } else {
        // type 0: two legacy Huffman streams back to back, each M32 coded (LsDecoder12.java:119-124)
        // type 1: two zlib streams; the second starts where the first one's input ended (:126-145)
        if (h.nInitCodes < nInit || h.nInitCodes > 6 * nInit || h.nInteriorCodes < nInterior || h.nInteriorCodes > 6 * nInterior)
          status = G4_ERR_FORMAT;
        else {
          uint8_t* m32a = a.scratch + size_t(blockIdx.x) * a.scratchStride;
          uint8_t* m32b = m32a + ((size_t(h.nInitCodes) + 31) & ~size_t(15));
          uint32_t endBit = 0;
          bool entropyOk;
          if (h.type == 1) {
            __shared__ int sInfOk;
            if (warp == 0) {
              const uint8_t* z = packing + h.headerSize;
              const uint32_t zLen = len - h.headerSize;
              uint32_t produced = 0, consumed = 0;
              int rc = inflate_warp(S.inf, z, zLen, m32a, h.nInitCodes, &produced, &consumed);
              bool ok = rc == kInfOk && produced == h.nInitCodes && consumed <= zLen;
              if (ok) {
                uint32_t c1 = consumed;
                rc = inflate_warp(S.inf, z + c1, zLen - c1, m32b, h.nInteriorCodes, &produced, &consumed);
                ok = rc == kInfOk && produced == h.nInteriorCodes;
              }
              if (lane == 0) sInfOk = ok ? 1 : 0;
            }
            __syncthreads();
            entropyOk = sInfOk != 0;
          } else {
            entropyOk = huffman_decode_stream(S.h, src, 0, h.nInitCodes, m32a, &endBit) &&
                        huffman_decode_stream(S.h, src, endBit, h.nInteriorCodes, m32b, &endBit);
          }
          if (!entropyOk) status = G4_ERR_FORMAT;
          else {
            __syncthreads();
            if (!m32_parse_to_cells(m32a, h.nInitCodes, kStreamLsopInit, t, nInit, S.h.scan)) status = G4_ERR_FORMAT;
            else if (!m32_parse_to_cells(m32b, h.nInteriorCodes, kStreamLsopInterior, t, nInterior, S.h.scan)) status = G4_ERR_FORMAT;
          }
        }
      }
    }
    if (status == G4_OK) {
      // LsDecoder12.unpackInitializers (:204-241): row 0 and column 0 by differencing, row 1 and column 1 by Triangle
      __syncthreads();
      if (tid == 0) t.at(0, 0) = h.seed;
      __syncthreads();
      uint32_t* scan = S.h.scan;
      if (warp == 0) row_scan_warp(t.row(0), C, 0);
      column0_scan(t, scan);  // ends with __syncthreads
      if (warp == 0) {
        // row 1: T[c] = v[1][c] - v[0][c];  T[c] = T[c-1] + residual(1,c)
        uint32_t carry = uint32_t(t.at(1, 0)) - uint32_t(t.at(0, 0));
        for (int c0 = 1; c0 < C; c0 += 32) {
          int c = c0 + lane;
          uint32_t x = c < C ? uint32_t(t.at(1, c)) : 0u;
          uint32_t inc = warp_inclusive_scan(x);
          if (c < C) t.at(1, c) = int32_t(carry + inc + uint32_t(t.at(0, c)));
          carry += __shfl_sync(0xffffffffu, inc, 31);
        }
      }
      __syncthreads();
      {
        // column 1, rows 2..: U[r] = v[r][1] - v[r][0];  U[r] = U[r-1] + residual(r,1)
        uint32_t carry = uint32_t(t.at(1, 1)) - uint32_t(t.at(1, 0));
        for (int r0 = 2; r0 < R; r0 += kThreads) {
          int r = r0 + tid;
          uint32_t x = r < R ? uint32_t(t.at(r, 1)) : 0u;
          uint32_t tot;
          uint32_t ex = block_exclusive_scan(x, scan, &tot);
          if (r < R) t.at(r, 1) = int32_t(carry + ex + x + uint32_t(t.at(r, 0)));
          carry += tot;
        }
      }
    }
    __syncthreads();
    if (tid == 0) a.status[tIdx] = status;
  }
}

// ---- kernel B: wavefront ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) lsop_wavefront_kernel(DecodeArgs a, const float* coef) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int li = blockIdx.x * kWarps + warp;
  if (li >= *a.listCount) return;
  const int tIdx = a.list[li];
  if (a.status[tIdx] != G4_OK) return;
  const TileView t = tile_view(a.band, a.grid, tIdx);
  const int R = t.R, C = t.C;
  float u[12];
#pragma unroll
  for (int i = 0; i < 12; i++) u[i] = coef[size_t(tIdx) * 12 + i];

  for (int r0 = 2; r0 < R; r0 += 32) {
    const int r = r0 + lane;
    const bool active = r < R;
    const int rr = active ? r : R - 1;  // clamp so inactive lanes read valid memory
    int32_t* rowp = t.row(rr);
    const int32_t* up1 = t.row(rr - 1);
    const int32_t* up2 = t.row(rr - 2);
    // windows positioned for column c = -1: [c-2 .. c+2] = [-3 .. 1]
    float z1 = float(rowp[1]), z6 = float(rowp[0]);
    int32_t o1 = rowp[1];
    float z7 = 0.f, z2 = 0.f, z3 = 0.f, z4 = float(up1[0]), z5 = float(up1[1]);
    int32_t i2 = 0, i3 = 0, i4 = up1[0], i5 = up1[1];
    float z8 = 0.f, z9 = 0.f, z10 = 0.f, z11 = float(up2[0]), z12 = float(up2[1]);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    int32_t lastOut = 0;
    const int lastLane = (R - 1 - r0) < 31 ? (R - 1 - r0) : 31;
    const int nSteps = C + 3 * lastLane;  // lane l works on column s - 3l at step s; the last lane ends at column C-1
    __syncwarp();
    for (int s = 0; s < nSteps; s++) {
      const int c = s - 3 * lane;
      int32_t recvA = __shfl_up_sync(0xffffffffu, lastOut, 1);
      float recvB = __shfl_up_sync(0xffffffffu, a2, 1);
      const bool shifting = c >= 0 && c < C;
      if (lane == 0 && shifting) {
        // rows r0-1 and r0-2 were finished by the previous row group (or by kernel A for rows 0,1)
        int cc = c + 2 < C ? c + 2 : C - 1;
        recvA = __ldcg(up1 + cc);
        recvB = float(__ldcg(up2 + cc));
      }
      if (shifting) {
        float fa = float(recvA);
        z7 = z2; z2 = z3; z3 = z4; z4 = z5; z5 = fa;
        i2 = i3; i3 = i4; i4 = i5; i5 = recvA;
        z8 = z9; z9 = z10; z10 = z11; z11 = z12; z12 = recvB;
        a2 = a1; a1 = a0; a0 = fa;
      }
      if (active && c >= 2 && c < C) {
        int32_t res = rowp[c];
        int32_t val;
        if (c < C - 2) {
          // LsDecoder12.java:424-438 -- evaluated left to right in float32, no fused multiply-add
          float p = u[0] * z1;
          p = p + u[1] * z2;
          p = p + u[2] * z3;
          p = p + u[3] * z4;
          p = p + u[4] * z5;
          p = p + u[5] * z6;
          p = p + u[6] * z7;
          p = p + u[7] * z8;
          p = p + u[8] * z9;
          p = p + u[9] * z10;
          p = p + u[10] * z11;
          p = p + u[11] * z12;
          val = int32_t(uint32_t(java_round(p)) + uint32_t(res));
        } else {
          // last two columns: Triangle predictor (LsDecoder12.java:459-468)
          val = int32_t(uint32_t(res) + ((uint32_t(o1) + uint32_t(i3)) - uint32_t(i2)));
        }
        rowp[c] = val;
        lastOut = val;
        z6 = z1;
        z1 = float(val);
        o1 = val;
      }
    }
    __syncwarp();
    __threadfence_block();
  }
}

// =====================================================================================================
// Encode (LsEncoder12.encode :122-219 with the canonical-Huffman body; LsOptimalPredictor12.encode :109-292)
// =====================================================================================================
namespace {

constexpr int kMomentQuantities = 104;  // 13 sums + 91 products (upper triangle incl. diagonal)
constexpr int kMomentPerRole = 26;

// quantity q: q < 13 -> z[q]; otherwise the (i,j) product with i <= j in row-major order of the upper triangle
__host__ __device__ constexpr int moment_i(int q) {
  int k = q - 13, i = 0;
  while (k >= 13 - i) { k -= 13 - i; i++; }
  return i;
}
__host__ __device__ constexpr int moment_j(int q) {
  int k = q - 13, i = 0;
  while (k >= 13 - i) { k -= 13 - i; i++; }
  return i + k;
}

template <int ROLE>
__device__ __forceinline__ void moment_accumulate(const double (&z)[13], double (&acc)[kMomentPerRole]) {
#pragma unroll
  for (int a = 0; a < kMomentPerRole; a++) {
    constexpr int dummy = 0;
    (void)dummy;
    const int q = a * 4 + ROLE;
    if (q < 13) acc[a] += z[q];
    else acc[a] += z[moment_i(q)] * z[moment_j(q)];
  }
}

struct LsopEncShared {
  CanonEncShared E;
  BitWindow W;
  double part[kWarps][kMomentPerRole];
  double sums[kMomentQuantities];
  double M[13][13];
  double rhs[13];
  float u[12];
  int singular;
};

// JAMA LUDecomposition ctor + solve (util/jama/LUDecomposition.java:70-135, :253-286), one thread, FP64,
// identical operation order (compiled with -fmad=false).
__device__ bool lu_solve13(double (*LU)[13], double* X) {
  const int n = 13;
  int piv[13];
  for (int i = 0; i < n; i++) piv[i] = i;
  double col[13];
  for (int j = 0; j < n; j++) {
    for (int i = 0; i < n; i++) col[i] = LU[i][j];
    for (int i = 0; i < n; i++) {
      int kmax = i < j ? i : j;
      double s = 0.0;
      for (int k = 0; k < kmax; k++) s += LU[i][k] * col[k];
      col[i] -= s;
      LU[i][j] = col[i];
    }
    int p = j;
    for (int i = j + 1; i < n; i++)
      if (fabs(col[i]) > fabs(col[p])) p = i;
    if (p != j) {
      for (int k = 0; k < n; k++) { double t = LU[p][k]; LU[p][k] = LU[j][k]; LU[j][k] = t; }
      int k = piv[p]; piv[p] = piv[j]; piv[j] = k;
    }
    if (LU[j][j] != 0.0)
      for (int i = j + 1; i < n; i++) LU[i][j] /= LU[j][j];
  }
  for (int j = 0; j < n; j++)
    if (LU[j][j] == 0.0) return false;
  double B[13];
  for (int i = 0; i < n; i++) B[i] = X[piv[i]];
  for (int i = 0; i < n; i++) X[i] = B[i];
  for (int k = 0; k < n; k++)
    for (int i = k + 1; i < n; i++) X[i] -= X[k] * LU[i][k];
  for (int k = n - 1; k >= 0; k--) {
    X[k] /= LU[k][k];
    for (int i = 0; i < k; i++) X[i] -= X[k] * LU[i][k];
  }
  return true;
}

struct LsInitGet {  // initializer residuals in stream order (LsOptimalPredictor12.java:143-209)
  TileView t;
  __device__ __forceinline__ int32_t operator()(uint32_t k) const {
    int r, c;
    stream_to_cell(kStreamLsopInit, int(k), t.R, t.C, &r, &c);
    return residual_at(G4_PRED_TRIANGLE, t, r, c);  // row 0 / column 0 degrade to plain differences
  }
};

struct LsInteriorGet {  // interior residuals (LsOptimalPredictor12.java:254-282)
  TileView t;
  const float* u;
  __device__ __forceinline__ int32_t operator()(uint32_t k) const {
    const int w = t.C - 4;
    int rr = int(k) / w;
    int r = 2 + rr, c = 2 + int(k) - rr * w;
    const int32_t* row = t.row(r);
    const int32_t* up1 = row - t.pitch;
    const int32_t* up2 = up1 - t.pitch;
    float p = u[0] * float(row[c - 1]);
    p = p + u[1] * float(up1[c - 1]);
    p = p + u[2] * float(up1[c]);
    p = p + u[3] * float(up1[c + 1]);
    p = p + u[4] * float(up1[c + 2]);
    p = p + u[5] * float(row[c - 2]);
    p = p + u[6] * float(up1[c - 2]);
    p = p + u[7] * float(up2[c - 2]);
    p = p + u[8] * float(up2[c - 1]);
    p = p + u[9] * float(up2[c]);
    p = p + u[10] * float(up2[c + 1]);
    p = p + u[11] * float(up2[c + 2]);
    return int32_t(uint32_t(row[c]) - uint32_t(java_round(p)));
  }
};

}  // namespace

__global__ void __launch_bounds__(kThreads) lsop_encode_kernel(EncodeArgs a) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  LsopEncShared& S = *reinterpret_cast<LsopEncShared*>(smemRaw);
  __shared__ int sTile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nTiles = a.band.tiles_down * a.band.tiles_across;
  uint8_t* pm = a.scratch + size_t(blockIdx.x) * a.scratchStride;
  for (;;) {
    __syncthreads();
    if (tid == 0) sTile = atomicAdd(a.counter, 1);
    __syncthreads();
    const int tIdx = sTile;
    if (tIdx >= nTiles) break;
    const TileView t = tile_view(a.band, a.grid, tIdx);
    const int R = t.R, C = t.C;
    uint32_t* outWords = reinterpret_cast<uint32_t*>(a.slots + size_t(tIdx) * a.slotBytes);
    const uint32_t capWords = uint32_t(a.slotBytes / 4);
    if (R < 6 || C < 6) {  // LsOptimalPredictor12.java:114-116
      if (tid == 0) { a.lens[tIdx] = 0; a.preds[tIdx] = 0; a.status[tIdx] = G4_DECLINED; }
      continue;
    }
    // ---- normal equations: s[i] += z[i], c[i][j] += z[i]*z[j] over the interior cells (:319-344) --------
    {
      const int role = warp & 3, group = warp >> 2;
      const int w = C - 4;
      const int nInterior = (R - 2) * w;
      double acc[kMomentPerRole];
#pragma unroll
      for (int i = 0; i < kMomentPerRole; i++) acc[i] = 0.0;
      for (int m = group * 32 + lane; m < nInterior; m += 64) {
        int rr = m / w;
        int r = 2 + rr, c = 2 + m - rr * w;
        const int32_t* row = t.row(r);
        const int32_t* up1 = row - t.pitch;
        const int32_t* up2 = up1 - t.pitch;
        double z[13];
        z[0] = row[c]; z[1] = row[c - 1]; z[2] = up1[c - 1]; z[3] = up1[c]; z[4] = up1[c + 1]; z[5] = up1[c + 2];
        z[6] = row[c - 2]; z[7] = up1[c - 2]; z[8] = up2[c - 2]; z[9] = up2[c - 1]; z[10] = up2[c]; z[11] = up2[c + 1];
        z[12] = up2[c + 2];
        switch (role) {
          case 0: moment_accumulate<0>(z, acc); break;
          case 1: moment_accumulate<1>(z, acc); break;
          case 2: moment_accumulate<2>(z, acc); break;
          default: moment_accumulate<3>(z, acc); break;
        }
      }
#pragma unroll
      for (int i = 0; i < kMomentPerRole; i++) {
        double v = acc[i];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0) S.part[warp][i] = v;
      }
      __syncthreads();
      if (tid < kMomentQuantities) S.sums[tid] = S.part[tid & 3][tid >> 2] + S.part[(tid & 3) + 4][tid >> 2];
      __syncthreads();
    }
    if (tid == 0) {
      // design matrix (:346-368): 12x12 moment block bordered by the sums (Lagrange constraint)
      double cfull[13][13];
      for (int q = 13; q < kMomentQuantities; q++) {
        int i = moment_i(q), j = moment_j(q);
        cfull[i][j] = S.sums[q];
        cfull[j][i] = S.sums[q];
      }
      for (int i = 0; i < 13; i++) for (int j = 0; j < 13; j++) S.M[i][j] = 0.0;
      for (int i = 1; i < 13; i++) {
        for (int j = 1; j < 13; j++) S.M[i - 1][j - 1] = cfull[i][j];
        S.M[i - 1][12] = S.sums[i];
      }
      for (int j = 1; j < 13; j++) S.M[12][j - 1] = S.sums[j];
      for (int i = 1; i < 13; i++) S.rhs[i - 1] = cfull[0][i];
      S.rhs[12] = S.sums[0];
      bool ok = lu_solve13(S.M, S.rhs);
      S.singular = ok ? 0 : 1;
      if (ok)
        for (int i = 0; i < 12; i++) S.u[i] = __double2float_rn(S.rhs[i]);
    }
    __syncthreads();
    if (S.singular) {  // RuntimeException("Matrix is singular.") -> null (:377-381)
      if (tid == 0) { a.lens[tIdx] = 0; a.preds[tIdx] = 0; a.status[tIdx] = G4_DECLINED; }
      continue;
    }
    // ---- packing: revised header (LsHeader.java:210-265) + two canonical streams in one bit store --------
    const uint32_t nInit = uint32_t(4 * R + 2 * C - 9);
    const uint32_t nInterior = uint32_t(R - 2) * uint32_t(C - 4);
    BitOut o;
    bitwin_reset(S.W, o, outWords, capWords);
    if (tid == 0) {
      WinSink sink{S.W.win, 0};
      sink.put(uint32_t(a.codecIndex) & 0xffu, 8);
      sink.put(0x40u | 2u, 8);  // revision flag | COMPRESSION_TYPE_CANON_HUFFMAN
      sink.put(12, 8);
      sink.put(uint32_t(t.at(0, 0)), 32);
      for (int i = 0; i < 12; i++) sink.put(__float_as_uint(S.u[i]), 32);
    }
    o.bitPos = 55u * 8u;
    __syncthreads();
    LsInitGet g1{t};
    LsInteriorGet g2{t, S.u};
    bool bad = false;
    if (!canon_histogram(S.E, g1, nInit)) bad = true;
    else {
      canon_build_code(S.E, pm);
      canon_emit_stream(S.E, S.W, o, g1, nInit);
      if (!canon_histogram(S.E, g2, nInterior)) bad = true;
      else {
        canon_build_code(S.E, pm);
        canon_emit_stream(S.E, S.W, o, g2, nInterior);
      }
    }
    if (bad) {
      if (tid == 0) { a.lens[tIdx] = 0; a.preds[tIdx] = 0; a.status[tIdx] = G4_DECLINED; }
      continue;
    }
    bitwin_finish(S.W, o);
    if (tid == 0) {
      uint32_t len = (o.bitPos + 7) >> 3;
      a.lens[tIdx] = len;
      a.preds[tIdx] = 2;  // compression type: canonical Huffman
      a.status[tIdx] = len <= a.slotBytes ? G4_OK : G4_ERR_CAPACITY;
    }
  }
}

// ---- Deflate alternative (LsEncoder12.java:170-218) ---------------------------------------------------------------
// After lsop_encode_kernel has written the canonical-Huffman packing (type 2) into the tile's slot, the initializer
// and interior residuals are M32 coded (LsOptimalPredictor12.java:143-282 fills both the int and the M32 forms),
// deflated at level 6 by the stream workers, and the packing is replaced by the type-1 form when
// insideN < canonLength and initN + insideN < canonLength (the reference compares bodies only, headers excluded).
namespace {
__device__ __forceinline__ void load_slot_coefficients(const uint8_t* slot, float* u) {
  if (threadIdx.x < 12) u[threadIdx.x] = __uint_as_float(load_le32(slot + 7 + 4 * threadIdx.x));
  __syncthreads();
}
}  // namespace

__global__ void __launch_bounds__(kThreads) lsop_m32_size_kernel(EncodeArgs a, uint32_t* inLen) {
  __shared__ uint32_t scan[kWarps + 1];
  __shared__ float u[12];
  const int nTiles = a.band.tiles_down * a.band.tiles_across;
  for (int tIdx = blockIdx.x; tIdx < nTiles; tIdx += gridDim.x) {
    __syncthreads();
    if (a.status[tIdx] != G4_OK || a.lens[tIdx] < 55) {  // declined (too small / singular): nothing to deflate
      if (threadIdx.x == 0) inLen[2 * tIdx] = inLen[2 * tIdx + 1] = 0;
      continue;
    }
    const TileView t = tile_view(a.band, a.grid, tIdx);
    load_slot_coefficients(a.slots + size_t(tIdx) * a.slotBytes, u);
    const uint32_t nInit = uint32_t(4 * t.R + 2 * t.C - 9);
    const uint32_t nInterior = uint32_t(t.R - 2) * uint32_t(t.C - 4);
    uint32_t s0 = m32_stream_size(LsInitGet{t}, nInit, scan);
    uint32_t s1 = m32_stream_size(LsInteriorGet{t, u}, nInterior, scan);
    if (threadIdx.x == 0) { inLen[2 * tIdx] = s0; inLen[2 * tIdx + 1] = s1; }
  }
}

__global__ void __launch_bounds__(kThreads) lsop_m32_write_kernel(EncodeArgs a, const uint32_t* inLen, const uint64_t* inOff,
                                                                  uint8_t* inBuf) {
  __shared__ uint32_t scan[kWarps + 1];
  __shared__ float u[12];
  const int nTiles = a.band.tiles_down * a.band.tiles_across;
  for (int tIdx = blockIdx.x; tIdx < nTiles; tIdx += gridDim.x) {
    __syncthreads();
    if (inLen[2 * tIdx + 1] == 0) continue;
    const TileView t = tile_view(a.band, a.grid, tIdx);
    load_slot_coefficients(a.slots + size_t(tIdx) * a.slotBytes, u);
    const uint32_t nInit = uint32_t(4 * t.R + 2 * t.C - 9);
    const uint32_t nInterior = uint32_t(t.R - 2) * uint32_t(t.C - 4);
    uint8_t* d0 = inBuf + inOff[2 * tIdx];
    uint8_t* d1 = inBuf + inOff[2 * tIdx + 1];
    uint32_t w0 = m32_stream_write(LsInitGet{t}, nInit, d0, scan);
    uint32_t w1 = m32_stream_write(LsInteriorGet{t, u}, nInterior, d1, scan);
    if (threadIdx.x < 16) { d0[w0 + threadIdx.x] = 0; d1[w1 + threadIdx.x] = 0; }
  }
}

__global__ void __launch_bounds__(kThreads) lsop_pick_kernel(EncodeArgs a, const uint32_t* inLen, const uint64_t* inOff,
                                                             const uint8_t* outBuf, const uint32_t* outLen) {
  const int tid = threadIdx.x;
  const int nTiles = a.band.tiles_down * a.band.tiles_across;
  for (int tIdx = blockIdx.x; tIdx < nTiles; tIdx += gridDim.x) {
    if (inLen[2 * tIdx + 1] == 0) continue;
    const uint32_t canonLength = a.lens[tIdx] - 55u;
    const uint32_t initN = outLen[2 * tIdx], insideN = outLen[2 * tIdx + 1];
    if (insideN == 0 || insideN >= canonLength) continue;         // LsEncoder12.java:185-190
    if (initN == 0 || initN + insideN >= canonLength) continue;   // :197-201
    const uint32_t len = 63u + initN + insideN;
    if (len > a.slotBytes) {
      if (tid == 0) a.status[tIdx] = G4_ERR_CAPACITY;
      continue;
    }
    uint8_t* slot = a.slots + size_t(tIdx) * a.slotBytes;
    if (tid == 0) {
      slot[1] = uint8_t(0x40u | 1u);  // revision flag | COMPRESSION_TYPE_DEFLATE (LsHeader.java:220-245)
      const uint32_t n0 = inLen[2 * tIdx], n1 = inLen[2 * tIdx + 1];
      for (int k = 0; k < 4; k++) { slot[55 + k] = uint8_t(n0 >> (8 * k)); slot[59 + k] = uint8_t(n1 >> (8 * k)); }
      a.lens[tIdx] = len;
      a.preds[tIdx] = 1;
    }
    const uint8_t* s0 = outBuf + inOff[2 * tIdx] + 112ull * uint64_t(2 * tIdx);
    const uint8_t* s1 = outBuf + inOff[2 * tIdx + 1] + 112ull * uint64_t(2 * tIdx + 1);
    for (uint32_t i = tid; i < initN; i += kThreads) slot[63 + i] = s0[i];
    for (uint32_t i = tid; i < insideN; i += kThreads) slot[63 + initN + i] = s1[i];
  }
}

cudaError_t launch_lsop_m32_size(const EncodeArgs& a, uint32_t* inLen, int nCtas, cudaStream_t s) {
  lsop_m32_size_kernel<<<nCtas, kThreads, 0, s>>>(a, inLen);
  return cudaGetLastError();
}
cudaError_t launch_lsop_m32_write(const EncodeArgs& a, const uint32_t* inLen, const uint64_t* inOff, uint8_t* inBuf, int nCtas,
                                  cudaStream_t s) {
  lsop_m32_write_kernel<<<nCtas, kThreads, 0, s>>>(a, inLen, inOff, inBuf);
  return cudaGetLastError();
}
cudaError_t launch_lsop_pick(const EncodeArgs& a, const uint32_t* inLen, const uint64_t* inOff, const uint8_t* outBuf,
                             const uint32_t* outLen, int nCtas, cudaStream_t s) {
  lsop_pick_kernel<<<nCtas, kThreads, 0, s>>>(a, inLen, inOff, outBuf, outLen);
  return cudaGetLastError();
}

cudaError_t launch_lsop_encode(const EncodeArgs& a, int nCtas, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(lsop_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(LsopEncShared)));
    if (e != cudaSuccess) return e;
    attr = true;
  }
  lsop_encode_kernel<<<nCtas, kThreads, sizeof(LsopEncShared), s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_lsop_decode(const DecodeArgs& a, float* coef, int nCtas, int nTilesUpper, cudaStream_t s) {
  lsop_decode_entropy_kernel<<<nCtas, kThreads, 0, s>>>(a, coef);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  lsop_wavefront_kernel<<<(nTilesUpper + kWarps - 1) / kWarps, kThreads, 0, s>>>(a, coef);
  return cudaGetLastError();
}

}  // namespace g4
