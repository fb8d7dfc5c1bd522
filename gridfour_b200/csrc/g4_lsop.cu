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

namespace g4 {

namespace {

union LsopDecShared {
  HuffDecShared h;
  CanonDecShared c;
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
    else if (h.type == 1) status = G4_ERR_UNSUPPORTED;  // TODO(next): zlib streams (needs the GPU inflate)
    if (status == G4_OK) {
      BitSrc src;
      src.init(packing + h.headerSize, len - h.headerSize);
      if (h.type == 2) {
        uint32_t endBit = 0, nv = 0;
        CellSink s1{t, kStreamLsopInit};
        CellSink s2{t, kStreamLsopInterior};
        if (!canon_decode_stream(S.c, src, 0, nInit, 32768u, s1, &endBit, &nv) || nv != nInit) status = G4_ERR_FORMAT;
        else if (!canon_decode_stream(S.c, src, endBit, nInterior, 0u, s2, &endBit, &nv) || nv != nInterior) status = G4_ERR_FORMAT;
      } else {  // type 0: two legacy Huffman streams back to back, each M32 coded (LsDecoder12.java:119-124)
        if (h.nInitCodes < nInit || h.nInitCodes > 6 * nInit || h.nInteriorCodes < nInterior || h.nInteriorCodes > 6 * nInterior)
          status = G4_ERR_FORMAT;
        else {
          uint8_t* m32a = a.scratch + size_t(blockIdx.x) * a.scratchStride;
          uint8_t* m32b = m32a + ((size_t(h.nInitCodes) + 31) & ~size_t(15));
          uint32_t endBit = 0;
          if (!huffman_decode_stream(S.h, src, 0, h.nInitCodes, m32a, &endBit)) status = G4_ERR_FORMAT;
          else if (!huffman_decode_stream(S.h, src, endBit, h.nInteriorCodes, m32b, &endBit)) status = G4_ERR_FORMAT;
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

cudaError_t launch_lsop_decode(const DecodeArgs& a, float* coef, int nCtas, int nTilesUpper, cudaStream_t s) {
  lsop_decode_entropy_kernel<<<nCtas, kThreads, 0, s>>>(a, coef);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  lsop_wavefront_kernel<<<(nTilesUpper + kWarps - 1) / kWarps, kThreads, 0, s>>>(a, coef);
  return cudaGetLastError();
}

}  // namespace g4
