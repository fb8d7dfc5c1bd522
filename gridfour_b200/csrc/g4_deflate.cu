// g4_deflate.cu -- CodecDeflate and CodecFloat on sm_100a (decode side; zlib streams via g4_inflate.cuh).
//
// Reference (under /root/reference/core/src/main/java/org/gridfour/compress/):
//   CodecDeflate.java:109-154 (decode), :157-228 (encode/compress)
//   CodecFloat.java:300-325 (byte deltas), :328-458 (encodeFloats/decodeFloats)
//
// Decode runs as two kernels per codec: (1) one WARP per zlib stream inflates into an HBM staging region,
// (2) one CTA per tile turns the staged bytes into samples (M32 parse + inverse predictor for CodecDeflate;
// byte-plane delta scans + bit merge for CodecFloat).
#include "g4_kernels.h"
#include "g4_predict.cuh"
#include "g4_inflate.cuh"

namespace g4 {

#ifndef G4_INF_GROUP
#define G4_INF_GROUP 16
#endif
constexpr int kInfGroup = G4_INF_GROUP;                      // lanes per zlib stream in the stand-alone inflate kernels (measured on config 4: 8 -> 27.5 ms, 16 -> 23.1 ms, 32 -> 25.5 ms)
constexpr int kInfPerWarp = 32 / kInfGroup;       // streams a warp inflates side by side

// ---- CodecDeflate --------------------------------------------------------------------------------------
// One warp per CTA: streams differ a lot in length, and a CTA of eight warps keeps its slot until the longest is through
// (measured on CodecFloat: 24 % warps active with eight-warp CTAs).  A warp inflates 32 / kInfGroup streams side by side.
__global__ void __launch_bounds__(32) deflate_inflate_kernel(DecodeArgs a, uint8_t* region, size_t regionStride) {
  constexpr int kG = kInfGroup, kPerWarp = 32 / kG;  // streams per warp (g4_inflate.cuh)
  __shared__ InflateWarpShared S[kPerWarp];
  const int lane = threadIdx.x & 31, grp = lane / kG, sub = lane % kG;
  const int li = blockIdx.x * kPerWarp + grp;
  if (li >= *a.listCount) return;
  const int tIdx = a.list[li];
  const int n = a.band.tile_rows * a.band.tile_cols;
  const uint8_t* packing = a.arena + a.offsets[tIdx];
  const uint32_t len = a.lens[tIdx];
  int status = G4_OK;
  const int pred = len >= 10 ? int(packing[1]) : 0;
  const uint32_t nM32 = len >= 10 ? load_le32(packing + 6) : 0;
  const uint32_t expect = pred == G4_PRED_DIFF_NULLS ? uint32_t(n) : uint32_t(n - 1);
  if (len < 12 || pred < 1 || pred > 4 || nM32 < expect || nM32 > uint32_t(6 * n)) status = G4_ERR_FORMAT;
  else {
    uint32_t produced = 0, consumed = 0;
    int rc = inflate_group<kG>(S[grp], packing + 10, len - 10, region + size_t(li) * regionStride, nM32, &produced, &consumed);
    if (rc != kInfOk || produced != nM32) status = G4_ERR_FORMAT;  // DataFormatException -> IOException (:150-152)
  }
  if (sub == 0) a.status[tIdx] = status;
}

__global__ void __launch_bounds__(kThreads) deflate_finish_kernel(DecodeArgs a, const uint8_t* region, size_t regionStride) {
  __shared__ uint32_t scan[kWarps + 1];
  __shared__ int sTile;
  const int tid = threadIdx.x;
  for (;;) {
    __syncthreads();
    if (tid == 0) sTile = atomicAdd(a.counter, 1);
    __syncthreads();
    const int li = sTile;
    if (li >= *a.listCount) break;
    const int tIdx = a.list[li];
    if (a.status[tIdx] != G4_OK) continue;
    const TileView t = tile_view(a.band, a.grid, tIdx);
    const int n = t.R * t.C;
    const uint8_t* packing = a.arena + a.offsets[tIdx];
    const int pred = int(packing[1]);
    const int32_t seed = int32_t(load_le32(packing + 2));
    const uint32_t nM32 = load_le32(packing + 6);
    const uint8_t* m32 = region + size_t(li) * regionStride;
    int status = G4_OK;
    const uint32_t expect = pred == G4_PRED_DIFF_NULLS ? uint32_t(n) : uint32_t(n - 1);
    if (!m32_parse_to_cells(m32, nM32, pred, t, expect, scan)) status = G4_ERR_FORMAT;
    else if (pred == G4_PRED_DIFF_NULLS) predictor_inverse_nulls(t, seed);
    else {
      __syncthreads();
      if (tid == 0) t.at(0, 0) = seed;
      __syncthreads();
      predictor_inverse(pred, t, scan);
    }
    if (tid == 0 && status != G4_OK) a.status[tIdx] = status;
  }
}

// ---- CodecFloat ----------------------------------------------------------------------------------------
// packing = [codecIndex][0] + 5 x ([len:int32 LE][zlib stream]): sign bitmap, exponent bytes, three
// row-delta coded mantissa byte planes (CodecFloat.java:377-387).  Staging per tile: 5 planes of n bytes.
// Jobs are plane-major: the four streams of a warp are the SAME plane of four consecutive tiles, so that they have the same
// character (the exponent plane is all matches, the low mantissa plane all literals) and the four walking lanes stay in
// step; mixing the planes of one tile in a warp serialises them (each waits at the others' events) and was slower than one
// stream per warp.
__global__ void __launch_bounds__(32) float_inflate_kernel(DecodeArgs a, uint8_t* region, size_t regionStride, int nTilesUpper) {
  __shared__ InflateWarpShared S[kInfPerWarp];
  const int lane = threadIdx.x & 31, grp = lane / kInfGroup, sub = lane % kInfGroup;
  const int perPlane = (nTilesUpper + kInfPerWarp - 1) / kInfPerWarp;  // warps per plane
  const int plane = blockIdx.x / perPlane;
  const int li = (blockIdx.x - plane * perPlane) * kInfPerWarp + grp;
  if (li >= *a.listCount) return;
  const int tIdx = a.list[li];
  const uint32_t n = uint32_t(a.band.tile_rows) * uint32_t(a.band.tile_cols);
  const size_t planeStride = (size_t(n) + 15) & ~size_t(15);
  const uint8_t* packing = a.arena + a.offsets[tIdx];
  const uint32_t len = a.lens[tIdx];
  uint32_t off = 2;
  bool ok = len >= 2 + 5 * 4;
  uint32_t secLen = 0;
  for (int p = 0; ok && p <= plane; p++) {
    if (off + 4 > len) { ok = false; break; }
    secLen = load_le32(packing + off);
    off += 4;
    if (secLen > len - off) { ok = false; break; }
    if (p < plane) off += secLen;
  }
  int status = G4_OK;
  if (!ok) status = G4_ERR_FORMAT;
  else {
    const uint32_t want = plane == 0 ? (n + 7) / 8 : n;
    uint32_t produced = 0, consumed = 0;
    int rc = inflate_group<kInfGroup>(S[grp], packing + off, secLen, region + size_t(li) * regionStride + size_t(plane) * planeStride, want,
                                      &produced, &consumed);
    if (rc != kInfOk || produced != want) status = G4_ERR_FORMAT;  // "Inflate failed" (CodecFloat.java:285-298)
  }
  if (sub == 0 && status != G4_OK) atomicMin(&a.status[tIdx], status);
}

__global__ void __launch_bounds__(kThreads) float_finish_kernel(DecodeArgs a, const uint8_t* region, size_t regionStride) {
  __shared__ uint32_t scan[kWarps + 1];
  __shared__ int sTile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (;;) {
    __syncthreads();
    if (tid == 0) sTile = atomicAdd(a.counter, 1);
    __syncthreads();
    const int li = sTile;
    if (li >= *a.listCount) break;
    const int tIdx = a.list[li];
    if (a.status[tIdx] < 0) continue;  // an inflate warp reported an error
    const TileView t = tile_view(a.band, a.grid, tIdx);
    const int R = t.R, C = t.C;
    const uint32_t n = uint32_t(R) * uint32_t(C);
    const size_t planeStride = (size_t(n) + 15) & ~size_t(15);
    // the staging planes are mutable scratch: the three mantissa planes are delta-decoded in place
    uint8_t* base = const_cast<uint8_t*>(region) + size_t(li) * regionStride;
    const uint8_t* sign = base;
    const uint8_t* expo = base + planeStride;
    // CodecFloat.decodeDeltas (:315-325): the first byte of row r continues from the first byte of row r-1
    for (int p = 2; p < 5; p++) {
      uint8_t* pl = base + size_t(p) * planeStride;
      uint32_t carry = 0;
      for (int r0 = 0; r0 < R; r0 += kThreads) {
        int r = r0 + tid;
        uint32_t x = r < R ? pl[size_t(r) * C] : 0u;
        uint32_t tot;
        uint32_t ex = block_exclusive_scan(x, scan, &tot);
        if (r < R) pl[size_t(r) * C] = uint8_t(carry + ex + x);
        carry += tot;
      }
    }
    __syncthreads();
    // rows: byte prefix sums (mod 256) from the row's first byte; then merge the five planes into IEEE bits
    const int seg = (C + 31) / 32;
    for (int r = warp; r < R; r += kWarps) {
      const int c0 = lane * seg;
      uint32_t pre[3];
#pragma unroll
      for (int p = 0; p < 3; p++) {
        const uint8_t* pl = base + size_t(p + 2) * planeStride + size_t(r) * C;
        uint32_t loc = 0;
        for (int i = 0; i < seg; i++) {
          int c = c0 + i;
          if (c < C && c > 0) loc += pl[c];
        }
        uint32_t inc = warp_inclusive_scan(loc);
        pre[p] = inc - loc + pl[0];  // decoded value just before this lane's segment (row start included)
      }
      for (int i = 0; i < seg; i++) {
        int c = c0 + i;
        if (c >= C) break;
        uint32_t m[3];
#pragma unroll
        for (int p = 0; p < 3; p++) {
          const uint8_t* pl = base + size_t(p + 2) * planeStride + size_t(r) * C;
          if (c > 0) pre[p] += pl[c];
          m[p] = pre[p] & 0xffu;
        }
        uint32_t k = uint32_t(r) * C + c;
        uint32_t bits = (uint32_t((sign[k >> 3] >> (k & 7)) & 1u) << 31) | (uint32_t(expo[k]) << 23) | ((m[0] & 0x7fu) << 16) |
                        (m[1] << 8) | m[2];
        t.at(r, c) = int32_t(bits);
      }
    }
    if (tid == 0) a.status[tIdx] = G4_OK;
  }
}

cudaError_t launch_deflate_decode(const DecodeArgs& a, uint8_t* region, size_t regionStride, int nCtas, int nTilesUpper,
                                  cudaStream_t s) {
  deflate_inflate_kernel<<<(nTilesUpper + kInfPerWarp - 1) / kInfPerWarp, 32, 0, s>>>(a, region, regionStride);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  deflate_finish_kernel<<<nCtas, kThreads, 0, s>>>(a, region, regionStride);
  return cudaGetLastError();
}

cudaError_t launch_float_decode(const DecodeArgs& a, uint8_t* region, size_t regionStride, int nCtas, int nTilesUpper,
                                cudaStream_t s) {
  float_inflate_kernel<<<5 * ((nTilesUpper + kInfPerWarp - 1) / kInfPerWarp), 32, 0, s>>>(a, region, regionStride, nTilesUpper);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  float_finish_kernel<<<nCtas, kThreads, 0, s>>>(a, region, regionStride);
  return cudaGetLastError();
}

}  // namespace g4
