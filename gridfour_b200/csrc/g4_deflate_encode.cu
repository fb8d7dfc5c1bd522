// g4_deflate_encode.cu -- encode side of CodecDeflate and CodecFloat, and the generic zlib-stream worker kernel.
//
// Reference (under /root/reference/core/src/main/java/org/gridfour/compress/):
//   CodecDeflate.java:157-202 (encode: best of the three predictors by packing length, strict <),
//   CodecDeflate.java:204-228 (compress: Deflater(6), finish(), one deflate() into nM32+128 bytes at offset 10),
//   CodecFloat.java:300-313 (encodeDeltas), :328-392 (encodeFloats), :268-283 (doDeflate: Deflater(9), n+128 bytes)
//
// Every codec that needs zlib streams runs the same four stages:
//   size   one CTA per tile computes the byte length of each stream it will hand to DEFLATE
//   scan   exclusive scan of the padded lengths -> stream offsets (exact-size staging, no 6n worst-case slabs)
//   write  one CTA per tile writes the streams (M32 bytes in predictor stream order / float byte planes)
//   zlib   one THREAD per stream replays zlib's deflate_slow (g4_deflate_enc.cuh); LZ77 hash-chain matching is
//          serial per stream, so the parallelism is across the tens of thousands of streams of a band
//   pick   one CTA per tile selects / assembles the packing in the tile's candidate slot
#include "g4_kernels.h"
#include "g4_device.cuh"
#include "g4_predict.cuh"
#include "g4_m32stream.cuh"
#include "g4_deflate_enc.cuh"

namespace g4 {

size_t deflate_work_bytes() { return sizeof(DeflateWork); }

// ---- stream offsets -----------------------------------------------------------------------------------------
// inOff[j] = sum_{i<j} align16(inLen[i] + 16); the output slot of stream j starts at inOff[j] + 112*j in the output
// buffer (capacity inLen[j] + 128).  One CTA.
__global__ void __launch_bounds__(kThreads) stream_offsets_kernel(const uint32_t* inLen, uint64_t* inOff, int n, uint64_t* total) {
  __shared__ unsigned long long sm[kWarps];
  __shared__ unsigned long long carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int t0 = 0; t0 < n; t0 += kThreads) {
    int t = t0 + threadIdx.x;
    unsigned long long x = t < n ? ((unsigned long long)(inLen[t]) + 31ull) & ~15ull : 0ull;
    unsigned long long inc = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      unsigned long long y = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += y;
    }
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    unsigned long long base = carry;
    for (int w = 0; w < warp; w++) base += sm[w];
    if (t < n) inOff[t] = base + inc - x;
    __syncthreads();
    if (threadIdx.x == kThreads - 1) carry = base + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

// ---- one thread per zlib stream ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) deflate_streams_kernel(StreamArgs a) {
  DeflateWork* W = static_cast<DeflateWork*>(a.work) + (size_t(blockIdx.x) * blockDim.x + threadIdx.x);
  for (;;) {
    const int j = atomicAdd(a.counter, 1);
    if (j >= a.nStreams) break;
    const uint32_t n = a.inLen[j];
    if (n == 0) { a.outLen[j] = 0; continue; }  // stream not produced (tile declined by the size stage)
    const uint64_t off = a.inOff[j];
    a.outLen[j] = deflate_stream(a.inBuf + off, n, a.outBuf + off + 112ull * uint64_t(j), n + uint32_t(a.capExtra), W, a.level);
  }
}

// ---- CodecDeflate ---------------------------------------------------------------------------------------------
namespace {
struct PredResidualGet {
  TileView t;
  int pred;
  __device__ __forceinline__ int32_t operator()(uint32_t k) const {
    int r, c;
    stream_to_cell(pred, int(k), t.R, t.C, &r, &c);
    return residual_at(pred, t, r, c);
  }
};
struct NullsResidualGet {  // PredictorModelDifferencingWithNulls: one residual per cell, row-major
  TileView t;
  int32_t seed;
  __device__ __forceinline__ int32_t operator()(uint32_t k) const {
    int r = int(k) / t.C, c = int(k) - r * t.C;
    return residual_nulls_at(t, r, c, seed);
  }
};
}  // namespace

__global__ void __launch_bounds__(kThreads) deflate_m32_size_kernel(EncodeArgs a, uint32_t* inLen) {
  __shared__ uint32_t scan[kWarps + 1];
  const int tid = threadIdx.x;
  const int nTiles = a.band.tiles_down * a.band.tiles_across;
  for (int tIdx = blockIdx.x; tIdx < nTiles; tIdx += gridDim.x) {
    const TileView t = tile_view(a.band, a.grid, tIdx);
    const int n = t.R * t.C;
    bool sawNull = false, sawValid = false;
    for (int i = tid; i < n; i += kThreads) {
      int r = i / t.C, c = i - r * t.C;
      if (t.at(r, c) == kNull) sawNull = true; else sawValid = true;
    }
    const bool anyNull = __syncthreads_or(sawNull) != 0;
    const bool anyValid = __syncthreads_or(sawValid) != 0;
    if (!anyValid) {  // all-null tile -> null (:167-169)
      if (tid == 0) {
        a.lens[tIdx] = 0; a.preds[tIdx] = 0; a.status[tIdx] = G4_DECLINED;
        inLen[3 * tIdx] = inLen[3 * tIdx + 1] = inLen[3 * tIdx + 2] = 0;
      }
      continue;
    }
    if (anyNull) {  // only PredictorModelDifferencingWithNulls applies (:178-186): one stream, a.preds marks the case
      int nStart;
      const int32_t seed = nulls_seed(t, &nStart);
      uint32_t sz = m32_stream_size(NullsResidualGet{t, seed}, uint32_t(n), scan);
      if (tid == 0) {
        inLen[3 * tIdx] = sz; inLen[3 * tIdx + 1] = inLen[3 * tIdx + 2] = 0;
        a.preds[tIdx] = G4_PRED_DIFF_NULLS; a.status[tIdx] = G4_OK;
      }
      continue;
    }
    if (tid == 0) a.preds[tIdx] = 0;
    for (int p = 0; p < 3; p++) {
      PredResidualGet get{t, p + 1};
      uint32_t sz = m32_stream_size(get, uint32_t(n - 1), scan);
      if (tid == 0) inLen[3 * tIdx + p] = sz;
    }
    if (tid == 0) a.status[tIdx] = G4_OK;
  }
}

__global__ void __launch_bounds__(kThreads) deflate_m32_write_kernel(EncodeArgs a, const uint32_t* inLen, const uint64_t* inOff,
                                                                     uint8_t* inBuf) {
  __shared__ uint32_t scan[kWarps + 1];
  const int nTiles = a.band.tiles_down * a.band.tiles_across;
  for (int tIdx = blockIdx.x; tIdx < nTiles; tIdx += gridDim.x) {
    if (inLen[3 * tIdx] == 0) continue;
    const TileView t = tile_view(a.band, a.grid, tIdx);
    const int n = t.R * t.C;
    if (a.preds[tIdx] == G4_PRED_DIFF_NULLS) {
      int nStart;
      const int32_t seed = nulls_seed(t, &nStart);
      uint8_t* dst = inBuf + inOff[3 * tIdx];
      uint32_t w = m32_stream_write(NullsResidualGet{t, seed}, uint32_t(n), dst, scan);
      if (threadIdx.x < 16) dst[w + threadIdx.x] = 0;
      continue;
    }
    for (int p = 0; p < 3; p++) {
      PredResidualGet get{t, p + 1};
      uint8_t* dst = inBuf + inOff[3 * tIdx + p];
      uint32_t w = m32_stream_write(get, uint32_t(n - 1), dst, scan);
      if (threadIdx.x < 16) dst[w + threadIdx.x] = 0;  // the pre-filter of longest_match may look 2 bytes past the end
    }
  }
}

__global__ void __launch_bounds__(kThreads) deflate_pick_kernel(EncodeArgs a, const uint32_t* inLen, const uint64_t* inOff,
                                                                const uint8_t* outBuf, const uint32_t* outLen) {
  const int tid = threadIdx.x;
  const int nTiles = a.band.tiles_down * a.band.tiles_across;
  for (int tIdx = blockIdx.x; tIdx < nTiles; tIdx += gridDim.x) {
    if (inLen[3 * tIdx] == 0) continue;
    // CodecDeflate.encode: Differencing, Linear, Triangle; keep the strictly smallest packing (:176-199)
    uint32_t best = 0xffffffffu;
    int win = -1;
    for (int p = 0; p < 3; p++) {
      uint32_t dN = outLen[3 * tIdx + p];
      if (dN > 0 && dN < best) { best = dN; win = p; }  // dN <= 0: "deflate failed" -> candidate skipped (:211-214)
    }
    if (win < 0) {
      if (tid == 0) { a.lens[tIdx] = 0; a.preds[tIdx] = 0; a.status[tIdx] = G4_DECLINED; }
      continue;
    }
    const int j = 3 * tIdx + win;
    const uint32_t len = best + 10u;
    const bool fits = len <= a.slotBytes;
    const bool nulls = a.preds[tIdx] == G4_PRED_DIFF_NULLS;
    const TileView t = tile_view(a.band, a.grid, tIdx);
    int nStart = 0;
    const int32_t seedNulls = nulls ? nulls_seed(t, &nStart) : 0;
    const int predCode = nulls ? G4_PRED_DIFF_NULLS : win + 1;
    if (fits) {
      uint8_t* slot = a.slots + size_t(tIdx) * a.slotBytes;
      const uint8_t* src = outBuf + inOff[j] + 112ull * uint64_t(j);
      if (tid == 0) {
        const uint32_t seed = nulls ? uint32_t(seedNulls) : uint32_t(t.at(0, 0)), nM32 = inLen[j];
        slot[0] = uint8_t(a.codecIndex);
        slot[1] = uint8_t(predCode);
        for (int k = 0; k < 4; k++) { slot[2 + k] = uint8_t(seed >> (8 * k)); slot[6 + k] = uint8_t(nM32 >> (8 * k)); }
      }
      for (uint32_t i = tid; i < best; i += kThreads) slot[10 + i] = src[i];
    }
    __syncthreads();
    if (tid == 0) { a.lens[tIdx] = len; a.preds[tIdx] = uint8_t(predCode); a.status[tIdx] = fits ? G4_OK : G4_ERR_CAPACITY; }
  }
}

// ---- CodecFloat -------------------------------------------------------------------------------------------------
// Streams of tile t: 5t+0 sign bitmap ((n+7)/8 bytes, cell i -> bit i LSB-first), 5t+1 exponent bytes, 5t+2..4 the
// mantissa bytes (high 7 bits, middle, low) as row-wise byte differences; the first byte of a row differences
// against the first byte of the previous row (row 0 against 0).
__global__ void float_plane_size_kernel(int nTiles, uint32_t n, uint32_t* inLen) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= 5 * nTiles) return;
  inLen[j] = (j % 5 == 0) ? (n + 7u) / 8u : n;
}

__global__ void __launch_bounds__(kThreads) float_plane_write_kernel(EncodeArgs a, const uint64_t* inOff, uint8_t* inBuf) {
  const int tid = threadIdx.x;
  const int nTiles = a.band.tiles_down * a.band.tiles_across;
  for (int tIdx = blockIdx.x; tIdx < nTiles; tIdx += gridDim.x) {
    const TileView t = tile_view(a.band, a.grid, tIdx);
    const int R = t.R, C = t.C, n = R * C;
    uint8_t* pl[5];
    for (int p = 0; p < 5; p++) pl[p] = inBuf + inOff[5 * tIdx + p];
    const int nSignBytes = (n + 7) / 8;
    for (int b = tid; b < nSignBytes; b += kThreads) {
      uint32_t v = 0;
      for (int i = 0; i < 8; i++) {
        int k = 8 * b + i;
        if (k < n) { int r = k / C, c = k - r * C; v |= (uint32_t(t.at(r, c)) >> 31) << i; }
      }
      pl[0][b] = uint8_t(v);
    }
    for (int k = tid; k < n; k += kThreads) {
      int r = k / C, c = k - r * C;
      uint32_t bits = uint32_t(t.at(r, c));
      uint32_t prior = c > 0 ? uint32_t(t.at(r, c - 1)) : r > 0 ? uint32_t(t.at(r - 1, 0)) : 0u;
      pl[1][k] = uint8_t(bits >> 23);
      pl[2][k] = uint8_t(((bits >> 16) & 0x7fu) - ((prior >> 16) & 0x7fu));
      pl[3][k] = uint8_t((bits >> 8) - (prior >> 8));
      pl[4][k] = uint8_t(bits - prior);
    }
    if (tid < 16) {
      pl[0][nSignBytes + tid] = 0;
      for (int p = 1; p < 5; p++) pl[p][n + tid] = 0;
    }
  }
}

__global__ void __launch_bounds__(kThreads) float_pick_kernel(EncodeArgs a, const uint64_t* inOff, const uint8_t* outBuf,
                                                              const uint32_t* outLen) {
  const int tid = threadIdx.x;
  const int nTiles = a.band.tiles_down * a.band.tiles_across;
  for (int tIdx = blockIdx.x; tIdx < nTiles; tIdx += gridDim.x) {
    uint32_t dN[5], total = 2;
    bool failed = false;
    for (int p = 0; p < 5; p++) {
      dN[p] = outLen[5 * tIdx + p];
      if (dN[p] == 0) failed = true;  // doDeflate throws "Deflate failed"
      total += 4u + dN[p];
    }
    const bool fits = total <= a.slotBytes;
    if (!failed && fits) {
      uint8_t* slot = a.slots + size_t(tIdx) * a.slotBytes;
      if (tid == 0) { slot[0] = uint8_t(a.codecIndex); slot[1] = 0; }
      uint32_t off = 2;
      for (int p = 0; p < 5; p++) {
        const int j = 5 * tIdx + p;
        const uint8_t* src = outBuf + inOff[j] + 112ull * uint64_t(j);
        if (tid < 4) slot[off + tid] = uint8_t(dN[p] >> (8 * tid));
        off += 4;
        for (uint32_t i = tid; i < dN[p]; i += kThreads) slot[off + i] = src[i];
        off += dN[p];
      }
    }
    if (tid == 0) {
      a.lens[tIdx] = failed ? 0u : total;
      a.preds[tIdx] = 0;
      a.status[tIdx] = failed ? G4_DECLINED : fits ? G4_OK : G4_ERR_CAPACITY;
    }
  }
}

// ---- launchers ------------------------------------------------------------------------------------------------------
cudaError_t launch_stream_offsets(const uint32_t* inLen, uint64_t* inOff, int nStreams, uint64_t* total, cudaStream_t s) {
  stream_offsets_kernel<<<1, kThreads, 0, s>>>(inLen, inOff, nStreams, total);
  return cudaGetLastError();
}
cudaError_t launch_deflate_streams(const StreamArgs& a, int nWorkers, cudaStream_t s) {
  deflate_streams_kernel<<<(nWorkers + 31) / 32, 32, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_deflate_m32_size(const EncodeArgs& a, uint32_t* inLen, int nCtas, cudaStream_t s) {
  deflate_m32_size_kernel<<<nCtas, kThreads, 0, s>>>(a, inLen);
  return cudaGetLastError();
}
cudaError_t launch_deflate_m32_write(const EncodeArgs& a, const uint32_t* inLen, const uint64_t* inOff, uint8_t* inBuf, int nCtas,
                                     cudaStream_t s) {
  deflate_m32_write_kernel<<<nCtas, kThreads, 0, s>>>(a, inLen, inOff, inBuf);
  return cudaGetLastError();
}
cudaError_t launch_deflate_pick(const EncodeArgs& a, const uint32_t* inLen, const uint64_t* inOff, const uint8_t* outBuf,
                                const uint32_t* outLen, int nCtas, cudaStream_t s) {
  deflate_pick_kernel<<<nCtas, kThreads, 0, s>>>(a, inLen, inOff, outBuf, outLen);
  return cudaGetLastError();
}
cudaError_t launch_float_plane_size(int nTiles, uint32_t n, uint32_t* inLen, cudaStream_t s) {
  float_plane_size_kernel<<<(5 * nTiles + 255) / 256, 256, 0, s>>>(nTiles, n, inLen);
  return cudaGetLastError();
}
cudaError_t launch_float_plane_write(const EncodeArgs& a, const uint64_t* inOff, uint8_t* inBuf, int nCtas, cudaStream_t s) {
  float_plane_write_kernel<<<nCtas, kThreads, 0, s>>>(a, inOff, inBuf);
  return cudaGetLastError();
}
cudaError_t launch_float_pick(const EncodeArgs& a, const uint64_t* inOff, const uint8_t* outBuf, const uint32_t* outLen, int nCtas,
                              cudaStream_t s) {
  float_pick_kernel<<<nCtas, kThreads, 0, s>>>(a, inOff, outBuf, outLen);
  return cudaGetLastError();
}

}  // namespace g4
