// g4_batch.cu -- the selection layer and data-movement kernels around the codec kernels.
//
// Reference semantics reproduced here (paths under /root/reference/core/src/main/java/org/gridfour/):
//   CodecMaster.encodeSingleThread  gvrs/CodecMaster.java:150-169   strictly smallest, ties -> lowest index
//   TileElementInt.encode/decode    gvrs/TileElementInt.java:196-219  raw little-endian when len >= 4n
//   CodecMaster.decode              gvrs/CodecMaster.java:195-203   dispatch on packing[0]
#include "g4_kernels.h"
#include "g4_device.cuh"
#include "../../include/g4terrain.h"

namespace g4 {

__global__ void fill_terrain_kernel(int elemType, uint64_t seed, int64_t row0, int64_t col0, int64_t nRows, int64_t nCols,
                                    void* out) {
  const int64_t total = nRows * nCols;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    int64_t r = i / nCols, c = i - r * nCols;
    if (elemType == G4_ELEM_I32) static_cast<int32_t*>(out)[i] = g4_terrain_m(seed, row0 + r, col0 + c);
    else static_cast<float*>(out)[i] = float(g4_terrain_dm(seed, row0 + r, col0 + c)) * 0.1f;
  }
}

// Best-of selection.  cand arrays are [nCand][nTiles]; candIndex[c] = codec list position of candidate c
// (ascending, so the first strictly-smallest candidate is the lowest index).
__global__ void select_kernel(SelectArgs a) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.nTiles) return;
  uint32_t best = 0xffffffffu;
  int bc = -1;
  int st = G4_OK;
  for (int c = 0; c < a.nCand; c++) {
    int s = a.candStatus[size_t(c) * a.nTiles + t];
    uint32_t l = a.candLens[size_t(c) * a.nTiles + t];
    if (s == G4_OK && l > 0 && l < best) { best = l; bc = c; }
    else if (s < 0 && s != G4_ERR_CAPACITY && st == G4_OK) st = s;  // a candidate failed (not "declined")
  }
  const uint32_t rawLen = a.rawLen;
  if (bc < 0 || best >= rawLen) {  // TileElementInt.java:198-204
    a.lens[t] = rawLen;
    a.codecOut[t] = G4_CODEC_RAW;
    a.predOut[t] = 0;
    a.src[t] = -1;
  } else {
    a.lens[t] = best;
    a.codecOut[t] = uint8_t(a.candIndex[bc]);
    a.predOut[t] = a.candPreds[size_t(bc) * a.nTiles + t];
    a.src[t] = bc;
  }
  a.status[t] = st;
}

// Exclusive scan of the 8-byte-aligned payload lengths -> offsets; one CTA.
__global__ void __launch_bounds__(kThreads) offsets_kernel(const uint32_t* lens, uint64_t* offsets, int nTiles, uint64_t* total) {
  __shared__ unsigned long long sm[kWarps];
  __shared__ unsigned long long carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int t0 = 0; t0 < nTiles; t0 += kThreads) {
    int t = t0 + threadIdx.x;
    unsigned long long x = t < nTiles ? ((unsigned long long)(lens[t]) + 7ull) & ~7ull : 0ull;
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
    if (t < nTiles) offsets[t] = base + inc - x;
    __syncthreads();
    if (threadIdx.x == kThreads - 1) carry = base + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

// One CTA per tile: copy the winner's packing (or the raw samples) to arena + offsets[t].
// Destination offsets are multiples of 8; the tail of the last 8-byte word is zero filled.
__global__ void __launch_bounds__(kThreads) compact_kernel(CompactArgs a) {
  const int t = blockIdx.x;
  const uint32_t len = a.lens[t];
  const uint64_t off = a.offsets[t];
  if (off + ((uint64_t(len) + 7) & ~7ull) > a.arenaCap) {
    if (threadIdx.x == 0) a.status[t] = G4_ERR_CAPACITY;
    return;
  }
  uint8_t* dst = a.arena + off;
  const int src = a.src[t];
  if (src < 0 && a.raw16) {  // TileElementShort raw form: little-endian shorts, zero padded to a multiple of 4 bytes
    const int R = a.band.tile_rows, C = a.band.tile_cols, n = R * C;
    const int tr = t / a.band.tiles_across, tc = t - tr * a.band.tiles_across;
    const int16_t* base = a.raw16 + int64_t(tr) * R * a.raw16Pitch + int64_t(tc) * C;
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
    const int nWords = int((len + 3u) >> 2);
    for (int i = threadIdx.x; i < ((nWords + 1) & ~1); i += kThreads) {
      uint32_t w = 0;
      for (int h = 0; h < 2; h++) {
        int k = 2 * i + h;
        if (k < n) { int r = k / C, c = k - r * C; w |= uint32_t(uint16_t(base[int64_t(r) * a.raw16Pitch + c])) << (16 * h); }
      }
      d32[i] = w;
    }
  } else if (src < 0) {
    const TileView tv = tile_view(a.band, a.grid, t);
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
    const int n = tv.R * tv.C;
    for (int i = threadIdx.x; i < n; i += kThreads) {
      int r = i / tv.C, c = i - r * tv.C;
      d32[i] = uint32_t(tv.at(r, c));
    }
    if ((n & 1) && threadIdx.x == 0) d32[n] = 0;
  } else {
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(a.slots[src] + size_t(t) * a.slotBytes);
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
    const uint32_t nWords = ((len + 7u) & ~7u) >> 2;
    for (uint32_t i = threadIdx.x; i < nWords; i += kThreads) {
      uint32_t w = 0;
      if (i * 4u < len) {
        w = s32[i];
        uint32_t valid = len - i * 4u;
        if (valid < 4u) w &= (1u << (8u * valid)) - 1u;
      }
      d32[i] = w;
    }
  }
}

// Decode-side classification: one thread per tile appends the tile to the list of its codec kind.
// lists layout: [G4_CODEC_COUNT + 1][nTiles]; kind G4_CODEC_COUNT == raw.
__global__ void classify_kernel(ClassifyArgs a) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.nTiles) return;
  const uint32_t len = a.lens[t];
  int kind;
  int st = G4_OK;
  // a payload must lie inside the arena; one longer than the raw tile is never written (TileElementInt.encode :196-205
  // stores the raw samples instead) and would overflow the decoders' per-tile bounds
  if (len > a.rawLen || a.offsets[t] > a.arenaLen || uint64_t(len) > a.arenaLen - a.offsets[t]) { kind = -1; st = G4_ERR_FORMAT; }
  else if (len == a.rawLen) kind = G4_CODEC_COUNT;
  else if (len == 0) { kind = -1; st = G4_ERR_FORMAT; }
  else {
    int index = a.arena[a.offsets[t]];
    if (index >= a.codecs.n_codecs) { kind = -1; st = G4_ERR_FORMAT; }  // CodecMaster.java:197-199
    else {
      kind = a.codecs.codec_ids[index];
      bool isFloatCodec = kind == G4_CODEC_FLOAT;
      if (kind < 0 || kind >= G4_CODEC_COUNT || isFloatCodec != (a.elemType == G4_ELEM_F32)) { kind = -1; st = G4_ERR_FORMAT; }
    }
  }
  a.status[t] = st;
  if (kind >= 0) {
    int pos = atomicAdd(&a.counts[kind], 1);
    a.lists[size_t(kind) * a.nTiles + pos] = t;
  }
}

__global__ void __launch_bounds__(kThreads) raw_decode_kernel(DecodeArgs a) {
  const int li = blockIdx.x;
  if (li >= *a.listCount) return;
  const int t = a.list[li];
  const TileView tv = tile_view(a.band, a.grid, t);
  const uint8_t* src = a.arena + a.offsets[t];
  const int n = tv.R * tv.C;
  if (a.rawShorts) {
    for (int i = threadIdx.x; i < n; i += kThreads) {
      int r = i / tv.C, c = i - r * tv.C;
      tv.at(r, c) = int32_t(int16_t(uint16_t(src[2 * size_t(i)]) | (uint16_t(src[2 * size_t(i) + 1]) << 8)));
    }
    return;
  }
  const bool aligned = (reinterpret_cast<uintptr_t>(src) & 3) == 0;
  for (int i = threadIdx.x; i < n; i += kThreads) {
    int r = i / tv.C, c = i - r * tv.C;
    uint32_t v = aligned ? reinterpret_cast<const uint32_t*>(src)[i] : load_le32(src + 4 * size_t(i));
    tv.at(r, c) = int32_t(v);
  }
}

// TileElementShort.encode (:213-220): widen, fill value -> INT4_NULL_CODE
__global__ void widen_i16_kernel(const int16_t* src, int64_t srcPitch, int32_t* dst, int64_t rows, int64_t cols, int32_t fill) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    int64_t r = i / cols, c = i - r * cols;
    int32_t v = src[r * srcPitch + c];
    dst[i] = v == fill ? kNull : v;
  }
}
// TileElementShort.decode (:236-245): INT4_NULL_CODE -> SHORT_NULL_CODE, everything else narrowed with a (short) cast
__global__ void narrow_i16_kernel(const int32_t* src, int16_t* dst, int64_t dstPitch, int64_t rows, int64_t cols) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    int64_t r = i / cols, c = i - r * cols;
    int32_t v = src[i];
    dst[r * dstPitch + c] = v == kNull ? int16_t(-32768) : int16_t(v);
  }
}

cudaError_t launch_fill_terrain(int elemType, uint64_t seed, int64_t row0, int64_t col0, int64_t nRows, int64_t nCols, void* out,
                                cudaStream_t s) {
  fill_terrain_kernel<<<148 * 8, 256, 0, s>>>(elemType, seed, row0, col0, nRows, nCols, out);
  return cudaGetLastError();
}
cudaError_t launch_widen_i16(const int16_t* src, int64_t srcPitch, int32_t* dst, int64_t rows, int64_t cols, int32_t fill, cudaStream_t s) {
  widen_i16_kernel<<<148 * 8, 256, 0, s>>>(src, srcPitch, dst, rows, cols, fill);
  return cudaGetLastError();
}
cudaError_t launch_narrow_i16(const int32_t* src, int16_t* dst, int64_t dstPitch, int64_t rows, int64_t cols, cudaStream_t s) {
  narrow_i16_kernel<<<148 * 8, 256, 0, s>>>(src, dst, dstPitch, rows, cols);
  return cudaGetLastError();
}
cudaError_t launch_select(const SelectArgs& a, cudaStream_t s) {
  select_kernel<<<(a.nTiles + 255) / 256, 256, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_offsets(const uint32_t* lens, uint64_t* offsets, int nTiles, uint64_t* total, cudaStream_t s) {
  offsets_kernel<<<1, kThreads, 0, s>>>(lens, offsets, nTiles, total);
  return cudaGetLastError();
}
cudaError_t launch_compact(const CompactArgs& a, int nTiles, cudaStream_t s) {
  compact_kernel<<<nTiles, kThreads, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_classify(const ClassifyArgs& a, cudaStream_t s) {
  classify_kernel<<<(a.nTiles + 255) / 256, 256, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_raw_decode(const DecodeArgs& a, int nTiles, cudaStream_t s) {
  raw_decode_kernel<<<nTiles, kThreads, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace g4
