// g4_canon.cu -- CodecCanonHuffman (canonical Huffman over integer predictor residuals) on sm_100a.
//
// Reference: compress/canonicalHuffman/CodecCanonHuffman.java:79-195 (+ the classes cited in g4_canon.cuh),
// under /root/reference/core/src/main/java/org/gridfour/.
#include "g4_kernels.h"
#include "g4_predict.cuh"
#include "g4_canon.cuh"
#include "g4_canon_fast.cuh"
#include "g4_canon_enc.cuh"

namespace g4 {

namespace {
struct PredCellSink {
  TileView t;
  int pred;
  __device__ __forceinline__ void operator()(uint32_t k, int32_t v) const {
    int r, c;
    stream_to_cell(pred, int(k), t.R, t.C, &r, &c);
    t.at(r, c) = v;
  }
};
}  // namespace

// Decode: the packing is staged in shared memory and decoded with g4_canon_fast.cuh when it fits (<= 28 KB, 4-byte
// aligned); otherwise the global-memory decoder of g4_canon.cuh.
union CanonDecodeSmem {
  CanonFastShared f;
  CanonDecShared s;
};

__global__ void __launch_bounds__(kThreads, 4) canon_decode_kernel(DecodeArgs a) {
  extern __shared__ __align__(16) unsigned char canonDecSmem[];
  CanonFastShared& F = *reinterpret_cast<CanonFastShared*>(canonDecSmem);
  CanonDecShared& S = *reinterpret_cast<CanonDecShared*>(canonDecSmem);
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
    const TileView t = tile_view(a.band, a.grid, tIdx);
    const int n = t.R * t.C;
    const uint8_t* packing = a.arena + a.offsets[tIdx];
    const uint32_t len = a.lens[tIdx];
    int status = G4_OK;
    // CodecCanonHuffman.decode (:162-195)
    const int pred = len >= 6 ? int(int8_t(packing[1])) : -1;
    const int32_t seed = len >= 6 ? int32_t(load_le32(packing + 2)) : 0;
    if (len < 6) status = G4_ERR_FORMAT;
    else if (pred == 0 && len == 6) {  // uniform tile shortcut (:170-175)
      for (int i = tid; i < n; i += kThreads) {
        int r = i / t.C, c = i - r * t.C;
        t.at(r, c) = seed;
      }
    } else if (pred < 1 || pred > 4) status = G4_ERR_FORMAT;
    else {
      uint32_t endBit = 0, nv = 0;
      const uint32_t expect = pred == G4_PRED_DIFF_NULLS ? uint32_t(n) : uint32_t(n - 1);
      bool ok;
      if (len <= uint32_t(kFastStageWords) * 4u && (reinterpret_cast<uintptr_t>(packing) & 3) == 0) {
        canon_fast_stage(F, packing, len);
        CellRunSink sink{t, pred, 0};
        ok = canon_fast_decode_stream(F, len * 8u, 48u, expect, 0u, sink, &endBit, &nv);
      } else {
        BitSrc src;
        src.init(packing + 6, len - 6);
        PredCellSink sink{t, pred};
        ok = canon_decode_stream(S, src, 0, expect, 0u, sink, &endBit, &nv);
      }
      if (!ok || nv != expect) status = G4_ERR_FORMAT;
      else if (pred == G4_PRED_DIFF_NULLS) predictor_inverse_nulls(t, seed);
      else {
        __syncthreads();
        if (tid == 0) t.at(0, 0) = seed;
        __syncthreads();
        predictor_inverse(pred, t, scan);
      }
    }
    if (tid == 0) a.status[tIdx] = status;
  }
}

// ---- encode ---------------------------------------------------------------------------------------------
namespace {
struct CanonEncKernelShared {
  CanonEncShared E;
  BitWindow W;
};
struct PredResidualGet {
  TileView t;
  int pred;
  int32_t seed;  // PredictorModelDifferencingWithNulls only
  __device__ __forceinline__ int32_t operator()(uint32_t k) const {
    int r, c;
    stream_to_cell(pred, int(k), t.R, t.C, &r, &c);
    return pred == G4_PRED_DIFF_NULLS ? residual_nulls_at(t, r, c, seed) : residual_at(pred, t, r, c);
  }
};
}  // namespace

// CodecCanonHuffman.encode (:79-142) + compress (:144-159)
__global__ void __launch_bounds__(kThreads) canon_encode_kernel(EncodeArgs a) {
  __shared__ CanonEncKernelShared S;
  __shared__ int sTile;
  const int tid = threadIdx.x;
  const int nTiles = a.band.tiles_down * a.band.tiles_across;
  uint8_t* pm = a.scratch + size_t(blockIdx.x) * a.scratchStride;
  for (;;) {
    __syncthreads();
    if (tid == 0) sTile = atomicAdd(a.counter, 1);
    __syncthreads();
    const int tIdx = sTile;
    if (tIdx >= nTiles) break;
    const TileView t = tile_view(a.band, a.grid, tIdx);
    const int n = t.R * t.C;
    uint32_t* outWords = reinterpret_cast<uint32_t*>(a.slots + size_t(tIdx) * a.slotBytes);
    const uint32_t capWords = uint32_t(a.slotBytes / 4);
    // nulls / uniformity scan (:83-110)
    bool sawNull = false, sawValid = false, differs = false;
    const int32_t v0 = t.at(0, 0);
    for (int i = tid; i < n; i += kThreads) {
      int r = i / t.C, c = i - r * t.C;
      int32_t v = t.at(r, c);
      if (v == kNull) sawNull = true; else sawValid = true;
      if (v != v0) differs = true;
    }
    const bool anyNull = __syncthreads_or(sawNull) != 0;
    const bool anyValid = __syncthreads_or(sawValid) != 0;
    const bool anyDiff = __syncthreads_or(differs) != 0;
    if (!anyValid) {
      if (tid == 0) { a.lens[tIdx] = 0; a.preds[tIdx] = 0; a.status[tIdx] = G4_DECLINED; }
      continue;
    }
    if (!anyDiff) {  // uniform tile: 6-byte packing, predictor 0
      if (tid == 0) {
        outWords[0] = (uint32_t(a.codecIndex) & 0xffu) | ((uint32_t(v0) & 0xffffu) << 16);
        outWords[1] = uint32_t(v0) >> 16;
        a.lens[tIdx] = 6; a.preds[tIdx] = 0; a.status[tIdx] = G4_OK;
      }
      continue;
    }
    // nulls: only PredictorModelDifferencingWithNulls applies (:117-125), one residual per cell
    int nStart = 0;
    const int32_t seedNulls = anyNull ? nulls_seed(t, &nStart) : 0;
    const uint32_t nRes = anyNull ? uint32_t(n) : uint32_t(n - 1);
    const int firstPred = anyNull ? 3 : 0, lastPred = anyNull ? 4 : 3;
    unsigned long long best = ~0ull;
    int win = -1, built = -1;
    bool bug = false;
    for (int p = firstPred; p < lastPred; p++) {
      PredResidualGet get{t, p + 1, seedNulls};
      const bool clean = canon_histogram(S.E, get, nRes);
      canon_build_code(S.E, pm);
      if (!clean && !canon_reconcile_escapes(S.E, get, nRes)) { bug = true; break; }
      built = p;
      unsigned long long bytes = 6ull + (S.E.totalBits + 7ull) / 8ull;
      if (bytes < best) { best = bytes; win = p; }
    }
    if (bug) {  // reference escape-range inconsistency with no code for the written symbol (see g4_canon_enc.cuh): decline
      if (tid == 0) { a.lens[tIdx] = 0; a.preds[tIdx] = 0; a.status[tIdx] = G4_DECLINED; }
      continue;
    }
    PredResidualGet get{t, win + 1, seedNulls};
    if (built != win) {
      canon_histogram(S.E, get, nRes);
      canon_build_code(S.E, pm);
    }
    BitOut o;
    bitwin_reset(S.W, o, outWords, capWords);
    if (tid == 0) {
      WinSink sink{S.W.win, 0};
      sink.put(uint32_t(a.codecIndex) & 0xffu, 8);
      sink.put(uint32_t(win + 1), 8);
      sink.put(anyNull ? uint32_t(seedNulls) : uint32_t(v0), 32);
    }
    o.bitPos = 48;
    __syncthreads();
    canon_emit_stream(S.E, S.W, o, get, nRes);
    bitwin_finish(S.W, o);
    if (tid == 0) {
      uint32_t len = (o.bitPos + 7) >> 3;
      a.lens[tIdx] = len;
      a.preds[tIdx] = uint8_t(win + 1);
      a.status[tIdx] = len <= a.slotBytes ? G4_OK : G4_ERR_CAPACITY;
    }
  }
}

cudaError_t launch_canon_encode(const EncodeArgs& a, int nCtas, cudaStream_t s) {
  canon_encode_kernel<<<nCtas, kThreads, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_canon_decode(const DecodeArgs& a, int nCtas, cudaStream_t s) {
  static std::atomic<uint64_t> attr{0};
  cudaError_t ea = once_per_device(attr, [] {
    return cudaFuncSetAttribute(canon_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(CanonDecodeSmem)));
  });
  if (ea != cudaSuccess) return ea;
  canon_decode_kernel<<<nCtas, kThreads, sizeof(CanonDecodeSmem), s>>>(a);
  return cudaGetLastError();
}

}  // namespace g4
