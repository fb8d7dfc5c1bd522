// g4_canon.cu -- CodecCanonHuffman (canonical Huffman over integer predictor residuals) on sm_100a.
//
// Reference: compress/canonicalHuffman/CodecCanonHuffman.java:79-195 (+ the classes cited in g4_canon.cuh),
// under /root/reference/core/src/main/java/org/gridfour/.
#include "g4_kernels.h"
#include "g4_predict.cuh"
#include "g4_canon.cuh"

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

__global__ void __launch_bounds__(kThreads) canon_decode_kernel(DecodeArgs a) {
  __shared__ CanonDecShared S;
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
    else if (pred == G4_PRED_DIFF_NULLS) status = G4_ERR_UNSUPPORTED;  // TODO(next): nulls on the GPU
    else {
      BitSrc src;
      src.init(packing + 6, len - 6);
      uint32_t endBit = 0, nv = 0;
      PredCellSink sink{t, pred};
      const uint32_t expect = uint32_t(n - 1);
      if (!canon_decode_stream(S, src, 0, expect, 0u, sink, &endBit, &nv) || nv != expect) status = G4_ERR_FORMAT;
      else {
        __syncthreads();
        if (tid == 0) t.at(0, 0) = seed;
        __syncthreads();
        predictor_inverse(pred, t, S.scan);
      }
    }
    if (tid == 0) a.status[tIdx] = status;
  }
}

cudaError_t launch_canon_decode(const DecodeArgs& a, int nCtas, cudaStream_t s) {
  canon_decode_kernel<<<nCtas, kThreads, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace g4
