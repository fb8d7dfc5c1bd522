// g4_predictor.cu -- the predictor models on their own (IPredictorModel, compress/IPredictorModel.java:42-173).
//
// The codecs fuse their predictor into the entropy stage; these kernels expose the four models' encode / decode /
// encodeInt / decodeInt for a band of tiles, so that the predictor byte streams (M32, compress/CodecM32.java:257-356)
// and residual arrays can be compared with the reference on their own:
//   PredictorModelDifferencing.java:112-225, PredictorModelLinear.java:66-223, PredictorModelTriangle.java:62-217,
//   PredictorModelDifferencingWithNulls.java:66-269   (paths under /root/reference/core/src/main/java/org/gridfour/).
// Encode is pointwise on the original values (coalesced sweeps, residuals produced in stream order, M32 byte offsets by
// block scan); decode is the M32 START/CONT automaton scan followed by the predictors' prefix scans (g4_predict.cuh).
#include "g4_kernels.h"
#include "g4_predict.cuh"
#include "g4_m32stream.cuh"

namespace g4 {

namespace {

struct PredGet {  // residual k of the stream (A.5 of SURVEY.md)
  TileView t;
  int pred;
  int32_t seed;
  __device__ __forceinline__ int32_t operator()(uint32_t k) const {
    int r, c;
    stream_to_cell(pred, int(k), t.R, t.C, &r, &c);
    return pred == G4_PRED_DIFF_NULLS ? residual_nulls_at(t, r, c, seed) : residual_at(pred, t, r, c);
  }
};

}  // namespace

// One CTA per tile.  intFlavour: residual ints to out (4 bytes each), else M32 bytes.  lens[t] = number of bytes / ints
// (IPredictorModel.encode / encodeInt return value), seeds[t] = getSeed(); status G4_DECLINED where the model returns -1.
__global__ void __launch_bounds__(kThreads) predictor_encode_kernel(PredictorArgs a) {
  __shared__ uint32_t scan[kWarps + 1];
  const int nTiles = a.band.tiles_down * a.band.tiles_across;
  for (int tIdx = blockIdx.x; tIdx < nTiles; tIdx += gridDim.x) {
    __syncthreads();
    const TileView t = tile_view(a.band, a.grid, tIdx);
    const int n = t.R * t.C;
    uint8_t* out = a.slots + size_t(tIdx) * a.slotBytes;
    int32_t seed = t.at(0, 0);
    int status = G4_OK;
    uint32_t N = uint32_t(n - 1);
    if (a.model == G4_PRED_TRIANGLE && (t.R < 2 || t.C < 2)) status = G4_DECLINED;  // PredictorModelTriangle.java:107-109
    if (a.model == G4_PRED_DIFF_NULLS) {
      int nStart = 0;
      seed = nulls_seed(t, &nStart);
      N = uint32_t(n);
      if (nStart == 0) status = G4_DECLINED;  // nothing but nulls
    }
    uint32_t produced = 0;
    if (status == G4_OK) {
      const PredGet get{t, a.model, seed};
      if (a.intFlavour) {
        int32_t* o = reinterpret_cast<int32_t*>(out);
        for (uint32_t k = threadIdx.x; k < N; k += kThreads) o[k] = get(k);
        produced = N;
      } else produced = m32_stream_write(get, N, out, scan);
    }
    if (threadIdx.x == 0) {
      a.lens[tIdx] = produced;
      a.seeds[tIdx] = seed;
      a.status[tIdx] = status;
    }
  }
}

// Inverse.  Input of tile t at slots + t * slotBytes (M32 bytes, 16-byte aligned and padded, or residual ints), lens[t]
// of them; seeds[t] = the seed.  The raster receives the values.
__global__ void __launch_bounds__(kThreads) predictor_decode_kernel(PredictorArgs a) {
  __shared__ uint32_t scan[kWarps + 1];
  const int nTiles = a.band.tiles_down * a.band.tiles_across;
  for (int tIdx = blockIdx.x; tIdx < nTiles; tIdx += gridDim.x) {
    __syncthreads();
    const TileView t = tile_view(a.band, a.grid, tIdx);
    const int n = t.R * t.C;
    const uint8_t* in = a.slots + size_t(tIdx) * a.slotBytes;
    const int32_t seed = a.seeds[tIdx];
    const uint32_t len = a.lens[tIdx];
    const uint32_t expect = a.model == G4_PRED_DIFF_NULLS ? uint32_t(n) : uint32_t(n - 1);
    int status = G4_OK;
    if (a.intFlavour) {
      if (len != expect) status = G4_ERR_FORMAT;
      else {
        const int32_t* r32 = reinterpret_cast<const int32_t*>(in);
        for (uint32_t k = threadIdx.x; k < expect; k += kThreads) {
          int r, c;
          stream_to_cell(a.model, int(k), t.R, t.C, &r, &c);
          t.at(r, c) = r32[k];
        }
      }
    } else if (len < expect || len > 6u * uint32_t(n) || !m32_parse_to_cells(in, len, a.model, t, expect, scan)) status = G4_ERR_FORMAT;
    __syncthreads();
    if (status == G4_OK) {
      if (a.model == G4_PRED_DIFF_NULLS) predictor_inverse_nulls(t, seed);
      else {
        if (threadIdx.x == 0) t.at(0, 0) = seed;
        __syncthreads();
        predictor_inverse(a.model, t, scan);
      }
    }
    if (threadIdx.x == 0) a.status[tIdx] = status;
  }
}

cudaError_t launch_predictor(const PredictorArgs& a, int decode, int nCtas, cudaStream_t s) {
  if (decode) predictor_decode_kernel<<<nCtas, kThreads, 0, s>>>(a);
  else predictor_encode_kernel<<<nCtas, kThreads, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace g4
