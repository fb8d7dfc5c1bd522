// g4_analyze.cu -- ICompressionDecoder.analyze on the device: the statistics CodecHuffman.analyze (:172-199) and
// CodecDeflate.analyze (:71-106) gather per tile with compress/CodecStats.java:100-155 (paths under
// /root/reference/core/src/main/java/org/gridfour/compress/).
//
// One CTA per tile: the entropy stage runs as in the decoders (legacy Huffman tree walk / inflate) into a per-CTA scratch,
// then the M32 bytes are tallied -- 256-bin histogram in shared memory (distinct symbols, first-order entropy in FP64,
// summed over the bins in the reference's order), successor pairs into the 65,536-bin table of the tile's predictor
// (CodecStats.sB; sA is its column sum).  The host mirror adds the per-tile records up in tile order, as the reference's
// loop over the tiles of a file does (GvrsFile.summarize), and prints reportAnalysisData's table.
#include "g4_kernels.h"
#include "g4_huffdec.cuh"
#include "g4_inflate.cuh"

namespace g4 {

namespace {
union AnalyzeShared {
  HuffDecShared h;
  InflateWarpShared inf;
};
}  // namespace

__global__ void __launch_bounds__(kThreads) analyze_kernel(AnalyzeArgs a) {
  extern __shared__ __align__(16) unsigned char analyzeSmem[];
  AnalyzeShared& S = *reinterpret_cast<AnalyzeShared*>(analyzeSmem);
  __shared__ uint32_t hist[256];
  __shared__ int sInfOk;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint8_t* m32 = a.scratch + size_t(blockIdx.x) * a.scratchStride;
  for (int t = blockIdx.x; t < a.nTiles; t += gridDim.x) {
    __syncthreads();
    g4_tile_stats st{};
    const uint32_t len = a.lens[t];
    const uint8_t* packing = a.arena + a.offsets[t];
    st.codec_kind = -1;
    st.status = G4_DECLINED;  // no M32 statistics: a raw tile, or a codec that keeps none
    int kind = -1;
    if (len > a.rawLen || a.offsets[t] > a.arenaLen || uint64_t(len) > a.arenaLen - a.offsets[t] || len == 0) st.status = G4_ERR_FORMAT;
    else if (len != a.rawLen) {
      const int index = packing[0];
      if (index >= a.codecs.n_codecs) st.status = G4_ERR_FORMAT;
      else kind = a.codecs.codec_ids[index];
      st.codec_kind = kind;
    }
    if (kind == G4_CODEC_HUFFMAN || kind == G4_CODEC_DEFLATE) {
      const int pred = len >= 10 ? int(packing[1]) : 0;
      const uint32_t nM32 = len >= 10 ? load_le32(packing + 6) : 0;
      const uint32_t n = a.nCells;
      bool ok = len >= 12 && pred >= 1 && pred <= 4 && nM32 >= n - 1 && nM32 <= 6u * n;
      uint32_t overhead = 0;
      if (ok) {
        if (kind == G4_CODEC_HUFFMAN) {
          BitSrc src;
          src.init(packing + 10, len - 10);
          uint32_t endBit = 0;
          ok = huffman_decode_stream(S.h, src, 0, nM32, m32, &endBit);
          overhead = S.h.treeBits;  // HuffmanDecoder.getBitsInTreeCount (:195)
        } else {
          if (warp == 0) {
            uint32_t produced = 0, consumed = 0;
            const int rc = inflate_warp(S.inf, packing + 10, len - 10, m32, nM32, &produced, &consumed);
            if (lane == 0) sInfOk = (rc == kInfOk && produced == nM32) ? 1 : 0;
          }
          __syncthreads();
          ok = sInfOk != 0;
        }
      }
      __syncthreads();
      if (!ok) st.status = G4_ERR_FORMAT;
      else {
        hist[tid] = 0;  // kThreads == 256
        __syncthreads();
        unsigned long long* pairs = a.pairs ? a.pairs + (size_t(kind == G4_CODEC_DEFLATE ? 1 : 0) * 5 + size_t(pred)) * 65536 : nullptr;
        for (uint32_t i = tid; i < nM32; i += kThreads) {
          const uint32_t v = m32[i];
          atomicAdd(&hist[v], 1u);
          if (pairs && i > 0) atomicAdd(&pairs[(uint32_t(m32[i - 1]) << 8) | v], 1ull);
        }
        __syncthreads();
        if (tid == 0) {
          // CodecStats.addCountsForM32 (:100-131): distinct symbols, entropy = -sum p log2 p over the bins in index order
          uint32_t observed = 0;
          const double d = double(nM32), LOG2 = log(2.0);
          double s = 0;
          for (int i = 0; i < 256; i++)
            if (hist[i] > 0) {
              observed++;
              const double p = double(hist[i]) / d;
              s += p * log(p) / LOG2;
            }
          st.status = G4_OK;
          st.predictor = pred;
          st.n_bytes = len - 10;
          st.n_symbols = n;
          st.n_bits_overhead = overhead;
          st.n_m32 = nM32;
          st.observed = observed;
          st.entropy = -s;
        }
      }
    }
    if (tid == 0) a.stats[t] = st;
  }
}

cudaError_t launch_analyze(const AnalyzeArgs& a, int nCtas, cudaStream_t s) {
  static std::atomic<uint64_t> attr{0};
  cudaError_t ea = once_per_device(attr, [] {
    return cudaFuncSetAttribute(analyze_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(AnalyzeShared)));
  });
  if (ea != cudaSuccess) return ea;
  analyze_kernel<<<nCtas, kThreads, sizeof(AnalyzeShared), s>>>(a);
  return cudaGetLastError();
}

}  // namespace g4
