// g4_m32stream.cuh -- CTA-cooperative production of an M32 byte stream from a residual sequence.
//
// The reference fills `byte[] mCode` with one serial CodecM32.encode call per residual
// (compress/CodecM32.java:257-311, driven by the predictor models' encode loops, e.g.
// compress/PredictorModelTriangle.java:101-145; paths under /root/reference/core/src/main/java/org/gridfour/).
// Here every thread encodes a few consecutive residuals; byte offsets come from a block scan of the code
// lengths ("warp-scan byte offsets"), so the stream is byte-identical.
#pragma once
#include "g4_device.cuh"

namespace g4 {

// Number of M32 bytes of the stream get(0..N).  All threads call; sm: kWarps+1 words.
template <class Get>
__device__ inline uint32_t m32_stream_size(Get get, uint32_t N, uint32_t* sm) {
  uint32_t my = 0;
  for (uint32_t k = threadIdx.x; k < N; k += kThreads) my += uint32_t(m32_length(get(k)));
  uint32_t total;
  block_exclusive_scan(my, sm, &total);
  __syncthreads();
  return total;
}

// Writes the M32 bytes of get(0..N) to dst (any alignment).  Returns the byte count.  All threads call.
template <class Get>
__device__ inline uint32_t m32_stream_write(Get get, uint32_t N, uint8_t* dst, uint32_t* sm) {
  constexpr int kIpt = 4;
  uint32_t base = 0;
  for (uint32_t k0 = 0; k0 < N; k0 += kThreads * kIpt) {
    uint64_t packed[kIpt];
    int nb[kIpt];
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < kIpt; j++) {
      uint32_t k = k0 + threadIdx.x * kIpt + j;
      nb[j] = 0;
      packed[j] = 0;
      if (k < N) { nb[j] = m32_encode(get(k), &packed[j]); mine += uint32_t(nb[j]); }
    }
    uint32_t total;
    uint32_t p = base + block_exclusive_scan(mine, sm, &total);
#pragma unroll
    for (int j = 0; j < kIpt; j++)
      for (int q = 0; q < nb[j]; q++) dst[p++] = uint8_t(packed[j] >> (8 * q));
    base += total;
  }
  __syncthreads();
  return base;
}

}  // namespace g4
