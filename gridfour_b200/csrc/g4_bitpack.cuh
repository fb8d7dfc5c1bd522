// g4_bitpack.cuh -- CTA-cooperative LSB-first bit packing (the GPU counterpart of io/BitOutputStore.java).
//
// Every thread knows the bit offset of its contribution from a block scan of bit counts.  Bits are OR-ed
// into a shared-memory window (atomicOr only on the two words a thread shares with its neighbours, plain
// stores for the words it owns) and the window is flushed to HBM with coalesced 32-bit stores.
#pragma once
#include "g4_device.cuh"

namespace g4 {

constexpr int kPackWinWords = 4096;  // 16 KB window

struct BitWindow {
  uint32_t win[kPackWinWords + 4];
};

// Per-tile output cursor (identical in every thread of the CTA).
struct BitOut {
  uint32_t* outWords;  // tile slot in HBM, 4-byte aligned
  uint32_t capWords;
  uint32_t bitPos;     // bits produced so far
  uint32_t gbase;      // global word index of win[0]
};

__device__ inline void bitwin_reset(BitWindow& W, BitOut& o, uint32_t* outWords, uint32_t capWords) {
  for (int i = threadIdx.x; i < kPackWinWords + 4; i += kThreads) W.win[i] = 0;
  o.outWords = outWords;
  o.capWords = capWords;
  o.bitPos = 0;
  o.gbase = 0;
  __syncthreads();
}

// Writes the complete words to HBM and re-bases the window on the current partial word.  All threads call.
__device__ inline void bitwin_flush(BitWindow& W, BitOut& o) {
  __syncthreads();
  uint32_t nWords = (o.bitPos >> 5) - o.gbase;
  for (uint32_t i = threadIdx.x; i < nWords; i += kThreads)
    if (o.gbase + i < o.capWords) o.outWords[o.gbase + i] = W.win[i];
  uint32_t carry = W.win[nWords];
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < kPackWinWords + 4; i += kThreads) W.win[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) W.win[0] = carry;
  o.gbase += nWords;
  __syncthreads();
}

// Makes room for `bits` more bits.  Returns false when they cannot fit even into an empty window.
__device__ inline bool bitwin_reserve(BitWindow& W, BitOut& o, uint32_t bits) {
  if (((o.bitPos + bits + 31) >> 5) + 1 - o.gbase > uint32_t(kPackWinWords)) {
    bitwin_flush(W, o);
    if (((o.bitPos + bits + 31) >> 5) + 1 - o.gbase > uint32_t(kPackWinWords)) return false;
  }
  return true;
}

// Final flush including the partial last word (zero padded).  All threads call.
__device__ inline void bitwin_finish(BitWindow& W, BitOut& o) {
  __syncthreads();
  uint32_t nWords = ((o.bitPos + 31) >> 5) - o.gbase;
  for (uint32_t i = threadIdx.x; i < nWords; i += kThreads)
    if (o.gbase + i < o.capWords) o.outWords[o.gbase + i] = W.win[i];
  __syncthreads();
}

// Single-thread serial writer positioned inside the window (headers, trees, tables).
struct WinSink {
  uint32_t* win;
  uint32_t pos;  // bit position relative to win[0]
  __device__ __forceinline__ void put(uint32_t v, int n) {  // n in 1..32, v < 2^n
    uint32_t w = pos >> 5, o = pos & 31;
    win[w] |= v << o;
    if (o + n > 32) win[w + 1] |= v >> (32 - o);
    pos += n;
  }
};

// Per-thread appender for a contiguous bit range that starts at absolute bit `start`.
struct ThreadBits {
  uint32_t* win;
  uint32_t w;
  uint64_t acc;
  int nacc;
  bool first;
  __device__ __forceinline__ void begin(BitWindow& W, const BitOut& o, uint32_t start) {
    win = W.win;
    w = (start >> 5) - o.gbase;
    nacc = int(start & 31);
    acc = 0;
    first = true;
  }
  __device__ __forceinline__ void put(uint32_t v, int n) {  // n in 0..32, v < 2^n
    acc |= uint64_t(v) << nacc;
    nacc += n;
    if (nacc >= 32) {
      if (first) { atomicOr(&win[w], uint32_t(acc)); first = false; }
      else win[w] = uint32_t(acc);
      w++;
      acc >>= 32;
      nacc -= 32;
    }
  }
  __device__ __forceinline__ void end() {
    if (nacc > 0 && acc != 0) atomicOr(&win[w], uint32_t(acc));
  }
};

}  // namespace g4
