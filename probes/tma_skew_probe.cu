// Probe (round 2): does cuTensorMapEncodeTiled accept a row stride SMALLER than the row extent (overlapping rows), and
// does a 32 x 64-byte box with SWIZZLE_64B land in shared memory where the wavefront kernel expects it?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_skew_probe tma_skew_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe_kernel(const __grid_constant__ CUtensorMap tm, const int* x0s, int n, uint8_t* out, int swz) {
  __shared__ __align__(1024) uint8_t box[2048];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(&bar), sd = (uint32_t)__cvta_generic_to_shared(box);
  const int lane = threadIdx.x;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncwarp();
  uint32_t phase = 0;
  for (int k = 0; k < n; k++) {
    if (lane == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(2048));
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(sd),
                   "l"(&tm), "r"(x0s[k]), "r"(0), "r"(sb)
                   : "memory");
    }
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(sb), "r"(phase) : "memory");
    phase ^= 1;
    // de-swizzle: row l = lane, 16-byte chunk c stored at chunk c ^ ((l >> 1) & 3)
    for (int c = 0; c < 4; c++) {
      const uint4 v = *reinterpret_cast<const uint4*>(box + 64 * lane + 16 * (swz == 2 ? (c ^ ((lane >> 1) & 3)) : c));
      *reinterpret_cast<uint4*>(out + size_t(k) * 2048 + 64 * lane + 16 * c) = v;
    }
    __syncwarp();
  }
}

int main(int argc, char** argv) {
  const size_t total = 1 << 20;
  const int stride = argc > 1 ? atoi(argv[1]) : 1440;
  const int swz = argc > 2 ? atoi(argv[2]) : 2;
  const size_t dim0 = argc > 3 ? size_t(atol(argv[3])) : total - 64 * 1024;
  std::vector<uint8_t> h(total);
  for (size_t i = 0; i < total; i++) h[i] = uint8_t((i * 2654435761u) >> 13);
  uint8_t* d;
  cudaMalloc(&d, total);
  cudaMemcpy(d, h.data(), total, cudaMemcpyHostToDevice);
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr);
  printf("entry point: %s qr=%d fn=%p\n", cudaGetErrorString(e), int(qr), (void*)enc);
  if (!enc) return 1;
  CUtensorMap tm;
  cuuint64_t dims[2] = {dim0, 32};
  cuuint64_t strides[1] = {cuuint64_t(stride)};
  cuuint32_t box[2] = {64, 32};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("cuTensorMapEncodeTiled (stride %d < extent): CUresult=%d\n", stride, int(r));
  if (r != CUDA_SUCCESS) return 2;
  const int n = 6;
  int hx[n] = {0, 4, 64, 100, 46208 + 8, 499996};
  if (dim0 < 500000) for (int k = 0; k < n; k++) hx[k] %= int(dim0 - 64);
  int* dx;
  cudaMalloc(&dx, sizeof(hx));
  cudaMemcpy(dx, hx, sizeof(hx), cudaMemcpyHostToDevice);
  uint8_t* dout;
  cudaMalloc(&dout, n * 2048);
  probe_kernel<<<1, 32>>>(tm, dx, n, dout, swz);
  e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<uint8_t> o(n * 2048);
  cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int k = 0; k < n; k++)
    for (int l = 0; l < 32; l++)
      for (int x = 0; x < 64; x++) {
        const uint8_t want = h[size_t(hx[k]) + size_t(stride) * l + x];
        if (o[size_t(k) * 2048 + 64 * l + x] != want) bad++;
      }
  printf("skewed box check: %ld mismatches of %d\n", bad, n * 2048);
  return bad ? 3 : 0;
}
