#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ void wait_bar(uint32_t sb, uint32_t phase) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(sb), "r"(phase) : "memory");
}
__global__ void k_bulk(const uint8_t* src, uint8_t* out) {
  __shared__ __align__(1024) uint8_t box[2048];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(&bar), sd = (uint32_t)__cvta_generic_to_shared(box);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb)); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncwarp();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(2048));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sd), "l"(src), "r"(2048), "r"(sb) : "memory");
  }
  wait_bar(sb, 0);
  for (int i = threadIdx.x; i < 2048; i += 32) out[i] = box[i];
}
__global__ void k_tensor(const __grid_constant__ CUtensorMap tm, int x0, int y0, int bytes, uint8_t* out) {
  __shared__ __align__(1024) uint8_t box[8192];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(&bar), sd = (uint32_t)__cvta_generic_to_shared(box);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb)); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncwarp();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(bytes));
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(sd),
                 "l"(reinterpret_cast<uint64_t>(&tm)), "r"(x0), "r"(y0), "r"(sb) : "memory");
  }
  wait_bar(sb, 0);
  for (int i = threadIdx.x; i < bytes; i += 32) out[i] = box[i];
}
int run_tensor(EncodeFn enc, void* d, CUtensorMapDataType dt, int es, cuuint64_t d0, cuuint64_t d1, cuuint64_t stride, cuuint32_t b0, cuuint32_t b1,
               CUtensorMapSwizzle sw, CUtensorMapL2promotion l2, const char* name, uint8_t* dout) {
  alignas(64) CUtensorMap tm;
  cuuint64_t dims[2] = {d0, d1};
  cuuint64_t strides[1] = {stride};
  cuuint32_t box[2] = {b0, b1};
  cuuint32_t est[2] = {1, 1};
  CUresult r = enc(&tm, dt, 2, d, dims, strides, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  k_tensor<<<1, 32>>>(tm, 0, 0, int(b0 * b1 * es), dout);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-40s encode=%d kernel=%s\n", name, int(r), cudaGetErrorString(e));
  return e == cudaSuccess;
}
__global__ void k_loop(const __grid_constant__ CUtensorMap tm, const int* x0s, int n, uint8_t* out) {
  __shared__ __align__(1024) uint8_t box[2][2048];
  __shared__ __align__(8) uint64_t bar[2];
  const int lane = threadIdx.x;
  if (lane == 0) {
    for (int s = 0; s < 2; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncwarp();
  auto issue = [&](int k) {
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(&bar[k & 1]), sd = (uint32_t)__cvta_generic_to_shared(box[k & 1]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(2048));
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(sd),
                 "l"(reinterpret_cast<uint64_t>(&tm)), "r"(x0s[k]), "r"(0), "r"(sb) : "memory");
  };
  if (lane == 0) issue(0);
  for (int k = 0; k < n; k++) {
    if (lane == 0 && k + 1 < n) issue(k + 1);
    wait_bar((uint32_t)__cvta_generic_to_shared(&bar[k & 1]), (k >> 1) & 1);
    for (int c = 0; c < 4; c++) {
      const uint4 v = *reinterpret_cast<const uint4*>(box[k & 1] + 64 * lane + 16 * (c ^ ((lane >> 1) & 3)));
      *reinterpret_cast<uint4*>(out + size_t(k) * 2048 + 64 * lane + 16 * c) = v;
    }
    __syncwarp();
  }
}
int main() {
  int drv = 0, rt = 0;
  cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt);
  printf("driver %d runtime %d\n", drv, rt);
  const size_t total = 1 << 20;
  uint8_t *d, *dout;
  cudaMalloc(&d, total); cudaMemset(d, 7, total);
  cudaMalloc(&dout, 8192);
  k_bulk<<<1, 32>>>(d, dout);
  printf("bulk 1D copy: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr);
  run_tensor(enc, d, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 256, 64, 1024, 16, 32, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, "f32 256x64 box16x32 noswz", dout);
  run_tensor(enc, d, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 4096, 32, 4096, 64, 32, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, "u8 4096x32 box64x32 noswz", dout);
  run_tensor(enc, d, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 4096, 32, 4096, 64, 32, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "u8 4096x32 box64x32 swz64", dout);
  run_tensor(enc, d, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 900000, 32, 1440, 64, 32, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "u8 900000x32 stride1440 swz64", dout);
  {
    std::vector<uint8_t> h(total);
    for (size_t i = 0; i < total; i++) h[i] = uint8_t((i * 2654435761u) >> 13);
    cudaMemcpy(d, h.data(), total, cudaMemcpyHostToDevice);
    const int stride = 1440, n = 7;
    int hx[n] = {0, 16, 64, 128, 46208 + 64, 499968, 1440 * 3 + 32};
    int* dx; cudaMalloc(&dx, sizeof(hx)); cudaMemcpy(dx, hx, sizeof(hx), cudaMemcpyHostToDevice);
    uint8_t* o2; cudaMalloc(&o2, n * 2048);
    alignas(64) CUtensorMap tm;
    cuuint64_t dims[2] = {900000, 32}; cuuint64_t strides[1] = {cuuint64_t(stride)}; cuuint32_t box[2] = {64, 32}; cuuint32_t est[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    k_loop<<<1, 32>>>(tm, dx, n, o2);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint8_t> o(n * 2048);
    cudaMemcpy(o.data(), o2, o.size(), cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int k = 0; k < n; k++) for (int l = 0; l < 32; l++) for (int x = 0; x < 64; x++)
      if (o[size_t(k) * 2048 + 64 * l + x] != h[size_t(hx[k]) + size_t(stride) * l + x]) bad++;
    printf("skew loop: encode=%d kernel=%s mismatches=%ld of %d\n", int(r), cudaGetErrorString(e), bad, n * 2048);
  }
  return 0;
}
