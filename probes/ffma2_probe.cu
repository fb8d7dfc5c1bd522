// probes/ffma2_probe.cu -- issue rate and dependent latency of the packed float32 instructions of sm_100a
// (fma.rn.f32x2 / add.rn.f32x2 / add.rm.f32x2 -> FFMA2 / FADD2) against their scalar forms, and a check that
// mul-as-fma(-0) + add stays UNFUSED (ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o ffma2_probe ffma2_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b, uint64_t nz) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int CHAINS, bool PACKED>
__global__ void rate(float* out, int iters, long long* cyc) {
  float s = threadIdx.x * 1e-3f;
  long long t0 = 0;
  if (PACKED) {
    uint64_t acc[CHAINS];
    const uint64_t u = pk(1.0001f, 0.9999f), nz = pk(-0.f, -0.f);
    for (int i = 0; i < CHAINS; i++) acc[i] = pk(s + i, s - i);
    t0 = clock64();
    for (int k = 0; k < iters; k++)
#pragma unroll
      for (int i = 0; i < CHAINS; i++) acc[i] = add2(mul2(acc[i], u, nz), u);
    long long t1 = clock64();
    float a, b, r = 0;
    for (int i = 0; i < CHAINS; i++) { unpk(acc[i], a, b); r += a + b; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  } else {
    float acc[CHAINS];
    for (int i = 0; i < CHAINS; i++) acc[i] = s + i;
    t0 = clock64();
    for (int k = 0; k < iters; k++)
#pragma unroll
      for (int i = 0; i < CHAINS; i++) acc[i] = __fadd_rn(__fmul_rn(acc[i], 1.0001f), 0.9999f);
    long long t1 = clock64();
    float r = 0;
    for (int i = 0; i < CHAINS; i++) r += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  }
}

__global__ void exact(const float* a, const float* b, const float* c, uint32_t* bad, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t nz = pk(-0.f, -0.f);
  float x, y;
  unpk(add2(mul2(pk(a[i], a[i]), pk(b[i], b[i]), nz), pk(c[i], c[i])), x, y);
  const float want = __fadd_rn(__fmul_rn(a[i], b[i]), c[i]);
  if (__float_as_uint(x) != __float_as_uint(want) || __float_as_uint(y) != __float_as_uint(want)) atomicAdd(bad, 1u);
  uint64_t t; float lo, hi;
  asm volatile("add.rm.f32x2 %0, %1, %2;" : "=l"(t) : "l"(pk(a[i], b[i])), "l"(pk(c[i], c[i])));
  unpk(t, lo, hi);
  if (__float_as_uint(lo) != __float_as_uint(__fadd_rd(a[i], c[i])) || __float_as_uint(hi) != __float_as_uint(__fadd_rd(b[i], c[i]))) atomicAdd(bad + 1, 1u);
}

template <int CHAINS, bool PACKED>
void run(const char* name, int warpsPerSm) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  rate<CHAINS, PACKED><<<148, warpsPerSm * 32>>>(out, iters, cyc);
  rate<CHAINS, PACKED><<<148, warpsPerSm * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double instr = double(iters) * CHAINS * 2;  // per warp
  printf("%-8s chains %d warps/SM %2d: %.2f cycles per warp-instruction per scheduler-slot (%.2f cyc/instr/warp)\n", name, CHAINS, warpsPerSm,
         double(h) / (instr * warpsPerSm / 4.0), double(h) / instr);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<1, false>("scalar", 4); run<1, true>("packed", 4);     // dependent latency (1 warp per scheduler)
  run<8, false>("scalar", 4); run<8, true>("packed", 4);     // issue rate, one warp per scheduler
  run<8, false>("scalar", 16); run<8, true>("packed", 16);   // issue rate, 4 warps per scheduler
  run<4, false>("scalar", 32); run<4, true>("packed", 32);
  const int n = 1 << 22;
  float *a, *b, *c; uint32_t* bad;
  cudaMallocManaged(&a, n * 4); cudaMallocManaged(&b, n * 4); cudaMallocManaged(&c, n * 4); cudaMallocManaged(&bad, 8);
  uint64_t s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  for (int i = 0; i < n; i++) {
    a[i] = float(int64_t(rnd() % 2000001) - 1000000) * 1e-3f; b[i] = float(int64_t(rnd() % 40001) - 20000) * 1e-4f; c[i] = float(int64_t(rnd() % 2000001) - 1000000) * 0.37f;
  }
  bad[0] = bad[1] = 0;
  exact<<<n / 256, 256>>>(a, b, c, bad, n);
  cudaDeviceSynchronize();
  printf("unfused mul(-0)+add mismatches vs scalar: %u of %d; add.rm.f32x2 mismatches vs __fadd_rd: %u\n", bad[0], n, bad[1]);
  return 0;
}
