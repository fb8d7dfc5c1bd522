// probes/ffma2_probe.cu -- issue rate and dependent latency of the packed float32 instructions of sm_100a
// (fma.rn.f32x2 / add.rn.f32x2 / add.rm.f32x2 -> FFMA2 / FADD2) against their scalar forms, and a check that
// mul-as-fma(-0) + add stays UNFUSED: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even with
// --fmad=false, and also fma(a, b, -0) + add when the -0 is a compile-time constant; with the -0 pair as a KERNEL ARGUMENT it
// cannot.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o ffma2_probe ffma2_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b, uint64_t nz) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int CHAINS, bool PACKED>
__global__ void rate(float* out, int iters, long long* cyc, uint64_t nz) {
  float s = threadIdx.x * 1e-3f;
  long long t0 = 0;
  if (PACKED) {
    uint64_t acc[CHAINS];
    const uint64_t u = pk(1.0001f, 0.9999f);
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

__global__ void exact(const float* a, const float* b, const float* c, uint32_t* bad, int n, uint64_t nz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x, y;
  unpk(add2(mul2(pk(a[i], a[i]), pk(b[i], b[i]), nz), pk(c[i], c[i])), x, y);
  const float want = __fadd_rn(__fmul_rn(a[i], b[i]), c[i]);
  if (__float_as_uint(x) != __float_as_uint(want) || __float_as_uint(y) != __float_as_uint(want)) atomicAdd(bad, 1u);
  uint64_t t; float lo, hi;
  asm volatile("add.rm.f32x2 %0, %1, %2;" : "=l"(t) : "l"(pk(a[i], b[i])), "l"(pk(c[i], c[i])));
  unpk(t, lo, hi);
  if (__float_as_uint(lo) != __float_as_uint(__fadd_rd(a[i], c[i])) || __float_as_uint(hi) != __float_as_uint(__fadd_rd(b[i], c[i]))) atomicAdd(bad + 1, 1u);
}


// issue-slot test: per FP instruction (packed: one FFMA2/FADD2 = two pipe cycles; scalar: two FMUL/FADD) INTS independent
// integer instructions.  If the slot behind a packed instruction is free for other pipes, packed + INTS=1 costs what packed alone costs.
template <int INTS, bool PACKED>
__global__ void mixed(float* out, int iters, long long* cyc, uint64_t nz, uint32_t seed) {
  float s = threadIdx.x * 1e-3f;
  uint32_t iv[8];
  for (int i = 0; i < 8; i++) iv[i] = seed * (i + 1) + threadIdx.x;
  uint64_t acc[8];
  float fa[16];
  const uint64_t u = pk(1.0001f, 0.9999f);
  for (int i = 0; i < 8; i++) { acc[i] = pk(s + i, s - i); fa[2 * i] = s + i; fa[2 * i + 1] = s - i; }
  long long t0 = clock64();
  for (int k = 0; k < iters; k++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (PACKED) acc[i] = add2(mul2(acc[i], u, nz), u);
      else { fa[2 * i] = __fadd_rn(__fmul_rn(fa[2 * i], 1.0001f), 0.9999f); fa[2 * i + 1] = __fadd_rn(__fmul_rn(fa[2 * i + 1], 0.9999f), 1.0001f); }
#pragma unroll
      for (int j = 0; j < 2 * INTS; j++) iv[(i + j) & 7] = (iv[(i + j) & 7] ^ seed) + (iv[(i + j + 3) & 7] >> 3);
    }
  }
  long long t1 = clock64();
  float a, b, r = 0;
  for (int i = 0; i < 8; i++) { unpk(acc[i], a, b); r += a + b + fa[2 * i] + fa[2 * i + 1] + float(iv[i]); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int INTS, bool PACKED>
void runmix(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2048, warps = 16;
  for (int r = 0; r < 2; r++) mixed<INTS, PACKED><<<148, warps * 32>>>(out, iters, cyc, 0x8000000080000000ull, 12345u);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  // per scheduler: 4 warps; per iteration and warp: 8 x (2 FP2 or 4 FP) + 8 x 2 INTS x (about 2 integer instructions)
  printf("%-7s ints/fp-pair %d: %.2f cycles per scheduler per (1 chain step = 2 mul + 2 add on two floats + %d int statements)\n", name, INTS,
         double(h) / (double(iters) * 8 * 4), 2 * INTS);
  cudaFree(out); cudaFree(cyc);
}

template <int CHAINS, bool PACKED>
void run(const char* name, int warpsPerSm) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  rate<CHAINS, PACKED><<<148, warpsPerSm * 32>>>(out, iters, cyc, 0x8000000080000000ull);
  rate<CHAINS, PACKED><<<148, warpsPerSm * 32>>>(out, iters, cyc, 0x8000000080000000ull);
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
  runmix<0, false>("scalar"); runmix<0, true>("packed"); runmix<1, false>("scalar"); runmix<1, true>("packed"); runmix<2, false>("scalar"); runmix<2, true>("packed");
  const int n = 1 << 22;
  float *a, *b, *c; uint32_t* bad;
  cudaMallocManaged(&a, n * 4); cudaMallocManaged(&b, n * 4); cudaMallocManaged(&c, n * 4); cudaMallocManaged(&bad, 8);
  uint64_t s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  for (int i = 0; i < n; i++) {
    a[i] = float(int64_t(rnd() % 2000001) - 1000000) * 1e-3f; b[i] = float(int64_t(rnd() % 40001) - 20000) * 1e-4f; c[i] = float(int64_t(rnd() % 2000001) - 1000000) * 0.37f;
  }
  bad[0] = bad[1] = 0;
  exact<<<n / 256, 256>>>(a, b, c, bad, n, 0x8000000080000000ull);
  cudaDeviceSynchronize();
  printf("unfused mul(-0)+add mismatches vs scalar: %u of %d; add.rm.f32x2 mismatches vs __fadd_rd: %u\n", bad[0], n, bad[1]);
  return 0;
}
