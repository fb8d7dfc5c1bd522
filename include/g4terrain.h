/* g4terrain.h -- position-addressable synthetic fractal terrain (benchmark / test input generator).
 *
 * Not part of the reference: the reference's benchmark inputs are GEBCO/ETOPO NetCDF files
 * (demo/src/main/java/org/gridfour/demo/globalDEM/PackageData.java:144-600) which are not available
 * offline.  This generator produces elevation-like int32 / float32 grids of the BASELINE.json shapes
 * (SURVEY.md section 8d).  Integer-only arithmetic so the CPU (oracle side) and the GPU (bench side)
 * produce identical samples for any (row, col) without touching floating point:
 *
 *   z(r,c) = sum_{o=0..13} amp[o] * valuenoise_o(r / 2^(13-o), c / 2^(13-o))
 *
 * valuenoise = smoothstep-interpolated lattice of SplitMix64-hashed int16 values; amp[o] = 2^(-0.85 o)
 * in Q16.  Range about [-11000, +9000] "metres"; neighbouring samples differ by a few metres, so
 * predictor residuals are mostly one M32 byte, as for real 15-arc-second bathymetry.
 */
#ifndef G4TERRAIN_H
#define G4TERRAIN_H
#include <stdint.h>

#if defined(__CUDACC__)
#define G4T_HD __host__ __device__ __forceinline__
#else
#define G4T_HD static inline
#endif

#define G4_TERRAIN_SEED 0x9E3779B97F4A7C15ull

G4T_HD uint64_t g4t_mix(uint64_t x) { /* SplitMix64 finaliser */
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

G4T_HD int32_t g4t_lattice(uint64_t seed, int o, int64_t i, int64_t j) {
  uint64_t x = seed + (uint64_t)(o + 1) * 0x9E3779B97F4A7C15ull + (uint64_t)i * 0xD1B54A32D192ED03ull +
               (uint64_t)j * 0x8CB92BA72F3D8DD7ull;
  return (int32_t)(int16_t)(g4t_mix(x) >> 48); /* [-32768, 32767] */
}

G4T_HD int64_t g4t_smooth(int64_t t) { /* t in Q16 [0,65536) -> 3t^2 - 2t^3 in Q16 */
  int64_t t2 = (t * t) >> 16;
  return (t2 * ((3ll << 16) - 2 * t)) >> 16;
}

/* sum of octaves, Q16 amplitude * Q15 noise; |acc| <= 147161 * 32768 */
G4T_HD int64_t g4t_acc(uint64_t seed, int64_t r, int64_t c) {
  const int32_t amp[14] = {65536, 36358, 20171, 11191, 6208, 3444, 1911, 1060, 588, 326, 181, 100, 56, 31};
  int64_t acc = 0;
  for (int o = 0; o < 14; o++) {
    int sh = 13 - o;
    int64_t i = r >> sh, j = c >> sh;
    int64_t fr = (r - (i << sh)) << (16 - sh);
    int64_t fc = (c - (j << sh)) << (16 - sh);
    int64_t sr = g4t_smooth(fr), sc = g4t_smooth(fc);
    int64_t h00 = g4t_lattice(seed, o, i, j), h01 = g4t_lattice(seed, o, i, j + 1);
    int64_t h10 = g4t_lattice(seed, o, i + 1, j), h11 = g4t_lattice(seed, o, i + 1, j + 1);
    int64_t a = (h00 << 16) + (h01 - h00) * sc; /* Q16 of a Q15 value */
    int64_t b = (h10 << 16) + (h11 - h10) * sc;
    int64_t n = ((a << 16) + (b - a) * sr) >> 32; /* back to Q15 integer */
    acc += (int64_t)amp[o] * n;
  }
  return acc;
}

/* elevation in whole metres (int32 configs 1,2,3,5) */
G4T_HD int32_t g4_terrain_m(uint64_t seed, int64_t r, int64_t c) {
  return (int32_t)((g4t_acc(seed, r, c) * 12520ll) >> 32) - 1050;
}

/* elevation in decimetres; the float32 config 4 sample is (float)dm * 0.1f */
G4T_HD int32_t g4_terrain_dm(uint64_t seed, int64_t r, int64_t c) {
  return (int32_t)((g4t_acc(seed, r, c) * 125200ll) >> 32) - 10500;
}

#endif
