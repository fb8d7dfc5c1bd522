// Host-side pin of the hand-written DEFLATE encoder (gridfour_b200/csrc/g4_deflate_enc.cuh) against the
// system zlib: same bytes for the same input at levels 6 and 9.  Test infrastructure only; built by
// tests/test_deflate_host.py with `nvcc -x cu` (host code path of the header) and linked with -lz.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <zlib.h>
#include "../../gridfour_b200/csrc/g4_deflate_enc.cuh"

static uint64_t rng = 0x9E3779B97F4A7C15ull;
static uint32_t next() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return uint32_t(rng >> 11); }

static std::vector<uint8_t> make_input(int kind, size_t n) {
  std::vector<uint8_t> v(n + 16, 0);
  switch (kind) {
    case 0: for (size_t i = 0; i < n; i++) v[i] = uint8_t(next()); break;                       // incompressible
    case 1: for (size_t i = 0; i < n; i++) v[i] = uint8_t(int(next() % 7) - 3); break;           // small residuals
    case 2: { int x = 0; for (size_t i = 0; i < n; i++) { x += int(next() % 5) - 2; v[i] = uint8_t(x); } break; }
    case 3: for (size_t i = 0; i < n; i++) v[i] = uint8_t((i / 300) & 0xff); break;              // long runs
    case 4: for (size_t i = 0; i < n; i++) v[i] = uint8_t((next() % 100 < 90) ? 0 : next() % 256); break;
    case 5: { for (size_t i = 0; i < n; i++) v[i] = uint8_t(next() % 3); for (size_t i = 4000; i + 5000 < n; i += 9000) memcpy(&v[i], &v[i - 3333], 700); break; }
    case 6: for (size_t i = 0; i < n; i++) v[i] = 0; break;
    default: for (size_t i = 0; i < n; i++) v[i] = uint8_t((i * 7 + (next() % 4 == 0)) & 0x7f); break;
  }
  for (size_t i = n; i < n + 16; i++) v[i] = 0;
  return v;
}

int main(int argc, char** argv) {
  const size_t sizes[] = {0, 1, 2, 3, 4, 10, 100, 257, 258, 259, 1000, 16383, 16384, 16385, 32767, 32768, 32769, 43199, 65535, 65536, 65537, 70000, 100000, 131072, 259200, 300000};
  g4::DeflateWork* W = new g4::DeflateWork();
  int fails = 0, runs = 0;
  for (int level : {6, 9}) {
    for (int kind = 0; kind < 8; kind++) {
      for (size_t n : sizes) {
        if (argc > 1 && n > size_t(atoi(argv[1]))) continue;
        std::vector<uint8_t> in = make_input(kind, n);
        for (int capMode = 0; capMode < 2; capMode++) {
          size_t cap = capMode == 0 ? n + 118 : (n / 2 + 20);
          std::vector<uint8_t> ref(cap + 8, 0xAA), got(cap + 8, 0xAA);
          z_stream s;
          memset(&s, 0, sizeof(s));
          deflateInit(&s, level);
          s.next_in = in.data(); s.avail_in = uInt(n); s.next_out = ref.data(); s.avail_out = uInt(cap);
          deflate(&s, Z_FINISH);
          size_t refLen = cap - s.avail_out;
          deflateEnd(&s);
          uint32_t gotLen = g4::deflate_stream(in.data(), uint32_t(n), got.data(), uint32_t(cap), W, level);
          runs++;
          if (n <= g4::kDefStagedMax) {  // staged form: sorted position list -> match table -> table-driven lazy loop
            std::vector<uint16_t> sorted(n + 8), rank(n + 8);
            std::vector<uint2> table(n + 8);
            g4::def_sort_positions_host(in.data(), uint32_t(n), sorted.data(), rank.data());
            const g4::DeflateLevel L = g4::deflate_level(level);
            for (size_t slot = 0; slot + 2 < n; slot++)
              table[sorted[slot]] = g4::def_find_match(in.data(), uint32_t(n), L, sorted.data(), uint32_t(slot), rank[slot]);
            std::vector<uint8_t> got2(cap + 8, 0xAA);
            uint32_t got2Len = g4::deflate_stream_table(in.data(), uint32_t(n), got2.data(), uint32_t(cap), W, level, table.data());
            runs++;
            {  // the split form the device uses: decisions, then block emission from the symbol list
              std::vector<uint16_t> sd(n + 8);
              std::vector<uint8_t> sl(n + 8), got3(cap + 8, 0xAA);
              g4::DeflateBlocks B;
              g4::deflate_decide_table(in.data(), uint32_t(n), level, table.data(), sd.data(), sl.data(), &B);
              uint32_t got3Len = g4::deflate_emit_blocks_serial(in.data(), uint32_t(n), got3.data(), uint32_t(cap), W, level, sd.data(), sl.data(), B);
              runs++;
              if (got3Len != refLen || memcmp(ref.data(), got3.data(), refLen) != 0) {
                printf("MISMATCH (decide+emit) level=%d kind=%d n=%zu cap=%zu ref=%zu got=%u\n", level, kind, n, cap, refLen, got3Len);
                fails++;
              }
            }
            if (got2Len != refLen || memcmp(ref.data(), got2.data(), refLen) != 0) {
              size_t d = 0;
              while (d < refLen && d < got2Len && ref[d] == got2[d]) d++;
              printf("MISMATCH (staged) level=%d kind=%d n=%zu cap=%zu ref=%zu got=%u firstdiff=%zu\n", level, kind, n, cap, refLen, got2Len, d);
              fails++;
            }
          }
          if (gotLen != refLen || memcmp(ref.data(), got.data(), refLen) != 0) {
            size_t d = 0;
            while (d < refLen && d < gotLen && ref[d] == got[d]) d++;
            printf("MISMATCH level=%d kind=%d n=%zu cap=%zu ref=%zu got=%u firstdiff=%zu\n", level, kind, n, cap, refLen, gotLen, d);
            fails++;
          }
        }
      }
    }
  }
  printf("%d runs, %d mismatches (zlib %s)\n", runs, fails, zlibVersion());
  return fails ? 1 : 0;
}
