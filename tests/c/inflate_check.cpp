// Checks guac_inflate::inflate_member against zlib: raw deflate streams of every block type (stored / fixed / dynamic), levels
// 0-9, all strategies, sizes 0 .. 65,280 (a BGZF member's maximum), five kinds of content; then corrupted and truncated
// streams must be refused or decoded without touching a byte outside the output buffer.  Prints "ok <streams> <fast>".
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../guacamole_b200/csrc/guac_inflate.h"

static std::vector<uint8_t> deflate_raw(const std::vector<uint8_t>& in, int level, int strategy) {
  z_stream zs{};
  if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, strategy) != Z_OK) exit(2);
  std::vector<uint8_t> out(deflateBound(&zs, in.size()) + 64);
  zs.next_in = const_cast<Bytef*>(in.data());
  zs.avail_in = (uInt)in.size();
  zs.next_out = out.data();
  zs.avail_out = (uInt)out.size();
  if (deflate(&zs, Z_FINISH) != Z_STREAM_END) exit(3);
  out.resize(zs.total_out);
  deflateEnd(&zs);
  return out;
}

int main() {
  std::mt19937 rng(12345);
  guac_inflate::Tables* T = new guac_inflate::Tables;
  const size_t sizes[] = {0, 1, 2, 3, 7, 8, 9, 31, 100, 257, 258, 259, 1000, 4096, 20000, 65280};
  long streams = 0, fast = 0;
  for (int kind = 0; kind < 5; ++kind)
    for (size_t n : sizes) {
      std::vector<uint8_t> data(n);
      for (size_t i = 0; i < n; ++i) {
        switch (kind) {
          case 0: data[i] = (uint8_t)rng(); break;                                        // incompressible
          case 1: data[i] = "ACGT"[rng() & 3]; break;                                      // sequence
          case 2: data[i] = (uint8_t)("the quick brown fox "[i % 20] + ((rng() % 50) == 0)); break;  // text with long matches
          case 3: data[i] = 0; break;                                                      // one long run (distance 1)
          default: data[i] = (uint8_t)((i & 64) ? (i * 7) : (rng() % 5 + 30)); break;       // quality-like with structure
        }
      }
      for (int level = 0; level <= 9; ++level)
        for (int strategy : {Z_DEFAULT_STRATEGY, Z_FILTERED, Z_HUFFMAN_ONLY, Z_RLE, Z_FIXED}) {
          std::vector<uint8_t> c = deflate_raw(data, level, strategy);
          const size_t clen = c.size();
          c.insert(c.end(), 8, 0xFF);  // the member's trailer: readable, not data
          std::vector<uint8_t> out(n + 32, 0xA5);
          const bool ok = guac_inflate::inflate_member(c.data(), clen, out.data() + 16, n, *T);
          ++streams;
          for (int g = 0; g < 16; ++g)
            if (out[g] != 0xA5 || out[16 + n + g] != 0xA5) { printf("guard bytes touched (kind %d n %zu level %d strategy %d)\n", kind, n, level, strategy); return 1; }
          if (ok) {
            ++fast;
            if (n && memcmp(out.data() + 16, data.data(), n) != 0) { printf("wrong bytes (kind %d n %zu level %d strategy %d)\n", kind, n, level, strategy); return 1; }
          } else {
            printf("refused a valid stream (kind %d n %zu level %d strategy %d clen %zu)\n", kind, n, level, strategy, clen);
            return 1;
          }
          // wrong ISIZE must be refused
          if (n > 0) {
            std::vector<uint8_t> o2(n + 40, 0xA5);
            if (guac_inflate::inflate_member(c.data(), clen, o2.data() + 16, n - 1, *T)) { printf("accepted a short output\n"); return 1; }
            if (guac_inflate::inflate_member(c.data(), clen, o2.data() + 16, n + 1, *T)) { printf("accepted a long output\n"); return 1; }
            for (int g = 0; g < 16; ++g)
              if (o2[g] != 0xA5 || o2[16 + n + 1 + g] != 0xA5) { printf("guard bytes touched (ISIZE)\n"); return 1; }
          }
          // truncation and bit flips: refused, or decoded inside the buffer
          if (clen > 4 && (streams % 7) == 0) {
            for (int trial = 0; trial < 6; ++trial) {
              std::vector<uint8_t> bad(c);
              size_t blen = clen;
              if (trial < 2) blen = clen - 1 - rng() % (clen / 2);
              else bad[rng() % clen] ^= (uint8_t)(1u << (rng() & 7));
              std::vector<uint8_t> o3(n + 32, 0xA5);
              (void)guac_inflate::inflate_member(bad.data(), blen, o3.data() + 16, n, *T);
              for (int g = 0; g < 16; ++g)
                if (o3[g] != 0xA5 || o3[16 + n + g] != 0xA5) { printf("guard bytes touched (corrupt stream)\n"); return 1; }
            }
          }
        }
    }
  printf("ok %ld %ld\n", streams, fast);
  delete T;
  return 0;
}
