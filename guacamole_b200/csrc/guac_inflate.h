// guac_inflate.h — raw DEFLATE (RFC 1951) decoder for BGZF members, host code (no dependencies).
//
// The BAM front end (guac_bam.cuh) spends nine tenths of its time in zlib's inflate().  A BGZF member is a complete deflate
// stream of at most 64 KB whose inflated size is known beforehand (ISIZE), and its compressed bytes are always followed by
// eight more readable bytes (CRC32 + ISIZE): that allows a decoder without zlib's streaming state machine — a 64-bit bit
// buffer refilled with one unaligned 8-byte load per symbol pair, table look-ups that resolve literals / lengths / distances
// with their extra-bit counts in one entry, and 8-byte match copies.  Anything irregular (a code the tables reject, input or
// output overrun, an incomplete code) makes it return false and the caller inflates that member with zlib instead: this file
// can only make the loader faster, never change what it returns (tests/test_inflate.py compares the two byte for byte over
// streams of every block type, level and strategy).
#pragma once

#include <cstdint>
#include <cstring>

namespace guac_inflate {

constexpr int kLitBits = 11, kDistBits = 8, kPreBits = 7;
constexpr int kLitEnough = 2400, kDistEnough = 420;  // main table + the largest set of subtables (zlib's ENOUGH figures: 2342 / 402)

// table entry: payload << 16 | kind << 12 | extra bits (or subtable bits) << 8 | code length consumed by this look-up
enum Kind : uint32_t { kLength = 1, kEndOfBlock = 2, kSubtable = 3, kInvalid = 4, kDistance = 5, kLiteral = 8 };  // (a literal is bit 15 of the entry)
constexpr uint32_t kLiteralFlag = 0x8000u;
constexpr uint32_t entry(uint32_t payload, uint32_t kind, uint32_t extra, uint32_t len) { return payload << 16 | kind << 12 | extra << 8 | len; }

struct Tables {
  uint32_t lit[kLitEnough];
  uint32_t dist[kDistEnough];
  uint32_t pre[1 << kPreBits];
};

inline uint32_t reverse_bits(uint32_t v, int n) {
  uint32_t r = 0;
  for (int i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1 - i);
  return r;
}

// Canonical Huffman code of `lens[0..n)` -> look-up table of `tb` main bits with subtables behind it (`cap` entries in all).
// make(symbol) gives the entry's payload / kind / extra fields.  False: over-subscribed, incomplete (other than the one-code
// distance tree deflate allows) or out of table space.
template <typename Make>
inline bool build_table(const uint8_t* lens, int n, int tb, uint32_t* table, int cap, Make make, bool allow_single) {
  int count[16] = {0};
  for (int i = 0; i < n; ++i) count[lens[i]]++;
  count[0] = 0;
  int used = 0, max_len = 0;
  long left = 1;
  for (int l = 1; l <= 15; ++l) {
    left = (left << 1) - count[l];
    if (left < 0) return false;  // over-subscribed
    used += count[l];
    if (count[l]) max_len = l;
  }
  for (int i = 0; i < (1 << tb); ++i) table[i] = entry(0, kInvalid, 0, 0);
  if (used == 0) return allow_single;  // no codes at all (a block without matches may send an empty distance tree)
  if (left > 0 && !(allow_single && used == 1)) return false;  // incomplete
  uint32_t next_code[16];
  {
    uint32_t code = 0;
    for (int l = 1; l <= 15; ++l) {
      code = (code + (uint32_t)count[l - 1]) << 1;
      next_code[l] = code;
    }
  }
  // codes longer than the main table: the widest code behind each main-table prefix sizes its subtable
  uint8_t sub_bits[1 << kLitBits];
  if (max_len > tb) memset(sub_bits, 0, (size_t)1 << tb);
  uint32_t codes[320];
  for (int s = 0; s < n; ++s) {
    const int l = lens[s];
    if (!l) continue;
    const uint32_t rev = reverse_bits(next_code[l]++, l);
    codes[s] = rev;
    if (l > tb) {
      uint8_t& sb = sub_bits[rev & ((1u << tb) - 1u)];
      if (l - tb > sb) sb = (uint8_t)(l - tb);
    }
  }
  int next_free = 1 << tb;
  if (max_len > tb) {
    for (int p = 0; p < (1 << tb); ++p) {
      if (!sub_bits[p]) continue;
      const int size = 1 << sub_bits[p];
      if (next_free + size > cap) return false;
      table[p] = entry((uint32_t)next_free, kSubtable, sub_bits[p], (uint32_t)tb);
      for (int i = 0; i < size; ++i) table[next_free + i] = entry(0, kInvalid, 0, 0);
      next_free += size;
    }
  }
  for (int s = 0; s < n; ++s) {
    const int l = lens[s];
    if (!l) continue;
    const uint32_t rev = codes[s];
    if (l <= tb) {
      const uint32_t e = make(s, (uint32_t)l);
      for (uint32_t i = rev; i < (1u << tb); i += 1u << l) {
        if (((table[i] >> 12) & 15u) == kSubtable) return false;  // a short code that is a prefix of a long one: not a prefix code
        table[i] = e;
      }
    } else {
      const uint32_t main = table[rev & ((1u << tb) - 1u)];
      if (((main >> 12) & 15u) != kSubtable) return false;
      const uint32_t base = main >> 16, sb = (main >> 8) & 15u;
      const uint32_t e = make(s, (uint32_t)(l - tb));
      for (uint32_t i = rev >> tb; i < (1u << sb); i += 1u << (l - tb)) table[base + i] = e;
    }
  }
  return true;
}

constexpr uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
constexpr uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
constexpr uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
constexpr uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

inline bool build_litlen(const uint8_t* lens, int n, Tables& T) {
  return build_table(lens, n, kLitBits, T.lit, kLitEnough, [](int s, uint32_t l) {
    if (s < 256) return entry((uint32_t)s, kLiteral, 0, l);
    if (s == 256) return entry(0, kEndOfBlock, 0, l);
    if (s > 285) return entry(0, kInvalid, 0, l);
    return entry(kLenBase[s - 257], kLength, kLenExtra[s - 257], l);
  }, false);
}
inline bool build_dist(const uint8_t* lens, int n, Tables& T) {
  return build_table(lens, n, kDistBits, T.dist, kDistEnough, [](int s, uint32_t l) {
    if (s > 29) return entry(0, kInvalid, 0, l);
    return entry(kDistBase[s], kDistance, kDistExtra[s], l);
  }, true);
}

inline uint64_t load64(const uint8_t* p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return v;  // (little-endian hosts: x86-64 / aarch64)
}

// Inflates one raw deflate stream of exactly `out_len` bytes.  `in[in_len .. in_len + 8)` must be readable (a BGZF member's
// CRC32 + ISIZE follow its data); nothing is written outside out[0 .. out_len).  False = not decoded (use zlib).
inline bool inflate_member(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len, Tables& T) {
  const uint8_t* ip = in;
  const uint8_t* const in_end = in + in_len;
  uint8_t* op = out;
  uint8_t* const out_end = out + out_len;
  uint64_t bitbuf = 0;
  unsigned bitcnt = 0;
  bool overrun = false;
  // at least 56 valid bits after a refill while input remains; behind the input's end the buffer runs dry and `overrun` is set
  // by whoever needs more bits than are left
  auto refill = [&]() {
    if (ip <= in_end) {  // (the 8 bytes behind in_end are readable; bits taken from them are caught by the final check)
      bitbuf |= load64(ip) << bitcnt;
      ip += (63u - bitcnt) >> 3;
      bitcnt |= 56u;
    }
  };
  auto need = [&](unsigned n) {
    if (bitcnt < n) { refill(); if (bitcnt < n) overrun = true; }
  };
  auto take = [&](unsigned n) -> uint32_t {
    const uint32_t v = (uint32_t)(bitbuf & ((1ull << n) - 1ull));
    bitbuf >>= n;
    bitcnt -= n;
    return v;
  };
  for (;;) {
    need(3);
    if (overrun) return false;
    const uint32_t final_block = take(1), type = take(2);
    if (type == 0) {  // stored
      take(bitcnt & 7u);
      need(32);
      if (overrun) return false;
      const uint32_t len = take(16), nlen = take(16);
      if ((len ^ nlen) != 0xFFFFu) return false;
      // the bytes still in the bit buffer belong to the stored data: hand them back
      ip -= bitcnt >> 3;
      bitbuf = 0;
      bitcnt = 0;
      if (ip + len > in_end || op + len > out_end) return false;
      memcpy(op, ip, len);
      ip += len;
      op += len;
    } else if (type == 1 || type == 2) {
      if (type == 1) {  // fixed Huffman codes
        uint8_t lens[288 + 32];
        int i = 0;
        for (; i < 144; ++i) lens[i] = 8;
        for (; i < 256; ++i) lens[i] = 9;
        for (; i < 280; ++i) lens[i] = 7;
        for (; i < 288; ++i) lens[i] = 8;
        for (i = 0; i < 32; ++i) lens[288 + i] = 5;
        if (!build_litlen(lens, 288, T) || !build_dist(lens + 288, 32, T)) return false;
      } else {  // dynamic Huffman codes
        need(14);
        if (overrun) return false;
        const uint32_t hlit = take(5) + 257, hdist = take(5) + 1, hclen = take(4) + 4;
        if (hlit > 286 || hdist > 30) return false;
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint8_t pre_lens[19] = {0};
        for (uint32_t i = 0; i < hclen; ++i) {
          need(3);
          if (overrun) return false;
          pre_lens[order[i]] = (uint8_t)take(3);
        }
        if (!build_table(pre_lens, 19, kPreBits, T.pre, 1 << kPreBits, [](int s, uint32_t l) { return entry((uint32_t)s, kLiteral, 0, l); }, false))
          return false;
        uint8_t lens[286 + 30 + 140];
        uint32_t i = 0;
        while (i < hlit + hdist) {
          need(7 + 7);
          if (overrun) return false;
          const uint32_t e = T.pre[bitbuf & ((1u << kPreBits) - 1u)];
          if (((e >> 12) & 15u) != kLiteral) return false;
          take(e & 255u);
          const uint32_t sym = e >> 16;
          if (sym < 16) {
            lens[i++] = (uint8_t)sym;
          } else {
            uint32_t rep, val = 0;
            if (sym == 16) {
              if (i == 0) return false;
              val = lens[i - 1];
              rep = 3 + take(2);
            } else if (sym == 17) {
              rep = 3 + take(3);
            } else {
              rep = 11 + take(7);
            }
            if (i + rep > hlit + hdist) return false;
            memset(lens + i, (int)val, rep);
            i += rep;
          }
        }
        if (lens[256] == 0) return false;  // no end-of-block code
        if (!build_litlen(lens, (int)hlit, T) || !build_dist(lens + hlit, (int)hdist, T)) return false;
      }
      // ---- the symbols of the block
      constexpr uint64_t kLitMask = (1u << kLitBits) - 1u, kDistMask = (1u << kDistBits) - 1u;
      for (;;) {
        refill();
        if (bitcnt >= 48u && (size_t)(out_end - op) >= 3u + 258u + 8u) {
          // The fast iteration: 48 valid bits cover three literals (15 bits each at most) or a length with its extra bits, a
          // distance code and its extra bits (15 + 5 + 15 + 13), and the output has room for either plus the copy's overshoot,
          // so nothing below needs a bounds check except the distance.
          uint32_t f = T.lit[bitbuf & kLitMask];
          if (f & kLiteralFlag) {
            bitbuf >>= (f & 255u);
            bitcnt -= (f & 255u);
            *op++ = (uint8_t)(f >> 16);
            f = T.lit[bitbuf & kLitMask];
            if (f & kLiteralFlag) {
              bitbuf >>= (f & 255u);
              bitcnt -= (f & 255u);
              *op++ = (uint8_t)(f >> 16);
              f = T.lit[bitbuf & kLitMask];
              if (f & kLiteralFlag) {
                bitbuf >>= (f & 255u);
                bitcnt -= (f & 255u);
                *op++ = (uint8_t)(f >> 16);
              }
            }
            continue;
          }
          uint32_t fk = (f >> 12) & 15u;
          if (fk == kSubtable) {
            bitbuf >>= kLitBits;
            bitcnt -= kLitBits;
            f = T.lit[(f >> 16) + (uint32_t)(bitbuf & ((1u << ((f >> 8) & 15u)) - 1u))];
            fk = (f >> 12) & 15u;
          }
          bitbuf >>= (f & 255u);
          bitcnt -= (f & 255u);
          if (fk == kLiteral) {
            *op++ = (uint8_t)(f >> 16);
            continue;
          }
          if (fk == kEndOfBlock) break;
          if (fk != kLength) return false;
          const uint32_t flx = (f >> 8) & 15u;
          const uint32_t flen = (f >> 16) + (uint32_t)(bitbuf & ((1u << flx) - 1u));
          bitbuf >>= flx;
          bitcnt -= flx;
          uint32_t fd = T.dist[bitbuf & kDistMask];
          if (((fd >> 12) & 15u) == kSubtable) {
            bitbuf >>= kDistBits;
            bitcnt -= kDistBits;
            fd = T.dist[(fd >> 16) + (uint32_t)(bitbuf & ((1u << ((fd >> 8) & 15u)) - 1u))];
          }
          if (((fd >> 12) & 15u) != kDistance) return false;
          bitbuf >>= (fd & 255u);
          bitcnt -= (fd & 255u);
          const uint32_t fdx = (fd >> 8) & 15u;
          const uint32_t fdist = (fd >> 16) + (uint32_t)(bitbuf & ((1u << fdx) - 1u));
          bitbuf >>= fdx;
          bitcnt -= fdx;
          if (fdist > (size_t)(op - out)) return false;
          const uint8_t* fsrc = op - fdist;
          uint8_t* fdst = op;
          op += flen;
          if (fdist >= 8) {
            do {  // (up to 7 bytes past the match, inside the room checked above)
              memcpy(fdst, fsrc, 8);
              fdst += 8;
              fsrc += 8;
            } while (fdst < op);
          } else if (fdist == 1) {
            memset(fdst, *fsrc, flen);
          } else {
            do { *fdst++ = *fsrc++; } while (fdst < op);
          }
          continue;
        }
        // the careful iteration (the last bytes of the input or of the output): every step checked
        uint32_t e = T.lit[bitbuf & ((1u << kLitBits) - 1u)];
        if (((e >> 12) & 15u) == kSubtable) {
          if (bitcnt < (unsigned)kLitBits) return false;
          bitbuf >>= kLitBits;
          bitcnt -= kLitBits;
          e = T.lit[(e >> 16) + (uint32_t)(bitbuf & ((1u << ((e >> 8) & 15u)) - 1u))];
        }
        const uint32_t kind = (e >> 12) & 15u, len_bits = e & 255u;
        if (bitcnt < len_bits) return false;  // (input exhausted)
        bitbuf >>= len_bits;
        bitcnt -= len_bits;
        if (kind == kLiteral) {
          if (op >= out_end) return false;
          *op++ = (uint8_t)(e >> 16);
          // a second literal from the same refill, the common case in sequence data
          uint32_t e2 = T.lit[bitbuf & ((1u << kLitBits) - 1u)];
          if (((e2 >> 12) & 15u) == kLiteral && bitcnt >= 32u && op < out_end) {
            bitbuf >>= (e2 & 255u);
            bitcnt -= (e2 & 255u);
            *op++ = (uint8_t)(e2 >> 16);
          }
          continue;
        }
        if (kind == kEndOfBlock) break;
        if (kind != kLength) return false;
        const uint32_t lx = (e >> 8) & 15u;
        if (bitcnt < lx + 15u + 13u + 15u) {  // (a length's extra bits, a distance code and its extra bits: 48 at most)
          refill();
        }
        if (bitcnt < lx) return false;
        const uint32_t length = (e >> 16) + take(lx);
        uint32_t d = T.dist[bitbuf & ((1u << kDistBits) - 1u)];
        if (((d >> 12) & 15u) == kSubtable) {
          if (bitcnt < (unsigned)kDistBits) return false;
          bitbuf >>= kDistBits;
          bitcnt -= kDistBits;
          d = T.dist[(d >> 16) + (uint32_t)(bitbuf & ((1u << ((d >> 8) & 15u)) - 1u))];
        }
        if (((d >> 12) & 15u) != kDistance) return false;
        const uint32_t dl = d & 255u, dx = (d >> 8) & 15u;
        if (bitcnt < dl + dx) return false;
        bitbuf >>= dl;
        bitcnt -= dl;
        const uint32_t distance = (d >> 16) + take(dx);
        if (distance > (size_t)(op - out) || length > (size_t)(out_end - op)) return false;
        const uint8_t* src = op - distance;
        if (distance >= 8 && (size_t)(out_end - op) >= (size_t)length + 8) {
          uint8_t* dst = op;
          uint8_t* const stop = op + length;
          do {  // (may write up to 7 bytes past the match, still inside this member's output)
            memcpy(dst, src, 8);
            dst += 8;
            src += 8;
          } while (dst < stop);
        } else {
          for (uint32_t i = 0; i < length; ++i) op[i] = src[i];
        }
        op += length;
      }
    } else {
      return false;
    }
    if (final_block) break;
  }
  // every bit taken must have been real input, and the output must be exactly what ISIZE promised
  const uint8_t* consumed_to = ip - (bitcnt >> 3);
  return consumed_to <= in_end && op == out_end;
}

}  // namespace guac_inflate
