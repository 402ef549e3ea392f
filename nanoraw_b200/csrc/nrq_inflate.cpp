// nrq_inflate.cpp -- inflate (RFC 1951 inside the zlib container of RFC 1950) for the raw signals of FAST5 files.
//
// An HDF5 gzip chunk is a zlib stream.  The raw signal of a read (about 180 KB of int16 DAC codes) deflates to
// mostly literals and short matches, which zlib's inflate decodes at about 150 MB/s: 1.2 ms per read, the largest
// item of the per-read host work once the file handling is out of Python.  This decoder does the same job with a
// 64-bit bit buffer that is refilled once per symbol by an unaligned load, 11-bit / 8-bit first-level tables with
// second-level tables behind them, and word-wise match copies.  It never reads outside [src, src + n) nor writes
// outside [dst, dst + cap): close to either end it switches to a byte-wise loop.  The Adler-32 of the output is
// checked like zlib's uncompress does.  Part of libnrqio.so.
#include <cstdint>
#include <cstring>

#include "../../include/nrq_io.h"

namespace {

constexpr int kLitBits = 11, kDistBits = 8, kPreBits = 7;
constexpr int kLitTable = (1 << kLitBits) + 1024, kDistTable = (1 << kDistBits) + 512;

// table entry: bits 0-3 code bits consumed at this level, bits 4-7 kind (one flag each), bits 8-12 extra bits (or second-level
// index bits), bits 16-31 value (literal / base length / base distance / offset of the second-level table)
enum : uint32_t { kInvalid = 0, kLiteral = 1 << 4, kBase = 1 << 5, kEnd = 1 << 6, kSub = 1 << 7, kKindMask = 15 << 4 };
// a literal entry of the first level may carry the literal that follows it as well, when both codes fit the index
// (raw signals: a short code for the high byte of a sample, then 7-9 bits for the low byte): bit 13, value = two bytes
constexpr uint32_t kDouble = 1u << 13;
constexpr int kDoubleShift = 13;

struct Tables {
  uint32_t lit[kLitTable];
  uint32_t dist[kDistTable];
};

const uint16_t kLenBase[29] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23, 27,
                               31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1,   2,   3,   4,   5,   7,    9,    13,   17,   25,   33,   49,   65,    97,    129,
                                193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

inline uint32_t reverse_bits(uint32_t code, int len) {
  uint32_t r = 0;
  for (int i = 0; i < len; ++i) r |= ((code >> i) & 1u) << (len - 1 - i);
  return r;
}

// what a symbol decodes to, without the code length: kind 0 = literal / length alphabet, 1 = distances, 2 = precode
inline uint32_t symbol_entry(int kind, int sym) {
  if (kind == 2) return kLiteral | ((uint32_t)sym << 16);
  if (kind == 1) return sym < 30 ? (kBase | ((uint32_t)kDistExtra[sym] << 8) | ((uint32_t)kDistBase[sym] << 16)) : kInvalid;
  if (sym < 256) return kLiteral | ((uint32_t)sym << 16);
  if (sym == 256) return kEnd;
  if (sym < 286) return kBase | ((uint32_t)kLenExtra[sym - 257] << 8) | ((uint32_t)kLenBase[sym - 257] << 16);
  return kInvalid;
}

// canonical Huffman code of lens[0..n) -> lookup table with `bits` first-level index bits; false if the lengths
// over-subscribe the code space or the table space does not suffice.  Incomplete codes leave invalid entries.
bool build_table(const uint8_t* lens, int n, int kind, int bits, uint32_t* table, int table_size) {
  int count[16] = {0};
  for (int i = 0; i < n; ++i) ++count[lens[i]];
  count[0] = 0;
  int left = 1;
  for (int len = 1; len <= 15; ++len) {
    left = (left << 1) - count[len];
    if (left < 0) return false;
  }
  uint32_t next[16];
  uint32_t code = 0;
  for (int len = 1; len <= 15; ++len) {
    code = (code + (uint32_t)count[len - 1]) << 1;
    next[len] = code;
  }
  const int first = 1 << bits;
  for (int i = 0; i < first; ++i) table[i] = kInvalid;
  // index bits of the second-level table behind each first-level slot: the longest code that starts there
  uint8_t sub_bits[1 << kLitBits];
  bool any_long = false;
  uint32_t codes[288];
  for (int s = 0; s < n; ++s) {
    const int len = lens[s];
    if (!len) continue;
    codes[s] = reverse_bits(next[len]++, len);
    if (len > bits) {
      if (!any_long) {
        std::memset(sub_bits, 0, (size_t)first);
        any_long = true;
      }
      uint8_t& sb = sub_bits[codes[s] & (uint32_t)(first - 1)];
      if (len - bits > sb) sb = (uint8_t)(len - bits);
    }
  }
  int used = first;
  if (any_long) {
    for (int i = 0; i < first; ++i) {
      if (!sub_bits[i]) continue;
      const int size = 1 << sub_bits[i];
      if (used + size > table_size) return false;
      table[i] = kSub | (uint32_t)bits | ((uint32_t)sub_bits[i] << 8) | ((uint32_t)used << 16);
      for (int k = 0; k < size; ++k) table[used + k] = kInvalid;
      used += size;
    }
  }
  for (int s = 0; s < n; ++s) {
    const int len = lens[s];
    if (!len) continue;
    const uint32_t e = symbol_entry(kind, s);
    if (len <= bits) {
      const uint32_t entry = e | (uint32_t)len;
      for (uint32_t i = codes[s]; i < (uint32_t)first; i += 1u << len) table[i] = entry;
    } else {
      const uint32_t slot = table[codes[s] & (uint32_t)(first - 1)];
      const uint32_t base = slot >> 16, sb = (slot >> 8) & 31u;
      const uint32_t entry = e | (uint32_t)(len - bits);
      for (uint32_t i = codes[s] >> bits; i < (1u << sb); i += 1u << (len - bits)) table[base + i] = entry;
    }
  }
  if (kind == 0) {
    uint32_t orig[1 << kLitBits];
    std::memcpy(orig, table, sizeof(uint32_t) * (size_t)first);
    for (int i = 0; i < first; ++i) {
      const uint32_t e1 = orig[i];
      if (!(e1 & kLiteral)) continue;
      const int len1 = (int)(e1 & 15u);
      if (len1 >= bits) continue;
      const uint32_t e2 = orig[(uint32_t)i >> len1];  // the bits after the first code, zeros above them
      const int len2 = (int)(e2 & 15u);
      if (!(e2 & kLiteral) || len1 + len2 > bits) continue;  // (an entry of len2 bits does not depend on the bits above)
      table[i] = kLiteral | kDouble | (uint32_t)(len1 + len2) | (e1 & 0x00ff0000u) | ((e2 & 0x00ff0000u) << 8);
    }
  }
  return true;
}

struct Fixed {
  Tables t;
  Fixed() {
    uint8_t lens[288];
    for (int i = 0; i < 144; ++i) lens[i] = 8;
    for (int i = 144; i < 256; ++i) lens[i] = 9;
    for (int i = 256; i < 280; ++i) lens[i] = 7;
    for (int i = 280; i < 288; ++i) lens[i] = 8;
    build_table(lens, 288, 0, kLitBits, t.lit, kLitTable);
    for (int i = 0; i < 32; ++i) lens[i] = 5;
    build_table(lens, 32, 1, kDistBits, t.dist, kDistTable);
  }
};

struct Bits {
  const uint8_t* in;
  const uint8_t* end;
  uint64_t buf = 0;
  int cnt = 0;  // valid bits in buf
  // fills up to at least 56 bits where the input allows; beyond the input the buffer is fed zeros and `over`
  // counts the bytes that were not there, so that a stream that needs them is refused at the end
  int over = 0;
  inline void refill() {
    if (end - in >= 8) {
      uint64_t w;
      std::memcpy(&w, in, 8);
      buf |= w << cnt;
      in += (63 - cnt) >> 3;
      cnt |= 56;
    } else {
      while (cnt <= 56) {
        if (in < end) buf |= (uint64_t)(*in++) << cnt;
        else ++over;
        cnt += 8;
      }
    }
  }
  inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
  inline void drop(int n) {
    buf >>= n;
    cnt -= n;
  }
  inline uint32_t take(int n) {
    const uint32_t v = peek(n);
    drop(n);
    return v;
  }
};

inline uint32_t lookup(const uint32_t* table, int bits, Bits& b) {
  uint32_t e = table[b.peek(bits)];
  if (e & kSub) {
    b.drop(bits);
    e = table[(e >> 16) + b.peek((int)((e >> 8) & 31u))];
  }
  b.drop((int)(e & 15u));
  return e;
}

bool read_dynamic(Bits& b, Tables& t) {
  static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  b.refill();
  const int hlit = (int)b.take(5) + 257, hdist = (int)b.take(5) + 1, hclen = (int)b.take(4) + 4;
  if (hlit > 286 || hdist > 30) return false;
  uint8_t pre[19] = {0};
  for (int i = 0; i < hclen; ++i) {
    if (b.cnt < 3) b.refill();
    pre[order[i]] = (uint8_t)b.take(3);
  }
  uint32_t ptab[1 << kPreBits];
  if (!build_table(pre, 19, 2, kPreBits, ptab, 1 << kPreBits)) return false;
  uint8_t lens[286 + 30 + 138];
  int i = 0;
  const int total = hlit + hdist;
  while (i < total) {
    b.refill();
    const uint32_t e = lookup(ptab, kPreBits, b);
    if ((e & kKindMask) != kLiteral) return false;
    const int sym = (int)(e >> 16);
    if (sym < 16) {
      lens[i++] = (uint8_t)sym;
    } else {
      int rep;
      uint8_t v = 0;
      if (sym == 16) {
        if (i == 0) return false;
        v = lens[i - 1];
        rep = 3 + (int)b.take(2);
      } else if (sym == 17) {
        rep = 3 + (int)b.take(3);
      } else {
        rep = 11 + (int)b.take(7);
      }
      if (i + rep > total) return false;
      std::memset(lens + i, v, (size_t)rep);
      i += rep;
    }
  }
  if (lens[256] == 0) return false;  // no end-of-block code
  return build_table(lens, hlit, 0, kLitBits, t.lit, kLitTable) &&
         build_table(lens + hlit, hdist, 1, kDistBits, t.dist, kDistTable);
}

// raw deflate stream -> out[0..cap); returns the number of bytes produced, or -1.  *used = bytes of input consumed.
int64_t inflate_raw(const uint8_t* src, size_t n, uint8_t* out, size_t cap, size_t* used) {
  static const Fixed fixed;
  Tables dyn;
  Bits b{src, src + n};
  uint8_t* o = out;
  uint8_t* const o_end = out + cap;
  for (;;) {
    b.refill();
    const uint32_t last = b.take(1), type = b.take(2);
    if (type == 0) {  // stored
      b.drop(b.cnt & 7);
      // give back the whole bytes still in the bit buffer (those that came from the input)
      const int real = (b.cnt >> 3) - b.over;
      if (real < 0) return -1;
      const uint8_t* p = b.in - real;
      b.buf = 0;
      b.cnt = 0;
      b.over = 0;
      if (b.end - p < 4) return -1;
      const uint32_t len = (uint32_t)p[0] | ((uint32_t)p[1] << 8), nlen = (uint32_t)p[2] | ((uint32_t)p[3] << 8);
      if ((len ^ 0xffffu) != nlen) return -1;
      p += 4;
      if ((size_t)(b.end - p) < len || (size_t)(o_end - o) < len) return -1;
      std::memcpy(o, p, len);
      o += len;
      b.in = p + len;
    } else if (type == 1 || type == 2) {
      const Tables* t = &fixed.t;
      if (type == 2) {
        if (!read_dynamic(b, dyn)) return -1;
        t = &dyn;
      }
      for (;;) {
        // fast loop: a whole (length, distance) pair fits one refill, copies may run 8 bytes over
        while (b.end - b.in >= 8 && o_end - o >= 258 + 8) {
          b.refill();
          uint32_t e = lookup(t->lit, kLitBits, b);
          if (e & kLiteral) {  // up to three lookups per refill (3 x 15 bits of the 56), one or two literals each
            uint16_t two = (uint16_t)(e >> 16);
            std::memcpy(o, &two, 2);
            o += 1 + ((e >> kDoubleShift) & 1u);
            e = lookup(t->lit, kLitBits, b);
            if (e & kLiteral) {
              two = (uint16_t)(e >> 16);
              std::memcpy(o, &two, 2);
              o += 1 + ((e >> kDoubleShift) & 1u);
              e = lookup(t->lit, kLitBits, b);
              if (e & kLiteral) {
                two = (uint16_t)(e >> 16);
                std::memcpy(o, &two, 2);
                o += 1 + ((e >> kDoubleShift) & 1u);
                continue;
              }
            }
            if (b.cnt < 48) b.refill();
          }
          if (!(e & kBase)) {
            if (e & kEnd) goto block_done;
            return -1;
          }
          const uint32_t len = (e >> 16) + b.take((int)((e >> 8) & 31u));
          const uint32_t d = lookup(t->dist, kDistBits, b);
          if ((d & kKindMask) != kBase) return -1;
          const uint32_t dist = (d >> 16) + b.take((int)((d >> 8) & 31u));
          if (dist > (size_t)(o - out)) return -1;
          const uint8_t* from = o - dist;
          uint8_t* const stop = o + len;
          if (dist >= 8) {
            do {
              std::memcpy(o, from, 8);
              o += 8;
              from += 8;
            } while (o < stop);
          } else if (dist == 1) {
            std::memset(o, *from, len);
          } else {
            do *o++ = *from++; while (o < stop);
          }
          o = stop;
        }
        // careful step near either end
        b.refill();
        const uint32_t e = lookup(t->lit, kLitBits, b);
        if (e & kLiteral) {
          const size_t k = 1 + ((e >> kDoubleShift) & 1u);
          if ((size_t)(o_end - o) < k) return -1;
          o[0] = (uint8_t)(e >> 16);
          if (k == 2) o[1] = (uint8_t)(e >> 24);
          o += k;
          continue;
        }
        if (e & kEnd) break;
        if (!(e & kBase)) return -1;
        const uint32_t len = (e >> 16) + b.take((int)((e >> 8) & 31u));
        if (b.cnt < 32) b.refill();
        const uint32_t d = lookup(t->dist, kDistBits, b);
        if ((d & kKindMask) != kBase) return -1;
        const uint32_t dist = (d >> 16) + b.take((int)((d >> 8) & 31u));
        if (dist > (size_t)(o - out) || len > (size_t)(o_end - o)) return -1;
        const uint8_t* from = o - dist;
        for (uint32_t k = 0; k < len; ++k) o[k] = from[k];
        o += len;
      }
    block_done:;
    } else {
      return -1;
    }
    if (b.over > 0 && b.cnt < 8 * b.over) return -1;  // bits that were never in the input have been used
    if (last) break;
  }
  if (b.over > 0 && b.cnt < 8 * b.over) return -1;
  if (used) {
    // whole bytes still unread in the bit buffer belong to what follows (the Adler-32)
    const int spare = (b.cnt - 8 * b.over) >> 3;
    *used = (size_t)(b.in - src) - (size_t)(spare > 0 ? spare : 0);
  }
  return (int64_t)(o - out);
}

}  // namespace

extern "C" int64_t nrqio_inflate(const uint8_t* src, int64_t n, uint8_t* dst, int64_t cap) {
  if (!src || n < 6 || !dst || cap < 0) return -1;
  // zlib header: deflate, window <= 32 K, no preset dictionary, check bits
  if ((src[0] & 0x0f) != 8 || (src[0] >> 4) > 7 || (src[1] & 0x20) || (((unsigned)src[0] << 8) | src[1]) % 31 != 0) return -1;
  size_t used = 0;
  const int64_t got = inflate_raw(src + 2, (size_t)n - 2, dst, (size_t)cap, &used);
  if (got < 0) return -1;
  const uint8_t* tail = src + 2 + used;
  if (src + n - tail < 4) return -1;
  const uint32_t want = ((uint32_t)tail[0] << 24) | ((uint32_t)tail[1] << 16) | ((uint32_t)tail[2] << 8) | tail[3];
  const uint32_t a = nrqio_adler32(dst, got);
  if (a != want) return -1;
  return got;
}
