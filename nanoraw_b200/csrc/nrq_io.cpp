// libnrqio.so (include/nrq_io.h): host-side byte work of the FAST5 write-back -- a deflate encoder without match
// search and the packing of a result batch into the gzip chunks write_new_fast5_group stores
// (nanoraw/resquiggle.py:112-124).  Plain C++17, no CUDA, no zlib.
#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nrq_io.h"

namespace {

constexpr int kSyms = 257;     // 256 literals + end of block
constexpr int kMaxBits = 15;   // deflate's limit on a code length
constexpr int64_t kStoredMax = 65535;

uint32_t adler32_scalar(uint32_t a, uint32_t b, const uint8_t* p, int64_t n) {
  while (n > 0) {
    const int64_t k = n < 5552 ? n : 5552;  // largest run before the sums can overflow 32 bits
    for (int64_t i = 0; i < k; ++i) { a += p[i]; b += a; }
    a %= 65521u;
    b %= 65521u;
    p += k;
    n -= k;
  }
  return (b << 16) | a;
}

// 32 bytes per step: byte sums by PSADBW, position-weighted sums by PMADDUBSW / PMADDWD against the weights 32 .. 1,
// and the sum of the byte sums so far (times 32) for what earlier steps of a run contribute to the second sum.
// Runs of at most 173 steps keep every 32-bit lane below 2^32.  (An Events table is checksummed once when it is
// deflated and a raw signal once when it is inflated: at one byte per cycle that was a fifth of either.)
__attribute__((target("avx2"))) uint32_t adler32_avx2(const uint8_t* p, int64_t n) {
  uint32_t a = 1, b = 0;
  const __m256i weights = _mm256_setr_epi8(32, 31, 30, 29, 28, 27, 26, 25, 24, 23, 22, 21, 20, 19, 18, 17, 16, 15, 14,
                                           13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1);
  const __m256i ones = _mm256_set1_epi16(1), zero = _mm256_setzero_si256();
  while (n >= 32) {
    int64_t steps = n / 32 < 173 ? n / 32 : 173;
    n -= 32 * steps;
    __m256i v_ps = _mm256_setzero_si256(), v_s1 = zero, v_s2 = zero;
    const uint64_t lead = (uint64_t)a * (uint64_t)(32 * steps);  // what the sum so far adds to the second sum
    do {
      const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
      p += 32;
      v_ps = _mm256_add_epi32(v_ps, v_s1);
      v_s1 = _mm256_add_epi32(v_s1, _mm256_sad_epu8(x, zero));
      v_s2 = _mm256_add_epi32(v_s2, _mm256_madd_epi16(_mm256_maddubs_epi16(x, weights), ones));
    } while (--steps);
    alignas(32) uint32_t l1[8], l2[8], lp[8];
    _mm256_store_si256(reinterpret_cast<__m256i*>(l1), v_s1);
    _mm256_store_si256(reinterpret_cast<__m256i*>(l2), v_s2);
    _mm256_store_si256(reinterpret_cast<__m256i*>(lp), v_ps);
    uint64_t s1 = 0, s2 = 0;
    for (int k = 0; k < 8; ++k) { s1 += l1[k]; s2 += (uint64_t)l2[k] + 32ull * lp[k]; }
    b = (uint32_t)(((uint64_t)b + lead + s2) % 65521u);
    a = (uint32_t)(((uint64_t)a + s1) % 65521u);
  }
  return adler32_scalar(a, b, p, n);
}

uint32_t adler32(const uint8_t* p, int64_t n) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  return avx2 ? adler32_avx2(p, n) : adler32_scalar(1, 0, p, n);
}

// Huffman code lengths (<= kMaxBits, complete code) for the symbols with freq > 0
void code_lengths(const uint32_t* freq, uint8_t* len) {
  struct Leaf { uint32_t f; int sym; };
  Leaf leaf[kSyms];
  int m = 0;
  for (int s = 0; s < kSyms; ++s) {
    len[s] = 0;
    if (freq[s]) leaf[m++] = {freq[s], s};
  }
  if (m == 1) { len[leaf[0].sym] = 1; return; }
  std::sort(leaf, leaf + m, [](const Leaf& x, const Leaf& y) { return x.f < y.f || (x.f == y.f && x.sym < y.sym); });
  // two queues: leaves in ascending order, internal nodes in the order they are made (ascending as well)
  uint64_t weight[2 * kSyms];
  int parent[2 * kSyms];
  for (int i = 0; i < m; ++i) weight[i] = leaf[i].f;
  int li = 0, ni = m, made = m;
  auto take = [&]() {
    if (li < m && (ni >= made || weight[li] <= weight[ni])) return li++;
    return ni++;
  };
  while (made < 2 * m - 1) {
    const int x = take(), y = take();
    weight[made] = weight[x] + weight[y];
    parent[x] = parent[y] = made;
    ++made;
  }
  int depth[2 * kSyms];
  depth[2 * m - 2] = 0;
  for (int i = 2 * m - 3; i >= 0; --i) depth[i] = depth[parent[i]] + 1;
  int count[64] = {0};
  for (int i = 0; i < m; ++i) ++count[depth[i] < 63 ? depth[i] : 63];
  // limit to kMaxBits: fold the deeper leaves into the last level, then pay the Kraft sum back one leaf at a time
  for (int d = kMaxBits + 1; d < 64; ++d) { count[kMaxBits] += count[d]; count[d] = 0; }
  uint32_t total = 0;
  for (int d = kMaxBits; d >= 1; --d) total += (uint32_t)count[d] << (kMaxBits - d);
  while (total != (1u << kMaxBits)) {
    --count[kMaxBits];
    for (int d = kMaxBits - 1; d >= 1; --d)
      if (count[d]) { --count[d]; count[d + 1] += 2; break; }
    --total;
  }
  int j = 0;  // least frequent symbols take the longest codes
  for (int d = kMaxBits; d >= 1; --d)
    for (int c = count[d]; c > 0; --c) len[leaf[j++].sym] = (uint8_t)d;
}

inline uint32_t reverse_bits(uint32_t v, int n) {
  uint32_t r = 0;
  for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
  return r;
}

struct BitWriter {
  uint8_t* p;
  uint64_t bits = 0;
  int n = 0;
  explicit BitWriter(uint8_t* dst) : p(dst) {}
  inline void put(uint32_t v, int len) {  // len <= 16
    bits |= (uint64_t)v << n;
    n += len;
    if (n >= 32) {
      std::memcpy(p, &bits, 4);
      p += 4;
      bits >>= 32;
      n -= 32;
    }
  }
  uint8_t* finish() {
    while (n > 0) { *p++ = (uint8_t)bits; bits >>= 8; n -= 8; }
    n = 0;
    return p;
  }
  // write out the whole bytes that are pending; fewer than 8 bits stay in *b / *cnt; returns the write position
  uint8_t* finish_to_byte_state(uint64_t* b, int* cnt) {
    while (n >= 8) { *p++ = (uint8_t)bits; bits >>= 8; n -= 8; }
    *b = bits;
    *cnt = n;
    return p;
  }
};

int64_t stored_size(int64_t n) { return 2 + n + 5 * ((n + kStoredMax - 1) / kStoredMax + (n == 0)) + 4; }

int64_t write_stored(const uint8_t* src, int64_t n, uint8_t* dst) {
  uint8_t* p = dst;
  *p++ = 0x78; *p++ = 0x01;
  int64_t left = n;
  do {
    const int64_t k = left < kStoredMax ? left : kStoredMax;
    *p++ = left == k ? 1 : 0;  // BFINAL, BTYPE = 00
    *p++ = (uint8_t)(k & 255); *p++ = (uint8_t)(k >> 8);
    *p++ = (uint8_t)(~k & 255); *p++ = (uint8_t)((~k >> 8) & 255);
    std::memcpy(p, src + (n - left), (size_t)k);
    p += k;
    left -= k;
  } while (left > 0);
  const uint32_t ad = adler32(src, n);
  *p++ = (uint8_t)(ad >> 24); *p++ = (uint8_t)(ad >> 16); *p++ = (uint8_t)(ad >> 8); *p++ = (uint8_t)ad;
  return p - dst;
}

// one dynamic-Huffman block of literals; returns the stream length (dst must hold nrqio_deflate_bound(n))
int64_t write_huffman(const uint8_t* src, int64_t n, uint8_t* dst) {
  uint32_t h[4][256];
  std::memset(h, 0, sizeof(h));
  int64_t i = 0;
  for (; i + 3 < n; i += 4) { ++h[0][src[i]]; ++h[1][src[i + 1]]; ++h[2][src[i + 2]]; ++h[3][src[i + 3]]; }
  for (; i < n; ++i) ++h[0][src[i]];
  uint32_t freq[kSyms];
  for (int s = 0; s < 256; ++s) freq[s] = h[0][s] + h[1][s] + h[2][s] + h[3][s];
  freq[256] = 1;
  uint8_t len[kSyms];
  code_lengths(freq, len);
  // canonical codes (RFC 1951 3.2.2), bit-reversed for the LSB-first stream
  uint32_t next[kMaxBits + 2] = {0}, cnt[kMaxBits + 2] = {0};
  for (int s = 0; s < kSyms; ++s) ++cnt[len[s]];
  cnt[0] = 0;
  uint32_t code = 0;
  for (int b = 1; b <= kMaxBits; ++b) { code = (code + cnt[b - 1]) << 1; next[b] = code; }
  uint32_t enc[kSyms];  // reversed code | length << 16
  uint64_t payload_bits = 0;
  for (int s = 0; s < kSyms; ++s) {
    enc[s] = 0;
    if (len[s]) {
      enc[s] = reverse_bits(next[len[s]]++, len[s]) | ((uint32_t)len[s] << 16);
      payload_bits += (uint64_t)freq[s] * len[s];
    }
  }
  const int64_t huff_bytes = 2 + (int64_t)((3 + 5 + 5 + 4 + 19 * 3 + 258 * 4 + payload_bits + 7) / 8) + 4;
  if (huff_bytes >= stored_size(n)) return write_stored(src, n, dst);

  uint8_t* p = dst;
  *p++ = 0x78; *p++ = 0x01;
  BitWriter bw(p);
  bw.put(1, 1);    // BFINAL
  bw.put(2, 2);    // BTYPE = dynamic Huffman
  bw.put(0, 5);    // HLIT: 257 literal / length codes
  bw.put(0, 5);    // HDIST: 1 distance code (of length 0: the block holds literals only)
  bw.put(15, 4);   // HCLEN: all 19 code-length code lengths follow
  // code-length alphabet: symbols 0..15 (a length sent as itself) get 4 bits each, 16..18 (run lengths) are unused.
  // Order of transmission: 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15.
  for (int k = 0; k < 19; ++k) bw.put(k < 3 ? 0 : 4, 3);
  for (int s = 0; s < kSyms; ++s) bw.put(reverse_bits(len[s], 4), 4);
  bw.put(reverse_bits(0, 4), 4);  // the one distance code: length 0
  // literals, four at a time when no code is longer than 14 bits (four codes on top of fewer than 8 pending bits
  // then fit 64 bits), else three at a time; whole bytes leave with one unaligned 8-byte store (the bound leaves
  // room for the bytes past the end)
  {
    int max_len = 0;
    for (int s = 0; s < kSyms; ++s) max_len = std::max<int>(max_len, len[s]);
    uint8_t* q = bw.finish_to_byte_state(&bw.bits, &bw.n);
    uint64_t acc = bw.bits;
    int nb = bw.n;  // < 8
    auto flush = [&]() {
      std::memcpy(q, &acc, 8);
      const int bytes = nb >> 3;
      q += bytes;
      acc >>= 8 * bytes;  // bytes <= 7
      nb &= 7;
    };
    i = 0;
    if (max_len <= 14) {
      for (; i + 3 < n; i += 4) {
        // the four codes are put together first (no dependence on the bits pending), so that the chain from one
        // group of four to the next is one shift, one or and the flush
        const uint32_t e0 = enc[src[i]], e1 = enc[src[i + 1]], e2 = enc[src[i + 2]], e3 = enc[src[i + 3]];
        const int l0 = (int)(e0 >> 16), l1 = (int)(e1 >> 16), l2 = (int)(e2 >> 16), l3 = (int)(e3 >> 16);
        const uint64_t c01 = (uint64_t)(e0 & 0xffffu) | ((uint64_t)(e1 & 0xffffu) << l0);
        const uint64_t c23 = (uint64_t)(e2 & 0xffffu) | ((uint64_t)(e3 & 0xffffu) << l2);
        acc |= (c01 | (c23 << (l0 + l1))) << nb;
        nb += l0 + l1 + l2 + l3;
        flush();
      }
    } else {
      for (; i + 2 < n; i += 3) {
        const uint32_t e0 = enc[src[i]], e1 = enc[src[i + 1]], e2 = enc[src[i + 2]];
        acc |= (uint64_t)(e0 & 0xffffu) << nb; nb += (int)(e0 >> 16);
        acc |= (uint64_t)(e1 & 0xffffu) << nb; nb += (int)(e1 >> 16);
        acc |= (uint64_t)(e2 & 0xffffu) << nb; nb += (int)(e2 >> 16);
        flush();
      }
    }
    bw.p = q;
    bw.bits = acc;
    bw.n = nb;
  }
  for (; i < n; ++i) { const uint32_t e0 = enc[src[i]]; bw.put(e0 & 0xffffu, (int)(e0 >> 16)); }
  bw.put(enc[256] & 0xffffu, (int)(enc[256] >> 16));
  p = bw.finish();
  const uint32_t ad = adler32(src, n);
  *p++ = (uint8_t)(ad >> 24); *p++ = (uint8_t)(ad >> 16); *p++ = (uint8_t)(ad >> 8); *p++ = (uint8_t)ad;
  return p - dst;
}

struct ReadView {
  int64_t ev_rows, ev_off, align_n, align_off, seg_n, seg_off;
  bool ok;
};

ReadView view_of(const nrq_batch* b, const nrq_results* res, int r) {
  ReadView v;
  v.ok = res->status[r] == NRQ_ST_OK;
  v.align_off = b->align_off[r];
  v.align_n = b->align_off[r + 1] - b->align_off[r];
  v.ev_off = v.align_off;
  v.ev_rows = v.ok ? res->n_bases[r] : 0;
  v.seg_off = b->starts_off[r];
  v.seg_n = b->starts_off[r + 1] - b->starts_off[r];
  return v;
}

constexpr int kRow = 25;  // norm_mean f64, norm_stdev f64, start u32, length u32, base S1 (resquiggle.py:69-70)

}  // namespace

extern "C" {

int64_t nrqio_deflate_bound(int64_t n) { return n < 0 ? -1 : stored_size(n) + 16; }

uint32_t nrqio_adler32(const uint8_t* src, int64_t n) { return adler32(src, n < 0 ? 0 : n); }

int64_t nrqio_deflate(const uint8_t* src, int64_t n, uint8_t* dst, int64_t cap) {
  if (n < 0 || !dst || (n > 0 && !src) || cap < nrqio_deflate_bound(n)) return -1;
  if (n < 64) return write_stored(src, n, dst);
  return write_huffman(src, n, dst);
}

int64_t nrqio_pack_corrected_bound(const nrq_batch* b, const nrq_results* res, int32_t first, int32_t count) {
  if (!b || !res || first < 0 || count < 0 || first + count > b->n_reads) return -1;
  int64_t total = 0;
  for (int i = 0; i < count; ++i) {
    const ReadView v = view_of(b, res, first + i);
    if (!v.ok) continue;
    total += nrqio_deflate_bound(kRow * v.ev_rows) + 2 * nrqio_deflate_bound(v.align_n) +
             nrqio_deflate_bound(8 * v.seg_n);
  }
  return total;
}

int nrqio_pack_corrected(const nrq_batch* b, const nrq_results* res, int32_t first, int32_t count,
                         int32_t n_threads, uint8_t* out, int64_t out_cap, int64_t* chunk_off,
                         int64_t* chunk_len) {
  if (!b || !res || !chunk_off || !chunk_len || first < 0 || count < 0 || first + count > b->n_reads) return -1;
  if (count == 0) return 0;
  if (!out || !res->status || !res->n_bases || !res->norm_mean || !res->norm_stdev || !res->ev_length ||
      !res->ev_base || !b->read_align || !b->genome_align || !b->starts)
    return -1;
  // where each chunk goes: worst-case sizes, so that the reads can be packed independently
  int64_t pos = 0;
  for (int i = 0; i < count; ++i) {
    const ReadView v = view_of(b, res, first + i);
    const int64_t raw[4] = {kRow * v.ev_rows, v.align_n, v.align_n, 8 * v.seg_n};
    for (int j = 0; j < 4; ++j) {
      chunk_off[4 * i + j] = pos;
      chunk_len[4 * i + j] = 0;
      if (v.ok) pos += nrqio_deflate_bound(raw[j]);
    }
  }
  if (pos > out_cap) return -1;
  std::atomic<int> next{0};
  auto work = [&]() {
    std::vector<uint8_t> rows;
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= count) break;
      const ReadView v = view_of(b, res, first + i);
      if (!v.ok) continue;
      rows.resize((size_t)std::max<int64_t>(kRow * v.ev_rows, 8 * v.seg_n));
      uint8_t* q = rows.data();
      // `start` is the running sum of `length` from the read's first boundary (new_segs[0] = starts[0]), so a caller
      // that wants fewer bytes back from the device leaves ev_start out (NULL)
      uint32_t run = v.seg_n > 0 ? (uint32_t)b->starts[v.seg_off] : 0u;
      for (int64_t k = 0; k < v.ev_rows; ++k, q += kRow) {
        const uint32_t len = res->ev_length[v.ev_off + k];
        const uint32_t start = res->ev_start ? res->ev_start[v.ev_off + k] : run;
        run = start + len;
        std::memcpy(q, res->norm_mean + v.ev_off + k, 8);
        std::memcpy(q + 8, res->norm_stdev + v.ev_off + k, 8);
        std::memcpy(q + 16, &start, 4);
        std::memcpy(q + 20, &len, 4);
        q[24] = res->ev_base[v.ev_off + k];
      }
      chunk_len[4 * i] = nrqio_deflate(rows.data(), kRow * v.ev_rows, out + chunk_off[4 * i], out_cap - chunk_off[4 * i]);
      chunk_len[4 * i + 1] = nrqio_deflate(b->read_align + v.align_off, v.align_n, out + chunk_off[4 * i + 1],
                                           out_cap - chunk_off[4 * i + 1]);
      chunk_len[4 * i + 2] = nrqio_deflate(b->genome_align + v.align_off, v.align_n, out + chunk_off[4 * i + 2],
                                           out_cap - chunk_off[4 * i + 2]);
      int64_t* seg = reinterpret_cast<int64_t*>(rows.data());
      for (int64_t k = 0; k < v.seg_n; ++k) seg[k] = b->starts[v.seg_off + k];
      chunk_len[4 * i + 3] = nrqio_deflate(rows.data(), 8 * v.seg_n, out + chunk_off[4 * i + 3],
                                           out_cap - chunk_off[4 * i + 3]);
    }
  };
  const int nt = std::min<int>(std::max(n_threads, 1), count);
  if (nt <= 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) pool.emplace_back(work);
    for (auto& t : pool) t.join();
  }
  for (int i = 0; i < 4 * count; ++i)
    if (chunk_len[i] < 0) return -1;
  return 0;
}

}  // extern "C"

// ===================================================================================================== front end
namespace {

struct CigarOp { int64_t n; char op; };

inline bool is_op(char c) { return std::strchr("MIDNSHP=X", c) != nullptr && c != 0; }
inline bool consumes_ref(char c) { return c == 'M' || c == 'D' || c == 'N' || c == '=' || c == 'X'; }
inline bool consumes_read(char c) { return c == 'M' || c == 'I' || c == 'P' || c == '=' || c == 'X'; }
inline bool aligned(char c) { return c == 'M' || c == '=' || c == 'X'; }
inline bool read_only(char c) { return c == 'I' || c == 'P'; }

// re.findall(r'(\d+)([MIDNSHP=X])'): every digit run that is directly followed by an operation character
void parse_cigar(const char* c, std::vector<CigarOp>& ops) {
  ops.clear();
  for (const char* p = c; *p;) {
    if (*p < '0' || *p > '9') { ++p; continue; }
    int64_t n = 0;
    const char* q = p;
    while (*q >= '0' && *q <= '9') { n = n * 10 + (*q - '0'); ++q; }
    if (is_op(*q)) { ops.push_back({n, *q}); p = q + 1; }
    else p = q;
  }
}

inline char comp(char b) {  // nanoraw_helper.py:39-45: anything unknown becomes N
  switch (b) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    case '-': return '-';
    default: return 'N';
  }
}

void rev_comp(std::string& s) {
  std::reverse(s.begin(), s.end());
  for (char& ch : s) ch = comp(ch);
}

// Python slice s[a:b] with non-negative a, b clamped to the string
inline std::string slice(const std::string& s, int64_t a, int64_t b) {
  const int64_t n = (int64_t)s.size();
  a = std::max<int64_t>(0, std::min(a, n));
  b = std::max<int64_t>(a, std::min(b, n));
  return s.substr((size_t)a, (size_t)(b - a));
}
// s[n:] and s[:-n] for n > 0 (n == 0 never reaches the callers below the way Python's s[:-0] would matter)
inline void drop_front(std::string& s, int64_t n) { s = slice(s, n, (int64_t)s.size()); }
inline void drop_back(std::string& s, int64_t n) { s = n > 0 ? slice(s, 0, (int64_t)s.size() - n) : std::string(); }

struct SamRows {
  std::string read_row, genome_row;
  int64_t clip[2];
};

int sam_rows(const char* cigar, int32_t flag, int64_t pos, const char* seq, int64_t seq_len, const char* contig,
             int64_t contig_len, std::vector<CigarOp>& ops, SamRows& out) {
  parse_cigar(cigar, ops);
  if (ops.empty()) return NRQIO_FE_BAD_CIGAR;
  const bool reverse = (flag & 0x10) != 0;
  std::string read(seq, (size_t)seq_len);
  if (reverse) {
    std::reverse(ops.begin(), ops.end());
    rev_comp(read);
  }
  out.clip[0] = out.clip[1] = 0;
  // hard clips first, both ends (each end looked at once, in this order, as the reference does)
  if (ops.front().op == 'H') { out.clip[0] += ops.front().n; ops.erase(ops.begin()); }
  if (ops.empty()) return NRQIO_FE_BAD_INPUT;
  if (ops.back().op == 'H') { out.clip[1] += ops.back().n; ops.pop_back(); }
  if (ops.empty()) return NRQIO_FE_BAD_INPUT;
  if (ops.front().op == 'S') {
    const int64_t n = ops.front().n;
    ops.erase(ops.begin());
    out.clip[0] += n;
    drop_front(read, n);
  }
  if (ops.empty()) return NRQIO_FE_BAD_INPUT;
  if (ops.back().op == 'S') {
    const int64_t n = ops.back().n;
    ops.pop_back();
    out.clip[1] += n;
    drop_back(read, n);
  }
  if (ops.empty()) return NRQIO_FE_BAD_INPUT;
  int64_t ref_len = 0;
  for (const CigarOp& o : ops)
    if (consumes_ref(o.op)) ref_len += o.n;
  std::string ref;
  {  // genome_index[rName][pos - 1 : pos - 1 + ref_len] (a negative start would wrap in Python: not produced by a mapper)
    const int64_t a = std::max<int64_t>(0, std::min(pos - 1, contig_len));
    const int64_t b = std::max<int64_t>(a, std::min(pos - 1 + ref_len, contig_len));
    ref.assign(contig + a, (size_t)(b - a));
  }
  if (reverse) rev_comp(ref);
  while (!aligned(ops.front().op)) {
    const CigarOp o = ops.front();
    ops.erase(ops.begin());
    if (read_only(o.op)) drop_front(ref, o.n);
    else { drop_front(read, o.n); out.clip[0] += o.n; }
    if (ops.empty()) return NRQIO_FE_BAD_INPUT;
  }
  while (!aligned(ops.back().op)) {
    const int64_t first_len = ops.front().n;  // (:636 adds cigar[0][0], not the stripped operation's own length)
    const CigarOp o = ops.back();
    ops.pop_back();
    if (read_only(o.op)) drop_back(ref, o.n);
    else { drop_back(read, o.n); out.clip[1] += first_len; }
    if (ops.empty()) return NRQIO_FE_BAD_INPUT;
  }
  int64_t need_read = 0;
  for (const CigarOp& o : ops)
    if (consumes_read(o.op)) need_read += o.n;
  if ((int64_t)read.size() != need_read) return NRQIO_FE_SEQ_CIGAR_DISAGREE;
  out.read_row.clear();
  out.genome_row.clear();
  int64_t ri = 0, gi = 0;
  for (const CigarOp& o : ops) {
    // zip(...) of the reference (:645-656): as many pairs as the shorter side has (the reference sequence runs
    // short when a leading / trailing insertion was trimmed from it, :626, :632)
    const std::string rp = slice(read, ri, ri + o.n), gp = slice(ref, gi, gi + o.n);
    if (aligned(o.op)) {
      const size_t pairs = std::min(rp.size(), gp.size());
      out.read_row.append(rp, 0, pairs);
      out.genome_row.append(gp, 0, pairs);
      ri += o.n;
      gi += o.n;
    } else if (read_only(o.op)) {
      out.read_row += rp;
      out.genome_row.append(rp.size(), '-');
      ri += o.n;
    } else {
      out.read_row.append(gp.size(), '-');
      out.genome_row += gp;
      gi += o.n;
    }
  }
  return NRQIO_FE_OK;
}

void tally(const std::string& r, const std::string& g, int64_t* t) {
  int64_t ins = 0, del = 0, match = 0, both = 0;
  const size_t n = std::min(r.size(), g.size());
  for (size_t i = 0; i < n; ++i) {
    const bool rg = r[i] == '-', gg = g[i] == '-';
    del += rg;
    ins += gg && !rg;
    if (!rg && !gg) { ++both; match += r[i] == g[i]; }
  }
  t[0] = ins; t[1] = del; t[2] = match; t[3] = both - match;
}

inline double col(const void* p, int32_t is_f32, int64_t i) {
  return is_f32 ? (double)static_cast<const float*>(p)[i] : static_cast<const double*>(p)[i];
}
inline bool moved(const void* p, int32_t size, int64_t i) {
  switch (size) {
    case 1: return static_cast<const int8_t*>(p)[i] > 0;
    case 2: return static_cast<const int16_t*>(p)[i] > 0;
    case 4: return static_cast<const int32_t*>(p)[i] > 0;
    default: return static_cast<const int64_t*>(p)[i] > 0;
  }
}

}  // namespace

extern "C" {

const char* nrqio_fe_message(int32_t st) {
  switch (st) {
    case NRQIO_FE_OK: return "";
    case NRQIO_FE_EVENTS_BEFORE_RAW: return "Events appear to start before the raw signal.";
    case NRQIO_FE_TOO_FEW_SEGMENTS: return "One or no segments or signal present in read.";
    case NRQIO_FE_ZERO_LENGTH: return "Zero length event present in input data.";
    case NRQIO_FE_ALL_STAYS: return "Read is composed entirely of stay model states and cannot be processed";
    case NRQIO_FE_BAD_CIGAR: return "Invalid cigar string produced.";
    case NRQIO_FE_SEQ_CIGAR_DISAGREE: return "Read sequence from SAM and cooresponding cigar string do not agree.";
    case NRQIO_FE_CAPACITY: return "Front-end output arrays too small.";
    default: return "Inconsistent front-end input.";
  }
}

int nrqio_events_to_starts(const void* ev_start, int32_t start_is_f32, const void* ev_length, int32_t length_is_f32,
                           const uint8_t* model_state, int32_t state_len, const void* move, int32_t move_size,
                           int64_t n, int32_t albacore_after_1_0, int64_t sampling_rate, uint64_t raw_start_time,
                           int64_t* starts, uint8_t* bases, int64_t* n_starts, int64_t* read_start) {
  if (!ev_start || !ev_length || !model_state || !move || !starts || !bases || !n_starts || !read_start ||
      n < 1 || state_len < 3)
    return NRQIO_FE_BAD_INPUT;
  const double rate = (double)sampling_rate;
  // abs_event_start = round(Events.start[0] * rate) (half to even, as np.round) as uint64 (:851)
  const double first_d = std::nearbyint(col(ev_start, start_is_f32, 0) * rate);
  const uint64_t first_sample = first_d > 0 ? (uint64_t)first_d : 0ull;
  int64_t rstart;
  if (first_sample >= raw_start_time) rstart = (int64_t)(first_sample - raw_start_time);
  // float noise in the stored start time (:852-855).  start_time is a uint64 attribute and `start_time - 2` is taken
  // in that type: for start_time < 2 it wraps around and the test fails, as in the reference
  else if (first_sample >= raw_start_time - 2ull) rstart = 0;
  else return NRQIO_FE_EVENTS_BEFORE_RAW;
  std::vector<int64_t> all((size_t)n + 1);
  if (albacore_after_1_0) {
    // round(length * rate), the product taken in the column's own precision (numpy: float32 * int stays float32)
    int64_t run = 0;
    all[0] = 0;
    for (int64_t i = 0; i < n; ++i) {
      int64_t obs;
      if (length_is_f32) obs = (int64_t)std::nearbyintf(static_cast<const float*>(ev_length)[i] * (float)sampling_rate);
      else obs = (int64_t)std::nearbyint(static_cast<const double*>(ev_length)[i] * rate);
      run += obs;
      all[(size_t)i + 1] = run;
    }
  } else {
    // round(append(start, start[-1] + length[-1]) * rate) - abs_event_start (:878-887); the last edge is formed
    // in the wider of the two column types
    for (int64_t i = 0; i < n; ++i)
      all[(size_t)i] = (int64_t)std::nearbyint(col(ev_start, start_is_f32, i) * rate) - (int64_t)first_sample;
    double last_edge;
    if (start_is_f32 && length_is_f32)
      last_edge = (double)(static_cast<const float*>(ev_start)[n - 1] + static_cast<const float*>(ev_length)[n - 1]);
    else
      last_edge = col(ev_start, start_is_f32, n - 1) + col(ev_length, length_is_f32, n - 1);
    all[(size_t)n] = (int64_t)std::nearbyint(last_edge * rate) - (int64_t)first_sample;
  }
  if (n <= 1) return NRQIO_FE_TOO_FEW_SEGMENTS;  // min(len(starts), len(basecalls)) <= 1 (:891-894)
  for (int64_t i = 0; i < n; ++i)
    if (all[(size_t)i + 1] - all[(size_t)i] < 1) return NRQIO_FE_ZERO_LENGTH;
  // ---- fix_stay_states (:767-805): moves = move[1:] > 0
  const int64_t nm = n - 1;
  int64_t first = -1, last = -1;
  for (int64_t j = 0; j < nm; ++j)
    if (moved(move, move_size, j + 1)) { if (first < 0) first = j; last = j; }
  if (first < 0 || first >= nm - 1) return NRQIO_FE_ALL_STAYS;
  const int64_t cut = nm - 1 - last;               // trailing stays
  const int64_t ns = (n + 1 - cut) - first;         // len(starts[first : len - cut]) = L + 2
  const int64_t L = last - first + 1;               // len(moves[first : last + 1])
  const int64_t shift = first > 0 ? all[(size_t)first] : 0;
  if (first > 0) rstart += shift;
  // keep = [True] + moves[:-1] + [True] + moves[-1:]
  int64_t w = 0;
  for (int64_t j = 0; j < ns; ++j) {
    bool keep;
    if (j == 0 || j == L) keep = true;
    else if (j < L) keep = moved(move, move_size, first + (j - 1) + 1);
    else keep = moved(move, move_size, first + (L - 1) + 1);
    if (!keep) continue;
    starts[w] = all[(size_t)(first + j)] - shift;
    if (j < ns - 1) {  // calls[keep[:-1]]: the call of event first + j
      bases[w] = model_state[(size_t)(first + j) * (size_t)state_len + 2];
    }
    ++w;
  }
  *n_starts = w;
  *read_start = rstart;
  return NRQIO_FE_OK;
}

int nrqio_sam_to_rows(const char* cigar, int32_t flag, int64_t pos, const char* seq, int64_t seq_len,
                      const char* contig, int64_t contig_len, char* read_row, char* genome_row, int64_t row_cap,
                      int64_t* n_cols, int64_t* clip, int64_t* tallies) {
  if (!cigar || !seq || !contig || !read_row || !genome_row || !n_cols || !clip || seq_len < 0 || contig_len < 0)
    return NRQIO_FE_BAD_INPUT;
  std::vector<CigarOp> ops;
  SamRows rows;
  const int st = sam_rows(cigar, flag, pos, seq, seq_len, contig, contig_len, ops, rows);
  if (st != NRQIO_FE_OK) return st;
  if ((int64_t)rows.read_row.size() > row_cap || (int64_t)rows.genome_row.size() > row_cap) return NRQIO_FE_CAPACITY;
  std::memcpy(read_row, rows.read_row.data(), rows.read_row.size());
  std::memcpy(genome_row, rows.genome_row.data(), rows.genome_row.size());
  *n_cols = (int64_t)rows.read_row.size();
  clip[0] = rows.clip[0];
  clip[1] = rows.clip[1];
  if (tallies) tally(rows.read_row, rows.genome_row, tallies);
  return NRQIO_FE_OK;
}

int nrqio_clip_starts(int64_t* starts, int64_t* n_starts, int64_t* read_start, int64_t clip_start, int64_t clip_end) {
  if (!starts || !n_starts || !read_start || clip_start < 0 || clip_end < 0) return NRQIO_FE_BAD_INPUT;
  int64_t n = *n_starts;
  if (clip_start > 0) {
    if (clip_start >= n) return NRQIO_FE_BAD_INPUT;  // (an IndexError in the reference)
    const int64_t shift = starts[clip_start];
    for (int64_t i = clip_start; i < n; ++i) starts[i - clip_start] = starts[i] - shift;
    n -= clip_start;
    *read_start += shift;
  }
  if (clip_end > 0) n = std::max<int64_t>(0, n - clip_end);
  *n_starts = n;
  return NRQIO_FE_OK;
}

int nrqio_frontend_batch(const nrqio_fe_read* reads, int32_t n_reads, int32_t n_threads, int32_t* status,
                         int32_t* read_start, int32_t* starts, int64_t starts_cap, int64_t* starts_off,
                         uint8_t* read_align, uint8_t* genome_align, int64_t align_cap, int64_t* align_off,
                         int64_t* clip, int64_t* tallies) {
  if (n_reads < 0 || !status || !read_start || !starts_off || !align_off || !clip || !tallies ||
      (n_reads > 0 && (!reads || !starts || !read_align || !genome_align)))
    return NRQIO_FE_BAD_INPUT;
  std::vector<SamRows> rows((size_t)n_reads);
  std::atomic<int> next{0};
  auto parse = [&]() {
    std::vector<CigarOp> ops;
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n_reads) break;
      const nrqio_fe_read& rd = reads[i];
      int st = NRQIO_FE_BAD_INPUT;
      if (rd.cigar && rd.seq && rd.contig && rd.starts && rd.n_starts >= 0)
        st = sam_rows(rd.cigar, rd.flag, rd.pos, rd.seq, rd.seq_len, rd.contig, rd.contig_len, ops, rows[(size_t)i]);
      if (st == NRQIO_FE_OK && rows[(size_t)i].clip[0] > 0 && rows[(size_t)i].clip[0] >= rd.n_starts)
        st = NRQIO_FE_BAD_INPUT;
      status[i] = st;
      if (st != NRQIO_FE_OK) { rows[(size_t)i].read_row.clear(); rows[(size_t)i].genome_row.clear(); }
    }
  };
  auto run = [&](auto fn) {
    const int nt = std::min<int>(std::max(n_threads, 1), std::max(n_reads, 1));
    next = 0;
    if (nt <= 1) { fn(); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) pool.emplace_back(fn);
    for (auto& t : pool) t.join();
  };
  run(parse);
  int64_t so = 0, ao = 0;
  for (int i = 0; i < n_reads; ++i) {
    starts_off[i] = so;
    align_off[i] = ao;
    if (status[i] != NRQIO_FE_OK) continue;
    const SamRows& r = rows[(size_t)i];
    so += std::max<int64_t>(0, reads[i].n_starts - r.clip[0] - r.clip[1]);
    ao += (int64_t)r.read_row.size();
  }
  starts_off[n_reads] = so;
  align_off[n_reads] = ao;
  if (so > starts_cap || ao > align_cap) return NRQIO_FE_CAPACITY;
  auto fill = [&]() {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n_reads) break;
      clip[2 * i] = clip[2 * i + 1] = 0;
      tallies[4 * i] = tallies[4 * i + 1] = tallies[4 * i + 2] = tallies[4 * i + 3] = 0;
      read_start[i] = 0;
      if (status[i] != NRQIO_FE_OK) continue;
      const SamRows& r = rows[(size_t)i];
      const nrqio_fe_read& rd = reads[i];
      clip[2 * i] = r.clip[0];
      clip[2 * i + 1] = r.clip[1];
      tally(r.read_row, r.genome_row, tallies + 4 * i);
      std::memcpy(read_align + align_off[i], r.read_row.data(), r.read_row.size());
      std::memcpy(genome_align + align_off[i], r.genome_row.data(), r.genome_row.size());
      // fix_raw_starts_for_clipped_bases (:447-460), written as int32 straight into the batch
      const int64_t c0 = r.clip[0];
      const int64_t shift = c0 > 0 ? rd.starts[c0] : 0;
      const int64_t n_out = starts_off[i + 1] - starts_off[i];
      int32_t* o = starts + starts_off[i];
      for (int64_t k = 0; k < n_out; ++k) o[k] = (int32_t)(rd.starts[c0 + k] - shift);
      read_start[i] = (int32_t)(rd.read_start + shift);
    }
  };
  run(fill);
  return NRQIO_FE_OK;
}

}  // extern "C"

