// libnrqio.so (include/nrq_io.h): host-side byte work of the FAST5 write-back -- a deflate encoder without match
// search and the packing of a result batch into the gzip chunks write_new_fast5_group stores
// (nanoraw/resquiggle.py:112-124).  Plain C++17, no CUDA, no zlib.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/nrq_io.h"

namespace {

constexpr int kSyms = 257;     // 256 literals + end of block
constexpr int kMaxBits = 15;   // deflate's limit on a code length
constexpr int64_t kStoredMax = 65535;

uint32_t adler32(const uint8_t* p, int64_t n) {
  uint32_t a = 1, b = 0;
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
  for (i = 0; i + 1 < n; i += 2) {
    const uint32_t e0 = enc[src[i]], e1 = enc[src[i + 1]];
    bw.put(e0 & 0xffffu, (int)(e0 >> 16));
    bw.put(e1 & 0xffffu, (int)(e1 >> 16));
  }
  if (i < n) { const uint32_t e0 = enc[src[i]]; bw.put(e0 & 0xffffu, (int)(e0 >> 16)); }
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

int64_t nrqio_deflate_bound(int64_t n) { return n < 0 ? -1 : stored_size(n) + 8; }

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
  if (!out || !res->status || !res->n_bases || !res->norm_mean || !res->norm_stdev || !res->ev_start ||
      !res->ev_length || !res->ev_base || !b->read_align || !b->genome_align || !b->starts)
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
      for (int64_t k = 0; k < v.ev_rows; ++k, q += kRow) {
        std::memcpy(q, res->norm_mean + v.ev_off + k, 8);
        std::memcpy(q + 8, res->norm_stdev + v.ev_off + k, 8);
        std::memcpy(q + 16, res->ev_start + v.ev_off + k, 4);
        std::memcpy(q + 20, res->ev_length + v.ev_off + k, 4);
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
