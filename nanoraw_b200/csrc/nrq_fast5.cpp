// nrq_fast5.cpp -- the FAST5 (HDF5) file work of the re-squiggle worker, without libhdf5 (part of libnrqio.so).
//
// What the worker does to a file (nanoraw/resquiggle.py:327-347 read, :80-135 write):
//   read   /Raw/Reads/<the read>/Signal (int16, usually one or more gzip chunks) and the channel_id attributes
//   write  /Analyses/<corrected group>/<subgroup> {attributes, Alignment/{3 datasets, attributes}, Events}
// Both are implemented here from the HDF5 File Format Specification (version 3.0) for the on-disk forms FAST5 files
// come in: superblock 0-3, object headers v1 / v2, symbol-table groups and compact link messages, contiguous /
// compact / chunked (v1 B-tree or single chunk) datasets, deflate + shuffle + fletcher32 filters, attribute
// messages v1-v3.  Anything outside that slice is reported as NRQIO_F5_UNSUPPORTED and the caller takes the
// Python reader / writer (nanoraw_b200/hdf5_min.py), whose forms these functions mirror byte for byte on the
// write side (v1 object headers, symbol-table groups, one gzip chunk under a v1 B-tree, v1 attribute messages,
// variable-length strings in a global heap collection).
//
// The write is an append: the new objects go behind the end-of-file address, the group that receives the link gets
// a new local heap / symbol node / B-tree there as well, then the end-of-file address and the two places that
// point at the group's B-tree and heap (its symbol-table message, and the cached copy in its parent's symbol table
// entry) are patched.  The file is valid at every point in between.  (hdf5_min.py re-serialises the whole file
// instead: 1.7 ms of Python per read, which was what bounded the FAST5-inclusive throughput.)
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nrq_io.h"

namespace {

constexpr uint64_t kUndef = ~0ull;

struct Fail {
  int code;
};
#define F5_REQUIRE(cond, code_) \
  do {                          \
    if (!(cond)) throw Fail{code_}; \
  } while (0)

inline uint64_t rd_le(const uint8_t* p, int n) {
  uint64_t v = 0;
  for (int i = n - 1; i >= 0; --i) v = (v << 8) | p[i];
  return v;
}

template <class Fn>
void parallel_for(int n, int n_threads, Fn fn) {
  if (n_threads <= 1 || n <= 1) {
    for (int i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<int> next(0);
  auto body = [&]() {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n) break;
      fn(i);
    }
  };
  std::vector<std::thread> pool;
  const int nt = std::min(n_threads, n);
  for (int t = 1; t < nt; ++t) pool.emplace_back(body);
  body();
  for (auto& th : pool) th.join();
}

// ------------------------------------------------------------------------------------------------ file image
struct File {
  std::vector<uint8_t> d;
  uint64_t base = 0;  // addresses in the file are relative to this (user block)
  int osz = 8, lsz = 8;
  int sb_ver = 0;
  uint64_t sb_off = 0;
  uint64_t root = 0;       // root group's object header (absolute)
  uint64_t eof_field = 0;  // where the end-of-file address is stored
  int leaf_k = 4, int_k = 16, chunk_k = 32;

  const uint8_t* at(uint64_t off, uint64_t n) const {
    F5_REQUIRE(off <= d.size() && n <= d.size() - off, NRQIO_F5_FORMAT);
    return d.data() + off;
  }
  uint64_t u(uint64_t off, int n) const { return rd_le(at(off, (uint64_t)n), n); }
  uint64_t undef() const { return osz >= 8 ? kUndef : ((1ull << (8 * osz)) - 1); }
  uint64_t addr(uint64_t off) const {
    const uint64_t v = u(off, osz);
    return v == undef() ? kUndef : v + base;
  }
  uint64_t len(uint64_t off) const { return u(off, lsz); }
};

void load_file(const char* path, int flags, File& f, int* fd_out) {
  const int fd = ::open(path, flags);
  F5_REQUIRE(fd >= 0, NRQIO_F5_IO);
  struct stat st;
  if (fstat(fd, &st) != 0) {
    ::close(fd);
    throw Fail{NRQIO_F5_IO};
  }
  f.d.resize((size_t)st.st_size);
  size_t got = 0;
  while (got < f.d.size()) {
    const ssize_t k = pread(fd, f.d.data() + got, f.d.size() - got, (off_t)got);
    if (k <= 0) {
      ::close(fd);
      throw Fail{NRQIO_F5_IO};
    }
    got += (size_t)k;
  }
  if (fd_out) *fd_out = fd;
  else ::close(fd);
}

void parse_superblock(File& f) {
  static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
  uint64_t pos = 0;
  for (;;) {
    F5_REQUIRE(pos + 8 <= f.d.size(), NRQIO_F5_FORMAT);
    if (std::memcmp(f.d.data() + pos, sig, 8) == 0) break;
    pos = pos == 0 ? 512 : pos * 2;
  }
  f.sb_off = pos;
  const int ver = (int)f.u(pos + 8, 1);
  f.sb_ver = ver;
  if (ver == 0 || ver == 1) {
    f.osz = (int)f.u(pos + 13, 1);
    f.lsz = (int)f.u(pos + 14, 1);
    F5_REQUIRE((f.osz == 2 || f.osz == 4 || f.osz == 8) && (f.lsz == 2 || f.lsz == 4 || f.lsz == 8), NRQIO_F5_FORMAT);
    f.leaf_k = (int)f.u(pos + 16, 2);
    f.int_k = (int)f.u(pos + 18, 2);
    if (ver == 1) f.chunk_k = (int)f.u(pos + 24, 2);
    uint64_t p = pos + 24 + (ver == 1 ? 4 : 0);
    f.base = f.u(p, f.osz);
    f.eof_field = p + 2 * (uint64_t)f.osz;
    p += 4 * (uint64_t)f.osz;
    f.root = f.addr(p + f.osz);
  } else if (ver == 2 || ver == 3) {
    f.osz = (int)f.u(pos + 9, 1);
    f.lsz = (int)f.u(pos + 10, 1);
    F5_REQUIRE((f.osz == 2 || f.osz == 4 || f.osz == 8) && (f.lsz == 2 || f.lsz == 4 || f.lsz == 8), NRQIO_F5_FORMAT);
    f.base = f.u(pos + 12, f.osz);
    f.eof_field = 0;  // checksummed superblock: not patched here
    f.root = f.addr(pos + 12 + 3 * (uint64_t)f.osz);
  } else {
    throw Fail{NRQIO_F5_UNSUPPORTED};
  }
  F5_REQUIRE(f.root != kUndef, NRQIO_F5_FORMAT);
}

// ------------------------------------------------------------------------------------------------ object headers
struct Msg {
  int type;
  int flags;
  uint64_t off;  // of the message body in the file image
  uint64_t n;
};

// messages of the object header at addr, continuation blocks followed; *version = 1 or 2
void read_messages(const File& f, uint64_t addr, std::vector<Msg>& out, int* version) {
  out.clear();
  F5_REQUIRE(addr != kUndef, NRQIO_F5_FORMAT);
  std::vector<std::pair<uint64_t, uint64_t>> blocks;
  if (std::memcmp(f.at(addr, 4), "OHDR", 4) == 0) {
    F5_REQUIRE(f.u(addr + 4, 1) == 2, NRQIO_F5_UNSUPPORTED);
    if (version) *version = 2;
    const int flags = (int)f.u(addr + 5, 1);
    uint64_t p = addr + 6;
    if (flags & 0x20) p += 16;
    if (flags & 0x10) p += 4;
    const int nsz = 1 << (flags & 3);
    const uint64_t size0 = f.u(p, nsz);
    p += nsz;
    blocks.emplace_back(p, size0);
    const bool track = (flags & 0x04) != 0;
    for (size_t bi = 0; bi < blocks.size(); ++bi) {
      F5_REQUIRE(blocks.size() < 4096, NRQIO_F5_FORMAT);
      uint64_t q = blocks[bi].first;
      const uint64_t end = q + blocks[bi].second;
      F5_REQUIRE(end <= f.d.size(), NRQIO_F5_FORMAT);
      while (q + 4 <= end) {
        const int mtype = (int)f.u(q, 1);
        const uint64_t msize = f.u(q + 1, 2);
        const int mflags = (int)f.u(q + 3, 1);
        q += 4 + (track ? 2 : 0);
        if (q + msize > end) break;  // (the gap in front of the checksum)
        if (mtype == 0x10) {
          const uint64_t ca = f.addr(q), cl = f.len(q + f.osz);
          F5_REQUIRE(ca != kUndef && cl >= 8 && std::memcmp(f.at(ca, 4), "OCHK", 4) == 0, NRQIO_F5_FORMAT);
          blocks.emplace_back(ca + 4, cl - 8);
        } else if (mtype != 0) {
          out.push_back(Msg{mtype, mflags, q, msize});
        }
        q += msize;
      }
    }
    return;
  }
  F5_REQUIRE(f.u(addr, 1) == 1, NRQIO_F5_UNSUPPORTED);
  if (version) *version = 1;
  int nmsg = (int)f.u(addr + 2, 2);
  blocks.emplace_back(addr + 16, f.u(addr + 8, 4));
  for (size_t bi = 0; bi < blocks.size() && nmsg > 0; ++bi) {
    F5_REQUIRE(blocks.size() < 4096, NRQIO_F5_FORMAT);
    uint64_t q = blocks[bi].first;
    const uint64_t end = q + blocks[bi].second;
    F5_REQUIRE(end <= f.d.size(), NRQIO_F5_FORMAT);
    while (q + 8 <= end && nmsg > 0) {
      const int mtype = (int)f.u(q, 2);
      const uint64_t msize = f.u(q + 2, 2);
      const int mflags = (int)f.u(q + 4, 1);
      F5_REQUIRE(q + 8 + msize <= f.d.size(), NRQIO_F5_FORMAT);
      --nmsg;
      if (mtype == 0x10) {
        const uint64_t ca = f.addr(q + 8), cl = f.len(q + 8 + f.osz);
        F5_REQUIRE(ca != kUndef, NRQIO_F5_FORMAT);
        blocks.emplace_back(ca, cl);
      } else if (mtype != 0) {
        out.push_back(Msg{mtype, mflags, q + 8, msize});
      }
      q += 8 + msize;
    }
  }
}

// ------------------------------------------------------------------------------------------------ groups
struct Entry {
  std::string name;
  uint64_t hdr = kUndef;      // object header (absolute)
  uint32_t cache = 0;         // symbol table entry: cache type
  uint8_t scratch[16] = {0};  // ... and scratch-pad
  uint64_t entry_off = 0;     // of the symbol table entry in the file image (0: a link message)
};

std::string cstr_at(const File& f, uint64_t off) {
  F5_REQUIRE(off < f.d.size(), NRQIO_F5_FORMAT);
  const void* z = std::memchr(f.d.data() + off, 0, f.d.size() - off);
  F5_REQUIRE(z != nullptr, NRQIO_F5_FORMAT);
  return std::string(reinterpret_cast<const char*>(f.d.data() + off), (const char*)z - (const char*)(f.d.data() + off));
}

uint64_t heap_data_addr(const File& f, uint64_t heap) {
  F5_REQUIRE(heap != kUndef && std::memcmp(f.at(heap, 4), "HEAP", 4) == 0, NRQIO_F5_FORMAT);
  return f.addr(heap + 8 + 2 * (uint64_t)f.lsz);
}

void list_symtab(const File& f, uint64_t node, uint64_t heap_data, std::vector<Entry>& out, int depth) {
  F5_REQUIRE(depth < 16 && node != kUndef, NRQIO_F5_FORMAT);
  F5_REQUIRE(std::memcmp(f.at(node, 4), "TREE", 4) == 0 && f.u(node + 4, 1) == 0, NRQIO_F5_FORMAT);
  const int level = (int)f.u(node + 5, 1), n = (int)f.u(node + 6, 2);
  uint64_t p = node + 8 + 2 * (uint64_t)f.osz;
  for (int i = 0; i < n; ++i) {
    const uint64_t child = f.addr(p + f.lsz);
    p += (uint64_t)f.lsz + f.osz;
    if (level > 0) {
      list_symtab(f, child, heap_data, out, depth + 1);
      continue;
    }
    F5_REQUIRE(child != kUndef && std::memcmp(f.at(child, 4), "SNOD", 4) == 0, NRQIO_F5_FORMAT);
    const int ns = (int)f.u(child + 6, 2);
    uint64_t q = child + 8;
    for (int s = 0; s < ns; ++s) {
      Entry e;
      e.name = cstr_at(f, heap_data + f.u(q, f.osz));
      e.hdr = f.addr(q + f.osz);
      e.cache = (uint32_t)f.u(q + 2 * (uint64_t)f.osz, 4);
      F5_REQUIRE(e.cache != 2, NRQIO_F5_UNSUPPORTED);  // soft link
      std::memcpy(e.scratch, f.at(q + 2 * (uint64_t)f.osz + 8, 16), 16);
      e.entry_off = q;
      out.push_back(e);
      q += 2 * (uint64_t)f.osz + 24;
    }
  }
}

// members of the group whose header messages are msgs; false if the object is not a group.
// *stab_msg: index of its symbol-table message (old-style group) or -1
bool list_group(const File& f, const std::vector<Msg>& msgs, std::vector<Entry>& out, int* stab_msg) {
  bool is_group = false;
  out.clear();
  if (stab_msg) *stab_msg = -1;
  for (size_t k = 0; k < msgs.size(); ++k) {
    const Msg& m = msgs[k];
    if (m.type == 0x11) {
      is_group = true;
      F5_REQUIRE(m.n >= 2 * (uint64_t)f.osz, NRQIO_F5_FORMAT);
      if (stab_msg) *stab_msg = (int)k;
      list_symtab(f, f.addr(m.off), heap_data_addr(f, f.addr(m.off + f.osz)), out, 0);
    } else if (m.type == 0x02) {
      is_group = true;
      const int flags = (int)f.u(m.off + 1, 1);
      const uint64_t p = m.off + 2 + ((flags & 1) ? 8 : 0);
      F5_REQUIRE(f.addr(p) == kUndef, NRQIO_F5_UNSUPPORTED);  // dense link storage (fractal heap)
    } else if (m.type == 0x06) {
      is_group = true;
      const int flags = (int)f.u(m.off + 1, 1);
      uint64_t p = m.off + 2;
      int ltype = 0;
      if (flags & 0x08) ltype = (int)f.u(p++, 1);
      if (flags & 0x04) p += 8;
      if (flags & 0x10) p += 1;
      const int nsz = 1 << (flags & 3);
      const uint64_t nlen = f.u(p, nsz);
      p += nsz;
      Entry e;
      e.name.assign(reinterpret_cast<const char*>(f.at(p, nlen)), (size_t)nlen);
      p += nlen;
      F5_REQUIRE(ltype == 0, NRQIO_F5_UNSUPPORTED);  // soft / external link
      e.hdr = f.addr(p);
      out.push_back(e);
    }
  }
  return is_group;
}

bool find_entry(const std::vector<Entry>& es, const char* name, Entry& out) {
  for (const Entry& e : es)
    if (e.name == name) {
      out = e;
      return true;
    }
  return false;
}

// walk `path` ('/'-separated) from the root; msgs / entry describe the last component
void walk(const File& f, const char* const* parts, int n_parts, std::vector<Msg>& msgs, Entry* last, int* version) {
  read_messages(f, f.root, msgs, version);
  std::vector<Entry> es;
  for (int k = 0; k < n_parts; ++k) {
    F5_REQUIRE(list_group(f, msgs, es, nullptr), NRQIO_F5_MISSING);
    Entry e;
    F5_REQUIRE(find_entry(es, parts[k], e), NRQIO_F5_MISSING);
    if (last) *last = e;
    read_messages(f, e.hdr, msgs, version);
  }
}

// ------------------------------------------------------------------------------------------------ datasets
struct Filter {
  int id;
  uint32_t cd0;
  bool has_cd;
};

struct Dataset {
  int rank = -1;
  uint64_t dims[4] = {0, 0, 0, 0};
  int dt_class = -1, dt_size = 0;
  bool dt_signed = false, dt_be = false;
  int layout = -1;  // 0 compact, 1 contiguous, 2 chunked
  uint64_t addr = kUndef;  // contiguous data / chunk B-tree / the single chunk
  uint64_t compact_off = 0, compact_n = 0;
  bool single = false;
  uint64_t single_size = 0;
  uint32_t single_mask = 0;
  uint64_t cdim = 0;  // chunk length (rank 1)
  std::vector<Filter> filters;
};

void parse_datatype(const File& f, uint64_t off, int* cls, int* size, bool* is_signed, bool* be, bool* vlen_str) {
  const int b0 = (int)f.u(off, 1);
  *cls = b0 & 0x0f;
  const int bits0 = (int)f.u(off + 1, 1), bits1 = (int)f.u(off + 2, 1);
  *size = (int)f.u(off + 4, 4);
  *be = (bits0 & 1) != 0;
  *is_signed = (bits0 & 8) != 0;
  if (vlen_str) *vlen_str = (*cls == 9) && ((bits0 & 0x0f) == 1);
  (void)bits1;
}

bool parse_dataspace(const File& f, const Msg& m, int* rank, uint64_t* dims) {  // false: null dataspace
  const int ver = (int)f.u(m.off, 1);
  *rank = (int)f.u(m.off + 1, 1);
  uint64_t p;
  if (ver == 1) p = m.off + 8;
  else if (ver == 2) {
    if (f.u(m.off + 3, 1) == 2) return false;
    p = m.off + 4;
  } else throw Fail{NRQIO_F5_UNSUPPORTED};
  F5_REQUIRE(*rank <= 4, NRQIO_F5_UNSUPPORTED);
  for (int k = 0; k < *rank; ++k) dims[k] = f.u(p + (uint64_t)f.lsz * k, f.lsz);
  return true;
}

bool parse_dataset(const File& f, const std::vector<Msg>& msgs, Dataset& ds) {
  const Msg* layout = nullptr;
  for (const Msg& m : msgs) {
    if ((m.type == 0x01 || m.type == 0x03) && (m.flags & 0x02)) throw Fail{NRQIO_F5_UNSUPPORTED};  // shared message
    if (m.type == 0x01) {
      F5_REQUIRE(parse_dataspace(f, m, &ds.rank, ds.dims), NRQIO_F5_UNSUPPORTED);
    } else if (m.type == 0x03) {
      parse_datatype(f, m.off, &ds.dt_class, &ds.dt_size, &ds.dt_signed, &ds.dt_be, nullptr);
    } else if (m.type == 0x08) {
      layout = &m;
    } else if (m.type == 0x0B) {
      const int ver = (int)f.u(m.off, 1), nf = (int)f.u(m.off + 1, 1);
      uint64_t p = m.off + (ver == 1 ? 8 : 2);
      for (int k = 0; k < nf; ++k) {
        const int fid = (int)f.u(p, 2);
        p += 2;
        uint64_t nlen = 0;
        if (ver == 1 || fid >= 256) {
          nlen = f.u(p, 2);
          p += 2;
        }
        const int ncd = (int)f.u(p + 2, 2);
        p += 4;
        p += ver == 1 ? ((nlen + 7) & ~7ull) : nlen;
        Filter fl{fid, 0, ncd > 0};
        if (ncd > 0) fl.cd0 = (uint32_t)f.u(p, 4);
        p += 4 * (uint64_t)ncd;
        if (ver == 1 && (ncd & 1)) p += 4;
        ds.filters.push_back(fl);
      }
    }
  }
  if (ds.rank < 0 || ds.dt_class < 0 || layout == nullptr) return false;
  const uint64_t lo = layout->off;
  const int ver = (int)f.u(lo, 1);
  if (ver == 1 || ver == 2) {
    const int ndim = (int)f.u(lo + 1, 1), cls = (int)f.u(lo + 2, 1);
    uint64_t p = lo + 8;
    uint64_t a = kUndef;
    if (cls != 0) {
      a = f.addr(p);
      p += f.osz;
    }
    ds.layout = cls;
    if (cls == 2) {
      F5_REQUIRE(ndim == 2, NRQIO_F5_UNSUPPORTED);  // one data dimension + the element size
      ds.addr = a;
      ds.cdim = f.u(p, 4);
    } else if (cls == 1) {
      ds.addr = a;
    } else {
      p += 4 * (uint64_t)ndim;
      ds.compact_n = f.u(p, 4);
      ds.compact_off = p + 4;
    }
    return true;
  }
  F5_REQUIRE(ver == 3 || ver == 4, NRQIO_F5_UNSUPPORTED);
  const int cls = (int)f.u(lo + 1, 1);
  ds.layout = cls;
  if (cls == 0) {
    ds.compact_n = f.u(lo + 2, 2);
    ds.compact_off = lo + 4;
  } else if (cls == 1) {
    ds.addr = f.addr(lo + 2);
  } else if (cls == 2 && ver == 3) {
    const int ndim = (int)f.u(lo + 2, 1);
    F5_REQUIRE(ndim == 2, NRQIO_F5_UNSUPPORTED);
    ds.addr = f.addr(lo + 3);
    ds.cdim = f.u(lo + 3 + f.osz, 4);
  } else if (cls == 2) {
    const int flags = (int)f.u(lo + 2, 1), ndim = (int)f.u(lo + 3, 1), enc = (int)f.u(lo + 4, 1);
    F5_REQUIRE(ndim == 2 && enc >= 1 && enc <= 8, NRQIO_F5_UNSUPPORTED);
    ds.cdim = f.u(lo + 5, enc);
    uint64_t p = lo + 5 + (uint64_t)enc * ndim;
    F5_REQUIRE(f.u(p, 1) == 1, NRQIO_F5_UNSUPPORTED);  // chunk index: only the single chunk
    ++p;
    ds.single = true;
    ds.single_size = ds.dims[0] * (uint64_t)ds.dt_size;
    if (flags & 0x02) {
      ds.single_size = f.len(p);
      ds.single_mask = (uint32_t)f.u(p + f.lsz, 4);
      p += (uint64_t)f.lsz + 4;
    }
    ds.addr = f.addr(p);
  } else {
    throw Fail{NRQIO_F5_UNSUPPORTED};
  }
  return true;
}

struct Chunk {
  uint64_t offset, size, addr;
  uint32_t mask;
};

void list_chunks(const File& f, uint64_t node, std::vector<Chunk>& out, int depth) {
  F5_REQUIRE(depth < 16 && node != kUndef, NRQIO_F5_FORMAT);
  F5_REQUIRE(std::memcmp(f.at(node, 4), "TREE", 4) == 0 && f.u(node + 4, 1) == 1, NRQIO_F5_FORMAT);
  const int level = (int)f.u(node + 5, 1), n = (int)f.u(node + 6, 2);
  uint64_t p = node + 8 + 2 * (uint64_t)f.osz;
  const uint64_t ksz = 8 + 8 * 2;  // rank 1: (offset, 0)
  for (int i = 0; i < n; ++i) {
    Chunk c;
    c.size = f.u(p, 4);
    c.mask = (uint32_t)f.u(p + 4, 4);
    c.offset = f.u(p + 8, 8);
    c.addr = f.addr(p + ksz);
    p += ksz + f.osz;
    if (level > 0) list_chunks(f, c.addr, out, depth + 1);
    else out.push_back(c);
  }
}

void inflate_into(const uint8_t* src, uint64_t n_src, uint8_t* dst, uint64_t n_dst) {
  if (nrqio_inflate(src, (int64_t)n_src, dst, (int64_t)n_dst) == (int64_t)n_dst) return;
  // refused by the fast decoder: zlib has the last word on whether the stream is valid
  uLongf got = (uLongf)n_dst;
  const int rc = uncompress(dst, &got, src, (uLong)n_src);
  F5_REQUIRE(rc == Z_OK && (uint64_t)got == n_dst, NRQIO_F5_FORMAT);
}

// one stored chunk -> cbytes bytes of elements (filters undone in reverse order, as H5Z does)
void unfilter(const File& f, const Dataset& ds, const Chunk& c, uint64_t cbytes, uint8_t* out, std::vector<uint8_t>& t0,
              std::vector<uint8_t>& t1) {
  const uint8_t* cur = f.at(c.addr, c.size);
  uint64_t n = c.size;
  const int nf = (int)ds.filters.size();
  int last = -1;  // the last filter that is actually undone writes into `out`
  for (int k = 0; k < nf; ++k)
    if (!(c.mask & (1u << k)) && ds.filters[k].id != 3) { last = k; break; }
  for (int k = nf - 1; k >= 0; --k) {
    if (c.mask & (1u << k)) continue;
    const Filter& fl = ds.filters[k];
    if (fl.id == 3) {  // fletcher32: checksum behind the data
      F5_REQUIRE(n >= 4, NRQIO_F5_FORMAT);
      n -= 4;
    } else if (fl.id == 1) {
      uint8_t* dst = out;
      if (k != last) {
        std::vector<uint8_t>& t = (cur == t0.data()) ? t1 : t0;
        t.resize(cbytes);
        dst = t.data();
      }
      inflate_into(cur, n, dst, cbytes);
      cur = dst;
      n = cbytes;
    } else if (fl.id == 2) {
      const uint64_t sz = fl.has_cd && fl.cd0 ? fl.cd0 : (uint64_t)ds.dt_size;
      const uint64_t cnt = n / sz;
      uint8_t* dst = out;
      if (k != last) {
        std::vector<uint8_t>& t = (cur == t0.data()) ? t1 : t0;
        t.resize(n);
        dst = t.data();
      } else {
        F5_REQUIRE(n == cbytes, NRQIO_F5_FORMAT);
      }
      for (uint64_t j = 0; j < sz; ++j)
        for (uint64_t i = 0; i < cnt; ++i) dst[i * sz + j] = cur[j * cnt + i];
      std::memcpy(dst + cnt * sz, cur + cnt * sz, n - cnt * sz);
      cur = dst;
    } else {
      throw Fail{NRQIO_F5_UNSUPPORTED};  // e.g. vbz (32020)
    }
  }
  if (last < 0) {  // nothing to undo: stored as it is
    F5_REQUIRE(n == cbytes, NRQIO_F5_FORMAT);
    std::memcpy(out, cur, cbytes);
  }
}

void require_int16(const Dataset& ds) {
  F5_REQUIRE(ds.rank == 1, NRQIO_F5_UNSUPPORTED);
  F5_REQUIRE(ds.dt_class == 0 && ds.dt_size == 2 && ds.dt_signed && !ds.dt_be, NRQIO_F5_UNSUPPORTED);
}

void read_int16(const File& f, const Dataset& ds, int16_t* dst) {
  const uint64_t n = ds.dims[0], nbytes = 2 * n;
  if (n == 0) return;
  uint8_t* out = reinterpret_cast<uint8_t*>(dst);
  if (ds.layout == 0) {
    F5_REQUIRE(ds.compact_n >= nbytes, NRQIO_F5_FORMAT);
    std::memcpy(out, f.at(ds.compact_off, nbytes), nbytes);
    return;
  }
  if (ds.layout == 1) {
    if (ds.addr == kUndef) std::memset(out, 0, nbytes);
    else std::memcpy(out, f.at(ds.addr, nbytes), nbytes);
    return;
  }
  std::vector<Chunk> chunks;
  if (ds.single) {
    if (ds.addr != kUndef) chunks.push_back(Chunk{0, ds.single_size, ds.addr, ds.single_mask});
  } else if (ds.addr != kUndef) {
    list_chunks(f, ds.addr, chunks, 0);
  }
  F5_REQUIRE(ds.cdim > 0, NRQIO_F5_FORMAT);
  const uint64_t cbytes = 2 * ds.cdim;
  std::vector<uint8_t> t0, t1, whole;
  uint64_t covered = 0;
  for (const Chunk& c : chunks) {
    F5_REQUIRE(c.offset < n || n == 0, NRQIO_F5_FORMAT);
    if (c.offset + ds.cdim <= n) {
      unfilter(f, ds, c, cbytes, out + 2 * c.offset, t0, t1);
      covered += ds.cdim;
    } else {  // the chunk at the end of the dataset reaches past it
      whole.resize(cbytes);
      unfilter(f, ds, c, cbytes, whole.data(), t0, t1);
      std::memcpy(out + 2 * c.offset, whole.data(), 2 * (n - c.offset));
      covered += n - c.offset;
    }
  }
  if (covered < n) {  // chunks that were never written read as the fill value (0)
    std::vector<uint8_t> have((size_t)((n + ds.cdim - 1) / ds.cdim), 0);
    for (const Chunk& c : chunks) have[(size_t)(c.offset / ds.cdim)] = 1;
    for (size_t k = 0; k < have.size(); ++k)
      if (!have[k]) {
        const uint64_t lo = k * ds.cdim, hi = std::min(n, lo + ds.cdim);
        std::memset(out + 2 * lo, 0, 2 * (hi - lo));
      }
  }
}

// ------------------------------------------------------------------------------------------------ attributes
struct Attr {
  int dt_class = -1, dt_size = 0;
  bool dt_signed = false, dt_be = false, vlen_str = false;
  int rank = 0;
  uint64_t dims[4] = {0, 0, 0, 0};
  uint64_t data_off = 0, data_n = 0;
};

bool find_attr(const File& f, const std::vector<Msg>& msgs, const char* name, Attr& a) {
  for (const Msg& m : msgs) {
    if (m.type == 0x15) {  // attribute info: dense storage?
      const int flags = (int)f.u(m.off + 1, 1);
      const uint64_t p = m.off + 2 + ((flags & 1) ? 2 : 0);
      F5_REQUIRE(f.addr(p) == kUndef, NRQIO_F5_UNSUPPORTED);
    }
    if (m.type != 0x0C) continue;
    F5_REQUIRE(!(m.flags & 0x02), NRQIO_F5_UNSUPPORTED);
    const int ver = (int)f.u(m.off, 1);
    const uint64_t nsz = f.u(m.off + 2, 2), tsz = f.u(m.off + 4, 2), ssz = f.u(m.off + 6, 2);
    uint64_t p, tp, sp, dp;
    if (ver == 1) {
      p = m.off + 8;
      tp = p + ((nsz + 7) & ~7ull);
      sp = tp + ((tsz + 7) & ~7ull);
      dp = sp + ((ssz + 7) & ~7ull);
    } else if (ver == 2 || ver == 3) {
      F5_REQUIRE(!(f.u(m.off + 1, 1) & 0x03), NRQIO_F5_UNSUPPORTED);
      p = m.off + 8 + (ver == 3 ? 1 : 0);
      tp = p + nsz;
      sp = tp + tsz;
      dp = sp + ssz;
    } else {
      throw Fail{NRQIO_F5_UNSUPPORTED};
    }
    F5_REQUIRE(dp <= m.off + m.n, NRQIO_F5_FORMAT);
    if (cstr_at(f, p) != name) continue;
    parse_datatype(f, tp, &a.dt_class, &a.dt_size, &a.dt_signed, &a.dt_be, &a.vlen_str);
    Msg sm{0x01, 0, sp, ssz};
    F5_REQUIRE(parse_dataspace(f, sm, &a.rank, a.dims), NRQIO_F5_UNSUPPORTED);
    a.data_off = dp;
    a.data_n = m.off + m.n - dp;
    return true;
  }
  return false;
}

double attr_number(const File& f, const Attr& a) {
  uint64_t count = 1;
  for (int k = 0; k < a.rank; ++k) count *= a.dims[k];
  F5_REQUIRE(count == 1, NRQIO_F5_UNSUPPORTED);
  F5_REQUIRE(a.data_n >= (uint64_t)a.dt_size, NRQIO_F5_FORMAT);
  uint8_t raw[8];
  F5_REQUIRE(a.dt_size >= 1 && a.dt_size <= 8, NRQIO_F5_UNSUPPORTED);
  std::memcpy(raw, f.at(a.data_off, a.dt_size), a.dt_size);
  if (a.dt_be) std::reverse(raw, raw + a.dt_size);
  if (a.dt_class == 1) {
    if (a.dt_size == 8) {
      double v;
      std::memcpy(&v, raw, 8);
      return v;
    }
    if (a.dt_size == 4) {
      float v;
      std::memcpy(&v, raw, 4);
      return (double)v;
    }
    throw Fail{NRQIO_F5_UNSUPPORTED};
  }
  F5_REQUIRE(a.dt_class == 0, NRQIO_F5_UNSUPPORTED);
  const uint64_t v = rd_le(raw, a.dt_size);
  if (!a.dt_signed) return (double)v;
  if (a.dt_size == 8) return (double)(int64_t)v;
  const uint64_t sign = 1ull << (8 * a.dt_size - 1);
  return (double)((int64_t)(v ^ sign) - (int64_t)sign);
}

// ================================================================================================ writer
struct Out {
  std::vector<uint8_t> b;
  uint64_t origin = 0;  // address (as stored in the file) of b[0]; a multiple of 8
  uint64_t alloc(const uint8_t* p, size_t n) {
    while (b.size() % 8) b.push_back(0);
    const uint64_t a = origin + b.size();
    b.insert(b.end(), p, p + n);
    return a;
  }
  uint64_t alloc(const std::vector<uint8_t>& v) { return alloc(v.data(), v.size()); }
};

inline void put(std::vector<uint8_t>& v, uint64_t x, int n) {
  for (int i = 0; i < n; ++i) v.push_back((uint8_t)(x >> (8 * i)));
}
inline void put_at(std::vector<uint8_t>& v, size_t off, uint64_t x, int n) {
  for (int i = 0; i < n; ++i) v[off + i] = (uint8_t)(x >> (8 * i));
}
inline void pad8(std::vector<uint8_t>& v) {
  while (v.size() % 8) v.push_back(0);
}

struct M {
  int type, flags;
  std::vector<uint8_t> body;
};

uint64_t object_header(Out& o, const std::vector<M>& msgs) {
  std::vector<uint8_t> body;
  for (const M& m : msgs) {
    const size_t n = (m.body.size() + 7) & ~(size_t)7;
    F5_REQUIRE(n <= 65528, NRQIO_F5_UNSUPPORTED);
    put(body, (uint64_t)m.type, 2);
    put(body, n, 2);
    put(body, (uint64_t)m.flags, 1);
    put(body, 0, 3);
    body.insert(body.end(), m.body.begin(), m.body.end());
    body.resize(body.size() + (n - m.body.size()), 0);
  }
  std::vector<uint8_t> h;
  put(h, 1, 1);
  put(h, 0, 1);
  put(h, msgs.size(), 2);
  put(h, 1, 4);
  put(h, body.size(), 4);
  put(h, 0, 4);
  h.insert(h.end(), body.begin(), body.end());
  return o.alloc(h);
}

// datatype messages (version 1 encodings)
std::vector<uint8_t> dt_int(int size, bool is_signed) {
  std::vector<uint8_t> v;
  put(v, 0x10, 1);
  put(v, is_signed ? 8 : 0, 1);
  put(v, 0, 2);
  put(v, (uint64_t)size, 4);
  put(v, 0, 2);
  put(v, 8 * (uint64_t)size, 2);
  return v;
}
std::vector<uint8_t> dt_f64() {
  std::vector<uint8_t> v;
  put(v, 0x11, 1);
  put(v, 0x20, 1);
  put(v, 63, 1);
  put(v, 0, 1);
  put(v, 8, 4);
  put(v, 0, 2);
  put(v, 64, 2);
  put(v, 52, 1);
  put(v, 11, 1);
  put(v, 0, 1);
  put(v, 52, 1);
  put(v, 1023, 4);
  return v;
}
std::vector<uint8_t> dt_s1() {
  std::vector<uint8_t> v;
  put(v, 0x13, 1);
  put(v, 0x01, 1);
  put(v, 0, 2);
  put(v, 1, 4);
  return v;
}
std::vector<uint8_t> dt_vlen_utf8() {
  std::vector<uint8_t> v;
  put(v, 0x19, 1);
  put(v, 0x01, 1);
  put(v, 0x01, 1);
  put(v, 0, 1);
  put(v, 16, 4);
  put(v, 0x13, 1);
  put(v, 0x10, 1);
  put(v, 0, 2);
  put(v, 1, 4);
  return v;
}
// the Events compound (resquiggle.py:69-70): norm_mean f8, norm_stdev f8, start u4, length u4, base S1 -- 25 bytes
std::vector<uint8_t> dt_events() {
  struct Member {
    const char* name;
    int off;
    std::vector<uint8_t> dt;
  };
  const Member members[5] = {{"norm_mean", 0, dt_f64()},
                             {"norm_stdev", 8, dt_f64()},
                             {"start", 16, dt_int(4, false)},
                             {"length", 20, dt_int(4, false)},
                             {"base", 24, dt_s1()}};
  std::vector<uint8_t> v;
  put(v, 0x16, 1);
  put(v, 5, 1);
  put(v, 0, 2);
  put(v, 25, 4);
  for (const Member& m : members) {
    const size_t nl = std::strlen(m.name) + 1;
    v.insert(v.end(), m.name, m.name + nl);
    v.resize(v.size() + (((nl + 7) & ~(size_t)7) - nl), 0);
    put(v, (uint64_t)m.off, 4);
    put(v, 0, 1);
    put(v, 0, 3);
    put(v, 0, 4);
    put(v, 0, 4);
    v.resize(v.size() + 16, 0);
    v.insert(v.end(), m.dt.begin(), m.dt.end());
  }
  return v;
}

std::vector<uint8_t> ds_scalar() {
  std::vector<uint8_t> v;
  put(v, 1, 1);
  put(v, 0, 1);
  put(v, 0, 1);
  put(v, 0, 5);
  return v;
}
std::vector<uint8_t> ds_1d(uint64_t n) {
  std::vector<uint8_t> v;
  put(v, 1, 1);
  put(v, 1, 1);
  put(v, 1, 1);
  put(v, 0, 5);
  put(v, n, 8);
  put(v, n, 8);
  return v;
}

M attr_message(const char* name, const std::vector<uint8_t>& dt, const std::vector<uint8_t>& ds,
               const std::vector<uint8_t>& data) {
  M m{0x0C, 0, {}};
  const size_t nl = std::strlen(name) + 1;
  put(m.body, 1, 1);
  put(m.body, 0, 1);
  put(m.body, nl, 2);
  put(m.body, dt.size(), 2);
  put(m.body, ds.size(), 2);
  m.body.insert(m.body.end(), name, name + nl);
  pad8(m.body);
  m.body.insert(m.body.end(), dt.begin(), dt.end());
  pad8(m.body);
  m.body.insert(m.body.end(), ds.begin(), ds.end());
  pad8(m.body);
  m.body.insert(m.body.end(), data.begin(), data.end());
  return m;
}
M attr_f64(const char* name, double x) {
  std::vector<uint8_t> d(8);
  std::memcpy(d.data(), &x, 8);
  return attr_message(name, dt_f64(), ds_scalar(), d);
}
M attr_i64(const char* name, int64_t x) {
  std::vector<uint8_t> d;
  put(d, (uint64_t)x, 8);
  return attr_message(name, dt_int(8, true), ds_scalar(), d);
}

// variable-length strings of one write: one global heap collection (4096 bytes at least, as libhdf5 makes them)
struct Strings {
  uint64_t gcol = 0;
  std::vector<std::string> items;
  M attr(const char* name, const char* s) {
    items.emplace_back(s);
    std::vector<uint8_t> d;
    put(d, std::strlen(s), 4);
    put(d, gcol, 8);
    put(d, items.size(), 4);
    return attr_message(name, dt_vlen_utf8(), ds_scalar(), d);
  }
};

struct Child {
  std::string name;
  uint64_t hdr;
  uint32_t cache;
  uint8_t scratch[16];
};

// local heap + symbol table nodes + B-tree over `children` (hdf5_min.Writer._group); -> (B-tree, heap) addresses
void build_symtab(Out& o, std::vector<Child>& children, int leaf_k, int int_k, uint64_t* btree, uint64_t* heap_addr) {
  std::sort(children.begin(), children.end(), [](const Child& a, const Child& b) { return a.name < b.name; });
  std::vector<uint8_t> heap(8, 0);
  std::vector<uint64_t> name_off(children.size());
  for (size_t k = 0; k < children.size(); ++k) {
    name_off[k] = heap.size();
    heap.insert(heap.end(), children[k].name.begin(), children[k].name.end());
    heap.push_back(0);
    pad8(heap);
  }
  const uint64_t free_off = heap.size();
  put(heap, 1, 8);
  put(heap, 16, 8);
  const uint64_t daddr = o.alloc(heap);
  std::vector<uint8_t> hh;
  hh.insert(hh.end(), {'H', 'E', 'A', 'P'});
  put(hh, 0, 1);
  put(hh, 0, 3);
  put(hh, heap.size(), 8);
  put(hh, free_off, 8);
  put(hh, daddr, 8);
  *heap_addr = o.alloc(hh);
  const size_t cap = 2 * (size_t)leaf_k;
  std::vector<std::pair<uint64_t, uint64_t>> nodes;  // (address, heap offset of the last name below it)
  for (size_t lo = 0; lo < children.size(); lo += cap) {
    const size_t hi = std::min(children.size(), lo + cap);
    std::vector<uint8_t> node(8 + cap * 40, 0);
    std::memcpy(node.data(), "SNOD", 4);
    node[4] = 1;
    put_at(node, 6, hi - lo, 2);
    for (size_t k = lo; k < hi; ++k) {
      const size_t q = 8 + 40 * (k - lo);
      put_at(node, q, name_off[k], 8);
      put_at(node, q + 8, children[k].hdr, 8);
      put_at(node, q + 16, children[k].cache, 4);
      std::memcpy(node.data() + q + 24, children[k].scratch, 16);
    }
    nodes.emplace_back(o.alloc(node), name_off[hi - 1]);
  }
  const size_t cap_i = 2 * (size_t)int_k;
  int level = 0;
  for (;;) {
    std::vector<std::pair<uint64_t, uint64_t>> parents;
    std::vector<size_t> where;  // offsets of the parents inside o.b (for the sibling pointers)
    uint64_t low_key = 0;
    for (size_t lo = 0; lo < std::max<size_t>(nodes.size(), 1); lo += cap_i) {
      const size_t hi = std::min(nodes.size(), lo + cap_i);
      std::vector<uint8_t> node(24 + (cap_i + 1) * 8 + cap_i * 8, 0);
      std::memcpy(node.data(), "TREE", 4);
      node[4] = 0;
      node[5] = (uint8_t)level;
      put_at(node, 6, hi > lo ? hi - lo : 0, 2);
      put_at(node, 8, kUndef, 8);
      put_at(node, 16, kUndef, 8);
      size_t p = 24;
      put_at(node, p, low_key, 8);
      p += 8;
      for (size_t k = lo; k < hi; ++k) {
        put_at(node, p, nodes[k].first, 8);
        put_at(node, p + 8, nodes[k].second, 8);
        p += 16;
      }
      low_key = hi > lo ? nodes[hi - 1].second : 0;
      const uint64_t a = o.alloc(node);
      where.push_back((size_t)(a - o.origin));
      parents.emplace_back(a, hi > lo ? nodes[hi - 1].second : 0);
    }
    if (parents.size() == 1) {
      *btree = parents[0].first;
      return;
    }
    for (size_t k = 0; k < parents.size(); ++k) {
      put_at(o.b, where[k] + 8, k > 0 ? parents[k - 1].first : kUndef, 8);
      put_at(o.b, where[k] + 16, k + 1 < parents.size() ? parents[k + 1].first : kUndef, 8);
    }
    nodes = parents;
    ++level;
  }
}

uint64_t build_group(Out& o, std::vector<Child>& children, const std::vector<M>& attrs, int leaf_k, int int_k,
                     uint64_t* btree, uint64_t* heap) {
  build_symtab(o, children, leaf_k, int_k, btree, heap);
  std::vector<M> msgs;
  M st{0x11, 0, {}};
  put(st.body, *btree, 8);
  put(st.body, *heap, 8);
  msgs.push_back(st);
  msgs.insert(msgs.end(), attrs.begin(), attrs.end());
  return object_header(o, msgs);
}

// a 1-D dataset held as one gzip chunk that is already deflated (hdf5_min.Writer._dataset)
uint64_t build_chunked(Out& o, uint64_t n, const std::vector<uint8_t>& dt, int itemsize, const uint8_t* chunk,
                       uint64_t chunk_n, int chunk_k, const std::vector<M>& attrs) {
  std::vector<M> msgs;
  msgs.push_back(M{0x01, 0, ds_1d(n)});
  msgs.push_back(M{0x03, 1, dt});
  if (n > 0) {
    const uint64_t caddr = o.alloc(chunk, (size_t)chunk_n);
    const size_t ksz = 8 + 8 * 2;
    std::vector<uint8_t> node(24 + (2 * (size_t)chunk_k + 1) * ksz + 2 * (size_t)chunk_k * 8, 0);
    std::memcpy(node.data(), "TREE", 4);
    node[4] = 1;
    node[5] = 0;
    put_at(node, 6, 1, 2);
    put_at(node, 8, kUndef, 8);
    put_at(node, 16, kUndef, 8);
    size_t p = 24;
    put_at(node, p, chunk_n, 4);  // key 0: the chunk at the origin
    p += ksz;
    put_at(node, p, caddr, 8);
    p += 8;
    put_at(node, p + 8, n, 8);    // key 1: one chunk past the end
    const uint64_t baddr = o.alloc(node);
    msgs.push_back(M{0x05, 0, {2, 3, 2, 0}});  // fill value: incremental, if set, undefined
    M fp{0x0B, 0, {}};
    put(fp.body, 1, 1);
    put(fp.body, 1, 1);
    put(fp.body, 0, 6);
    put(fp.body, 1, 2);
    put(fp.body, 8, 2);
    put(fp.body, 1, 2);
    put(fp.body, 1, 2);
    fp.body.insert(fp.body.end(), {'d', 'e', 'f', 'l', 'a', 't', 'e', 0});
    put(fp.body, 4, 4);
    put(fp.body, 0, 4);
    msgs.push_back(fp);
    M lay{0x08, 0, {}};
    put(lay.body, 3, 1);
    put(lay.body, 2, 1);
    put(lay.body, 2, 1);
    put(lay.body, baddr, 8);
    put(lay.body, n, 4);
    put(lay.body, (uint64_t)itemsize, 4);
    msgs.push_back(lay);
  } else {  // an empty dataset is contiguous with no storage (as hdf5_min writes it)
    msgs.push_back(M{0x05, 0, {2, 2, 2, 0}});
    M lay{0x08, 0, {}};
    put(lay.body, 3, 1);
    put(lay.body, 1, 1);
    put(lay.body, kUndef, 8);
    put(lay.body, 0, 8);
    msgs.push_back(lay);
  }
  msgs.insert(msgs.end(), attrs.begin(), attrs.end());
  return object_header(o, msgs);
}

void write_all(int fd, const uint8_t* p, size_t n, uint64_t off) {
  size_t done = 0;
  while (done < n) {
    const ssize_t k = pwrite(fd, p + done, n - done, (off_t)(off + done));
    F5_REQUIRE(k > 0, NRQIO_F5_IO);
    done += (size_t)k;
  }
}

void write_corrected_one(const nrqio_f5_corrected_batch* w, int i) {
  const char* path = w->strtab + w->path_off[i];
  File f;
  int fd = -1;
  load_file(path, O_RDWR, f, &fd);
  struct Closer {
    int fd;
    ~Closer() { if (fd >= 0) ::close(fd); }
  } closer{fd};
  parse_superblock(f);
  // the append needs: 8-byte addresses, a superblock whose end-of-file address can be patched, v1 headers
  F5_REQUIRE((f.sb_ver == 0 || f.sb_ver == 1) && f.osz == 8 && f.lsz == 8, NRQIO_F5_UNSUPPORTED);
  std::vector<Msg> msgs;
  Entry corr;
  int version = 0;
  const char* parts[2] = {"Analyses", w->corrected_group};
  walk(f, parts, 2, msgs, &corr, &version);
  F5_REQUIRE(version == 1 && corr.entry_off != 0, NRQIO_F5_UNSUPPORTED);
  std::vector<Entry> members;
  int stab = -1;
  F5_REQUIRE(list_group(f, msgs, members, &stab), NRQIO_F5_MISSING);
  F5_REQUIRE(stab >= 0, NRQIO_F5_UNSUPPORTED);  // a group of link messages: the Python writer rewrites the file
  Entry dummy;
  F5_REQUIRE(!find_entry(members, w->subgroup, dummy), NRQIO_F5_EXISTS);

  const uint64_t eof_rel = f.u(f.eof_field, 8);
  Out o;
  o.origin = (std::max<uint64_t>(f.d.size() > f.base ? f.d.size() - f.base : 0, eof_rel) + 7) & ~7ull;
  const int leaf_k = f.leaf_k > 0 ? f.leaf_k : 4, int_k = f.int_k > 0 ? f.int_k : 16;
  const int chunk_k = f.chunk_k > 0 ? f.chunk_k : 32;

  Strings strs;
  {
    std::vector<uint8_t> blank(4096, 0);
    strs.gcol = o.alloc(blank);
  }
  const uint8_t* chunks = w->chunks;
  const int64_t* coff = w->chunk_off + 4 * (int64_t)i;
  const int64_t* clen = w->chunk_len + 4 * (int64_t)i;
  const int64_t* iv = w->ints + 8 * (int64_t)i;
  const double* sv = w->scale_values + 4 * (int64_t)i;
  const uint64_t n_bases = (uint64_t)w->n_bases[i], n_align = (uint64_t)w->n_align[i], n_segs = (uint64_t)w->n_segs[i];

  // Alignment: three datasets (insertion order of the reference, :112-113), nine attributes
  std::vector<Child> al_children;
  auto add_ds = [&](std::vector<Child>& into, const char* name, uint64_t hdr) {
    Child c{name, hdr, 0, {0}};
    into.push_back(c);
  };
  add_ds(al_children, "read_alignment", build_chunked(o, n_align, dt_s1(), 1, chunks + coff[1], (uint64_t)clen[1], chunk_k, {}));
  add_ds(al_children, "genome_alignment", build_chunked(o, n_align, dt_s1(), 1, chunks + coff[2], (uint64_t)clen[2], chunk_k, {}));
  add_ds(al_children, "read_segments", build_chunked(o, n_segs, dt_int(8, true), 8, chunks + coff[3], (uint64_t)clen[3], chunk_k, {}));
  std::vector<M> al_attrs;
  al_attrs.push_back(attr_i64("mapped_start", iv[0]));
  al_attrs.push_back(strs.attr("mapped_strand", w->strtab + w->strand_off[i]));
  al_attrs.push_back(strs.attr("mapped_chrom", w->strtab + w->chrom_off[i]));
  al_attrs.push_back(attr_i64("clipped_bases_start", iv[1]));
  al_attrs.push_back(attr_i64("clipped_bases_end", iv[2]));
  al_attrs.push_back(attr_i64("num_insertions", iv[3]));
  al_attrs.push_back(attr_i64("num_deletions", iv[4]));
  al_attrs.push_back(attr_i64("num_matches", iv[5]));
  al_attrs.push_back(attr_i64("num_mismatches", iv[6]));
  uint64_t al_bt = 0, al_heap = 0;
  const uint64_t al_hdr = build_group(o, al_children, al_attrs, leaf_k, int_k, &al_bt, &al_heap);

  std::vector<M> ev_attrs;
  ev_attrs.push_back(attr_i64("read_start_rel_to_raw", iv[7]));
  const uint64_t ev_hdr = build_chunked(o, n_bases, dt_events(), 25, chunks + coff[0], (uint64_t)clen[0], chunk_k, ev_attrs);

  std::vector<Child> sub_children;
  {
    Child c{"Alignment", al_hdr, 1, {0}};
    for (int k = 0; k < 8; ++k) c.scratch[k] = (uint8_t)(al_bt >> (8 * k)), c.scratch[8 + k] = (uint8_t)(al_heap >> (8 * k));
    sub_children.push_back(c);
  }
  add_ds(sub_children, "Events", ev_hdr);
  std::vector<M> sub_attrs;
  sub_attrs.push_back(attr_f64("shift", sv[0]));
  sub_attrs.push_back(attr_f64("scale", sv[1]));
  sub_attrs.push_back(attr_f64("lower_lim", sv[2]));
  sub_attrs.push_back(attr_f64("upper_lim", sv[3]));
  sub_attrs.push_back(strs.attr("norm_type", w->norm_type));
  if (w->outlier_is_int) sub_attrs.push_back(attr_i64("outlier_threshold", (int64_t)w->outlier_threshold));
  else sub_attrs.push_back(attr_f64("outlier_threshold", w->outlier_threshold));
  uint64_t sub_bt = 0, sub_heap = 0;
  const uint64_t sub_hdr = build_group(o, sub_children, sub_attrs, leaf_k, int_k, &sub_bt, &sub_heap);

  // the global heap collection of the three strings
  {
    std::vector<uint8_t> col;
    col.insert(col.end(), {'G', 'C', 'O', 'L'});
    put(col, 1, 1);
    put(col, 0, 3);
    put(col, 4096, 8);
    for (size_t k = 0; k < strs.items.size(); ++k) {
      put(col, k + 1, 2);
      put(col, 1, 2);
      put(col, 0, 4);
      put(col, strs.items[k].size(), 8);
      col.insert(col.end(), strs.items[k].begin(), strs.items[k].end());
      pad8(col);
    }
    F5_REQUIRE(col.size() + 16 <= 4096, NRQIO_F5_UNSUPPORTED);  // (chromosome names are short)
    const uint64_t free_n = 4096 - col.size();
    put(col, 0, 2);
    put(col, 0, 2);
    put(col, 0, 4);
    put(col, free_n, 8);
    std::memcpy(o.b.data() + (strs.gcol - o.origin), col.data(), col.size());
  }

  // the corrected group's members again, plus the new one
  std::vector<Child> corr_children;
  for (const Entry& e : members) {
    Child c{e.name, e.hdr - f.base, e.cache, {0}};
    std::memcpy(c.scratch, e.scratch, 16);
    corr_children.push_back(c);
  }
  {
    Child c{w->subgroup, sub_hdr, 1, {0}};
    for (int k = 0; k < 8; ++k) c.scratch[k] = (uint8_t)(sub_bt >> (8 * k)), c.scratch[8 + k] = (uint8_t)(sub_heap >> (8 * k));
    corr_children.push_back(c);
  }
  uint64_t corr_bt = 0, corr_heap = 0;
  build_symtab(o, corr_children, leaf_k, int_k, &corr_bt, &corr_heap);
  pad8(o.b);

  // 1. the new objects, 2. the end-of-file address, 3. the pointers to the group's new B-tree and heap
  write_all(fd, o.b.data(), o.b.size(), f.base + o.origin);
  uint8_t w8[16];
  const uint64_t new_eof = o.origin + o.b.size();
  for (int k = 0; k < 8; ++k) w8[k] = (uint8_t)(new_eof >> (8 * k));
  write_all(fd, w8, 8, f.eof_field);
  for (int k = 0; k < 8; ++k) w8[k] = (uint8_t)(corr_bt >> (8 * k)), w8[8 + k] = (uint8_t)(corr_heap >> (8 * k));
  write_all(fd, w8, 16, msgs[(size_t)stab].off);
  if (corr.cache == 1) write_all(fd, w8, 16, corr.entry_off + 2 * 8 + 8);
}

}  // namespace

// ================================================================================================ C ABI
struct nrqio_f5_batch {
  std::vector<File> files;
  std::vector<Dataset> sig;
  std::vector<int32_t> status;
  std::vector<std::string> paths;
};

extern "C" {

int nrqio_f5_open_raw_batch(const char* strtab, const int64_t* path_off, int32_t n, int32_t n_threads,
                            nrqio_f5_batch** out, int64_t* n_samples, double* channel, int32_t* status) {
  if (!strtab || !path_off || n < 0 || !out || !n_samples || !status) return -1;
  nrqio_f5_batch* h = new nrqio_f5_batch();
  h->files.resize((size_t)n);
  h->sig.resize((size_t)n);
  h->status.assign((size_t)n, NRQIO_F5_OK);
  parallel_for(n, n_threads, [&](int i) {
    File& f = h->files[(size_t)i];
    n_samples[i] = 0;
    try {
      load_file(strtab + path_off[i], O_RDONLY, f, nullptr);
      parse_superblock(f);
      std::vector<Msg> msgs;
      // /Raw/Reads/<the first read>/Signal (resquiggle.py:331-333)
      const char* parts[2] = {"Raw", "Reads"};
      walk(f, parts, 2, msgs, nullptr, nullptr);
      std::vector<Entry> es;
      F5_REQUIRE(list_group(f, msgs, es, nullptr) && !es.empty(), NRQIO_F5_MISSING);
      const Entry* first = &es[0];
      for (const Entry& e : es)
        if (e.name < first->name) first = &e;
      read_messages(f, first->hdr, msgs, nullptr);
      F5_REQUIRE(list_group(f, msgs, es, nullptr), NRQIO_F5_MISSING);
      Entry sig;
      F5_REQUIRE(find_entry(es, "Signal", sig), NRQIO_F5_MISSING);
      read_messages(f, sig.hdr, msgs, nullptr);
      Dataset& ds = h->sig[(size_t)i];
      F5_REQUIRE(parse_dataset(f, msgs, ds), NRQIO_F5_MISSING);
      require_int16(ds);
      n_samples[i] = (int64_t)ds.dims[0];
      // /UniqueGlobalKey/channel_id attributes (nanoraw_helper.py:154-166)
      const char* cparts[2] = {"UniqueGlobalKey", "channel_id"};
      walk(f, cparts, 2, msgs, nullptr, nullptr);
      static const char* names[4] = {"offset", "range", "digitisation", "sampling_rate"};
      for (int k = 0; k < 4; ++k) {
        Attr a;
        F5_REQUIRE(find_attr(f, msgs, names[k], a), NRQIO_F5_MISSING);
        const double v = attr_number(f, a);
        if (channel) channel[4 * (int64_t)i + k] = v;
      }
      Attr a;
      F5_REQUIRE(find_attr(f, msgs, "channel_number", a), NRQIO_F5_MISSING);
    } catch (const Fail& e) {
      h->status[(size_t)i] = e.code;
      f.d.clear();
      f.d.shrink_to_fit();
    } catch (...) {
      h->status[(size_t)i] = NRQIO_F5_FORMAT;
      f.d.clear();
    }
    status[i] = h->status[(size_t)i];
  });
  *out = h;
  return 0;
}

int nrqio_f5_read_raw_batch(nrqio_f5_batch* h, int16_t* dst, const int64_t* dst_off, int32_t n_threads,
                            int32_t* status) {
  if (!h || !dst || !dst_off || !status) return -1;
  const int n = (int)h->files.size();
  parallel_for(n, n_threads, [&](int i) {
    status[i] = h->status[(size_t)i];
    if (status[i] != NRQIO_F5_OK || dst_off[i] < 0) return;
    try {
      read_int16(h->files[(size_t)i], h->sig[(size_t)i], dst + dst_off[i]);
    } catch (const Fail& e) {
      status[i] = e.code;
    } catch (...) {
      status[i] = NRQIO_F5_FORMAT;
    }
    h->files[(size_t)i].d.clear();
    h->files[(size_t)i].d.shrink_to_fit();
  });
  return 0;
}

void nrqio_f5_close_batch(nrqio_f5_batch* h) { delete h; }

int nrqio_f5_write_corrected_batch(const nrqio_f5_corrected_batch* w, int32_t n_threads, int32_t* status) {
  if (!w || !status || w->n < 0 || !w->strtab || !w->path_off || !w->chrom_off || !w->strand_off || !w->scale_values ||
      !w->ints || !w->n_bases || !w->n_align || !w->n_segs || !w->chunk_off || !w->chunk_len || !w->corrected_group ||
      !w->subgroup || !w->norm_type)
    return -1;
  parallel_for(w->n, n_threads, [&](int i) {
    if (w->skip && w->skip[i]) {
      status[i] = NRQIO_F5_OK;
      return;
    }
    try {
      write_corrected_one(w, i);
      status[i] = NRQIO_F5_OK;
    } catch (const Fail& e) {
      status[i] = e.code;
    } catch (...) {
      status[i] = NRQIO_F5_FORMAT;
    }
  });
  return 0;
}

}  // extern "C"
