// C ABI of libnrq.so (include/nrq.h).  Unity build: the kernels live in the .cu files included
// below; this file only validates arguments, carves the workspace and launches.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "nrq_indels.cu"
#include "nrq_norm.cu"
#include "nrq_resegment.cu"
#include "nrq_events.cu"

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(const char* what, cudaError_t e = cudaSuccess) {
  g_err = what;
  if (e != cudaSuccess) {
    g_err += ": ";
    g_err += cudaGetErrorString(e);
  }
  return e != cudaSuccess ? -2 : -1;
}

#define NRQ_CHECK_LAUNCH(name)                          \
  do {                                                  \
    g_launches.fetch_add(1, std::memory_order_relaxed); \
    cudaError_t e__ = cudaGetLastError();               \
    if (e__ != cudaSuccess) return fail(name, e__);     \
  } while (0)

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

int resegment_grid(int n_reads) {
  static int cached = 0;
  if (!cached) {
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nrq::k_resegment, nrq::kSegWarps * 32, 0);
    if (per_sm < 1) per_sm = 1;
    cached = sms * per_sm;
  }
  int need = (n_reads + nrq::kSegWarps - 1) / nrq::kSegWarps;
  return need < cached ? (need > 0 ? need : 1) : cached;
}

struct Workspace {
  int64_t* indel_off;
  int32_t* work_counter;
  int32_t* indels;
  int32_t* groups;
  int64_t indel_cap;
  double* big;
  int64_t big_doubles;
  int64_t bytes;
};

int64_t default_indel_cap(int32_t n_reads, int64_t total_align) { return total_align / 4 + n_reads + 64; }

Workspace carve(void* base, int32_t n_reads, int64_t total_align, const nrq_params* p, int64_t indel_cap) {
  Workspace w;
  int64_t off = 0;
  char* b = static_cast<char*>(base);
  w.indel_off = reinterpret_cast<int64_t*>(b + off);
  off = align_up(off + 8 * ((int64_t)n_reads + 1), 256);
  w.work_counter = reinterpret_cast<int32_t*>(b + off);
  off = align_up(off + 256, 256);
  w.indel_cap = indel_cap;
  w.indels = reinterpret_cast<int32_t*>(b + off);
  off = align_up(off + 16 * indel_cap, 256);
  w.groups = reinterpret_cast<int32_t*>(b + off);
  off = align_up(off + 16 * indel_cap, 256);
  int grid = resegment_grid(n_reads);
  w.big_doubles = (int64_t)grid * nrq::kSegWarps * ((int64_t)p->big_region_cap + 1);
  w.big = reinterpret_cast<double*>(b + off);
  off = align_up(off + 8 * w.big_doubles, 256);
  w.bytes = off;
  (void)total_align;
  return w;
}

int check_params(const nrq_params* p) {
  if (!p) return fail("params is NULL");
  if (p->min_seg_len < 1 || p->min_seg_len > 16) return fail("min_seg_len must be in [1, 16]");
  if (p->big_region_cap < 0) return fail("big_region_cap must be >= 0");
  if (p->norm_type != NRQ_NORM_MEDIAN && p->norm_type != NRQ_NORM_GIVEN) return fail("unknown norm_type");
  return 0;
}

int check_batch(const nrq_batch* b) {
  if (!b) return fail("batch is NULL");
  if (b->n_reads < 0) return fail("n_reads < 0");
  if (b->n_reads > 0 && (!b->raw || !b->raw_off || !b->read_start || !b->starts || !b->starts_off ||
                         !b->read_align || !b->genome_align || !b->align_off))
    return fail("batch has NULL arrays");
  return 0;
}

}  // namespace

extern "C" {

int nrq_version(void) { return NRQ_VERSION; }
const char* nrq_last_error(void) { return g_err.c_str(); }
int64_t nrq_launch_count(void) { return g_launches.load(); }

const char* nrq_status_message(int32_t st) {
  switch (st) {
    case NRQ_ST_OK: return "";
    case NRQ_ST_CPTS_LIMIT: return "Reached maximum number of changepoints for a single indel";
    case NRQ_ST_TIMEOUT: return "Read took too long to re-segment.";
    case NRQ_ST_ZERO_LEN: return "New segments include zero length events.";
    case NRQ_ST_NEG_START: return "New segments start with negative index.";
    case NRQ_ST_PAST_END: return "New segments end past raw signal values.";
    case NRQ_ST_SEQ_MISMATCH: return "Aligned sequence does not match number of segments produced.";
    case NRQ_ST_FPE_DIVIDE: return "divide by zero encountered in divide";
    case NRQ_ST_FPE_INVALID: return "invalid value encountered in divide";
    case NRQ_ST_UNSATISFIABLE:
      return "Indel region spans the whole read and still cannot be re-segmented (the reference does not "
             "terminate here).";
    case NRQ_ST_REGION_TOO_LARGE: return "Indel region exceeds the re-squiggle workspace capacity.";
    case NRQ_ST_BAD_ALIGNMENT: return "Alignment does not start and end with matched bases.";
    case NRQ_ST_BAD_INPUT: return "Inconsistent or empty read data.";
    case NRQ_ST_WORKSPACE: return "Re-squiggle workspace too small for this read.";
    default: return "Unknown re-squiggle status.";
  }
}

int64_t nrq_workspace_bytes(int32_t n_reads, int64_t total_align, const nrq_params* p) {
  if (check_params(p)) return -1;
  Workspace w = carve(nullptr, n_reads, total_align, p, default_indel_cap(n_reads, total_align));
  return w.bytes;
}

int nrq_find_indels(const nrq_batch* b, int32_t* status, int32_t* n_indels, int32_t* n_bases,
                    int64_t* indel_off, int32_t* indels, int64_t indel_capacity, uint8_t* ev_base,
                    void* stream) {
  if (int e = check_batch(b)) return e;
  if (!status || !n_indels || !n_bases || !indel_off) return fail("nrq_find_indels: NULL output");
  if (b->n_reads == 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int grid = (b->n_reads + nrq::kWarpsPerCta - 1) / nrq::kWarpsPerCta;
  nrq::k_align_count<<<grid, nrq::kWarpsPerCta * 32, 0, s>>>(*b, status, n_indels, n_bases);
  NRQ_CHECK_LAUNCH("k_align_count");
  nrq::k_scan_counts<<<1, 1024, 0, s>>>(n_indels, b->n_reads, indel_off);
  NRQ_CHECK_LAUNCH("k_scan_counts");
  if (indels) {
    nrq::k_align_fill<<<grid, nrq::kWarpsPerCta * 32, 0, s>>>(
        *b, status, indel_off, reinterpret_cast<int4*>(indels), indel_capacity, ev_base, status);
    NRQ_CHECK_LAUNCH("k_align_fill");
  }
  return 0;
}

int nrq_norm_stats(const nrq_batch* b, const nrq_params* p, int32_t* status, double* scale_values,
                   void* stream) {
  if (int e = check_batch(b)) return e;
  if (int e = check_params(p)) return e;
  if (!status || !scale_values) return fail("nrq_norm_stats: NULL output");
  if (p->norm_type == NRQ_NORM_GIVEN && !b->scale_values_in) return fail("NRQ_NORM_GIVEN needs scale_values_in");
  if (b->n_reads == 0) return 0;
  nrq::k_norm_stats<<<b->n_reads, nrq::kNormThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      *b, *p, status, scale_values);
  NRQ_CHECK_LAUNCH("k_norm_stats");
  return 0;
}

int nrq_normalize_signal(const nrq_batch* b, const double* scale_values, const int32_t* status,
                         const int64_t* norm_off, double* norm_signal, void* stream) {
  if (int e = check_batch(b)) return e;
  if (!scale_values || !status || !norm_off || !norm_signal) return fail("nrq_normalize_signal: NULL argument");
  if (b->n_reads == 0) return 0;
  dim3 grid(b->n_reads, 16);
  nrq::k_normalize_signal<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*b, scale_values, status,
                                                                               norm_off, norm_signal);
  NRQ_CHECK_LAUNCH("k_normalize_signal");
  return 0;
}

int nrq_resegment(const nrq_batch* b, const nrq_params* p, const double* scale_values,
                  const int64_t* indel_off, const int32_t* indels, const int32_t* n_bases,
                  int32_t* group_scratch, double* big_scratch, int64_t big_scratch_doubles,
                  int32_t* work_counter, int32_t* status, int32_t* new_segs, int32_t* region_samples,
                  int32_t* n_groups, void* stream) {
  if (int e = check_batch(b)) return e;
  if (int e = check_params(p)) return e;
  if (!scale_values || !indel_off || !indels || !n_bases || !group_scratch || !work_counter || !status ||
      !new_segs)
    return fail("nrq_resegment: NULL argument");
  if (b->n_reads == 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int grid = resegment_grid(b->n_reads);
  int64_t need = (int64_t)grid * nrq::kSegWarps * ((int64_t)p->big_region_cap + 1);
  if (p->big_region_cap > 0 && (!big_scratch || big_scratch_doubles < need))
    return fail("nrq_resegment: big_scratch too small");
  cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(int32_t), s);
  if (e != cudaSuccess) return fail("memset work_counter", e);
  nrq::k_resegment<<<grid, nrq::kSegWarps * 32, 0, s>>>(
      *b, *p, scale_values, indel_off, reinterpret_cast<const int4*>(indels), n_bases,
      reinterpret_cast<int4*>(group_scratch), big_scratch, work_counter, status, new_segs, region_samples,
      n_groups);
  NRQ_CHECK_LAUNCH("k_resegment");
  return 0;
}

int nrq_event_table(const nrq_batch* b, const nrq_params* p, const double* scale_values,
                    const int32_t* n_bases, const int32_t* new_segs, const int32_t* status,
                    double* norm_mean, double* norm_stdev, uint32_t* ev_start, uint32_t* ev_length,
                    uint8_t* ev_base, void* stream) {
  if (int e = check_batch(b)) return e;
  if (int e = check_params(p)) return e;
  if (!scale_values || !n_bases || !new_segs || !status || !norm_mean || !norm_stdev || !ev_start || !ev_length)
    return fail("nrq_event_table: NULL argument");
  (void)ev_base;  // written by nrq_find_indels
  if (b->n_reads == 0) return 0;
  if ((p->debug_flags & 2) || b->norm_signal)
    nrq::k_event_table_exact<<<b->n_reads, nrq::kEvThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        *b, *p, scale_values, n_bases, new_segs, status, norm_mean, norm_stdev, ev_start, ev_length);
  else
    nrq::k_event_table<<<(b->n_reads + nrq::kEvWarps - 1) / nrq::kEvWarps, nrq::kEvWarps * 32, 0,
                         static_cast<cudaStream_t>(stream)>>>(
        *b, *p, scale_values, n_bases, new_segs, status, norm_mean, norm_stdev, ev_start, ev_length);
  NRQ_CHECK_LAUNCH("k_event_table");
  return 0;
}

static int resquiggle_batch_impl(const nrq_batch* b, const nrq_params* p, const nrq_results* out,
                                 void* workspace, int64_t workspace_bytes, void* stream, cudaEvent_t* ev) {
  if (int e = check_batch(b)) return e;
  if (int e = check_params(p)) return e;
  if (!out || !out->status || !out->n_bases || !out->n_indels || !out->scale_values || !out->new_segs ||
      !out->norm_mean || !out->norm_stdev || !out->ev_start || !out->ev_length || !out->ev_base)
    return fail("nrq_resquiggle_batch: NULL output array");
  if (b->n_reads == 0) return 0;
  if (!workspace) return fail("nrq_resquiggle_batch: NULL workspace");
  // total_align is not known on the host here; the indel capacity is whatever fits the workspace
  Workspace probe = carve(nullptr, b->n_reads, 0, p, 0);
  int64_t spare = workspace_bytes - probe.bytes - 1024;
  if (spare < 32) return fail("nrq_resquiggle_batch: workspace too small");
  int64_t indel_cap = spare / 32;
  Workspace w = carve(workspace, b->n_reads, 0, p, indel_cap);
  if (w.bytes > workspace_bytes) return fail("nrq_resquiggle_batch: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto mark = [&](int i) { if (ev) cudaEventRecord(ev[i], s); };

  mark(0);
  if (int e = nrq_find_indels(b, out->status, out->n_indels, out->n_bases, w.indel_off, w.indels, w.indel_cap,
                              out->ev_base, stream))
    return e;
  mark(1);
  if (int e = nrq_norm_stats(b, p, out->status, out->scale_values, stream)) return e;
  mark(2);
  if (int e = nrq_resegment(b, p, out->scale_values, w.indel_off, w.indels, out->n_bases, w.groups, w.big,
                            w.big_doubles, w.work_counter, out->status, out->new_segs, out->region_samples, nullptr, stream))
    return e;
  mark(3);
  if (int e = nrq_event_table(b, p, out->scale_values, out->n_bases, out->new_segs, out->status, out->norm_mean,
                              out->norm_stdev, out->ev_start, out->ev_length, out->ev_base, stream))
    return e;
  mark(4);
  return 0;
}

int nrq_resquiggle_batch(const nrq_batch* b, const nrq_params* p, const nrq_results* out, void* workspace,
                         int64_t workspace_bytes, void* stream) {
  return resquiggle_batch_impl(b, p, out, workspace, workspace_bytes, stream, nullptr);
}

int nrq_resquiggle_batch_profile(const nrq_batch* b, const nrq_params* p, const nrq_results* out,
                                 void* workspace, int64_t workspace_bytes, void* stream, float* stage_ms) {
  if (!stage_ms) return fail("nrq_resquiggle_batch_profile: stage_ms is NULL");
  thread_local cudaEvent_t ev[5];
  thread_local bool have = false;
  if (!have) {
    for (auto& e : ev)
      if (cudaEventCreate(&e) != cudaSuccess) return fail("cudaEventCreate failed");
    have = true;
  }
  if (int e = resquiggle_batch_impl(b, p, out, workspace, workspace_bytes, stream, ev)) return e;
  cudaError_t ce = cudaEventSynchronize(ev[4]);
  if (ce != cudaSuccess) return fail("nrq_resquiggle_batch_profile: sync", ce);
  for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&stage_ms[i], ev[i], ev[i + 1]);
  return 0;
}

// ------------------------------------------------------------------ host-buffer entry point
struct nrq_handle {
  int device;
  cudaStream_t stream;
  std::vector<std::pair<void*, size_t>> pool;  // grow-only device buffers, by slot
};

static void* slot(nrq_handle* h, size_t idx, size_t bytes, cudaError_t* err) {
  if (h->pool.size() <= idx) h->pool.resize(idx + 1, {nullptr, 0});
  auto& s = h->pool[idx];
  if (s.second < bytes) {
    if (s.first) cudaFree(s.first);
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&s.first, want);
    if (e != cudaSuccess) { s = {nullptr, 0}; *err = e; return nullptr; }
    s.second = want;
  }
  return s.first;
}

nrq_handle* nrq_create(int32_t device) {
  if (cudaSetDevice(device) != cudaSuccess) { fail("nrq_create: cudaSetDevice failed"); return nullptr; }
  nrq_handle* h = new nrq_handle();
  h->device = device;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    fail("nrq_create: stream creation failed");
    return nullptr;
  }
  return h;
}

void nrq_destroy(nrq_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (auto& s : h->pool) if (s.first) cudaFree(s.first);
  cudaStreamDestroy(h->stream);
  delete h;
}

int nrq_resquiggle_batch_host(nrq_handle* h, const nrq_batch* b, const nrq_params* p, const nrq_results* out) {
  if (!h) return fail("handle is NULL");
  if (int e = check_batch(b)) return e;
  if (int e = check_params(p)) return e;
  if (!out) return fail("results is NULL");
  const int R = b->n_reads;
  if (R == 0) return 0;
  cudaError_t ce = cudaSetDevice(h->device);
  if (ce != cudaSuccess) return fail("cudaSetDevice", ce);
  const int64_t n_raw = b->raw_off[R], n_starts = b->starts_off[R], n_align = b->align_off[R];
  cudaStream_t s = h->stream;
  cudaError_t err = cudaSuccess;
  size_t k = 0;
  auto up = [&](const void* src, size_t bytes) -> void* {
    void* d = slot(h, k++, bytes ? bytes : 1, &err);
    if (d && bytes && src) {
      cudaError_t e = cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, s);
      if (e != cudaSuccess) err = e;
    }
    return d;
  };
  nrq_batch d = *b;
  d.norm_signal = nullptr;  // the host entry point always normalises from raw
  d.norm_off = nullptr;
  d.raw = (const int16_t*)up(b->raw, 2 * (size_t)n_raw);
  d.raw_off = (const int64_t*)up(b->raw_off, 8 * (size_t)(R + 1));
  d.read_start = (const int32_t*)up(b->read_start, 4 * (size_t)R);
  d.starts = (const int32_t*)up(b->starts, 4 * (size_t)n_starts);
  d.starts_off = (const int64_t*)up(b->starts_off, 8 * (size_t)(R + 1));
  d.read_align = (const uint8_t*)up(b->read_align, (size_t)n_align);
  d.genome_align = (const uint8_t*)up(b->genome_align, (size_t)n_align);
  d.align_off = (const int64_t*)up(b->align_off, 8 * (size_t)(R + 1));
  d.scale_values_in = b->scale_values_in ? (const double*)up(b->scale_values_in, 32 * (size_t)R) : nullptr;
  if (!b->scale_values_in) ++k;
  nrq_results o;
  o.status = (int32_t*)up(nullptr, 4 * (size_t)R);
  o.n_bases = (int32_t*)up(nullptr, 4 * (size_t)R);
  o.n_indels = (int32_t*)up(nullptr, 4 * (size_t)R);
  o.scale_values = (double*)up(nullptr, 32 * (size_t)R);
  o.new_segs = (int32_t*)up(nullptr, 4 * (size_t)(n_align + R));
  o.norm_mean = (double*)up(nullptr, 8 * (size_t)n_align);
  o.norm_stdev = (double*)up(nullptr, 8 * (size_t)n_align);
  o.ev_start = (uint32_t*)up(nullptr, 4 * (size_t)n_align);
  o.ev_length = (uint32_t*)up(nullptr, 4 * (size_t)n_align);
  o.ev_base = (uint8_t*)up(nullptr, (size_t)n_align);
  o.region_samples = (int32_t*)up(nullptr, 4 * (size_t)R);
  int64_t ws_bytes = nrq_workspace_bytes(R, n_align, p);
  void* ws = up(nullptr, (size_t)ws_bytes);
  if (err != cudaSuccess) return fail("nrq_resquiggle_batch_host: staging", err);
  if (int e = nrq_resquiggle_batch(&d, p, &o, ws, ws_bytes, s)) return e;
  auto down = [&](void* dst, const void* src, size_t bytes) {
    if (dst && bytes) {
      cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s);
      if (e != cudaSuccess) err = e;
    }
  };
  down(out->status, o.status, 4 * (size_t)R);
  down(out->n_bases, o.n_bases, 4 * (size_t)R);
  down(out->n_indels, o.n_indels, 4 * (size_t)R);
  down(out->scale_values, o.scale_values, 32 * (size_t)R);
  down(out->new_segs, o.new_segs, 4 * (size_t)(n_align + R));
  down(out->norm_mean, o.norm_mean, 8 * (size_t)n_align);
  down(out->norm_stdev, o.norm_stdev, 8 * (size_t)n_align);
  down(out->ev_start, o.ev_start, 4 * (size_t)n_align);
  down(out->ev_length, o.ev_length, 4 * (size_t)n_align);
  down(out->ev_base, o.ev_base, (size_t)n_align);
  down(out->region_samples, o.region_samples, 4 * (size_t)R);
  if (err != cudaSuccess) return fail("nrq_resquiggle_batch_host: readback", err);
  ce = cudaStreamSynchronize(s);
  if (ce != cudaSuccess) return fail("nrq_resquiggle_batch_host: sync", ce);
  return 0;
}

}  // extern "C"
