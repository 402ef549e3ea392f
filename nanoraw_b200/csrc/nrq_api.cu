// C ABI of libnrq.so (include/nrq.h).  Unity build: the kernels live in the .cu files included
// below; this file only validates arguments, carves the workspace and launches.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only: ranges cost nothing unless a profiler is attached

#include "nrq_indels.cu"
#include "nrq_norm.cu"
#include "nrq_speculate.cu"
#include "nrq_resegment.cu"
#include "nrq_events.cu"
#include "nrq_aggregate.cu"

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

// NVTX range over a scope: the four stages of a batch and the copy pipeline of the host entry point show up as
// named ranges on the timeline (tracing)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// device-filling grids of the two persistent kernels, worked out once per device (a process may drive several)
struct DeviceGrids {
  bool ready = false;
  int seg = 1;   // k_resegment: CTAs resident on the whole device
  int spec = 1;  // k_speculate
};
constexpr int kMaxDevices = 64;

const DeviceGrids& device_grids() {
  static DeviceGrids grids[kMaxDevices];
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevices) dev = 0;
  std::lock_guard<std::mutex> lock(mu);
  DeviceGrids& g = grids[dev];
  if (!g.ready) {
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nrq::k_resegment, nrq::kSegWarps * 32, 0);
    g.seg = sms * (per_sm < 1 ? 1 : per_sm);
    cudaFuncSetAttribute(nrq::k_speculate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(nrq::SpecSm));
    per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nrq::k_speculate, nrq::kSpecThreads,
                                                  sizeof(nrq::SpecSm));
    g.spec = sms * (per_sm < 1 ? 1 : per_sm);
    g.ready = true;
  }
  return g;
}

int resegment_grid(int n_reads) {
  int cached = device_grids().seg;
  // NRQ_SEG_CTAS_PER_SM (experiments: the walk sharing the SMs with another kernel): fewer resident CTAs per SM
  const char* env = getenv("NRQ_SEG_CTAS_PER_SM");
  const int limit = env ? atoi(env) : 0;
  if (limit > 0) {
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (limit * sms < cached) cached = limit * sms;
  }
  int need = (n_reads + nrq::kSegWarps - 1) / nrq::kSegWarps;
  return need < cached ? (need > 0 ? need : 1) : cached;
}

int speculate_grid(int n_reads) {
  const int cached = device_grids().spec;
  return n_reads < cached ? (n_reads > 0 ? n_reads : 1) : cached;
}

struct Workspace {
  int64_t* indel_off;
  int32_t* work_counter;
  int32_t* indels;
  int32_t* groups;
  int32_t* spec;
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
  w.spec = reinterpret_cast<int32_t*>(b + off);
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
    case NRQ_ST_REGION_TOO_LARGE:
      // (not a reference message: the reference re-segments a region of any size, serially)
      return "Indel region exceeds the re-squiggle workspace capacity (nrq_params.big_region_cap samples; the reference has no such limit).";
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
  // 128-bit row loads need both row buffers on 16-byte boundaries (any cudaMalloc'ed buffer is); otherwise bytes
  const bool wide = (((uintptr_t)b->read_align | (uintptr_t)b->genome_align) & 15u) == 0;
  if (wide) nrq::k_align_count<true><<<grid, nrq::kWarpsPerCta * 32, 0, s>>>(*b, status, n_indels, n_bases);
  else nrq::k_align_count<false><<<grid, nrq::kWarpsPerCta * 32, 0, s>>>(*b, status, n_indels, n_bases);
  NRQ_CHECK_LAUNCH("k_align_count");
  nrq::k_scan_counts<<<1, 1024, 0, s>>>(n_indels, b->n_reads, indel_off);
  NRQ_CHECK_LAUNCH("k_scan_counts");
  if (indels) {
    if (wide)
      nrq::k_align_fill<true><<<grid, nrq::kWarpsPerCta * 32, 0, s>>>(
          *b, status, indel_off, reinterpret_cast<int4*>(indels), indel_capacity, ev_base);
    else
      nrq::k_align_fill<false><<<grid, nrq::kWarpsPerCta * 32, 0, s>>>(
          *b, status, indel_off, reinterpret_cast<int4*>(indels), indel_capacity, ev_base);
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

}  // extern "C"

// `mid` (optional): recorded between the speculative search and the walk
static int resegment_impl(const nrq_batch* b, const nrq_params* p, const double* scale_values,
                          const int64_t* indel_off, const int32_t* indels, const int32_t* n_bases,
                          int32_t* group_scratch, int32_t* spec_scratch, int64_t spec_capacity, double* big_scratch,
                          int64_t big_scratch_doubles, int32_t* work_counter, int32_t* status, int32_t* new_segs,
                          int32_t* region_samples, int32_t* n_groups, void* stream, cudaEvent_t mid) {
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
  cudaError_t e = cudaMemsetAsync(work_counter, 0, 2 * sizeof(int32_t), s);
  if (e != cudaSuccess) return fail("memset work_counter", e);
  // K2a: first change-point search of every indel group, one region per thread (nrq_speculate.cu).  Optional:
  // whatever it does not cover is searched by k_resegment.  debug_flags bit 2 turns it off (tests compare both).
  const bool speculate = spec_scratch && spec_capacity > 0 && !b->norm_signal && p->min_seg_len == nrq::kSpecM && !(p->debug_flags & 4);
  // spec_scratch: an int2 entry per indel, then kSpecCptBytesPerIndel bytes per indel for the found change-points
  int2* spec_entries = reinterpret_cast<int2*>(spec_scratch);
  unsigned char* spec_bytes = reinterpret_cast<unsigned char*>(spec_scratch) + 8 * spec_capacity;
  // debug_flags bit 4 / bit 5 (experiments with the two kernels on different streams): only the first / only the
  // second kernel of the stage (the second then trusts the entries a bit-4 call left in spec_scratch)
  if (speculate && !(p->debug_flags & 32)) {
    nrq::k_speculate<<<speculate_grid(b->n_reads), nrq::kSpecThreads, sizeof(nrq::SpecSm), s>>>(
        *b, *p, scale_values, indel_off, reinterpret_cast<const int4*>(indels), work_counter + 1, status,
        spec_entries, spec_bytes);
    NRQ_CHECK_LAUNCH("k_speculate");
  }
  if (mid) cudaEventRecord(mid, s);
  if (p->debug_flags & 16) return 0;
  nrq::k_resegment<<<grid, nrq::kSegWarps * 32, 0, s>>>(
      *b, *p, scale_values, indel_off, reinterpret_cast<const int4*>(indels), n_bases,
      reinterpret_cast<int4*>(group_scratch), big_scratch, work_counter, status, new_segs, region_samples,
      n_groups, speculate ? spec_entries : nullptr, spec_bytes);
  NRQ_CHECK_LAUNCH("k_resegment");
  return 0;
}

extern "C" {

int nrq_resegment(const nrq_batch* b, const nrq_params* p, const double* scale_values,
                  const int64_t* indel_off, const int32_t* indels, const int32_t* n_bases,
                  int32_t* group_scratch, int32_t* spec_scratch, int64_t spec_capacity, double* big_scratch,
                  int64_t big_scratch_doubles, int32_t* work_counter, int32_t* status, int32_t* new_segs,
                  int32_t* region_samples, int32_t* n_groups, void* stream) {
  return resegment_impl(b, p, scale_values, indel_off, indels, n_bases, group_scratch, spec_scratch, spec_capacity,
                        big_scratch,
                        big_scratch_doubles, work_counter, status, new_segs, region_samples, n_groups, stream,
                        nullptr);
}

int nrq_running_diffs(const double* signal, int32_t n, int32_t min_seg_len, double* scratch, double* out,
                      void* stream) {
  if (!signal || !scratch || !out) return fail("nrq_running_diffs: NULL argument");
  if (n < 0 || min_seg_len < 1) return fail("nrq_running_diffs: bad size");
  if (n - 2 * min_seg_len + 1 <= 0) return 0;
  nrq::k_running_diffs<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(signal, n, min_seg_len, scratch, out);
  NRQ_CHECK_LAUNCH("k_running_diffs");
  return 0;
}

int nrq_genome_aggregate(const int32_t* read_start, const int32_t* read_len, const int64_t* ev_off,
                         const uint8_t* read_rev, int32_t n_track_reads, int32_t max_len, int32_t track_len,
                         const double* norm_mean, const double* norm_stdev, const uint32_t* ev_length,
                         int32_t* coverage, double* mean_of_means, double* mean_of_sds, double* mean_of_lengths,
                         void* stream) {
  if (!coverage || !mean_of_means || !mean_of_sds || !mean_of_lengths)
    return fail("nrq_genome_aggregate: NULL output");
  if (n_track_reads > 0 && (!read_start || !read_len || !ev_off || !read_rev || !norm_mean || !norm_stdev ||
                            !ev_length))
    return fail("nrq_genome_aggregate: NULL input");
  if (track_len <= 0) return 0;
  nrq::k_genome_aggregate<<<(track_len + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      read_start, read_len, ev_off, read_rev, n_track_reads, max_len, track_len, norm_mean, norm_stdev, ev_length,
      coverage, mean_of_means, mean_of_sds, mean_of_lengths);
  NRQ_CHECK_LAUNCH("k_genome_aggregate");
  return 0;
}

int nrq_genome_event_lists(const int32_t* read_start, const int32_t* read_len, const int64_t* ev_off,
                           const uint8_t* read_rev, int32_t n_track_reads, int32_t max_len, int32_t track_len,
                           const double* norm_mean, const int64_t* pos_off, double* values, void* stream) {
  if (!pos_off || !values) return fail("nrq_genome_event_lists: NULL output");
  if (n_track_reads > 0 && (!read_start || !read_len || !ev_off || !read_rev || !norm_mean))
    return fail("nrq_genome_event_lists: NULL input");
  if (track_len <= 0 || n_track_reads <= 0) return 0;
  nrq::k_genome_event_lists<<<(track_len + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      read_start, read_len, ev_off, read_rev, n_track_reads, max_len, track_len, norm_mean, pos_off, values);
  NRQ_CHECK_LAUNCH("k_genome_event_lists");
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
  else if (p->long_read_parts > 1)
    // gridDim.y: parts a long read (> 262 144 samples) is streamed in, one warp each; the warps of the other parts
    // of a shorter read leave at once
    nrq::k_event_table<true><<<dim3((b->n_reads + nrq::kEvWarps - 1) / nrq::kEvWarps,
                                    p->long_read_parts > 64 ? 64 : p->long_read_parts),
                               nrq::kEvWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        *b, *p, scale_values, n_bases, new_segs, status, norm_mean, norm_stdev, ev_start, ev_length);
  else
    nrq::k_event_table<false><<<(b->n_reads + nrq::kEvWarps - 1) / nrq::kEvWarps, nrq::kEvWarps * 32, 0,
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
  int64_t indel_cap = spare / 48;
  Workspace w = carve(workspace, b->n_reads, 0, p, indel_cap);
  if (w.bytes > workspace_bytes) return fail("nrq_resquiggle_batch: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto mark = [&](int i) { if (ev) cudaEventRecord(ev[i], s); };

  NvtxRange batch_range("nrq batch (a1..a7)");
  mark(0);
  nvtxRangePushA("a2 indel scan");
  if (int e = nrq_find_indels(b, out->status, out->n_indels, out->n_bases, w.indel_off, w.indels, w.indel_cap,
                              out->ev_base, stream))
    { nvtxRangePop(); return e; }
  nvtxRangePop();
  mark(1);
  {
    NvtxRange r("a1 order statistics");
    if (int e = nrq_norm_stats(b, p, out->status, out->scale_values, stream)) return e;
  }
  mark(2);
  nvtxRangePushA("a3-a6 re-segmentation");
  if (int e = resegment_impl(b, p, out->scale_values, w.indel_off, w.indels, out->n_bases, w.groups, w.spec, w.indel_cap, w.big,
                             w.big_doubles, w.work_counter, out->status, out->new_segs, out->region_samples, nullptr,
                             stream, ev ? ev[5] : nullptr))
    { nvtxRangePop(); return e; }
  nvtxRangePop();
  mark(3);
  NvtxRange ev_range("a7 event table");
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
  thread_local cudaEvent_t ev[6];  // 0..4 around the stages, 5 between the two kernels of the re-segmentation
  thread_local bool have = false;
  if (!have) {
    for (auto& e : ev)
      if (cudaEventCreate(&e) != cudaSuccess) return fail("cudaEventCreate failed");
    have = true;
  }
  if (int e = resquiggle_batch_impl(b, p, out, workspace, workspace_bytes, stream, ev)) return e;
  cudaError_t ce = cudaEventSynchronize(ev[4]);
  if (ce != cudaSuccess) return fail("nrq_resquiggle_batch_profile: sync", ce);
  cudaEventElapsedTime(&stage_ms[0], ev[0], ev[1]);
  cudaEventElapsedTime(&stage_ms[1], ev[1], ev[2]);
  cudaEventElapsedTime(&stage_ms[2], ev[2], ev[5]);
  cudaEventElapsedTime(&stage_ms[3], ev[5], ev[3]);
  cudaEventElapsedTime(&stage_ms[4], ev[3], ev[4]);
  return 0;
}

// ------------------------------------------------------------------ host-buffer entry point
// The batch is cut into chunks of reads that flow through a three-stage pipeline: one host->device copy
// stream, one kernel stream per slot, one device->host copy stream, chained by events.  Each PCIe direction is
// a FIFO, so chunk c is on the device as early as possible, its kernels overlap chunk c+1's upload and chunk
// c-1's download, and the two directions stay busy together.  A slot is a set of device buffers; a chunk may
// reuse it once the slot's previous download has finished.  Per-chunk CSR offsets are rebased on the host into
// pinned staging.
constexpr int kMaxSlots = 8;
constexpr int kSlotsDefault = 6;  // NRQ_HOST_SLOTS overrides (1..kMaxSlots), read at nrq_create
// Input bytes per chunk: an eighth of the call, within [kChunkBytesMin, kChunkBytesMax]; NRQ_HOST_CHUNK_BYTES
// forces a size.  A chunk's kernels take about 3 ms however small it is (K2 walks a read's indel groups one after
// the other), so slots x chunk >= ~300 MB must be in flight to keep both PCIe directions busy, while a small
// call still wants several chunks to overlap at all.  Measured sweep in profiles/README.md.
constexpr int64_t kChunkBytesMin = 32ll << 20, kChunkBytesMax = 128ll << 20;
constexpr int64_t kReadsPerChunkMin = 192, kChunkBytesLongMax = 768ll << 20;  // long reads: see the submit call

struct nrq_slot {
  cudaStream_t stream = nullptr;  // kernels
  cudaEvent_t up_done = nullptr, k_done = nullptr, done = nullptr;
  std::vector<std::pair<void*, size_t>> dev;  // grow-only device buffers
  int64_t* off_stage = nullptr;               // pinned: 3 rebased offset arrays
  size_t off_stage_elems = 0;
};

constexpr int kTickets = 16;  // completion events of the most recent submissions

struct nrq_handle {
  int device;
  int n_slots = kSlotsDefault;
  cudaStream_t h2d = nullptr, d2h = nullptr;
  nrq_slot slots[kMaxSlots];
  cudaEvent_t ticket[kTickets] = {};
  int64_t next_ticket = 0;
  int64_t chunk_index = 0;  // slots are taken round-robin across calls, so a new call queues behind the last one
};

static void* dev_buf(nrq_slot& sl, size_t idx, size_t bytes, cudaError_t* err) {
  if (sl.dev.size() <= idx) sl.dev.resize(idx + 1, {nullptr, 0});
  auto& s = sl.dev[idx];
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
  if (const char* env = getenv("NRQ_HOST_SLOTS")) {
    int v = atoi(env);
    if (v >= 1 && v <= kMaxSlots) h->n_slots = v;
  }
  bool ok = cudaStreamCreateWithFlags(&h->h2d, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&h->d2h, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < h->n_slots; ++i) {
    nrq_slot& sl = h->slots[i];
    if (!ok || cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&sl.up_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&sl.k_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming) != cudaSuccess) {
      nrq_destroy(h);
      fail("nrq_create: stream / event creation failed");
      return nullptr;
    }
  }
  return h;
}

void nrq_destroy(nrq_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (auto& sl : h->slots) {
    for (auto& s : sl.dev) if (s.first) cudaFree(s.first);
    if (sl.off_stage) cudaFreeHost(sl.off_stage);
    for (cudaEvent_t e : {sl.up_done, sl.k_done, sl.done}) if (e) cudaEventDestroy(e);
    if (sl.stream) cudaStreamDestroy(sl.stream);
  }
  for (cudaEvent_t e : h->ticket) if (e) cudaEventDestroy(e);
  if (h->h2d) cudaStreamDestroy(h->h2d);
  if (h->d2h) cudaStreamDestroy(h->d2h);
  delete h;
}

int64_t nrq_resquiggle_batch_host_submit(nrq_handle* h, const nrq_batch* b, const nrq_params* p,
                                         const nrq_results* out) {
  if (!h) return fail("handle is NULL");
  if (int e = check_batch(b)) return e;
  if (int e = check_params(p)) return e;
  if (!out) return fail("results is NULL");
  const int R = b->n_reads;
  cudaError_t ce = cudaSetDevice(h->device);
  if (ce != cudaSuccess) return fail("cudaSetDevice", ce);
  const int64_t id = h->next_ticket;
  cudaEvent_t& done_ev = h->ticket[id % kTickets];
  // the slot still belongs to ticket id - kTickets: a 17th batch is refused while that one is in flight
  if (done_ev && id >= kTickets && cudaEventQuery(done_ev) == cudaErrorNotReady)
    return fail("nrq_resquiggle_batch_host_submit: 16 batches are outstanding already (wait for the oldest first)");
  if (!done_ev && cudaEventCreateWithFlags(&done_ev, cudaEventDisableTiming) != cudaSuccess)
    return fail("nrq_resquiggle_batch_host_submit: event creation failed");
  ++h->next_ticket;
  if (R == 0) {
    cudaEventRecord(done_ev, h->d2h);
    return id;
  }
  cudaError_t err = cudaSuccess;
  int rc = 0;

  int64_t kChunkBytes = (2 * b->raw_off[R] + 4 * b->starts_off[R] + 2 * b->align_off[R] + 7) / 8;
  kChunkBytes = kChunkBytes < kChunkBytesMin ? kChunkBytesMin : kChunkBytes > kChunkBytesMax ? kChunkBytesMax : kChunkBytes;
  // a chunk's kernels take as long as its slowest read (one warp walks one read: about 30 ms for a 100 kb read), so a
  // chunk of long reads must hold enough of them to be worth that wait
  {
    const int64_t per_read = (2 * b->raw_off[R] + 4 * b->starts_off[R] + 2 * b->align_off[R]) / R;
    int64_t want = kReadsPerChunkMin * per_read;
    if (want > kChunkBytesLongMax) want = kChunkBytesLongMax;
    if (want > kChunkBytes) kChunkBytes = want;
  }
  if (const char* env = getenv("NRQ_HOST_CHUNK_BYTES")) {
    long long v = atoll(env);
    if (v > 0) kChunkBytes = v;
  }
  int64_t& chunk_index = h->chunk_index;
  for (int lo = 0; lo < R && rc == 0 && err == cudaSuccess; ++chunk_index) {
    // ---- next chunk: reads [lo, hi) holding about kChunkBytes of input
    int hi = lo;
    auto in_bytes = [&](int a, int z) {
      return 2 * (b->raw_off[z] - b->raw_off[a]) + 4 * (b->starts_off[z] - b->starts_off[a]) +
             2 * (b->align_off[z] - b->align_off[a]);
    };
    while (hi < R && (hi == lo || in_bytes(lo, hi + 1) <= kChunkBytes)) ++hi;
    const int n = hi - lo;
    NvtxRange chunk_range("host chunk: H2D, kernels, D2H queued");
    nrq_slot& sl = h->slots[chunk_index % h->n_slots];
    cudaStream_t s = sl.stream;
    // the slot's buffers and pinned offset staging are reused below: wait for its previous chunk's download
    ce = cudaEventSynchronize(sl.done);
    if (ce != cudaSuccess) { err = ce; break; }
    const size_t need = 3 * (size_t)(n + 1);
    if (sl.off_stage_elems < need) {
      if (sl.off_stage) cudaFreeHost(sl.off_stage);
      sl.off_stage = nullptr;
      sl.off_stage_elems = 0;
      ce = cudaMallocHost(&sl.off_stage, 8 * (need + need / 4));
      if (ce != cudaSuccess) { err = ce; break; }
      sl.off_stage_elems = need + need / 4;
    }
    int64_t* raw_off = sl.off_stage;
    int64_t* starts_off = raw_off + (n + 1);
    int64_t* align_off = starts_off + (n + 1);
    const int64_t raw0 = b->raw_off[lo], st0 = b->starts_off[lo], al0 = b->align_off[lo];
    for (int i = 0; i <= n; ++i) {
      raw_off[i] = b->raw_off[lo + i] - raw0;
      starts_off[i] = b->starts_off[lo + i] - st0;
      align_off[i] = b->align_off[lo + i] - al0;
    }
    const int64_t n_raw = raw_off[n], n_starts = starts_off[n], n_align = align_off[n];

    size_t k = 0;
    auto up = [&](const void* src, size_t bytes) -> void* {
      void* d = dev_buf(sl, k++, bytes ? bytes : 1, &err);
      if (d && bytes && src) {
        cudaError_t e = cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, h->h2d);
        if (e != cudaSuccess) err = e;
      }
      return d;
    };
    nrq_batch d = *b;
    d.n_reads = n;
    d.norm_signal = nullptr;  // the host entry point always normalises from raw
    d.norm_off = nullptr;
    d.raw = (const int16_t*)up(b->raw + raw0, 2 * (size_t)n_raw);
    d.raw_off = (const int64_t*)up(raw_off, 8 * (size_t)(n + 1));
    d.read_start = (const int32_t*)up(b->read_start + lo, 4 * (size_t)n);
    d.starts = (const int32_t*)up(b->starts + st0, 4 * (size_t)n_starts);
    d.starts_off = (const int64_t*)up(starts_off, 8 * (size_t)(n + 1));
    d.read_align = (const uint8_t*)up(b->read_align + al0, (size_t)n_align);
    d.genome_align = (const uint8_t*)up(b->genome_align + al0, (size_t)n_align);
    d.align_off = (const int64_t*)up(align_off, 8 * (size_t)(n + 1));
    d.scale_values_in = b->scale_values_in ? (const double*)up(b->scale_values_in + 4 * (size_t)lo, 32 * (size_t)n)
                                           : nullptr;
    if (!b->scale_values_in) ++k;
    nrq_results o;
    o.status = (int32_t*)up(nullptr, 4 * (size_t)n);
    o.n_bases = (int32_t*)up(nullptr, 4 * (size_t)n);
    o.n_indels = (int32_t*)up(nullptr, 4 * (size_t)n);
    o.scale_values = (double*)up(nullptr, 32 * (size_t)n);
    o.new_segs = (int32_t*)up(nullptr, 4 * (size_t)(n_align + n));
    o.norm_mean = (double*)up(nullptr, 8 * (size_t)n_align);
    o.norm_stdev = (double*)up(nullptr, 8 * (size_t)n_align);
    o.ev_start = (uint32_t*)up(nullptr, 4 * (size_t)n_align);
    o.ev_length = (uint32_t*)up(nullptr, 4 * (size_t)n_align);
    o.ev_base = (uint8_t*)up(nullptr, (size_t)n_align);
    o.region_samples = (int32_t*)up(nullptr, 4 * (size_t)n);
    // worst case: every alignment column starts an indel run, so no read can hit NRQ_ST_WORKSPACE here
    int64_t ws_bytes = carve(nullptr, n, n_align, p, n_align + n + 64).bytes + 2048;
    void* ws = up(nullptr, (size_t)ws_bytes);
    if (err != cudaSuccess) break;
    cudaEventRecord(sl.up_done, h->h2d);
    cudaStreamWaitEvent(s, sl.up_done, 0);
    nrq_params pp = *p;
    if (pp.long_read_parts == 0) {  // the read lengths are in host memory here: parts only where a long read is
      int64_t longest = 0;
      for (int i = 0; i < n; ++i) longest = std::max<int64_t>(longest, raw_off[i + 1] - raw_off[i]);
      pp.long_read_parts = longest > 262144 ? 16 : 1;
    }
    rc = nrq_resquiggle_batch(&d, &pp, &o, ws, ws_bytes, s);
    if (rc) break;
    cudaEventRecord(sl.k_done, s);
    cudaStreamWaitEvent(h->d2h, sl.k_done, 0);
    auto down = [&](void* dst, const void* src, size_t bytes) {
      if (dst && bytes) {
        cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->d2h);
        if (e != cudaSuccess) err = e;
      }
    };
    auto at = [](auto* base, int64_t elems) { return base ? base + elems : base; };
    down(at(out->status, lo), o.status, 4 * (size_t)n);
    down(at(out->n_bases, lo), o.n_bases, 4 * (size_t)n);
    down(at(out->n_indels, lo), o.n_indels, 4 * (size_t)n);
    down(at(out->scale_values, 4 * (int64_t)lo), o.scale_values, 32 * (size_t)n);
    down(at(out->new_segs, al0 + lo), o.new_segs, 4 * (size_t)(n_align + n));
    down(at(out->norm_mean, al0), o.norm_mean, 8 * (size_t)n_align);
    down(at(out->norm_stdev, al0), o.norm_stdev, 8 * (size_t)n_align);
    down(at(out->ev_start, al0), o.ev_start, 4 * (size_t)n_align);
    down(at(out->ev_length, al0), o.ev_length, 4 * (size_t)n_align);
    down(at(out->ev_base, al0), o.ev_base, (size_t)n_align);
    down(at(out->region_samples, lo), o.region_samples, 4 * (size_t)n);
    cudaEventRecord(sl.done, h->d2h);
    lo = hi;
  }
  if (rc || err != cudaSuccess) {  // leave nothing of a failed submission in flight
    for (int i = -2; i < h->n_slots; ++i)
      cudaStreamSynchronize(i == -2 ? h->h2d : i == -1 ? h->d2h : h->slots[i].stream);
    if (rc) return rc;
    return fail("nrq_resquiggle_batch_host", err);
  }
  // every download of this batch is queued on the one device->host stream: its end marks the batch's end
  ce = cudaEventRecord(done_ev, h->d2h);
  if (ce != cudaSuccess) return fail("nrq_resquiggle_batch_host_submit: event record", ce);
  return id;
}

int nrq_resquiggle_batch_host_wait(nrq_handle* h, int64_t ticket) {
  if (!h) return fail("handle is NULL");
  if (ticket < 0 || ticket >= h->next_ticket) return fail("nrq_resquiggle_batch_host_wait: unknown ticket");
  // An older ticket's event slot has been taken over by a later batch; that batch was recorded later on the same
  // FIFO download stream, so waiting for it covers the older one as well.
  cudaError_t ce = cudaSetDevice(h->device);
  if (ce == cudaSuccess) ce = cudaEventSynchronize(h->ticket[ticket % kTickets]);
  if (ce != cudaSuccess) return fail("nrq_resquiggle_batch_host_wait", ce);
  return 0;
}

int nrq_resquiggle_batch_host(nrq_handle* h, const nrq_batch* b, const nrq_params* p, const nrq_results* out) {
  const int64_t ticket = nrq_resquiggle_batch_host_submit(h, b, p, out);
  if (ticket < 0) return (int)ticket;
  return nrq_resquiggle_batch_host_wait(h, ticket);
}

}  // extern "C"
