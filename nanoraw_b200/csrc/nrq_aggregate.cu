// f-3 -- genome-anchored aggregation of corrected event tables (nanoraw/nanoraw_helper.py:298-444):
// per genomic position, the read coverage and the mean over covering reads of the per-base event mean, sd
// and length (get_coverage / get_base_means / get_base_sds / get_base_lengths).
//
// One thread per genomic position of one (chromosome, strand) track.  The track's reads are sorted by mapped
// start, so the reads that can cover position p start in (p - max_len, p]: two binary searches give that range
// and the thread adds their contributions in sorted order -- no atomics, bit-identical from run to run.  (The
// reference adds in file order, so results agree to float64 rounding, not bitwise.)
#include "nrq_common.cuh"

namespace nrq {

// read k of the track: genome positions [start[k], start[k] + n[k]); event row of position p is
// ev_off[k] + (rev[k] ? n[k] - 1 - (p - start[k]) : p - start[k])   (reverse strand: nanoraw_helper.py:322-323)
__global__ void __launch_bounds__(256)
k_genome_aggregate(const int32_t* __restrict__ start, const int32_t* __restrict__ n, const int64_t* __restrict__ ev_off,
                   const uint8_t* __restrict__ rev, int n_track_reads, int max_len, int track_len,
                   const double* __restrict__ norm_mean, const double* __restrict__ norm_stdev,
                   const uint32_t* __restrict__ ev_length, int32_t* __restrict__ coverage,
                   double* __restrict__ mean_of_means, double* __restrict__ mean_of_sds,
                   double* __restrict__ mean_of_lengths) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= track_len) return;
  // first read with start > p - max_len, and first read with start > p
  int lo = 0, hi = n_track_reads;
  const int first_start = p - max_len + 1;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (start[mid] < first_start) lo = mid + 1; else hi = mid; }
  const int k0 = lo;
  hi = n_track_reads;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (start[mid] <= p) lo = mid + 1; else hi = mid; }
  const int k1 = lo;
  int cov = 0;
  double sm = 0.0, ss = 0.0, sl = 0.0;
  for (int k = k0; k < k1; ++k) {
    const int d = p - start[k];
    const int len = n[k];
    if (d >= len) continue;
    const int64_t row = ev_off[k] + (rev[k] ? len - 1 - d : d);
    sm = __dadd_rn(sm, norm_mean[row]);
    ss = __dadd_rn(ss, norm_stdev[row]);
    sl = __dadd_rn(sl, (double)ev_length[row]);
    ++cov;
  }
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  coverage[p] = cov;
  // sums / coverage with numpy's 0 / 0 = nan where nothing covers the base (nanoraw_helper.py:340)
  mean_of_means[p] = cov ? __ddiv_rn(sm, (double)cov) : qnan;
  mean_of_sds[p] = cov ? __ddiv_rn(ss, (double)cov) : qnan;
  mean_of_lengths[p] = cov ? __ddiv_rn(sl, (double)cov) : qnan;
}

// get_reads_events (nanoraw_helper.py:446-484): the event means of every read that covers a position, per
// position.  pos_off = exclusive prefix of the coverage (one entry more than the track); the means of position p go
// to values[pos_off[p] ..), in the order of the track's reads (sorted by mapped start; the reference's order inside
// one position is that of an unstable argsort, i.e. unspecified).
__global__ void __launch_bounds__(256)
k_genome_event_lists(const int32_t* __restrict__ start, const int32_t* __restrict__ n,
                     const int64_t* __restrict__ ev_off, const uint8_t* __restrict__ rev, int n_track_reads,
                     int max_len, int track_len, const double* __restrict__ norm_mean,
                     const int64_t* __restrict__ pos_off, double* __restrict__ values) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= track_len) return;
  int lo = 0, hi = n_track_reads;
  const int first_start = p - max_len + 1;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (start[mid] < first_start) lo = mid + 1; else hi = mid; }
  const int k0 = lo;
  hi = n_track_reads;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (start[mid] <= p) lo = mid + 1; else hi = mid; }
  const int k1 = lo;
  double* o = values + pos_off[p];
  const int64_t cap = pos_off[p + 1] - pos_off[p];
  int64_t w = 0;
  for (int k = k0; k < k1 && w < cap; ++k) {
    const int d = p - start[k];
    const int len = n[k];
    if (d >= len) continue;
    o[w++] = norm_mean[ev_off[k] + (rev[k] ? len - 1 - d : d)];
  }
}

}  // namespace nrq
