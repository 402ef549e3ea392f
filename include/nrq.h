/*
 * nrq.h -- C ABI of libnrq.so: B200 (sm_100a) kernels for the per-read hot path of
 * `nanoraw genome_resquiggle`.
 *
 * The reference has no FFI: the path is plain Python called from one place,
 * resquiggle_worker -> resquiggle_read (nanoraw/resquiggle.py:403-440, 318-401).
 * Each entry point below states which reference lines it replaces.  A whole
 * batch of reads is processed per call; reads are packed in ragged CSR buffers
 * (SURVEY.md section 7 step 1).  All pointers in the *device* entry points are
 * device pointers; the caller owns every buffer, the library owns nothing
 * between calls.  `stream` is a cudaStream_t passed as void*.
 *
 * Return value: 0 on success, negative on an infrastructure error (bad argument,
 * CUDA error; see nrq_last_error()).  Per-read outcomes never fail a call: they
 * are reported in status[r] (NRQ_ST_*), mirroring the reference's per-read
 * exception -> failed_reads_q routing (resquiggle.py:434-438).
 *
 * Tie order of the change-point selection: larger running difference first,
 * larger index first among equal values == argsort(kind='stable')[::-1]
 * (the reference's default argsort, resquiggle.py:246, is not reproducible
 * across numpy builds).
 */
#ifndef NRQ_H
#define NRQ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRQ_VERSION 1

/* ---- per-read status codes (status[r]) and the reference message each maps to ---- */
enum {
  NRQ_ST_OK = 0,
  NRQ_ST_CPTS_LIMIT = 1,     /* resquiggle.py:234-236 */
  NRQ_ST_TIMEOUT = 2,        /* resquiggle.py:288-289 */
  NRQ_ST_ZERO_LEN = 3,       /* resquiggle.py:373-375 */
  NRQ_ST_NEG_START = 4,      /* resquiggle.py:376-378 */
  NRQ_ST_PAST_END = 5,       /* resquiggle.py:379-381 */
  NRQ_ST_SEQ_MISMATCH = 6,   /* resquiggle.py:385-387 */
  NRQ_ST_FPE_DIVIDE = 7,     /* np.seterr(all='raise'), resquiggle.py:8 + nanoraw_helper.py:248: scale == 0 */
  NRQ_ST_FPE_INVALID = 8,    /* same, 0/0 only (constant signal) */
  NRQ_ST_UNSATISFIABLE = 9,  /* region spans the whole read and still fails: the reference never returns (resquiggle.py:210-215, 262-273) */
  NRQ_ST_REGION_TOO_LARGE = 10, /* region larger than the workspace's big-region capacity (library limit) */
  NRQ_ST_BAD_ALIGNMENT = 11, /* alignment rows violate what parse_sam_record guarantees (resquiggle.py:623-637) */
  NRQ_ST_BAD_INPUT = 12,     /* empty signal / inconsistent CSR lengths */
  NRQ_ST_WORKSPACE = 13      /* the caller's indel workspace was too small for this read: re-run it with a larger one */
};

/* ---- normalisation types (nanoraw_helper.py:25, option_parsers.py:77-90) ---- */
enum {
  NRQ_NORM_MEDIAN = 0, /* shift = median, scale = MAD, computed on device */
  NRQ_NORM_GIVEN = 1   /* shift/scale supplied per read: 'none' (0,1), 'pA_raw', 'pA', or explicit shift/scale */
};

/* Options shared by all reads of a batch (the CLI-level arguments of resquiggle_read). */
typedef struct nrq_params {
  int32_t norm_type;        /* NRQ_NORM_* */
  int32_t use_outlier;      /* 1: clip at med +- mad*outlier_thresh (nanoraw_helper.py:252-262); 0: outlier_thresh is None */
  double  outlier_thresh;
  int32_t given_limits;     /* 1: lower_lim/upper_lim supplied per read in scale_values_in (plot path, nanoraw_helper.py:250-251) */
  int32_t compute_sd;       /* 0 == --skip-event-stdev: norm_stdev is NaN (resquiggle.py:63) */
  int32_t num_cpts_limit;   /* --cpts-limit; < 0: no limit */
  int32_t min_seg_len;      /* min_event_obs, 4 in the reference (resquiggle.py:323) */
  int64_t timeout_ns;       /* --timeout on the device clock; <= 0: none */
  int32_t big_region_cap;   /* samples per warp of global scratch for regions that do not fit shared memory */
  int32_t long_read_parts;  /* parts (<= 64) in which nrq_event_table streams a read of more than 262 144 samples, one
                               warp each (BASELINE configs[3]); <= 1: whole reads only.  0 lets the host entry points,
                               which see the read lengths, decide per chunk; the device entry points treat 0 as 1. */
  int32_t debug_flags;      /* bit0: force the general (bisection) order-statistic path in nrq_norm_stats;
                               bit1: event statistics in numpy's float64 summation order (bit-identical to
                               numpy, slower) instead of exact integer prefix sums (within a few ulps);
                               bit2: no speculative first search in nrq_resegment (walk only; same results, slower);
                               bit4 / bit5 (experiments with the stage's two kernels on different streams,
                               profiles/overlap_probe2.py): nrq_resegment runs only its first kernel (the speculative
                               searches) / only its second (the walk, trusting the entries a bit-4 call over the
                               same reads left in spec_scratch) */
} nrq_params;

/* Ragged batch, inputs.  R = n_reads.
 *   raw          int16[raw_off[R]]      all_raw_signal of every read, back to back (resquiggle.py:332-333)
 *   raw_off      int64[R+1]
 *   read_start   int32[R]               read_start_rel_to_raw
 *   starts       int32[starts_off[R]]   starts_rel_to_read, Br+1 entries per read
 *   starts_off   int64[R+1]
 *   read_align / genome_align  uint8[align_off[R]]  the two rows of alignVals, ASCII, '-' = gap
 *   align_off    int64[R+1]
 *   scale_values_in double[R*4] or NULL  (shift, scale, lower_lim, upper_lim) for NRQ_NORM_GIVEN / given_limits
 * Output capacity is derived from the alignment: a read with A alignment columns
 * owns event rows [align_off[r], align_off[r]+A) and new_segs entries
 * [align_off[r]+r, align_off[r]+r+A+1); n_bases[r] (= Bg) rows are valid. */
typedef struct nrq_batch {
  int32_t n_reads;
  const int16_t* raw;
  const int64_t* raw_off;
  const int32_t* read_start;
  const int32_t* starts;
  const int64_t* starts_off;
  const uint8_t* read_align;
  const uint8_t* genome_align;
  const int64_t* align_off;
  const double*  scale_values_in;
  /* Optional: an already normalised float64 signal per read (norm_off[R+1] offsets).  When non-NULL,
   * nrq_resegment and the exact event-table kernel read it instead of normalising `raw`: this is how the
   * reference's get_indel_groups(alignVals, align_segs, raw_signal, ...) call surface (resquiggle.py:139-141),
   * which receives the normalised signal, is served.  NULL in the batch pipeline. */
  const double*  norm_signal;
  const int64_t* norm_off;
} nrq_batch;

/* Outputs (device).  Event table columns are the fields of the reference's
 * structured dtype (resquiggle.py:69-70), kept as struct-of-arrays on device. */
typedef struct nrq_results {
  int32_t*  status;       /* [R] */
  int32_t*  n_bases;      /* [R]   Bg = genome bases in the alignment = rows of the event table */
  int32_t*  n_indels;     /* [R] */
  double*   scale_values; /* [R*4] shift, scale, lower_lim, upper_lim (NaN where the reference has None) */
  int32_t*  new_segs;     /* [align_off[R] + R] */
  double*   norm_mean;    /* [align_off[R]] */
  double*   norm_stdev;   /* [align_off[R]] */
  uint32_t* ev_start;     /* [align_off[R]] */
  uint32_t* ev_length;    /* [align_off[R]] */
  uint8_t*  ev_base;      /* [align_off[R]] */
  int32_t*  region_samples; /* [R] or NULL: samples covered by the final indel groups (accounting aid) */
} nrq_results;

int  nrq_version(void);
const char* nrq_last_error(void);
/* reference message for a status code ("" for NRQ_ST_OK) */
const char* nrq_status_message(int32_t status);

/* Bytes of device scratch nrq_resquiggle_batch needs for a batch with the given totals. */
int64_t nrq_workspace_bytes(int32_t n_reads, int64_t total_align, const nrq_params* params);

/* a2: get_all_indels (resquiggle.py:142-201) for every read.
 *   out_n_indels[R], out_n_bases[R] always written.
 *   indel_off int64[R+1] (device, written), indels int32[4*capacity]: (start, end, diff, diff_prefix) per indel.
 *   ev_base (optional): the `base` column of the event table = genome row without gaps (resquiggle.py:384).
 * Two phases so the caller can size `indels`: call with indels == NULL to get counts + offsets only.
 * Reads whose indels do not fit indel_capacity get NRQ_ST_WORKSPACE. */
int nrq_find_indels(const nrq_batch* b, int32_t* status, int32_t* out_n_indels, int32_t* out_n_bases,
                    int64_t* indel_off, int32_t* indels, int64_t indel_capacity, uint8_t* ev_base,
                    void* stream);

/* a1 / a1': the order statistics of normalize_raw_signal (nanoraw_helper.py:240-256):
 * writes scale_values[R*4]; sets status for scale == 0 and for signals shorter than starts[-1]. */
int nrq_norm_stats(const nrq_batch* b, const nrq_params* p, int32_t* status, double* scale_values, void* stream);

/* a1: the normalised, winsorised signal itself (nanoraw_helper.py:248, 257-262) as float64,
 * CSR by norm_off[R+1] (device, starts[-1] samples per read).  Not used by the batch pipeline (which
 * never materialises it); exists for the normalize_raw_signal call surface and for tests. */
int nrq_normalize_signal(const nrq_batch* b, const double* scale_values, const int32_t* status,
                         const int64_t* norm_off, double* norm_signal, void* stream);

/* a3-a6: get_indel_groups + splice (resquiggle.py:203-316, 362-387) -> new_segs, status.
 * group_scratch: int32[4*indel_capacity]; spec_scratch: int32[4*spec_capacity] or NULL, spec_capacity >= the
 * number of indels of the batch (scratch of the speculative first search of every group, which runs ahead of the
 * read-by-read walk; NULL = walk only, same results);
 * work_counter: int32[2]; big_scratch: see nrq_workspace_bytes; region_samples and n_groups may be NULL.
 * On return group_scratch holds, per read at int4 index indel_off[r], n_groups[r] records
 * (group_start, group_stop, first indel index, max indel end) = the reference's indelGroupStats. */
int nrq_resegment(const nrq_batch* b, const nrq_params* p, const double* scale_values,
                  const int64_t* indel_off, const int32_t* indels, const int32_t* n_bases,
                  int32_t* group_scratch, int32_t* spec_scratch, int64_t spec_capacity, double* big_scratch,
                  int64_t big_scratch_doubles, int32_t* work_counter, int32_t* status, int32_t* new_segs,
                  int32_t* region_samples, int32_t* n_groups, void* stream);

/* f-4: the running-difference trace of plot_correction (plot_commands.py:225-229) over one window of an already
 * normalised signal (device pointers): out[i] = |2 cs[i+m] - cs[i] - cs[i+2m]|, i in [0, n - 2m], cs the
 * sequential float64 cumulative sum.  scratch: n + 1 doubles. */
int nrq_running_diffs(const double* signal, int32_t n, int32_t min_seg_len, double* scratch, double* out,
                      void* stream);

/* a7: per-base mean / sd / start / length / base (resquiggle.py:60-70). */
int nrq_event_table(const nrq_batch* b, const nrq_params* p, const double* scale_values,
                    const int32_t* n_bases, const int32_t* new_segs, const int32_t* status,
                    double* norm_mean, double* norm_stdev, uint32_t* ev_start, uint32_t* ev_length,
                    uint8_t* ev_base, void* stream);

/* f-3 (downstream of the hot path): genome-anchored aggregation of corrected event tables, the arithmetic of
 * get_coverage / get_base_means / get_base_sds / get_base_lengths (nanoraw_helper.py:298-444), for ONE
 * (chromosome, strand) track.  The track's reads must be sorted by mapped start; read k covers genome positions
 * [read_start[k], read_start[k] + read_len[k]), its event rows start at ev_off[k] in the three event arrays and
 * are taken in reverse order when read_rev[k] != 0.  max_len >= every read_len.  Outputs have track_len entries;
 * the means are NaN where coverage is 0.  All pointers are device pointers. */
int nrq_genome_aggregate(const int32_t* read_start, const int32_t* read_len, const int64_t* ev_off,
                         const uint8_t* read_rev, int32_t n_track_reads, int32_t max_len, int32_t track_len,
                         const double* norm_mean, const double* norm_stdev, const uint32_t* ev_length,
                         int32_t* coverage, double* mean_of_means, double* mean_of_sds, double* mean_of_lengths,
                         void* stream);

/* f-3, the per-position event lists of get_reads_events (nanoraw_helper.py:446-484) for one track laid out as for
 * nrq_genome_aggregate: pos_off int64[track_len + 1] = exclusive prefix sum of that call's coverage; the event means
 * of the reads covering position p are written to values[pos_off[p] .. pos_off[p + 1]) in the order of the track's
 * reads.  All pointers are device pointers. */
int nrq_genome_event_lists(const int32_t* read_start, const int32_t* read_len, const int64_t* ev_off,
                           const uint8_t* read_rev, int32_t n_track_reads, int32_t max_len, int32_t track_len,
                           const double* norm_mean, const int64_t* pos_off, double* values, void* stream);

/* The whole path a1..a7 for a batch resident in device memory (the body of resquiggle_worker's loop,
 * resquiggle.py:424-433, minus FAST5 I/O).  workspace: device, >= nrq_workspace_bytes(). */
int nrq_resquiggle_batch(const nrq_batch* b, const nrq_params* p, const nrq_results* out,
                         void* workspace, int64_t workspace_bytes, void* stream);

/* nrq_resquiggle_batch with CUDA events recorded on `stream` around each stage; waits for the batch and
 * writes the device time in ms of {indel scan (a2), order statistics (a1), re-segmentation (a3-a6): speculative
 * first searches, then the walk, event table (a7)} to stage_ms[5].  Measurement aid for bench.py. */
int nrq_resquiggle_batch_profile(const nrq_batch* b, const nrq_params* p, const nrq_results* out,
                                 void* workspace, int64_t workspace_bytes, void* stream, float* stage_ms);

/* Same, but every pointer in `b` and `out` is a HOST pointer (pinned memory recommended).  The library
 * stages through device buffers it owns (grown on demand, per handle) and returns after the results
 * are back in host memory.  device: CUDA device ordinal.  This is the call a non-Python host binds. */
typedef struct nrq_handle nrq_handle;
nrq_handle* nrq_create(int32_t device);
void nrq_destroy(nrq_handle* h);
int nrq_resquiggle_batch_host(nrq_handle* h, const nrq_batch* b, const nrq_params* p, const nrq_results* out);
/* The same call in two halves, so that a host that keeps reads coming (the re-squiggle worker's loop,
 * resquiggle.py:410-440) never lets the copy / kernel pipeline drain between batches: _submit queues the whole
 * batch and returns a ticket (>= 0; negative = error as above), _wait blocks until that batch's results are in the
 * caller's host arrays.  Batches complete in submission order.  The host buffers of a submitted batch must stay
 * valid and untouched until its ticket has been waited for; at most 16 tickets may be outstanding.  A handle is
 * used from one host thread at a time (one handle per worker; handles are independent of each other). */
int64_t nrq_resquiggle_batch_host_submit(nrq_handle* h, const nrq_batch* b, const nrq_params* p,
                                         const nrq_results* out);
int nrq_resquiggle_batch_host_wait(nrq_handle* h, int64_t ticket);
/* number of kernel launches the library has issued on this process (all entry points) */
int64_t nrq_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NRQ_H */
