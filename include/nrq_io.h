/*
 * nrq_io.h -- C ABI of libnrqio.so: host-side (no CUDA) byte work of the FAST5 write-back of
 * `nanoraw genome_resquiggle`.
 *
 * The reference writes the corrected read with h5py, every dataset gzip-filtered
 * (nanoraw/resquiggle.py:112-124: create_dataset(..., compression='gzip')).  An HDF5 gzip chunk is a zlib
 * stream (RFC 1950) around raw deflate (RFC 1951); the reference leaves the compression to libhdf5's zlib at
 * level 4, which costs 8 of the 9.4 ms a write-back takes here and finds next to nothing to match in float64
 * event statistics.  nrqio_deflate writes the same container with one dynamic-Huffman block of literals
 * (no match search): any inflate reads it, the size is within a few percent of zlib's on these tables, and it
 * runs at memory speed.  nrqio_pack_corrected builds the four chunks of every read of a result batch on a
 * pool of threads, straight from the host arrays nrq_resquiggle_batch_host filled.
 */
#ifndef NRQ_IO_H
#define NRQ_IO_H

#include <stdint.h>
#include "nrq.h"

#ifdef __cplusplus
extern "C" {
#endif

/* bytes nrqio_deflate may need for n input bytes */
int64_t nrqio_deflate_bound(int64_t n);
/* zlib stream of src[0..n) into dst; returns its length, or -1 if cap < nrqio_deflate_bound(n) */
int64_t nrqio_deflate(const uint8_t* src, int64_t n, uint8_t* dst, int64_t cap);
/* zlib stream src[0..n) -> dst; returns the number of bytes written (<= cap), or -1 if the stream is damaged, ends
 * early, or holds more than cap bytes (nothing outside dst[0..cap) is touched).  Any deflate stream is accepted
 * (stored / fixed / dynamic blocks); the Adler-32 is verified.  About three times zlib's speed on raw signals. */
int64_t nrqio_inflate(const uint8_t* src, int64_t n, uint8_t* dst, int64_t cap);
/* Adler-32 of src[0..n) (the checksum a zlib stream ends with) */
uint32_t nrqio_adler32(const uint8_t* src, int64_t n);

/* The datasets write_new_fast5_group creates for one read (resquiggle.py:112-124), as gzip chunk payloads:
 *   chunk 0  Events            n_bases rows of (norm_mean f64, norm_stdev f64, start u32, length u32, base S1) = 25 B
 *   chunk 1  read_alignment    S1[A]
 *   chunk 2  genome_alignment  S1[A]
 *   chunk 3  read_segments     int64[Br + 1]   (the basecaller's starts_rel_to_read)
 * for reads first .. first + count - 1 of a batch whose arrays are all in HOST memory (b: inputs, res: what
 * nrq_resquiggle_batch_host returned).  chunk_off and chunk_len have 4 * count entries each: chunk j of read
 * first + i is out[chunk_off[4 i + j] .. + chunk_len[4 i + j]) (chunks do not overlap; there may be gaps between
 * them).  Reads whose status is not NRQ_ST_OK get four empty chunks.  res->ev_start may be NULL (not copied back
 * from the device): `start` is then the running sum of `length` from starts[0].  n_threads <= 1: in the calling
 * thread.
 * Returns 0, or -1 (bad argument / out smaller than nrqio_pack_corrected_bound says). */
int64_t nrqio_pack_corrected_bound(const nrq_batch* b, const nrq_results* res, int32_t first, int32_t count);
int nrqio_pack_corrected(const nrq_batch* b, const nrq_results* res, int32_t first, int32_t count,
                         int32_t n_threads, uint8_t* out, int64_t out_cap, int64_t* chunk_off,
                         int64_t* chunk_len);

/* ------------------------------------------------------------------------------------------------------------
 * Front end: what turns a basecalled read and its SAM record into the inputs of the
 * hot path.  Integer / byte work, one read at a time or a whole batch on threads, written into the CSR arrays of
 * nrq_batch directly.  Per-read outcomes are status codes; nrqio_fe_message gives the reference's message.
 * ------------------------------------------------------------------------------------------------------------ */
enum {
  NRQIO_FE_OK = 0,
  NRQIO_FE_EVENTS_BEFORE_RAW = 1,  /* resquiggle.py:856-858 */
  NRQIO_FE_TOO_FEW_SEGMENTS = 2,   /* resquiggle.py:891-894 */
  NRQIO_FE_ZERO_LENGTH = 3,        /* resquiggle.py:895-898 */
  NRQIO_FE_ALL_STAYS = 4,          /* resquiggle.py:774-777 */
  NRQIO_FE_BAD_CIGAR = 5,          /* resquiggle.py:586-588 */
  NRQIO_FE_SEQ_CIGAR_DISAGREE = 6, /* resquiggle.py:639-641 */
  NRQIO_FE_CAPACITY = 7,           /* an output array of the caller is too small */
  NRQIO_FE_BAD_INPUT = 8
};
const char* nrqio_fe_message(int32_t status);

/* Basecall events -> starts_rel_to_read, basecalls, read_start_rel_to_raw: the arithmetic of get_read_data
 * (resquiggle.py:848-898) followed by fix_stay_states (:767-805).
 *   ev_start / ev_length   the Events columns `start`, `length` (seconds), float32 (is_f32 != 0) or float64
 *   model_state            n_events x state_len bytes (the k-mer of each event; its third character is the call)
 *   move                   n_events integers of move_size bytes (1, 2, 4, 8; signed)
 *   albacore_after_1_0     group attribute version > "1.0": boundaries from `length` (:872-877), else from `start`
 *   sampling_rate          channel_info.sampling_rate (an integer, nanoraw_helper.py:164)
 *   raw_start_time         /Raw/Reads/<read>.attrs['start_time']
 * Outputs: starts[*n_starts] (capacity n_events + 1), bases[*n_starts - 1] (capacity n_events), *read_start. */
int nrqio_events_to_starts(const void* ev_start, int32_t start_is_f32, const void* ev_length, int32_t length_is_f32,
                           const uint8_t* model_state, int32_t state_len, const void* move, int32_t move_size,
                           int64_t n_events, int32_t albacore_after_1_0, int64_t sampling_rate,
                           uint64_t raw_start_time, int64_t* starts, uint8_t* bases, int64_t* n_starts,
                           int64_t* read_start);

/* One SAM record -> the two alignment rows in read orientation (parse_sam_record, resquiggle.py:583-660):
 * CIGAR (\d+)([MIDNSHP=X]), flag & 0x10 = reverse strand (CIGAR reversed, sequences reverse-complemented), hard /
 * soft clips counted and trimmed, leading / trailing non-match operations stripped (with the reference's
 * cigar[0][0] quirk at :636), '-' for gaps.  contig: the reference sequence named by the record (rName).
 * read_row / genome_row: capacity row_cap each (the sum of all CIGAR lengths always suffices).
 * clip[2]: bases clipped from the read's start / end; tallies[4]: insertions, deletions, matches, mismatches as
 * fix_all_clipped_bases counts them (:476-485).  Genome position of the alignment = pos - 1 (:658). */
int nrqio_sam_to_rows(const char* cigar, int32_t flag, int64_t pos, const char* seq, int64_t seq_len,
                      const char* contig, int64_t contig_len, char* read_row, char* genome_row, int64_t row_cap,
                      int64_t* n_cols, int64_t* clip, int64_t* tallies);

/* fix_raw_starts_for_clipped_bases (resquiggle.py:447-460), in place */
int nrqio_clip_starts(int64_t* starts, int64_t* n_starts, int64_t* read_start, int64_t clip_start, int64_t clip_end);

/* A batch of reads, SAM records + stay-free starts in, CSR arrays of nrq_batch out (on n_threads threads). */
typedef struct nrqio_fe_read {
  const char* cigar;
  int32_t flag;
  int64_t pos;
  const char* seq;
  int64_t seq_len;
  const char* contig;
  int64_t contig_len;
  const int64_t* starts;  /* starts_rel_to_read after fix_stay_states, n_starts entries */
  int64_t n_starts;
  int64_t read_start;     /* read_start_rel_to_raw after fix_stay_states */
} nrqio_fe_read;
/* Fills, for read i: status[i]; read_start[i]; starts[starts_off[i] .. starts_off[i + 1]) (int32, clip-fixed);
 * read_align / genome_align[align_off[i] .. align_off[i + 1]); clip[2 i ..]; tallies[4 i ..].  starts_off and
 * align_off get n_reads + 1 entries; a read that fails gets empty ranges.  Returns 0, or NRQIO_FE_CAPACITY if
 * starts_cap / align_cap are too small (nothing useful written). */
int nrqio_frontend_batch(const nrqio_fe_read* reads, int32_t n_reads, int32_t n_threads, int32_t* status,
                         int32_t* read_start, int32_t* starts, int64_t starts_cap, int64_t* starts_off,
                         uint8_t* read_align, uint8_t* genome_align, int64_t align_cap, int64_t* align_off,
                         int64_t* clip, int64_t* tallies);

/* ------------------------------------------------------------------------------------------------------------
 * FAST5 files without libhdf5: the two things the re-squiggle worker does to a file,
 * for whole device batches on n_threads threads, straight from / into the arrays of the C ABI above.
 *   read   /Raw/Reads/<the read>/Signal and the channel_id attributes          (nanoraw/resquiggle.py:327-347)
 *   write  /Analyses/<corrected group>/<subgroup> with its attributes, Alignment/{read_alignment, genome_alignment,
 *          read_segments} and Events                                            (nanoraw/resquiggle.py:80-135)
 * On-disk forms read: superblock 0-3, object headers v1 / v2, symbol-table groups and compact link messages,
 * compact / contiguous / chunked int16 signals (v1 B-tree or single chunk; deflate, shuffle, fletcher32).  The write
 * appends to files with a version 0 / 1 superblock whose <corrected group> is a symbol-table group (what libhdf5
 * writes by default and what nanoraw_b200/hdf5_min.py writes), in the forms hdf5_min.py uses.  Per-file outcomes are
 * status codes; on NRQIO_F5_UNSUPPORTED (a valid file outside this slice) the caller takes the Python reader /
 * writer, which handles more forms, for that file.
 * ------------------------------------------------------------------------------------------------------------ */
enum {
  NRQIO_F5_OK = 0,
  NRQIO_F5_IO = 1,          /* cannot open / read / write the file */
  NRQIO_F5_FORMAT = 2,      /* not an HDF5 file, or damaged */
  NRQIO_F5_UNSUPPORTED = 3, /* valid HDF5 outside the slice implemented here */
  NRQIO_F5_MISSING = 4,     /* a group, dataset or attribute the worker needs is not in the file */
  NRQIO_F5_EXISTS = 5       /* <subgroup> is already there (h5py's create_group raises) */
};

typedef struct nrqio_f5_batch nrqio_f5_batch;

/* Opens the n files strtab + path_off[i] (NUL-terminated names), finds each raw signal and fills n_samples[i],
 * status[i] and channel[4 i ..] = offset, range, digitisation, sampling_rate (channel may be NULL).  The file images
 * stay in *out until nrqio_f5_read_raw_batch has inflated them.  Returns 0, or -1 on a bad argument. */
int nrqio_f5_open_raw_batch(const char* strtab, const int64_t* path_off, int32_t n, int32_t n_threads,
                            nrqio_f5_batch** out, int64_t* n_samples, double* channel, int32_t* status);
/* Signal of file i -> dst[dst_off[i] .. dst_off[i] + n_samples[i]) (e.g. nrq_batch.raw in pinned memory);
 * dst_off[i] < 0 skips the file.  status[i] as above. */
int nrqio_f5_read_raw_batch(nrqio_f5_batch* h, int16_t* dst, const int64_t* dst_off, int32_t n_threads,
                            int32_t* status);
void nrqio_f5_close_batch(nrqio_f5_batch* h);

/* The corrected reads of one result batch, as nrqio_pack_corrected left them. */
typedef struct nrqio_f5_corrected_batch {
  int32_t n;
  int32_t outlier_is_int;       /* outlier_threshold is written as int64 (else float64) */
  const char* corrected_group;  /* e.g. "RawGenomeCorrected_000": exists under /Analyses (prep_fast5, :942-988) */
  const char* subgroup;         /* e.g. "BaseCalled_template": created */
  const char* norm_type;
  double outlier_threshold;
  const char* strtab;           /* NUL-terminated strings the three offset arrays point into */
  const int64_t* path_off;      /* [n] file names */
  const int64_t* chrom_off;     /* [n] mapped_chrom */
  const int64_t* strand_off;    /* [n] mapped_strand */
  const double* scale_values;   /* [4 n] shift, scale, lower_lim, upper_lim */
  const int64_t* ints;          /* [8 n] mapped_start, clipped_bases_start, clipped_bases_end, num_insertions,
                                   num_deletions, num_matches, num_mismatches, read_start_rel_to_raw */
  const int32_t* n_bases;       /* [n] rows of Events */
  const int64_t* n_align;       /* [n] columns of the alignment */
  const int64_t* n_segs;        /* [n] entries of read_segments */
  const uint8_t* chunks;        /* nrqio_pack_corrected: out, chunk_off, chunk_len (4 per read) */
  const int64_t* chunk_off;
  const int64_t* chunk_len;
  const int32_t* skip;          /* [n] or NULL: nonzero = leave this file alone */
} nrqio_f5_corrected_batch;
/* Appends the group of every read to its file; status[i] as above.  Returns 0, or -1 on a bad argument. */
int nrqio_f5_write_corrected_batch(const nrqio_f5_corrected_batch* w, int32_t n_threads, int32_t* status);

#ifdef __cplusplus
}
#endif
#endif /* NRQ_IO_H */
