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
 * them).  Reads whose status is not NRQ_ST_OK get four empty chunks.  n_threads <= 1: in the calling thread.
 * Returns 0, or -1 (bad argument / out smaller than nrqio_pack_corrected_bound says). */
int64_t nrqio_pack_corrected_bound(const nrq_batch* b, const nrq_results* res, int32_t first, int32_t count);
int nrqio_pack_corrected(const nrq_batch* b, const nrq_results* res, int32_t first, int32_t count,
                         int32_t n_threads, uint8_t* out, int64_t out_cap, int64_t* chunk_off,
                         int64_t* chunk_len);

#ifdef __cplusplus
}
#endif
#endif /* NRQ_IO_H */
