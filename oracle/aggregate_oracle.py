"""CPU oracle for the genome-anchored aggregation (nanoraw_helper.py:298-444) -- TEST INFRASTRUCTURE ONLY.
Operates on in-memory event columns instead of re-opening FAST5 files; pinned by
tests/golden/aggregate_v1.npz, which the real reference computed (tests/golden/make_aggregate_golden.py)."""
import numpy as np


def get_coverage(reads):
    """reads: [(start, end)] of one (chrom, strand).  nanoraw_helper.py:298-308"""
    max_end = max(e for _, e in reads)
    cov = np.zeros(max_end, dtype=np.int_)
    for s, e in reads:
        cov[s:e] += 1
    return cov


def mean_over_reads(reads, chrm_len, rev_strand):
    """reads: [(start, column values)].  nanoraw_helper.py:327-340 (and :363-386, :404-428): sums / coverage,
    NaN where nothing covers the base."""
    sums = np.zeros(chrm_len)
    cov = np.zeros(chrm_len, dtype=np.int_)
    for start, values in reads:
        if rev_strand:
            values = values[::-1]
        sums[start:start + len(values)] += values
        cov[start:start + len(values)] += 1
    with np.errstate(all='ignore'):
        return sums / cov
