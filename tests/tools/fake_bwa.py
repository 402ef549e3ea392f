#!/usr/bin/env python
"""Stand-in for `bwa mem -x ont2d -v 1 -t T genome reads.fa` (resquiggle.py:700-702) used by the
integration test: prints, for every read in the FASTA, the SAM record stored for it in the JSON file
named by $FAKE_BWA_RECORDS (made by the test from synthetic alignments).  The real mappers are external
programs in the reference too; only their SAM output is consumed."""
import json
import os
import sys

records = json.load(open(os.environ['FAKE_BWA_RECORDS']))
fasta = sys.argv[-1]
fields = ('qName', 'flag', 'rName', 'pos', 'mapq', 'cigar', 'rNext', 'pNext', 'tLen', 'seq', 'qual')
print('@SQ\tSN:synth_chr1\tLN:1')
for line in open(fasta):
    if line.startswith('>'):
        name = line[1:].strip()
        rec = records.get(name)
        if rec is None:
            continue
        rec = dict(rec, qName=name)
        print('\t'.join(rec[k] for k in fields))
