"""Command-line options of `nanoraw genome_resquiggle` (reference: option_parsers.py:9-136,
280-283, 302-362).  Flags, destinations, types, defaults and choices are the reference's; the
GPU knobs at the bottom are additive.  The other twelve parsers of the reference belong to
plotting / text export and are out of scope."""
import argparse

NORM_CHOICES = ('median', 'pA', 'pA_raw', 'none')

# (flags, kwargs) in the order the reference's parser adds them
_POSITIONAL = (
    (('fast5_basedir',), dict(help='Directory containing fast5 files.')),
    (('genome_fasta',), dict(help='Path to fasta file for mapping.')),
)
_MAPPER = (
    (('--graphmap-executable',), dict(help='graphmap executable (path or command name).')),
    (('--bwa-mem-executable',), dict(help='bwa executable (path or command name).')),
)
_NORMALIZATION = (
    (('--normalization-type',), dict(
        default='median', choices=NORM_CHOICES,
        help='How raw signal is normalised before the event statistics: "median" shifts by the '
             'read median and scales by the MAD; "pA_raw" uses the channel offset / range / '
             'digitisation; "pA" additionally fits the k-mer model (needs --pore-model-filename); '
             '"none" keeps the 16-bit DAC values. Default: %(default)s')),
    (('--pore-model-filename',), dict(
        help='k-mer model file (level_mean, level_stdv) for the "pA" normalisation.')),
    (('--outlier-threshold',), dict(
        default=5, type=float,
        help='Clip the normalised signal at this many MADs from the median; a negative value '
             'disables clipping. Default: %(default)d')),
)
_FILTERING = (
    (('--timeout',), dict(default=None, type=int,
                          help='Per-read time limit in seconds. Default: none.')),
    (('--cpts-limit',), dict(default=None, type=int,
                             help='Largest number of change-points searched within one indel '
                                  'group. Default: no limit.')),
)
_IO = (
    (('--fast5-pattern',), dict(help='Glob (relative to fast5_basedir) selecting the files; quote '
                                     'it on the command line.')),
    (('--recursive',), dict(default=False, action='store_true',
                            help='Look for *.fast5 one directory level down (same as '
                                 '--fast5-pattern "*/*.fast5").')),
    (('--overwrite',), dict(default=False, action='store_true',
                            help='Replace an existing corrected group (--corrected-group only).')),
    (('--failed-reads-filename',), dict(help='Write failed reads, grouped by error, to this file.')),
)
_FAST5 = (
    (('--corrected-group',), dict(default='RawGenomeCorrected_000',
                                  help='FAST5 group that receives the result. Default: %(default)s')),
    (('--basecall-group',), dict(default='Basecall_1D_000',
                                 help='FAST5 group (under Analyses) holding the basecalls. '
                                      'Default: %(default)s')),
)
_SUBGROUPS = (  # mutually exclusive
    (('--basecall-subgroups',), dict(default=['BaseCalled_template', ], nargs='+',
                                     help='Basecall subgroup(s) to correct. Default: %(default)s')),
    (('--2d',), dict(dest='basecall_subgroups', action='store_const',
                     const=['BaseCalled_template', 'BaseCalled_complement'],
                     help='2D reads: template and complement subgroups.')),
)
_PROCESSES = (
    (('--processes',), dict(default=2, type=int, help='Number of processes. Default: %(default)d')),
    (('--align-processes',), dict(default=1, type=int,
                                  help='Processes that align and parse basecalls (each loads the '
                                       'genome). Default: %(default)d')),
    (('--align-threads-per-process',), dict(
        type=int, help='Mapper threads per alignment process. '
                       'Default: [--processes] / (2 * [--align-processes])')),
    (('--resquiggle-processes',), dict(
        type=int, help='Re-squiggle (GPU) worker processes. Default: [--processes] / 2')),
)
_MISC = (
    (('--alignment-batch-size',), dict(default=500, type=int,
                                       help='Reads per mapper call. Default: %(default)d')),
    (('--skip-event-stdev',), dict(default=False, action='store_true',
                                   help='Do not compute the per-base standard deviation.')),
    (('--quiet', '-q'), dict(default=False, action='store_true', help="Don't print status.")),
    (('--help', '-h'), dict(action='help', help='Print this help message and exit')),
)


def get_resquiggle_parser():
    """option_parsers.py:313-362: same option groups, in the same order."""
    parser = argparse.ArgumentParser(
        description='Re-segment raw nanopore signal to match a genome alignment and store '
                    'the corrected events in the FAST5 files.',
        add_help=False)
    layout = (('Required Arguments', _POSITIONAL, None),
              ('Mapper Arguments (One mapper is required)', _MAPPER, None),
              ('Read Filtering Arguments', _FILTERING, None),
              ('Read Normalization Arguments', _NORMALIZATION, None),
              ('Input/Output Arguments', _IO, None),
              ('FAST5 Data Arguments', _FAST5, _SUBGROUPS),
              ('Multiprocessing Arguments', _PROCESSES, None),
              ('Miscellaneous Arguments', _MISC, None))
    for title, options, exclusive in layout:
        grp = parser.add_argument_group(title)
        for flags, kw in options:
            grp.add_argument(*flags, **kw)
        if exclusive:
            mx = grp.add_mutually_exclusive_group()
            for flags, kw in exclusive:
                mx.add_argument(*flags, **kw)
    return parser
