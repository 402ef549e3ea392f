"""User-visible strings of the genome_resquiggle path, kept byte-identical to the reference so
that failed-read summaries (resquiggle.py:1189-1196) group the same way.  Each constant cites
the reference line that owns the text."""

# hot path -------------------------------------------------------------------------------
EVENTS = 'Error computing new events.'  # resquiggle.py:77
OPEN_FOR_WRITE = (  # :82-86
    'Error opening file for new group writing. This should have been caught during the '
    'alignment phase. Check that there are no other nanoraw processes or processes accessing '
    'these HDF5 files running simultaneously.')
WRITE_BACK = 'Error writing resquiggle information back into fast5 file.'  # :126-127
CLOSE_AFTER_WRITE = 'Error closing fast5 file after writing resquiggle information.'  # :133-135
OPEN_FOR_RESQUIGGLE = (  # :343-347
    'Error opening file for re-squiggle. This should have been caught during the alignment '
    'phase. Check that there are no other nanoraw processes or processes accessing these HDF5 '
    'files running simultaneously.')

# front end ------------------------------------------------------------------------------
BAD_CIGAR = 'Invalid cigar string produced.'  # :589
SEQ_CIGAR_DISAGREE = 'Read sequence from SAM and cooresponding cigar string do not agree.'  # :641-642
NO_ALIGNMENT = 'Alignment not produced. Potentially failed to locate BWA index files.'  # :684-685
MAPPER_FAILED = (  # :747-749 (the missing space after "installed." is the reference's)
    'Problem running/parsing genome mapper. Ensure you have a compatible version installed.'
    'Potentially failed to locate BWA index files.')
MAPPER_UNSUPPORTED = 'Mapper not supported.'  # :733
FORMAT_UNSUPPORTED = 'Mapper output type not supported.'  # :760
ALL_STAYS = 'Read is composed entirely of stay model states and cannot be processed'  # :775-777
OPEN_FOR_ALIGNMENT = (  # :811-815
    'Error opening file for alignment. This should have been caught during the HDF5 prep '
    'phase. Check that there are no other nanoraw processes or processes accessing these HDF5 '
    'files running simultaneously.')
NO_EVENTS = (  # :827-830
    'No events or corrupted events in file. Likely a segmentation error or mis-specified '
    'basecall-subgroups (--2d?).')
NO_RAW = 'Raw data is not stored in Raw/Reads/Read_[read#] so new segments cannot be identified.'  # :835-837
CHANNEL_INFO = 'Error getting channel information and closing fast5 file.'  # :843-844
EVENTS_BEFORE_RAW = 'Events appear to start before the raw signal.'  # :857-858
TOO_FEW_SEGMENTS = 'One or no segments or signal present in read.'  # :894-895
ZERO_LENGTH_INPUT = 'Zero length event present in input data.'  # :897-898

# FAST5 preparation ----------------------------------------------------------------------
NOT_IN_PLACE = 'Not currently implementing new hdf5 file writing.'  # :947
NOT_WRITABLE = 'FAST5 file is not writable.'  # :951
NO_BASECALLS = (  # :955-961
    'FAST5 basecall-group or Analyses group does not exist. Likely these reads either have '
    'not been basecalled (with event storage) or are mux scan files. Also check '
    '--basecall-group and --basecall-subgroups arguments against these files.')
ALREADY_CORRECTED = 'Raw genome corrected data exists for this read and --overwrite is not set.'  # :964-966
CORRUPT_FILE = 'Error opening file. Likely a corrupted file.'  # :968
CREATE_GROUP = 'Error creating new fast5 group or writing to fast5 file.'  # :984-985

# command line ---------------------------------------------------------------------------
BANNER = '*' * 60
NEED_MAPPER = 'ERROR: Must provide either a graphmap or bwa-mem executable.'  # :1123-1125
DIR_INACCESSIBLE = ('ERROR: One of the directories does not appear to be accessible. Check '
                    'directory permissions.')  # :1145-1148
NO_FILES = (  # :1151-1156
    'ERROR: No files identified in the specified directories. If files are stored in '
    'subdirectories (e.g. 0/, 1/, ...) specify the --recursive or --fast5-pattern "*/*.fast5" '
    'options appropriately.')
