"""Run the REAL reference hot path (Python 2 source) under Python 3 -- test infrastructure.

/root/reference/nanoraw is Python 2 only and imports h5py, neither of which
exists in this image.  This module loads ``nanoraw_helper.py`` and
``resquiggle.py`` from where they lie, applies a handful of mechanical,
semantics-preserving source rewrites IN MEMORY (nothing from the reference is
copied into this repository), and executes them with

* Python-2 builtins restored in the module namespace (``zip``/``map``/``filter``
  returning lists, so ``zip(...)[0]`` works as written),
* ``h5py`` replaced by the in-memory store of ``nanoraw_b200.fast5_io``,
* ``distutils.version.LooseVersion`` and ``option_parsers`` stubbed.

Rewrites (all purely syntactic):
  ``raise X, expr``  ->  ``raise X(expr)``        (tokenizer based)
  ``import Queue``   ->  ``import queue as Queue``
  ``izip``           ->  ``zip``  (and dropped from the itertools import)
  ``.iteritems()``   ->  ``.items()``;  ``xrange`` -> ``range``
Optional, explicit switch (never silent):
  ``stable_sort=True`` rewrites ``running_diffs.argsort()`` (resquiggle.py:246)
  to ``running_diffs.argsort(kind='stable')`` -- the deterministic tie order the
  CUDA path is held to.

The result is the reference's own code computing the answers: it is what pins
``oracle/resquiggle_oracle.py`` and generates ``tests/golden/*.npz``
(``tests/golden/make_golden.py``).  It only works where /root/reference exists.
"""
import builtins
import io
import os
import sys
import tokenize
import types
import warnings

import numpy as np

REFERENCE_DIR = os.environ.get('NANORAW_REFERENCE', '/root/reference/nanoraw')


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, 'resquiggle.py'))


def _rewrite_raise(src):
    """``raise Name, expr`` -> ``raise Name(expr)`` using token positions."""
    lines = src.splitlines(keepends=True)
    offs = [0]
    for ln in lines:
        offs.append(offs[-1] + len(ln))

    def absolute(pos):
        return offs[pos[0] - 1] + pos[1]

    toks = list(tokenize.generate_tokens(io.StringIO(src).readline))
    edits = []  # (offset, n_delete, insert)
    i = 0
    skip = (tokenize.COMMENT, tokenize.NL, tokenize.NEWLINE,
            tokenize.INDENT, tokenize.DEDENT)
    while i < len(toks):
        t = toks[i]
        if t.type == tokenize.NAME and t.string == 'raise' and \
                i + 2 < len(toks) and toks[i + 1].type == tokenize.NAME and \
                toks[i + 2].type == tokenize.OP and toks[i + 2].string == ',':
            comma = toks[i + 2]
            j = i + 3
            last = comma
            while toks[j].type != tokenize.NEWLINE:
                if toks[j].type not in skip:
                    last = toks[j]
                j += 1
            edits.append((absolute(comma.start), 1, '('))
            edits.append((absolute(last.end), 0, ')'))
            i = j
        i += 1
    for off, ndel, ins in sorted(edits, reverse=True):
        src = src[:off] + ins + src[off + ndel:]
    return src


def _py3_source(src, stable_sort):
    src = _rewrite_raise(src)
    src = src.replace('import Queue\n', 'import queue as Queue\n')
    src = src.replace('from itertools import groupby, izip, repeat',
                      'from itertools import groupby, repeat')
    src = src.replace('from itertools import izip\n', '\n')
    src = src.replace('izip(', 'zip(')
    src = src.replace('.iteritems()', '.items()')
    src = src.replace('xrange(', 'range(')
    if stable_sort:
        assert 'running_diffs.argsort()' in src
        src = src.replace('running_diffs.argsort()',
                          "running_diffs.argsort(kind='stable')")
    return src


def _py2_builtins(ns):
    ns['zip'] = lambda *a: list(builtins.zip(*a))
    ns['map'] = lambda f, *a: list(builtins.map(f, *a))
    ns['filter'] = lambda f, a: list(builtins.filter(f, a))


class _LooseVersion(object):
    def __init__(self, s):
        if isinstance(s, bytes):
            s = s.decode()
        self.parts = [int(p) if p.isdigit() else p for p in str(s).split('.')]

    def __gt__(self, other):
        return self.parts > other.parts

    def __lt__(self, other):
        return self.parts < other.parts


def _fake_h5py():
    from nanoraw_b200 import fast5_io
    mod = types.ModuleType('h5py')
    mod.File = fast5_io.MemFile
    return mod


def load_reference(stable_sort=False):
    """Returns (resquiggle_module, nanoraw_helper_module) built from the
    reference sources.  numpy's error state is restored afterwards (the
    reference sets np.seterr(all='raise') at import, resquiggle.py:8); callers
    should wrap calls in ``np.errstate(all='raise')`` to keep its semantics."""
    if not reference_available():
        raise RuntimeError('reference sources not found under ' + REFERENCE_DIR)
    saved_err = np.geterr()
    saved_mods = {k: sys.modules.get(k) for k in (
        'h5py', 'distutils', 'distutils.version', 'option_parsers',
        'nanoraw_helper')}
    try:
        sys.modules['h5py'] = _fake_h5py()
        dist = types.ModuleType('distutils')
        dist_v = types.ModuleType('distutils.version')
        dist_v.LooseVersion = _LooseVersion
        dist.version = dist_v
        sys.modules['distutils'] = dist
        sys.modules['distutils.version'] = dist_v
        sys.modules['option_parsers'] = types.ModuleType('option_parsers')

        mods = []
        for name in ('nanoraw_helper', 'resquiggle'):
            path = os.path.join(REFERENCE_DIR, name + '.py')
            with open(path) as fp:
                src = _py3_source(fp.read(), stable_sort and name == 'resquiggle')
            mod = types.ModuleType('nanoraw_ref_' + name)
            mod.__file__ = path
            _py2_builtins(mod.__dict__)
            with warnings.catch_warnings():
                warnings.simplefilter('ignore', SyntaxWarning)
                code = compile(src, path, 'exec')
            exec(code, mod.__dict__)
            if name == 'nanoraw_helper':
                sys.modules['nanoraw_helper'] = mod
            mods.append(mod)
        return mods[1], mods[0]
    finally:
        np.seterr(**saved_err)
        for k, v in saved_mods.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def run_hot_path(ref_rsq, ref_nh, read, norm_type='median', outlier_thresh=5,
                 timeout=None, num_cpts_limit=None, compute_sd=True,
                 channel_info=None, min_event_obs=4):
    """normalize_raw_signal -> get_indel_groups -> splice -> event stats, each
    step executed by the reference's own functions (the splice and event-table
    lines live inside resquiggle_read / write_new_fast5_group, which need a
    FAST5 file: see run_resquiggle_read for that route)."""
    with np.errstate(all='raise'):
        norm_signal, scale_values = ref_nh.normalize_raw_signal(
            read.raw, read.read_start_rel_to_raw, read.starts[-1], norm_type,
            channel_info, outlier_thresh)
        groups = ref_rsq.get_indel_groups(
            read.align_vals, read.starts, norm_signal, min_event_obs, timeout,
            num_cpts_limit)
    return norm_signal, scale_values, groups


def make_mem_fast5(name, read, channel=None, basecall_group='Basecall_1D_000',
                   corrected_group='RawGenomeCorrected_000'):
    """A minimal FAST5 in the in-memory store holding what resquiggle_read opens
    (resquiggle.py:328-341, nanoraw_helper.py:154-166) and what prep_fast5
    creates (resquiggle.py:980-982)."""
    from nanoraw_b200 import fast5_io
    f = fast5_io.MemFile(name, 'w')
    rd = f.create_group('Raw/Reads/Read_1')
    rd.create_dataset('Signal', data=np.asarray(read.raw, dtype=np.int16))
    rd.attrs['read_id'] = read.read_id
    rd.attrs['start_time'] = np.uint64(0)
    ch = f.create_group('UniqueGlobalKey/channel_id')
    channel = channel or {}
    ch.attrs['offset'] = np.float64(channel.get('offset', 10.0))
    ch.attrs['range'] = np.float64(channel.get('range', 1400.0))
    ch.attrs['digitisation'] = np.float64(channel.get('digitisation', 8192.0))
    ch.attrs['channel_number'] = channel.get('channel_number', '1')
    ch.attrs['sampling_rate'] = np.float64(channel.get('sampling_rate', 4000.0))
    f.create_group('Analyses/' + basecall_group)
    corr = f.create_group('Analyses/' + corrected_group)
    corr.attrs['nanoraw_version'] = '0.5'
    corr.attrs['basecall_group'] = basecall_group
    f.close()
    return name


def run_resquiggle_read(ref_rsq, name, read, norm_type='median',
                        outlier_thresh=5, timeout=None, num_cpts_limit=None,
                        compute_sd=True, subgroup='BaseCalled_template',
                        basecall_group='Basecall_1D_000',
                        corrected_group='RawGenomeCorrected_000'):
    """The reference's resquiggle_read (resquiggle.py:318-401) end to end against
    the in-memory FAST5; returns the written corrected subgroup."""
    from nanoraw_b200 import fast5_io
    n_ins, n_del, n_match, n_mm = read.tallies()
    info = ref_rsq.readInfo(read.read_id, subgroup,
                            read.extras.get('clip_start', 0),
                            read.extras.get('clip_end', 0),
                            n_ins, n_del, n_match, n_mm)
    loc = ref_rsq.genomeLoc(read.genome_start, read.strand, read.chrom)
    with np.errstate(all='raise'):
        ref_rsq.resquiggle_read(
            name, read.read_start_rel_to_raw, read.starts, norm_type,
            outlier_thresh, read.align_vals, timeout, num_cpts_limit, loc, info,
            basecall_group, corrected_group, compute_sd, None)
    f = fast5_io.MemFile(name, 'r')
    return f['/Analyses/' + corrected_group + '/' + subgroup]
