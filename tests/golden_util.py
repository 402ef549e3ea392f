"""Loader for tests/golden/hotpath_v1.npz (made by tests/golden/make_golden.py
from the real reference)."""
import json
import os

import numpy as np

from nanoraw_b200.synth import SynthRead

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'hotpath_v1.npz')


class GoldenCase(object):
    def __init__(self, meta, arrays):
        self.meta = meta
        self.name = meta['name']
        self.params = meta['params']
        self.a = arrays
        self.read = SynthRead(
            raw=arrays['raw'], read_start_rel_to_raw=meta['read_start_rel_to_raw'],
            starts=arrays['starts'], read_align=arrays['read_align'].tobytes(),
            genome_align=arrays['genome_align'].tobytes())

    @property
    def error(self):
        """error of the compute path: function-level error, else a full-route
        error that is not the FAST5 write quirk"""
        if self.meta['error']:
            return self.meta['error']
        full = self.meta['error_full']
        if full and 'Error writing resquiggle information' not in full:
            return full
        return ''

    def kwargs(self):
        return dict(norm_type=self.params.get('norm_type', 'median'),
                    outlier_thresh=self.params.get('outlier_thresh', 5),
                    num_cpts_limit=self.params.get('num_cpts_limit'),
                    compute_sd=self.params.get('compute_sd', True))


def load_golden():
    z = np.load(GOLDEN)
    info = json.loads(z['meta'].tobytes().decode())
    cases = []
    for m in info['cases']:
        pre = m['key'] + '/'
        arrays = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
        cases.append(GoldenCase(m, arrays))
    return info, cases
