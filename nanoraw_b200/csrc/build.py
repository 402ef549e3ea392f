"""Build libnrq.so (sm_100a) in-tree with nvcc.  python -m nanoraw_b200.csrc.build [--force]"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, 'libnrq.so')
SOURCES = ['nrq_api.cu', 'nrq_indels.cu', 'nrq_norm.cu', 'nrq_speculate.cu', 'nrq_resegment.cu', 'nrq_events.cu', 'nrq_aggregate.cu',
           'nrq_common.cuh', os.path.join('..', '..', 'include', 'nrq.h')]

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
    # boundary decisions are float64 and must match numpy bit for bit: no FMA contraction
    '-fmad=false', '--shared', '-Xcompiler', '-fPIC', '-Xptxas', '-v',
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(HERE, s)) > t for s in SOURCES)


IO_LIB = os.path.join(PKG, 'libnrqio.so')
IO_SOURCES = ['nrq_io.cpp', 'nrq_fast5.cpp', 'nrq_inflate.cpp', os.path.join('..', '..', 'include', 'nrq_io.h'), os.path.join('..', '..', 'include', 'nrq.h')]


def build_io(force=False, verbose=False):
    """libnrqio.so: host-side C++ (deflate + chunk packing of the FAST5 write-back, front end, FAST5 files); g++, no CUDA"""
    if not force and os.path.exists(IO_LIB) and all(
            os.path.getmtime(os.path.join(HERE, s)) <= os.path.getmtime(IO_LIB) for s in IO_SOURCES):
        return IO_LIB
    cmd = [os.environ.get('CXX', 'g++'), '-O3', '-std=c++17', '-shared', '-fPIC', '-pthread', '-o', IO_LIB,
           os.path.join(HERE, 'nrq_io.cpp'), os.path.join(HERE, 'nrq_fast5.cpp'),
           os.path.join(HERE, 'nrq_inflate.cpp'), '-lz']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError('g++ failed building libnrqio.so')
    return IO_LIB


def build(force=False, verbose=False):
    build_io(force, verbose)
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + ['-o', LIB, os.path.join(HERE, 'nrq_api.cu')]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed building libnrq.so')
    with open(os.path.join(HERE, 'ptxas_info.txt'), 'w') as fp:
        fp.write(res.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
