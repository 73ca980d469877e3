"""Build libb200fft.so in-tree with nvcc for sm_100a (no torch extension machinery:
the library is a plain C-ABI shared object loaded with ctypes).

    python -m mpi4py_fft_b200.build [--force]

Objects go to ``mpi4py_fft_b200/csrc/_build/`` and the library next to this
file; both are git-ignored but travel to the GPU box with the snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
BUILD = os.path.join(CSRC, '_build')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')
LIB = os.path.join(HERE, 'libb200fft.so')

NVCC_FLAGS = ['-std=c++17', '-O3', '-lineinfo',
              '-gencode', 'arch=compute_100a,code=sm_100a',
              '--compress-mode=size',        # ~9x smaller fatbin: the library travels to the GPU box with every snapshot
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
              '-I', INCLUDE, '-I', CSRC]


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return 'nvcc'


def _sources():
    """longest first (the mixed-radix and real-transform units take minutes, most others seconds):
    with a fixed number of workers the build then ends when the work does, not when a late big unit does"""
    def weight(name):
        for key, w in (('real_mixed', 9), ('pow2_mixed', 8), ('fft_real', 6), ('cpa_c', 4), ('cpa_b', 4), ('tma_b', 3),
                       ('pow2_large', 3), ('pow2_mid', 3), ('pow2_small', 3), ('rot', 2), ('chirpz', 2)):
            if key in name:
                return -w
        return 0
    return sorted((f for f in os.listdir(CSRC) if f.endswith('.cu')), key=lambda f: (weight(f), f))


def _digest():
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)) + ['../../include/b200fft.h']:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path) and name.endswith(('.cu', '.cuh', '.h')):
            h.update(name.encode())
            with open(path, 'rb') as f:
                h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src):
    obj = os.path.join(BUILD, src[:-3] + '.o')
    cmd = [_nvcc()] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ and link libb200fft.so; no-op when the
    sources have not changed since the last build."""
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, 'digest.txt')
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return LIB
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(_compile, srcs))
    cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ['-ldl', '-Xlinker', '--no-undefined']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, 'w') as f:
        f.write(digest)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True)
