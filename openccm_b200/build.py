"""Builds liboccm_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the source tree)."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(CSRC, 'liboccm_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '--threads', '0',
              '-shared', '-Xcompiler', '-fPIC', '--use_fast_math=false']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.h')) + glob.glob(os.path.join(CSRC, '*.cuh')) + \
        [os.path.join(os.path.dirname(HERE), 'include', 'occm_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    flags = [f for f in NVCC_FLAGS if not f.startswith('--use_fast_math')]
    cmd = [nvcc, *flags, *(['-Xptxas', '-v'] if verbose else []), '-o', LIB, *sources()]
    print('[build]', ' '.join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    build_library(force='--force' in sys.argv, verbose='-v' in sys.argv)
