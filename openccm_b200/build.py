"""Builds liboccm_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the source tree).

Objects are compiled in parallel and cached under csrc/_obj/ (git-ignored).  occm_solver.cu is compiled nine times: once
for the host API and controller kernels and once per Dormand-Prince / RHS stage (-DOCCM_STAGE_TU=k), because every stage
instantiates a dozen fused kernels and they dominate the compile time."""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
# kernel experiments: OCCM_BUILD_VARIANT=name OCCM_NVCC_EXTRA='-D...' builds build/variants/liboccm_b200_<name>.so (load it with OCCM_B200_LIB)
VARIANT = os.environ.get('OCCM_BUILD_VARIANT', '')
OBJ = os.path.join(CSRC, '_obj' + ('_' + VARIANT if VARIANT else ''))
LIB = os.path.join(CSRC, 'liboccm_b200.so') if not VARIANT else \
    os.path.join(os.path.dirname(HERE), 'build', 'variants', f'liboccm_b200_{VARIANT}.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC', '-diag-suppress', '177'] + \
    os.environ.get('OCCM_NVCC_EXTRA', '').split()          # experiments: e.g. OCCM_NVCC_EXTRA='-DOCCM_FAST_BLOCK=128 -DOCCM_FAST_CTAS=4'
N_STAGES = 8


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _headers():
    return glob.glob(os.path.join(CSRC, '*.h')) + glob.glob(os.path.join(CSRC, '*.cuh')) + \
        [os.path.join(os.path.dirname(HERE), 'include', 'occm_b200.h'), os.path.abspath(__file__)]


def _units():
    """(object path, source, extra flags) of every translation unit."""
    units = []
    for src in sources():
        stem = os.path.splitext(os.path.basename(src))[0]
        units.append((os.path.join(OBJ, stem + '.o'), src, []))
        if stem == 'occm_solver':
            units += [(os.path.join(OBJ, f'{stem}_stage{k}.o'), src, [f'-DOCCM_STAGE_TU={k}']) for k in range(N_STAGES)]
    return units


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get('NVCC', 'nvcc')
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    hdrs = _headers()
    units = _units()
    todo = [u for u in units if force or _newer(u[0], [u[1]] + hdrs)]

    def compile_one(u):
        obj, src, extra = u
        cmd = [nvcc, *NVCC_FLAGS, *extra, *(['-Xptxas', '-v'] if verbose else []), '-c', '-o', obj, src]
        print('[build]', ' '.join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)

    if todo:
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as pool:
            list(pool.map(compile_one, todo))
    objs = [u[0] for u in units]
    if todo or _newer(LIB, objs):
        cmd = [nvcc, '-arch=sm_100a', '-shared', '-o', LIB, *objs]
        print('[build]', ' '.join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    build_library(force='--force' in sys.argv, verbose='-v' in sys.argv)
