"""nvcc build of the in-tree C-ABI library (importable without the library being present)."""
from __future__ import annotations

import os
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, 'libmvx_b200.so')
CSRC = os.path.join(_PKG, 'csrc')


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC']
OBJ_DIR = os.path.join(_ROOT, 'build', 'obj')     # git-ignored scratch: one object per source, rebuilt only when stale


def nvcc_command(out: str = LIB_PATH):
    """The equivalent one-shot command (what the incremental build below amounts to)."""
    return ['nvcc'] + NVCC_FLAGS + ['-shared', '-o', out] + sources()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into libmvx_b200.so (cross-compiles without a GPU).
    One object per .cu, compiled in parallel and only when the source or any header is newer; then one link."""
    from concurrent.futures import ThreadPoolExecutor
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith('.cu')] + [os.path.join(_ROOT, 'include', 'mvx_b200.h')]
    hdr_time = max(os.path.getmtime(h) for h in headers)
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= max([hdr_time] + [os.path.getmtime(x) for x in sources()]):
        return LIB_PATH     # up to date (the GPU box: the prebuilt library travels with the snapshot, the objects do not)
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            jobs.append(['nvcc'] + NVCC_FLAGS + ['-c', '-o', obj, src])

    def run(cmd):
        if verbose:
            print(' '.join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB_PATH) or any(os.path.getmtime(LIB_PATH) < os.path.getmtime(o) for o in objs):
        run(['nvcc'] + NVCC_FLAGS + ['-shared', '-o', LIB_PATH] + objs)
    return LIB_PATH


if __name__ == '__main__':
    build(force='-f' in sys.argv, verbose=True)
